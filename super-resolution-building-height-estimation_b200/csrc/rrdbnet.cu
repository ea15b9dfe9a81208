// rrdbnet.cu — the whole RRDBNet x4 forward as one C call: a fixed schedule of
// bhsr_conv_tc launches over a caller-provided workspace.
//
// Reference: RRDBNet.forward / forward_feature (SR/rrdbnet_arch.py:208-240), RRDB (:162-167),
// ResidualDenseBlock (:136-143).  What the schedule removes relative to the reference:
//   * torch.cat: every RDB conv writes its 32-channel result straight into its slice of a
//     192-channel NHWC concat buffer, the next conv reads channels [0, cin) of the same buffer;
//   * the `x5*0.2 + x` / `out*0.2 + x` / `feat + body_feat` elementwise kernels: folded into the
//     producing conv's epilogue (RES1 / RES2);
//   * F.interpolate(nearest x2): conv3x3(nearest_x2(x)) is computed as its four 2x2-tap
//     sub-pixel phases on the SOURCE grid (2.25x fewer MACs, no upsampled tensor in memory).
//
// Workspace (all NHWC fp16 hi/lo plane pairs):
//   buf[3]  [nb][h][w][192]   concat buffers, rotated rdb1 -> rdb2 -> rdb3
//   feat    [nb][h][w][64]    conv_first output (the long skip)
//   trunk   [nb][h][w][64]    feat + conv_body(body(feat))
//   up1     [nb][2h][2w][64]
//   up2     [nb][4h][4w][64]
//   hr      [nb][4h][4w][64]  only for forward(): lrelu(conv_hr) feeding conv_last, + conv_last's packed weights / bias
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <string.h>

#include "common.h"

namespace bhsr {

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Planes {
  __half* hi;
  __half* lo;
};

struct Workspace {
  Planes buf[3], feat, trunk, up1, up2, hr;
  void* last_w;      // forward(): conv_last packed for the tensor-core kernel (3 outputs padded to 32)
  float* last_b;     // ... and its bias padded to 32 floats
  size_t total;
};

static Workspace carve(void* base, int nb, int h, int w, bool feature_only) {
  Workspace ws{};
  size_t off = 0;
  auto take = [&](size_t pixels, int ch) {
    Planes p;
    const size_t bytes = align_up(pixels * ch * sizeof(__half), 1024);
    p.hi = reinterpret_cast<__half*>(static_cast<char*>(base) + off);
    off += bytes;
    p.lo = reinterpret_cast<__half*>(static_cast<char*>(base) + off);
    off += bytes;
    return p;
  };
  const size_t px = static_cast<size_t>(nb) * h * w;
  for (int i = 0; i < 3; ++i) ws.buf[i] = take(px, 192);
  ws.feat = take(px, 64);
  ws.trunk = take(px, 64);
  ws.up1 = take(px * 4, 64);
  ws.up2 = take(px * 16, 64);
  if (!feature_only) {
    ws.hr = take(px * 16, 64);
    ws.last_w = static_cast<char*>(base) + off;
    off += align_up(bhsr_packed_conv_weight_bytes(32, 64, 9, BHSR_NUMERICS_EXACT_F16X3), 1024);
    ws.last_b = reinterpret_cast<float*>(static_cast<char*>(base) + off);
    off += 1024;
  }
  ws.total = off;
  return ws;
}

// the tensor-core convs in schedule order: 15 per block, conv_body, conv_up1, conv_up2, conv_hr
static int tc_conv_count(int num_block) { return 15 * num_block + 4; }
static int rdb_cin(int c) { return 64 + 32 * c; }          // c = 0..4
static int rdb_cout(int c) { return c < 4 ? 32 : 64; }

struct PackedLayout {
  // byte offset of conv i's packed blob (conv_up1/up2 hold 4 phase blobs back to back)
  size_t conv_off(int num_block, int numerics, int idx) const {
    size_t off = 0;
    const int n = tc_conv_count(num_block);
    for (int i = 0; i < idx && i < n; ++i) off += conv_bytes(num_block, numerics, i);
    return off;
  }
  static size_t conv_bytes(int num_block, int numerics, int i) {
    const int body = 15 * num_block;
    if (i < body) {
      const int c = i % 5;
      return bhsr_packed_conv_weight_bytes(rdb_cout(c), rdb_cin(c), 9, numerics);
    }
    const int k = i - body;  // 0 body, 1 up1, 2 up2, 3 hr
    if (k == 1 || k == 2) return 4 * bhsr_packed_conv_weight_bytes(64, 64, 4, numerics);
    return bhsr_packed_conv_weight_bytes(64, 64, 9, numerics);
  }
};

static size_t bias_off(int num_block, int idx) {
  size_t off = 0;
  const int body = 15 * num_block;
  for (int i = 0; i < idx; ++i) off += (i < body) ? rdb_cout(i % 5) : 64;
  return off;
}

static void plain_taps(BhsrConvTcDesc& d) {
  d.ntaps = 9;
  for (int t = 0; t < 9; ++t) {
    d.dy[t] = static_cast<int8_t>(t / 3 - 1);
    d.dx[t] = static_cast<int8_t>(t % 3 - 1);
  }
}
static void phase_taps(BhsrConvTcDesc& d, int a, int b) {
  d.ntaps = 4;
  for (int t = 0; t < 4; ++t) {
    d.dy[t] = static_cast<int8_t>(a - 1 + (t >> 1));
    d.dx[t] = static_cast<int8_t>(b - 1 + (t & 1));
  }
}

}  // namespace bhsr

using namespace bhsr;

extern "C" size_t bhsr_rrdbnet_packed_bytes(int32_t num_block, int32_t numerics) {
  return PackedLayout().conv_off(num_block, numerics, tc_conv_count(num_block));
}

extern "C" size_t bhsr_rrdbnet_bias_floats(int32_t num_block) {
  return bias_off(num_block, tc_conv_count(num_block));
}

extern "C" size_t bhsr_rrdbnet_workspace_bytes(int32_t nb, int32_t h, int32_t w,
                                               int32_t feature_only) {
  return carve(nullptr, nb, h, w, feature_only != 0).total;
}

extern "C" int bhsr_rrdbnet_pack(const float* const* params, int32_t num_block, int32_t numerics,
                                 void* packed, float* biases, void* stream) {
  BHSR_REQUIRE(params && packed && biases && num_block >= 0, "rrdbnet_pack: bad arguments");
  const int n = tc_conv_count(num_block);
  const int body = 15 * num_block;
  size_t off = 0, boff = 0;
  char* base = static_cast<char*>(packed);
  for (int i = 0; i < n; ++i) {
    const float* w = params[2 * i];
    const float* b = params[2 * i + 1];
    BHSR_REQUIRE(w && b, "rrdbnet_pack: null parameter %d", i);
    int cout = 64, cin = 64;
    if (i < body) { cout = rdb_cout(i % 5); cin = rdb_cin(i % 5); }
    const int k = i - body;
    if (i >= body && (k == 1 || k == 2)) {
      const size_t one = bhsr_packed_conv_weight_bytes(64, 64, 4, numerics);
      for (int ph = 0; ph < 4; ++ph) {
        int rc = bhsr_pack_conv_weights(w, 64, 64, ph, numerics, base + off + ph * one, stream);
        if (rc) return rc;
      }
    } else {
      int rc = bhsr_pack_conv_weights(w, cout, cin, -1, numerics, base + off, stream);
      if (rc) return rc;
    }
    BHSR_CUDA_CHECK(cudaMemcpyAsync(biases + boff, b, cout * sizeof(float),
                                    cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
    off += PackedLayout::conv_bytes(num_block, numerics, i);
    boff += cout;
  }
  return 0;
}

extern "C" int bhsr_rrdbnet_forward(const BhsrRrdbNetDesc* dp, const float* x, int64_t sn,
                                    int64_t sc, int64_t sh, int64_t sw, float* y, int32_t feature,
                                    void* stream) {
  BHSR_REQUIRE(dp && x && y, "rrdbnet_forward: null pointer");
  const BhsrRrdbNetDesc& d = *dp;
  BHSR_REQUIRE(d.nb > 0 && d.h > 0 && d.w > 0 && d.num_block >= 0 && d.num_in_ch > 0,
               "rrdbnet_forward: bad shape");
  BHSR_REQUIRE(d.conv_first_w && d.conv_first_b && d.packed && d.biases && d.workspace,
               "rrdbnet_forward: null parameter/workspace");
  BHSR_REQUIRE(feature || (d.conv_last_w && d.conv_last_b && d.num_out_ch >= 1 && d.num_out_ch <= 8),
               "rrdbnet_forward: forward() needs conv_last with 1..8 outputs");
  BHSR_REQUIRE((reinterpret_cast<uintptr_t>(d.workspace) & 1023) == 0,
               "rrdbnet_forward: workspace must be 1024-byte aligned");
  const size_t need = bhsr_rrdbnet_workspace_bytes(d.nb, d.h, d.w, feature);
  BHSR_REQUIRE(d.workspace_bytes >= need, "rrdbnet_forward: workspace too small (%zu < %zu)",
               d.workspace_bytes, need);
  Workspace ws = carve(d.workspace, d.nb, d.h, d.w, feature != 0);
  const bool exact = d.numerics == BHSR_NUMERICS_EXACT_F16X3;
  const char* packed = static_cast<const char*>(d.packed);
  size_t poff = 0, boff = 0;
  int rc;

  // conv_first -> the long-skip copy and the first RDB's x0, one launch (K = 27: CUDA cores)
  rc = bhsr_conv3x3_first(x, sn, sc, sh, sw, d.nb, d.num_in_ch, d.h, d.w, d.conv_first_w,
                          d.conv_first_b, 64, ws.feat.hi, ws.feat.lo, 64, 0, ws.buf[0].hi,
                          ws.buf[0].lo, 192, 0, stream);
  if (rc) return rc;

  auto base_desc = [&](const Planes& in, int in_ctot, int cin, int cout, int h, int w) {
    BhsrConvTcDesc c;
    memset(&c, 0, sizeof(c));
    c.in_hi = in.hi; c.in_lo = in.lo;
    c.nb = d.nb; c.h = h; c.w = w;
    c.in_ctot = in_ctot; c.in_choff = 0; c.cin = cin;
    c.cout = cout;
    c.w_packed = packed + poff;
    c.bias = d.biases + boff;
    c.oh = h; c.ow = w; c.out_scale = 1;
    c.numerics = d.numerics;
    c.mblocks = d.mblocks;
    c.alpha1 = 1.f; c.alpha2 = 1.f;
    (void)exact;
    return c;
  };
  const int body = 15 * d.num_block;
  int idx = 0;
  auto advance = [&](int cout) {
    poff += PackedLayout::conv_bytes(d.num_block, d.numerics, idx);
    boff += cout;
    ++idx;
  };

  for (int blk = 0; blk < d.num_block; ++blk) {
    for (int r = 0; r < 3; ++r) {
      const Planes& cur = ws.buf[r];
      const Planes& nxt = ws.buf[(r + 1) % 3];
      for (int c = 0; c < 5; ++c) {
        BhsrConvTcDesc cd = base_desc(cur, 192, rdb_cin(c), rdb_cout(c), d.h, d.w);
        plain_taps(cd);
        if (c < 4) {
          cd.out_hi = cur.hi; cd.out_lo = cur.lo; cd.out_ctot = 192; cd.out_choff = rdb_cin(c);
          cd.epilogue = BHSR_EPI_LRELU;
        } else {
          cd.out_hi = nxt.hi; cd.out_lo = nxt.lo; cd.out_ctot = 192; cd.out_choff = 0;
          cd.epilogue = BHSR_EPI_RES1;
          cd.alpha1 = 0.2f;
          cd.res1_hi = cur.hi; cd.res1_lo = cur.lo; cd.res1_ctot = 192; cd.res1_choff = 0;
          if (r == 2) {  // RRDB tail: (rdb3_out) * 0.2 + rrdb_in, written over rrdb_in in place
            cd.epilogue |= BHSR_EPI_RES2;
            cd.alpha2 = 0.2f;
            cd.res2_hi = ws.buf[0].hi; cd.res2_lo = ws.buf[0].lo; cd.res2_ctot = 192;
            cd.res2_choff = 0;
          }
        }
        rc = bhsr_conv_tc(&cd, stream);
        if (rc) return rc;
        advance(rdb_cout(c));
      }
    }
  }
  // conv_body + long skip
  {
    BhsrConvTcDesc cd = base_desc(ws.buf[0], 192, 64, 64, d.h, d.w);
    plain_taps(cd);
    cd.out_hi = ws.trunk.hi; cd.out_lo = ws.trunk.lo; cd.out_ctot = 64; cd.out_choff = 0;
    cd.epilogue = BHSR_EPI_RES1; cd.alpha1 = 1.f;
    cd.res1_hi = ws.feat.hi; cd.res1_lo = ws.feat.lo; cd.res1_ctot = 64; cd.res1_choff = 0;
    rc = bhsr_conv_tc(&cd, stream);
    if (rc) return rc;
    advance(64);
  }
  // conv_up1 / conv_up2 as four sub-pixel phases each
  const Planes* src[2] = {&ws.trunk, &ws.up1};
  const Planes* dst[2] = {&ws.up1, &ws.up2};
  for (int u = 0; u < 2; ++u) {
    const int h = d.h << u, w = d.w << u;
    const size_t one = bhsr_packed_conv_weight_bytes(64, 64, 4, d.numerics);
    for (int ph = 0; ph < 4; ++ph) {
      BhsrConvTcDesc cd = base_desc(*src[u], 64, 64, 64, h, w);
      cd.w_packed = packed + poff + ph * one;
      phase_taps(cd, ph >> 1, ph & 1);
      cd.oh = 2 * h; cd.ow = 2 * w; cd.out_scale = 2; cd.out_oy = ph >> 1; cd.out_ox = ph & 1;
      cd.out_hi = dst[u]->hi; cd.out_lo = dst[u]->lo; cd.out_ctot = 64; cd.out_choff = 0;
      cd.epilogue = BHSR_EPI_LRELU;
      rc = bhsr_conv_tc(&cd, stream);
      if (rc) return rc;
    }
    advance(64);
  }
  // conv_hr
  {
    const int h = d.h * 4, w = d.w * 4;
    BhsrConvTcDesc cd = base_desc(ws.up2, 64, 64, 64, h, w);
    plain_taps(cd);
    if (feature) {
      cd.out_f32 = y; cd.out_ctot = 64; cd.out_choff = 0;
      cd.epilogue = BHSR_EPI_OUT_NCHW_F32;
    } else {
      cd.out_hi = ws.hr.hi; cd.out_lo = ws.hr.lo; cd.out_ctot = 64; cd.out_choff = 0;
      cd.epilogue = BHSR_EPI_LRELU;        // conv_last reads lrelu(conv_hr) (SR/rrdbnet_arch.py:222)
    }
    rc = bhsr_conv_tc(&cd, stream);
    if (rc) return rc;
    advance(64);
    if (!feature) {
      // conv_last (64 -> num_out_ch) on the tensor-core kernel: outputs padded to 32, fp32 NCHW epilogue writes the real
      // ones.  (Round 1 ran it as a thread-per-pixel CUDA-core kernel: 1.9 ms at B = 12, 256x256 — 60x its HBM time.)
      cudaStream_t st = static_cast<cudaStream_t>(stream);
      rc = pack_conv_weights_padded(d.conv_last_w, d.num_out_ch, 32, 64, d.numerics, ws.last_w, stream);
      if (rc) return rc;
      BHSR_CUDA_CHECK(cudaMemsetAsync(ws.last_b, 0, 32 * sizeof(float), st));
      BHSR_CUDA_CHECK(cudaMemcpyAsync(ws.last_b, d.conv_last_b, d.num_out_ch * sizeof(float), cudaMemcpyDeviceToDevice, st));
      BhsrConvTcDesc cl = base_desc(ws.hr, 64, 64, 32, h, w);
      plain_taps(cl);
      cl.w_packed = ws.last_w;
      cl.bias = ws.last_b;
      cl.out_f32 = y; cl.out_ctot = d.num_out_ch; cl.out_choff = 0;
      cl.cout_valid = d.num_out_ch;
      cl.epilogue = BHSR_EPI_OUT_NCHW_F32;
      rc = bhsr_conv_tc(&cl, stream);
      if (rc) return rc;
    }
  }
  (void)body;
  return 0;
}
