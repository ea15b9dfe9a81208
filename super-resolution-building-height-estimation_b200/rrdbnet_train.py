"""Autograd for RRDBNet on the B200 kernels — the generator side of the SR fine-tune step.

Reference: `RealESRGAN.optimize_parameters` (SR/rrdbnet_arch.py:538-592) runs `self.output = self.net_g(self.lq)`,
sums pixel / perceptual / GAN losses, calls `backward()` and steps Adam on the generator, then updates the EMA copy
(:531-536).  The reference gets the backward pass from `nn.Conv2d` autograd; here `RRDBNet.forward / forward_feature`
under autograd dispatch to `RRDBNetTrainFn`, whose forward runs the network layer by layer through the same C-ABI
kernels as the frozen path (exact numerics) while keeping what the backward pass needs, and whose backward is built
from the same kernels:

  data gradient   dX = conv3x3(dY, W^T rotated 180 deg): `bhsr_conv_tc` with the fp32 NCHW accumulate epilogue, so the
                  five contributions a ResidualDenseBlock's concat channels receive (SR/rrdbnet_arch.py:137-143) add up
                  in one 192-channel gradient buffer; the gradient is pre-scaled by a power of two before the hi/lo split
                  (fp16 range) and un-scaled by the conv's per-channel `scale` vector;
  weight gradient dW = sum_pixels dY * X: `bhsr_head_wgrad_tc` (tcgen05, K = pixels; head_tc.cu) on groups of 64 (32-output
                  convs) or 32 (64-output convs) input channels (its TMEM budget), reading channel slices of the fp32 copy of the saved concat buffer;
  LeakyReLU       mask from the SAVED post-activation values (sign-preserving), nearest-x2 upsample backward = 2x2 sum.

Saved activations are the NHWC hi/lo planes the forward pass produces anyway: one 192-channel concat buffer per RDB
(2 x 192 x 2 B per pixel: 13.9 GB for 23 blocks at B = 64, 64x64 — sized for 180 GB of HBM, no recomputation).
This path is functional and parity-tested, not yet tuned: weight packing and fp32 copies happen per call.

Stand-alone `ResidualDenseBlock` / `RRDB` modules (SR/rrdbnet_arch.py:113-167) get the same treatment from
`RDBChainTrainFn` (one or three dense blocks, same kernels, same saved planes).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import ops
from ._lib import NUMERICS_EXACT

_SLOPE = 0.2


def _planes(nb: int, h: int, w: int, c: int, dev) -> Tuple[Tensor, Tensor]:
    return (torch.zeros((nb, h, w, c), dtype=torch.float16, device=dev),
            torch.zeros((nb, h, w, c), dtype=torch.float16, device=dev))


def _grad_scale(g: Tensor) -> Tensor:
    """Power-of-two device scalar that brings max|g| into [8, 16) (see hrfuse._grad_scale)."""
    amax = torch.linalg.vector_norm(g.detach(), ord=float("inf")).clamp_min(1e-30)
    return torch.exp2(torch.floor(4.0 - torch.log2(amax))).to(torch.float32).reshape(1)


def _lrelu_mask(post: Tensor) -> Tensor:
    """d lrelu(z) / dz from the stored lrelu(z): the activation keeps the sign."""
    return torch.where(post > 0, torch.ones((), dtype=post.dtype, device=post.device),
                       torch.full((), _SLOPE, dtype=post.dtype, device=post.device))


def _pad_oihw(w: Tensor, cout_pad: int, cin_pad: int) -> Tensor:
    full = torch.zeros((cout_pad, cin_pad, 3, 3), dtype=torch.float32, device=w.device)
    full[: w.shape[0], : w.shape[1]] = w
    return full


def conv_backward(x_f32: Tensor, x_choff: int, cin: int, g: Tensor, weight: Tensor, dx: Optional[Tensor],
                  dx_choff: int = 0, need_dw: bool = True):
    """Backward of y = conv3x3(x[:, x_choff:x_choff+cin], weight) + b given g = dL/dy (fp32 NCHW, contiguous).

    Returns (dW, db); accumulates the data gradient into dx[:, dx_choff:dx_choff+cin] (fp32 NCHW) when dx is given.
    """
    nb, cout, h, w = g.shape
    dev = g.device
    assert x_f32.is_contiguous() and g.is_contiguous() and g.dtype == torch.float32
    gs = _grad_scale(g)
    dw = db = None
    if need_dw:
        dw = torch.empty((cout, cin, 3, 3), dtype=torch.float32, device=dev)
        # TMEM budget of wgrad_tc_kernel: 3 * ceil(3 cin / 128) * 2 * nco <= 512 with nco = 16 / 32 / 64 padded outputs
        group = 32 if cout > 32 else 64
        for c0 in range(0, cin, group):
            cg = min(group, cin - c0)
            xf = ops.head_xform(x_f32, cg, h, w)
            xf.x_choff = x_choff + c0
            gf = ops.head_xform(g, cout, h, w, premul=gs)
            part, dbp = ops.head_wgrad_tc(x_f32, xf, gf, nb, cg, cout, 3, want_db=(c0 == 0))
            dw[:, c0:c0 + cg] = part
            if c0 == 0:
                db = dbp
    if dx is not None:
        assert dx.is_contiguous() and dx.dtype == torch.float32 and dx.shape[0] == nb and dx.shape[2:] == (h, w)
        cpad = (cout + 31) // 32 * 32            # plane channels (exact numerics: 32-channel chunks)
        ck = (cout + 15) // 16 * 16              # reduction length the kernel walks
        gp = _planes(nb, h, w, cpad, dev)
        ops.head_to_planes(g, ops.head_xform(g, cout, h, w, premul=gs), gp[0], gp[1], 0, cpad)
        wt = weight.detach().float().flip(2, 3).transpose(0, 1).contiguous()      # [cin, cout, 3, 3]
        unscale = (1.0 / gs).expand(64).contiguous()
        for o0 in range(0, cin, 64):
            co = min(64, cin - o0)
            wp = ops.pack_conv_weights(_pad_oihw(wt[o0:o0 + co], 64, ck), NUMERICS_EXACT)
            ops.conv_tc(gp[0], gp[1], 0, ck, wp, 64, None, ops.PLAIN_TAPS, None, None, out_choff=dx_choff + o0,
                        out_f32=dx, cout_valid=co, scale=unscale, accumulate=True, numerics=NUMERICS_EXACT)
    return dw, db


class RRDBNetTrainFn(torch.autograd.Function):
    """y = RRDBNet.forward(x) (feature=False) or forward_feature(x) (feature=True) with a backward pass.

    `convs`: [conv_first, (conv1..conv5) x 3 RDBs x num_block RRDBs, conv_body, conv_up1, conv_up2, conv_hr, conv_last]
    as nn.Conv2d parameter containers; their weights and biases are passed flat in `params` (same order, weight then
    bias) so autograd routes the gradients."""

    @staticmethod
    def forward(ctx, convs: Sequence[torch.nn.Conv2d], x: Tensor, feature: bool, *params: Tensor) -> Tensor:
        num = NUMERICS_EXACT
        n_rdb = (len(convs) - 6) // 5
        assert (len(convs) - 6) % 15 == 0 and len(params) == 2 * len(convs)
        x = x.detach().float()
        nb, _, h, w = x.shape
        dev = x.device
        P = [p.detach().to(dev, torch.float32).contiguous() for p in params]
        W = lambda i: P[2 * i]
        B = lambda i: P[2 * i + 1]
        with torch.cuda.device(dev):
            cats = [_planes(nb, h, w, 192, dev) for _ in range(n_rdb)]
            body_out = _planes(nb, h, w, 64, dev)
            ops.conv3x3_first(x, W(0), B(0), cats[0][0], cats[0][1], 0) if n_rdb else None
            if not n_rdb:                                  # num_block = 0: conv_first feeds conv_body directly
                ops.conv3x3_first(x, W(0), B(0), body_out[0], body_out[1], 0)
            for r in range(n_rdb):
                cur = cats[r]
                nxt = cats[r + 1] if r + 1 < n_rdb else body_out
                for k in range(5):
                    ci = 1 + 5 * r + k
                    cin = 64 + 32 * k
                    wp = ops.pack_conv_weights(W(ci), num)
                    if k < 4:
                        ops.conv_tc(cur[0], cur[1], 0, cin, wp, 32, B(ci), ops.PLAIN_TAPS, cur[0], cur[1],
                                    out_choff=cin, lrelu=True, numerics=num)
                    else:
                        kw = {}
                        if r % 3 == 2:                     # third RDB of an RRDB: out * 0.2 + RRDB input
                            rin = cats[r - 2]
                            kw = dict(res2=(rin[0], rin[1], 0), alpha2=_SLOPE)
                        ops.conv_tc(cur[0], cur[1], 0, cin, wp, 64, B(ci), ops.PLAIN_TAPS, nxt[0], nxt[1], out_choff=0,
                                    res1=(cur[0], cur[1], 0), alpha1=_SLOPE, numerics=num, **kw)
            i_body = 1 + 5 * n_rdb
            feat_src = cats[0] if n_rdb else body_out
            feat2 = _planes(nb, h, w, 64, dev)
            # feat + conv_body(body(feat)); with no blocks body(feat) = feat
            ops.conv_tc(body_out[0], body_out[1], 0, 64, ops.pack_conv_weights(W(i_body), num), 64, B(i_body),
                        ops.PLAIN_TAPS, feat2[0], feat2[1], out_choff=0, res1=(feat_src[0], feat_src[1], 0), alpha1=1.0,
                        numerics=num)
            up1 = _planes(nb, 2 * h, 2 * w, 64, dev)
            up2 = _planes(nb, 4 * h, 4 * w, 64, dev)
            for src, dst, ci in ((feat2, up1, i_body + 1), (up1, up2, i_body + 2)):
                for a in range(2):
                    for b in range(2):
                        wp = ops.pack_conv_weights(W(ci), num, fold_phase=2 * a + b)
                        ops.conv_tc(src[0], src[1], 0, 64, wp, 64, B(ci), ops.phase_taps(a, b), dst[0], dst[1],
                                    out_scale=2, out_oy=a, out_ox=b, lrelu=True, numerics=num)
            i_hr = i_body + 3
            hr = None
            if feature:
                y = torch.empty((nb, 64, 4 * h, 4 * w), dtype=torch.float32, device=dev)
                ops.conv_tc(up2[0], up2[1], 0, 64, ops.pack_conv_weights(W(i_hr), num), 64, B(i_hr), ops.PLAIN_TAPS,
                            None, None, out_f32=y, numerics=num)
            else:
                hr = _planes(nb, 4 * h, 4 * w, 64, dev)
                ops.conv_tc(up2[0], up2[1], 0, 64, ops.pack_conv_weights(W(i_hr), num), 64, B(i_hr), ops.PLAIN_TAPS,
                            hr[0], hr[1], out_choff=0, lrelu=True, numerics=num)
                # conv_last on the tensor-core kernel (outputs padded to 32; the fp32 NCHW epilogue stores the real ones)
                n_out = W(i_hr + 1).shape[0]
                if n_out > 32:
                    raise NotImplementedError("RRDBNet under autograd: at most 32 output channels")
                b_last = torch.zeros(32, dtype=torch.float32, device=dev)
                b_last[:n_out] = B(i_hr + 1)
                y = torch.empty((nb, n_out, 4 * h, 4 * w), dtype=torch.float32, device=dev)
                ops.conv_tc(hr[0], hr[1], 0, 64, ops.pack_conv_weights(_pad_oihw(W(i_hr + 1), 32, 64), num), 32, b_last,
                            ops.PLAIN_TAPS, None, None, out_f32=y, cout_valid=n_out, numerics=num)
        ctx.n_rdb, ctx.feature = n_rdb, feature
        ctx.x = x
        ctx.weights = [W(i) for i in range(len(convs))]
        ctx.acts = (cats, body_out, feat2, up1, up2, hr)
        ctx.need_dx = None
        return y

    @staticmethod
    def backward(ctx, gy: Tensor):
        n_rdb, feature = ctx.n_rdb, ctx.feature
        cats, body_out, feat2, up1, up2, hr = ctx.acts
        Wt = ctx.weights
        x = ctx.x
        nb, cin0, h, w = x.shape
        dev = x.device
        n_convs = len(Wt)
        grads: List[Optional[Tensor]] = [None] * (2 * n_convs)
        need = ctx.needs_input_grad          # (convs, x, feature, *params)
        gy = gy.contiguous().float()

        def put(i, dw, db):
            if need[3 + 2 * i]:
                grads[2 * i] = dw
            if need[3 + 2 * i + 1]:
                grads[2 * i + 1] = db

        def f32(pl, c):
            return ops.planes_to_nchw(pl[0], pl[1], c, 0)

        with torch.cuda.device(dev):
            i_body = 1 + 5 * n_rdb
            i_hr = i_body + 3
            # ---- conv_last / conv_hr (SR/rrdbnet_arch.py:221-222, 238)
            up2_f = f32(up2, 64)
            if feature:
                g_hr = gy
            else:
                hr_f = f32(hr, 64)
                d_hr = torch.zeros_like(hr_f)
                dw, db = conv_backward(hr_f, 0, 64, gy, Wt[i_hr + 1], d_hr)
                put(i_hr + 1, dw, db)
                g_hr = d_hr * _lrelu_mask(hr_f)
                del hr_f, d_hr
            d_up2 = torch.zeros_like(up2_f)
            dw, db = conv_backward(up2_f, 0, 64, g_hr, Wt[i_hr], d_up2)
            put(i_hr, dw, db)
            g = d_up2 * _lrelu_mask(up2_f)
            del d_up2, up2_f, g_hr
            # ---- the two nearest-x2 + conv + lrelu stages (:236-237)
            for src, ci in ((up1, i_body + 2), (feat2, i_body + 1)):
                src_f = f32(src, 64)
                xin = src_f.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3).contiguous()
                d_xin = torch.zeros_like(xin)
                dw, db = conv_backward(xin, 0, 64, g.contiguous(), Wt[ci], d_xin)
                put(ci, dw, db)
                sh = src_f.shape
                d_src = d_xin.view(sh[0], sh[1], sh[2], 2, sh[3], 2).sum(dim=(3, 5))
                g = d_src * _lrelu_mask(src_f) if src is up1 else d_src      # feat2 is not activated
                del xin, d_xin, src_f
            d_feat2 = g.contiguous()                       # = gradient of feat (skip) and of conv_body's output
            # ---- conv_body (:234-235)
            body_f = f32(body_out, 64)
            d_body = torch.zeros_like(body_f)
            dw, db = conv_backward(body_f, 0, 64, d_feat2, Wt[i_body], d_body)
            put(i_body, dw, db)
            del body_f
            # ---- RRDB trunk in reverse (:137-143, 160-167)
            d = d_body
            for r in range(n_rdb - 1, -1, -1):
                if r % 3 == 2:
                    d_rrdb = d                              # gradient of the RRDB output
                    d = d * _SLOPE                          # into rdb3's output
                cat_f = f32(cats[r], 192)
                d_cat = torch.zeros_like(cat_f)
                d_cat[:, :64] += d                          # x5 * 0.2 + x: identity branch
                gk = (d * _SLOPE).contiguous()
                for k in range(4, -1, -1):
                    ci = 1 + 5 * r + k
                    cin = 64 + 32 * k
                    if k < 4:
                        sl = slice(cin, cin + 32)
                        gk = (d_cat[:, sl] * _lrelu_mask(cat_f[:, sl])).contiguous()
                    dw, db = conv_backward(cat_f, 0, cin, gk, Wt[ci], d_cat)
                    put(ci, dw, db)
                d = d_cat[:, :64].contiguous()
                if r % 3 == 0:
                    d = d + d_rrdb                          # RRDB: out * 0.2 + x
                del cat_f, d_cat
            d_feat = d + d_feat2 if n_rdb else d_body + d_feat2
            # ---- conv_first (:232)
            xc = x.contiguous()
            dx = torch.zeros_like(xc) if need[1] else None
            dw, db = conv_backward(xc, 0, cin0, d_feat.contiguous(), Wt[0], dx)
            put(0, dw, db)
        return (None, dx, None, *grads)


class RDBChainTrainFn(torch.autograd.Function):
    """A stand-alone `ResidualDenseBlock` (rrdb=False, five convs; SR/rrdbnet_arch.py:137-143) or `RRDB` (rrdb=True,
    fifteen convs, `out * 0.2 + x` around three dense blocks; :160-167) with a backward pass — the trunk section of
    `RRDBNetTrainFn` on its own.  `convs`: conv1..conv5 of every dense block in order; `params`: their weights and
    biases flat (weight then bias)."""

    @staticmethod
    def forward(ctx, convs: Sequence[torch.nn.Conv2d], x: Tensor, rrdb: bool, *params: Tensor) -> Tensor:
        num = NUMERICS_EXACT
        n_rdb = len(convs) // 5
        assert len(convs) == (15 if rrdb else 5) and len(params) == 2 * len(convs)
        x = x.detach().float().contiguous()
        nb, _, h, w = x.shape
        dev = x.device
        P = [p.detach().to(dev, torch.float32).contiguous() for p in params]
        with torch.cuda.device(dev):
            cats = [_planes(nb, h, w, 192, dev) for _ in range(n_rdb)]
            out = _planes(nb, h, w, 64, dev)
            ops.nchw_to_planes(x, cats[0][0], cats[0][1], 0)
            for r in range(n_rdb):
                cur = cats[r]
                nxt = cats[r + 1] if r + 1 < n_rdb else out
                for k in range(5):
                    ci = 5 * r + k
                    cin = 64 + 32 * k
                    wp = ops.pack_conv_weights(P[2 * ci], num)
                    if k < 4:
                        ops.conv_tc(cur[0], cur[1], 0, cin, wp, 32, P[2 * ci + 1], ops.PLAIN_TAPS, cur[0], cur[1],
                                    out_choff=cin, lrelu=True, numerics=num)
                    else:
                        kw = {}
                        if rrdb and r == 2:                # third dense block of an RRDB: out * 0.2 + RRDB input
                            kw = dict(res2=(cats[0][0], cats[0][1], 0), alpha2=_SLOPE)
                        ops.conv_tc(cur[0], cur[1], 0, cin, wp, 64, P[2 * ci + 1], ops.PLAIN_TAPS, nxt[0], nxt[1],
                                    out_choff=0, res1=(cur[0], cur[1], 0), alpha1=_SLOPE, numerics=num, **kw)
            y = ops.planes_to_nchw(out[0], out[1], 64, 0)
        ctx.rrdb = rrdb
        ctx.weights = [P[2 * i] for i in range(len(convs))]
        ctx.cats = cats
        return y

    @staticmethod
    def backward(ctx, gy: Tensor):
        cats, Wt, rrdb = ctx.cats, ctx.weights, ctx.rrdb
        n_rdb = len(cats)
        need = ctx.needs_input_grad          # (convs, x, rrdb, *params)
        grads: List[Optional[Tensor]] = [None] * (2 * len(Wt))
        d = gy.contiguous().float()
        with torch.cuda.device(d.device):
            d_out = d                                       # gradient of the block's output
            if rrdb:
                d = d * _SLOPE                              # into the third dense block's output
            for r in range(n_rdb - 1, -1, -1):
                cat_f = ops.planes_to_nchw(cats[r][0], cats[r][1], 192, 0)
                d_cat = torch.zeros_like(cat_f)
                d_cat[:, :64] += d                          # x5 * 0.2 + x: identity branch
                gk = (d * _SLOPE).contiguous()
                for k in range(4, -1, -1):
                    ci = 5 * r + k
                    cin = 64 + 32 * k
                    if k < 4:
                        sl = slice(cin, cin + 32)
                        gk = (d_cat[:, sl] * _lrelu_mask(cat_f[:, sl])).contiguous()
                    need_dw = need[3 + 2 * ci] or need[3 + 2 * ci + 1]
                    dw, db = conv_backward(cat_f, 0, cin, gk, Wt[ci], d_cat, need_dw=need_dw)
                    if need[3 + 2 * ci]:
                        grads[2 * ci] = dw
                    if need[3 + 2 * ci + 1]:
                        grads[2 * ci + 1] = db
                d = d_cat[:, :64].contiguous()
                del cat_f, d_cat
            if rrdb:
                d = d + d_out                               # RRDB: out * 0.2 + x
        return (None, d if need[1] else None, None, *grads)
