"""Python-side wrappers of the C-ABI kernels: argument marshalling only, no arithmetic.

Activation "planes" are NHWC fp16 tensors [NB, H, W, C]; `hi` carries fp16(v) and `lo` carries
fp16((v - hi) * 2**11), so hi + lo * 2**-11 reproduces the fp32 value to ~22 bits.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import (EPI_ACCUM, EPI_LRELU, EPI_OUT_NCHW_F32, EPI_RELU, EPI_RES1, EPI_RES2, EPI_SHUFFLE2, NUMERICS,
                   NUMERICS_EXACT, NUMERICS_FAST, ConvTcDesc)

PLAIN_TAPS: Tuple[Tuple[int, int], ...] = tuple((ky - 1, kx - 1) for ky in range(3) for kx in range(3))


def phase_taps(a: int, b: int) -> Tuple[Tuple[int, int], ...]:
    """(dy, dx) of the 2x2 taps of conv3x3(nearest_x2(x)) at output parity (a, b).

    F.interpolate(scale_factor=2, mode='nearest') maps destination index d to source d // 2
    (SR/rrdbnet_arch.py:236-237), so the three kernel rows of an output row 2y+a read source
    rows {y-1, y, y} (a = 0) or {y, y, y+1} (a = 1); same for columns.
    """
    return tuple((a - 1 + iy, b - 1 + ix) for iy in range(2) for ix in range(2))


def numerics_code(mode) -> int:
    if isinstance(mode, str):
        return NUMERICS[mode]
    return int(mode)


def packed_conv_weight_elems(cout: int, cin: int, ntaps: int, numerics: int) -> int:
    return _lib.load().bhsr_packed_conv_weight_bytes(cout, cin, ntaps, numerics) // 2


@_lib.device_guarded
def pack_conv_weights(w: torch.Tensor, numerics: int, fold_phase: int = -1,
                      out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """OIHW fp32 [cout, cin, 3, 3] -> packed fp16 blob for conv_tc (see bhsr.h)."""
    _lib.require_cuda(w, "weight")
    assert w.dtype == torch.float32 and w.dim() == 4 and w.shape[2:] == (3, 3)
    w = w.contiguous()
    cout, cin = w.shape[0], w.shape[1]
    ntaps = 9 if fold_phase < 0 else 4
    n = packed_conv_weight_elems(cout, cin, ntaps, numerics)
    if out is None:
        out = torch.empty(n, dtype=torch.float16, device=w.device)
    assert out.numel() == n and out.dtype == torch.float16 and out.is_contiguous()
    _lib.check(_lib.load().bhsr_pack_conv_weights(w.data_ptr(), cout, cin, fold_phase, numerics,
                                                  out.data_ptr(), _lib.stream_ptr(w.device)),
               "bhsr_pack_conv_weights")
    return out


@_lib.device_guarded
def conv_tc(in_hi: torch.Tensor, in_lo: Optional[torch.Tensor], in_choff: int, cin: int,
            w_packed: torch.Tensor, cout: int, bias: Optional[torch.Tensor],
            taps: Sequence[Tuple[int, int]],
            out_hi: Optional[torch.Tensor], out_lo: Optional[torch.Tensor], out_choff: int = 0,
            out_f32: Optional[torch.Tensor] = None,
            out_scale: int = 1, out_oy: int = 0, out_ox: int = 0,
            lrelu: bool = False,
            res1: Optional[Tuple[torch.Tensor, Optional[torch.Tensor], int]] = None, alpha1: float = 1.0,
            res2: Optional[Tuple[torch.Tensor, Optional[torch.Tensor], int]] = None, alpha2: float = 1.0,
            numerics: int = NUMERICS_EXACT, mblocks: int = 0, max_ctas: int = 0,
            desc_mode: int = 0, scale: Optional[torch.Tensor] = None, cout_valid: int = 0,
            relu: bool = False, shuffle2: bool = False, accumulate: bool = False) -> None:
    """Enqueue one tensor-core convolution on the current stream (bhsr_conv_tc)."""
    _lib.require_cuda(in_hi, "in_hi")
    nb, h, w, ctot = in_hi.shape
    d = ConvTcDesc()
    d.in_hi = in_hi.data_ptr()
    d.in_lo = _lib.ptr(in_lo)
    d.nb, d.h, d.w = nb, h, w
    d.in_ctot, d.in_choff, d.cin = ctot, in_choff, cin
    d.w_packed = w_packed.data_ptr()
    d.cout = cout
    d.bias = _lib.ptr(bias)
    d.scale = _lib.ptr(scale)
    d.cout_valid = cout_valid
    d.ntaps = len(taps)
    for i, (dy, dx) in enumerate(taps):
        d.dy[i] = dy
        d.dx[i] = dx
    epi = 0
    if out_f32 is not None:
        epi |= EPI_OUT_NCHW_F32
        assert out_f32.dtype == torch.float32 and out_f32.is_contiguous()
        _, octot, oh, ow = out_f32.shape
        d.out_f32 = out_f32.data_ptr()
    else:
        assert out_hi is not None and out_hi.dtype == torch.float16 and out_hi.is_contiguous()
        _, oh, ow, octot = out_hi.shape
        d.out_hi = out_hi.data_ptr()
        d.out_lo = _lib.ptr(out_lo)
    d.oh, d.ow, d.out_scale, d.out_oy, d.out_ox = oh, ow, out_scale, out_oy, out_ox
    d.out_ctot, d.out_choff = octot, out_choff
    if lrelu:
        epi |= EPI_LRELU
    if relu:
        epi |= EPI_RELU
    if shuffle2:
        epi |= EPI_SHUFFLE2
    if accumulate:
        epi |= EPI_ACCUM
    if res1 is not None:
        epi |= EPI_RES1
        d.res1_hi, d.res1_lo = res1[0].data_ptr(), _lib.ptr(res1[1])
        d.res1_ctot, d.res1_choff = res1[0].shape[3], res1[2]
    if res2 is not None:
        epi |= EPI_RES2
        d.res2_hi, d.res2_lo = res2[0].data_ptr(), _lib.ptr(res2[1])
        d.res2_ctot, d.res2_choff = res2[0].shape[3], res2[2]
    d.epilogue = epi
    d.alpha1, d.alpha2 = alpha1, alpha2
    d.numerics, d.mblocks, d.max_ctas, d.desc_mode = numerics, mblocks, max_ctas, desc_mode
    _lib.check(_lib.load().bhsr_conv_tc(d, _lib.stream_ptr(in_hi.device)), "bhsr_conv_tc")


@_lib.device_guarded
def nchw_to_planes(x: torch.Tensor, out_hi: torch.Tensor, out_lo: Optional[torch.Tensor],
                   choff: int = 0) -> None:
    _lib.require_cuda(x, "x")
    assert x.dtype == torch.float32 and x.is_contiguous()
    nb, c, h, w = x.shape
    assert out_hi.shape[:3] == (nb, h, w)
    _lib.check(_lib.load().bhsr_nchw_f32_to_planes(x.data_ptr(), nb, c, h, w, out_hi.data_ptr(),
                                                   _lib.ptr(out_lo), out_hi.shape[3], choff,
                                                   _lib.stream_ptr(x.device)),
               "bhsr_nchw_f32_to_planes")


@_lib.device_guarded
def planes_to_nchw(in_hi: torch.Tensor, in_lo: Optional[torch.Tensor], c: int, choff: int = 0,
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _lib.require_cuda(in_hi, "in_hi")
    nb, h, w, ctot = in_hi.shape
    if out is None:
        out = torch.empty((nb, c, h, w), dtype=torch.float32, device=in_hi.device)
    _lib.check(_lib.load().bhsr_planes_to_nchw_f32(in_hi.data_ptr(), _lib.ptr(in_lo), nb, c, h, w,
                                                   ctot, choff, out.data_ptr(),
                                                   _lib.stream_ptr(in_hi.device)),
               "bhsr_planes_to_nchw_f32")
    return out


@_lib.device_guarded
def conv3x3_first(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor],
                  out_hi: torch.Tensor, out_lo: Optional[torch.Tensor], out_choff: int = 0) -> None:
    """fp32 NCHW (any strides) -> planes; SR/rrdbnet_arch.py:232."""
    _lib.require_cuda(x, "x")
    assert x.dtype == torch.float32
    nb, cin, h, w = x.shape
    sn, sc, sh, sw = x.stride()
    cout = weight.shape[0]
    _lib.check(_lib.load().bhsr_conv3x3_first(x.data_ptr(), sn, sc, sh, sw, nb, cin, h, w,
                                              weight.data_ptr(), _lib.ptr(bias), cout,
                                              out_hi.data_ptr(), _lib.ptr(out_lo), out_hi.shape[3],
                                              out_choff, None, None, 0, 0, _lib.stream_ptr(x.device)),
               "bhsr_conv3x3_first")


@_lib.device_guarded
def conv3x3_last(in_hi: torch.Tensor, in_lo: Optional[torch.Tensor], in_choff: int, cin: int,
                 weight: torch.Tensor, bias: Optional[torch.Tensor], lrelu_in: bool,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """planes -> fp32 NCHW with a small output channel count; SR/rrdbnet_arch.py:221-222."""
    _lib.require_cuda(in_hi, "in_hi")
    nb, h, w, ctot = in_hi.shape
    cout = weight.shape[0]
    if out is None:
        out = torch.empty((nb, cout, h, w), dtype=torch.float32, device=in_hi.device)
    _lib.check(_lib.load().bhsr_conv3x3_last(in_hi.data_ptr(), _lib.ptr(in_lo), ctot, in_choff, nb,
                                             cin, h, w, int(lrelu_in), weight.data_ptr(),
                                             _lib.ptr(bias), cout, out.data_ptr(),
                                             _lib.stream_ptr(in_hi.device)),
               "bhsr_conv3x3_last")
    return out


# ------------------------------------------------------------------ tensor-core training head (head_tc.cu)
def head_xform(x: torch.Tensor, c: int, h: int, w: int, *, unshuffle: bool = False, in_affine=None,
               in_relu: bool = False, premul: Optional[torch.Tensor] = None) -> "_lib.HeadXform":
    """How a conv reads the contiguous fp32 NCHW tensor `x` (BhsrHeadXform).  The returned struct holds raw
    pointers: keep `x`, the affine vectors and `premul` alive until the call that uses it has been enqueued."""
    assert x.dtype == torch.float32 and x.is_contiguous()
    t = _lib.HeadXform()
    t.x, t.x_ctot, t.x_choff = x.data_ptr(), x.shape[1], 0
    t.c, t.h, t.w = c, h, w
    t.unshuffle = int(unshuffle)
    if in_affine is not None:
        t.in_scale, t.in_shift = in_affine[0].data_ptr(), in_affine[1].data_ptr()
    t.in_relu = int(in_relu)
    t.premul = _lib.ptr(premul)
    return t


@_lib.device_guarded
def head_to_planes(x: torch.Tensor, xf: "_lib.HeadXform", out_hi: torch.Tensor, out_lo: torch.Tensor,
                   choff: int, cpad: int) -> None:
    """fp32 NCHW -> NHWC hi/lo planes with the fused input transform; channels [choff, choff+cpad) of the
    planes receive the xf.c values followed by zeros (bhsr_head_to_planes)."""
    import ctypes as C
    nb = x.shape[0]
    assert out_hi.shape[:3] == (nb, xf.h, xf.w) and out_hi.dtype == torch.float16
    _lib.check(_lib.load().bhsr_head_to_planes(C.byref(xf), nb, out_hi.data_ptr(), out_lo.data_ptr(),
                                               out_hi.shape[3], choff, cpad, _lib.stream_ptr(x.device)),
               "bhsr_head_to_planes")


@_lib.device_guarded
def head_from_planes(in_hi: torch.Tensor, in_lo: torch.Tensor, c: int, y: torch.Tensor, *, choff: int = 0,
                     unscale: Optional[torch.Tensor] = None, accumulate: bool = False,
                     stats: Optional[torch.Tensor] = None) -> torch.Tensor:
    """NHWC hi/lo planes -> fp32 NCHW `y` (bhsr_head_from_planes): y (=|+=) value / *unscale, BatchNorm statistics."""
    nb, h, w, ctot = in_hi.shape
    assert y.dtype == torch.float32 and y.is_contiguous() and y.shape == (nb, c, h, w)
    assert stats is None or (stats.dtype == torch.float64 and stats.numel() == 2 * c)
    _lib.check(_lib.load().bhsr_head_from_planes(in_hi.data_ptr(), in_lo.data_ptr(), nb, h, w, ctot, choff, c,
                                                 _lib.ptr(unscale), y.data_ptr(), c, 0, int(accumulate),
                                                 _lib.ptr(stats), _lib.stream_ptr(in_hi.device)),
               "bhsr_head_from_planes")
    return y


@_lib.device_guarded
def channel_stats(y: torch.Tensor, stats: torch.Tensor) -> None:
    """stats[0:c] += sum, stats[c:2c] += sum of squares of the fp32 NCHW tensor y (BatchNorm batch statistics)."""
    nb, c, h, w = y.shape
    assert y.dtype == torch.float32 and y.is_contiguous() and stats.dtype == torch.float64 and stats.numel() == 2 * c
    _lib.check(_lib.load().bhsr_channel_stats(y.data_ptr(), c, 0, nb, c, h * w, stats.data_ptr(),
                                              _lib.stream_ptr(y.device)), "bhsr_channel_stats")


@_lib.device_guarded
def head_wgrad_tc(x: torch.Tensor, xf: "_lib.HeadXform", gf: "_lib.HeadXform", nb: int, cin: int, cout: int, k: int,
                  want_db: bool):
    """dw [cout, cin, k, k] (and db) on the tensor cores (bhsr_head_wgrad_tc)."""
    import ctypes as C
    lib = _lib.load()
    dev = x.device
    n = lib.bhsr_head_wgrad_workspace_bytes(nb, cin, cout, k, xf.h, xf.w)
    ws = torch.empty(n + 1024, dtype=torch.uint8, device=dev)
    ws_ptr = (ws.data_ptr() + 1023) // 1024 * 1024
    dw = torch.empty((cout, cin, k, k), dtype=torch.float32, device=dev)
    db_sum = torch.zeros(cout, dtype=torch.float64, device=dev) if want_db else None
    _lib.check(lib.bhsr_head_wgrad_tc(C.byref(xf), C.byref(gf), nb, k, dw.data_ptr(), _lib.ptr(db_sum), ws_ptr,
                                      ws.numel() - (ws_ptr - ws.data_ptr()), _lib.stream_ptr(dev)),
               "bhsr_head_wgrad_tc")
    return dw, (db_sum.float() if want_db else None)
