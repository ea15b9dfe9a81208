"""ctypes binding of libbhsr.so (C ABI declared in include/bhsr.h).

The library is the product: if it is missing, or a call fails, this module raises — there is no
PyTorch/CPU fallback anywhere in the package.  torch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# BHSR_LIB selects another build of the same library (e.g. the -DBHSR_TIMING profiling build)
LIB_PATH = os.environ.get("BHSR_LIB") or os.path.join(_HERE, "lib", "libbhsr.so")

NUMERICS_EXACT = 0  # BHSR_NUMERICS_EXACT_F16X3
NUMERICS_FAST = 1   # BHSR_NUMERICS_FAST_F16
NUMERICS = {"exact": NUMERICS_EXACT, "fast": NUMERICS_FAST}

EPI_LRELU = 1
EPI_RES1 = 2
EPI_RES2 = 4
EPI_OUT_NCHW_F32 = 8
EPI_RELU = 16
EPI_SHUFFLE2 = 32
EPI_ACCUM = 64


class BhsrError(RuntimeError):
    pass


class ConvTcDesc(C.Structure):
    """Mirror of `BhsrConvTcDesc` (include/bhsr.h)."""

    _fields_ = [
        ("in_hi", C.c_void_p), ("in_lo", C.c_void_p),
        ("nb", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
        ("in_ctot", C.c_int32), ("in_choff", C.c_int32), ("cin", C.c_int32),
        ("w_packed", C.c_void_p),
        ("cout", C.c_int32),
        ("bias", C.c_void_p),
        ("scale", C.c_void_p),
        ("cout_valid", C.c_int32),
        ("ntaps", C.c_int32),
        ("dy", C.c_int8 * 9), ("dx", C.c_int8 * 9),
        ("oh", C.c_int32), ("ow", C.c_int32), ("out_scale", C.c_int32),
        ("out_oy", C.c_int32), ("out_ox", C.c_int32),
        ("out_hi", C.c_void_p), ("out_lo", C.c_void_p),
        ("out_ctot", C.c_int32), ("out_choff", C.c_int32),
        ("out_f32", C.c_void_p),
        ("epilogue", C.c_int32),
        ("alpha1", C.c_float), ("alpha2", C.c_float),
        ("res1_hi", C.c_void_p), ("res1_lo", C.c_void_p),
        ("res1_ctot", C.c_int32), ("res1_choff", C.c_int32),
        ("res2_hi", C.c_void_p), ("res2_lo", C.c_void_p),
        ("res2_ctot", C.c_int32), ("res2_choff", C.c_int32),
        ("numerics", C.c_int32), ("mblocks", C.c_int32), ("max_ctas", C.c_int32),
        ("desc_mode", C.c_int32),
    ]


class RrdbNetDesc(C.Structure):
    """Mirror of `BhsrRrdbNetDesc` (include/bhsr.h)."""

    _fields_ = [
        ("num_in_ch", C.c_int32), ("num_out_ch", C.c_int32), ("num_block", C.c_int32),
        ("numerics", C.c_int32),
        ("nb", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
        ("conv_first_w", C.c_void_p), ("conv_first_b", C.c_void_p),
        ("conv_last_w", C.c_void_p), ("conv_last_b", C.c_void_p),
        ("packed", C.c_void_p), ("biases", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("mblocks", C.c_int32),
    ]


class HeadConvDesc(C.Structure):
    """Mirror of `BhsrHeadConvDesc` (include/bhsr.h)."""

    _fields_ = [
        ("x", C.c_void_p), ("x_ctot", C.c_int32), ("x_choff", C.c_int32),
        ("nb", C.c_int32), ("cin", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
        ("x_unshuffle", C.c_int32),
        ("in_scale", C.c_void_p), ("in_shift", C.c_void_p), ("in_relu", C.c_int32),
        ("weight", C.c_void_p), ("bias", C.c_void_p),
        ("cout", C.c_int32), ("ksize", C.c_int32),
        ("y", C.c_void_p), ("y_ctot", C.c_int32), ("y_choff", C.c_int32),
        ("y_shuffle", C.c_int32),
        ("stats", C.c_void_p),
        ("accumulate", C.c_int32),
    ]


class HeadXform(C.Structure):
    """Mirror of `BhsrHeadXform` (include/bhsr.h)."""

    _fields_ = [
        ("x", C.c_void_p), ("x_ctot", C.c_int32), ("x_choff", C.c_int32),
        ("c", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
        ("unshuffle", C.c_int32),
        ("in_scale", C.c_void_p), ("in_shift", C.c_void_p), ("in_relu", C.c_int32),
        ("premul", C.c_void_p),
    ]


# symbol -> (restype, argtypes); tests check every symbol of include/bhsr.h is listed here
# and exported by the shared object.
_SIGNATURES = {
    "bhsr_version": (C.c_int, []),
    "bhsr_last_error": (C.c_char_p, []),
    "bhsr_device_sm_count": (C.c_int, []),
    "bhsr_device_cc": (C.c_int, []),
    "bhsr_conv_tc": (C.c_int, [C.POINTER(ConvTcDesc), C.c_void_p]),
    "bhsr_debug_timing": (C.c_int, [C.c_void_p, C.c_int32]),
    "bhsr_packed_conv_weight_bytes": (C.c_size_t, [C.c_int32] * 4),
    "bhsr_pack_conv_weights": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                         C.c_void_p, C.c_void_p]),
    "bhsr_nchw_f32_to_planes": (C.c_int, [C.c_void_p] + [C.c_int32] * 4 + [C.c_void_p, C.c_void_p,
                                          C.c_int32, C.c_int32, C.c_void_p]),
    "bhsr_planes_to_nchw_f32": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int32] * 6 +
                                [C.c_void_p, C.c_void_p]),
    "bhsr_conv3x3_first": (C.c_int, [C.c_void_p] + [C.c_int64] * 4 + [C.c_int32] * 4 +
                           [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                            C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "bhsr_conv3x3_last": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int32] * 7 +
                          [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "bhsr_rrdbnet_packed_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "bhsr_rrdbnet_bias_floats": (C.c_size_t, [C.c_int32]),
    "bhsr_rrdbnet_workspace_bytes": (C.c_size_t, [C.c_int32] * 4),
    "bhsr_rrdbnet_pack": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                    C.c_void_p]),
    "bhsr_head_conv": (C.c_int, [C.POINTER(HeadConvDesc), C.c_void_p]),
    "bhsr_head_conv_wgrad": (C.c_int, [C.POINTER(HeadConvDesc), C.c_void_p, C.c_int32, C.c_int32,
                                       C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bhsr_bn_finalize": (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_void_p,
                                   C.c_float, C.c_float] + [C.c_void_p] * 7),
    "bhsr_bn_eval_affine": (C.c_int, [C.c_int32] + [C.c_void_p] * 4 + [C.c_float] + [C.c_void_p] * 4),
    "bhsr_affine_add_relu": (C.c_int, [C.c_void_p] * 4 + [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                       C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                       C.c_int32, C.c_void_p]),
    "bhsr_bn_bwd_reduce": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                     C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "bhsr_bn_bwd_coeffs": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_double] + [C.c_void_p] * 4 +
                           [C.c_int32] + [C.c_void_p] * 6),
    "bhsr_bn_bwd_apply": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32] +
                          [C.c_void_p] * 8 + [C.c_int32, C.c_int32] + [C.c_void_p] * 4 +
                          [C.c_int32] * 6 + [C.c_void_p]),
    "bhsr_head_to_planes": (C.c_int, [C.POINTER(HeadXform), C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                      C.c_int32, C.c_void_p]),
    "bhsr_head_from_planes": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int32] * 6 + [C.c_void_p, C.c_void_p, C.c_int32,
                                        C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "bhsr_channel_stats": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                     C.c_void_p]),
    "bhsr_head_wgrad_workspace_bytes": (C.c_size_t, [C.c_int32] * 6),
    "bhsr_head_wgrad_tc": (C.c_int, [C.POINTER(HeadXform), C.POINTER(HeadXform), C.c_int32, C.c_int32, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "bhsr_predict_postproc": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int32] * 4 + [C.c_void_p] * 3),
    "bhsr_weighted_mse": (C.c_int, [C.c_void_p] * 3 + [C.c_int64] + [C.c_void_p] * 6),
    "bhsr_ce_dice": (C.c_int, [C.c_void_p] * 3 + [C.c_int32] * 4 + [C.c_void_p] * 6),
    "bhsr_aggregate": (C.c_int, [C.c_void_p] + [C.c_int32] * 4 + [C.c_float, C.c_int32, C.c_void_p,
                                 C.c_void_p]),
    "bhsr_rrdbnet_forward": (C.c_int, [C.POINTER(RrdbNetDesc), C.c_void_p] + [C.c_int64] * 4 +
                             [C.c_void_p, C.c_int32, C.c_void_p]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load libbhsr.so (built by `__graft_entry__.build()` / `build.py`).  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BhsrError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` — this package has no fallback path")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def exported_symbols():
    return sorted(_SIGNATURES)


def check(rc: int, what: str = "") -> None:
    if rc < 0:
        msg = load().bhsr_last_error().decode("utf-8", "replace")
        raise BhsrError(f"{what or 'bhsr call'} failed ({rc}): {msg}")


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def on_device(t: torch.Tensor):
    """Context that makes `t`'s GPU the current CUDA device for the ctypes call inside it: the library
    launches with `<<<>>>` on the caller's stream and builds tensor maps / reads the SM count of the
    CURRENT device, so a tensor living on another GPU than the current one must switch first."""
    return torch.cuda.device(t.device)


def device_guarded(fn):
    """Decorator: run `fn` with the GPU of its first CUDA-tensor argument as the current device."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, torch.Tensor) and a.is_cuda:
                if a.device.index == torch.cuda.current_device():
                    break
                with torch.cuda.device(a.device):
                    return fn(*args, **kwargs)
        return fn(*args, **kwargs)

    return wrapper


# ------------------------------------------------------------------ packed-weight cache keys
# Packed weights / folded BatchNorm vectors are private caches keyed on (data_ptr, _version) of the
# source tensors.  `_version` is bumped by every in-place op on the tensor itself (optimizer steps,
# `copy_`, `load_state_dict`, `.to()` re-allocation) but NOT by writes through `.data` / `.detach()`
# aliases (EMA loops of the form `p.data.mul_(d).add_(...)`), nor by raw-pointer writes from a kernel.
# Three safety nets: (1) modules drop their caches in `_apply` and `_load_from_state_dict`;
# (2) `invalidate_cache(module)` for callers that mutate through `.data`; (3) BHSR_STRICT_CACHE=1 adds a
# device-side content fingerprint to every key (one host sync per forward — correctness over speed).
STRICT_CACHE = os.environ.get("BHSR_STRICT_CACHE", "0") == "1"


def tensor_key(tensors) -> tuple:
    ts = [t for t in tensors if t is not None]
    key = tuple((t.data_ptr(), t._version) for t in ts)
    if STRICT_CACHE and ts:
        fl = [t.detach().float() for t in ts if t.is_floating_point() and t.numel()]
        if fl:
            norms = torch._foreach_norm(fl)
            sums = [t.sum() for t in fl[:8]]           # sign-sensitive part on a few tensors
            key += (float(torch.stack(norms).double().sum()), float(torch.stack(sums).double().sum()))
    return key


def bump_version(*tensors) -> None:
    """Mark tensors a kernel wrote through their raw pointer as modified (autograd version counter)."""
    inc = getattr(torch._C, "_increment_version", None)
    for t in tensors:
        if t is None:
            continue
        if inc is not None:
            try:
                inc(t)
                continue
            except Exception:
                pass
        t.add_(0)


def invalidate_cache(module) -> None:
    """Drop every packed-weight / folded-BatchNorm cache under `module` (call after mutating parameters
    or buffers through `.data`, e.g. an EMA update `p.data.mul_(d).add_(q.data, alpha=1-d)`)."""
    for m in module.modules():
        m.__dict__.pop("_tc_cache", None)
        m.__dict__.pop("_tc_own", None)
        if "_packed" in m.__dict__:
            m.__dict__["_packed"] = None
        m.__dict__.pop("_graphs", None)


class CacheMixin:
    """nn.Module mixin: caches die with `.to()/.cuda()/.half()` (`_apply`) and with `load_state_dict`."""

    def invalidate_cache(self):
        invalidate_cache(self)

    def _apply(self, fn, *args, **kwargs):
        invalidate_cache(self)
        return super()._apply(fn, *args, **kwargs)

    def _load_from_state_dict(self, *args, **kwargs):
        invalidate_cache(self)
        return super()._load_from_state_dict(*args, **kwargs)


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise BhsrError(
            f"{name} must be a CUDA tensor (got device {t.device}); the B200 kernels have no CPU "
            "fallback — run the reference modules for a CPU oracle")
