"""Drop-in RRDBNet / Real-ESRGAN generator on the B200 kernels.

Mirrors the reference's nn.Module surface — constructor kwargs, attribute names, state_dict keys
and shapes, forward / forward_feature signatures (SR/rrdbnet_arch.py:113-240; the older
ESRGAN-style names of SR/RRDBNet.py:14-78 are in `OldRRDBNet`).  The nn.Conv2d children are
kept only as parameter containers (so `load_state_dict`, `.to()`, `.parameters()` and the
reference's RNG-order-dependent init behave identically); the arithmetic runs in libbhsr.so
through `bhsr_rrdbnet_forward`.  There is no PyTorch fallback: CPU tensors raise.

Numerics: `net.numerics = "exact"` (default; fp16x3 split products, matches the fp32 reference
within rtol 1e-3 / atol 1e-4) or `"fast"` (single fp16 product, TF32-class error).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional

import torch
from torch import nn
from torch.nn import init
from torch.nn.modules.batchnorm import _BatchNorm

from . import _lib, ops
from ._lib import NUMERICS


# ------------------------------------------------------------------ init helpers (reference API)
@torch.no_grad()
def default_init_weights(module_list, scale=1, bias_fill=0, **kwargs):
    """Same contract as SR/rrdbnet_arch.py:20-48 (kaiming_normal_ * scale, constant bias)."""
    if not isinstance(module_list, list):
        module_list = [module_list]
    for module in module_list:
        for m in module.modules():
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                init.kaiming_normal_(m.weight, **kwargs)
                m.weight.data *= scale
                if m.bias is not None:
                    m.bias.data.fill_(bias_fill)
            elif isinstance(m, _BatchNorm):
                init.constant_(m.weight, 1)
                if m.bias is not None:
                    m.bias.data.fill_(bias_fill)


def make_layer(basic_block, num_basic_block, **kwarg):
    """SR/rrdbnet_arch.py:51-64."""
    return nn.Sequential(*[basic_block(**kwarg) for _ in range(num_basic_block)])


def pixel_unshuffle(x, scale):
    """SR/rrdbnet_arch.py:94-110 — a pure index permutation, kept as torch view ops (bit-exact)."""
    b, c, hh, hw = x.size()
    out_channel = c * (scale ** 2)
    assert hh % scale == 0 and hw % scale == 0
    h = hh // scale
    w = hw // scale
    x_view = x.view(b, c, h, scale, w, scale)
    return x_view.permute(0, 1, 3, 5, 2, 4).reshape(b, out_channel, h, w)


def _default_numerics() -> str:
    mode = os.environ.get("BHSR_NUMERICS", "exact")
    if mode not in NUMERICS:
        raise ValueError(f"BHSR_NUMERICS must be one of {sorted(NUMERICS)} (got {mode!r})")
    return mode


def _wants_autograd(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def _no_autograd(name: str, *tensors) -> None:
    if _wants_autograd(*tensors):
        raise NotImplementedError(f"{name}: this entry is the frozen (no-grad) schedule; the autograd path is "
                                  "rrdbnet_train.py — call the module's forward / forward_feature instead")


def _block_autograd(convs, x, rrdb: bool):
    """A stand-alone dense block / RRDB under autograd: exact numerics, tensor-core backward (rrdbnet_train.py)."""
    from .rrdbnet_train import RDBChainTrainFn
    params = []
    for c in convs:
        if c.bias is None:
            raise NotImplementedError("ResidualDenseBlock / RRDB under autograd: convs without bias are not supported")
        params += [c.weight, c.bias]
    return RDBChainTrainFn.apply(convs, x, rrdb, *params)


# ------------------------------------------------------------------ blocks (parameter containers)
class _PlanesRunner:
    """Runs single RDB / RRDB blocks through the per-conv C entry (used when a block is called on
    its own; the full net goes through the one-call schedule instead)."""

    @staticmethod
    def rdb(convs: List[nn.Conv2d], cur, nxt, numerics: int, rrdb_in=None) -> None:
        """convs: conv1..conv5; cur/nxt: (hi, lo) 192-channel planes; result -> nxt[..., :64]."""
        for c, conv in enumerate(convs):
            cin = 64 + 32 * c
            wp = ops.pack_conv_weights(conv.weight.detach(), numerics)
            bias = conv.bias.detach() if conv.bias is not None else None
            if c < 4:
                ops.conv_tc(cur[0], cur[1], 0, cin, wp, 32, bias, ops.PLAIN_TAPS, cur[0], cur[1],
                            out_choff=cin, lrelu=True, numerics=numerics)
            else:
                kw = {}
                if rrdb_in is not None:
                    kw = dict(res2=(rrdb_in[0], rrdb_in[1], 0), alpha2=0.2)
                ops.conv_tc(cur[0], cur[1], 0, cin, wp, 64, bias, ops.PLAIN_TAPS, nxt[0], nxt[1],
                            out_choff=0, res1=(cur[0], cur[1], 0), alpha1=0.2, numerics=numerics, **kw)


def _new_planes(nb, h, w, c, device):
    return (torch.zeros((nb, h, w, c), dtype=torch.float16, device=device),
            torch.zeros((nb, h, w, c), dtype=torch.float16, device=device))


class ResidualDenseBlock(nn.Module):
    """SR/rrdbnet_arch.py:113-143."""

    def __init__(self, num_feat=64, num_grow_ch=32):
        super().__init__()
        self.conv1 = nn.Conv2d(num_feat, num_grow_ch, 3, 1, 1)
        self.conv2 = nn.Conv2d(num_feat + num_grow_ch, num_grow_ch, 3, 1, 1)
        self.conv3 = nn.Conv2d(num_feat + 2 * num_grow_ch, num_grow_ch, 3, 1, 1)
        self.conv4 = nn.Conv2d(num_feat + 3 * num_grow_ch, num_grow_ch, 3, 1, 1)
        self.conv5 = nn.Conv2d(num_feat + 4 * num_grow_ch, num_feat, 3, 1, 1)
        self.lrelu = nn.LeakyReLU(negative_slope=0.2, inplace=True)
        default_init_weights([self.conv1, self.conv2, self.conv3, self.conv4, self.conv5], 0.1)
        self.numerics = _default_numerics()

    def _convs(self):
        return [self.conv1, self.conv2, self.conv3, self.conv4, self.conv5]

    def forward(self, x):
        _lib.require_cuda(x, "x")
        _check_widths(self.conv1.in_channels, self.conv1.out_channels)
        if _wants_autograd(x, *self.parameters()):
            return _block_autograd(self._convs(), x, rrdb=False)
        nb, _, h, w = x.shape
        cur = _new_planes(nb, h, w, 192, x.device)
        nxt = _new_planes(nb, h, w, 192, x.device)
        ops.nchw_to_planes(x.float().contiguous(), cur[0], cur[1], 0)
        _PlanesRunner.rdb(self._convs(), cur, nxt, NUMERICS[self.numerics])
        return ops.planes_to_nchw(nxt[0], nxt[1], 64, 0)


class RRDB(nn.Module):
    """SR/rrdbnet_arch.py:146-167."""

    def __init__(self, num_feat, num_grow_ch=32):
        super().__init__()
        self.rdb1 = ResidualDenseBlock(num_feat, num_grow_ch)
        self.rdb2 = ResidualDenseBlock(num_feat, num_grow_ch)
        self.rdb3 = ResidualDenseBlock(num_feat, num_grow_ch)
        self.numerics = _default_numerics()

    def forward(self, x):
        _lib.require_cuda(x, "x")
        _check_widths(self.rdb1.conv1.in_channels, self.rdb1.conv1.out_channels)
        if _wants_autograd(x, *self.parameters()):
            return _block_autograd(self.rdb1._convs() + self.rdb2._convs() + self.rdb3._convs(), x, rrdb=True)
        nb, _, h, w = x.shape
        bufs = [_new_planes(nb, h, w, 192, x.device) for _ in range(3)]
        ops.nchw_to_planes(x.float().contiguous(), bufs[0][0], bufs[0][1], 0)
        num = NUMERICS[self.numerics]
        _PlanesRunner.rdb(self.rdb1._convs(), bufs[0], bufs[1], num)
        _PlanesRunner.rdb(self.rdb2._convs(), bufs[1], bufs[2], num)
        _PlanesRunner.rdb(self.rdb3._convs(), bufs[2], bufs[0], num, rrdb_in=bufs[0])
        return ops.planes_to_nchw(bufs[0][0], bufs[0][1], 64, 0)


def _check_widths(num_feat, num_grow_ch):
    if num_feat != 64 or num_grow_ch != 32:
        raise NotImplementedError(
            f"the B200 kernels are built for num_feat=64, num_grow_ch=32 (the only widths the "
            f"reference instantiates); got {num_feat}/{num_grow_ch}")


# ------------------------------------------------------------------ the network
class _RRDBNetBase(_lib.CacheMixin, nn.Module):
    """Shared engine: subclasses define the child-module names."""

    # attribute names: (conv_first, body, conv_body, conv_up1, conv_up2, conv_hr, conv_last)
    _NAMES = ("conv_first", "body", "conv_body", "conv_up1", "conv_up2", "conv_hr", "conv_last")
    _RDB_NAMES = ("rdb1", "rdb2", "rdb3")

    def _engine_init(self):
        self.numerics = _default_numerics()
        self.mblocks = int(os.environ.get("BHSR_MBLOCKS", "0"))  # 0 = auto (2)
        self._packed = None       # (key, packed, biases)
        self._workspace = {}      # (device, nb, h, w, feature) -> uint8 tensor
        self.use_cuda_graph = False   # opt-in: replay the whole forward from a CUDA graph (see _run)

    # -- parameter plumbing
    def _child(self, i):
        return getattr(self, self._NAMES[i])

    def _tc_convs(self) -> List[nn.Conv2d]:
        convs = []
        for blk in self._child(1):
            for rn in self._RDB_NAMES:
                rdb = getattr(blk, rn)
                convs += [rdb.conv1, rdb.conv2, rdb.conv3, rdb.conv4, rdb.conv5]
        convs += [self._child(2), self._child(3), self._child(4), self._child(5)]
        return convs

    def _cache_key(self, convs, device, numerics):
        # see _lib.tensor_key for what invalidates the packed-weight cache (and what cannot: `.data` writes)
        return (str(device), numerics) + _lib.tensor_key([t for c in convs for t in (c.weight, c.bias)])

    def _get_packed(self, device):
        lib = _lib.load()
        convs = self._tc_convs()
        numerics = NUMERICS[self.numerics]
        key = self._cache_key(convs, device, numerics)
        if self._packed is not None and self._packed[0] == key:
            return self._packed[1], self._packed[2]
        nblk = len(self._child(1))
        plist = []
        keep = []
        for c in convs:
            if c.bias is None:
                raise _lib.BhsrError("RRDBNet convs must have biases (the reference's all do)")
            for p in (c.weight, c.bias):
                t = p.detach()
                if t.dtype != torch.float32 or not t.is_contiguous() or t.device != device:
                    t = t.to(device=device, dtype=torch.float32).contiguous()
                keep.append(t)
                plist.append(t.data_ptr())
        arr = (C.c_void_p * len(plist))(*plist)
        packed = torch.empty(lib.bhsr_rrdbnet_packed_bytes(nblk, numerics), dtype=torch.uint8, device=device)
        biases = torch.empty(lib.bhsr_rrdbnet_bias_floats(nblk), dtype=torch.float32, device=device)
        _lib.check(lib.bhsr_rrdbnet_pack(arr, nblk, numerics, packed.data_ptr(), biases.data_ptr(),
                                         _lib.stream_ptr(device)), "bhsr_rrdbnet_pack")
        del keep
        self._packed = (key, packed, biases)
        return packed, biases

    def _get_workspace(self, device, nb, h, w, feature):
        key = (str(device), nb, h, w, bool(feature))
        ws = self._workspace.get(key)
        if ws is None:
            if len(self._workspace) > 4:  # a few shapes (train / val / tail batch) at most
                self._workspace.clear()
            n = _lib.load().bhsr_rrdbnet_workspace_bytes(nb, h, w, int(feature))
            # zero-filled once: unused channel chunks are read by TMA (never by an MMA)
            ws = torch.zeros(n + 1024, dtype=torch.uint8, device=device)
            self._workspace[key] = ws
        return ws

    # -- the hot call
    def _run(self, x: torch.Tensor, feature: bool, scale: int) -> torch.Tensor:
        """Eager schedule (356 launches at 23 blocks), or — with `self.use_cuda_graph = True` — one CUDA-graph launch.

        Graph mode keeps one captured graph per (input address / shape / strides, feature, numerics, weights) and
        replays it: the kernels, their tensor maps and the OUTPUT buffer are baked in, so the returned tensor is
        overwritten by the next call with the same input buffer (callers that keep results must clone them).  It is
        opt-in because of that aliasing; bench.py and the sharded predictor, which consume each result before the
        next step, turn it on."""
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            return self._run_train(x, feature, scale)
        if getattr(self, "use_cuda_graph", False) and x.is_cuda and not torch.cuda.is_current_stream_capturing():
            y = self._run_graphed(x, feature, scale)
            if y is not None:
                return y
        return self._run_eager(x, feature, scale)

    def _run_graphed(self, x, feature, scale):
        convs = self._tc_convs()
        numerics = NUMERICS[self.numerics]
        key = (x.data_ptr(), tuple(x.shape), tuple(x.stride()), x.dtype, bool(feature), scale, self.mblocks,
               self._cache_key(convs, x.device, numerics))
        graphs = self.__dict__.setdefault("_graphs", {})
        entry = graphs.get(key)
        if entry is None:
            if len(graphs) >= 8:
                graphs.clear()
            try:
                self._run_eager(x, feature, scale)            # warm the packed-weight / workspace caches
                torch.cuda.current_stream(x.device).synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    y = self._run_eager(x, feature, scale)
                # everything whose ADDRESS the graph baked in must outlive it: the input buffer, the packed weights
                # / biases (re-packed into new tensors when `numerics` or the parameters change) and the workspaces
                entry = (g, y, (x, self._packed, dict(self._workspace)))
            except Exception as e:                             # capture refused: remember, fall back to eager launches
                entry = (None, None, str(e))
                torch.cuda.synchronize(x.device)
            graphs[key] = entry
        if entry[0] is None:
            return None
        entry[0].replay()
        return entry[1]

    def _run_train(self, x: torch.Tensor, feature: bool, scale: int) -> torch.Tensor:
        """Forward under autograd (SR fine-tuning, SR/rrdbnet_arch.py:538-592): layer by layer through the same
        kernels in exact numerics, keeping the activations; backward = tensor-core dgrad / wgrad (rrdbnet_train.py).
        The frozen pipeline (torch.no_grad(), train.py:243-244) never comes here."""
        from .rrdbnet_train import RRDBNetTrainFn
        _lib.require_cuda(x, "x")
        first, last = self._child(0), self._child(6)
        _check_widths(first.out_channels, self._tc_convs()[0].out_channels if len(self._child(1)) else 32)
        if x.dtype != torch.float32:
            x = x.float()
        if scale == 2:
            x = pixel_unshuffle(x, 2)
        elif scale == 1:
            x = pixel_unshuffle(x, 4)
        if x.shape[1] != first.in_channels:
            raise RuntimeError(f"expected input with {first.in_channels} channels, got {x.shape[1]}")
        if x.shape[0] == 0 or x.shape[2] == 0 or x.shape[3] == 0:
            raise RuntimeError("RRDBNet under autograd needs a non-empty batch")
        convs = [first] + self._tc_convs() + [last]
        for c in convs:
            if c.bias is None:
                raise _lib.BhsrError("RRDBNet convs must have biases (the reference's all do)")
        params = [t for c in convs for t in (c.weight, c.bias)]
        return RRDBNetTrainFn.apply(convs, x, bool(feature), *params)

    def _run_eager(self, x: torch.Tensor, feature: bool, scale: int) -> torch.Tensor:
        _lib.require_cuda(x, "x")
        _no_autograd(type(self).__name__, x, *self.parameters())
        first, last = self._child(0), self._child(6)
        _check_widths(first.out_channels, self._tc_convs()[0].out_channels if len(self._child(1)) else 32)
        if x.dtype != torch.float32:
            x = x.float()
        if scale == 2:
            x = pixel_unshuffle(x, 2)
        elif scale == 1:
            x = pixel_unshuffle(x, 4)
        nb, cin, h, w = x.shape
        if cin != first.in_channels:
            raise RuntimeError(f"expected input with {first.in_channels} channels, got {cin}")
        dev = x.device
        if nb == 0 or h == 0 or w == 0:  # empty batch: nothing to launch (nn.Conv2d returns empty too)
            return torch.empty((nb, 64 if feature else last.out_channels, 4 * h, 4 * w), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            packed, biases = self._get_packed(dev)
            ws = self._get_workspace(dev, nb, h, w, feature)
            ws_ptr = (ws.data_ptr() + 1023) // 1024 * 1024
            d = _lib.RrdbNetDesc()
            d.num_in_ch, d.num_out_ch = cin, last.out_channels
            d.num_block = len(self._child(1))
            d.numerics = NUMERICS[self.numerics]
            d.nb, d.h, d.w = nb, h, w
            fw = first.weight.detach().to(dev, torch.float32).contiguous()
            fb = first.bias.detach().to(dev, torch.float32).contiguous()
            lw = last.weight.detach().to(dev, torch.float32).contiguous()
            lb = last.bias.detach().to(dev, torch.float32).contiguous()
            d.conv_first_w, d.conv_first_b = fw.data_ptr(), fb.data_ptr()
            d.conv_last_w, d.conv_last_b = lw.data_ptr(), lb.data_ptr()
            d.packed, d.biases = packed.data_ptr(), biases.data_ptr()
            d.workspace = ws_ptr
            d.workspace_bytes = ws.numel() - (ws_ptr - ws.data_ptr())
            d.mblocks = self.mblocks
            cout = 64 if feature else last.out_channels
            y = torch.empty((nb, cout, 4 * h, 4 * w), dtype=torch.float32, device=dev)
            sn, sc, sh, sw = x.stride()
            _lib.check(_lib.load().bhsr_rrdbnet_forward(C.byref(d), x.data_ptr(), sn, sc, sh, sw,
                                                        y.data_ptr(), int(feature),
                                                        _lib.stream_ptr(dev)),
                       "bhsr_rrdbnet_forward")
        return y


class RRDBNet(_RRDBNetBase):
    """SR/rrdbnet_arch.py:170-240 — same ctor, children, state_dict (702 tensors at 23 blocks)."""

    def __init__(self, num_in_ch, num_out_ch, scale=4, num_feat=64, num_block=23, num_grow_ch=32):
        super().__init__()
        self.scale = scale
        if scale == 2:
            num_in_ch = num_in_ch * 4
        elif scale == 1:
            num_in_ch = num_in_ch * 16
        self.conv_first = nn.Conv2d(num_in_ch, num_feat, 3, 1, 1)
        self.body = make_layer(RRDB, num_block, num_feat=num_feat, num_grow_ch=num_grow_ch)
        self.conv_body = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.conv_up1 = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.conv_up2 = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.conv_hr = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.conv_last = nn.Conv2d(num_feat, num_out_ch, 3, 1, 1)
        self.lrelu = nn.LeakyReLU(negative_slope=0.2, inplace=True)
        self._engine_init()

    def forward(self, x):
        """[B, C, H, W] -> [B, num_out_ch, 4H', 4W'] (rrdbnet_arch.py:208-223)."""
        return self._run(x, feature=False, scale=self.scale)

    def forward_feature(self, x):
        """[B, C, H, W] -> pre-activation conv_hr feature map [B, 64, 4H', 4W'] (:225-240)."""
        return self._run(x, feature=True, scale=self.scale)


class _OldRDB(nn.Module):
    """ResidualDenseBlock_5C, SR/RRDBNet.py:14-34 (parameter container)."""

    def __init__(self, nf=64, gc=32, bias=True):
        super().__init__()
        self.conv1 = nn.Conv2d(nf, gc, 3, 1, 1, bias=bias)
        self.conv2 = nn.Conv2d(nf + gc, gc, 3, 1, 1, bias=bias)
        self.conv3 = nn.Conv2d(nf + 2 * gc, gc, 3, 1, 1, bias=bias)
        self.conv4 = nn.Conv2d(nf + 3 * gc, gc, 3, 1, 1, bias=bias)
        self.conv5 = nn.Conv2d(nf + 4 * gc, nf, 3, 1, 1, bias=bias)
        self.lrelu = nn.LeakyReLU(negative_slope=0.2, inplace=True)


class _OldRRDB(nn.Module):
    """RRDB, SR/RRDBNet.py:37-50."""

    def __init__(self, nf, gc=32):
        super().__init__()
        self.RDB1 = _OldRDB(nf, gc)
        self.RDB2 = _OldRDB(nf, gc)
        self.RDB3 = _OldRDB(nf, gc)


class OldRRDBNet(_RRDBNetBase):
    """SR/RRDBNet.py:53-78 — the ESRGAN-era naming of the same network (no forward_feature)."""

    _NAMES = ("conv_first", "RRDB_trunk", "trunk_conv", "upconv1", "upconv2", "HRconv", "conv_last")
    _RDB_NAMES = ("RDB1", "RDB2", "RDB3")

    def __init__(self, in_nc=4, out_nc=3, nf=64, nb=23, gc=32):
        super().__init__()
        self.conv_first = nn.Conv2d(in_nc, nf, 3, 1, 1, bias=True)
        self.RRDB_trunk = nn.Sequential(*[_OldRRDB(nf, gc) for _ in range(nb)])
        self.trunk_conv = nn.Conv2d(nf, nf, 3, 1, 1, bias=True)
        self.upconv1 = nn.Conv2d(nf, nf, 3, 1, 1, bias=True)
        self.upconv2 = nn.Conv2d(nf, nf, 3, 1, 1, bias=True)
        self.HRconv = nn.Conv2d(nf, nf, 3, 1, 1, bias=True)
        self.conv_last = nn.Conv2d(nf, out_nc, 3, 1, 1, bias=True)
        self.lrelu = nn.LeakyReLU(negative_slope=0.2, inplace=True)
        self._engine_init()

    def forward(self, x):
        return self._run(x, feature=False, scale=4)


# ------------------------------------------------------------------ RealESRGAN shell
class RealESRGAN:
    """SR/rrdbnet_arch.py:437-592 on the B200 path: `.net_g` for the height pipeline and, with `is_train=True`, the
    generator half of the SR fine-tune step.

    The callers on the hot path (train.py:133-140, predict_realesanet_feature_globe.py:95-102) only touch `net_g`
    (load_state_dict / eval / parameters / forward_feature); the default construction therefore stays free of the
    reference constructor's side effects (it always builds an EMA copy, a U-Net discriminator, USM sharpening on CUDA, a
    VGG19 perceptual loss — a network download — and two optimisers).

    `is_train=True` adds what `optimize_parameters` (:538-592) needs for the GENERATOR update, which
    is where the B200 kernels are (forward + tensor-core backward of RRDBNet, rrdbnet_train.py): `net_g_ema` (:459-478),
    `cri_pix = nn.L1Loss()` (:493), `optimizer_g = Adam(lr 1e-4, betas (0.9, 0.99))` (:502), its MultiStepLR (:505),
    `feed_data`, `model_ema` (:531-536) and `optimize_parameters`.  The discriminator, the VGG perceptual loss, the GAN
    loss and the USM sharpener are stock PyTorch modules outside this repo's scope: assign `net_d` / `optimizer_d` /
    `cri_perceptual` / `cri_gan` / `usm_sharpener` to use them — `optimize_parameters` then runs the reference's
    sequence including the discriminator update; unset, their terms are skipped (pixel loss only) and said so in the
    returned dict.
    """

    _OPTIONAL = ("net_d", "optimizer_d", "usm_sharpener", "cri_perceptual", "cri_gan")
    _TRAIN_ONLY = ("net_g_ema", "cri_pix", "optimizer_g", "optimizers", "schedulers")

    def __init__(self, in_ch=3, out_ch=3, num_block=23, device='cuda', scale=4, ema_decay=0.999,
                 pretrain_g_path=None, pretrain_d_path=None, is_train=False):
        self.device = device
        self.scale = scale
        self.ema_decay = ema_decay
        self.is_train = bool(is_train)

        def make():
            return RRDBNet(num_in_ch=in_ch, num_out_ch=out_ch, num_feat=64, num_block=num_block,
                           num_grow_ch=32, scale=scale).to(device)

        def pretrained():
            weights = torch.load(pretrain_g_path, map_location=device)['params_ema']
            if in_ch == 1:  # same averaging as rrdbnet_arch.py:451-454
                weights['conv_first.weight'] = torch.mean(weights['conv_first.weight'], dim=1, keepdim=True)
                weights['conv_last.weight'] = torch.mean(weights['conv_last.weight'], dim=0, keepdim=True)
                weights['conv_last.bias'] = torch.mean(weights['conv_last.bias'], dim=0, keepdim=True)
            return weights

        self.net_g = make()
        if pretrain_g_path is not None:
            self.net_g.load_state_dict(pretrained())
        if self.ema_decay > 0:
            print(f'Use Exponential Moving Average with decay: {self.ema_decay}')
        self.net_g.train()
        if not self.is_train:
            return
        if pretrain_d_path is not None:
            raise NotImplementedError("RealESRGAN(pretrain_d_path=...): the discriminator is a stock PyTorch module outside "
                                      "this repo; build it and assign `.net_d` / `.optimizer_d`")
        if self.ema_decay > 0:
            self.net_g_ema = make()
            if pretrain_g_path is not None:
                self.net_g_ema.load_state_dict(pretrained())
            else:
                self.model_ema(0)  # copy net_g weight (:476)
            for p in self.net_g_ema.parameters():
                p.requires_grad = False
        self.cri_pix = nn.L1Loss().to(device)
        self.net_d_iters = 1
        self.net_d_init_iters = 0
        # capturable: the step counters live on the device, so the whole generator step can replay from a CUDA graph
        self.optimizer_g = torch.optim.Adam(params=self.net_g.parameters(), lr=1e-4, betas=(0.9, 0.99), weight_decay=0,
                                            capturable=str(device).startswith("cuda"))
        self.use_cuda_graph = False      # opt-in: see _optimize_parameters_graphed
        self.optimizers = [self.optimizer_g]
        self.schedulers = [torch.optim.lr_scheduler.MultiStepLR(self.optimizer_g, milestones=[400000], gamma=0.5)]

    def __getattr__(self, name):
        if name in RealESRGAN._OPTIONAL:
            return None                      # stock PyTorch pieces the caller may plug in
        if name in RealESRGAN._TRAIN_ONLY:
            raise AttributeError(
                f"RealESRGAN.{name}: SR fine-tuning state is only built with is_train=True"
                + (" and ema_decay > 0" if name == "net_g_ema" else ""))
        raise AttributeError(name)

    @torch.no_grad()
    def feed_data(self, data):
        """:522-528 (USM sharpening only when a `usm_sharpener` module was assigned)."""
        self.lq = data['lq'].to(self.device, non_blocking=True)
        self.gt = data['gt'].to(self.device, non_blocking=True)
        self.gt_usm = self.usm_sharpener(self.gt) if self.usm_sharpener is not None else self.gt

    @torch.no_grad()
    def model_ema(self, decay=0.999):
        """:530-536.  The reference writes through `.data`, which does not bump tensor versions: the EMA copy's packed-
        weight cache is invalidated explicitly."""
        net_g_params = dict(self.net_g.named_parameters())
        net_g_ema_params = dict(self.net_g_ema.named_parameters())
        for k in net_g_ema_params.keys():
            net_g_ema_params[k].data.mul_(decay).add_(net_g_params[k].data, alpha=1 - decay)
        self.net_g_ema.invalidate_cache()

    def optimize_parameters(self):
        """:538-592.  Generator: output = net_g(lq) -> pixel (+ perceptual + GAN when plugged in) -> backward through the
        B200 kernels -> Adam; discriminator update when `net_d` / `optimizer_d` / `cri_gan` are plugged in; EMA."""
        from collections import OrderedDict
        l1_gt = percep_gt = self.gt_usm
        gan_gt = self.gt
        have_d = self.net_d is not None and self.cri_gan is not None
        if self.use_cuda_graph and not have_d and self.cri_perceptual is None and self.lq.is_cuda:
            return self._optimize_parameters_graphed()
        if have_d:
            for p in self.net_d.parameters():
                p.requires_grad = False
        self.optimizer_g.zero_grad()
        self.output = self.net_g(self.lq)
        loss_dict = OrderedDict()
        l_g_pix = self.cri_pix(self.output, l1_gt)
        l_g_total = l_g_pix
        loss_dict['l_g_pix'] = l_g_pix.item()
        if self.cri_perceptual is not None:
            l_g_percep = self.cri_perceptual(self.output, percep_gt)
            l_g_total = l_g_total + l_g_percep
            loss_dict['l_g_percep'] = l_g_percep.item()
        if have_d:
            fake_g_pred = self.net_d(self.output)
            l_g_gan = self.cri_gan(fake_g_pred, True, is_disc=False)
            l_g_total = l_g_total + l_g_gan
            loss_dict['l_g_gan'] = l_g_gan.item()
        l_g_total.backward()
        self.optimizer_g.step()
        if have_d and self.optimizer_d is not None:
            for p in self.net_d.parameters():
                p.requires_grad = True
            self.optimizer_d.zero_grad()
            real_d_pred = self.net_d(gan_gt)
            l_d_real = self.cri_gan(real_d_pred, True, is_disc=True)
            loss_dict['l_d_real'] = l_d_real.item()
            loss_dict['out_d_real'] = torch.mean(real_d_pred.detach())
            l_d_real.backward()
            fake_d_pred = self.net_d(self.output.detach().clone())
            l_d_fake = self.cri_gan(fake_d_pred, False, is_disc=True)
            loss_dict['l_d_fake'] = l_d_fake.item()
            loss_dict['out_d_fake'] = torch.mean(fake_d_pred.detach())
            l_d_fake.backward()
            self.optimizer_d.step()
        else:
            loss_dict['skipped'] = 'perceptual / GAN / discriminator terms: no net_d / cri_perceptual / cri_gan assigned'
        if self.ema_decay > 0:
            self.model_ema(decay=self.ema_decay)
        return loss_dict

    def _optimize_parameters_graphed(self):
        """The pixel-loss generator step (forward, L1, backward, Adam, EMA) as ONE CUDA-graph launch.  Eagerly the step is
        ~9,000 launches issued from Python (260 ms at any batch size on a B200 host); the kernels themselves need a
        fraction of that.  The first two calls run eagerly (allocator / cuBLAS / tensor-map warm-up, real training
        steps), the third captures, later calls copy `lq` / `gt_usm` into the captured buffers and replay.  The graph is
        re-captured when the input shape or the learning rate changes.  Replays update the parameters without bumping
        their autograd versions, so the packed-weight caches of both networks are dropped after every step."""
        from collections import OrderedDict
        st = self.__dict__.setdefault("_gg", {"calls": 0, "graph": None})
        lr_now = tuple(g["lr"] for g in self.optimizer_g.param_groups)
        key = (tuple(self.lq.shape), tuple(self.gt_usm.shape), lr_now)
        if st["graph"] is not None and st["key"] != key:
            st.update(graph=None, calls=1)
        if st["graph"] is None and st["calls"] < 2:
            st["calls"] += 1
            flag, self.use_cuda_graph = self.use_cuda_graph, False
            # on a side stream: gradients first allocated on the legacy default stream would make autograd tie the
            # capture to that stream ("legacy stream depend on a capturing blocking stream")
            for p in self.net_g.parameters():
                p.grad = None
            side = st.setdefault("side", torch.cuda.Stream())
            side.wait_stream(torch.cuda.current_stream())
            try:
                with torch.cuda.stream(side):
                    out = self.optimize_parameters()
            finally:
                self.use_cuda_graph = flag
            torch.cuda.current_stream().wait_stream(side)
            return out
        if st["graph"] is None:
            st["lq"], st["gt"] = self.lq.clone(), self.gt_usm.clone()
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self.optimizer_g.zero_grad(set_to_none=False)
                out = self.net_g(st["lq"])
                loss = self.cri_pix(out, st["gt"])
                loss.backward()
                self.optimizer_g.step()
                if self.ema_decay > 0:
                    self.model_ema(decay=self.ema_decay)
                st["out"], st["loss"] = out.detach(), loss.detach()
            st.update(graph=graph, key=key)
        st["lq"].copy_(self.lq, non_blocking=True)
        st["gt"].copy_(self.gt_usm, non_blocking=True)
        st["graph"].replay()
        self.net_g.invalidate_cache()
        if self.ema_decay > 0:
            self.net_g_ema.invalidate_cache()
        self.output = st["out"]
        return OrderedDict(l_g_pix=st["loss"].item(), launch="cuda-graph",
                           skipped='perceptual / GAN / discriminator terms: no net_d / cri_perceptual / cri_gan assigned')

    def save(self, epoch, current_iter, respath):
        """:508-519 (generator file; the discriminator file only when a `net_d` was assigned)."""
        torch.save({'params': self.net_g.state_dict(),
                    'params_ema': self.net_g_ema.state_dict() if 'net_g_ema' in self.__dict__ else None,
                    'epoch': epoch, 'current_iter': current_iter}, os.path.join(respath, 'net_g.tar'))
        if self.net_d is not None:
            torch.save({'params': self.net_d.state_dict(), 'epoch': epoch, 'current_iter': current_iter},
                       os.path.join(respath, 'net_d.tar'))
