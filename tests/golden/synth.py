"""Seeded synthetic parameters and inputs shared by the golden-vector generator
(make_golden.py, which feeds them to the REFERENCE modules) and by the tests (which feed the
same arrays to the oracle and to the CUDA path).  numpy RandomState only, so the arrays do not
depend on torch's RNG or on module construction order.

Scales are chosen to look like a trained net rather than the reference's 0.1-scaled init: RDB
convs have fan-in-normalised weights with gain ~0.7 so every layer contributes visibly to the
output, all biases are non-zero, BatchNorm affine/running stats are non-trivial.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict

import numpy as np


def _conv_w(rng, cout, cin, k, gain):
    std = gain / np.sqrt(cin * k * k)
    return (rng.standard_normal((cout, cin, k, k)) * std).astype(np.float32)


def _bias(rng, c, s=0.05):
    return (rng.standard_normal(c) * s).astype(np.float32)


def rrdbnet_state(num_in_ch=3, num_out_ch=3, scale=4, num_feat=64, num_block=23, num_grow_ch=32,
                  seed=0, rdb_gain=0.7) -> "OrderedDict[str, np.ndarray]":
    """state_dict of RRDBNet (SR/rrdbnet_arch.py:190-206), reference key order."""
    rng = np.random.RandomState(seed)
    cin = num_in_ch * (4 if scale == 2 else 16 if scale == 1 else 1)
    sd = OrderedDict()
    sd["conv_first.weight"] = _conv_w(rng, num_feat, cin, 3, 1.0)
    sd["conv_first.bias"] = _bias(rng, num_feat)
    for b in range(num_block):
        for r in (1, 2, 3):
            for c in range(1, 6):
                ci = num_feat + (c - 1) * num_grow_ch
                co = num_grow_ch if c < 5 else num_feat
                sd[f"body.{b}.rdb{r}.conv{c}.weight"] = _conv_w(rng, co, ci, 3, rdb_gain)
                sd[f"body.{b}.rdb{r}.conv{c}.bias"] = _bias(rng, co)
    for name in ("conv_body", "conv_up1", "conv_up2", "conv_hr"):
        sd[f"{name}.weight"] = _conv_w(rng, num_feat, num_feat, 3, 1.0)
        sd[f"{name}.bias"] = _bias(rng, num_feat)
    sd["conv_last.weight"] = _conv_w(rng, num_out_ch, num_feat, 3, 1.0)
    sd["conv_last.bias"] = _bias(rng, num_out_ch)
    return sd


_NEW_TO_OLD = (("body.", "RRDB_trunk."), (".rdb1.", ".RDB1."), (".rdb2.", ".RDB2."),
               (".rdb3.", ".RDB3."), ("conv_body.", "trunk_conv."), ("conv_up1.", "upconv1."),
               ("conv_up2.", "upconv2."), ("conv_hr.", "HRconv."))


def to_old_rrdbnet_keys(sd):
    """Rename to the SR/RRDBNet.py:53-67 attribute names."""
    out = OrderedDict()
    for k, v in sd.items():
        if k.startswith("body."):
            k = "RRDB_trunk." + k[len("body."):]
            for a, b in _NEW_TO_OLD[1:4]:
                k = k.replace(a, b)
        else:
            for a, b in _NEW_TO_OLD[4:]:
                if k.startswith(a):
                    k = b + k[len(a):]
        out[k] = v
    return out


def _bn(rng, sd, prefix, c):
    sd[prefix + ".weight"] = (1.0 + 0.2 * rng.standard_normal(c)).astype(np.float32)
    sd[prefix + ".bias"] = (0.1 * rng.standard_normal(c)).astype(np.float32)
    sd[prefix + ".running_mean"] = (0.1 * rng.standard_normal(c)).astype(np.float32)
    sd[prefix + ".running_var"] = (0.5 + rng.rand(c)).astype(np.float32)
    sd[prefix + ".num_batches_tracked"] = np.array(3, dtype=np.int64)


def basic_block_state(rng, sd, prefix, inplanes, planes):
    """BasicBlock (SR/HRfuse.py:109-141) keys under `prefix`."""
    sd[prefix + ".conv1.weight"] = _conv_w(rng, planes, inplanes, 3, 1.4)
    _bn(rng, sd, prefix + ".bn1", planes)
    sd[prefix + ".conv2.weight"] = _conv_w(rng, planes, planes, 3, 1.4)
    _bn(rng, sd, prefix + ".bn2", planes)
    if inplanes != planes:
        sd[prefix + ".downsample.0.weight"] = _conv_w(rng, planes, inplanes, 1, 1.0)
        _bn(rng, sd, prefix + ".downsample.1", planes)


def hrfeature_state(in_chans=64, mid=16, out=16, seed=0, prefix=""):
    rng = np.random.RandomState(seed)
    sd = OrderedDict()
    pre = prefix + "." if prefix else ""
    basic_block_state(rng, sd, pre + "0", in_chans, mid)
    basic_block_state(rng, sd, pre + "1", mid, mid)
    basic_block_state(rng, sd, pre + "2", mid, out)
    return sd


def upsampler_state(rng, sd, prefix, n_feats=16, scale=4):
    for s in range(int(np.log2(scale))):
        sd[f"{prefix}.{2 * s}.weight"] = _conv_w(rng, 4 * n_feats, n_feats, 3, 1.0)
        sd[f"{prefix}.{2 * s}.bias"] = _bias(rng, 4 * n_feats)


def hrfuse_residual_state(hr=16, lr=16, mid=16, out=1, upscale=4, seed=0, prefix=""):
    rng = np.random.RandomState(seed)
    sd = OrderedDict()
    pre = prefix + "." if prefix else ""
    upsampler_state(rng, sd, pre + "upsampler", lr, upscale)
    basic_block_state(rng, sd, pre + "fuse.0", hr + lr, mid)
    basic_block_state(rng, sd, pre + "fuse.1", mid, mid)
    basic_block_state(rng, sd, pre + "fuse.2", mid, mid)
    sd[pre + "conv_last.weight"] = _conv_w(rng, out, mid, 3, 1.0)
    sd[pre + "conv_last.bias"] = _bias(rng, out)
    return sd


def hrfuse_plain_state(hr=16, lr=16, mid=16, out=3, upscale=4, seed=0):
    """HRfuse / HRfuse_x2 (SR/HRfuse.py:47-89): `fuse` = conv-BN-ReLU twice, `upsampler`, `conv_last`."""
    rng = np.random.RandomState(seed)
    sd = OrderedDict()
    sd["fuse.0.weight"] = _conv_w(rng, mid, hr + lr, 3, 1.4)
    _bn(rng, sd, "fuse.1", mid)
    sd["fuse.3.weight"] = _conv_w(rng, mid, mid, 3, 1.4)
    _bn(rng, sd, "fuse.4", mid)
    upsampler_state(rng, sd, "upsampler", mid, upscale)
    sd["conv_last.weight"] = _conv_w(rng, out, mid, 3, 1.0)
    sd["conv_last.bias"] = _bias(rng, out)
    return sd


def head_state(super_in=64, super_mid=16, chans_build=7, isaggre=True, seed=0):
    """The reference-owned (non-smp) parameters of SRRegress_Cls_feature (mymodels.py:259-268)."""
    sd = OrderedDict()
    sd.update(hrfuse_residual_state(super_mid, 16, 16, 1, 4, seed + 1, "reg"))
    sd.update(hrfuse_residual_state(super_mid, 16, 16, chans_build, 4, seed + 2, "seg"))
    sd.update(hrfeature_state(super_in, super_mid, super_mid, seed + 3, "hrfeat"))
    if isaggre:
        rng = np.random.RandomState(seed + 4)
        sd["aggre_height.weight"] = _conv_w(rng, 1, super_mid, 3, 1.0)
        sd["aggre_height.bias"] = _bias(rng, 1)
    return sd


SRREGRESS_SEED = 4321  # torch seed under which the a16 golden model is constructed


def perturb_smp_state(model, seed=77):
    """Make the third-party (encoder / decoder) part of an SRRegress_Cls_feature non-trivial and
    identical wherever it is built: after a `torch.manual_seed(SRREGRESS_SEED)` construction the
    BatchNorm affine parameters and running statistics of `encoder.*`, `decoder1.*`, `decoder2.*`
    are overwritten in state_dict key order from a numpy RandomState (a fresh BatchNorm is the
    identity in eval mode, which would hide wiring errors).  Used by make_golden.py on the
    REFERENCE class and by the tests on the drop-in class."""
    import torch
    rng = np.random.RandomState(seed)
    sd = model.state_dict()
    new = {}
    for k in sd:
        if k.split(".")[0] not in ("encoder", "decoder1", "decoder2"):
            continue
        v = sd[k]
        base = k.rsplit(".", 1)[0]
        if k.endswith("running_mean") and (base + ".running_var") in sd:
            c = v.numel()
            new[k] = torch.from_numpy((0.1 * rng.standard_normal(c)).astype(np.float32))
            new[base + ".running_var"] = torch.from_numpy((0.6 + 0.8 * rng.rand(c)).astype(np.float32))
            new[base + ".weight"] = torch.from_numpy((1.0 + 0.2 * rng.standard_normal(c)).astype(np.float32))
            new[base + ".bias"] = torch.from_numpy((0.1 * rng.standard_normal(c)).astype(np.float32))
    model.load_state_dict(new, strict=False)
    return model


def smp_checksum(model) -> np.ndarray:
    """[sum, abs-sum, count] over the encoder / decoder tensors: guards the seeded construction."""
    s = a = n = 0.0
    for k, v in model.state_dict().items():
        if k.split(".")[0] in ("encoder", "decoder1", "decoder2") and v.dtype.is_floating_point:
            vd = v.double()
            s += float(vd.sum()); a += float(vd.abs().sum()); n += v.numel()
    return np.array([s, a, n])


def tiles(nb, c, h=64, w=64, seed=1337) -> np.ndarray:
    """Synthetic Sentinel tiles: uniform [0,1] like the loader's clipped min-max normalisation
    (BH_loader.py:367-369)."""
    return np.random.RandomState(seed).rand(nb, c, h, w).astype(np.float32)


def features(nb, c, h, w, seed=7, scale=1.0) -> np.ndarray:
    return (np.random.RandomState(seed).standard_normal((nb, c, h, w)) * scale).astype(np.float32)


def subsample(y: np.ndarray, cs=4, ss=8) -> np.ndarray:
    """The slice of a big NCHW output that goldens store."""
    return np.ascontiguousarray(y[:, ::cs, ::ss, ::ss])


def stats(y: np.ndarray) -> np.ndarray:
    yd = y.astype(np.float64)
    return np.array([yd.sum(), np.abs(yd).sum(), (yd * yd).sum(), yd.max(), yd.min()])
