"""CPU oracle, torch-functional flavour — TEST / BASELINE INFRASTRUCTURE ONLY.

The same restatement as oracle/ref_numpy.py, but every primitive is the stock
torch.nn.functional call the reference's nn.Modules dispatch to on CPU (oneDNN conv, native
batch_norm, pixel_shuffle, nearest interpolate).  It exists so bench.py's `cpu_baseline` and
`--impl reference` legs time what the reference itself would execute on the GPU box's host cores
(where /root/reference does not exist), rather than a slower numpy port.  Pinned against the
reference-generated goldens in tests/test_oracle_golden.py.  Never imported by the product path.

Parameters: `{state_dict key: torch.Tensor (cpu, fp32)}` with the reference's key names.
"""
from __future__ import annotations

from typing import Mapping

import torch
import torch.nn.functional as F

Params = Mapping[str, torch.Tensor]


def _conv(x, p, name, padding=1):
    return F.conv2d(x, p[name + ".weight"], p.get(name + ".bias"), stride=1, padding=padding)


def residual_dense_block(x, p: Params, prefix: str):
    """SR/rrdbnet_arch.py:136-143."""
    x1 = F.leaky_relu(_conv(x, p, prefix + ".conv1"), 0.2)
    x2 = F.leaky_relu(_conv(torch.cat((x, x1), 1), p, prefix + ".conv2"), 0.2)
    x3 = F.leaky_relu(_conv(torch.cat((x, x1, x2), 1), p, prefix + ".conv3"), 0.2)
    x4 = F.leaky_relu(_conv(torch.cat((x, x1, x2, x3), 1), p, prefix + ".conv4"), 0.2)
    x5 = _conv(torch.cat((x, x1, x2, x3, x4), 1), p, prefix + ".conv5")
    return x5 * 0.2 + x


def rrdb(x, p: Params, prefix: str):
    """SR/rrdbnet_arch.py:162-167."""
    out = residual_dense_block(x, p, prefix + ".rdb1")
    out = residual_dense_block(out, p, prefix + ".rdb2")
    out = residual_dense_block(out, p, prefix + ".rdb3")
    return out * 0.2 + x


def pixel_unshuffle(x, scale):
    """SR/rrdbnet_arch.py:94-110."""
    b, c, hh, hw = x.size()
    h, w = hh // scale, hw // scale
    return x.view(b, c, h, scale, w, scale).permute(0, 1, 3, 5, 2, 4).reshape(b, c * scale * scale, h, w)


def _trunk(x, p: Params, scale: int):
    if scale == 2:
        x = pixel_unshuffle(x, 2)
    elif scale == 1:
        x = pixel_unshuffle(x, 4)
    feat = _conv(x, p, "conv_first")
    nblk = 1 + max([int(k.split(".")[1]) for k in p if k.startswith("body.")], default=-1)
    body = feat
    for i in range(nblk):
        body = rrdb(body, p, f"body.{i}")
    feat = feat + _conv(body, p, "conv_body")
    feat = F.leaky_relu(_conv(F.interpolate(feat, scale_factor=2, mode="nearest"), p, "conv_up1"), 0.2)
    feat = F.leaky_relu(_conv(F.interpolate(feat, scale_factor=2, mode="nearest"), p, "conv_up2"), 0.2)
    return _conv(feat, p, "conv_hr")


@torch.no_grad()
def rrdbnet_forward_feature(x, p: Params, scale: int = 4):
    """SR/rrdbnet_arch.py:225-240."""
    return _trunk(x, p, scale)


@torch.no_grad()
def rrdbnet_forward(x, p: Params, scale: int = 4):
    """SR/rrdbnet_arch.py:208-223."""
    return _conv(F.leaky_relu(_trunk(x, p, scale), 0.2), p, "conv_last")


def _bn(x, p: Params, prefix: str, training: bool):
    # functional batch_norm updates the running buffers in place when training, like nn.BatchNorm2d
    return F.batch_norm(x, p[prefix + ".running_mean"], p[prefix + ".running_var"], p[prefix + ".weight"],
                        p[prefix + ".bias"], training=training, momentum=0.1, eps=1e-5)


def basic_block(x, p: Params, prefix: str, training: bool = False):
    """SR/HRfuse.py:143-159."""
    out = F.relu(_bn(_conv(x, p, prefix + ".conv1"), p, prefix + ".bn1", training))
    out = _bn(_conv(out, p, prefix + ".conv2"), p, prefix + ".bn2", training)
    if prefix + ".downsample.0.weight" in p:
        identity = _bn(_conv(x, p, prefix + ".downsample.0", padding=0), p, prefix + ".downsample.1", training)
    else:
        identity = x
    return F.relu(out + identity)


def hrfeature(x, p: Params, prefix: str = "", training: bool = False):
    """SR/HRfuse.py:164-169."""
    pre = prefix + "." if prefix else ""
    for i in range(3):
        x = basic_block(x, p, f"{pre}{i}", training)
    return x


def hrfuse_residual(x_lr, x_hr, p: Params, prefix: str = "", upscale: int = 4, training: bool = False):
    """SR/HRfuse.py:185-190."""
    pre = prefix + "." if prefix else ""
    s = 0
    while (1 << s) < upscale:
        x_lr = F.pixel_shuffle(_conv(x_lr, p, f"{pre}upsampler.{2 * s}"), 2)
        s += 1
    x = torch.cat([x_lr, x_hr], dim=1)
    for i in range(3):
        x = basic_block(x, p, f"{pre}fuse.{i}", training)
    return _conv(x, p, pre + "conv_last")


def srregress_head(height_fea, build_fea, super_fea, p: Params, isaggre: bool, training: bool = False):
    """Reference-owned part of SRRegress_Cls_feature.forward (mymodels.py:270-293)."""
    sf = hrfeature(super_fea, p, "hrfeat", training)
    out = []
    if isaggre:
        aggre = _conv(height_fea, p, "aggre_height")
    height = hrfuse_residual(height_fea, sf, p, "reg", 4, training)
    build = hrfuse_residual(build_fea, sf, p, "seg", 4, training)
    out = [height, build]
    if isaggre:
        out.append(aggre)
    return tuple(out)
