"""CPU oracle — TEST INFRASTRUCTURE ONLY (never imported by the product path).

A numpy restatement of the reference's hot-path arithmetic, one function per reference symbol,
each citing the reference file:line it follows.  Parameters are passed as a flat
`{state_dict key: ndarray}` mapping with the reference's own key names, so a reference
checkpoint (e.g. RealESRGAN_x4plus.pth['params_ema']) drives it directly.

Pinned (tests/test_oracle_golden.py) against outputs of the reference modules themselves, run on
CPU in the build container from /root/reference by tests/golden/make_golden.py; the generated
vectors are committed under tests/golden/.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.

The smp encoder / U-Net decoders that mymodels.py pulls from segmentation_models_pytorch are a
third-party dependency absent from /root/reference (requirements.txt:15, unpinned): they are NOT
restated here — the head oracle takes the decoder feature maps as inputs (parity unpinned for
that third-party part).
"""
from __future__ import annotations

from typing import Dict, Mapping, Optional, Tuple

import numpy as np

Params = Mapping[str, np.ndarray]


# --------------------------------------------------------------------------- primitives
def conv2d(x: np.ndarray, w: np.ndarray, b: Optional[np.ndarray] = None, padding: int = 0,
           stride: int = 1, acc_dtype=np.float32) -> np.ndarray:
    """nn.Conv2d forward (cross-correlation), NCHW, square kernel, zero padding.

    im2col + one GEMM per image; accumulation in `acc_dtype` (float64 gives a tighter oracle)."""
    n, c, h, wd = x.shape
    o, ci, kh, kw = w.shape
    assert ci == c
    xp = np.pad(x, ((0, 0), (0, 0), (padding, padding), (padding, padding))) if padding else x
    oh = (h + 2 * padding - kh) // stride + 1
    ow = (wd + 2 * padding - kw) // stride + 1
    wm = w.reshape(o, -1).astype(acc_dtype)
    out = np.empty((n, o, oh, ow), dtype=np.float32)
    for i in range(n):
        cols = np.empty((c, kh, kw, oh, ow), dtype=acc_dtype)
        for ky in range(kh):
            for kx in range(kw):
                cols[:, ky, kx] = xp[i, :, ky:ky + stride * oh:stride, kx:kx + stride * ow:stride]
        y = wm @ cols.reshape(c * kh * kw, oh * ow)
        if b is not None:
            y += b.astype(acc_dtype)[:, None]
        out[i] = y.reshape(o, oh, ow).astype(np.float32)
    return out


def leaky_relu(x: np.ndarray, slope: float = 0.2) -> np.ndarray:
    return np.where(x > 0, x, np.float32(slope) * x).astype(np.float32)


def relu(x: np.ndarray) -> np.ndarray:
    return np.maximum(x, 0).astype(np.float32)


def nearest_up2(x: np.ndarray) -> np.ndarray:
    """F.interpolate(scale_factor=2, mode='nearest'): dst index d reads src d // 2
    (SR/rrdbnet_arch.py:219-220, 236-237)."""
    return x.repeat(2, axis=2).repeat(2, axis=3)


def pixel_shuffle(x: np.ndarray, r: int) -> np.ndarray:
    """nn.PixelShuffle(r): out[b, c, h*r+i, w*r+j] = in[b, c*r*r + i*r + j, h, w]
    (SR/HRfuse.py:24, 34)."""
    b, c, h, w = x.shape
    oc = c // (r * r)
    return x.reshape(b, oc, r, r, h, w).transpose(0, 1, 4, 2, 5, 3).reshape(b, oc, h * r, w * r)


def pixel_unshuffle(x: np.ndarray, scale: int) -> np.ndarray:
    """SR/rrdbnet_arch.py:94-110."""
    b, c, hh, hw = x.shape
    assert hh % scale == 0 and hw % scale == 0
    h, w = hh // scale, hw // scale
    return x.reshape(b, c, h, scale, w, scale).transpose(0, 1, 3, 5, 2, 4).reshape(
        b, c * scale * scale, h, w)


def batch_norm(x: np.ndarray, p: Params, prefix: str, training: bool, eps: float = 1e-5,
               momentum: float = 0.1, new_stats: Optional[Dict[str, np.ndarray]] = None) -> np.ndarray:
    """nn.BatchNorm2d forward (SR/HRfuse.py:129, 132, 138).  Training mode normalises with the
    biased batch variance and, if `new_stats` is given, records the running-stat update (unbiased
    variance, momentum 0.1) under the reference's buffer names."""
    g = p[prefix + ".weight"].astype(np.float32)
    bt = p[prefix + ".bias"].astype(np.float32)
    if training:
        xd = x.astype(np.float64)
        mean = xd.mean(axis=(0, 2, 3))
        var = xd.var(axis=(0, 2, 3))
        if new_stats is not None:
            cnt = x.shape[0] * x.shape[2] * x.shape[3]
            unbiased = var * cnt / max(cnt - 1, 1)
            new_stats[prefix + ".running_mean"] = (
                (1 - momentum) * p[prefix + ".running_mean"] + momentum * mean).astype(np.float32)
            new_stats[prefix + ".running_var"] = (
                (1 - momentum) * p[prefix + ".running_var"] + momentum * unbiased).astype(np.float32)
            new_stats[prefix + ".num_batches_tracked"] = p[prefix + ".num_batches_tracked"] + 1
    else:
        mean = p[prefix + ".running_mean"].astype(np.float64)
        var = p[prefix + ".running_var"].astype(np.float64)
    inv = 1.0 / np.sqrt(var + eps)
    y = (x.astype(np.float64) - mean[None, :, None, None]) * inv[None, :, None, None]
    y = y * g[None, :, None, None] + bt[None, :, None, None]
    return y.astype(np.float32)


# --------------------------------------------------------------------------- RRDBNet
def residual_dense_block(x: np.ndarray, p: Params, prefix: str, **kw) -> np.ndarray:
    """ResidualDenseBlock.forward, SR/rrdbnet_arch.py:136-143."""
    def c(i, inp):
        return conv2d(inp, p[f"{prefix}.conv{i}.weight"], p[f"{prefix}.conv{i}.bias"], padding=1, **kw)
    x1 = leaky_relu(c(1, x))
    x2 = leaky_relu(c(2, np.concatenate((x, x1), 1)))
    x3 = leaky_relu(c(3, np.concatenate((x, x1, x2), 1)))
    x4 = leaky_relu(c(4, np.concatenate((x, x1, x2, x3), 1)))
    x5 = c(5, np.concatenate((x, x1, x2, x3, x4), 1))
    return (x5 * np.float32(0.2) + x).astype(np.float32)


def rrdb(x: np.ndarray, p: Params, prefix: str, **kw) -> np.ndarray:
    """RRDB.forward, SR/rrdbnet_arch.py:162-167."""
    out = residual_dense_block(x, p, prefix + ".rdb1", **kw)
    out = residual_dense_block(out, p, prefix + ".rdb2", **kw)
    out = residual_dense_block(out, p, prefix + ".rdb3", **kw)
    return (out * np.float32(0.2) + x).astype(np.float32)


def _num_blocks(p: Params, body: str = "body") -> int:
    idx = {int(k.split(".")[1]) for k in p if k.startswith(body + ".")}
    return max(idx) + 1 if idx else 0


def _rrdbnet_trunk(x: np.ndarray, p: Params, scale: int, **kw) -> np.ndarray:
    if scale == 2:
        feat = pixel_unshuffle(x, 2)
    elif scale == 1:
        feat = pixel_unshuffle(x, 4)
    else:
        feat = x
    feat = conv2d(feat, p["conv_first.weight"], p["conv_first.bias"], padding=1, **kw)
    body = feat
    for i in range(_num_blocks(p)):
        body = rrdb(body, p, f"body.{i}", **kw)
    body = conv2d(body, p["conv_body.weight"], p["conv_body.bias"], padding=1, **kw)
    feat = feat + body
    feat = leaky_relu(conv2d(nearest_up2(feat), p["conv_up1.weight"], p["conv_up1.bias"], padding=1, **kw))
    feat = leaky_relu(conv2d(nearest_up2(feat), p["conv_up2.weight"], p["conv_up2.bias"], padding=1, **kw))
    return conv2d(feat, p["conv_hr.weight"], p["conv_hr.bias"], padding=1, **kw)


def rrdbnet_forward_feature(x: np.ndarray, p: Params, scale: int = 4, **kw) -> np.ndarray:
    """RRDBNet.forward_feature, SR/rrdbnet_arch.py:225-240: conv_hr output WITHOUT activation."""
    return _rrdbnet_trunk(x, p, scale, **kw)


def rrdbnet_forward(x: np.ndarray, p: Params, scale: int = 4, **kw) -> np.ndarray:
    """RRDBNet.forward, SR/rrdbnet_arch.py:208-223."""
    feat = leaky_relu(_rrdbnet_trunk(x, p, scale, **kw))
    return conv2d(feat, p["conv_last.weight"], p["conv_last.bias"], padding=1, **kw)


# The older ESRGAN-style class (SR/RRDBNet.py:53-78) is the same arithmetic under other names.
_OLD_TO_NEW = (("RRDB_trunk.", "body."), (".RDB1.", ".rdb1."), (".RDB2.", ".rdb2."),
               (".RDB3.", ".rdb3."), ("trunk_conv.", "conv_body."), ("upconv1.", "conv_up1."),
               ("upconv2.", "conv_up2."), ("HRconv.", "conv_hr."))


def rename_old_rrdbnet_keys(p: Params) -> Dict[str, np.ndarray]:
    out = {}
    for k, v in p.items():
        for a, b in _OLD_TO_NEW:
            k = k.replace(a, b)
        out[k] = v
    return out


def old_rrdbnet_forward(x: np.ndarray, p: Params, **kw) -> np.ndarray:
    """SR/RRDBNet.py:69-78 (== rrdbnet_forward at scale 4)."""
    return rrdbnet_forward(x, rename_old_rrdbnet_keys(p), scale=4, **kw)


# --------------------------------------------------------------------------- HR fusion head
def upsampler(x: np.ndarray, p: Params, prefix: str, scale: int = 4, **kw) -> np.ndarray:
    """Upsampler, SR/HRfuse.py:17-44: (conv3x3 n->4n + PixelShuffle(2)) x log2(scale); the
    Sequential indices of the convs are 0, 2, ..."""
    assert scale & (scale - 1) == 0
    steps = int(np.log2(scale))
    for s in range(steps):
        x = conv2d(x, p[f"{prefix}.{2 * s}.weight"], p.get(f"{prefix}.{2 * s}.bias"), padding=1, **kw)
        x = pixel_shuffle(x, 2)
    return x


def basic_block(x: np.ndarray, p: Params, prefix: str, training: bool = False,
                new_stats: Optional[Dict[str, np.ndarray]] = None, **kw) -> np.ndarray:
    """BasicBlock.forward, SR/HRfuse.py:143-159 (stride 1)."""
    out = conv2d(x, p[prefix + ".conv1.weight"], None, padding=1, **kw)
    out = relu(batch_norm(out, p, prefix + ".bn1", training, new_stats=new_stats))
    out = conv2d(out, p[prefix + ".conv2.weight"], None, padding=1, **kw)
    out = batch_norm(out, p, prefix + ".bn2", training, new_stats=new_stats)
    if prefix + ".downsample.0.weight" in p:
        identity = conv2d(x, p[prefix + ".downsample.0.weight"], None, padding=0, **kw)
        identity = batch_norm(identity, p, prefix + ".downsample.1", training, new_stats=new_stats)
    else:
        identity = x
    return relu(out + identity)


def hrfeature(x: np.ndarray, p: Params, prefix: str = "", training: bool = False,
              new_stats=None, **kw) -> np.ndarray:
    """HRfeature, SR/HRfuse.py:164-169: three BasicBlocks (Sequential keys 0,1,2)."""
    pre = prefix + "." if prefix else ""
    for i in range(3):
        x = basic_block(x, p, f"{pre}{i}", training, new_stats, **kw)
    return x


def hrfuse_residual(x_lr: np.ndarray, x_hr: np.ndarray, p: Params, prefix: str = "",
                    upscale: int = 4, training: bool = False, new_stats=None, **kw) -> np.ndarray:
    """HRfuse_residual.forward, SR/HRfuse.py:185-190: upsample LR, cat([LR, HR]) (LR first),
    three BasicBlocks, 3x3 conv_last with bias."""
    pre = prefix + "." if prefix else ""
    x_lr = upsampler(x_lr, p, pre + "upsampler", upscale, **kw)
    x = np.concatenate([x_lr, x_hr], axis=1)
    for i in range(3):
        x = basic_block(x, p, f"{pre}fuse.{i}", training, new_stats, **kw)
    return conv2d(x, p[pre + "conv_last.weight"], p[pre + "conv_last.bias"], padding=1, **kw)


def srregress_head(height_fea: np.ndarray, build_fea: np.ndarray, super_fea: np.ndarray, p: Params,
                   isaggre: bool, upscale: int = 4, training: bool = False, new_stats=None, **kw):
    """The part of SRRegress_Cls_feature.forward (mymodels.py:270-293) that the reference owns:
    hrfeat, aggre_height, reg, seg — given the two 16-channel U-Net decoder outputs
    (third-party smp, not restated)."""
    sf = hrfeature(super_fea, p, "hrfeat", training, new_stats, **kw)
    out = []
    height = hrfuse_residual(height_fea, sf, p, "reg", upscale, training, new_stats, **kw)
    build = hrfuse_residual(build_fea, sf, p, "seg", upscale, training, new_stats, **kw)
    out = [height, build]
    if isaggre:
        out.append(conv2d(height_fea, p["aggre_height.weight"], p["aggre_height.bias"], padding=1, **kw))
    return tuple(out)


# --------------------------------------------------------------------------- aggregation
def aggregate_torch(data: np.ndarray, scale: float) -> np.ndarray:
    """aggregate_torch, aggregate_utils.py:29-41: ones-conv(k=s=step) sum / count(data >= 0),
    then squeeze."""
    step = int(1 / scale)
    ones = np.ones((1, 1, step, step), dtype=np.float32)
    s1 = conv2d(data.astype(np.float32), ones, None, 0, stride=step)
    s2 = conv2d((data >= 0).astype(np.float32), ones, None, 0, stride=step)
    return np.squeeze(s1 / (s2 + np.float32(1e-10)))


def aggregate_torch_gpu(data: np.ndarray, scale: float) -> np.ndarray:
    """aggregate_torch_gpu, aggregate_utils.py:44-59: mask is data > 1.0, no squeeze."""
    step = int(1 / scale)
    ones = np.ones((1, 1, step, step), dtype=np.float32)
    s1 = conv2d(data.astype(np.float32), ones, None, 0, stride=step)
    s2 = conv2d((data > 1.0).astype(np.float32), ones, None, 0, stride=step)
    return s1 / (s2 + np.float32(1e-10))


def aggregate(data: np.ndarray, scale: float) -> np.ndarray:
    """aggregate, aggregate_utils.py:11-26: loop version, mean over pixels > 0; keeps the
    reference's `nc = int(r*scale)` (rows used for both output dims)."""
    r, c = data.shape
    nr, nc = int(r * scale), int(r * scale)
    step = int(1 / scale)
    res = np.zeros((nr, nc))
    data = data.astype("float")
    for i in range(0, r, step):
        for j in range(0, c, step):
            patch = data[i:i + step, j:j + step]
            res[int(i / step), int(j / step)] = patch.sum() / ((patch > 0).sum() + 1e-6)
    return res


# --------------------------------------------------------------------------- loss-side KAT
def hierweight(stats: np.ndarray, hir: Tuple[int, ...]) -> np.ndarray:
    """hierweight, BH_loader.py:30-41: inverse-sqrt-frequency class weights over height levels."""
    stats = np.asarray(stats, dtype=np.float64)
    num_hier = len(hir) - 1
    stats = stats / stats.sum()
    pre = np.zeros((num_hier,))
    for i in range(num_hier):
        pre[i] = stats[hir[i]:hir[i + 1]].sum()
    pre = 1 / np.sqrt(pre)
    pre /= pre.sum()
    scaling = num_hier / np.sum(pre)
    return scaling * pre


def weighted_mse_adapt(pred: np.ndarray, target: np.ndarray, weight: np.ndarray, log_var: float) -> float:
    """MSE_adapt_weight.forward, losses_pytorch/selfloss.py:81-90:
    mean(weight * (pred-target)^2) * exp(-log_var) + log_var."""
    loss = np.mean(weight.astype(np.float64) * (pred.astype(np.float64) - target.astype(np.float64)) ** 2)
    return float(loss * np.exp(-log_var) + log_var)


# --------------------------------------------------------------------------- predictor post-processing / loss
def predict_postprocess(ypred: np.ndarray, build_pred: np.ndarray):
    """predict_realesanet_feature_globe.py:172-177: `ypred[ypred<0] = 0; round(ypred*10).astype(uint16)` and
    `round(softmax(build_pred, dim=1) * 255).astype(uint16)` (numpy rounding: half to even), float32 arithmetic."""
    y = ypred.astype(np.float32).copy()
    y[y < 0] = 0
    y = np.round(y * np.float32(10)).astype(np.uint16)
    b = build_pred.astype(np.float32)
    e = np.exp(b - b.max(axis=1, keepdims=True), dtype=np.float32)
    p = e / e.sum(axis=1, keepdims=True, dtype=np.float32)
    return y, np.round(p * np.float32(255)).astype(np.uint16)


def mse_adapt_weight(pred: np.ndarray, target: np.ndarray, weight: np.ndarray, log_var: float):
    """losses_pytorch/selfloss.py:81-90 with its gradients (float64): loss, d loss/d pred, d loss/d log_var."""
    d = pred.astype(np.float64) - target.astype(np.float64)
    w = weight.astype(np.float64)
    mean = (w * d * d).mean()
    prec = np.exp(-float(log_var))
    return mean * prec + float(log_var), 2.0 * w * d * prec / d.size, 1.0 - mean * prec


def ce_dice_adapt_weight(logits: np.ndarray, labels: np.ndarray, weight: np.ndarray, log_var: float):
    """CE_DICE_adapt_weight.forward (losses_pytorch/selfloss.py:145-168, Dice :6-17) with its gradients (float64):
    loss = (mean(weight * CE(logits, labels)) + Dice(p, labels > 0)) * exp(-log_var) + log_var with
    p = softmax(logits)[:, 1:].sum(1) and Dice(m1, m2) = 1 - (2 sum(m1 m2) + 1) / (sum(m1) + sum(m2) + 1).
    logits [N,C,H,W], labels [N,H,W] int in [0, C), weight [N,H,W].  Returns loss, d loss/d logits, d loss/d log_var."""
    z = logits.astype(np.float64)
    t = labels.astype(np.int64)
    w = weight.astype(np.float64)
    c = z.shape[1]
    e = np.exp(z - z.max(axis=1, keepdims=True))
    sm = e / e.sum(axis=1, keepdims=True)
    onehot = np.moveaxis(np.eye(c)[t], -1, 1)
    ce = -np.log((sm * onehot).sum(axis=1))
    loss_ce = (ce * w).mean()
    p = 1.0 - sm[:, 0]
    m2 = (t > 0).astype(np.float64)
    inter, den = (p * m2).sum(), p.sum() + m2.sum() + 1.0
    dice = 1.0 - (2.0 * inter + 1.0) / den
    prec = np.exp(-float(log_var))
    loss = (loss_ce + dice) * prec + float(log_var)
    # d dice / d p_i = (2 inter + 1) / den^2 - 2 m2_i / den;  d p / d z_c = softmax_0 (softmax_c - [c == 0])
    ddice_dp = (2.0 * inter + 1.0) / den ** 2 - 2.0 * m2 / den
    first = np.zeros((1, c, 1, 1))
    first[0, 0] = 1.0
    grad = prec * ((w / ce.size)[:, None] * (sm - onehot) + (ddice_dp * sm[:, 0])[:, None] * (sm - first))
    return loss, grad, 1.0 - (loss_ce + dice) * prec

