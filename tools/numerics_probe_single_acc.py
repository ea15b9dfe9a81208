#!/usr/bin/env python
"""CPU emulation of the split-precision conv numerics through the whole RRDBNet-23 trunk
(design probe for DESIGN.md §8b item 1; not part of the product path).

  scaled   (shipped): x = hi + lo'/2^11, w = hi + lo'/2^11; hi*hi -> main accumulator,
           hi*lo' + lo'*hi -> second accumulator scaled by 2^-11 in the epilogue.
  unscaled (proposed): lo stored unscaled (fp16 subnormals), weights pre-scaled by 2^8; the three
           products accumulate into ONE fp32 accumulator, the epilogue multiplies by 2^-8.
  fast     one fp16 product.
  fp8corr  (probe, round 2): fp16 main product + BOTH correction products on the fp8 tensor path (e4m3 operands, 2x the
           fp16 rate): 2 product-equivalents instead of 3 — would lift the exact path's 33 % ceiling to 50 %.
  fp8half  (probe, round 2): fp16 main + fp16 hi*lo' + fp8 lo'*hi: 2.5 product-equivalents.

Prints max |err| / (1e-4 + 1e-3 |ref|) of the [1,64,256,256] feature map against fp32 for random-init
weights and (when present) the shipped RealESRGAN_x4plus checkpoint.
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def h16(t):
    return t.to(torch.float16).to(torch.float32)


def e4m3(t):
    """Round to fp8 e4m3 with a per-tensor power-of-two scale that puts max|t| just under the format's 448."""
    amax = float(t.abs().max())
    if amax == 0.0:
        return t
    sc = 2.0 ** (8 - int(torch.ceil(torch.log2(torch.tensor(amax))).item()))     # amax * sc in (128, 256]
    return (t * sc).to(torch.float8_e4m3fn).to(torch.float32) / sc


def conv(x, w, b, mode):
    if mode == "fp32":
        return F.conv2d(x, w, b, padding=1)
    if mode == "fast":
        return F.conv2d(h16(x), h16(w), b, padding=1)
    if mode == "scaled":
        xh = h16(x); xl = h16((x - xh) * 2048.0)
        wh = h16(w); wl = h16((w - wh) * 2048.0)
        main = F.conv2d(xh, wh, None, padding=1)
        corr = F.conv2d(torch.cat((xh, xl), 1), torch.cat((wl, wh), 1), None, padding=1)
        out = main + corr / 2048.0
        return out + b.view(1, -1, 1, 1) if b is not None else out
    if mode in ("fp8corr", "fp8half"):
        xh = h16(x); xl = (x - xh) * 2048.0
        wh = h16(w); wl = (w - wh) * 2048.0
        main = F.conv2d(xh, wh, None, padding=1)
        if mode == "fp8corr":
            corr = F.conv2d(torch.cat((e4m3(xh), e4m3(xl)), 1), torch.cat((e4m3(wl), e4m3(wh)), 1), None, padding=1)
        else:
            corr = F.conv2d(xh, h16(wl), None, padding=1) + F.conv2d(e4m3(xl), e4m3(wh), None, padding=1)
        out = main + corr / 2048.0
        return out + b.view(1, -1, 1, 1) if b is not None else out
    if mode == "unscaled":
        xh = h16(x); xl = h16(x - xh)
        ws = w * 256.0
        wh = h16(ws); wl = h16(ws - wh)
        acc = F.conv2d(torch.cat((xh, xh, xl), 1), torch.cat((wh, wl, wh), 1), None, padding=1)
        out = acc / 256.0
        return out + b.view(1, -1, 1, 1) if b is not None else out
    raise ValueError(mode)


def forward_feature(x, p, mode, nb=23):
    c = lambda t, name: conv(t, p[name + ".weight"], p[name + ".bias"], mode)
    lr = lambda t: F.leaky_relu(t, 0.2)
    feat = c(x, "conv_first")
    y = feat
    for i in range(nb):
        rin = y
        for r in ("rdb1", "rdb2", "rdb3"):
            pre = f"body.{i}.{r}"
            x0 = y
            x1 = lr(c(x0, pre + ".conv1"))
            x2 = lr(c(torch.cat((x0, x1), 1), pre + ".conv2"))
            x3 = lr(c(torch.cat((x0, x1, x2), 1), pre + ".conv3"))
            x4 = lr(c(torch.cat((x0, x1, x2, x3), 1), pre + ".conv4"))
            y = c(torch.cat((x0, x1, x2, x3, x4), 1), pre + ".conv5") * 0.2 + x0
        y = y * 0.2 + rin
    feat = feat + c(y, "conv_body")
    feat = lr(c(F.interpolate(feat, scale_factor=2, mode="nearest"), "conv_up1"))
    feat = lr(c(F.interpolate(feat, scale_factor=2, mode="nearest"), "conv_up2"))
    return c(feat, "conv_hr")


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    import bhsr  # noqa: F401
    from bhsr import rrdbnet
    x = torch.rand(1, 3, 64, 64)
    nets = {}
    torch.manual_seed(1337)
    nets["random-init"] = {k: v.detach() for k, v in rrdbnet.RRDBNet(3, 3, 4, 64, 23, 32).state_dict().items()}
    ck = os.path.join(ROOT, "oracle", "_ref", "RealESRGAN_x4plus.pth")
    if os.path.exists(ck):
        nets["x4plus"] = torch.load(ck, map_location="cpu")["params_ema"]
    with torch.no_grad():
        for name, p in nets.items():
            ref = forward_feature(x, p, "fp32")
            for mode in (sys.argv[1:] or ("scaled", "unscaled", "fast")):
                got = forward_feature(x, p, mode)
                err = (got - ref).abs()
                tol = 1e-4 + 1e-3 * ref.abs()
                print(f"{name:12s} {mode:9s} max err/tol {float((err / tol).max()):8.4f}  max|err| {float(err.max()):.3e}  "
                      f"rel-L2 {float((got - ref).norm() / ref.norm()):.3e}  max|ref| {float(ref.abs().max()):.2f}", flush=True)


if __name__ == "__main__":
    main()
