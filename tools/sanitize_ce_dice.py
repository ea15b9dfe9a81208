#!/usr/bin/env python
"""Small ragged cases of the fused CE + Dice loss (bhsr_ce_dice) and of the stand-alone dense-block backward, meant to run
under compute-sanitizer:  compute-sanitizer --tool memcheck python tools/sanitize_ce_dice.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bhsr  # noqa: E402,F401
from bhsr import dp, rrdbnet  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
for nb, c, h, w in ((3, 7, 33, 17), (1, 2, 5, 3), (2, 16, 8, 40)):
    z = (torch.randn(nb, c, h, w, generator=g) * 3).to(dev).requires_grad_(True)
    t = torch.randint(0, c, (nb, h, w), generator=g).to(dev)
    wt = (torch.rand(nb, h, w, generator=g) + 0.1).to(dev)
    crit = dp.CE_DICE_adapt_weight(0.2, dev)
    loss = crit(z, t, wt)
    loss.backward()
    torch.cuda.synchronize()
    print(f"ce_dice {nb}x{c}x{h}x{w}: loss {float(loss):.6f} |grad| {float(z.grad.abs().sum()):.6f}")
if "--blocks" in sys.argv:
    blk = rrdbnet.ResidualDenseBlock(64, 32).to(dev).train()
    x = torch.randn(1, 64, 8, 8, generator=g).to(dev).requires_grad_(True)
    blk(x).sum().backward()
    torch.cuda.synchronize()
    print("stand-alone ResidualDenseBlock backward ok", float(x.grad.abs().sum()))
