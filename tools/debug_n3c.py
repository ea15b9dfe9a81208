"""GPU debug: record every conv_backward call of the RRDBNet training path and compare the tail stages with the oracle."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np, torch, torch.nn.functional as F
import synth, bhsr
from bhsr import rrdbnet, rrdbnet_train
from oracle import ref_torch as T

dev = torch.device("cuda:0")
def rel(a, b):
    a = a.detach().cpu().double(); b = b.detach().cpu().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
rec = []
orig = rrdbnet_train.conv_backward
def spy(x_f32, x_choff, cin, g, weight, dx, dx_choff=0, need_dw=True):
    before = dx.clone() if dx is not None else None
    out = orig(x_f32, x_choff, cin, g, weight, dx, dx_choff, need_dw)
    rec.append(dict(x=x_f32[:, x_choff:x_choff + cin].clone(), g=g.clone(), dx=(dx[:, dx_choff:dx_choff + cin] - before[:, dx_choff:dx_choff + cin]).clone() if dx is not None else None, dw=out[0]))
    return out
rrdbnet_train.conv_backward = spy
for nb in (1, 2):
    rec.clear()
    num_block, hw = 1, 16
    sd = synth.rrdbnet_state(num_block=num_block, seed=8)
    net = rrdbnet.RRDBNet(3, 3, scale=4, num_block=num_block)
    net.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}, strict=True)
    net = net.to(dev).train()
    rng = np.random.RandomState(hw)
    x = rng.rand(nb, 3, hw, hw).astype(np.float32)
    wy = rng.standard_normal((nb, 64, 4 * hw, 4 * hw)).astype(np.float32)
    p = {k: torch.from_numpy(np.ascontiguousarray(v)).double().requires_grad_(True) for k, v in sd.items()}
    xt = torch.from_numpy(x).double()
    feat = T._conv(xt, p, "conv_first"); feat.retain_grad()
    body = T.rrdb(feat, p, "body.0"); body.retain_grad()
    cb = T._conv(body, p, "conv_body"); cb.retain_grad()
    feat2 = feat + cb; feat2.retain_grad()
    z1 = T._conv(F.interpolate(feat2, scale_factor=2, mode="nearest"), p, "conv_up1"); z1.retain_grad()
    u1 = F.leaky_relu(z1, 0.2); u1.retain_grad()
    z2 = T._conv(F.interpolate(u1, scale_factor=2, mode="nearest"), p, "conv_up2"); z2.retain_grad()
    u2 = F.leaky_relu(z2, 0.2); u2.retain_grad()
    y = T._conv(u2, p, "conv_hr")
    (y * torch.from_numpy(wy).double()).sum().backward()
    xg = torch.from_numpy(x).to(dev).requires_grad_(True)
    yg = net.forward_feature(xg)
    (yg * torch.from_numpy(wy).to(dev)).sum().backward()
    print(f"== nb={nb}: {len(rec)} conv_backward calls")
    names = ["conv_hr", "conv_up2", "conv_up1", "conv_body"]
    ox = [u2, F.interpolate(u1, scale_factor=2, mode="nearest"), F.interpolate(feat2, scale_factor=2, mode="nearest"), body]
    og = [torch.from_numpy(wy).double(), z2.grad, z1.grad, cb.grad]
    for i, nme in enumerate(names):
        r = rec[i]
        print(f"  {nme}: x rel {rel(r['x'], ox[i]):.2e}  g rel {rel(r['g'], og[i]):.2e}  dW rel {rel(r['dw'], p[nme + '.weight'].grad):.2e}")
    print(f"  d_body (call 3 dx) vs oracle body.grad: {rel(rec[3]['dx'], body.grad):.2e}; u1.grad check via call1 dx 2x2 sum: "
          f"{rel(rec[1]['dx'].view(nb, 64, 32, 2, 32, 2).sum((3, 5)), u1.grad):.2e}; feat2.grad via call2: {rel(rec[2]['dx'].view(nb, 64, 16, 2, 16, 2).sum((3, 5)), feat2.grad):.2e}")
    d_up1 = rec[1]['dx'].view(nb, 64, 32, 2, 32, 2).sum((3, 5)).cpu().double()
    ratio = rec[2]['g'].cpu().double() / d_up1
    mine_pos = ratio > 0.6
    ref_pos = z1.detach() > 0
    diff = (mine_pos != ref_pos)
    idx = diff.nonzero()
    print(f"  mask mismatches: {int(diff.sum())} of {diff.numel()}")
    for i in idx[:10]:
        n_, c_, y_, x_ = [int(v) for v in i]
        print(f"    at {n_, c_, y_, x_}: oracle z1 = {float(z1[n_, c_, y_, x_]):.3e}, ratio {float(ratio[n_, c_, y_, x_]):.3f}")
