"""GPU debug: RRDBNet training path vs fp64 oracle autograd, per-parameter errors for a few configurations."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np, torch, torch.nn.functional as F
import synth, bhsr
from bhsr import rrdbnet
from oracle import ref_torch as T

dev = torch.device("cuda:0")
def rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b), 1e-30))
for num_block, nb, hw, feature in [(0, 1, 16, True), (0, 2, 16, True), (1, 1, 16, True), (1, 2, 16, True), (1, 4, 16, True), (1, 2, 16, False)]:
    sd = synth.rrdbnet_state(num_block=num_block, seed=7 + num_block)
    net = rrdbnet.RRDBNet(3, 3, scale=4, num_block=num_block)
    net.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}, strict=True)
    net = net.to(dev).train()
    rng = np.random.RandomState(hw)
    x = rng.rand(nb, 3, hw, hw).astype(np.float32)
    wy = rng.standard_normal((nb, 64 if feature else 3, 4 * hw, 4 * hw)).astype(np.float32)
    p = {k: torch.from_numpy(np.ascontiguousarray(v)).double().requires_grad_(True) for k, v in sd.items()}
    xt = torch.from_numpy(x).double().requires_grad_(True)
    y = T._trunk(xt, p, 4)
    if not feature:
        y = T._conv(F.leaky_relu(y, 0.2), p, "conv_last")
    (y * torch.from_numpy(wy).double()).sum().backward()
    xg = torch.from_numpy(x).to(dev).requires_grad_(True)
    yg = net.forward_feature(xg) if feature else net(xg)
    (yg * torch.from_numpy(wy).to(dev)).sum().backward()
    print(f"== blocks={num_block} nb={nb} hw={hw} feature={feature}: y rel {rel(yg.detach().cpu().numpy(), y.detach().numpy()):.2e} dx rel {rel(xg.grad.cpu().numpy(), xt.grad.numpy()):.2e}")
    for name, prm in net.named_parameters():
        if p[name].grad is None or prm.grad is None:
            continue
        e = rel(prm.grad.cpu().numpy(), p[name].grad.numpy())
        if e > 1e-4 or name.endswith("conv_hr.weight") or name.endswith("conv_body.weight"):
            print(f"   {name}: {e:.2e}")
