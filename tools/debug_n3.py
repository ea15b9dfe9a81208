"""GPU debug: conv_backward (rrdbnet_train.py) vs torch autograd of F.conv2d in fp64, shape sweep."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import bhsr
from bhsr.rrdbnet_train import conv_backward

dev = torch.device("cuda:0")
def rel(a, b):
    return float((a.double() - b).norm() / b.norm().clamp_min(1e-30))
torch.manual_seed(0)
for nb, ctot, choff, cin, cout, h, w in [(2, 64, 0, 64, 64, 64, 64), (1, 64, 0, 64, 64, 64, 64), (2, 64, 0, 64, 3, 64, 64), (1, 64, 0, 64, 3, 64, 64),
                                          (2, 192, 0, 192, 64, 16, 16), (1, 192, 0, 192, 64, 16, 16), (2, 192, 0, 160, 32, 16, 16), (1, 192, 0, 160, 32, 16, 16),
                                          (2, 192, 0, 64, 32, 16, 16), (3, 192, 0, 64, 32, 16, 16), (2, 64, 0, 64, 64, 32, 32), (2, 3, 0, 3, 64, 16, 16), (1, 3, 0, 3, 64, 16, 16),
                                          (2, 192, 0, 96, 32, 24, 24), (4, 192, 0, 128, 32, 16, 16)]:
    x = torch.randn(nb, ctot, h, w, device=dev)
    wt = torch.randn(cout, cin, 3, 3, device=dev) * 0.05
    g = torch.randn(nb, cout, h, w, device=dev) * 1e-3
    xd = x.double().requires_grad_(True); wd = wt.double().requires_grad_(True)
    y = F.conv2d(xd[:, choff:choff + cin], wd, None, padding=1)
    (y * g.double()).sum().backward()
    dx = torch.zeros_like(x)
    dw, db = conv_backward(x, choff, cin, g, wt, dx, 0)
    torch.cuda.synchronize()
    print(f"nb={nb} cin={cin} cout={cout} {h}x{w}: dW rel {rel(dw, wd.grad):.2e}  db rel {rel(db, g.double().sum((0,2,3))):.2e}  dX rel {rel(dx[:, :cin], xd.grad[:, choff:choff+cin]):.2e}")
