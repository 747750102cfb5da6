"""x(t) and gradient errors of the float32 tensor-core path against the float64 kernels for several weight draws."""
import os, sys, torch
sys.path.insert(0, '.')
from neurallaplace_b200 import laplace_reconstruct
from neurallaplace_b200.rep_func import LaplaceRepresentationFunc
from tests import helpers as hp
dev = torch.device('cuda', 0)
def run(f, p, t, g):
    for q in f.parameters(): q.grad = None
    pr = p.clone().requires_grad_(True)
    x = laplace_reconstruct(f, pr, t, recon_dim=1); x.backward(g)
    return x.detach(), [q.grad.clone() for q in f.parameters()], pr.grad.clone()
B, T = 4096, 128
worst = [0.0] * 8
for seed in range(8):
    torch.manual_seed(seed)
    f = LaplaceRepresentationFunc(33, 1, 2).to(dev)
    f64 = LaplaceRepresentationFunc(33, 1, 2).to(dev).double(); f64.load_state_dict({k: v.double() for k, v in f.state_dict().items()})
    p = torch.randn(B, 2, device=dev); t = torch.linspace(20.0 / T, 20.0, T, device=dev); g = torch.randn(B, T, 1, device=dev).abs()
    x, gw, gp = run(f, p, t, g); x64, gw64, gp64 = run(f64, p.double(), t.double(), g.double())
    e = [hp.rel(x, x64)] + [hp.rel(a, b) for a, b in zip(gw, gw64)] + [hp.rel(gp, gp64)]
    worst = [max(a, b) for a, b in zip(worst, e)]
    print(seed, " ".join("%.1e" % v for v in e))
print("worst", " ".join("%.1e" % v for v in worst), "(x W1 b1 W2 b2 W3 b3 dp)")
