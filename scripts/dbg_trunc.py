"""float32 tensor-core gradients against the float64 SIMT kernels at sizes with many tiles per CTA."""
import os, sys, torch
sys.path.insert(0, '.')
from neurallaplace_b200 import laplace_reconstruct
from neurallaplace_b200.rep_func import LaplaceRepresentationFunc
from tests import helpers as hp
dev = torch.device('cuda', 0)
torch.manual_seed(1)
names = ["W1", "b1", "W2", "b2", "W3", "b3"]
def run(f, p, t, g):
    params = list(f.parameters())
    for q in params: q.grad = None
    pr = p.clone().requires_grad_(True)
    x = laplace_reconstruct(f, pr, t, recon_dim=1)
    x.backward(g)
    return x.detach(), [q.grad.clone() for q in params], pr.grad.clone()
for (B, T, sign) in [(4096, 1024, "pos"), (4096, 1024, "rand"), (4096, 128, "pos"), (70000, 40, "pos")]:
    f = LaplaceRepresentationFunc(33, 1, 2).to(dev)
    p = torch.randn(B, 2, device=dev); t = torch.linspace(20.0 / T, 20.0, T, device=dev); g = torch.randn(B, T, 1, device=dev)
    if sign == "pos": g = g.abs()
    f64 = LaplaceRepresentationFunc(33, 1, 2).to(dev).double(); f64.load_state_dict({k: v.double() for k, v in f.state_dict().items()})
    x64, g64, p64 = run(f64, p.double(), t.double(), g.double())
    for label, env in (("recompute", {"NL_B200_STASH": "0"}), ("saved-1stream", {"NL_B200_STASH": "1", "NL_B200_BS2": "0"}), ("saved-2stream", {"NL_B200_STASH": "1", "NL_B200_BS2": "1"})):
        os.environ.update(env)
        x, gw, gp = run(f, p, t, g)
        print(B, T, sign, "%-14s x %.1e " % (label, hp.rel(x, x64)) + " ".join("%s %.1e" % (n, hp.rel(a, b)) for n, a, b in zip(names, gw, g64)) + " dp %.1e" % hp.rel(gp, p64))
