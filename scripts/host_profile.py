"""cProfile of the eager host path (laplace_reconstruct + autograd) at the reference demos' size: where the Python layer
spends its time per step.  Run on the GPU box: python scripts/host_profile.py [B T]"""
import cProfile, pstats, sys, time
import torch
sys.path.insert(0, '.')
from neurallaplace_b200 import laplace_reconstruct
from neurallaplace_b200.rep_func import LaplaceRepresentationFunc

B, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (128, 100)
dev = torch.device('cuda', 0)
f = LaplaceRepresentationFunc(33, 1, 2).to(dev)
p = torch.randn(B, 2, device=dev, requires_grad=True)
t = torch.linspace(0.1, 10.0, T, device=dev)
g = torch.randn(B, T, 1, device=dev)


def step():
  for q in f.parameters():
    q.grad = None
  p.grad = None
  x = laplace_reconstruct(f, p, t, recon_dim=1)
  x.backward(g)


for _ in range(20):
  step()
torch.cuda.synchronize()
N = 300
t0 = time.perf_counter()
for _ in range(N):
  step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host issue time per step %.1f us, with final sync %.1f us" % (1e6 * (t1 - t0) / N, 1e6 * (t2 - t0) / N))
pr = cProfile.Profile()
pr.enable()
for _ in range(N):
  step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats('cumulative').print_stats(45)
