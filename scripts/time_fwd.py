import os, sys, torch
sys.path.insert(0, '.')
from neurallaplace_b200 import laplace_reconstruct
from neurallaplace_b200.rep_func import LaplaceRepresentationFunc
dev = torch.device("cuda", 0)
B, T, terms, D = 4096, 1024, 33, 1
f = LaplaceRepresentationFunc(terms, D, 2).to(dev)
t = torch.linspace(20.0 / T, 20.0, T, device=dev)
flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
for mode in ("nograd", "grad"):
  p = torch.randn(B, 2, device=dev, requires_grad=(mode == "grad"))
  ts = []
  for it in range(8):
    flush.zero_()
    torch.cuda._sleep(int(4e6))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if mode == "nograd":
      with torch.no_grad():
        x = laplace_reconstruct(f, p, t, recon_dim=D)
    else:
      x = laplace_reconstruct(f, p, t, recon_dim=D)
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
  print(mode, "forward ms: min %.3f median %.3f" % (min(ts), sorted(ts)[len(ts)//2]))
