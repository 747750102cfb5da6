# times the split-path De Hoog line integral (QD recurrence, thread per point) at the headline size
import torch, time
from neurallaplace_b200 import _ops, _lib
from neurallaplace_b200.inverse_laplace import DeHoog
dev = torch.device("cuda:0")
B, T, n = 4096, 1024, 33
ilt = DeHoog(ilt_reconstruction_terms=n)
consts = ilt._c
t = torch.linspace(0.1, 10.0, T, device=dev)
g = torch.Generator(device=dev); g.manual_seed(0)
F = (torch.randn(B, T, 1, n, 2, device=dev, generator=g) * 0.5).requires_grad_(True)
def run(bwd):
  x = _ops.line_integrate(consts, F, None, t, None, _lib.NL_F_COMPLEX_INTERLEAVED, 0)
  if bwd:
    x.backward(torch.ones_like(x))
for bwd in (False, True):
  for _ in range(2): run(bwd)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(3): run(bwd)
  e1.record(); torch.cuda.synchronize()
  print("bwd" if bwd else "fwd", e0.elapsed_time(e1) / 3, "ms")
