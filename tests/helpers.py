"""Shared test helpers (oracle access lives here: tests may use oracle/, the product may not)."""
import glob
import os

import numpy as np
import torch

from oracle import nl_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden_cases():
  return sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, "recon_*.npz")))


def load_case(name):
  z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
  dt = torch.float64 if "float64" in str(z["dtype"]) else torch.float32
  case = dict(name=name, alg=str(z["alg"]), terms=int(z["terms"]), D=int(z["D"]), K=int(z["K"]), H=int(z["H"]),
              sphere=bool(int(z["sphere"])), t_mode=str(z["t_mode"]), dtype=dt)
  for k in ("p", "t", "gout", "x"):
    case[k] = torch.from_numpy(z[k])
  case["weights"] = [torch.from_numpy(z[f"w{i}"]) for i in range(6)]
  case["grads"] = [torch.from_numpy(z[f"g{i}"]) for i in range(7)]
  opts = {k[4:]: float(z[k]) for k in z.files if k.startswith("opt_")}
  case["options"] = opts or None
  return case


def rel(a, b):
  a, b = a.detach().double().cpu(), b.detach().double().cpu()
  return float((a - b).norm() / b.norm().clamp_min(1e-300))


def oracle_run(weights, p, t, D, alg, sphere, terms, options, gout):
  """Oracle fwd + bwd on CPU; returns x and the 7 gradients (W1,b1,W2,b2,W3,b3,p)."""
  ws = [w.detach().cpu().clone().requires_grad_(True) for w in weights]
  po = p.detach().cpu().clone().requires_grad_(True)
  x = orc.laplace_reconstruct_oracle(ws, po, t.detach().cpu(), D, alg, sphere, terms, options)
  (x * gout.detach().cpu()).sum().backward()
  return x.detach(), [w.grad for w in ws] + [po.grad]


def make_module(weights, n, D, K, sphere=True, device="cpu", user_copy=False):
  """Canonical laplace_rep_func with the given weights.  ``user_copy`` builds a look-alike class
  that does NOT declare _nl_head (exercises the probe in core._head_kind)."""
  from neurallaplace_b200.rep_func import LaplaceRepresentationFunc, SurfaceModel
  H = weights[0].shape[0]
  cls = LaplaceRepresentationFunc if sphere else SurfaceModel
  if user_copy:
    cls = type("UserCopy", (cls,), {"_nl_head": None})
  m = cls(n, D, K, hidden_units=H)
  lin = [m.linear_tanh_stack[i] for i in (0, 2, 4)]
  with torch.no_grad():
    for i, l in enumerate(lin):
      l.weight = torch.nn.Parameter(weights[2 * i].clone())
      l.bias = torch.nn.Parameter(weights[2 * i + 1].clone())
  return m.to(device)


def module_params(m):
  l = [m.linear_tanh_stack[i] for i in (0, 2, 4)]
  return [l[0].weight, l[0].bias, l[1].weight, l[1].bias, l[2].weight, l[2].bias]
