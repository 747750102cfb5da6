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


# ---- De Hoog in float32: robust comparison protocol ---------------------------------------------------------------
# The float32 quotient-difference recurrence is ill-conditioned: a handful of near-singular points
# dominate every NORM, for the reference too -- an exact float32 re-implementation that merely sums in another order
# lands anywhere between 0.3x and 29x of the reference-float32 normwise distance from float64 (measured over seeds,
# DESIGN.md section 3).  SURVEY section 8d's bar "error vs the float64 oracle <= 4x the reference-float32 error vs the
# float64 oracle" is therefore evaluated on statistics that a few outliers cannot move:
#   x          median and 90th percentile of the pointwise |x - x64|               <= 4x the reference's
#   gradients  the batch is cut into G groups of rows (upstream gradient masked to one group at a time); per tensor
#              the MEDIAN over groups of the normwise error                          <= 4x the reference's (+ 1e-4 floor)
# (float32 De Hoog gradients are noise in the reference itself, 6e-3..O(1) relative; an outlier point contaminates
# only its own group.)
DEHOOG_FACTOR = 4.0


def fast_dehoog_batched(c, fp, ti, T, per_row):
  """line_integrate_batched for De Hoog / shared t with the (b, d) Python loop of the reference flattened into
  rows of ONE call of the 1-D routine (row-wise arithmetic, so the same values; 100x fewer ATen calls)."""
  assert c.alg == "dehoog" and not per_row
  B, TT, D, n = fp.shape
  f = fp.permute(0, 2, 1, 3).reshape(B * D * TT, n)
  y = orc._dehoog_1d(c, f, ti.repeat(B * D), T.repeat(B * D))
  return y.view(B, D, TT).permute(0, 2, 1)


class fast_dehoog_oracle:
  """Context: the oracle's batched De Hoog runs flattened (see fast_dehoog_batched)."""

  def __enter__(self):
    self._orig = orc.line_integrate_batched
    orig = self._orig
    orc.line_integrate_batched = lambda c, fp, ti, T, per_row: (fast_dehoog_batched(c, fp, ti, T, per_row) if
                                                                c.alg == "dehoog" and not per_row else orig(
                                                                    c, fp, ti, T, per_row))
    return self

  def __exit__(self, *a):
    orc.line_integrate_batched = self._orig


def pointwise_ratios(x, x_ref32, x64):
  """(median ratio, p90 ratio) of |x - x64| against |x_ref32 - x64| (pointwise, robust to outliers)."""
  e = (x.detach().double().cpu() - x64.double()).abs().flatten()
  e0 = (x_ref32.detach().double().cpu() - x64.double()).abs().flatten()
  floor = 1e-7 * float(x64.abs().median())
  return (float(e.median()) / max(float(e0.median()), floor), float(e.quantile(0.9)) / max(float(e0.quantile(0.9)), floor))


def group_mask(B, G, gi, dtype=torch.float32):
  m = torch.zeros(B, 1, 1, dtype=dtype)
  m[gi::G] = 1
  return m


def oracle_group_grads(weights, p, t, D, alg, sphere, terms, gout, G):
  """Oracle forward once, then one backward per group of rows: [G][7] gradients (W1,b1,W2,b2,W3,b3,p) and x."""
  ws = [w.detach().cpu().clone().requires_grad_(True) for w in weights]
  po = p.detach().cpu().clone().requires_grad_(True)
  with fast_dehoog_oracle():
    x = orc.laplace_reconstruct_oracle(ws, po, t.detach().cpu(), D, alg, sphere, terms)
    out = []
    for gi in range(G):
      m = group_mask(p.shape[0], G, gi, x.dtype)
      out.append([q.detach() for q in torch.autograd.grad(x, ws + [po], gout.to(x.dtype) * m, retain_graph=True)])
  return x.detach(), out


def group_median_ratios(groups, groups_ref32, groups64, floor=2.5e-5):
  """Per tensor: median over groups of the normwise error, divided by the reference's (floored)."""
  out = []
  for ti in range(len(groups64[0])):
    em = sorted(rel(g[ti], g64[ti]) for g, g64 in zip(groups, groups64))
    er = sorted(rel(g[ti], g64[ti]) for g, g64 in zip(groups_ref32, groups64))
    med = lambda v: 0.5 * (v[(len(v) - 1) // 2] + v[len(v) // 2])
    out.append((med(em), med(er), med(em) / max(med(er), floor)))
  return out
