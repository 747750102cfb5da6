"""Generate golden vectors from the LIVE reference (/root/reference) for the hot path.

Run in the build container only (the reference does not travel to the GPU box):

    python tests/golden/make_golden.py

It imports the reference ``torchlaplace`` (with the missing ``_iltcme`` module stubbed by the
oracle's synthetic table -- /root/reference/.MISSING_LARGE_BLOBS:1), runs
``laplace_reconstruct`` fwd + bwd and the class-level ILT API on seeded inputs, asserts that
``oracle/nl_oracle.py`` reproduces every output, and writes ``tests/golden/*.npz``.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import nl_oracle as orc  # noqa: E402

stub = types.ModuleType("torchlaplace._iltcme")
stub.cme_params_factory = orc.synthetic_cme_table
sys.modules["torchlaplace._iltcme"] = stub
sys.path.insert(0, "/root/reference")
import torchlaplace  # noqa: E402
from torchlaplace import inverse_laplace as ril  # noqa: E402
from torchlaplace import transformations as rtr  # noqa: E402

assert torchlaplace.__file__.startswith("/root/reference"), torchlaplace.__file__
torch.set_num_threads(1)  # deterministic reductions


class RefMLP(torch.nn.Module):
  """Canonical laplace_rep_func with weights injected (tests/test_laplace_rep_func.py:36-64)."""

  def __init__(self, weights, n, D, K, sphere=True):
    super().__init__()
    self.s_dim, self.output_dim, self.latent_dim, self.sphere = n, D, K, sphere
    H = weights[0].shape[0]
    self.linear_tanh_stack = torch.nn.Sequential(torch.nn.Linear(2 * n + K, H), torch.nn.Tanh(),
                                                 torch.nn.Linear(H, H), torch.nn.Tanh(),
                                                 torch.nn.Linear(H, 2 * n * D))
    lin = [self.linear_tanh_stack[i] for i in (0, 2, 4)]
    with torch.no_grad():
      for i, l in enumerate(lin):
        l.weight = torch.nn.Parameter(weights[2 * i].clone())
        l.bias = torch.nn.Parameter(weights[2 * i + 1].clone())
    self.phi_scale = torch.pi / 2.0 - -torch.pi / 2.0

  def params(self):
    l = [self.linear_tanh_stack[i] for i in (0, 2, 4)]
    return [l[0].weight, l[0].bias, l[1].weight, l[1].bias, l[2].weight, l[2].bias]

  def forward(self, i):
    out = self.linear_tanh_stack(i.view(-1, self.s_dim * 2 + self.latent_dim)).view(-1, 2 * self.output_dim,
                                                                                   self.s_dim)
    if not self.sphere:
      return out[:, :self.output_dim, :], out[:, self.output_dim:, :]
    theta = torch.nn.Tanh()(out[:, :self.output_dim, :]) * torch.pi
    phi = (torch.nn.Tanh()(out[:, self.output_dim:, :]) * self.phi_scale / 2.0 - torch.pi / 2.0 +
           self.phi_scale / 2.0)
    return theta, phi


def n_points(alg, terms):
  return orc.IltConsts(alg, terms).n


def recon_case(name, alg, terms, B, T, D, K, dtype, t_mode, seed, sphere=True, options=None, H=64):
  g = torch.Generator().manual_seed(seed)
  n = n_points(alg, terms)
  weights = orc.make_canonical_weights(n, D, K, H=H, dtype=dtype, seed=seed)
  p = torch.randn(B, K, generator=g).to(dtype)
  if t_mode == "shared":
    t = torch.linspace(20.0 / T, 20.0, T).to(dtype)
  elif t_mode == "shared_1T":
    t = (torch.rand(1, T, generator=g) * 10 + 0.1).to(dtype)
  else:
    t = (torch.rand(B, T, generator=g) * 19.9 + 0.1).to(dtype)
  gout = torch.randn(B, T, D, generator=g).to(dtype)

  # ---- live reference
  mlp = RefMLP(weights, n, D, K, sphere)
  pr = p.clone().requires_grad_(True)
  x_ref = torchlaplace.laplace_reconstruct(mlp, pr, t, recon_dim=D, ilt_algorithm=alg,
                                           use_sphere_projection=sphere, ilt_reconstruction_terms=terms,
                                           options=options)
  (x_ref * gout).sum().backward()
  grads_ref = [q.grad.detach().clone() for q in mlp.params()] + [pr.grad.detach().clone()]

  # ---- oracle must agree
  ws = [w.clone().requires_grad_(True) for w in weights]
  po = p.clone().requires_grad_(True)
  x_or = orc.laplace_reconstruct_oracle(ws, po, t, D, alg, sphere, terms, options)
  (x_or * gout).sum().backward()
  grads_or = [w.grad for w in ws] + [po.grad]
  tol = 1e-12 if dtype == torch.float64 else 2e-5
  if alg == "dehoog":
    tol = 1e-7 if dtype == torch.float64 else 5e-2
  err = rel(x_or, x_ref)
  assert err <= tol, (name, "x", err)
  for a, b in zip(grads_or, grads_ref):
    assert rel(a, b) <= tol * 10, (name, "grad", rel(a, b))
  exact = bool(torch.equal(x_or, x_ref))
  print(f"  {name:34s} x rel {err:.2e} exact={exact}")

  out = dict(alg=alg, terms=terms, D=D, K=K, H=H, sphere=int(sphere), t_mode=t_mode, dtype=str(dtype),
             p=p.numpy(), t=t.numpy(), gout=gout.numpy(), x=x_ref.detach().numpy())
  for i, w in enumerate(weights):
    out[f"w{i}"] = w.numpy()
  for i, gr in enumerate(grads_ref):
    out[f"g{i}"] = gr.numpy()
  if options:
    for k, v in options.items():
      out["opt_" + k] = v
  np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


def rel(a, b):
  return float((a - b).norm() / b.norm().clamp_min(1e-300))


def class_level():
  """tests/test_ilt.py:31-115 inputs, with outputs stored (the reference only smoke-tests)."""
  out = {}
  fs = lambda so: so / (so**2 + 1)
  for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
    cd = torch.cdouble if dtype == torch.float64 else torch.cfloat
    t = torch.linspace(0.0001, 10.0, 1000).to(dtype)
    out[f"t_{tag}"] = t.numpy()
    kw = dict(ilt_reconstruction_terms=33, torch_float_datatype=dtype, torch_complex_datatype=cd)
    for alg, cls in (("fixed_tablot", ril.FixedTablot), ("stehfest", ril.Stehfest), ("fourier", ril.Fourier),
                     ("dehoog", ril.DeHoog), ("cme", ril.CME), ("euler", ril.Euler), ("gaver", ril.Gaver)):
      dec = cls(**kw)
      y = dec(fs, t)
      r = dec.compute_s(t)
      s = r if alg == "stehfest" else r[0]
      c = orc.IltConsts(alg, 33, dtype)
      so, To = orc.compute_s(c, t)
      assert torch.equal(so, s), alg
      fh = fs(s)
      if alg in ("fourier", "dehoog"):
        y2 = dec.line_integrate(fh, t, r[1])
        yo = orc.line_integrate_1d(c, fh, t, To)
      elif alg in ("cme", "euler", "gaver"):
        y2 = dec.line_integrate(fh, r[1])
        yo = orc.line_integrate_1d(c, fh, To)
      else:
        y2 = dec.line_integrate(fh, t)
        yo = orc.line_integrate_1d(c, fh, t)
      e = rel(yo, y2)
      assert e < (1e-6 if alg != "dehoog" else 1e-2), (alg, e)
      rmse = float(torch.sqrt(torch.mean((torch.cos(t) - y2)**2)))
      print(f"  class {alg:13s} {tag} oracle-vs-ref {e:.1e}  rmse-vs-cos {rmse:.6g} exact={torch.equal(yo, y2)}")
      out[f"{alg}_{tag}_s_re"] = s.real.numpy()
      out[f"{alg}_{tag}_s_im"] = s.imag.numpy()
      out[f"{alg}_{tag}_y"] = y2.numpy()
      out[f"{alg}_{tag}_yfwd"] = y.numpy()
      out[f"{alg}_{tag}_rmse"] = rmse
  # sphere maps incl. the points at infinity and zero (tests/test_transformation.py:11-29)
  phi = torch.linspace(-torch.pi / 2 + 1e-6, torch.pi / 2 - 1e-6, 32).double()
  theta = torch.linspace(0 + 1e-6, torch.pi - 1e-6, 32).double()
  sr, si = rtr.spherical_riemann_to_complex(theta, phi)
  th2, ph2 = rtr.complex_to_spherical_riemann(sr, si)
  out.update(sph_theta=theta.numpy(), sph_phi=phi.numpy(), sph_sr=sr.numpy(), sph_si=si.numpy(),
             sph_theta_back=th2.numpy(), sph_phi_back=ph2.numpy())
  big = torch.tensor([float("inf"), 0.0, 1e30, -3.0]).double()
  big_i = torch.tensor([float("inf"), 0.0, -2e30, 0.0]).double()
  th3, ph3 = rtr.complex_to_spherical_riemann(big, big_i)
  out.update(edge_re=big.numpy(), edge_im=big_i.numpy(), edge_theta=th3.numpy(), edge_phi=ph3.numpy())
  np.savez_compressed(os.path.join(HERE, "class_level.npz"), **out)


def class_level_time_max():
  """The ``time_max`` / fixed-contour calls of tests/test_ilt.py:56-64, 101-106 with their OUTPUTS stored:
  FixedTablot.compute_s/line_integrate/line_integrate_multi/forward(time_max=) (inverse_laplace.py:272-304,
  328-339, 361-368, 393-400), DeHoog.compute_fixed_s / fixed_line_integrate / compute_s(time_max=) /
  forward(time_max=) (567-735, 897-1042), Fourier.compute_s(time_max=) / forward(time_max=) (1122-1147, 1269-1306).
  Two contours per dtype: the reference test's own (t up to 10, time_max = max t) and one whose time_max is not
  float32-representable (shows the ``torch.Tensor([time_max])`` rounding on the double path)."""
  out = {}
  fs = lambda so: so / (so**2 + 1)
  for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
    cd = torch.cdouble if dtype == torch.float64 else torch.cfloat
    kw = dict(ilt_reconstruction_terms=33, torch_float_datatype=dtype, torch_complex_datatype=cd)
    for case, t in (("a", torch.linspace(0.0001, 10.0, 1000).to(dtype)),
                    ("b", (torch.linspace(0.05, 7.3, 257, dtype=torch.float64)).to(dtype))):
      key = f"{tag}_{case}"
      tmax = torch.max(t)
      out[f"t_{key}"] = t.numpy()
      out[f"tmax_{key}"] = float(tmax)
      # ---- FixedTablot
      dec = ril.FixedTablot(**kw)
      c = orc.IltConsts("fixed_tablot", 33, dtype)
      s, r_over = dec.compute_s(t, time_max=tmax)
      so, ro = orc.compute_s_time_max(c, t, tmax)
      assert torch.equal(s, so) and torch.equal(r_over, ro), ("ft s", key)
      fh = fs(s)
      y = dec.line_integrate(fh, t, time_max=tmax.item())
      yo = orc.fixed_tablot_line_integrate_time_max(c, fh, t, tmax.item())
      yf = dec(fs, t, time_max=tmax)
      # (line_integrate_multi cannot be recorded: the reference broadcasts [T] against [T, D] there and raises
      # for any D != T, with and without time_max -- inverse_laplace.py:393-408; the product defines it as
      # line_integrate applied to every d slice and is tested against that.)
      print(f"  tmax fixed_tablot {key}: oracle-vs-ref {rel(yo, y):.1e} exact={torch.equal(yo, y)}  "
            f"fwd-vs-split {rel(yf, y):.1e}")
      assert rel(yo, y) < 1e-6
      out.update({f"ft_{key}_s_re": s.real.numpy(), f"ft_{key}_s_im": s.imag.numpy(), f"ft_{key}_y": y.numpy(),
                  f"ft_{key}_yfwd": yf.numpy(), f"ft_{key}_r_over": r_over.numpy()})
      # ---- DeHoog: fixed contour
      dec = ril.DeHoog(**kw)
      c = orc.IltConsts("dehoog", 33, dtype)
      sf = dec.compute_fixed_s(tmax)
      sfo = orc.dehoog_compute_fixed_s(c, tmax)
      assert torch.equal(sf, sfo), ("dh fixed s", key)
      fhf = fs(sf)
      yfix = dec.fixed_line_integrate(fhf, t, tmax)
      yfixo = orc.dehoog_fixed_line_integrate(c, fhf, t, tmax)
      s2, T2 = dec.compute_s(t, time_max=tmax)
      s2o, T2o = orc.compute_s_time_max(c, t, tmax)
      assert torch.equal(s2, s2o) and torch.equal(T2, T2o), ("dh s tmax", key)
      y2 = dec.line_integrate(fs(s2), t, T2)
      y2o = orc.line_integrate_1d(c, fs(s2), t, T2)
      yfw = dec(fs, t, time_max=tmax)
      print(f"  tmax dehoog       {key}: fixed oracle-vs-ref {rel(yfixo, yfix):.1e} exact={torch.equal(yfixo, yfix)}; "
            f"compute_s(time_max) {rel(y2o, y2):.1e} exact={torch.equal(y2o, y2)}; fwd-vs-split {rel(yfw, y2):.1e}; "
            f"rmse-vs-cos {float(torch.sqrt(torch.mean((torch.cos(t) - yfix)**2))):.3g}")
      assert rel(yfixo, yfix) < 1e-2 and rel(y2o, y2) < 1e-2
      out.update({f"dh_{key}_sfix_re": sf.real.numpy(), f"dh_{key}_sfix_im": sf.imag.numpy(),
                  f"dh_{key}_yfix": yfix.numpy(), f"dh_{key}_s_re": s2.real.numpy(), f"dh_{key}_s_im": s2.imag.numpy(),
                  f"dh_{key}_T": T2.real.numpy(), f"dh_{key}_y": y2.numpy(), f"dh_{key}_yfwd": yfw.numpy()})
      # ---- Fourier
      dec = ril.Fourier(**kw)
      c = orc.IltConsts("fourier", 33, dtype)
      s3, T3 = dec.compute_s(t, time_max=tmax)
      s3o, T3o = orc.compute_s_time_max(c, t, tmax)
      assert torch.equal(s3, s3o) and torch.equal(T3, T3o), ("fourier s tmax", key)
      y3 = dec.line_integrate(fs(s3), t, T3)
      y3o = orc.line_integrate_1d(c, fs(s3), t, T3)
      y3f = dec(fs, t, time_max=tmax)
      print(f"  tmax fourier      {key}: oracle-vs-ref {rel(y3o, y3):.1e} exact={torch.equal(y3o, y3)} "
            f"fwd-vs-split {rel(y3f, y3):.1e}")
      assert rel(y3o, y3) < 1e-6
      out.update({f"fo_{key}_s_re": s3.real.numpy(), f"fo_{key}_s_im": s3.imag.numpy(), f"fo_{key}_T": T3.real.numpy(),
                  f"fo_{key}_y": y3.numpy(), f"fo_{key}_yfwd": y3f.numpy()})
  np.savez_compressed(os.path.join(HERE, "class_level_tmax.npz"), **out)


def collate_cases():
  """torchlaplace.data_utils.basic_collate_fn (data_utils.py:51-74) on seeded trajectories: interpolation,
  extrapolation with steps, the ``hopper`` split."""
  from torchlaplace import data_utils as rdu
  g = torch.Generator().manual_seed(5)
  trajs = [torch.randn(24, 3, generator=g) for _ in range(5)]
  ts = torch.linspace(0.0, 2.3, 24)
  out = dict(trajs=torch.stack(trajs).numpy(), ts=ts.numpy())
  cases = dict(interp=dict(), extrap=dict(extrap=True), extrap_steps=dict(extrap=True, observe_step=2, predict_step=3),
               hopper=dict(extrap=True, data="hopper", data_type="test"))
  for name, kw in cases.items():
    d = rdu.basic_collate_fn(list(trajs), ts, **kw)
    assert d["mask_predicted_data"] is None and d["labels"] is None
    out[f"{name}_keys"] = ",".join(d.keys())
    out[f"{name}_mode"] = d["mode"]
    for k in ("observed_data", "observed_tp", "data_to_predict", "tp_to_predict", "observed_mask"):
      out[f"{name}_{k}"] = d[k].numpy()
    print(f"  collate {name}: " + " ".join(f"{k}{tuple(d[k].shape)}" for k in ("observed_data", "tp_to_predict")))
  np.savez_compressed(os.path.join(HERE, "collate.npz"), **out)


def main():
  f32, f64 = torch.float32, torch.float64
  if len(sys.argv) > 1 and sys.argv[1] == "tmax":  # only the time_max fixture (keeps the other files untouched)
    print("time_max / fixed-contour cases:")
    class_level_time_max()
    return
  if len(sys.argv) > 1 and sys.argv[1] == "collate":
    collate_cases()
    return
  print("laplace_reconstruct cases (reference vs oracle):")
  # the five shape cases of tests/test_laplace_rep_func.py:67-145 (B shrunk where it only adds bytes)
  recon_case("recon_fourier_f32_orig", "fourier", 33, 2, 3, 2, 2, f32, "shared", 1)
  recon_case("recon_fourier_f32_shared1T", "fourier", 33, 16, 25, 1, 2, f32, "shared_1T", 2)
  recon_case("recon_fourier_f32_b11t1d5k7", "fourier", 33, 11, 1, 5, 7, f32, "per_row", 3)
  recon_case("recon_fourier_f32_perrow", "fourier", 33, 16, 25, 1, 2, f32, "per_row", 4)
  recon_case("recon_fourier_f64_shared", "fourier", 33, 16, 25, 1, 2, f64, "shared", 5)
  recon_case("recon_fourier_f64_perrow", "fourier", 33, 8, 12, 2, 2, f64, "per_row", 6)
  recon_case("recon_fourier_f32_t65_d4", "fourier", 65, 8, 10, 4, 3, f32, "shared", 7)
  recon_case("recon_fourier_f32_opts", "fourier", 17, 8, 10, 1, 2, f32, "shared", 8,
             options=dict(alpha=1e-2, tol=5e-2, scale=3.0))
  recon_case("recon_cme_f32_shared", "cme", 33, 16, 25, 2, 2, f32, "shared", 9)
  recon_case("recon_cme_f64_perrow", "cme", 33, 8, 12, 1, 2, f64, "per_row", 10)
  recon_case("recon_cme_f32_even32", "cme", 32, 8, 10, 1, 2, f32, "shared", 11)
  recon_case("recon_euler_f32_shared", "euler", 33, 16, 25, 1, 2, f32, "shared", 12)
  recon_case("recon_euler_f64_shared", "euler", 17, 8, 12, 2, 2, f64, "shared", 13)
  recon_case("recon_dehoog_f64_shared", "dehoog", 33, 6, 10, 1, 2, f64, "shared", 14)
  recon_case("recon_dehoog_f32_shared", "dehoog", 17, 6, 10, 2, 2, f32, "shared", 15)
  recon_case("recon_fourier_f32_raw", "fourier", 33, 8, 10, 2, 2, f32, "shared", 16, sphere=False)
  recon_case("recon_fourier_f32_h32", "fourier", 17, 8, 10, 1, 2, f32, "shared", 17, H=32)
  print("class-level cases:")
  class_level()
  print("time_max / fixed-contour cases:")
  class_level_time_max()
  print("collate cases:")
  collate_cases()


if __name__ == "__main__":
  main()
