"""CPU: the oracle reproduces every golden vector generated from the live reference
(tests/golden/make_golden.py) and the reference's published known answers."""
import os

import numpy as np
import pytest
import torch

from oracle import nl_oracle as orc
from tests import helpers as hp


@pytest.mark.parametrize("name", hp.golden_cases())
def test_oracle_matches_reference_fixture(name):
  c = hp.load_case(name)
  x, grads = hp.oracle_run(c["weights"], c["p"], c["t"], c["D"], c["alg"], c["sphere"], c["terms"], c["options"],
                           c["gout"])
  # same torch build => bit identical; tolerate last-ulp drift across BLAS/thread settings
  tol = 1e-12 if c["dtype"] == torch.float64 else 1e-5
  if c["alg"] == "dehoog":
    tol = 1e-6 if c["dtype"] == torch.float64 else 5e-2  # ill-conditioned
  assert x.shape == c["x"].shape and x.dtype == c["dtype"]
  assert hp.rel(x, c["x"]) <= tol
  for g, gr in zip(grads, c["grads"]):
    assert hp.rel(g, gr) <= tol * 10


def test_oracle_class_level_fixtures(golden_dir):
  z = np.load(os.path.join(golden_dir, "class_level.npz"))
  fs = lambda so: so / (so**2 + 1)
  for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
    t = torch.from_numpy(z[f"t_{tag}"])
    for alg in ("fixed_tablot", "stehfest", "fourier", "dehoog", "cme", "euler", "gaver"):
      c = orc.IltConsts(alg, 33, dtype)
      s, T = orc.compute_s(c, t)
      s_ref = torch.complex(torch.from_numpy(z[f"{alg}_{tag}_s_re"]), torch.from_numpy(z[f"{alg}_{tag}_s_im"]))
      assert hp.rel(torch.view_as_real(s), torch.view_as_real(s_ref)) < 1e-6
      fh = fs(s_ref)
      y = orc.line_integrate_1d(c, fh, T if alg in ("cme", "euler", "gaver") else t,
                                T if alg in ("fourier", "dehoog") else None)
      y_ref = torch.from_numpy(z[f"{alg}_{tag}_y"])
      scale = 1e-2 if alg in ("dehoog", "stehfest", "gaver") else 1e-5
      assert hp.rel(y, y_ref) < scale, alg


def test_known_answers_from_reference_notebook(golden_dir):
  # docs/source/notebooks/user_ilt.ipynb:89,206: RMSE vs cos(t), 33 terms, fp32
  z = np.load(os.path.join(golden_dir, "class_level.npz"))
  t = torch.from_numpy(z["t_f32"])
  fs = lambda so: so / (so**2 + 1)
  for alg, published in (("fourier", 0.01714298), ("fixed_tablot", 0.4364859)):
    c = orc.IltConsts(alg, 33, torch.float32)
    s, T = orc.compute_s(c, t)
    y = orc.line_integrate_1d(c, fs(s), t, T if alg == "fourier" else None)
    rmse = float(torch.sqrt(torch.mean((torch.cos(t) - y)**2)))
    assert abs(rmse - published) / published < 1e-4, (alg, rmse)
  # order-of-magnitude pins (fp32-noise dominated)
  assert 1e-6 < float(z["dehoog_f32_rmse"]) < 1e-4
  assert 5 < float(z["stehfest_f32_rmse"]) < 100


def test_sphere_edge_cases(golden_dir):
  z = np.load(os.path.join(golden_dir, "class_level.npz"))
  th, ph = orc.complex_to_sphere(torch.from_numpy(z["edge_re"]), torch.from_numpy(z["edge_im"]))
  assert torch.equal(th, torch.from_numpy(z["edge_theta"]))
  assert torch.allclose(ph, torch.from_numpy(z["edge_phi"]))
  assert abs(float(ph[0]) - np.pi / 2) < 1e-15  # point at infinity
  assert abs(float(ph[1]) + np.pi / 2) < 1e-15  # origin
  F = orc.sphere_to_complex(torch.from_numpy(z["sph_theta"]), torch.from_numpy(z["sph_phi"]))
  assert torch.allclose(F.real, torch.from_numpy(z["sph_sr"]))
  assert torch.allclose(F.imag, torch.from_numpy(z["sph_si"]))


def test_flattened_dehoog_oracle_equals_the_reference_loop():
  """tests/helpers.fast_dehoog_batched (the (b, d) loop of DeHoog.line_integrate_all_multi flattened into rows of
  one call) gives the values of the loop form, forward and backward -- it only exists to keep the robust float32
  De Hoog protocol of the GPU tests fast."""
  import torch
  from oracle import nl_oracle as orc
  from tests import helpers as hp
  for dtype, tol in ((torch.float64, 1e-13), (torch.float32, 1e-5)):
    g = torch.Generator().manual_seed(3)
    w = orc.make_canonical_weights(17, 2, 2, dtype=dtype, seed=1)
    p = torch.randn(5, 2, generator=g).to(dtype)
    t = torch.linspace(0.5, 9.0, 6).to(dtype)
    gout = torch.randn(5, 6, 2, generator=g).to(dtype)
    x0, g0 = hp.oracle_run(w, p, t, 2, "dehoog", True, 17, None, gout)
    with hp.fast_dehoog_oracle():
      x1, g1 = hp.oracle_run(w, p, t, 2, "dehoog", True, 17, None, gout)
    assert hp.rel(x1, x0) <= tol
    if dtype == torch.float64:
      for a, b in zip(g1, g0):
        assert hp.rel(a, b) <= tol * 1e3
    # float32: the SAME op chain evaluated over a different row arrangement (another SIMD / scalar split of the
    # complex divisions) already moves the float32 gradients by tens of percent -- they are noise in the reference
    # itself, which is why the GPU tests compare float32 De Hoog on outlier-proof statistics against float64.
