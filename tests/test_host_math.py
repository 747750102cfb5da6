"""CPU: the device math (De Hoog recurrence + its hand-written reverse sweep, tanh/sphere head
chain rule, s points / sphere features / reduction coefficients) compiled for the HOST from the
very headers the CUDA kernels include, checked against the oracle."""
import ctypes
import math
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import nl_oracle as orc
from tests import helpers as hp

HOST_DIR = os.path.join(hp.ROOT, "tests", "host")
SO = os.path.join(HOST_DIR, "libnl_hosttest.so")
DP = ctypes.POINTER(ctypes.c_double)


@pytest.fixture(scope="module")
def hostlib():
  src = os.path.join(HOST_DIR, "dehoog_host.cu")
  hdrs = [os.path.join(hp.ROOT, "neurallaplace_b200", "csrc", h)
          for h in ("nl_common.cuh", "nl_dehoog.cuh", "nl_dehoog_sm.cuh")]
  if not os.path.exists(SO) or any(os.path.getmtime(f) > os.path.getmtime(SO) for f in [src] + hdrs):
    subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "--extended-lambda",
                           "--expt-relaxed-constexpr", "-diag-suppress", "20013,20015", "-o", SO, src])
  return ctypes.CDLL(SO)


def _p(a):
  return a.ctypes.data_as(DP)


@pytest.mark.parametrize("terms", [5, 9, 17, 33, 65])
def test_dehoog_forward_and_reverse_sweep(hostlib, terms):
  c = orc.IltConsts("dehoog", terms, torch.float64)
  t = torch.tensor([0.7, 3.1, 11.0], dtype=torch.float64)
  s, T = orc.compute_s(c, t)
  torch.manual_seed(terms)
  fp = (s / (s**2 + 1) + 0.05 * torch.randn(s.shape, dtype=torch.cdouble)).clone().requires_grad_(True)
  y = orc._dehoog_1d(c, fp, t, T)
  y.sum().backward()
  for row in range(t.numel()):
    a = torch.view_as_real(fp.detach()[row]).numpy().copy().reshape(-1)
    Tr = float(T[row].real)
    ang = math.pi * float(t[row]) / Tr
    w, dw, w2 = np.zeros(2), np.zeros(2 * terms), np.zeros(2)
    hostlib.nlh_dehoog(terms, ctypes.c_double(math.cos(ang)), ctypes.c_double(math.sin(ang)), _p(a), _p(w), _p(dw),
                       _p(w2))
    gamma = float(c.alpha) - float(torch.log(c.tol)) / (float(c.scale) * Tr)
    pref = math.exp(gamma * float(t[row])) / Tr
    tol = 1e-9 if terms <= 33 else 1e-6  # conditioning grows with M
    assert abs(pref * w[0] - float(y[row].detach())) <= tol * abs(float(y[row].detach()))
    assert w2[0] == w[0] and w2[1] == w[1]
    g = fp.grad[row]
    assert np.linalg.norm(pref * dw[0::2] - g.real.numpy()) <= tol * 100 * np.linalg.norm(g.real.numpy())
    assert np.linalg.norm(-pref * dw[1::2] - g.imag.numpy()) <= tol * 100 * np.linalg.norm(g.imag.numpy())
    # the shared-memory-strip variant the GPU launches (rolled loops, in-place adjoint columns, divisions in
    # batches of four) performs the same operations in the same order: identical results, bit for bit
    ws, dws, w2s = np.zeros(2), np.zeros(2 * terms), np.zeros(2)
    assert hostlib.nlh_dehoog_sm(terms, ctypes.c_double(math.cos(ang)), ctypes.c_double(math.sin(ang)), _p(a),
                                 _p(ws), _p(dws), _p(w2s)) == 0
    assert ws[0] == w[0] and ws[1] == w[1] and w2s[0] == w[0] and w2s[1] == w[1]
    assert np.array_equal(dws, dw), np.abs(dws - dw).max()



@pytest.mark.parametrize("terms", [9, 33, 65])
def test_dehoog_fixed_contour_parts(hostlib, terms):
  """The parts nl_dehoog_fixed.cu composes (table once per contour, Pade recurrence + reverse per time point,
  weighted sum of dw/dd_i, one reverse sweep) against autograd through the oracle's restatement of
  DeHoog.fixed_line_integrate (inverse_laplace.py:587-701)."""
  c = orc.IltConsts("dehoog", terms, torch.float64)
  tmax = 7.3
  t = torch.tensor([0.4, 2.2, 5.0, 7.3], dtype=torch.float64)
  s = orc.dehoog_compute_fixed_s(c, tmax)
  torch.manual_seed(terms)
  fp = (s / (s**2 + 1) + 0.02 * torch.randn(s.shape, dtype=torch.cdouble)).clone().requires_grad_(True)
  gout = torch.tensor([0.7, -1.3, 0.2, 2.0], dtype=torch.float64)
  y = orc.dehoog_fixed_line_integrate(c, fp, t, tmax)
  (y * gout).sum().backward()
  T = float(torch.Tensor([c.scale * torch.Tensor([tmax])]))
  gamma = float(c.alpha - torch.log(c.tol) / (c.scale * torch.Tensor([T])))
  ang = math.pi * t.numpy() / T
  z = np.stack([np.cos(ang), np.sin(ang)], 1).reshape(-1).copy()
  pref = np.exp(gamma * t.numpy()) / T
  g = (gout.numpy() * pref).copy()
  a = torch.view_as_real(fp.detach()).numpy().copy().reshape(-1)
  w, G, w2 = np.zeros(2 * t.numel()), np.zeros(2 * terms), np.zeros(2 * t.numel())
  assert hostlib.nlh_dehoog_fixed(terms, t.numel(), _p(z), _p(g), _p(a), _p(w), _p(G), _p(w2)) == 0
  tol = 1e-9 if terms <= 33 else 1e-6
  yk = pref * w[0::2]
  assert np.abs(yk - y.detach().numpy()).max() <= tol * np.abs(y.detach().numpy()).max()
  assert np.array_equal(w, w2)  # forward-only parts: same arithmetic
  gr = fp.grad
  assert np.linalg.norm(G[0::2] - gr.real.numpy()) <= 100 * tol * np.linalg.norm(gr.real.numpy())
  assert np.linalg.norm(-G[1::2] - gr.imag.numpy()) <= 100 * tol * np.linalg.norm(gr.imag.numpy())


def test_head_chain_rule(hostlib):
  rng = np.random.RandomState(0)
  for head in (0, 1):
    for _ in range(20):
      oa, ob, gre, gim = rng.randn(4)
      out = np.zeros(4)
      hostlib.nlh_head(head, ctypes.c_double(oa), ctypes.c_double(ob), ctypes.c_double(gre), ctypes.c_double(gim),
                       _p(out))
      a = torch.tensor(oa, dtype=torch.float64, requires_grad=True)
      b = torch.tensor(ob, dtype=torch.float64, requires_grad=True)
      if head == 0:
        F = orc.sphere_to_complex(torch.tanh(a) * torch.pi, torch.tanh(b) * torch.pi / 2)
      else:
        F = torch.complex(a, b)
      (F.real * gre + F.imag * gim).backward()
      assert abs(out[0] - float(F.real)) < 1e-12 * max(1, abs(float(F.real)))
      assert abs(out[1] - float(F.imag)) < 1e-12 * max(1, abs(float(F.imag)))
      assert abs(out[2] - float(a.grad)) < 1e-10 * max(1, abs(float(a.grad)))
      assert abs(out[3] - float(b.grad)) < 1e-10 * max(1, abs(float(b.grad)))


@pytest.mark.parametrize("alg", ["fourier", "dehoog", "euler", "cme", "fixed_tablot", "stehfest"])
def test_query_points_and_coefficients(hostlib, alg):
  from neurallaplace_b200 import _consts
  dtype = torch.float64
  c = _consts.get(alg, 33, dtype)
  o = orc.IltConsts(alg, 33, dtype)
  t = torch.tensor([0.31, 2.0, 17.5], dtype=dtype)
  s_ref, T = orc.compute_s(o, t)
  th, ph = orc.complex_to_sphere(s_ref.real, s_ref.imag)
  beta = torch.view_as_real(c.beta).numpy().copy().reshape(-1) if c.beta is not None else np.zeros(2)
  eta = torch.view_as_real(c.eta).numpy().copy().reshape(-1) if c.eta is not None else np.zeros(2)
  for i, tv in enumerate(t.tolist()):
    for k in (0, 1, c.n // 2, c.n - 1):
      out = np.zeros(7)
      hostlib.nlh_query(c.family, 0, c.n, ctypes.c_double(c.alpha), ctypes.c_double(c.log_tol),
                        ctypes.c_double(c.scale), _p(beta), _p(eta), ctypes.c_double(tv), k, _p(out))
      assert abs(out[0] - float(s_ref[i, k].real)) <= 1e-14 * abs(float(s_ref[i, k].real)) + 1e-300
      assert abs(out[1] - float(s_ref[i, k].imag)) <= 1e-14 * abs(float(s_ref[i, k].imag)) + 1e-300
      assert abs(out[2] - float(th[i, k])) < 1e-14 and abs(out[3] - float(ph[i, k])) < 1e-13
  # reduction coefficients: a unit impulse at term k through the oracle's line integration
  for k in (0, 1, c.n - 1):
    out = np.zeros(7)
    hostlib.nlh_query(c.family, 0, c.n, ctypes.c_double(c.alpha), ctypes.c_double(c.log_tol),
                      ctypes.c_double(c.scale), _p(beta), _p(eta), ctypes.c_double(2.0), k, _p(out))
    if alg == "dehoog":
      continue
    for comp, sign, idx in ((1.0 + 0j, 1.0, 4), (1j, -1.0, 5)):
      fp = torch.zeros(1, c.n, dtype=torch.cdouble)
      fp[0, k] = comp
      tt = torch.tensor([2.0], dtype=dtype)
      _, T1 = orc.compute_s(o, tt)
      y = orc.line_integrate_1d(o, fp, T1 if alg in ("cme", "euler") else tt, T1 if alg == "fourier" else None)
      assert abs(out[6] * sign * out[idx] - float(y[0])) <= 1e-13 * max(abs(float(y[0])), 1e-3), (alg, k, comp)
