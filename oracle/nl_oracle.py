"""CPU oracle for the torchlaplace reconstruction hot path.  TEST INFRASTRUCTURE ONLY.

This file is a CPU restatement (plain torch ops on CPU tensors + autograd) of the reference
algorithm, written from SURVEY.md App. A/B and checked against the reference itself.  It is
NOT part of the product: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline``
/ ``--impl reference`` legs of ``bench.py`` may import it, and only as the checker / CPU
baseline.  The product path (``neurallaplace_b200``) never imports it and fails loudly when
the CUDA extension is missing.

Parity status
-------------
* pinned: fourier, dehoog, fixed_tablot, stehfest, euler, gaver, sphere maps, canonical MLP,
  ``laplace_reconstruct`` orchestration -- bit-for-bit against the live reference in the build
  container (``tests/golden/make_golden.py`` imports /root/reference, with the missing
  ``_iltcme`` module stubbed) and through the committed fixtures in ``tests/golden/*.npz``;
  plus the reference's notebook known answers (Fourier RMSE 0.017143, FixedTablot 0.436486).
* CME *arithmetic* is pinned the same way; CME *table values* are **parity unpinned**: the
  reference's ``torchlaplace/_iltcme.py`` is absent from the checkout
  (/root/reference/.MISSING_LARGE_BLOBS:1).  Tests use a synthetic table of the same schema.

Every function cites the reference lines it follows (paths relative to /root/reference).
The float32 rounding of "double" constants is deliberate.
"""
from __future__ import annotations

import math

import numpy as np
import torch

PI = math.pi


def _f32(x):
  """float32-rounded python float, mirrors ``torch.Tensor([x])``."""
  return torch.tensor([x], dtype=torch.float32)


def _cdt(dtype):
  return torch.cdouble if dtype == torch.float64 else torch.cfloat


def _to_complex(v):
  # inverse_laplace.py:56-74 (real_vector_to_complex): stack a zero imaginary part.
  if v.is_complex():
    return v
  return torch.complex(v, torch.zeros_like(v))


# --------------------------------------------------------------------------------------
# synthetic CME table (schema of inverse_laplace.py:1622-1634; values are NOT the reference's)
# --------------------------------------------------------------------------------------
def synthetic_cme_table(max_n=200):
  out = []
  for n in range(1, max_n):
    ra = np.random.RandomState(n)
    rb = np.random.RandomState(n + 999)
    out.append(
        dict(n=n,
             cv2=1.0 / (n + 1),
             c=0.1,
             omega=0.3,
             mu1=float(n)**0.5,
             a=list(ra.randn(n) * 0.1),
             b=list(rb.randn(n) * 0.1)))
  return out


# --------------------------------------------------------------------------------------
# constants per ILT family
# --------------------------------------------------------------------------------------
class IltConsts:
  """Bag of constants for one (algorithm, terms, dtype, options) combination."""

  def __init__(self, alg, terms, dtype=torch.float32, cme_table=None, **options):
    self.alg = alg
    self.terms = int(terms)
    self.dtype = dtype
    self.cdtype = _cdt(dtype)
    self.M = int((terms - 1) / 2)
    build = getattr(self, "_build_" + alg)
    self.cme_table = cme_table
    build(**options)

  # inverse_laplace.py:1089-1120 (Fourier.__init__) / 539-565 (DeHoog.__init__)
  def _build_fourier(self, alpha=1.0e-3, tol=None, scale=2.0, eps=1e-6):
    self.alpha = _f32(alpha)
    self.tol = _f32(tol) if tol is not None else self.alpha * 10.0
    self.scale = _f32(scale)
    self.k = torch.arange(self.terms, dtype=self.dtype)
    self.n = self.terms

  def _build_dehoog(self, alpha=1.0e-10, tol=None, scale=2.0):
    self._build_fourier(alpha=alpha, tol=tol, scale=scale)

  # inverse_laplace.py:250-270 (FixedTablot.__init__)
  def _build_fixed_tablot(self):
    M = self.M
    self.rdt = _f32((2.0 / 5.0) * M)
    k = torch.arange(1, M, dtype=self.dtype)
    self.theta = (k * torch.pi) / M
    self.sdt = self.rdt * self.theta * (1 / torch.tan(self.theta) + 1j)
    self.n = M

  # inverse_laplace.py:431-492 (Stehfest.__init__/_coeff)
  def _build_stehfest(self):
    from scipy.special import factorial  # only used here, like the reference (l.44)
    M = self.M
    self.ln2 = torch.log(_f32(2.0))
    Mv = M + 1 if M % 2 > 0 else M
    M2 = int(Mv / 2.0)
    fac = lambda x: float(factorial(x, exact=True))
    V = torch.empty((Mv,), dtype=self.cdtype)
    for k in range(1, Mv + 1):
      z = torch.zeros((min(k, M2) + 1,), dtype=self.cdtype)
      for j in range(int((k + 1) / 2.0), min(k, M2) + 1):
        z[j] = (j**M2 * fac(2 * j) / (fac(M2 - j) * fac(j) * fac(j - 1) * fac(k - j) * fac(2 * j - k)))
      V[k - 1] = (-1)**(k + M2) * torch.sum(z)
    self.V = V
    p_real = torch.arange(1, M + 1, dtype=self.dtype) * self.ln2
    self.sdt = _to_complex(p_real)
    self.n = M

  # inverse_laplace.py:1445-1514 (Euler.__init__)
  def _build_euler(self):
    fd, cd = self.dtype, self.cdtype
    n_euler = int(np.floor((self.terms - 1) / 2))
    end_element = torch.tensor([2.0], dtype=fd).pow(-n_euler)
    eta = torch.concat((torch.tensor([0.5], dtype=fd), torch.ones(n_euler, dtype=cd),
                        torch.zeros(n_euler - 1, dtype=cd), end_element))
    logsum = torch.cumsum(torch.log(torch.arange(1, n_euler + 1, dtype=fd)), dim=0)
    for k in torch.arange(1, n_euler):
      eta[2 * n_euler - k] = eta[2 * n_euler - k + 1] + torch.exp(
          logsum[n_euler - 1] - n_euler * torch.log(torch.tensor(2.0, dtype=fd)) - logsum[k - 1] -
          logsum[n_euler - k - 1])
    k = torch.arange(2 * n_euler + 1)
    beta = n_euler * torch.log(torch.tensor(10.0, dtype=fd)) / 3.0 + 1j * torch.pi * k
    final_element = torch.tensor(10.0, dtype=fd).pow(n_euler / 3.0)
    eta = final_element * (1 - (k % 2) * 2) * eta
    self.eta = eta.to(cd)
    self.beta = beta.to(cd)
    self.n = 2 * n_euler + 1

  # inverse_laplace.py:1334-1366 (Gaver.__init__); tables go through float32 (l.77-85)
  def _build_gaver(self):
    nfe = self.terms - 1 if self.terms % 2 == 1 else self.terms
    ndiv2 = nfe // 2
    eta = np.zeros(nfe)
    beta = np.zeros(nfe)
    logsum = np.concatenate(([0], np.cumsum(np.log(np.arange(1, nfe + 1)))))
    for k in range(1, nfe + 1):
      acc = 0.0
      for j in range(int(np.floor((k + 1) / 2)), min(k, ndiv2) + 1):
        acc += np.exp((ndiv2 + 1) * np.log(j) - logsum[ndiv2 - j] + logsum[2 * j] - 2 * logsum[j] -
                      logsum[k - j] - logsum[2 * j - k])
      eta[k - 1] = np.log(2.0) * (-1)**(k + ndiv2) * acc
      beta[k - 1] = k * np.log(2.0)
    self.eta = _np_complex_via_f32(eta).to(self.cdtype)
    self.beta = _np_complex_via_f32(beta).to(self.cdtype)
    self.n = nfe

  # inverse_laplace.py:1609-1638 (CME.__init__)
  def _build_cme(self):
    table = self.cme_table if self.cme_table is not None else synthetic_cme_table()
    params = table[0]
    for p in table:
      if p["cv2"] < params["cv2"] and p["n"] + 1 <= self.terms:
        params = p
    eta = np.concatenate(([params["c"]], np.array(params["a"]) + 1j * np.array(params["b"]))) * params["mu1"]
    beta = np.concatenate(([1], 1 + 1j * np.arange(1, params["n"] + 1) * params["omega"])) * params["mu1"]
    self.eta = _np_complex_via_f32(eta).to(self.cdtype)
    self.beta = _np_complex_via_f32(beta).to(self.cdtype)
    self.n = params["n"] + 1


def _np_complex_via_f32(z):
  # inverse_laplace.py:77-85: torch.Tensor(np) == float32 cast of both parts.
  z = np.asarray(z)
  re = torch.tensor(np.real(z), dtype=torch.float32).flatten()
  im = torch.tensor(np.imag(z) if np.iscomplexobj(z) else np.zeros_like(z), dtype=torch.float32).flatten()
  return torch.complex(re, im)


AW_FAMILY = ("cme", "euler", "gaver")


# --------------------------------------------------------------------------------------
# query points  s(t)
# --------------------------------------------------------------------------------------
def compute_s(c: IltConsts, ti):
  """Returns (s[rows, n], T) exactly as each reference class does.

  fourier/dehoog: inverse_laplace.py:1122-1147 / 703-735;  cme/euler/gaver: 1640-1643;
  fixed_tablot: 316-344;  stehfest: 494-496 (the reference returns a bare tensor; T=None here).
  """
  t = _to_complex(ti)
  if c.alg in ("fourier", "dehoog"):
    T = (t * c.scale).to(c.cdtype)
    gamma = c.alpha - torch.log(c.tol) / (c.scale * T)
    si = torch.matmul(1 / T.view(-1, 1), 1j * torch.pi * c.k.view(1, -1))
    return gamma.view(-1, 1) + si, T
  if c.alg in AW_FAMILY:
    return torch.matmul((1 / t).view(-1, 1), c.beta.view(1, -1)), t
  if c.alg == "fixed_tablot":
    s = torch.matmul(1 / t.view(-1, 1), c.sdt.view(1, -1))
    s0 = c.rdt / t
    return torch.hstack((s0.view(-1, 1), s)), c.rdt / t
  if c.alg == "stehfest":
    return torch.matmul((1 / t.view(-1, 1)), c.sdt.view(1, -1)), None
  raise ValueError(c.alg)


# --------------------------------------------------------------------------------------
# Riemann-sphere maps (transformations.py:28-89)
# --------------------------------------------------------------------------------------
def complex_to_sphere(s_real, s_imag):
  a2 = s_imag**2 + s_real**2
  ratio = torch.where(torch.isinf(a2), torch.ones_like(a2), (a2 - 1) / (a2 + 1))
  return torch.atan2(s_imag, s_real), torch.asin(ratio)


def sphere_to_complex(theta, phi):
  r = torch.tan(phi / 2 + torch.pi / 4)
  return torch.complex(r * torch.cos(theta), r * torch.sin(theta))


# --------------------------------------------------------------------------------------
# canonical laplace_rep_func (tests/test_laplace_rep_func.py:36-64)
# --------------------------------------------------------------------------------------
def mlp_forward(weights, inputs, D, n, sphere=True):
  """weights = (W1,b1,W2,b2,W3,b3) in nn.Linear layout; inputs [..., 2n+K] -> two [N, D, n]."""
  W1, b1, W2, b2, W3, b3 = weights
  x = inputs.reshape(-1, W1.shape[1])
  h1 = torch.tanh(torch.addmm(b1, x, W1.t()))
  h2 = torch.tanh(torch.addmm(b2, h1, W2.t()))
  out = torch.addmm(b3, h2, W3.t()).view(-1, 2 * D, n)
  if not sphere:  # SurfaceModel, experiments/baseline_models/neural_laplace.py:102-155
    return out[:, :D, :], out[:, D:, :]
  phi_scale = torch.pi / 2.0 - -torch.pi / 2.0
  theta = torch.tanh(out[:, :D, :]) * torch.pi
  phi = torch.tanh(out[:, D:, :]) * phi_scale / 2.0 - torch.pi / 2.0 + phi_scale / 2.0
  return theta, phi


# --------------------------------------------------------------------------------------
# line integration (the reduction over reconstruction terms)
# --------------------------------------------------------------------------------------
def _dehoog_1d(c: IltConsts, fp, ti, T):
  """DeHoog.line_integrate, inverse_laplace.py:737-866.  fp [T, n] complex."""
  t = _to_complex(ti)
  M = c.M
  NT = c.terms
  gamma = c.alpha - torch.log(c.tol) / (c.scale * T)
  qs, es = [], []
  q = fp[:, 1:2 * M + 1] / fp[:, 0:2 * M]
  q[:, 0] = fp[:, 1] / (fp[:, 0] / 2.0)
  qs.append(q)
  rows = t.shape[0]
  for r in range(1, M + 1):
    mr = 2 * (M - r) + 1
    e = torch.zeros((rows, NT), dtype=c.cdtype)
    if r == 1:
      e[:, 0:mr] = qs[r - 1][:, 1:mr + 1] - qs[r - 1][:, 0:mr]
    else:
      e[:, 0:mr] = qs[r - 1][:, 1:mr + 1] - qs[r - 1][:, 0:mr] + es[r - 2][:, 1:mr + 1]
    es.append(e)
    if r != M:
      q = torch.zeros((rows, 2 * M), dtype=c.cdtype)
      q[:, 0:mr] = qs[r - 1][:, 1:mr + 1] * es[r - 1][:, 1:mr + 1] / es[r - 1][:, 0:mr]
      qs.append(q)
  dt = torch.empty((rows, NT), dtype=c.cdtype)  # strided columns, like the reference (l.772-833)
  dt[:, 0] = fp[:, 0] / 2.0
  for r in range(1, M + 1):
    dt[:, 2 * r - 1] = -qs[r - 1][:, 0]
    dt[:, 2 * r] = -es[r - 1][:, 0]
  d = [dt[:, i] for i in range(NT)]
  Aim1 = torch.view_as_complex(torch.Tensor([0, 0]))
  Ai = d[0]
  Bim1 = torch.view_as_complex(torch.Tensor([1, 0]))
  Bi = torch.view_as_complex(torch.Tensor([1, 0]))
  z = torch.exp(1j * torch.pi * t / T)
  for i in range(1, 2 * M):
    Ait = Ai + d[i] * Aim1 * z
    Aim1, Ai = Ai, Ait
    Bit = Bi + d[i] * Bim1 * z
    Bim1, Bi = Bi, Bit
  brem = (1.0 + (d[2 * M - 1] - d[2 * M]) * z) / 2.0
  rem = -brem * (1.0 - torch.sqrt(1.0 + d[2 * M] * z / brem**2))
  A2 = Ai + rem * Aim1
  B2 = Bi + rem * Bim1
  return (torch.exp(gamma * t) / T * (A2 / B2).real).real


def _fixed_tablot_terms(c, fp_tail):
  # inverse_laplace.py:370-375, same association: (exp(sdt) * fp) * (...);  exp(rdt) is evaluated
  # in float32 even on the double path.
  th = c.theta
  return torch.exp(c.sdt) * fp_tail * (1.0 + 1j * th * (1.0 + (1 / torch.tan(th)**2)) - 1j * (1 / torch.tan(th)))


def line_integrate_1d(c: IltConsts, fp, ti, T=None):
  """Class-level ``line_integrate`` for fp [T, n] (FixedTablot 346-376, Stehfest 498-502,
  Fourier 1149-1175, DeHoog 737-866, CME/Euler/Gaver 1645-1647)."""
  if c.alg == "fourier":
    if T is None:
      T = ti
    t = _to_complex(ti)
    gamma = c.alpha - torch.log(c.tol) / (c.scale * T)
    tw = torch.exp(torch.matmul((t / T).view(-1, 1), 1j * c.k[1:].view(1, -1) * torch.pi))
    return ((1 / T) * torch.exp(gamma * t) * (fp[:, 0] / 2.0 + torch.sum(fp[:, 1:] * tw, 1))).real
  if c.alg == "dehoog":
    if T is None:
      T = ti
    return _dehoog_1d(c, fp, ti, T)
  if c.alg in AW_FAMILY:
    return (torch.matmul(fp, c.eta.view(-1, 1)).view(-1) / ti).real
  if c.alg == "fixed_tablot":
    t = _to_complex(ti)
    f0 = torch.exp(c.rdt) * fp[:, 0] / 2.0
    return ((2.0 / 5.0) * (torch.sum(_fixed_tablot_terms(c, fp[:, 1:]), 1) + f0) / t).real
  if c.alg == "stehfest":
    t = _to_complex(ti)
    return (torch.matmul(fp, c.V) * c.ln2 / t).real
  raise ValueError(c.alg)


def line_integrate_batched(c: IltConsts, fp, ti, T, per_row):
  """fp [B,T,D,n] -> [B,T,D].

  fourier: inverse_laplace.py:1177-1241; cme/euler/gaver: 1649-1683; dehoog: 868-895 (loop over
  batch and d of the 1-D routine).  fixed_tablot / stehfest (and dehoog with per-row t) have no
  batched form in the reference; SURVEY App. C.8 defines them as the class-level 1-D routine
  applied to every (b, d) slice, which is what this does.
  """
  B, TT, D, n = fp.shape
  if c.alg == "fourier":
    t = _to_complex(ti)
    gamma = c.alpha - torch.log(c.tol) / (c.scale * T)
    tw = torch.exp(torch.matmul((t / T).view(-1, 1), 1j * c.k[1:].view(1, -1) * torch.pi))
    if per_row:
      pre = ((1 / T) * torch.exp(gamma * t)).view(B, TT, 1)
      tw = tw.view(B, TT, 1, n - 1)
    else:
      pre = ((1 / T) * torch.exp(gamma * t)).view(1, -1, 1)
      tw = tw.view(TT, 1, n - 1)
    return (pre * (fp[:, :, :, 0] / 2.0 + torch.sum(fp[:, :, :, 1:] * tw, 3))).real
  if c.alg in AW_FAMILY:
    acc = torch.matmul(fp, c.eta.view(-1, 1))
    if per_row:
      return ((1 / T).view(B, TT, 1) * acc.view(B, TT, D)).real
    return ((1 / T).view(1, -1, 1) * acc.squeeze(-1)).real
  # loop forms
  out = []
  for b in range(B):
    tb = ti[b] if per_row else ti
    Tb = None if T is None else (T[b] if per_row else T)
    cols = []
    for d in range(D):
      cols.append(line_integrate_1d(c, fp[b, :, d, :], tb, Tb) if c.alg == "dehoog" else line_integrate_1d(
          c, fp[b, :, d, :], tb))
    out.append(torch.stack(cols, 1))
  return torch.stack(out)


# --------------------------------------------------------------------------------------
# time_max / fixed-contour variants: ONE contour for every time point
# --------------------------------------------------------------------------------------
def compute_s_time_max(c: IltConsts, ti, time_max):
  """``compute_s(ti, time_max=...)``: Fourier inverse_laplace.py:1135-1147, DeHoog 718-735, FixedTablot
  328-339.  ``torch.Tensor([time_max])`` rounds ``time_max`` to float32 on both dtype paths."""
  t = _to_complex(ti)
  if c.alg in ("fourier", "dehoog"):
    tmax = torch.Tensor([time_max]) * torch.ones_like(t)
    T = (tmax * c.scale).to(c.cdtype)
    gamma = c.alpha - torch.log(c.tol) / (c.scale * T)
    si = torch.matmul(1 / T.view(-1, 1), 1j * torch.pi * c.k.view(1, -1))
    return gamma.view(-1, 1) + si, T
  if c.alg == "fixed_tablot":
    ttm = _to_complex(torch.ones(t.shape, dtype=c.dtype)) * time_max
    s = torch.matmul(1 / ttm.view(-1, 1), c.sdt.view(1, -1))
    s0 = c.rdt / ttm
    return torch.hstack((s0.view(-1, 1), s)), c.rdt / time_max
  raise ValueError(c.alg)


def fixed_tablot_line_integrate_time_max(c: IltConsts, fp, ti, time_max):
  """FixedTablot.line_integrate / line_integrate_multi with ``time_max`` (inverse_laplace.py:361-368, 393-400).
  fp [T, n] or [T, D, n]."""
  t = _to_complex(ti)
  th = c.theta
  shape = (1.0 + 1j * th * (1.0 + (1 / torch.tan(th)**2)) - 1j * (1 / torch.tan(th)))
  if fp.dim() == 2:
    f0 = torch.exp((c.rdt * t) / time_max) * fp[:, 0] / 2.0
    pans = torch.exp((c.rdt * t) / time_max).view(-1, 1) * fp[:, 1:] * shape
    return ((2.0 / 5.0) * (torch.sum(pans, 1) + f0) * (1 / time_max)).real
  f0 = torch.exp((c.rdt * t) / time_max) * fp[:, :, 0] / 2.0
  pans = torch.exp((c.rdt * t) / time_max).view(-1, 1) * fp[:, :, 1:] * shape
  return ((2.0 / 5.0) * (torch.sum(pans, 2) + f0) * (1 / time_max)).real


def dehoog_compute_fixed_s(c: IltConsts, time_max):
  """DeHoog.compute_fixed_s (inverse_laplace.py:567-585): T and gamma are float32 on both dtype paths."""
  tmax = torch.Tensor([time_max])
  T = torch.Tensor([c.scale * tmax])
  gamma = c.alpha - torch.log(c.tol) / (c.scale * T)
  return gamma + 1j * torch.pi * torch.arange(c.terms, dtype=c.dtype) / T


def dehoog_fixed_line_integrate(c: IltConsts, fp, ti, time_max):
  """DeHoog.fixed_line_integrate (inverse_laplace.py:587-701): ONE quotient-difference table for the contour
  ``fp [n]``, the Pade recurrence evaluated at every ``ti`` through z = exp(i pi t / T)."""
  tv = _to_complex(ti)
  M, NT = c.M, c.terms
  tmax = torch.Tensor([time_max])
  T = torch.Tensor([c.scale * tmax])
  gamma = c.alpha - torch.log(c.tol) / (c.scale * T)
  qs, es = [], []
  q = fp[1:2 * M + 1] / fp[0:2 * M]
  q[0] = fp[1] / (fp[0] / 2.0)
  qs.append(q)
  for r in range(1, M + 1):
    mr = 2 * (M - r) + 1
    e = torch.zeros((NT,), dtype=c.cdtype)
    if r == 1:
      e[0:mr] = qs[r - 1][1:mr + 1] - qs[r - 1][0:mr]
    else:
      e[0:mr] = qs[r - 1][1:mr + 1] - qs[r - 1][0:mr] + es[r - 2][1:mr + 1]
    es.append(e)
    if r != M:
      q = torch.zeros((2 * M,), dtype=c.cdtype)
      q[0:mr] = qs[r - 1][1:mr + 1] * es[r - 1][1:mr + 1] / es[r - 1][0:mr]
      qs.append(q)
  d = torch.empty((NT,), dtype=c.cdtype)
  d[0] = fp[0] / 2.0
  for r in range(1, M + 1):
    d[2 * r - 1] = -qs[r - 1][0]
    d[2 * r] = -es[r - 1][0]
  Aim1 = torch.view_as_complex(torch.Tensor([0, 0]))
  Ai = d[0]
  Bim1 = torch.view_as_complex(torch.Tensor([1, 0]))
  Bi = torch.view_as_complex(torch.Tensor([1, 0]))
  z = torch.exp(1j * torch.pi * ti / T)
  for i in range(1, 2 * M):
    Ait = Ai + d[i] * Aim1 * z
    Aim1, Ai = Ai, Ait
    Bit = Bi + d[i] * Bim1 * z
    Bim1, Bi = Bi, Bit
  brem = (1.0 + (d[2 * M - 1] - d[2 * M]) * z) / 2.0
  rem = -brem * (1.0 - torch.sqrt(1.0 + d[2 * M] * z / brem**2))
  A2 = Ai + rem * Aim1
  B2 = Bi + rem * Bim1
  return (torch.exp(gamma * tv) / T * (A2 / B2).real).real


# --------------------------------------------------------------------------------------
# laplace_reconstruct (core.py:25-150) for the canonical MLP
# --------------------------------------------------------------------------------------
def laplace_reconstruct_oracle(weights,
                               p,
                               t,
                               recon_dim=None,
                               ilt_algorithm="fourier",
                               use_sphere_projection=True,
                               ilt_reconstruction_terms=33,
                               options=None,
                               cme_table=None,
                               rep_func=None):
  """CPU oracle of ``laplace_reconstruct``.  ``weights`` are the six canonical-MLP tensors
  (or pass ``rep_func`` = any callable mapping inputs -> (theta, phi))."""
  options = {} if options is None else dict(options)
  D = p.shape[1] if recon_dim is None else recon_dim
  c = IltConsts(ilt_algorithm, ilt_reconstruction_terms, p.dtype, cme_table=cme_table, **options)
  if t.dim() == 0:
    time_dim = 1
  elif t.dim() == 1:
    time_dim = t.shape[0]
  elif t.dim() == 2:
    time_dim = t.shape[1]
  else:
    raise ValueError("Unsupported time tensor shape, please use (batch, time_dim)")
  B = p.shape[0]
  per_row = t.dim() == 2 and B == t.shape[0]
  ts = torch.squeeze(t)
  s, T = compute_s(c, ts)
  s = s.reshape(-1, c.n)
  n = c.n
  if use_sphere_projection:
    th, ph = complex_to_sphere(torch.flatten(s.real), torch.flatten(s.imag))
    feat = torch.cat((th.reshape(s.real.shape), ph.reshape(s.imag.shape)), 1)
  else:
    feat = torch.cat((s.real, s.imag), 1)
  if per_row:
    inputs = torch.cat((feat.view(B, time_dim, -1), p.view(B, 1, -1).repeat(1, time_dim, 1)), 2)
  else:
    inputs = torch.cat((feat.view(1, time_dim, -1).repeat(B, 1, 1), p.view(B, 1, -1).repeat(1, time_dim, 1)), 2)
  if rep_func is not None:
    a, b = rep_func(inputs)
  else:
    a, b = mlp_forward(weights, inputs, D, n, sphere=use_sphere_projection)
  if use_sphere_projection:
    F = sphere_to_complex(a.reshape(-1), b.reshape(-1)).reshape(a.shape)
  else:
    F = torch.complex(a.reshape(-1), b.reshape(-1)).reshape(a.shape)
  ss = F.view(-1, time_dim, D, n)
  if per_row:
    tt = t.view(-1, time_dim)
    TTv = None if T is None else T.view(-1, time_dim)
  else:
    tt = ts.reshape(-1)
    TTv = None if T is None else T.reshape(-1)
  if c.alg == "fixed_tablot":
    TTv = None
  return line_integrate_batched(c, ss, tt, TTv, per_row)


def make_canonical_weights(n, D, K, H=64, dtype=torch.float32, seed=0, sphere=True):
  """Same init as tests/test_laplace_rep_func.py:43-51 (xavier-uniform weights, default biases)."""
  g = torch.Generator().manual_seed(seed)
  stack = torch.nn.Sequential(torch.nn.Linear(2 * n + K, H), torch.nn.Tanh(), torch.nn.Linear(H, H),
                              torch.nn.Tanh(), torch.nn.Linear(H, 2 * n * D))
  with torch.no_grad():
    for m in stack:
      if isinstance(m, torch.nn.Linear):
        fan_out, fan_in = m.weight.shape
        a = math.sqrt(6.0 / (fan_in + fan_out))
        m.weight.copy_((torch.rand(m.weight.shape, generator=g) * 2 - 1) * a)
        bound = 1 / math.sqrt(fan_in)
        m.bias.copy_((torch.rand(m.bias.shape, generator=g) * 2 - 1) * bound)
  ws = [stack[0].weight, stack[0].bias, stack[2].weight, stack[2].bias, stack[4].weight, stack[4].bias]
  return [w.detach().to(dtype).clone() for w in ws]
