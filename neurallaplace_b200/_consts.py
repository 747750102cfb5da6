"""Host-side ILT constant tables.

The kernels take the family constants as plain tables / scalars; they are built here ONCE per
(algorithm, terms, dtype, options) with the same torch expressions the reference evaluates in
each class constructor, so that the float32 rounding the reference applies to its "double"
constants (/root/reference/torchlaplace/inverse_laplace.py:263, 446,
555-563, 1107-1115) is reproduced bit for bit.  The reference rebuilds them on every
``laplace_reconstruct`` call (core.py:84-93); here they are cached.
"""
from __future__ import annotations

import functools
import math

import numpy as np
import torch

from . import _iltcme
from ._lib import NL_AW, NL_DEHOOG, NL_FOURIER

ALGORITHMS = ("fourier", "dehoog", "cme", "euler", "gaver", "fixed_tablot", "stehfest")


def complex_dtype(dtype):
  return torch.cdouble if dtype == torch.float64 else torch.cfloat


class Consts:
  """family / n / scalars / (beta, eta) tables for one configuration (CPU tensors)."""
  __slots__ = ("alg", "terms", "dtype", "family", "n", "M", "alpha", "log_tol", "scale", "tol", "beta", "eta",
               "extra", "_dev")

  def __init__(self):
    self.alpha = self.log_tol = 0.0
    self.scale = 1.0
    self.tol = None
    self.beta = self.eta = None
    self.extra = {}
    self._dev = {}

  def tables_on(self, device):
    """(beta, eta) as interleaved real tensors on ``device`` (cached)."""
    if self.beta is None:
      return None, None
    key = str(device)
    if key not in self._dev:
      b = torch.view_as_real(self.beta.contiguous()).contiguous().to(device)
      e = torch.view_as_real(self.eta.contiguous()).contiguous().to(device)
      self._dev[key] = (b, e)
    return self._dev[key]


class FixedContour:
  """Fourier-type contour with a given abscissa (DeHoog.compute_fixed_s): what `_ops._desc` reads."""

  def __init__(self, base, gamma):
    self.n, self.family, self.scale = base.n, base.family, base.scale
    self.alpha, self.log_tol = float(gamma), 0.0

  def tables_on(self, device):
    return None, None


def _f32(x):
  return torch.tensor([x], dtype=torch.float32)


def _via_f32(z, cdtype):
  # the reference pushes numpy tables through torch.Tensor(...) == float32 (inverse_laplace.py:77-85)
  z = np.asarray(z)
  re = torch.tensor(np.real(z), dtype=torch.float32).flatten()
  im = torch.tensor(np.imag(z), dtype=torch.float32).flatten() if np.iscomplexobj(z) else torch.zeros_like(re)
  return torch.complex(re, im).to(cdtype)


def _fourier_like(c, alpha, tol, scale):
  a = _f32(alpha)
  tl = _f32(tol) if tol is not None else a * 10.0
  sc = _f32(scale)
  c.alpha = float(a)
  c.tol = float(tl)
  c.log_tol = float(torch.log(tl))
  c.scale = float(sc)
  c.n = c.terms


def _stehfest_weights(Mv, cdtype):
  # Salzer summation weights, each inner term rounded into the complex dtype before the sum
  # (inverse_laplace.py:461-492)
  from scipy.special import factorial
  M2 = int(Mv / 2.0)
  fac = lambda v: float(factorial(v, exact=True))
  V = torch.empty((Mv,), dtype=cdtype)
  for k in range(1, Mv + 1):
    z = torch.zeros((min(k, M2) + 1,), dtype=cdtype)
    for j in range(int((k + 1) / 2.0), min(k, M2) + 1):
      z[j] = j**M2 * fac(2 * j) / (fac(M2 - j) * fac(j) * fac(j - 1) * fac(k - j) * fac(2 * j - k))
    V[k - 1] = (-1)**(k + M2) * torch.sum(z)
  return V


@functools.lru_cache(maxsize=256)
def _build(alg, terms, dtype, opt_items, cme_version):
  options = dict(opt_items)
  c = Consts()
  c.alg, c.terms, c.dtype = alg, int(terms), dtype
  cd = complex_dtype(dtype)
  M = int((terms - 1) / 2)
  c.M = M
  if alg == "fourier":
    c.family = NL_FOURIER
    _fourier_like(c, options.pop("alpha", 1.0e-3), options.pop("tol", None), options.pop("scale", 2.0))
    options.pop("eps", None)
  elif alg == "dehoog":
    c.family = NL_DEHOOG
    _fourier_like(c, options.pop("alpha", 1.0e-10), options.pop("tol", None), options.pop("scale", 2.0))
  elif alg == "fixed_tablot":
    # inverse_laplace.py:261-269 and 370-375: beta_0 = r, beta_k = r*theta_k*(cot theta_k + i);
    # x = (2/5)/t * Re[ exp_f32(r) F_0/2 + sum exp(beta_k) F_k (1 + i theta_k (1+cot^2) - i cot) ]
    c.family = NL_AW
    rdt = _f32((2.0 / 5.0) * M)
    k = torch.arange(1, M, dtype=dtype)
    theta = (k * torch.pi) / M
    sdt = rdt * theta * (1 / torch.tan(theta) + 1j)
    shape = (1.0 + 1j * theta * (1.0 + (1 / torch.tan(theta)**2)) - 1j * (1 / torch.tan(theta)))
    w = torch.exp(sdt) * shape
    e0 = (torch.exp(rdt) / 2.0).to(cd)
    c.beta = torch.cat((rdt.to(cd), sdt.to(cd)))
    c.eta = (2.0 / 5.0) * torch.cat((e0, w.to(cd)))
    c.n = M
    c.extra = dict(rdt=rdt, theta=theta, sdt=sdt, shape=shape.to(cd))
  elif alg == "stehfest":
    c.family = NL_AW
    ln2 = torch.log(_f32(2.0))
    Mv = M + 1 if M % 2 > 0 else M
    V = _stehfest_weights(Mv, cd)
    beta_re = torch.arange(1, M + 1, dtype=dtype) * ln2
    c.beta = torch.complex(beta_re, torch.zeros_like(beta_re)).to(cd)
    c.extra = dict(V=V, ln2=ln2, Mv=Mv)
    c.eta = (V * ln2).to(cd) if Mv == M else None  # odd M: the reference itself is inconsistent (App. C.4)
    c.n = M
  elif alg == "euler":
    # inverse_laplace.py:1456-1514
    c.family = NL_AW
    ne = int(np.floor((terms - 1) / 2))
    end_element = torch.tensor([2.0], dtype=dtype).pow(-ne)
    eta = torch.concat((torch.tensor([0.5], dtype=dtype), torch.ones(ne, dtype=cd), torch.zeros(ne - 1, dtype=cd),
                        end_element))
    logsum = torch.cumsum(torch.log(torch.arange(1, ne + 1, dtype=dtype)), dim=0)
    two = torch.tensor(2.0, dtype=dtype)
    for k in torch.arange(1, ne):
      eta[2 * ne - k] = eta[2 * ne - k + 1] + torch.exp(logsum[ne - 1] - ne * torch.log(two) - logsum[k - 1] -
                                                        logsum[ne - k - 1])
    k = torch.arange(2 * ne + 1)
    beta = ne * torch.log(torch.tensor(10.0, dtype=dtype)) / 3.0 + 1j * torch.pi * k
    eta = torch.tensor(10.0, dtype=dtype).pow(ne / 3.0) * (1 - (k % 2) * 2) * eta
    c.eta, c.beta = eta.to(cd), beta.to(cd)
    c.n = 2 * ne + 1
  elif alg == "gaver":
    # inverse_laplace.py:1345-1366
    c.family = NL_AW
    nfe = terms - 1 if terms % 2 == 1 else terms
    nd2 = nfe // 2
    eta = np.zeros(nfe)
    beta = np.zeros(nfe)
    logsum = np.concatenate(([0], np.cumsum(np.log(np.arange(1, nfe + 1)))))
    for k in range(1, nfe + 1):
      acc = 0.0
      for j in range(int(np.floor((k + 1) / 2)), min(k, nd2) + 1):
        acc += np.exp((nd2 + 1) * np.log(j) - logsum[nd2 - j] + logsum[2 * j] - 2 * logsum[j] - logsum[k - j] -
                      logsum[2 * j - k])
      eta[k - 1] = np.log(2.0) * (-1)**(k + nd2) * acc
      beta[k - 1] = k * np.log(2.0)
    c.eta, c.beta = _via_f32(eta, cd), _via_f32(beta, cd)
    c.n = nfe
  elif alg == "cme":
    # inverse_laplace.py:1622-1638: entry with the smallest cv2 among n+1 <= terms
    c.family = NL_AW
    table = _iltcme.cme_params_factory()
    params = table[0]
    for p in table:
      if p["cv2"] < params["cv2"] and p["n"] + 1 <= terms:
        params = p
    eta = np.concatenate(([params["c"]], np.array(params["a"]) + 1j * np.array(params["b"]))) * params["mu1"]
    beta = np.concatenate(([1], 1 + 1j * np.arange(1, params["n"] + 1) * params["omega"])) * params["mu1"]
    c.eta, c.beta = _via_f32(eta, cd), _via_f32(beta, cd)
    c.n = params["n"] + 1
  else:
    raise ValueError(f'Invalid ILT algorithm "{alg}"')
  if options:
    raise TypeError(f"unexpected ILT option(s) for {alg}: {sorted(options)}")
  return c


def get(alg, terms, dtype, options=None) -> Consts:
  options = options or {}
  items = tuple(sorted((k, (float(v) if v is not None else None)) for k, v in options.items()))
  return _build(alg, int(terms), dtype, items, _iltcme.table_version())
