"""ILT algorithm classes -- same names, constructor arguments and methods as
``torchlaplace.inverse_laplace`` (/root/reference/torchlaplace/inverse_laplace.py:88-1689), with
``compute_s`` and every ``line_integrate*`` running as hand-written CUDA (``nl_query_points`` /
``nl_line_integrate_*`` in include/nl_b200.h).  Spellings follow the reference (``FixedTablot``).

Differences from the reference, all deliberate:
  * constants are built once per configuration on the host and cached, not on every call and not
    on a module-global device; tensors may live on any CUDA device (CPU inputs are staged to the
    current CUDA device and results returned on the input's device -- compute is always on GPU);
  * ``line_integrate_all_multi[_batch_time]`` exist for every class (the reference only has them
    for Fourier / CME / Euler / Gaver and a Python loop for DeHoog);
  * ``Stehfest`` with an odd ``(terms-1)/2`` raises ``ValueError`` (the reference builds
    mismatched tables and fails inside ``matmul``);
  * gradients flow to ``fp`` only (the reference can also differentiate w.r.t. ``ti``; no
    caller does).
"""
from __future__ import annotations

import torch
from torch import nn

from . import _consts, _lib, _ops

TORCH_FLOAT_DATATYPE = torch.float32
TORCH_COMPLEX_DATATYPE = torch.cfloat


def real_vector_to_complex(vector):
  """inverse_laplace.py:56-74."""
  if vector.is_complex():
    return vector
  return torch.complex(vector, torch.zeros_like(vector))


def _real(v):
  return v.real if v.is_complex() else v


def _stage(x, dtype, device=None):
  """-> contiguous CUDA tensor of ``dtype`` on ``device`` (default: where it is, or the current CUDA device for
  host tensors) + the device to return results on."""
  _lib.require_cuda()
  if not isinstance(x, torch.Tensor):
    x = torch.as_tensor(x)
  home = x.device
  x = _real(x).to(dtype)
  if device is not None:
    x = x.to(device)
  elif not x.is_cuda:
    x = x.cuda()
  return x.contiguous(), home


class InverseLaplaceTransformAlgorithmBase(nn.Module):
  """Base class (inverse_laplace.py:88-222): ILT-Query ``compute_s`` + ILT-Compute ``line_integrate*``."""
  _alg = None

  def __init__(self, ilt_reconstruction_terms=33, torch_float_datatype=TORCH_FLOAT_DATATYPE,
               torch_complex_datatype=TORCH_COMPLEX_DATATYPE, **options):
    super().__init__()
    self.ilt_reconstruction_terms = ilt_reconstruction_terms
    self.torch_float_datatype = torch_float_datatype
    self.torch_complex_datatype = torch_complex_datatype
    self.M = int((ilt_reconstruction_terms - 1) / 2)
    self._options = options
    if self._alg is not None:
      self._c = _consts.get(self._alg, ilt_reconstruction_terms, torch_float_datatype, options)

  # -- helpers -----------------------------------------------------------------------------
  def _query(self, ti, Tper=None):
    t, home = _stage(ti, self.torch_float_datatype)
    flat = t.reshape(-1)
    Tp = None if Tper is None else _stage(Tper, self.torch_float_datatype, flat.device)[0].reshape(-1)
    s, _ = _ops.query_points(self._c, flat, _lib.NL_HEAD_RAW, True, False, Tp)
    return s, flat, home

  def _integrate(self, fp, ti, Tper, batch_time, eta_override=None):
    """fp [..., n] complex with leading dims (T) / (T,D) / (B,T,D); returns the real reduction."""
    _lib.require_cuda()
    home = fp.device
    f = fp.to(self.torch_complex_datatype)
    if not f.is_cuda:
      f = f.cuda()
    n = f.shape[-1]
    lead = f.shape[:-1]
    t, _ = _stage(ti, self.torch_float_datatype, f.device)
    Tp = None if Tper is None else _stage(Tper, self.torch_float_datatype, f.device)[0]
    if eta_override is not None:
      eta_override = eta_override.to(f.device)
    if batch_time:
      B, T, D = lead
      t_mode = _lib.NL_T_PER_ROW
      t = t.reshape(B, T)
      Tp = None if Tp is None else Tp.reshape(B, T)
    else:
      t = t.reshape(-1)
      T = t.numel()
      t_mode = _lib.NL_T_SHARED
      if len(lead) == 1:
        B, D = 1, 1
      elif len(lead) == 2:
        B, D = 1, lead[1]
      else:
        B, D = lead[0], lead[2]
      if Tp is not None:
        Tp = Tp.reshape(-1)
        if Tp.numel() == 1 and T != 1:
          Tp = Tp.expand(T)
    if (lead[0] if len(lead) <= 2 else lead[1]) != T:
      raise ValueError("fp and ti disagree on the number of time points")
    f4 = torch.view_as_real(f.reshape(B, T, D, n).contiguous())
    x = _ops.line_integrate(self._c, f4, None, t, Tp, _lib.NL_F_COMPLEX_INTERLEAVED, t_mode, eta_override)
    return x.reshape(lead).to(home)

  # -- reference API (overridden where the signature differs) ----------------------------
  def compute_s(self, ti):
    raise NotImplementedError("compute_s method is not implemented yet for this ILT algorithm")

  def line_integrate(self, fp, ti):
    raise NotImplementedError("line_integrate method is not implemented yet for this ILT algorithm")

  def line_integrate_all_multi(self, fp, ti, T):
    raise NotImplementedError("line_integrate_all_multi method is not implemented yet for this ILT algorithm")

  def forward(self, fs, ti):
    raise NotImplementedError("Forward method is not implemented yet for this ILT algorithm")


# ------------------------------------------------------------------------------------------
class _ContourBase(InverseLaplaceTransformAlgorithmBase):
  """Fourier-series contour s_k = gamma + i*pi*k/T shared by Fourier and DeHoog."""

  def __init__(self, ilt_reconstruction_terms=33, alpha=None, tol=None, scale=2.0,
               torch_float_datatype=TORCH_FLOAT_DATATYPE, torch_complex_datatype=TORCH_COMPLEX_DATATYPE, **kw):
    opts = dict(alpha=alpha if alpha is not None else self._default_alpha, tol=tol, scale=scale)
    opts.update(kw)
    super().__init__(ilt_reconstruction_terms, torch_float_datatype, torch_complex_datatype, **opts)
    self.nt = ilt_reconstruction_terms
    c = self._c
    self.alpha = torch.tensor([c.alpha], dtype=torch.float32)
    self.tol = torch.tensor([c.tol], dtype=torch.float32)
    self.scale = torch.tensor([c.scale], dtype=torch.float32)

  def _T_for(self, ti, time_max, Ti):
    if Ti is not None:
      # the reference's isinstance(Ti, dtype) check can never pass (App. C.3)
      raise ValueError("Invalid Ti type")
    if time_max is None:
      return None
    t = _real(torch.as_tensor(ti))
    # `torch.Tensor([time_max])` (inverse_laplace.py:1136, 719): time_max is rounded to float32 on both dtype
    # paths; T = scale * tmax with the float32 scale, in the working dtype
    tm = float(torch.tensor(float(time_max), dtype=torch.float32))
    return torch.full(t.reshape(-1).shape, tm, dtype=self.torch_float_datatype) * float(self._c.scale)

  def compute_s(self, ti, time_max=None, Ti=None):
    """inverse_laplace.py:1122-1147 / 703-735: returns (s[T, n] complex, T complex)."""
    Tper = self._T_for(ti, time_max, Ti)
    s, flat, home = self._query(ti, Tper)
    T = (flat * float(self._c.scale)) if Tper is None else Tper.to(flat.device)
    return s.to(home), real_vector_to_complex(T).to(self.torch_complex_datatype).to(home)

  def line_integrate(self, fp, ti, T=None):
    """fp [T, n] -> [T]   (inverse_laplace.py:1149-1175 / 737-866; T defaults to ti as there)."""
    return self._integrate(fp, ti, ti if T is None else T, False)

  def line_integrate_multi(self, fp, ti, T):
    """fp [T, D, n] -> [T, D]   (inverse_laplace.py:1243-1267)."""
    return self._integrate(fp, ti, T, False)

  def line_integrate_all_multi(self, fp, ti, T):
    """fp [B, T, D, n], ti [T], T [T] -> [B, T, D]   (inverse_laplace.py:1177-1208 / 868-895)."""
    return self._integrate(fp, ti, T, False)

  def line_integrate_all_multi_batch_time(self, fp, ti, T):
    """fp [B, T, D, n], ti [B, T], T [B, T] -> [B, T, D]   (inverse_laplace.py:1210-1241)."""
    return self._integrate(fp, ti, T, True)

  def forward(self, fs, ti, time_max=None, Ti=None):
    """inverse_laplace.py:1269-1306 / 897-1042.  (The reference evaluates ``fs`` on the full ``[T, n]`` grid of
    s points also when ``time_max`` makes every row the same contour; so does this, ``fs`` is the caller's.)"""
    s, T = self.compute_s(ti, time_max, Ti)
    return self.line_integrate(fs(s), ti, T)


class Fourier(_ContourBase):
  """Fourier-series ILT (inverse_laplace.py:1045-1306)."""
  _alg = "fourier"
  _default_alpha = 1.0e-3

  def __init__(self, ilt_reconstruction_terms=33, alpha=1.0e-3, tol=None, scale=2.0, eps=1e-6,
               torch_float_datatype=TORCH_FLOAT_DATATYPE, torch_complex_datatype=TORCH_COMPLEX_DATATYPE):
    super().__init__(ilt_reconstruction_terms, alpha, tol, scale, torch_float_datatype, torch_complex_datatype)
    self.eps = eps
    self.k = torch.arange(self.nt, dtype=torch_float_datatype)


class DeHoog(_ContourBase):
  """De Hoog accelerated Fourier ILT (inverse_laplace.py:514-1042)."""
  _alg = "dehoog"
  _default_alpha = 1.0e-10

  def __init__(self, ilt_reconstruction_terms=33, alpha=1.0e-10, tol=None, scale=2.0,
               torch_float_datatype=TORCH_FLOAT_DATATYPE, torch_complex_datatype=TORCH_COMPLEX_DATATYPE):
    super().__init__(ilt_reconstruction_terms, alpha, tol, scale, torch_float_datatype, torch_complex_datatype)

  def _fixed_contour(self, time_max):
    """(T, gamma) of the single contour for ``time_max`` as the reference computes them: float32 on both dtype
    paths (``torch.Tensor([...])``, inverse_laplace.py:579-582, 640-642)."""
    c = self._c
    tmax = torch.tensor([float(time_max)], dtype=torch.float32)
    scale = torch.tensor([c.scale], dtype=torch.float32)
    T = (scale * tmax).to(torch.float32)
    gamma = torch.tensor([c.alpha], dtype=torch.float32) - torch.log(torch.tensor([c.tol], dtype=torch.float32)) / (
        scale * T)
    return float(T), float(gamma)

  def compute_fixed_s(self, time_max):
    """s[n] for the single time point ``time_max`` (inverse_laplace.py:567-585): gamma + i pi k / T with the
    float32 (T, gamma) of that contour, evaluated by ``nl_query_points``."""
    T, gamma = self._fixed_contour(time_max)
    home = time_max.device if isinstance(time_max, torch.Tensor) else torch.device("cpu")
    _lib.require_cuda()
    dev = home if home.type == "cuda" else torch.device("cuda", torch.cuda.current_device())
    one = torch.ones(1, dtype=self.torch_float_datatype, device=dev)
    # the kernel forms gamma = alpha - log_tol / (scale * T): alpha := gamma, log_tol := 0 pins it
    s, _ = _ops.query_points(_consts.FixedContour(self._c, gamma), one, _lib.NL_HEAD_RAW, True, False, one * T)
    return s.reshape(-1).to(home)

  def fixed_line_integrate(self, fp, ti, time_max):
    """One contour ``fp [n]`` (or ``[rows, n]``) evaluated at every ``ti`` (inverse_laplace.py:587-701): the
    quotient-difference table is built once per contour, the Pade recurrence runs per time point
    (``nl_dehoog_fixed_forward/backward``).  Returns ``[T]`` (or ``[rows, T]``)."""
    _lib.require_cuda()
    T, gamma = self._fixed_contour(time_max)
    home = fp.device
    f = fp.to(self.torch_complex_datatype)
    if not f.is_cuda:
      f = f.cuda()
    t, _ = _stage(ti, self.torch_float_datatype, f.device)
    t = t.reshape(-1)
    single = f.dim() == 1
    F = torch.view_as_real(f.reshape(-1, f.shape[-1]).contiguous())
    if F.shape[1] != self.nt:
      raise ValueError(f"fp has {F.shape[1]} terms, the contour has {self.nt}")
    x = _ops.dehoog_fixed(F, t, T, gamma)
    return (x.reshape(-1) if single else x).to(home)


# ------------------------------------------------------------------------------------------
class _AbateWhittBase(InverseLaplaceTransformAlgorithmBase):
  """s_k = beta_k / t,  x = Re(sum_k eta_k F_k) / t  (CME, Euler, Gaver)."""

  def __init__(self, ilt_reconstruction_terms=33, torch_float_datatype=TORCH_FLOAT_DATATYPE,
               torch_complex_datatype=TORCH_COMPLEX_DATATYPE):
    super().__init__(ilt_reconstruction_terms, torch_float_datatype, torch_complex_datatype)
    self.eta = self._c.eta
    self.beta = self._c.beta

  def compute_s(self, ti):
    """inverse_laplace.py:1640-1643: returns (s[T, n], t as complex)."""
    s, flat, home = self._query(ti)
    return s.to(home), real_vector_to_complex(flat).to(self.torch_complex_datatype).to(home)

  def line_integrate(self, fp, ti):
    return self._integrate(fp, ti, None, False)

  def line_integrate_all_multi(self, fp, ti, T):
    return self._integrate(fp, ti, T, False)

  def line_integrate_all_multi_batch_time(self, fp, ti, T):
    return self._integrate(fp, ti, T, True)

  def forward(self, fs, ti):
    s, _ = self.compute_s(ti)
    return self.line_integrate(fs(s), ti)


class CME(_AbateWhittBase):
  """Concentrated matrix-exponential ILT (inverse_laplace.py:1581-1689)."""
  _alg = "cme"


class Euler(_AbateWhittBase):
  """Euler ILT in the Abate-Whitt framework (inverse_laplace.py:1420-1578)."""
  _alg = "euler"


class Gaver(_AbateWhittBase):
  """Gaver(-Stehfest) ILT in the Abate-Whitt framework (inverse_laplace.py:1309-1417)."""
  _alg = "gaver"


class Stehfest(_AbateWhittBase):
  """Stehfest ILT (inverse_laplace.py:411-511); n = (terms-1)/2 real s points."""
  _alg = "stehfest"

  def __init__(self, ilt_reconstruction_terms=33, torch_float_datatype=TORCH_FLOAT_DATATYPE,
               torch_complex_datatype=TORCH_COMPLEX_DATATYPE):
    InverseLaplaceTransformAlgorithmBase.__init__(self, ilt_reconstruction_terms, torch_float_datatype,
                                                  torch_complex_datatype)
    if self._c.eta is None:
      raise ValueError("Stehfest needs an even (ilt_reconstruction_terms - 1) / 2 "
                       f"(got {self.M}); e.g. 33, 37, 41, 65, 129 terms")
    self.V = self._c.extra["V"]
    self.ln2 = self._c.extra["ln2"]
    self.sdt = self._c.beta
    self.eta = self._c.eta
    self.beta = self._c.beta

  def compute_s(self, ti):
    """Returns a bare ``s`` like the reference (inverse_laplace.py:494-496)."""
    s, _, home = self._query(ti)
    return s.to(home)

  def forward(self, fs, ti):
    return self.line_integrate(fs(self.compute_s(ti)), ti)


class FixedTablot(_AbateWhittBase):
  """Fixed Talbot ILT (inverse_laplace.py:225-408); n = (terms-1)/2 s points."""
  _alg = "fixed_tablot"

  def __init__(self, ilt_reconstruction_terms=33, torch_float_datatype=TORCH_FLOAT_DATATYPE,
               torch_complex_datatype=TORCH_COMPLEX_DATATYPE):
    super().__init__(ilt_reconstruction_terms, torch_float_datatype, torch_complex_datatype)
    ex = self._c.extra
    self.rdt, self.theta, self.sdt = ex["rdt"], ex["theta"], ex["sdt"]
    self._eta_tmax = None

  def _tmax_tables(self, ti, time_max):
    # time_max branch (inverse_laplace.py:361-368): x = (2/5)/tmax * exp(r t/tmax) * Re[F_0/2 + sum F_k c_k]
    # -> same kernel with eta' = [1/2, c_k] and the divisor tmax / ((2/5) exp(r t/tmax)).
    if self._eta_tmax is None:
      one = torch.tensor([0.5], dtype=self.torch_complex_datatype)
      self._eta_tmax = torch.view_as_real(torch.cat((one, self._c.extra["shape"]))).contiguous()
    t = _real(torch.as_tensor(ti)).to(self.torch_float_datatype).reshape(-1)
    tm = float(time_max)
    r = float(self.rdt)
    div = tm / ((2.0 / 5.0) * torch.exp((r * t) / tm))
    return self._eta_tmax, div

  def compute_s(self, ti, time_max=None):
    """inverse_laplace.py:316-344: returns (s[T, n], r/t) (or r/time_max)."""
    if time_max is None:
      s, flat, home = self._query(ti)
      return s.to(home), real_vector_to_complex(float(self.rdt) / flat).to(self.torch_complex_datatype).to(home)
    t = _real(torch.as_tensor(ti)).reshape(-1)
    tm = torch.full(t.shape, float(time_max), dtype=self.torch_float_datatype)
    s, _, home = self._query(tm)
    return s.to(ti.device if isinstance(ti, torch.Tensor) else home), self.rdt / float(time_max)

  def line_integrate(self, fp, ti, time_max=None):
    if time_max is None:
      return self._integrate(fp, ti, None, False)
    eta, div = self._tmax_tables(ti, time_max)
    return self._integrate(fp, ti, div, False, eta_override=eta)

  def line_integrate_multi(self, fp, ti, time_max=None):
    return self.line_integrate(fp, ti, time_max)

  def forward(self, fs, ti, time_max=None):
    s, _ = self.compute_s(ti, time_max)
    return self.line_integrate(fs(s), ti, time_max)
