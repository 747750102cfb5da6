"""``laplace_reconstruct`` -- drop-in for ``torchlaplace.core.laplace_reconstruct``
(/root/reference/torchlaplace/core.py:25-187): same signature, argument meaning, validation
errors and output shape/dtype; the work runs as hand-written sm_100a CUDA.

Two execution paths, both CUDA, neither falls back to torch ops or to the CPU:

* fused -- ``laplace_rep_func`` has the canonical structure every reference caller uses
  (``linear_tanh_stack = Sequential(Linear, Tanh, Linear, Tanh, Linear)`` + tanh sphere heads, or
  the raw head of ``SurfaceModel``): ONE forward kernel and ONE backward kernel do s-point
  generation, the Riemann-sphere projection, the MLP, sphere->complex and the reduction over
  reconstruction terms (``nl_reconstruct_forward/backward``).
* split -- any other ``nn.Module``: kernel A (s points + projection) -> the user's module in
  torch -> kernel B (sphere->complex + reduction, forward and backward).

Extensions over the reference: ``fixed_tablot`` and ``stehfest`` work (their
``n = (terms-1)/2`` s points must match ``laplace_rep_func.s_dim``), ``dehoog`` accepts per-row
time grids, and ``use_sphere_projection=False`` accepts per-row time grids.
"""
from __future__ import annotations

import weakref

import torch
from torch import Tensor, nn

from . import _consts, _lib, _ops
from .inverse_laplace import CME, DeHoog, Euler, FixedTablot, Fourier, Stehfest

ILT_ALGORITHMS = {
    "fourier": Fourier,
    "dehoog": DeHoog,
    "cme": CME,
    "euler": Euler,
    "fixed_tablot": FixedTablot,
    "stehfest": Stehfest,
}

_canonical_cache = weakref.WeakKeyDictionary()


def _canonical_stack(f):
  """Returns the three Linear layers if ``f`` has the canonical laplace_rep_func structure."""
  stack = getattr(f, "linear_tanh_stack", None)
  if not isinstance(stack, nn.Sequential) or len(stack) != 5:
    return None
  l1, a1, l2, a2, l3 = stack
  if not (isinstance(l1, nn.Linear) and isinstance(l2, nn.Linear) and isinstance(l3, nn.Linear) and
          isinstance(a1, nn.Tanh) and isinstance(a2, nn.Tanh)):
    return None
  if l1.bias is None or l2.bias is None or l3.bias is None:
    return None
  for attr in ("s_dim", "output_dim", "latent_dim"):
    if not isinstance(getattr(f, attr, None), int):
      return None
  if l1.in_features != 2 * f.s_dim + f.latent_dim or l3.out_features != 2 * f.s_dim * f.output_dim:
    return None
  if l2.in_features != l1.out_features or l2.out_features != l1.out_features or l3.in_features != l1.out_features:
    return None
  return l1, l2, l3


def _head_kind(f, layers):
  """'sphere' / 'raw' if ``f.forward`` provably equals the canonical head, else None.

  Modules from ``neurallaplace_b200.rep_func`` declare it; for user copies (the reference's
  examples paste the class) the module's own forward is probed once on a tiny batch and compared
  with both canonical heads -- a module with the same attributes but a different forward is
  therefore never silently replaced by the fused kernel.
  """
  # The declaration of the shipped classes is honoured only for the shipped forward itself: a subclass that
  # overrides forward, an instance whose phi_scale was changed (the kernel hard-codes pi) or a module with
  # forward hooks goes through the probe / the split path like any user module.
  from . import rep_func
  if f._forward_hooks or f._forward_pre_hooks or getattr(f, "_forward_hooks_with_kwargs", None):
    return None
  shipped = {rep_func.LaplaceRepresentationFunc.forward: "sphere", rep_func.SurfaceModel.forward: "raw"}
  kind = shipped.get(type(f).forward)
  if kind is not None and kind == type(f).__dict__.get("_nl_head", getattr(type(f), "_nl_head", None)):
    if kind == "raw" or getattr(f, "phi_scale", None) == torch.pi:
      return kind
  try:
    return _canonical_cache[f]
  except (KeyError, TypeError):
    pass
  l1, l2, l3 = layers
  kind = None
  try:
    with torch.no_grad():
      w = l1.weight
      nfe = getattr(f, "nfe", None)
      probe = torch.linspace(-1.0, 1.0, 3 * l1.in_features, dtype=w.dtype, device=w.device).view(3, -1)
      out = f(probe)
      if nfe is not None:
        f.nfe = nfe
      if isinstance(out, (tuple, list)) and len(out) == 2:
        o = l3(torch.tanh(l2(torch.tanh(l1(probe))))).view(-1, 2 * f.output_dim, f.s_dim)
        oa, ob = o[:, :f.output_dim, :], o[:, f.output_dim:, :]
        tol = 1e-5 if w.dtype == torch.float32 else 1e-12
        if out[0].shape == oa.shape and out[1].shape == ob.shape:
          if torch.allclose(out[0], torch.tanh(oa) * torch.pi, atol=tol) and torch.allclose(
              out[1], torch.tanh(ob) * (torch.pi / 2), atol=tol):
            kind = "sphere"
          elif torch.allclose(out[0], oa, atol=tol) and torch.allclose(out[1], ob, atol=tol):
            kind = "raw"
  except Exception:  # pylint: disable=broad-except
    kind = None
  try:
    _canonical_cache[f] = kind
  except TypeError:
    pass
  return kind


def laplace_reconstruct(
    laplace_rep_func,
    p,
    t,
    recon_dim=None,
    ilt_algorithm="fourier",
    use_sphere_projection=True,
    ilt_reconstruction_terms=33,
    options=None,
    compute_deriv=False,  # pylint: disable=unused-argument
    x0=None,  # pylint: disable=unused-argument
) -> Tensor:
  r"""Reconstruct trajectories :math:`\mathbf{x}(t)` from a Laplace representation (core.py:25-150).

  Args:
      laplace_rep_func (nn.Module): maps ``inputs[B, T, 2n+K]`` (sphere coordinates of the ``n`` query
          points followed by ``p``) to ``(theta, phi)`` (or ``(re, im)``), each ``[B*T, recon_dim, n]``.
      p (Tensor): latent initial-state encoding ``[B, K]``; its dtype selects float32 / float64.
      t (Tensor): time points ``[T]``, ``[1, T]`` (shared) or ``[B, T]`` (per trajectory).
      recon_dim (int): :math:`d_{obs}`; defaults to ``K``.
      ilt_algorithm (str): one of ``fourier, dehoog, cme, euler, fixed_tablot, stehfest``.
      use_sphere_projection (bool): Riemann-sphere in/out maps (default) or raw complex head.
      ilt_reconstruction_terms (int): reconstruction terms per time point (odd, except ``cme``).
      options (dict): forwarded to the ILT class (``alpha``, ``tol``, ``scale``, ``eps``).

  Returns:
      Tensor ``[B, T, recon_dim]`` on ``p``'s device.
  """
  options, recon_dim, p = _check_inputs(laplace_rep_func, p, t, recon_dim, ilt_algorithm, use_sphere_projection,
                                        ilt_reconstruction_terms, options)
  _lib.require_cuda()
  if t.requires_grad:
    raise NotImplementedError("gradients with respect to t are not provided by the fused kernels")
  dtype = torch.float64 if p.dtype == torch.float64 else torch.float32
  consts = _consts.get(ilt_algorithm, ilt_reconstruction_terms, dtype, options)
  n = consts.n
  if consts.family == _lib.NL_AW and consts.eta is None:
    raise ValueError("Stehfest needs an even (ilt_reconstruction_terms - 1) / 2")

  if len(t.shape) == 0:
    time_dim = 1
  elif len(t.shape) == 1:
    time_dim = t.shape[0]
  elif len(t.shape) == 2:
    time_dim = t.shape[1]
  else:
    raise ValueError("Unsupported time tensor shape, please use (batch, time_dim)")
  batch_dim = p.shape[0]
  per_row = len(t.shape) == 2 and batch_dim == t.shape[0]
  if not per_row and t.numel() != time_dim:
    raise ValueError(f"time tensor of shape {tuple(t.shape)} is neither shared (T) / (1, T) nor per-row (B, T) "
                     f"for a batch of {batch_dim}")
  t_mode = _lib.NL_T_PER_ROW if per_row else _lib.NL_T_SHARED
  head = _lib.NL_HEAD_SPHERE if use_sphere_projection else _lib.NL_HEAD_RAW

  home = p.device
  dev = home if home.type == "cuda" else torch.device("cuda", torch.cuda.current_device())
  pd = p.to(device=dev, dtype=dtype)
  td = t.detach().to(device=dev, dtype=dtype)
  td = td.reshape(batch_dim, time_dim) if per_row else td.reshape(time_dim)

  layers = _canonical_stack(laplace_rep_func)
  s_dim = getattr(laplace_rep_func, "s_dim", None)
  if isinstance(s_dim, int) and s_dim != n:
    raise ValueError(f'ilt_algorithm "{ilt_algorithm}" with {ilt_reconstruction_terms} reconstruction terms '
                     f"evaluates {n} s points per time point, but laplace_rep_func.s_dim is {s_dim}")
  if layers is not None:
    kind = _head_kind(laplace_rep_func, layers)
    want = "sphere" if use_sphere_projection else "raw"
    l1, l2, l3 = layers
    H, K = l1.out_features, pd.shape[1]
    if (kind == want and laplace_rep_func.output_dim == recon_dim and laplace_rep_func.latent_dim == K and
        l1.weight.dtype == dtype and
        _ops.fused_supported(consts, dtype, batch_dim, time_dim, K, H, recon_dim, t_mode, head, 0, dev)):
      ws = [w if w.device == dev else w.to(dev) for w in (l1.weight, l1.bias, l2.weight, l2.bias, l3.weight, l3.bias)]
      if hasattr(laplace_rep_func, "nfe"):
        laplace_rep_func.nfe += 1
      # keep the activation stash only when a backward can follow (needs_input_grad alone stays
      # True under torch.no_grad())
      stash = torch.is_grad_enabled() and (pd.requires_grad or any(w.requires_grad for w in ws))
      x = _ops.ReconstructFn.apply(pd, td, *ws, consts, recon_dim, t_mode, head, 0, stash)
      return x if home == dev else x.to(home)

  # ---- split path: kernel A -> user module -> kernel B
  _, feat = _ops.query_points(consts, td.reshape(-1), head, False, True)
  if per_row:
    feat = feat.view(batch_dim, time_dim, 2 * n)
  else:
    feat = feat.view(1, time_dim, 2 * n).expand(batch_dim, time_dim, 2 * n)
  inputs = torch.cat((feat, pd.view(batch_dim, 1, -1).expand(batch_dim, time_dim, pd.shape[1])), 2)
  a, b = laplace_rep_func(inputs)
  # like core.py:130 (`sr.view(-1, time_dim, recon_dim, s_terms_dim)`): the leading dimension is
  # whatever is left, which differs from B when the module's output_dim != recon_dim.
  a = a.reshape(-1, time_dim, recon_dim, n)
  b = b.reshape(-1, time_dim, recon_dim, n)
  if per_row and a.shape[0] != batch_dim:
    raise ValueError("laplace_rep_func output does not match (batch, time, recon_dim, terms) for per-row times")
  layout = _lib.NL_F_SPHERE_PLANAR if use_sphere_projection else _lib.NL_F_COMPLEX_PLANAR
  x = _ops.line_integrate(consts, a, b, td, None, layout, t_mode)
  return x if home == dev else x.to(home)


def _check_inputs(laplace_rep_func, p, t, recon_dim, ilt_algorithm, use_sphere_projection, ilt_reconstruction_terms,
                  options):
  # same checks, same exception types as core.py:153-187
  if not isinstance(laplace_rep_func, nn.Module):
    raise RuntimeError("laplace_rep_func must be a descendant of torch.nn.Module")
  if not isinstance(p, Tensor):
    raise RuntimeError("p must be a torch.Tensor type")
  if not isinstance(t, Tensor):
    raise RuntimeError("t must be a torch.Tensor type")
  if not isinstance(use_sphere_projection, bool):
    raise RuntimeError("use_sphere_projection must be a bool type")
  if ilt_algorithm not in ILT_ALGORITHMS:
    raise ValueError(f'Invalid ILT algorithm "{ilt_algorithm}". Must be one of "' +
                     '", "'.join(ILT_ALGORITHMS.keys()) + '".')
  if (ilt_reconstruction_terms % 2 == 0) and (ilt_algorithm != "cme"):
    raise ValueError("Invalid 'ilt_reconstruction_terms', must be an odd input number. "
                     f"Was given an even number of {ilt_reconstruction_terms}")
  options = {} if options is None else options.copy()
  if recon_dim is None:
    recon_dim = p.shape[1]
  return options, recon_dim, p
