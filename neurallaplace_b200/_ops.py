"""Thin Python wrappers over the C ABI + the autograd glue.

Everything here launches hand-written CUDA through ``libnl_b200.so`` on the current torch stream;
torch is only used for device memory and autograd bookkeeping.
"""
from __future__ import annotations

import ctypes
import os

import torch
from torch.autograd.function import once_differentiable

from . import _lib
from ._lib import NlDesc

_stats = {"launches": 0}


def launches():
  """Kernel launches issued through the fused entry points since import (bench.py reads this)."""
  return _stats["launches"]


def _desc(consts, dtype, B=0, T=0, K=1, H=1, D=1, t_mode=0, head=0, impl=0, n=None):
  d = NlDesc()
  d.B, d.T, d.K, d.H, d.D = int(B), int(T), int(K), int(H), int(D)
  d.n = int(consts.n if n is None else n)
  d.family = consts.family
  d.t_mode = t_mode
  d.head = head
  d.dtype = _lib.dtype_code(dtype)
  # NL_B200_IMPL=simt|tc forces a kernel family (benchmarks / debugging); default: library picks
  d.impl = impl or {"auto": 0, "simt": 1, "tc": 2}.get(os.environ.get("NL_B200_IMPL", "auto"), 0)
  d.alpha, d.log_tol, d.scale = consts.alpha, consts.log_tol, consts.scale
  return d


def _workspace(lib, d, which, device):
  with _lib.on(device):
    nbytes = lib.nl_workspace_bytes(ctypes.byref(d), which)
  return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device), nbytes


class _Plan:
  """What a fused call needs besides its tensors, for one (constants, dtype, shape, mode, device) combination:
  the descriptor, the workspace sizes of both passes, the stash size and whether the fused kernels cover the shape.
  The library's answers are pure functions of the descriptor (and of the device's SM count), so they are asked once:
  at the reference demos' size (B=128, T=100) the Python layer, not the GPU, sets the step time."""
  __slots__ = ("d", "ws_fwd", "ws_bwd", "saved_bytes", "supported")


_plans = {}
_NO_PLAN_CACHE = os.environ.get("NL_B200_NO_PLAN_CACHE", "0") == "1"  # (A/B of the host path)


def _plan(consts, dtype, B, T, K, H, D, t_mode, head, impl, device):
  env = (os.environ.get("NL_B200_IMPL", "auto"), os.environ.get("NL_B200_STASH", "1"))
  key = (id(consts), dtype, B, T, K, H, D, t_mode, head, impl, device.index, env)
  hit = _plans.get(key)
  if hit is not None and hit[0] is consts and not _NO_PLAN_CACHE:
    return hit[1]
  lib = _lib.load()
  pl = _Plan()
  pl.d = _desc(consts, dtype, B, T, K, H, D, t_mode, head, impl)
  with _lib.on(device):
    pl.ws_fwd = int(lib.nl_workspace_bytes(ctypes.byref(pl.d), 0))
    pl.ws_bwd = int(lib.nl_workspace_bytes(ctypes.byref(pl.d), 1))
    pl.saved_bytes = int(lib.nl_saved_bytes(ctypes.byref(pl.d))) if env[1] != "0" else 0
    pl.supported = (lib.nl_selected_impl(ctypes.byref(pl.d), 0) != 0 and
                    lib.nl_selected_impl(ctypes.byref(pl.d), 1) != 0)
  if len(_plans) > 256:
    _plans.clear()
  _plans[key] = (consts, pl)  # (the constants object is kept alive so that its id cannot be reused)
  return pl


# ------------------------------------------------------------------------------------------
# fused reconstruct
# ------------------------------------------------------------------------------------------
def launch_forward(lib, d, pc, tc, weights, beta, eta, x, ws, nbytes, saved):
  """nl_reconstruct_forward[_save] on the current stream of the tensors' device.  Allocation-free and
  stream-ordered: also what engine.GraphedReconstruct captures into a CUDA graph."""
  dev = pc.device
  with _lib.on(dev):
    if saved is not None:
      st = lib.nl_reconstruct_forward_save(ctypes.byref(d), _lib.ptr(pc), _lib.ptr(tc), _lib.ptr_array(weights),
                                           _lib.ptr(beta), _lib.ptr(eta), _lib.ptr(x), _lib.ptr(saved),
                                           saved.numel(), _lib.ptr(ws), nbytes, _lib.stream_ptr(dev))
    else:
      st = lib.nl_reconstruct_forward(ctypes.byref(d), _lib.ptr(pc), _lib.ptr(tc), _lib.ptr_array(weights),
                                      _lib.ptr(beta), _lib.ptr(eta), _lib.ptr(x), _lib.ptr(ws), nbytes,
                                      _lib.stream_ptr(dev))
  _lib.check(st)
  _stats["launches"] += lib.nl_last_launch_count()


def launch_backward(lib, d, pc, tc, weights, beta, eta, gxc, grads, ws, nbytes, saved):
  """nl_reconstruct_backward[_saved]; ``grads`` = [dW1, db1, dW2, db2, dW3, db3, dp], every one overwritten."""
  dev = pc.device
  with _lib.on(dev):
    if saved is not None:
      st = lib.nl_reconstruct_backward_saved(ctypes.byref(d), _lib.ptr(pc), _lib.ptr(tc), _lib.ptr_array(weights),
                                             _lib.ptr(beta), _lib.ptr(eta), _lib.ptr(gxc), _lib.ptr(saved),
                                             saved.numel(), _lib.ptr_array(grads), _lib.ptr(ws), nbytes,
                                             _lib.stream_ptr(dev))
    else:
      st = lib.nl_reconstruct_backward(ctypes.byref(d), _lib.ptr(pc), _lib.ptr(tc), _lib.ptr_array(weights),
                                       _lib.ptr(beta), _lib.ptr(eta), _lib.ptr(gxc), _lib.ptr_array(grads),
                                       _lib.ptr(ws), nbytes, _lib.stream_ptr(dev))
  _lib.check(st)
  _stats["launches"] += lib.nl_last_launch_count()


class ReconstructFn(torch.autograd.Function):
  """x = laplace_reconstruct(canonical MLP, p, t) as ONE fused forward and ONE fused backward."""

  @staticmethod
  def forward(ctx, p, t, W1, b1, W2, b2, W3, b3, consts, D, t_mode, head, impl, stash=True):
    lib = _lib.load()
    dev = p.device
    B, K = p.shape
    T = t.shape[-1] if t_mode == _lib.NL_T_PER_ROW else t.numel()
    H = W1.shape[0]
    pl = _plan(consts, p.dtype, B, T, K, H, D, t_mode, head, impl, dev)
    d, nbytes = pl.d, pl.ws_fwd
    ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=dev)
    beta, eta = consts.tables_on(dev)
    weights = [w if w.is_contiguous() else w.contiguous() for w in (W1, b1, W2, b2, W3, b3)]
    x = torch.empty((B, T, D), dtype=p.dtype, device=dev)
    pc, tc = p.contiguous(), t.contiguous()
    # Training: keep the two hidden activations for backward (the analogue of autograd's saved tensors,
    # 512 B per (b,t) point) so that backward does not recompute layers 1-2.  NL_B200_STASH=0 disables it.
    # ``stash`` is decided by the caller (grad mode on AND something requires grad): ctx.needs_input_grad stays
    # True under torch.no_grad() and would make inference pay for the stash.  Above NL_B200_STASH_MAX_GB or when
    # the allocation fails, backward recomputes instead.
    saved, saved_bytes = None, 0
    if stash and any(ctx.needs_input_grad):
      saved_bytes = pl.saved_bytes  # (0 under NL_B200_STASH=0)
      if saved_bytes and saved_bytes <= _stash_budget(dev):
        try:
          saved = torch.empty(saved_bytes, dtype=torch.uint8, device=dev)
        except torch.OutOfMemoryError:
          saved = None
    launch_forward(lib, d, pc, tc, weights, beta, eta, x, ws, nbytes, saved)
    ctx.save_for_backward(pc, tc, *weights)
    ctx.meta = (consts, D, t_mode, head, impl)
    ctx.saved_act = saved
    return x

  @staticmethod
  @once_differentiable
  def backward(ctx, gx):
    lib = _lib.load()
    pc, tc, *weights = ctx.saved_tensors
    consts, D, t_mode, head, impl = ctx.meta
    dev = pc.device
    B, K = pc.shape
    T = tc.shape[-1] if t_mode == _lib.NL_T_PER_ROW else tc.numel()
    H = weights[0].shape[0]
    pl = _plan(consts, pc.dtype, B, T, K, H, D, t_mode, head, impl, dev)
    d, nbytes = pl.d, pl.ws_bwd
    ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=dev)
    beta, eta = consts.tables_on(dev)
    # the six MLP gradients are views of ONE flat buffer (W1|b1|W2|b2|W3|b3): the data-parallel all-reduce and
    # the fused clip + Adam step then work on it in place, with no pack / unpack copies (parallel.py, optim.py)
    flat = torch.empty(sum(w.numel() for w in weights), dtype=pc.dtype, device=dev)
    grads, off = [], 0
    for w in weights:
      grads.append(flat[off:off + w.numel()].view(w.shape))
      off += w.numel()
    del flat
    grads.append(torch.empty_like(pc))
    gxc = gx.contiguous()
    launch_backward(lib, d, pc, tc, weights, beta, eta, gxc, grads, ws, nbytes, ctx.saved_act)
    gW1, gb1, gW2, gb2, gW3, gb3, gp = grads
    return gp, None, gW1, gb1, gW2, gb2, gW3, gb3, None, None, None, None, None, None


def _stash_budget(dev):  # pylint: disable=unused-argument
  """Largest activation stash (bytes) the training forward may allocate: NL_B200_STASH_MAX_GB, default no cap
  (an allocation that fails falls back to the recomputing backward; querying free memory per call would cost a
  driver round trip on the hot path)."""
  cap = os.environ.get("NL_B200_STASH_MAX_GB")
  return int(float(cap) * (1 << 30)) if cap is not None else (1 << 62)


def fused_supported(consts, dtype, B, T, K, H, D, t_mode, head, impl=0, device=None):
  if device is None:
    device = torch.device("cuda", torch.cuda.current_device())
  return _plan(consts, dtype, B, T, K, H, D, t_mode, head, impl, device).supported


def saved_bytes(consts, dtype, B, T, K, H, D, t_mode, head, impl=0):
  """Bytes of hidden activations the training forward keeps for backward (0: backward recomputes them)."""
  if os.environ.get("NL_B200_STASH", "1") == "0":
    return 0
  lib = _lib.load()
  d = _desc(consts, dtype, B, T, K, H, D, t_mode, head, impl)
  return int(lib.nl_saved_bytes(ctypes.byref(d)))


def selected_impl(consts, dtype, B, T, K, H, D, t_mode, head, which=0, impl=0):
  lib = _lib.load()
  d = _desc(consts, dtype, B, T, K, H, D, t_mode, head, impl)
  return lib.nl_selected_impl(ctypes.byref(d), which)


# ------------------------------------------------------------------------------------------
# split path
# ------------------------------------------------------------------------------------------
def query_points(consts, t_flat, head, want_s, want_feat, Tper=None):
  """compute_s (+ sphere projection) for a flat vector of time points (no grad)."""
  lib = _lib.load()
  dev, dtype = t_flat.device, t_flat.dtype
  rows = t_flat.numel()
  d = _desc(consts, dtype, head=head)
  beta, _ = consts.tables_on(dev)
  s = torch.empty((rows, consts.n, 2), dtype=dtype, device=dev) if want_s else None
  feat = torch.empty((rows, 2 * consts.n), dtype=dtype, device=dev) if want_feat else None
  tc = t_flat.contiguous()
  Tc = Tper.contiguous() if Tper is not None else None
  with _lib.on(dev):
    st = lib.nl_query_points(ctypes.byref(d), rows, _lib.ptr(tc), _lib.ptr(Tc), _lib.ptr(beta), _lib.ptr(s),
                             _lib.ptr(feat), _lib.stream_ptr(dev))
  _lib.check(st)
  return (torch.view_as_complex(s) if want_s else None), feat


def _line_scratch(lib, d, which, device):
  """De Hoog's line integral needs scratch (planes + recurrence table): the library never allocates, so it gets a
  buffer from torch's caching allocator, registered for this host thread for the duration of the call."""
  with _lib.on(device):
    need = int(lib.nl_line_integrate_scratch_bytes(ctypes.byref(d), which))
  if need == 0:
    return None
  buf = torch.empty(need, dtype=torch.uint8, device=device)
  lib.nl_set_scratch(_lib.ptr(buf), need)
  return buf


class LineIntegrateFn(torch.autograd.Function):
  """x[B,T,D] from F[B,T,D,n] given as (theta, phi) / (re, im) planes or an interleaved complex."""

  @staticmethod
  def forward(ctx, fa, fb, t, Tper, consts, layout, t_mode, eta_override):
    lib = _lib.load()
    dev, dtype = fa.device, fa.dtype
    if layout == _lib.NL_F_COMPLEX_INTERLEAVED:
      B, T, D, n, _two = fa.shape
    else:
      B, T, D, n = fa.shape
    d = _desc(consts, dtype, B, T, 1, 1, D, t_mode, 0, 0, n=n)
    if eta_override is not None:
      eta = eta_override
    else:
      _, eta = consts.tables_on(dev)
    fac = fa.contiguous()
    fbc = fb.contiguous() if fb is not None else None
    tc = t.contiguous()
    Tc = Tper.contiguous() if Tper is not None else None
    x = torch.empty((B, T, D), dtype=dtype, device=dev)
    scratch = _line_scratch(lib, d, 0, dev)
    with _lib.on(dev):
      st = lib.nl_line_integrate_forward(ctypes.byref(d), layout, _lib.ptr(fac), _lib.ptr(fbc), _lib.ptr(tc),
                                         _lib.ptr(Tc), _lib.ptr(eta), _lib.ptr(x), _lib.stream_ptr(dev))
    lib.nl_set_scratch(None, 0)
    del scratch
    _lib.check(st)
    ctx.save_for_backward(fac, fbc, tc, Tc, eta)
    ctx.meta = (consts, layout, t_mode, d)
    return x

  @staticmethod
  @once_differentiable
  def backward(ctx, gx):
    lib = _lib.load()
    fac, fbc, tc, Tc, eta = ctx.saved_tensors
    consts, layout, t_mode, d = ctx.meta
    ga = torch.empty_like(fac)
    gb = torch.empty_like(fbc) if fbc is not None else None
    gxc = gx.contiguous()
    dev = fac.device
    scratch = _line_scratch(lib, d, 1, dev)
    with _lib.on(dev):
      st = lib.nl_line_integrate_backward(ctypes.byref(d), layout, _lib.ptr(fac), _lib.ptr(fbc), _lib.ptr(tc),
                                          _lib.ptr(Tc), _lib.ptr(eta), _lib.ptr(gxc), _lib.ptr(ga), _lib.ptr(gb),
                                          _lib.stream_ptr(dev))
    lib.nl_set_scratch(None, 0)
    del scratch
    _lib.check(st)
    return ga, gb, None, None, None, None, None, None


def line_integrate(consts, fa, fb, t, Tper, layout, t_mode, eta_override=None):
  return LineIntegrateFn.apply(fa, fb, t, Tper, consts, layout, t_mode, eta_override)


class DehoogFixedFn(torch.autograd.Function):
  """x[rows, T] for ONE De Hoog contour per row: F[rows, n, 2] (interleaved complex), t[T]
  (DeHoog.fixed_line_integrate, inverse_laplace.py:587-701; nl_dehoog_fixed_forward/backward)."""

  @staticmethod
  def forward(ctx, F, t, contour_T, gamma):
    lib = _lib.load()
    dev, dtype = F.device, F.dtype
    rows, n, _two = F.shape
    Tn = t.numel()
    Fc, tc = F.contiguous(), t.contiguous()
    x = torch.empty((rows, Tn), dtype=dtype, device=dev)
    code = _lib.dtype_code(dtype)
    with _lib.on(dev):
      need = int(lib.nl_dehoog_fixed_scratch_bytes(code, rows, n, Tn, 0))
      scratch = torch.empty(max(need, 256), dtype=torch.uint8, device=dev)
      st = lib.nl_dehoog_fixed_forward(code, rows, n, Tn, _lib.ptr(Fc), _lib.ptr(tc), contour_T, gamma, _lib.ptr(x),
                                       _lib.ptr(scratch), need, _lib.stream_ptr(dev))
    _lib.check(st)
    ctx.save_for_backward(Fc, tc)
    ctx.meta = (contour_T, gamma)
    return x

  @staticmethod
  @once_differentiable
  def backward(ctx, gx):
    lib = _lib.load()
    Fc, tc = ctx.saved_tensors
    contour_T, gamma = ctx.meta
    dev = Fc.device
    rows, n, _two = Fc.shape
    Tn = tc.numel()
    gF = torch.empty_like(Fc)
    gxc = gx.contiguous()
    code = _lib.dtype_code(Fc.dtype)
    with _lib.on(dev):
      need = int(lib.nl_dehoog_fixed_scratch_bytes(code, rows, n, Tn, 1))
      scratch = torch.empty(max(need, 256), dtype=torch.uint8, device=dev)
      st = lib.nl_dehoog_fixed_backward(code, rows, n, Tn, _lib.ptr(Fc), _lib.ptr(tc), contour_T, gamma,
                                        _lib.ptr(gxc), _lib.ptr(gF), _lib.ptr(scratch), need, _lib.stream_ptr(dev))
    _lib.check(st)
    return gF, None, None, None


def dehoog_fixed(F, t, contour_T, gamma):
  return DehoogFixedFn.apply(F, t, float(contour_T), float(gamma))


def sphere_map(which, a, b):
  """which=0: (s_re, s_im) -> (theta, phi);  which=1: (theta, phi) -> (s_re, s_im).  Flat, no grad."""
  lib = _lib.load()
  ac, bc = a.contiguous(), b.contiguous()
  o0, o1 = torch.empty_like(ac), torch.empty_like(bc)
  fn = lib.nl_complex_to_sphere if which == 0 else lib.nl_sphere_to_complex
  with _lib.on(ac.device):
    st = fn(_lib.dtype_code(ac.dtype), ac.numel(), _lib.ptr(ac), _lib.ptr(bc), _lib.ptr(o0), _lib.ptr(o1),
            _lib.stream_ptr(ac.device))
  _lib.check(st)
  return o0, o1
