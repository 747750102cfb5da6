"""Riemann-sphere maps -- same four functions as ``torchlaplace.transformations``
(/root/reference/torchlaplace/transformations.py:28-151), evaluated by the CUDA kernels
``nl_complex_to_sphere`` / ``nl_sphere_to_complex`` (include/nl_b200.h).  Inside
``laplace_reconstruct`` these maps are fused into the reconstruction kernels; the functions here
serve direct callers.  CPU inputs are staged to the current CUDA device and results returned on
the input's device.
"""
from __future__ import annotations

import torch

from . import _lib, _ops


def _stage(x):
  _lib.require_cuda()
  return (x if x.is_cuda else x.cuda()).contiguous()


class _SphereToComplexFn(torch.autograd.Function):

  @staticmethod
  def forward(ctx, theta, phi):
    sr, si = _ops.sphere_map(1, theta, phi)
    ctx.save_for_backward(theta, sr, si)
    return sr, si

  @staticmethod
  def backward(ctx, gsr, gsi):
    # s = r e^{i theta}, r = tan(phi/2 + pi/4):  ds/dtheta = i s ;  dr/dphi = (1 + r^2)/2
    theta, sr, si = ctx.saved_tensors
    gtheta = gsi * sr - gsr * si
    r2 = sr * sr + si * si
    gphi = (gsr * torch.cos(theta) + gsi * torch.sin(theta)) * (1 + r2) * 0.5
    return gtheta, gphi


def spherical_riemann_to_complex(theta, phi):
  """(theta, phi) -> (Re s, Im s) with s = tan(phi/2 + pi/4) e^{i theta}  (transformations.py:28-53)."""
  if phi.shape != theta.shape:
    raise ValueError("Invalid phi theta shapes")
  home = theta.device
  sr, si = _SphereToComplexFn.apply(_stage(theta), _stage(phi))
  return sr.to(home), si.to(home)


def complex_to_spherical_riemann(s_real, s_imag):
  """(Re s, Im s) -> (theta, phi) = (atan2(Im, Re), asin((|s|^2-1)/(|s|^2+1)))  (transformations.py:56-89).
  The point at infinity maps to phi = pi/2.  Not differentiable here (the reference never needs it)."""
  if s_real.shape != s_imag.shape:
    raise ValueError("Invalid s_real s_imag shapes")
  home = s_real.device
  th, ph = _ops.sphere_map(0, _stage(s_real.detach()), _stage(s_imag.detach()))
  return th.to(home), ph.to(home)


def spherical_to_complex(theta, phi):
  """Shape-preserving complex version (transformations.py:95-121)."""
  sr, si = spherical_riemann_to_complex(torch.flatten(theta), torch.flatten(phi))
  return torch.reshape(torch.complex(sr, si), theta.shape)


def complex_to_spherical(s):
  """Shape-preserving version taking a complex tensor (transformations.py:124-151)."""
  th, ph = complex_to_spherical_riemann(torch.flatten(s.real), torch.flatten(s.imag))
  return torch.reshape(th, s.real.shape), torch.reshape(ph, s.imag.shape)
