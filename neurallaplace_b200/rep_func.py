"""The canonical ``laplace_rep_func`` modules the reference examples/tests/experiments define
(they are not part of the ``torchlaplace`` package itself, but every caller copies them):

* ``LaplaceRepresentationFunc`` / ``SphereSurfaceModel`` -- 3-layer tanh MLP with tanh sphere
  heads (/root/reference/tests/test_laplace_rep_func.py:9-64,
  experiments/baseline_models/neural_laplace.py:32-99);
* ``SurfaceModel`` -- same stack with a raw (re, im) head for ``use_sphere_projection=False``
  (experiments/baseline_models/neural_laplace.py:102-155).

``laplace_reconstruct`` recognises this structure (on these classes or on user copies of them)
and evaluates it inside the fused CUDA kernels; ``forward`` below is the plain definition used
by the split path and by anyone calling the module directly.
"""
from __future__ import annotations

import torch
from torch import nn


class LaplaceRepresentationFunc(nn.Module):
  _nl_head = "sphere"

  def __init__(self, s_dim, output_dim, latent_dim, hidden_units=64):
    super().__init__()
    self.s_dim = s_dim
    self.output_dim = output_dim
    self.latent_dim = latent_dim
    self.linear_tanh_stack = nn.Sequential(
        nn.Linear(s_dim * 2 + latent_dim, hidden_units),
        nn.Tanh(),
        nn.Linear(hidden_units, hidden_units),
        nn.Tanh(),
        nn.Linear(hidden_units, s_dim * 2 * output_dim),
    )
    for m in self.linear_tanh_stack.modules():
      if isinstance(m, nn.Linear):
        nn.init.xavier_uniform_(m.weight)
    self.phi_scale = torch.pi
    self.nfe = 0

  def forward(self, i):
    self.nfe += 1
    out = self.linear_tanh_stack(i.view(-1, self.s_dim * 2 + self.latent_dim)).view(-1, 2 * self.output_dim,
                                                                                   self.s_dim)
    theta = torch.tanh(out[:, :self.output_dim, :]) * torch.pi
    phi = torch.tanh(out[:, self.output_dim:, :]) * self.phi_scale / 2.0 - torch.pi / 2.0 + self.phi_scale / 2.0
    return theta, phi


SphereSurfaceModel = LaplaceRepresentationFunc


class SurfaceModel(LaplaceRepresentationFunc):
  _nl_head = "raw"

  def forward(self, i):
    self.nfe += 1
    out = self.linear_tanh_stack(i.view(-1, self.s_dim * 2 + self.latent_dim)).view(-1, 2 * self.output_dim,
                                                                                   self.s_dim)
    return out[:, :self.output_dim, :], out[:, self.output_dim:, :]
