"""GPU (-m gpu): the fused clip + Adam step (nl_clip_adam_step / optim.FlatClipAdam) against the stock pair the
reference's training loops use: torch.nn.utils.clip_grad_norm_ + torch.optim.Adam (examples/simple_demo.py:349-350)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.float64, 1e-13)])
@pytest.mark.parametrize("max_norm", [1.0, 1e3, None])
def test_flat_clip_adam_matches_torch(dtype, tol, max_norm):
  if not torch.cuda.is_available():
    pytest.skip("no CUDA device")
  dev = torch.device("cuda", 0)
  from neurallaplace_b200.optim import FlatClipAdam
  from neurallaplace_b200.rep_func import LaplaceRepresentationFunc
  torch.manual_seed(3)
  ours = LaplaceRepresentationFunc(33, 2, 3).to(dev).to(dtype)
  ref = copy.deepcopy(ours)
  opt = FlatClipAdam(ours.parameters(), lr=3e-3, max_norm=max_norm)
  ropt = torch.optim.Adam(ref.parameters(), lr=3e-3)
  g = torch.Generator(device=dev).manual_seed(0)
  for it in range(6):
    opt.zero_grad()
    ropt.zero_grad()
    scale = 10.0 ** (it - 2)  # gradient norms on both sides of the clip threshold
    for q, r in zip(ours.parameters(), ref.parameters()):
      gr = torch.randn(q.shape, generator=g, device=dev, dtype=dtype) * scale
      q.grad.add_(gr)  # accumulates into the flat buffer, like autograd does
      r.grad = gr.clone()
    if max_norm is not None:
      total = torch.nn.utils.clip_grad_norm_(ref.parameters(), max_norm)
    opt.step()
    ropt.step()
    if max_norm is not None:
      assert abs(float(opt.grad_norm) - float(total)) <= 1e-5 * float(total)
    for q, r in zip(ours.parameters(), ref.parameters()):
      err = float((q.detach() - r.detach()).norm() / r.detach().norm())
      assert err <= tol, (it, err)
      if max_norm is not None:  # clip_grad_norm_ scales the gradients in place; so do we
        assert float((q.grad - r.grad).norm() / r.grad.norm()) <= 1e-6


def test_flat_clip_adam_trains_the_reconstruction():
  """A few optimisation steps through laplace_reconstruct: the loss goes down and the parameters stay views."""
  if not torch.cuda.is_available():
    pytest.skip("no CUDA device")
  dev = torch.device("cuda", 0)
  from neurallaplace_b200 import laplace_reconstruct
  from neurallaplace_b200.optim import FlatClipAdam
  from neurallaplace_b200.rep_func import LaplaceRepresentationFunc
  torch.manual_seed(0)
  f = LaplaceRepresentationFunc(33, 1, 2).to(dev)
  opt = FlatClipAdam(f.parameters(), lr=1e-2, max_norm=1.0)
  p = torch.randn(128, 2, device=dev)
  t = torch.linspace(0.1, 5.0, 50, device=dev)
  target = torch.sin(t)[None, :, None].expand(128, 50, 1)
  losses = []
  for _ in range(25):
    opt.zero_grad()
    loss = ((laplace_reconstruct(f, p, t, recon_dim=1) - target)**2).mean()
    loss.backward()
    opt.step()
    losses.append(float(loss.detach()))
  assert losses[-1] < 0.7 * losses[0], losses
  base = opt.flat_param.data_ptr()
  assert all(base <= q.data_ptr() < base + opt.flat_param.numel() * 4 for q in f.parameters())
