"""GPU (-m gpu): engine.GraphedReconstruct (CUDA-graph replay of the fused forward / backward for fixed shapes)
gives the values of laplace_reconstruct + autograd, stays correct across replays with new inputs and updated
weights, and the gradients of the eager path tile one flat buffer (parallel.flat_gradient_view)."""
import pytest
import torch

from tests import helpers as hp

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("alg,terms,B,T,D,dtype", [("fourier", 33, 128, 100, 1, torch.float32),
                                                   ("fourier", 33, 8, 20, 2, torch.float64),
                                                   ("euler", 33, 300, 7, 2, torch.float32),
                                                   ("dehoog", 17, 64, 9, 1, torch.float32)])
def test_graphed_step_matches_the_eager_path(alg, terms, B, T, D, dtype):
  if not torch.cuda.is_available():
    pytest.skip("no CUDA device")
  dev = torch.device("cuda", 0)
  from neurallaplace_b200 import laplace_reconstruct, parallel
  from neurallaplace_b200.engine import GraphedReconstruct
  from oracle import nl_oracle as orc
  n = orc.IltConsts(alg, terms).n
  weights = orc.make_canonical_weights(n, D, 2, dtype=dtype, seed=3)
  f = hp.make_module([w.to(dev) for w in weights], n, D, 2, True, dev)
  step = GraphedReconstruct(f, B, T, recon_dim=D, ilt_algorithm=alg, ilt_reconstruction_terms=terms)
  assert step.launches_per_step >= 2
  g = torch.Generator().manual_seed(B + T)
  tol = 1e-10 if dtype == torch.float64 else 1e-6
  for it in range(3):
    p = torch.randn(B, 2, generator=g).to(dtype).to(dev)
    t = (torch.rand(T, generator=g) * 10 + 0.1).to(dtype).to(dev)
    gout = torch.randn(B, T, D, generator=g).to(dtype).to(dev)
    if it == 2:  # weights change between replays (an optimiser step): the graph reads them in place
      with torch.no_grad():
        for q in f.parameters():
          q.mul_(1.01)
    x = step.forward(p, t).clone()
    flat, dp = step.backward(gout)
    for q in f.parameters():
      q.grad = None
    pe = p.clone().requires_grad_(True)
    xe = laplace_reconstruct(f, pe, t, recon_dim=D, ilt_algorithm=alg, ilt_reconstruction_terms=terms)
    xe.backward(gout)
    # same kernels, same inputs.  (De Hoog float32 gradients are only reproducible to the recurrence's noise
    # when atomics reorder sums; everything else is deterministic.)
    gt = 5e-2 if alg == "dehoog" else tol
    assert hp.rel(x, xe) <= tol
    assert hp.rel(dp, pe.grad) <= gt
    eager_flat = parallel.flat_gradient_view(list(hp.module_params(f)))
    assert eager_flat is not None  # the eager backward's gradients tile ONE buffer
    assert hp.rel(flat, eager_flat) <= gt
  step.assign_grads()
  assert all(q.grad.data_ptr() == v.data_ptr() for q, v in zip(hp.module_params(f), step.grads))


def test_tensors_on_a_non_current_device():
  """every C-ABI call runs with the tensors' device current and on that device's stream -- a call with
  cuda:1 tensors while cuda:0 is the current device gives the values of the same call on cuda:0 (fused path,
  split path, class API, fused optimiser step).  Needs two GPUs."""
  if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
    pytest.skip("needs two CUDA devices")
  from neurallaplace_b200 import inverse_laplace as il
  from neurallaplace_b200 import laplace_reconstruct
  from neurallaplace_b200.optim import FlatClipAdam
  from oracle import nl_oracle as orc
  torch.cuda.set_device(0)
  d0, d1 = torch.device("cuda", 0), torch.device("cuda", 1)
  g = torch.Generator().manual_seed(1)
  weights = orc.make_canonical_weights(33, 1, 2, seed=5)
  p = torch.randn(160, 2, generator=g)
  t = torch.linspace(0.2, 9.0, 40)
  gout = torch.randn(160, 40, 1, generator=g)
  res = []
  for dev in (d0, d1):
    f = hp.make_module([w.to(dev) for w in weights], 33, 1, 2, True, dev)
    pd = p.to(dev).requires_grad_(True)
    x = laplace_reconstruct(f, pd, t.to(dev), recon_dim=1)
    x.backward(gout.to(dev))
    assert x.device == dev and pd.grad.device == dev
    opt = FlatClipAdam(list(f.parameters()), lr=1e-2, max_norm=1.0)
    opt.step()
    res.append((x.detach().cpu(), pd.grad.cpu(), [q.detach().cpu().clone() for q in f.parameters()]))
    s, T = il.Fourier(33)(lambda so: so / (so**2 + 1), t.to(dev)), None
    res[-1] += (s.cpu(),)
  assert torch.cuda.current_device() == 0
  (x0, g0, w0, y0), (x1, g1, w1, y1) = res
  assert torch.equal(x0, x1) and hp.rel(g1, g0) <= 1e-5 and hp.rel(y1, y0) <= 1e-6
  for a, b in zip(w1, w0):
    assert hp.rel(a, b) <= 1e-5


def test_graphed_host_step_matches_host_staged_step():
  """engine.GraphedHostStep (pinned host buffers in / out around the graph replays) against
  parallel.HostStagedStep (the eager path): same x, same flat gradient, over several steps."""
  if not torch.cuda.is_available():
    pytest.skip("no CUDA device")
  dev = torch.device("cuda", 0)
  from neurallaplace_b200 import parallel
  from neurallaplace_b200.engine import GraphedHostStep, GraphedReconstruct
  from oracle import nl_oracle as orc
  B, T = 256, 64
  weights = orc.make_canonical_weights(33, 1, 2, seed=9)
  f = hp.make_module([w.to(dev) for w in weights], 33, 1, 2, True, dev)
  n_par = sum(q.numel() for q in f.parameters())
  g = torch.Generator().manual_seed(4)
  step_g = GraphedHostStep(GraphedReconstruct(f, B, T, recon_dim=1))
  step_e = parallel.HostStagedStep(f, 1, device=dev)
  for _ in range(3):
    p = torch.randn(B, 2, generator=g).pin_memory()
    t = (torch.rand(T, generator=g) * 10 + 0.1).pin_memory()
    gx = torch.randn(B, T, 1, generator=g).pin_memory()
    xo_g, xo_e = torch.empty(B, T, 1).pin_memory(), torch.empty(B, T, 1).pin_memory()
    go_g, go_e = torch.empty(n_par).pin_memory(), torch.empty(n_par).pin_memory()
    dp = step_g(p, t, gx, xo_g, go_g)
    torch.cuda.synchronize()
    dp = dp.clone()
    pe = step_e(p, t, gx, xo_e, go_e)
    torch.cuda.synchronize()
    assert hp.rel(xo_g, xo_e) <= 1e-6
    assert hp.rel(go_g, go_e) <= 1e-5
    assert hp.rel(dp, pe.grad) <= 1e-5


def test_host_staged_step_prefetch_of_the_next_upstream_gradient():
  """parallel.HostStagedStep with next_grad_x_host (upload one step ahead): same results as without, step by step,
  including a step whose gradient was NOT the prefetched buffer (the prefetch is then dropped)."""
  if not torch.cuda.is_available():
    pytest.skip("no CUDA device")
  dev = torch.device("cuda", 0)
  from neurallaplace_b200 import parallel
  from oracle import nl_oracle as orc
  B, T = 192, 48
  weights = orc.make_canonical_weights(33, 1, 2, seed=3)
  f = hp.make_module([w.to(dev) for w in weights], 33, 1, 2, True, dev)
  n_par = sum(q.numel() for q in f.parameters())
  g = torch.Generator().manual_seed(8)
  p = torch.randn(B, 2, generator=g).pin_memory()
  t = (torch.rand(T, generator=g) * 10 + 0.1).pin_memory()
  gxs = [torch.randn(B, T, 1, generator=g).pin_memory() for _ in range(4)]
  plain, ahead = parallel.HostStagedStep(f, 1, device=dev), parallel.HostStagedStep(f, 1, device=dev)
  order = [0, 1, 3, 2]  # step 2 announces buffer 2 but step 3 then brings buffer 3: the prefetch must be ignored
  for i, k in enumerate(order):
    announced = gxs[[1, 2, 2, 0][i]]
    xo_a, xo_b = torch.empty(B, T, 1).pin_memory(), torch.empty(B, T, 1).pin_memory()
    go_a, go_b = torch.empty(n_par).pin_memory(), torch.empty(n_par).pin_memory()
    pa = plain(p, t, gxs[k], xo_a, go_a)
    torch.cuda.synchronize()
    pb = ahead(p, t, gxs[k], xo_b, go_b, next_grad_x_host=announced)
    torch.cuda.synchronize()
    assert torch.equal(xo_a, xo_b)
    assert hp.rel(go_b, go_a) <= 1e-6 and hp.rel(pb.grad, pa.grad) <= 1e-6
