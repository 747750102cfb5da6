"""CUDA-graph replay of the fused reconstruction step for fixed shapes (the host-bound regime).

``laplace_reconstruct`` + ``Tensor.backward`` cost ~0.2 ms of Python / autograd / launch overhead per step (two
``autograd.Function`` crossings, 7 kernel launches, a dozen small allocations).  At the BASELINE size
(B=4096, T=1024) the 2.3 ms of kernels hide it; at the reference demos' size (``examples/simple_demo.py``: B=128,
T=100, 33 terms) the kernels take 0.10 ms and the step is host-bound.  The C ABI is allocation-free and
stream-ordered (``include/nl_b200.h``), so for a FIXED (batch, time, recon_dim, algorithm, terms) the forward call and
the backward call are each captured once into a ``torch.cuda.CUDAGraph`` over static buffers and replayed:

    step = GraphedReconstruct(laplace_rep_func, batch=128, time_points=100, recon_dim=1)
    x = step.forward(p, t)              # copies p, t into the static inputs, replays the forward graph
    loss_grad = 2 * (x - target) / x.numel()
    step.backward(loss_grad)            # replays the backward graph: step.flat_grad (views: step.grads), step.dp
    optimizer.step()                    # e.g. optim.FlatClipAdam built over step.flat_grad

Same kernels, same arithmetic as ``laplace_reconstruct`` (both go through ``_ops.launch_forward/backward``); the
weights are read in place from ``laplace_rep_func`` at replay time, so optimiser updates are seen.  One process per
GPU: ``allreduce()`` sums ``flat_grad`` over the ranks in place (one NCCL launch, outside the graph).
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from . import _consts, _lib, _ops
from .core import _canonical_stack, _head_kind


class GraphedReconstruct:
  """Fixed-shape fused forward / backward of ``laplace_reconstruct`` replayed from CUDA graphs."""

  def __init__(self, laplace_rep_func, batch, time_points, recon_dim=None, ilt_algorithm="fourier",
               use_sphere_projection=True, ilt_reconstruction_terms=33, options=None, per_row_t=False, device=None,
               graphs=True):
    _lib.require_cuda()
    layers = _canonical_stack(laplace_rep_func)
    want = "sphere" if use_sphere_projection else "raw"
    if layers is None or _head_kind(laplace_rep_func, layers) != want:
      raise ValueError("GraphedReconstruct needs the canonical laplace_rep_func (Linear-Tanh-Linear-Tanh-Linear "
                       "with the matching head); other modules go through laplace_reconstruct's split path")
    l1, l2, l3 = layers
    self.f = laplace_rep_func
    self.weights = [l1.weight, l1.bias, l2.weight, l2.bias, l3.weight, l3.bias]
    dev = self.weights[0].device if device is None else torch.device(device)
    if dev.type != "cuda" or any(w.device != dev or not w.is_contiguous() for w in self.weights):
      raise ValueError("GraphedReconstruct: the module's parameters must be contiguous tensors on one CUDA device")
    dtype = self.weights[0].dtype
    K = laplace_rep_func.latent_dim
    D = K if recon_dim is None else int(recon_dim)
    self.consts = _consts.get(ilt_algorithm, ilt_reconstruction_terms, dtype, options or {})
    if self.consts.family == _lib.NL_AW and self.consts.eta is None:
      raise ValueError("Stehfest needs an even (ilt_reconstruction_terms - 1) / 2")
    if laplace_rep_func.s_dim != self.consts.n or laplace_rep_func.output_dim != D:
      raise ValueError("laplace_rep_func.s_dim / output_dim do not match the ILT configuration")
    self.device, self.dtype = dev, dtype
    self.B, self.T, self.K, self.D, self.H = int(batch), int(time_points), K, D, l1.out_features
    self.t_mode = _lib.NL_T_PER_ROW if per_row_t else _lib.NL_T_SHARED
    head = _lib.NL_HEAD_SPHERE if use_sphere_projection else _lib.NL_HEAD_RAW
    self._lib = lib = _lib.load()
    self._d = d = _ops._desc(self.consts, dtype, self.B, self.T, K, self.H, D, self.t_mode, head, 0)
    with _lib.on(dev):
      if lib.nl_selected_impl(ctypes.byref(d), 0) == 0 or lib.nl_selected_impl(ctypes.byref(d), 1) == 0:
        raise ValueError("shape not covered by the fused kernels")
      self._ws_bytes = [int(lib.nl_workspace_bytes(ctypes.byref(d), w)) for w in (0, 1)]
      saved_bytes = int(lib.nl_saved_bytes(ctypes.byref(d)))
    # ---- static buffers
    self.p = torch.zeros((self.B, K), dtype=dtype, device=dev)
    self.t = torch.ones((self.B, self.T) if per_row_t else (self.T,), dtype=dtype, device=dev)
    self.x = torch.empty((self.B, self.T, D), dtype=dtype, device=dev)
    self.grad_x = torch.zeros((self.B, self.T, D), dtype=dtype, device=dev)
    self.dp = torch.empty((self.B, K), dtype=dtype, device=dev)
    self.flat_grad = torch.zeros(sum(w.numel() for w in self.weights), dtype=dtype, device=dev)
    self.grads, off = [], 0
    for w in self.weights:
      self.grads.append(self.flat_grad[off:off + w.numel()].view(w.shape))
      off += w.numel()
    self._ws = [torch.empty(max(b, 256), dtype=torch.uint8, device=dev) for b in self._ws_bytes]
    self._saved = torch.empty(saved_bytes, dtype=torch.uint8, device=dev) if saved_bytes else None
    self._beta, self._eta = self.consts.tables_on(dev)
    self._graphs = [None, None]
    self.launches_per_step = 0
    # one eager pass (kernel attribute calls, lazy module loading), then capture
    n0 = _ops.launches()
    self._launch(0)
    self._launch(1)
    self.launches_per_step = _ops.launches() - n0
    torch.cuda.synchronize(dev)
    if graphs:
      with torch.cuda.device(dev):
        for which in (0, 1):
          g = torch.cuda.CUDAGraph()
          with torch.cuda.graph(g):  # captures on a side stream; _lib.stream_ptr(dev) picks it up
            self._launch(which)
          self._graphs[which] = g

  def _launch(self, which):
    w = [q.data for q in self.weights]
    if which == 0:
      _ops.launch_forward(self._lib, self._d, self.p, self.t, w, self._beta, self._eta, self.x, self._ws[0],
                          self._ws_bytes[0], self._saved)
    else:
      _ops.launch_backward(self._lib, self._d, self.p, self.t, w, self._beta, self._eta, self.grad_x,
                           self.grads + [self.dp], self._ws[1], self._ws_bytes[1], self._saved)

  def _run(self, which):
    if self._graphs[which] is not None:
      self._graphs[which].replay()
    else:
      self._launch(which)

  def forward(self, p=None, t=None):
    """x[B, T, D] for (p, t); ``None`` keeps what the static inputs ``self.p`` / ``self.t`` hold."""
    if p is not None:
      self.p.copy_(p, non_blocking=True)
    if t is not None:
      self.t.copy_(t.reshape(self.t.shape), non_blocking=True)
    self._run(0)
    return self.x

  def backward(self, grad_x=None):
    """Gradients of sum(grad_x * x): ``flat_grad`` = dW1|db1|dW2|db2|dW3|db3 (views in ``grads``) and ``dp``."""
    if grad_x is not None:
      self.grad_x.copy_(grad_x, non_blocking=True)
    self._run(1)
    return self.flat_grad, self.dp

  def allreduce(self, group=None, average=False):
    """Data parallel: sum ``flat_grad`` over the ranks, in place (one NCCL message)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
      dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=group)
      if average:
        self.flat_grad.div_(dist.get_world_size(group))

  def assign_grads(self):
    """Point every parameter's ``.grad`` at its view of ``flat_grad`` (for stock optimisers / FlatClipAdam)."""
    for q, g in zip(self.weights, self.grads):
      q.grad = g


class GraphedHostStep:
  """One training step of the reconstruction path with HOST (pinned) operands and results, on CUDA-graph replays.

  The fixed-shape counterpart of ``parallel.HostStagedStep``: p / t / upstream gradient come from pinned host
  memory, x and the flat MLP gradient go back to it, the copies run on two copy streams next to the kernels, and
  the host issues ~10 calls per step instead of the ~60 of the eager path -- with one process per GPU and fewer
  host cores than ranks x threads the eager step becomes host-bound (8 ranks on 16 cores: 2.5 ms against 2.1 ms
  of kernels).

    main    : H2D p, t -> forward graph -----------------> backward graph -> all-reduce -> D2H flat gradient
    upload  :      H2D upstream gradient (under the forward) ---^
    download:                      D2H x (under the backward)
  """

  def __init__(self, graphed: GraphedReconstruct, group=None):
    self.g = graphed
    self.group = group
    dev = graphed.device
    self.upload = torch.cuda.Stream(dev)
    self.download = torch.cuda.Stream(dev)
    self._fwd_done = torch.cuda.Event()
    self._up_done = torch.cuda.Event()
    self._bwd_done = torch.cuda.Event()

  def __call__(self, p_host, t_host, grad_x_host, x_host_out, grads_host_out=None):
    """Asynchronous: synchronise the device before reading the host buffers."""
    g = self.g
    main = torch.cuda.current_stream(g.device)
    # the previous step's backward must be done with grad_x before it is overwritten
    self.upload.wait_event(self._bwd_done)
    with torch.cuda.stream(self.upload):
      g.grad_x.copy_(grad_x_host, non_blocking=True)
      self._up_done.record(self.upload)
    main.wait_stream(self.download)  # the previous x download has read g.x
    g.forward(p_host, t_host)
    self._fwd_done.record(main)
    self.download.wait_event(self._fwd_done)
    with torch.cuda.stream(self.download):
      x_host_out.copy_(g.x, non_blocking=True)
    main.wait_event(self._up_done)
    g.backward()
    self._bwd_done.record(main)
    g.allreduce(self.group)
    if grads_host_out is not None:
      grads_host_out[:g.flat_grad.numel()].copy_(g.flat_grad, non_blocking=True)
    return g.dp
