"""Batch-sharded data parallelism for the reconstruction path (one process per GPU).

Every (batch, time) point of ``laplace_reconstruct`` is independent (core.py:112-127 builds the
MLP input row by row), so the minibatch is sharded over ranks with NO communication in the
forward pass; ``p`` gradients stay local to the shard.  The backward pass needs exactly one
exchange: a sum all-reduce of the ``laplace_rep_func`` parameter gradients (12 866 floats = 51 KB
at d_obs=1 / 33 terms ... 289 184 floats = 1.16 MB at d_obs=16 / 129 terms), done here on ONE
flat buffer so NCCL sees a single latency-bound message over NVLink / NVSwitch.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def bind_host_to_gpu(device_index):
  """Pin the calling thread to the CPU cores next to GPU ``device_index`` (NVML's ideal-CPU affinity), so that the
  pinned staging buffers it allocates afterwards are first-touched on that GPU's NUMA node.  With one process per
  GPU and host-resident minibatches, eight ranks otherwise stage 2 x 16.8 MB per 2 ms step through whichever
  node the processes happened to start on (measured on 8 B200: end-to-end step 2.60 ms against 2.09 ms on one
  GPU).  Returns True when the affinity was applied; never raises (NVML may be absent)."""
  try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
    pynvml.nvmlDeviceSetCpuAffinity(h)
    return True
  except Exception:  # pylint: disable=broad-except
    return False


def shard_range(n_rows, rank=None, world_size=None):
  """Contiguous [lo, hi) slice of the batch owned by ``rank`` (balanced to within one row)."""
  if world_size is None:
    world_size = dist.get_world_size() if dist.is_initialized() else 1
  if rank is None:
    rank = dist.get_rank() if dist.is_initialized() else 0
  base, rem = divmod(n_rows, world_size)
  lo = rank * base + min(rank, rem)
  return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(p, t=None, rank=None, world_size=None):
  """Slice ``p`` (and a per-row ``t``; a shared ``t`` is replicated) for this rank."""
  lo, hi = shard_range(p.shape[0], rank, world_size)
  ps = p[lo:hi]
  if t is not None and t.dim() == 2 and t.shape[0] == p.shape[0] and p.shape[0] != 1:
    return ps, t[lo:hi]
  return ps, t


def flat_gradient_view(params):
  """The ONE flat tensor the gradients of ``params`` tile, or None.

  The fused backward (``_ops.ReconstructFn.backward``) writes dW1|db1|dW2|db2|dW3|db3 as views of one flat buffer
  and autograd adopts those views as ``.grad`` without copying, so the data-parallel exchange can run on the
  buffer in place.  Returns None when any ``.grad`` is missing or the gradients do not tile one storage in
  parameter order (e.g. they were accumulated over several backward passes into separate tensors)."""
  grads = [q.grad for q in params]
  if not grads or any(g is None or not g.is_contiguous() for g in grads):
    return None
  g0 = grads[0]
  base, off = g0.untyped_storage().data_ptr(), g0.storage_offset()
  first = off
  for g in grads:
    if g.untyped_storage().data_ptr() != base or g.storage_offset() != off or g.dtype != g0.dtype:
      return None
    off += g.numel()
  flat = torch.empty(0, dtype=g0.dtype, device=g0.device)
  flat.set_(g0.untyped_storage(), first, (off - first,), (1,))
  return flat


class GradientAllReducer:
  """Sum-all-reduce the gradients of a parameter list as ONE message.

  When the gradients already tile one flat buffer (what the fused backward produces, see ``flat_gradient_view``)
  the all-reduce runs on it in place: one NCCL launch per step and nothing else.  Otherwise they are packed
  into a persistent flat buffer and unpacked afterwards (2 x len(params) small copies)."""

  def __init__(self, params, group=None, average=False):
    self.params = [q for q in params if q.requires_grad]
    self.group = group
    self.average = average
    self._flat = None
    self.in_place = None  # how the last call ran (True: directly on the backward's flat buffer)

  def _buffer(self):
    n = sum(q.numel() for q in self.params)
    ref = self.params[0]
    if self._flat is None or self._flat.numel() != n or self._flat.device != ref.device or \
        self._flat.dtype != ref.dtype:
      self._flat = torch.empty(n, dtype=ref.dtype, device=ref.device)
    return self._flat

  def __call__(self):
    if not self.params:
      return
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
      return
    flat = flat_gradient_view(self.params)
    self.in_place = flat is not None
    if flat is not None:
      dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
      if self.average:
        flat.div_(dist.get_world_size(self.group))
      return
    flat = self._buffer()
    off = 0
    for q in self.params:
      g = q.grad
      k = q.numel()
      if g is None:
        flat[off:off + k].zero_()
      else:
        flat[off:off + k].copy_(g.reshape(-1))
      off += k
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
    if self.average:
      flat.div_(dist.get_world_size(self.group))
    off = 0
    for q in self.params:
      k = q.numel()
      if q.grad is None:
        q.grad = flat[off:off + k].view_as(q).clone()
      else:
        q.grad.copy_(flat[off:off + k].view_as(q))
      off += k


def allreduce_gradients(module_or_params, group=None, average=False):
  """One-shot helper: all-reduce ``.grad`` of a module's (or an iterable's) parameters."""
  params = list(module_or_params.parameters()) if isinstance(module_or_params, torch.nn.Module) else list(
      module_or_params)
  GradientAllReducer(params, group, average)()


class HostStagedStep:
  """One training step of the reconstruction path with HOST (pinned) operands and results.

  ``laplace_reconstruct`` itself takes device tensors (like the reference).  A caller whose minibatch lives in
  host memory pays three PCIe transfers per step: p / t / upstream gradient in, x and the MLP gradients out.
  This helper keeps them off the kernels' critical path with two copy streams:

    main   : H2D p, t (small) -> fused forward ------------------> fused backward -> all-reduce -> D2H grads
    upload :      H2D upstream gradient  (concurrent with forward) ----^
    download:                            D2H x (concurrent with backward)

  Buffers that cross streams are registered with the caching allocator (``record_stream``).
  """

  def __init__(self, laplace_rep_func, recon_dim, ilt_algorithm="fourier", ilt_reconstruction_terms=33,
               use_sphere_projection=True, reducer=None, device=None):
    from .core import laplace_reconstruct  # local import: parallel.py is importable without the CUDA library
    self._reconstruct = laplace_reconstruct
    self.f = laplace_rep_func
    self.kw = dict(recon_dim=recon_dim, ilt_algorithm=ilt_algorithm,
                   ilt_reconstruction_terms=ilt_reconstruction_terms, use_sphere_projection=use_sphere_projection)
    self.params = [q for q in laplace_rep_func.parameters() if q.requires_grad]
    self.reducer = reducer
    self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    self.upload = torch.cuda.Stream(self.device)
    self.download = torch.cuda.Stream(self.device)
    self._prefetched = None  # (host tensor, its device copy in flight on the upload stream)

  def prefetch(self, grad_x_host):
    """Start uploading the NEXT step's upstream gradient now (input prefetch, like a data loader's): called right
    after a step, the copy runs under that step's backward too instead of under the next forward alone.  With
    eight ranks sharing the host links a 16.8 MB upload no longer fits under the 0.7 ms forward (measured on
    8 B200: 2.73 ms per end-to-end step against 1.9 ms of kernels)."""
    with torch.cuda.stream(self.upload):
      self._prefetched = (grad_x_host, grad_x_host.to(self.device, non_blocking=True))

  def __call__(self, p_host, t_host, grad_x_host, x_host_out, grads_host_out=None, next_grad_x_host=None):
    """Runs forward + backward of sum(grad_x * x); writes x into ``x_host_out`` and the flattened MLP gradients
    into ``grads_host_out`` (both pinned).  Returns the device ``p`` (its ``.grad`` holds dL/dp).
    ``next_grad_x_host``: the upstream gradient of the FOLLOWING call, uploaded ahead of time (see ``prefetch``).
    The copies are asynchronous: synchronise (``torch.cuda.synchronize``) before reading the host buffers."""
    dev = self.device
    main = torch.cuda.current_stream(dev)
    if self._prefetched is not None and self._prefetched[0] is grad_x_host:
      g_d = self._prefetched[1]
    else:
      with torch.cuda.stream(self.upload):
        g_d = grad_x_host.to(dev, non_blocking=True)
    self._prefetched = None
    p_d = p_host.to(dev, non_blocking=True).requires_grad_(True)
    t_d = t_host.to(dev, non_blocking=True)
    for q in self.params:
      q.grad = None
    x = self._reconstruct(self.f, p_d, t_d, **self.kw)
    self.download.wait_stream(main)
    with torch.cuda.stream(self.download):
      x_host_out.copy_(x.detach(), non_blocking=True)
    x.record_stream(self.download)
    main.wait_stream(self.upload)
    g_d.record_stream(main)
    if next_grad_x_host is not None:  # issued before the backward's launches: the copy engine starts at once
      self.prefetch(next_grad_x_host)
    x.backward(g_d)
    if self.reducer is not None:
      self.reducer()
    if grads_host_out is not None:
      flat = flat_gradient_view(self.params)
      if flat is not None:  # one D2H copy of the backward's flat gradient buffer
        grads_host_out[:flat.numel()].copy_(flat, non_blocking=True)
      else:
        off = 0
        for q in self.params:
          k = q.numel()
          grads_host_out[off:off + k].copy_(q.grad.reshape(-1), non_blocking=True)
          off += k
    return p_d
