"""Training-step tail on the GPU: global-norm gradient clip + Adam over ONE flat parameter buffer.

The reference's training loops finish a step with ``torch.nn.utils.clip_grad_norm_(params, max_norm)`` and
``torch.optim.Adam.step()`` (examples/simple_demo.py:349-350, experiments/utils.py:70-71): ~10 small launches per
parameter tensor.  ``FlatClipAdam`` keeps the parameters, their gradients and the Adam moments in flat buffers
(the parameters / ``.grad`` tensors become views of them, so autograd accumulates straight into the flat
gradient), all-reduces the flat gradient once when ``torch.distributed`` is initialised (the data-parallel
exchange of parallel.py), and updates everything with two kernels (``nl_clip_adam_step``).
Same arithmetic as the stock pair (Adam without amsgrad / weight decay); no CPU fallback.
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from . import _lib


class FlatClipAdam:
  """Drop-in for ``clip_grad_norm_(params, max_norm); Adam(params, lr, betas, eps).step()``.

  ``max_norm=None`` disables clipping.  ``allreduce`` (default: whenever a process group is initialised) sums
  the flat gradient over ranks before the update; ``average`` divides it by the world size.
  """

  def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, max_norm=None, allreduce=None, average=False,
               group=None):
    self.params = [q for q in params if q.requires_grad]
    if not self.params:
      raise ValueError("FlatClipAdam: no parameter requires grad")
    dev, dtype = self.params[0].device, self.params[0].dtype
    if dev.type != "cuda":
      raise _lib.NlError("FlatClipAdam needs CUDA parameters: the update is a CUDA kernel, there is no CPU fallback")
    if any(q.device != dev or q.dtype != dtype for q in self.params):
      raise ValueError("FlatClipAdam: all parameters must share one device and dtype")
    _lib.dtype_code(dtype)
    self.lr, self.betas, self.eps, self.max_norm = float(lr), (float(betas[0]), float(betas[1])), float(eps), max_norm
    self.allreduce, self.average, self.group = allreduce, average, group
    self.step_count = 0
    n = sum(q.numel() for q in self.params)
    self.flat_param = torch.empty(n, dtype=dtype, device=dev)
    self.flat_grad = torch.zeros(n, dtype=dtype, device=dev)
    self.exp_avg = torch.zeros(n, dtype=dtype, device=dev)
    self.exp_avg_sq = torch.zeros(n, dtype=dtype, device=dev)
    lib = _lib.load()
    self._scratch = torch.zeros(int(lib.nl_clip_adam_scratch_bytes()), dtype=torch.uint8, device=dev)
    off = 0
    with torch.no_grad():
      for q in self.params:
        k = q.numel()
        self.flat_param[off:off + k].copy_(q.reshape(-1))
        q.data = self.flat_param[off:off + k].view(q.shape)
        q.grad = self.flat_grad[off:off + k].view(q.shape)
        off += k

  def zero_grad(self):
    """Zero the flat gradient; every ``.grad`` stays a view of it (never set to None)."""
    self.flat_grad.zero_()
    off = 0
    for q in self.params:
      k = q.numel()
      if q.grad is None or q.grad.data_ptr() != self.flat_grad.data_ptr() + off * self.flat_grad.element_size():
        q.grad = self.flat_grad[off:off + k].view(q.shape)
      off += k

  @property
  def grad_norm(self):
    """Total gradient norm before clipping, of the last ``step`` that clipped (a device scalar)."""
    return self._scratch[8192:8200].view(torch.float64)[0]

  @torch.no_grad()
  def step(self):
    off = 0
    for q in self.params:  # a caller may have replaced .grad (e.g. set_to_none): fold it back into the flat buffer
      k = q.numel()
      view = self.flat_grad[off:off + k]
      if q.grad is None:
        view.zero_()
        q.grad = view.view(q.shape)
      elif q.grad.data_ptr() != view.data_ptr():
        view.copy_(q.grad.reshape(-1))
        q.grad = view.view(q.shape)
      off += k
    use_dist = self.allreduce if self.allreduce is not None else (dist.is_available() and dist.is_initialized())
    if use_dist and dist.get_world_size(self.group) > 1:
      dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.group)
      if self.average:
        self.flat_grad.div_(dist.get_world_size(self.group))
    self.step_count += 1
    lib = _lib.load()
    dev = self.flat_param.device
    with _lib.on(dev):
      st = lib.nl_clip_adam_step(_lib.dtype_code(self.flat_param.dtype), self.flat_param.numel(),
                                 _lib.ptr(self.flat_param), _lib.ptr(self.flat_grad), _lib.ptr(self.exp_avg),
                                 _lib.ptr(self.exp_avg_sq), self.lr, self.betas[0], self.betas[1], self.eps,
                                 self.step_count, float(self.max_norm) if self.max_norm else 0.0, 1,
                                 _lib.ptr(self._scratch), self._scratch.numel(), _lib.stream_ptr(dev))
    _lib.check(st)
