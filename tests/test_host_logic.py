"""CPU: host-side logic of the drop-in (validation errors, canonical-module detection, constant
tables vs the oracle, C-ABI library loads and exports every declared symbol, sharding helpers,
2-rank gloo all-reduce).  No compute call is made without a GPU."""
import ctypes
import os
import re

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn

from neurallaplace_b200 import _consts, _lib, core, parallel
from neurallaplace_b200 import laplace_reconstruct
from neurallaplace_b200.rep_func import LaplaceRepresentationFunc, SurfaceModel
from oracle import nl_oracle as orc
from tests import helpers as hp


def test_library_loads_and_exports_every_header_symbol():
  lib = _lib.load()
  assert lib.nl_abi_version() == 3
  header = open(os.path.join(hp.ROOT, "include", "nl_b200.h")).read()
  declared = set(re.findall(r"\b(nl_[a-z_0-9]+)\s*\(", header))
  declared -= {"nl_workspace_bytes"} - set(_lib.EXPORTS)
  assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
  for name in declared:
    assert hasattr(lib, name), name
  assert lib.nl_strerror(-3).decode().startswith("shape not covered")


def test_desc_struct_layout_matches_header():
  assert ctypes.sizeof(_lib.NlDesc) == 12 * 4 + 3 * 8
  assert _lib.NlDesc.alpha.offset == 48


def test_workspace_and_impl_queries():
  lib = _lib.load()
  d = _lib.NlDesc()
  d.B, d.T, d.K, d.H, d.D, d.n = 4096, 1024, 2, 64, 1, 33
  assert lib.nl_workspace_bytes(ctypes.byref(d), 0) >= 256
  assert lib.nl_workspace_bytes(ctypes.byref(d), 1) > 148 * 12866 * 4
  assert lib.nl_selected_impl(ctypes.byref(d), 0) in (1, 2)
  d.dtype = 1
  assert lib.nl_selected_impl(ctypes.byref(d), 1) == 1  # float64 always runs on the SIMT kernels
  d.family = 7
  assert lib.nl_selected_impl(ctypes.byref(d), 0) == 0
  # NULL pointers are rejected before any CUDA call
  d.family = 0
  assert lib.nl_reconstruct_forward(ctypes.byref(d), None, None, None, None, None, None, None, 0, None) == -1


def test_dehoog_and_optimizer_queries():
  """Host-side contract of the entry points added for De Hoog and the training-step tail (no compute calls)."""
  lib = _lib.load()
  d = _lib.NlDesc()
  d.B, d.T, d.K, d.H, d.D, d.n = 4096, 1024, 2, 64, 1, 33
  d.family = _lib.NL_FOURIER if hasattr(_lib, "NL_FOURIER") else 0
  assert lib.nl_line_integrate_scratch_bytes(ctypes.byref(d), 0) == 0  # only De Hoog needs scratch
  d.family = 2  # NL_DEHOOG
  planes = 4096 * 1024 * 33 * 8
  fwd = lib.nl_line_integrate_scratch_bytes(ctypes.byref(d), 0)
  bwd = lib.nl_line_integrate_scratch_bytes(ctypes.byref(d), 1)
  assert planes <= fwd < 2 * planes and bwd > 2 * planes  # F plane; F + dL/dF planes + the recurrence's table
  # De Hoog on the tensor-core path keeps F behind the activation stash; float64 and per-row grids do not use it
  assert lib.nl_selected_impl(ctypes.byref(d), 0) == 2 and lib.nl_selected_impl(ctypes.byref(d), 1) == 2
  assert lib.nl_saved_bytes(ctypes.byref(d)) > 4096 * 1024 * (512 + 33 * 8)
  d.dtype = 1
  assert lib.nl_selected_impl(ctypes.byref(d), 1) == 1 and lib.nl_saved_bytes(ctypes.byref(d)) == 0
  d.dtype = 0
  d.n = 34  # De Hoog needs an odd number of terms
  assert lib.nl_selected_impl(ctypes.byref(d), 0) == 0
  lib.nl_set_scratch(None, 0)
  assert lib.nl_clip_adam_scratch_bytes() >= 8200
  assert lib.nl_clip_adam_step(5, 10, None, None, None, None, 1e-3, 0.9, 0.999, 1e-8, 1, 1.0, 1, None, 0, None) == -6
  assert lib.nl_clip_adam_step(0, 10, None, None, None, None, 1e-3, 0.9, 0.999, 1e-8, 0, 1.0, 1, None, 0, None) == -2
  assert lib.nl_clip_adam_step(0, 10, None, None, None, None, 1e-3, 0.9, 0.999, 1e-8, 1, 1.0, 1, None, 0, None) == -1
  assert lib.nl_clip_adam_step(0, 0, None, None, None, None, 1e-3, 0.9, 0.999, 1e-8, 1, 1.0, 1, None, 0, None) == 0


def test_check_inputs_error_contract():
  f = LaplaceRepresentationFunc(33, 1, 2)
  p, t = torch.zeros(2, 2), torch.ones(3)
  with pytest.raises(RuntimeError, match="descendant of torch.nn.Module"):
    laplace_reconstruct(lambda i: i, p, t)
  with pytest.raises(RuntimeError, match="p must be"):
    laplace_reconstruct(f, [1.0], t)
  with pytest.raises(RuntimeError, match="t must be"):
    laplace_reconstruct(f, p, [1.0])
  with pytest.raises(RuntimeError, match="bool"):
    laplace_reconstruct(f, p, t, use_sphere_projection=1)
  with pytest.raises(ValueError, match="Invalid ILT algorithm"):
    laplace_reconstruct(f, p, t, ilt_algorithm="talbot")
  with pytest.raises(ValueError, match="odd input number"):
    laplace_reconstruct(f, p, t, ilt_reconstruction_terms=32)
  assert set(core.ILT_ALGORITHMS) == {"fourier", "dehoog", "cme", "euler", "fixed_tablot", "stehfest"}


def test_no_cpu_fallback_without_cuda():
  if torch.cuda.is_available():
    pytest.skip("CUDA present")
  f = LaplaceRepresentationFunc(33, 1, 2)
  with pytest.raises(_lib.NlError, match="no CPU fallback"):
    laplace_reconstruct(f, torch.zeros(2, 2), torch.ones(3))


def test_canonical_detection_and_probe():
  f = LaplaceRepresentationFunc(33, 1, 2)
  layers = core._canonical_stack(f)
  assert layers is not None and core._head_kind(f, layers) == "sphere"
  g = SurfaceModel(33, 1, 2)
  assert core._head_kind(g, core._canonical_stack(g)) == "raw"
  # a user copy of the class (no marker) is recognised by probing its forward once
  ws = orc.make_canonical_weights(17, 2, 3)
  u = hp.make_module(ws, 17, 2, 3, user_copy=True)
  assert core._head_kind(u, core._canonical_stack(u)) == "sphere"
  assert u.nfe == 0  # the probe does not disturb the NFE counter

  class Different(LaplaceRepresentationFunc):
    _nl_head = None

    def forward(self, i):
      th, ph = super().forward(i)
      return th * 0.5, ph

  d = Different(17, 2, 3)
  assert core._head_kind(d, core._canonical_stack(d)) is None

  class NotStack(nn.Module):

    def forward(self, i):
      return i, i

  assert core._canonical_stack(NotStack()) is None


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("alg,terms", [("fourier", 33), ("dehoog", 33), ("cme", 33), ("cme", 32), ("euler", 33),
                                       ("gaver", 33), ("fixed_tablot", 33), ("stehfest", 33), ("euler", 65),
                                       ("fixed_tablot", 65), ("stehfest", 65)])
def test_constant_tables_match_oracle_bit_for_bit(alg, terms, dtype):
  c = _consts.get(alg, terms, dtype)
  o = orc.IltConsts(alg, terms, dtype)
  assert c.n == o.n
  if alg in ("fourier", "dehoog"):
    assert c.alpha == float(o.alpha) and c.log_tol == float(torch.log(o.tol)) and c.scale == float(o.scale)
    return
  if alg in ("cme", "euler", "gaver"):
    assert torch.equal(c.beta, o.beta) and torch.equal(c.eta, o.eta)
  elif alg == "stehfest":
    assert torch.equal(c.beta, o.sdt) and torch.equal(c.eta, (o.V * o.ln2).to(o.cdtype))
  else:
    assert torch.equal(c.beta[1:], o.sdt.to(o.cdtype)) and float(c.beta[0].real) == float(o.rdt)
  # and the fp32-tainted double constants of SURVEY App. A.0
  if alg == "fixed_tablot" and dtype == torch.float64:
    assert float(c.eta[0].real) == 0.4 * float(torch.exp(o.rdt)) / 2.0


def test_options_reach_constants():
  c = _consts.get("fourier", 17, torch.float32, dict(alpha=1e-2, tol=5e-2, scale=3.0))
  assert abs(c.alpha - 1e-2) < 1e-9 and c.scale == 3.0
  assert c.log_tol == float(torch.log(torch.tensor([5e-2], dtype=torch.float32)))
  with pytest.raises(TypeError):
    _consts.get("cme", 33, torch.float32, dict(alpha=1.0))


def test_stehfest_odd_m_is_rejected():
  from neurallaplace_b200.inverse_laplace import Stehfest
  with pytest.raises(ValueError, match="even"):
    Stehfest(35)


def test_shard_helpers():
  assert [parallel.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
  p = torch.arange(20.).view(10, 2)
  t_shared, t_rows = torch.arange(5.), torch.arange(50.).view(10, 5)
  ps, ts = parallel.shard_batch(p, t_shared, 1, 4)
  assert ps.shape == (3, 2) and ts is t_shared
  ps, ts = parallel.shard_batch(p, t_rows, 3, 4)
  assert ps.shape == (2, 2) and ts.shape == (2, 5) and float(ts[0, 0]) == 40.0


def _gloo_worker(rank, world, port, q):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  torch.manual_seed(0)
  m = LaplaceRepresentationFunc(5, 1, 2, hidden_units=8)
  for i, q_ in enumerate(m.parameters()):
    q_.grad = torch.full_like(q_, float(rank + 1) * (i + 1))
  red = parallel.GradientAllReducer(m.parameters())
  red()
  ok = all(torch.allclose(q_.grad, torch.full_like(q_, 3.0 * (i + 1))) for i, q_ in enumerate(m.parameters()))
  ok = ok and red.in_place is False  # separate .grad tensors: packed into the persistent flat buffer
  # gradients that tile ONE flat buffer (what the fused backward produces): all-reduced in place, no copies
  params = list(m.parameters())
  flat = torch.empty(sum(q_.numel() for q_ in params))
  off = 0
  for i, q_ in enumerate(params):
    q_.grad = flat[off:off + q_.numel()].view(q_.shape)
    q_.grad.fill_(float(rank + 1) * (i + 1))
    off += q_.numel()
  ptr = flat.data_ptr()
  red()
  ok = ok and red.in_place is True and params[0].grad.data_ptr() == ptr
  ok = ok and all(torch.allclose(q_.grad, torch.full_like(q_, 3.0 * (i + 1))) for i, q_ in enumerate(params))
  view = parallel.flat_gradient_view(params)
  ok = ok and view is not None and view.data_ptr() == ptr and view.numel() == flat.numel()
  params[1].grad = params[1].grad.clone()  # no longer one buffer
  ok = ok and parallel.flat_gradient_view(params) is None
  lo, hi = parallel.shard_range(7)
  q.put((rank, ok, lo, hi))
  dist.destroy_process_group()


def test_two_rank_gloo_gradient_allreduce():
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = 29000 + os.getpid() % 2000
  procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
  for pr in procs:
    pr.start()
  res = sorted(q.get(timeout=120) for _ in procs)
  for pr in procs:
    pr.join(timeout=60)
  assert res == [(0, True, 0, 4), (1, True, 4, 7)]


def test_plan_cache_matches_direct_queries_and_follows_the_environment(monkeypatch):
  """_ops._plan (what every fused call uses instead of asking the library again): same answers as the direct queries,
  one object per shape, and a different plan when NL_B200_IMPL / NL_B200_STASH change (they are part of the key)."""
  import torch
  from neurallaplace_b200 import _consts, _ops
  lib = _lib.load()
  c = _consts.get("fourier", 33, torch.float32)
  dev = torch.device("cuda", 0)  # (only its index is used: the queries are host-side)
  monkeypatch.delenv("NL_B200_IMPL", raising=False)
  monkeypatch.delenv("NL_B200_STASH", raising=False)
  monkeypatch.setattr(_lib, "on", lambda device: _lib._SAME)  # no CUDA device here: the size queries are host-side
  pl = _ops._plan(c, torch.float32, 4096, 1024, 2, 64, 1, 0, 0, 0, dev)
  d = _ops._desc(c, torch.float32, 4096, 1024, 2, 64, 1, 0, 0, 0)
  assert pl.ws_fwd == lib.nl_workspace_bytes(ctypes.byref(d), 0)
  assert pl.ws_bwd == lib.nl_workspace_bytes(ctypes.byref(d), 1)
  assert pl.saved_bytes == lib.nl_saved_bytes(ctypes.byref(d)) > 4096 * 1024 * 512
  assert pl.supported and _ops.fused_supported(c, torch.float32, 4096, 1024, 2, 64, 1, 0, 0, 0, dev)
  assert _ops._plan(c, torch.float32, 4096, 1024, 2, 64, 1, 0, 0, 0, dev) is pl
  assert _ops._plan(c, torch.float32, 4096, 512, 2, 64, 1, 0, 0, 0, dev) is not pl
  monkeypatch.setenv("NL_B200_STASH", "0")
  assert _ops._plan(c, torch.float32, 4096, 1024, 2, 64, 1, 0, 0, 0, dev).saved_bytes == 0
  monkeypatch.delenv("NL_B200_STASH")
  monkeypatch.setenv("NL_B200_IMPL", "simt")
  ps = _ops._plan(c, torch.float32, 4096, 1024, 2, 64, 1, 0, 0, 0, dev)
  assert ps is not pl and ps.d.impl == 1 and ps.saved_bytes == 0
