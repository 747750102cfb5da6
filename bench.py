#!/usr/bin/env python
"""Benchmark of the reconstruction hot path (BASELINE.json metric): ILT-reconstructed
(batch x time) points per second, forward + backward.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--alg fourier] ...
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one fused forward + one fused backward of ``laplace_reconstruct`` over one synthetic
minibatch (canonical MLP, H=64, K=2; p ~ N(0,1), shared t = linspace(20/T, 20, T), upstream
gradient g ~ N(0,1); loss = sum(g*x)); with N > 1 every rank owns B rows (weak scaling) and the
step ends with the NCCL all-reduce of the MLP gradients.  Rank 0 prints ONE JSON line.

``--impl reference`` times the CPU restatement of the reference (oracle/nl_oracle.py: the same
torch op chain as torchlaplace, checked bit-for-bit against it) on the host cores, on a bounded
sample of the same workload; it never touches the GPU.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=20)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
  ap.add_argument("--alg", default="fourier")
  ap.add_argument("--terms", type=int, default=33)
  ap.add_argument("--batch", type=int, default=4096)
  ap.add_argument("--time", type=int, default=1024)
  ap.add_argument("--dobs", type=int, default=1)
  ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
  ap.add_argument("--tmode", default="shared", choices=["shared", "per_row"])
  ap.add_argument("--kernel", default="auto", choices=["auto", "simt", "tc"])
  ap.add_argument("--cpu-sample-batch", type=int, default=256)
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                  help="weak: --batch rows per rank; strong: --batch rows in total, sharded over the ranks")
  ap.add_argument("--engine", default="eager", choices=["eager", "graph"],
                  help="eager: laplace_reconstruct + autograd (the reference-facing API); graph: engine.GraphedReconstruct "
                       "(CUDA-graph replay of the same two C-ABI calls, for the host-bound small shapes)")
  ap.add_argument("--sustain-s", type=float, default=2.0,
                  help="seconds of back-to-back steps for the sustained leg (clocks under a long run); 0 = skip")
  ap.add_argument("--no-peaks", action="store_true", help="skip the TF32 / FP64 matmul peak measurement")
  return ap.parse_args()


def workload_name(a):
  return (f"{a.alg}_B{a.batch}_T{a.time}_terms{a.terms}_D{a.dobs}_{a.dtype}_{a.tmode}_t "
          f"(BASELINE.json configs[4]: synthetic throughput sweep, canonical MLP H=64 K=2)")


def gemm_flop_per_point(H, D, n, per_row):
  # fwd+bwd = 3 * (2H^2 + 2H*2Dn) = 6H^2 + 12HDn (+ 8Hn for per-row t)
  fwd = 2 * H * H + 4 * H * D * n + (2 * H * 2 * n if per_row else 0)
  return fwd, 2 * fwd


def make_inputs(a, dtype, seed, batch=None):
  from oracle import nl_oracle as orc  # weight init only (same init as the reference's tests)
  B = a.batch if batch is None else batch
  g = torch.Generator().manual_seed(seed)
  n = orc.IltConsts(a.alg, a.terms).n
  weights = orc.make_canonical_weights(n, a.dobs, 2, H=64, dtype=dtype, seed=0)
  p = torch.randn(B, 2, generator=g).to(dtype)
  if a.tmode == "shared":
    t = torch.linspace(20.0 / a.time, 20.0, a.time).to(dtype)
  else:
    t = (torch.rand(B, a.time, generator=g) * 19.9 + 0.1).to(dtype)
  gout = torch.randn(B, a.time, a.dobs, generator=g).to(dtype)
  return n, weights, p, t, gout


# ------------------------------------------------------------------------------------------
# CPU arm: the reference's op chain (oracle port) on the host cores
# ------------------------------------------------------------------------------------------
def _reference_package():
  """The UNMODIFIED reference package staged in oracle/_ref by oracle/make_ref.py (None if absent).  Imported under
  its own name from there; the repository's `torchlaplace` drop-in at the repo root is never on this path."""
  ref_root = os.path.join(ROOT, "oracle", "_ref")
  if not os.path.exists(os.path.join(ref_root, "torchlaplace", "core.py")):
    return None
  for name in [m for m in sys.modules if m == "torchlaplace" or m.startswith("torchlaplace.")]:
    del sys.modules[name]
  sys.path.insert(0, ref_root)
  try:
    import torchlaplace
    import torchlaplace.inverse_laplace as ril
    assert os.path.abspath(torchlaplace.__file__).startswith(ref_root), torchlaplace.__file__
    ril.device = "cpu"  # the reference puts its constants on this module-global device (inverse_laplace.py:53)
    return torchlaplace
  finally:
    sys.path.remove(ref_root)


class _RefMLP(torch.nn.Module):
  """The canonical laplace_rep_func every reference caller defines (tests/test_laplace_rep_func.py:36-64)."""

  def __init__(self, weights, n, D, K):
    super().__init__()
    self.s_dim, self.output_dim, self.latent_dim = n, D, K
    H = weights[0].shape[0]
    self.linear_tanh_stack = torch.nn.Sequential(torch.nn.Linear(2 * n + K, H), torch.nn.Tanh(), torch.nn.Linear(H, H),
                                                 torch.nn.Tanh(), torch.nn.Linear(H, 2 * n * D))
    with torch.no_grad():
      for i, l in enumerate(self.linear_tanh_stack[j] for j in (0, 2, 4)):
        l.weight.copy_(weights[2 * i])
        l.bias.copy_(weights[2 * i + 1])

  def forward(self, i):
    out = self.linear_tanh_stack(i.view(-1, self.s_dim * 2 + self.latent_dim)).view(-1, 2 * self.output_dim,
                                                                                   self.s_dim)
    theta = torch.tanh(out[:, :self.output_dim, :]) * torch.pi
    phi = torch.tanh(out[:, self.output_dim:, :]) * (torch.pi / 2)
    return theta, phi


def cpu_step_rate(a, steps, warmup, sample_batch, budget_s=25.0):
  """fwd+bwd of the reference on the host cores over a bounded sample of the workload.  The real reference
  (oracle/_ref) when it is staged and can run the configuration, else the oracle port (same op chain)."""
  from oracle import nl_oracle as orc
  dtype = torch.float64 if a.dtype == "f64" else torch.float32
  cores = os.cpu_count() or 1
  torch.set_num_threads(cores)
  if a.alg == "dehoog":
    sample_batch = min(sample_batch, 8)  # the reference loops over batch rows in Python
  n, weights, p, t, gout = make_inputs(a, dtype, 0, batch=sample_batch)
  ref = None
  # the reference itself cannot run fixed_tablot / stehfest through laplace_reconstruct, nor dehoog with per-row t
  # (SURVEY App. C.1): those are timed on the port, which defines them
  if a.alg in ("fourier", "cme", "euler") or (a.alg == "dehoog" and a.tmode == "shared"):
    ref = _reference_package()
  if ref is not None:
    f = _RefMLP(weights, n, a.dobs, 2).to(dtype)
    params = list(f.parameters())
    pr = p.clone().requires_grad_(True)

    def step():
      for w in params:
        w.grad = None
      pr.grad = None
      x = ref.laplace_reconstruct(f, pr, t, recon_dim=a.dobs, ilt_algorithm=a.alg, ilt_reconstruction_terms=a.terms)
      (x * gout).sum().backward()
    kind = "reference"
  else:
    ws = [w.clone().requires_grad_(True) for w in weights]
    pr = p.clone().requires_grad_(True)

    def step():
      for w in ws:
        w.grad = None
      pr.grad = None
      x = orc.laplace_reconstruct_oracle(ws, pr, t, a.dobs, a.alg, True, a.terms)
      (x * gout).sum().backward()
    kind = "port"

  t0 = time.perf_counter()
  step()
  first = time.perf_counter() - t0
  # bounded: the whole run (warm-up + timed steps) stays within ~budget_s of CPU work
  fit = max(1, int(budget_s / max(first, 1e-3)) - 1)
  warmup = max(0, min(warmup, fit // 3) - 1)
  steps = max(1, min(steps, fit - warmup))
  for _ in range(warmup):
    step()
  t0 = time.perf_counter()
  for _ in range(steps):
    step()
  dt = (time.perf_counter() - t0) / steps
  pts = sample_batch * a.time
  what = "torchlaplace (oracle/_ref, unmodified)" if kind == "reference" else "oracle port of torchlaplace"
  return dict(rate=pts / dt, ms=dt * 1e3, cores=cores, kind=kind, steps=steps, warmup=warmup + 1,
              sample=f"B={sample_batch} rows of the workload (T={a.time}), fwd+bwd, {what}, torch CPU {cores} threads")


def run_reference(a):
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  r = cpu_step_rate(a, a.steps, a.warmup, a.cpu_sample_batch, budget_s=120.0)
  line = {
      "impl": "reference", "metric": "ILT reconstructed batch*time points/sec fwd+bwd", "value": r["rate"],
      "unit": "points/s", "n_gpus": a.gpus, "steps": r["steps"], "warmup": r["warmup"], "ms_per_step": r["ms"],
      "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": a.dtype, "data": "synthetic",
      "config": {"workload": workload_name(a),
                 "note": ("the reference's own laplace_reconstruct (oracle/_ref: the unmodified torchlaplace package) "
                          if r["kind"] == "reference" else
                          "CPU restatement of torchlaplace (oracle port, bit-exact vs the reference in the build "
                          "container) ") + "on the host cores; steps / warm-up as asked unless that exceeds ~2 minutes "
                         "of CPU work, then as many as fit"},
      "cpu_baseline": {"value": r["rate"], "unit": "points/s", "cores": r["cores"], "kind": r["kind"],
                       "sample": r["sample"]},
      "e2e": {"value": r["rate"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
      "gpu_launches": 0,
  }
  print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
  """Samples SM clock / throttle reasons of one GPU (NVML, ~2 ms period) while the timed region runs."""
  REASONS = (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("hw_thermal_slowdown", 0x40),
             ("sw_thermal_slowdown", 0x20))

  def __init__(self, index):
    super().__init__(daemon=True)
    self.index, self.samples, self.stop_flag, self.active = index, [], False, False
    self.h = None
    try:
      import pynvml
      pynvml.nvmlInit()
      self.nv = pynvml
      self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
      self.max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
    except Exception:  # pylint: disable=broad-except
      self.h = None

  def run(self):
    if self.h is None:
      return
    nv = self.nv
    while not self.stop_flag:
      if self.active:
        try:
          clk = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
          try:
            rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
          except Exception:  # pylint: disable=broad-except
            rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
          self.samples.append((clk, rs))
        except Exception:  # pylint: disable=broad-except
          pass
      time.sleep(0.002)

  def summary(self):
    sm = sorted(c for c, _ in self.samples)
    reasons = set()
    for _, rs in self.samples:
      for name, bit in self.REASONS:
        if rs & bit:
          reasons.add(name)
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": getattr(self, "max", None) if sm else None,
            "reasons": sorted(reasons), "samples": len(sm)}


def _device_sleep(cycles=int(4e6)):
  """~2 ms of device-side spinning (private torch helper; skipped if a torch build does not have it)."""
  try:
    torch.cuda._sleep(cycles)  # pylint: disable=protected-access
  except Exception:  # pylint: disable=broad-except
    pass


def measure_matmul_peaks(dev):
  """torch.matmul peaks of the precision modes MEASURED_PEAKS.json does not record (TF32, FP64), best of 5."""
  out = {}
  for name, dt, nmat, tf32 in (("tf32_tflops", torch.float32, 8192, True), ("fp64_tflops", torch.float64, 4096, False)):
    try:
      old = torch.backends.cuda.matmul.allow_tf32
      torch.backends.cuda.matmul.allow_tf32 = tf32
      x = torch.randn(nmat, nmat, dtype=dt, device=dev)
      y = torch.randn(nmat, nmat, dtype=dt, device=dev)
      best = 1e30
      for i in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(x, y)
        e1.record()
        torch.cuda.synchronize()
        if i:
          best = min(best, e0.elapsed_time(e1))
      out[name] = 2.0 * nmat**3 / (best * 1e-3) / 1e12
      torch.backends.cuda.matmul.allow_tf32 = old
      del x, y
    except Exception as ex:  # pylint: disable=broad-except
      out[name] = f"failed: {ex}"
  out["how"] = "torch.matmul 8192^3 float32 with allow_tf32 / 4096^3 float64, best of 5, CUDA events"
  return out


def run_b200(a):
  import torch.distributed as dist
  from neurallaplace_b200 import _consts, _lib, _ops, laplace_reconstruct, parallel
  from neurallaplace_b200.rep_func import LaplaceRepresentationFunc

  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
  torch.cuda.set_device(local)
  dev = torch.device("cuda", local)
  # host staging buffers on the GPU's NUMA node (matters for the end-to-end leg at N > 1)
  # (opt-in: measured on 8 B200, 16 host cores: binding made the end-to-end step slower, 2.73 vs 2.52 ms -- the
  # ranks are short of host cores, not of memory bandwidth)
  numa_bound = parallel.bind_host_to_gpu(local) if os.environ.get("NL_BENCH_NUMA", "0") == "1" else False
  if world > 1:
    import datetime
    # a mismatched collective must fail in a minute, not after NCCL's default 10 (GPU-minutes are budgeted)
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=90))
  _lib.require_cuda()
  if a.alg == "cme":  # the reference's CME table is not redistributable: synthetic table of the same schema
    from neurallaplace_b200 import _iltcme
    from oracle import nl_oracle as orc
    _iltcme.set_table(orc.synthetic_cme_table())
  dtype = torch.float64 if a.dtype == "f64" else torch.float32

  rows = a.batch  # rows this rank owns
  if a.scaling == "strong":  # the global batch is fixed and sharded over the ranks (north_star: "shards the minibatch")
    lo, hi = parallel.shard_range(a.batch, rank, world)
    rows = hi - lo
  n, weights, p_h, t_h, g_h = make_inputs(a, dtype, 1000 + rank, batch=rows)
  f = LaplaceRepresentationFunc(n, a.dobs, 2, hidden_units=64).to(dev).to(dtype)
  with torch.no_grad():
    for q, w in zip([f.linear_tanh_stack[i] for i in (0, 2, 4)], range(3)):
      q.weight.copy_(weights[2 * w])
      q.bias.copy_(weights[2 * w + 1])
  params = list(f.parameters())
  reducer = parallel.GradientAllReducer(params)
  if a.kernel != "auto":
    os.environ["NL_B200_IMPL"] = a.kernel

  p_pin, t_pin, g_pin = p_h.pin_memory(), t_h.pin_memory(), g_h.pin_memory()
  x_pin = torch.empty_like(g_h).pin_memory()
  grads_pin = torch.empty(sum(q.numel() for q in params), dtype=dtype).pin_memory()
  p_d = p_pin.to(dev).requires_grad_(True)
  t_d, g_d = t_pin.to(dev), g_pin.to(dev)
  flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

  graphed = None
  if a.engine == "graph":
    from neurallaplace_b200.engine import GraphedReconstruct
    graphed = GraphedReconstruct(f, rows, a.time, recon_dim=a.dobs, ilt_algorithm=a.alg,
                                 ilt_reconstruction_terms=a.terms, per_row_t=(a.tmode == "per_row"))
    graphed.p.copy_(p_d.detach())
    graphed.t.copy_(t_d)
    graphed.grad_x.copy_(g_d)

  def step(p_in, t_in, g_in):
    if graphed is not None:  # inputs already sit in the static buffers: two graph replays + the all-reduce
      x = graphed.forward()
      graphed.backward()
      graphed.allreduce()
      return x
    for q in params:
      q.grad = None
    p_in.grad = None
    x = laplace_reconstruct(f, p_in, t_in, recon_dim=a.dobs, ilt_algorithm=a.alg, ilt_reconstruction_terms=a.terms)
    x.backward(g_in)
    reducer()
    return x

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  for _ in range(max(a.warmup, 3)):
    step(p_d, t_d, g_d)
  barrier()

  # ---- device-resident timing (value): per-step CUDA events, L2 flushed between steps
  sampler = ClockSampler(local)
  sampler.start()
  sampler.active = True
  ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
  l0 = _ops.launches()
  barrier()
  # keep the host ahead of the device: the timed region must not contain launch latency of the Python layer
  _device_sleep()
  for s0, s1 in ev:
    flush.zero_()
    s0.record()
    step(p_d, t_d, g_d)
    s1.record()
  barrier()
  launches = _ops.launches() - l0
  if graphed is not None:  # replayed launches do not pass through the Python wrappers
    launches = graphed.launches_per_step * a.steps
  sampler.active = False
  ms_total = sum(s0.elapsed_time(s1) for s0, s1 in ev)
  clocks_timed = sampler.summary()

  # ---- sustained leg: the same step back to back for >= --sustain-s seconds (no L2 flush, no host pauses): does the
  # clock of the short timed region survive a long run?  (MEASURED_PEAKS.json: 1327 MHz under a sustained GEMM.)
  sustained = None
  if a.sustain_s > 0:
    n_sus = max(a.steps, int(a.sustain_s * 1e3 / max(ms_total / a.steps, 1e-3)) + 1)
    if world > 1:  # every rank must run the SAME number of steps (each ends with a collective)
      n_t = torch.tensor([n_sus], dtype=torch.int64, device=dev)
      dist.broadcast(n_t, src=0)
      n_sus = int(n_t.item())
    sampler.samples = []
    barrier()
    sampler.active = True
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(n_sus):
      step(p_d, t_d, g_d)
    s1.record()
    barrier()
    sampler.active = False
    sus_ms = s0.elapsed_time(s1) / n_sus
    sustained = {"steps": n_sus, "seconds": s0.elapsed_time(s1) * 1e-3, "ms_per_step": sus_ms,
                 "clocks": sampler.summary(), "note": "back-to-back steps, L2 not flushed, host launch overhead included"}
  sampler.stop_flag = True

  # ---- per-kernel timing of forward and backward (roofline of the dominant kernel)
  fwd_list, bwd_list = [], []
  reps = max(3, min(a.steps, 10))
  for _ in range(reps):
    for q in params:
      q.grad = None
    p_d.grad = None
    _device_sleep()  # host runs ahead: the events then see kernels only
    flush.zero_()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    if graphed is not None:
      graphed.forward()
      e1.record()
      graphed.backward()
    else:
      x = laplace_reconstruct(f, p_d, t_d, recon_dim=a.dobs, ilt_algorithm=a.alg, ilt_reconstruction_terms=a.terms)
      e1.record()
      x.backward(g_d)
    e2.record()
    torch.cuda.synchronize()
    fwd_list.append(e0.elapsed_time(e1))
    bwd_list.append(e1.elapsed_time(e2))
  fwd_ms = sorted(fwd_list)[len(fwd_list) // 2]  # median: a step that had to grow the allocator pool is an outlier
  bwd_ms = sorted(bwd_list)[len(bwd_list) // 2]

  # ---- end to end: host (pinned) buffers in, host buffers out, copies inside the timed region, through the
  # package's host-staged step (parallel.HostStagedStep: uploads / downloads on copy streams)
  if graphed is not None:
    from neurallaplace_b200.engine import GraphedHostStep
    staged = GraphedHostStep(graphed)
  else:
    staged = parallel.HostStagedStep(f, a.dobs, a.alg, a.terms, reducer=reducer, device=dev)

  def e2e_step():
    if graphed is None:  # the next step's upstream gradient is uploaded one step ahead (input prefetch)
      return staged(p_pin, t_pin, g_pin, x_pin, grads_pin, next_grad_x_host=g_pin)
    return staged(p_pin, t_pin, g_pin, x_pin, grads_pin)

  # The copy streams grow their own allocator pools while the host runs ahead of the device: a cudaMalloc inside a
  # timed span stalls the host until the queued steps have drained (measured: one 20 MB cudaMalloc = 60-80 ms, a
  # 2-10x span).  Warm up with as many un-synchronised steps as the timed spans issue, twice, so the pools reach
  # their steady-state size first; span_diag still reports any cudaMalloc that happens inside a span.
  for _ in range(2):
    for _ in range(3 * max(3, min(a.steps, 10))):
      e2e_step()
    torch.cuda.synchronize()
  barrier()
  e2e_steps = max(3, min(a.steps, 10))
  # five spans of e2e_steps steps, the median span is reported (all five are in the JSON line): this leg is bound by
  # the host -- the Python layer needs about as long to issue a step (~60 calls) as the GPU to run it -- so a host
  # hiccup inside a span shows up in full (one span in ~10 was 2x, rarely 8x, the others of the same run normal)
  e2e_spans, e2e_diag = [], []
  for _ in range(5):
    ms0 = torch.cuda.memory_stats(dev)
    t_host0 = time.perf_counter()
    es, ee = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    es.record()
    for _ in range(e2e_steps):
      e2e_step()
    torch.cuda.current_stream().wait_stream(staged.download)  # the last x download belongs to the timed region
    ee.record()
    barrier()
    e2e_spans.append(es.elapsed_time(ee) / e2e_steps)
    ms1 = torch.cuda.memory_stats(dev)
    e2e_diag.append({"host_s": round(time.perf_counter() - t_host0, 5),
                     "cudaMalloc": ms1.get("num_device_alloc", 0) - ms0.get("num_device_alloc", 0),
                     "cudaFree": ms1.get("num_device_free", 0) - ms0.get("num_device_free", 0),
                     "alloc_retries": ms1.get("num_alloc_retries", 0) - ms0.get("num_alloc_retries", 0),
                     "reserved_gb": round(ms1.get("reserved_bytes.all.current", 0) / 2**30, 2)})
  e2e_ms = sorted(e2e_spans)[len(e2e_spans) // 2]
  h2d = p_pin.numel() * p_pin.element_size() + t_pin.numel() * t_pin.element_size() + g_pin.numel() * g_pin.element_size()
  d2h = x_pin.numel() * x_pin.element_size() + grads_pin.numel() * grads_pin.element_size()

  # ---- max over ranks
  stats = torch.tensor([ms_total, e2e_ms, fwd_ms, bwd_ms, sustained["ms_per_step"] if sustained else 0.0],
                       dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(stats, op=dist.ReduceOp.MAX)
  ms_total, e2e_ms, fwd_ms, bwd_ms, sus_ms = stats.tolist()
  pts_step = a.batch * a.time * (world if a.scaling == "weak" else 1)
  if sustained:
    sustained["ms_per_step"] = sus_ms
    sustained["value"] = pts_step / (sus_ms * 1e-3)
  value = pts_step * a.steps / (ms_total * 1e-3)
  e2e_value = pts_step / (e2e_ms * 1e-3)

  if rank == 0:
    peaks = {}
    try:
      peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # pylint: disable=broad-except
      pass
    peak_tf = float(peaks.get("bf16_tflops", 1590.0))
    peak_src = "measured bf16 burst (MEASURED_PEAKS.json)" if peaks else "fallback 1.59 PF (B200_PROFILING.md)"
    ffl, bfl = gemm_flop_per_point(64, a.dobs, n, a.tmode == "per_row")
    pts_rank = rows * a.time
    dom_ms, dom_fl, dom = (bwd_ms, bfl, "backward") if bwd_ms >= fwd_ms else (fwd_ms, ffl, "forward")
    achieved = dom_fl * pts_rank / (dom_ms * 1e-3) / 1e12
    tm = 1 if a.tmode == "per_row" else 0
    cst = _consts.get(a.alg, a.terms, dtype)
    impl_f = _ops.selected_impl(cst, dtype, rows, a.time, 2, 64, a.dobs, tm, 0, 1)
    # streaming share: x out / grad_x in (4*D bytes per point each) + the activation stash the training forward
    # writes and the backward reads back (512 B per point when the saved backward is used)
    stash = _ops.saved_bytes(cst, dtype, rows, a.time, 2, 64, a.dobs, tm, 0)
    peak_hbm = float(peaks.get("hbm_gbs", 6554.6))
    elt = 8 if a.dtype == "f64" else 4
    hbm_fwd = (elt * a.dobs * pts_rank + stash) / (fwd_ms * 1e-3) / 1e9
    hbm_bwd = (elt * a.dobs * pts_rank + stash) / (bwd_ms * 1e-3) / 1e9
    # precision mode of the tensor path: float32 parity needs every GEMM as bf16 x 3 passes (hi.hi + hi.lo + lo.hi),
    # so the tensor pipe does 3x the algorithmic FLOPs; f64 has no tcgen05 kind and runs on the CUDA cores (DFMA)
    if impl_f == 2:
      mode = {"name": "bf16x3 (float32 parity: 3 bf16 passes per GEMM, fp32 accumulate in TMEM)", "passes": 3,
              "frac_of_mode_peak": 3 * achieved / peak_tf}
    else:
      mode = {"name": "CUDA-core FFMA (float32)" if a.dtype == "f32" else
              "FP64 tensor pipe (DMMA mma.sync.m8n8k4) for the GEMMs, DFMA epilogue (float64)", "passes": 1,
              "frac_of_mode_peak": None}
    # transcendental (MUFU) share.  Algorithmic count, SURVEY section 8(d): 2H hidden tanh + per (d, k) two head tanh
    # + tan, sin, cos = 2H + 5 D n per point forward (293 at d_obs 1 / 33 terms).  Implementation count of the
    # tcgen05 kernels: 1.5 per hidden tanh (ex2 each, one rcp per pair), 5 per head pair (ex2 x2, sin, cos and one
    # rcp per two tanh pairs + one per two tangents): 3H + 5 D n forward; the saved backward re-evaluates the heads.
    mufu = None
    if impl_f == 2 and a.tmode == "shared" and a.alg != "dehoog":
      clk = clocks_timed.get("sm_mhz") or 1965
      mufu_peak = 148 * 16 * clk * 1e6
      alg_f = 2 * 64 + 5 * a.dobs * n
      mf, mb = 3 * 64 + 5 * a.dobs * n, 5 * a.dobs * n
      mufu = {"algorithmic_fwd_per_point": alg_f, "fwd_ops_per_point": mf, "bwd_ops_per_point": mb,
              "peak_ops_per_s": mufu_peak,
              "fwd_frac_algorithmic": alg_f * pts_rank / (fwd_ms * 1e-3) / mufu_peak,
              "fwd_frac": mf * pts_rank / (fwd_ms * 1e-3) / mufu_peak,
              "bwd_frac": mb * pts_rank / (bwd_ms * 1e-3) / mufu_peak}
    traffic = None
    try:
      tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
      key = f"{a.alg}_B{a.batch}_T{a.time}_terms{a.terms}_D{a.dobs}_{a.dtype}_{a.tmode}"
      traffic = tr.get(key, {}).get(dom)
    except Exception:  # pylint: disable=broad-except
      pass
    if a.dtype == "f64":
      # float64 runs on the FP64 tensor pipe (DMMA) / DFMA: judged against the FP64 matmul peak measured in this run
      # (SURVEY 7.2), not the bf16 one
      p64 = measure_matmul_peaks(dev).get("fp64_tflops")
      if p64:
        peak_tf, peak_src = float(p64), "FP64 matmul peak measured in this run (torch.matmul 4096^3, best of 5)"
    line = {
        "metric": "ILT reconstructed batch*time points/sec fwd+bwd", "value": value, "unit": "points/s",
        "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms_total / a.steps,
        "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": a.dtype, "data": "synthetic",
        "config": {"workload": workload_name(a), "per_gpu_batch": rows, "global_batch": rows * world if a.scaling ==
                   "weak" else a.batch, "time_points": a.time,
                   "terms": a.terms, "s_points": n, "d_obs": a.dobs, "parallelism": f"dp{world} (batch-sharded)",
                   "l2": "flushed (512 MB write) between timed steps",
                   "kernel_family": {1: "simt (CUDA-core FFMA)" if a.dtype == "f32" else "dmma (FP64 tensor pipe + DFMA)",
                                     2: "tcgen05"}.get(impl_f, "?"),
                   "host_path": ("engine.GraphedReconstruct (CUDA-graph replay of the two C-ABI calls)"
                                 if graphed is not None else "laplace_reconstruct + autograd"),
                   "host_numa_bound": bool(numa_bound), "fwd_ms": fwd_ms, "bwd_ms": bwd_ms},
        "e2e": {"value": e2e_value, "unit": "points/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms, "steps_per_span": e2e_steps, "spans": len(e2e_spans), "span_ms_per_step": e2e_spans,
                "reported": "median span", "span_diag": e2e_diag},
        "gpu_launches": launches,
        "clocks": clocks_timed,
        "sustained": sustained,
        "roofline": {"bound": "tensor", "kernel": f"fused {dom}", "achieved": achieved, "peak": peak_tf,
                     "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": traffic,
                     "peak_source": peak_src, "precision_mode": mode,
                     "algorithmic_flop_per_point": dom_fl,
                     "whole_step": {"achieved": (ffl + bfl) * pts_rank / ((fwd_ms + bwd_ms) * 1e-3) / 1e12,
                                    "frac": (ffl + bfl) * pts_rank / ((fwd_ms + bwd_ms) * 1e-3) / 1e12 / peak_tf,
                                    "frac_of_sustained_peak": (ffl + bfl) * pts_rank / ((fwd_ms + bwd_ms) * 1e-3) /
                                    1e12 / float(peaks.get("bf16_tflops_sustained", 1354.6))},
                     "streaming": {"activation_stash_bytes": stash, "fwd_gbs": hbm_fwd, "bwd_gbs": hbm_bwd,
                                   "peak_gbs": peak_hbm, "fwd_frac": hbm_fwd / peak_hbm,
                                   "bwd_frac": hbm_bwd / peak_hbm},
                     "mufu": mufu,
                     "note": ("De Hoog: the step is dominated by the per-thread quotient-difference recurrence "
                              "(dehoog_qd kernels: CUDA cores, IEEE divisions), the tensor-core MLP kernels take "
                              "~3 ms of it; see DESIGN.md section 4.9") if a.alg == "dehoog" else
                             ("backward: the tensor core is fed from shared memory (small-N SS-mode MMAs) and shares that "
                              "port with the epilogue -- ncu: shared-memory wavefronts LSU 36 % + tensor 47 % of peak, "
                              "tensor pipe active 37 %, issue slots 45 %; forward: issue slots 59 %, XU (MUFU) 53 %; "
                              "see DESIGN.md section 6 and profiles/r02c_*")},
    }
    if not a.no_peaks:  # the matmul peaks of the other precision modes, measured in the same run
      line["roofline"]["peaks_measured"] = measure_matmul_peaks(dev)
    if not a.no_cpu_baseline and world == 1:  # reported on rank 0 at N=1 only
      r = cpu_step_rate(a, 3, 1, a.cpu_sample_batch, budget_s=20.0)
      line["cpu_baseline"] = {"value": r["rate"], "unit": "points/s", "cores": r["cores"], "kind": r["kind"],
                              "sample": r["sample"]}
    print(json.dumps(line), flush=True)
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()


def main():
  a = parse()
  if a.impl == "reference":
    run_reference(a)
  else:
    run_b200(a)


if __name__ == "__main__":
  main()
