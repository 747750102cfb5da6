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
def cpu_step_rate(a, steps, warmup, sample_batch):
  from oracle import nl_oracle as orc
  if a.alg == "cme":
    pass  # the oracle falls back to its synthetic table
  dtype = torch.float64 if a.dtype == "f64" else torch.float32
  cores = os.cpu_count() or 1
  torch.set_num_threads(cores)
  if a.alg == "dehoog":
    sample_batch = min(sample_batch, 8)  # the reference loops over batch rows in Python
  n, weights, p, t, gout = make_inputs(a, dtype, 0, batch=sample_batch)
  ws = [w.clone().requires_grad_(True) for w in weights]
  pr = p.clone().requires_grad_(True)

  def step():
    for w in ws:
      w.grad = None
    pr.grad = None
    x = orc.laplace_reconstruct_oracle(ws, pr, t, a.dobs, a.alg, True, a.terms)
    (x * gout).sum().backward()

  for _ in range(warmup):
    step()
  t0 = time.perf_counter()
  for _ in range(steps):
    step()
  dt = (time.perf_counter() - t0) / steps
  pts = sample_batch * a.time
  return pts / dt, dt * 1e3, cores, f"B={sample_batch} rows of the workload (T={a.time}), fwd+bwd, torch CPU {cores} threads"


def run_reference(a):
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  steps = max(1, min(a.steps, 5))
  warm = max(1, min(a.warmup, 2))
  rate, ms, cores, sample = cpu_step_rate(a, steps, warm, a.cpu_sample_batch)
  line = {
      "impl": "reference", "metric": "ILT reconstructed batch*time points/sec fwd+bwd", "value": rate,
      "unit": "points/s", "n_gpus": a.gpus, "steps": steps, "warmup": warm, "ms_per_step": ms,
      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": a.dtype, "data": "synthetic",
      "config": {"workload": workload_name(a), "note": "CPU restatement of torchlaplace (oracle port, bit-exact vs "
                 "the reference in the build container); the Python reference itself cannot travel to the GPU box"},
      "cpu_baseline": {"value": rate, "unit": "points/s", "cores": cores, "kind": "port", "sample": sample},
      "e2e": {"value": rate, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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
  if world > 1:
    dist.init_process_group("nccl", device_id=dev)
  _lib.require_cuda()
  if a.alg == "cme":  # the reference's CME table is not redistributable: synthetic table of the same schema
    from neurallaplace_b200 import _iltcme
    from oracle import nl_oracle as orc
    _iltcme.set_table(orc.synthetic_cme_table())
  dtype = torch.float64 if a.dtype == "f64" else torch.float32

  n, weights, p_h, t_h, g_h = make_inputs(a, dtype, 1000 + rank)
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

  def step(p_in, t_in, g_in):
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
  sampler.active = False
  sampler.stop_flag = True
  ms_total = sum(s0.elapsed_time(s1) for s0, s1 in ev)

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
  staged = parallel.HostStagedStep(f, a.dobs, a.alg, a.terms, reducer=reducer, device=dev)

  def e2e_step():
    return staged(p_pin, t_pin, g_pin, x_pin, grads_pin)

  for _ in range(4):  # the copy streams grow their own allocator pools during the first steps
    e2e_step()
  barrier()
  e2e_steps = max(3, min(a.steps, 10))
  es, ee = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  es.record()
  for _ in range(e2e_steps):
    e2e_step()
  torch.cuda.current_stream().wait_stream(staged.download)  # the last x download belongs to the timed region
  ee.record()
  barrier()
  e2e_ms = es.elapsed_time(ee) / e2e_steps
  h2d = p_pin.numel() * p_pin.element_size() + t_pin.numel() * t_pin.element_size() + g_pin.numel() * g_pin.element_size()
  d2h = x_pin.numel() * x_pin.element_size() + grads_pin.numel() * grads_pin.element_size()

  # ---- max over ranks
  stats = torch.tensor([ms_total, e2e_ms, fwd_ms, bwd_ms], dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(stats, op=dist.ReduceOp.MAX)
  ms_total, e2e_ms, fwd_ms, bwd_ms = stats.tolist()
  pts_step = a.batch * a.time * world
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
    pts_rank = a.batch * a.time
    dom_ms, dom_fl, dom = (bwd_ms, bfl, "backward") if bwd_ms >= fwd_ms else (fwd_ms, ffl, "forward")
    achieved = dom_fl * pts_rank / (dom_ms * 1e-3) / 1e12
    impl_f = _ops.selected_impl(_consts.get(a.alg, a.terms, dtype), dtype, a.batch, a.time, 2, 64, a.dobs,
                                1 if a.tmode == "per_row" else 0, 0, 1)
    # streaming share: x out / grad_x in (4*D bytes per point each) + the activation stash the training forward
    # writes and the backward reads back (512 B per point when the saved backward is used)
    stash = _ops.saved_bytes(_consts.get(a.alg, a.terms, dtype), dtype, a.batch, a.time, 2, 64, a.dobs,
                             1 if a.tmode == "per_row" else 0, 0)
    peak_hbm = float(peaks.get("hbm_gbs", 6554.6))
    elt = 8 if a.dtype == "f64" else 4
    hbm_fwd = (elt * a.dobs * pts_rank + stash) / (fwd_ms * 1e-3) / 1e9
    hbm_bwd = (elt * a.dobs * pts_rank + stash) / (bwd_ms * 1e-3) / 1e9
    # transcendental (MUFU) share, SURVEY section 8(d): the tcgen05 kernels spend 2 MUFU ops per hidden tanh
    # (ex2 + rcp) and 7 per head pair (two tanh sharing one rcp, sin, cos, and sin / cos / rcp for the tangent);
    # the saved backward re-evaluates the heads only.  Peak = SMs x 16 MUFU lanes x SM clock.
    mufu = None
    if impl_f == 2 and a.tmode == "shared" and a.alg != "dehoog":
      clk = sampler.summary().get("sm_mhz") or 1965
      mufu_peak = 148 * 16 * clk * 1e6
      mf, mb = 4 * 64 + 7 * a.dobs * n, 7 * a.dobs * n
      mufu = {"fwd_ops_per_point": mf, "bwd_ops_per_point": mb, "peak_ops_per_s": mufu_peak,
              "fwd_frac": mf * pts_rank / (fwd_ms * 1e-3) / mufu_peak,
              "bwd_frac": mb * pts_rank / (bwd_ms * 1e-3) / mufu_peak}
    traffic = None
    try:
      tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
      key = f"{a.alg}_B{a.batch}_T{a.time}_terms{a.terms}_D{a.dobs}_{a.dtype}_{a.tmode}"
      traffic = tr.get(key, {}).get(dom)
    except Exception:  # pylint: disable=broad-except
      pass
    line = {
        "metric": "ILT reconstructed batch*time points/sec fwd+bwd", "value": value, "unit": "points/s",
        "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms_total / a.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": a.dtype, "data": "synthetic",
        "config": {"workload": workload_name(a), "per_gpu_batch": a.batch, "time_points": a.time,
                   "terms": a.terms, "s_points": n, "d_obs": a.dobs, "parallelism": f"dp{world} (batch-sharded)",
                   "l2": "flushed (512 MB write) between timed steps",
                   "kernel_family": {1: "simt (CUDA-core FFMA/DFMA)", 2: "tcgen05"}.get(impl_f, "?"),
                   "fwd_ms": fwd_ms, "bwd_ms": bwd_ms},
        "e2e": {"value": e2e_value, "unit": "points/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms},
        "gpu_launches": launches,
        "clocks": sampler.summary(),
        "roofline": {"bound": "tensor", "kernel": f"fused {dom}", "achieved": achieved, "peak": peak_tf,
                     "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": traffic,
                     "peak_source": peak_src,
                     "algorithmic_flop_per_point": dom_fl,
                     "streaming": {"activation_stash_bytes": stash, "fwd_gbs": hbm_fwd, "bwd_gbs": hbm_bwd,
                                   "peak_gbs": peak_hbm, "fwd_frac": hbm_fwd / peak_hbm,
                                   "bwd_frac": hbm_bwd / peak_hbm},
                     "mufu": mufu,
                     "note": ("De Hoog: the step is dominated by the per-thread quotient-difference recurrence "
                              "(dehoog_qd kernels: CUDA cores, IEEE divisions, 8 warps/SM), the tensor-core MLP "
                              "kernels take ~3 ms of it; see DESIGN.md section 4.9") if a.alg == "dehoog" else
                             ("co-limited by the CUDA-core epilogue (MUFU / issue slots) and shared-memory bandwidth "
                              "of the SS-mode MMAs, see DESIGN.md section 6 and profiles/")},
    }
    if not a.no_cpu_baseline and world == 1:  # reported on rank 0 at N=1 only
      rate, ms, cores, sample = cpu_step_rate(a, 2, 1, a.cpu_sample_batch)
      line["cpu_baseline"] = {"value": rate, "unit": "points/s", "cores": cores, "kind": "port", "sample": sample}
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
