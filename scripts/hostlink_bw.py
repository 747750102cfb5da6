"""Pinned host <-> device copy bandwidth with every rank of the box copying at once (torchrun), the traffic pattern of
the end-to-end leg of bench.py: 16.8 MB up and 16.8 MB down per step and rank, on two copy streams.
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 scripts/hostlink_bw.py"""
import os, time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
  dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 4096 * 1024
h_up, h_dn = torch.empty(n, dtype=torch.float32).pin_memory(), torch.empty(n, dtype=torch.float32).pin_memory()
d_up, d_dn = torch.empty(n, device="cuda"), torch.empty(n, device="cuda")
s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
res = {}
for mode in ("up", "down", "both"):
  for it in range(2):
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 40
    for _ in range(reps):
      if mode in ("up", "both"):
        with torch.cuda.stream(s_up):
          d_up.copy_(h_up, non_blocking=True)
      if mode in ("down", "both"):
        with torch.cuda.stream(s_dn):
          h_dn.copy_(d_dn, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
  res[mode] = dt
v = torch.tensor([res["up"], res["down"], res["both"]], dtype=torch.float64, device="cuda")
if world > 1:
  dist.all_reduce(v, op=dist.ReduceOp.MAX)
if rank == 0:
  mb = n * 4 / 1e6
  up, dn, both = v.tolist()
  print("ranks %d: 16.8 MB pinned copies, slowest rank: H2D alone %.3f ms (%.1f GB/s per GPU), D2H alone %.3f ms (%.1f GB/s), "
        "both at once %.3f ms per pair (%.1f GB/s each way)" % (world, up * 1e3, mb / up / 1e3, dn * 1e3, mb / dn / 1e3,
                                                             both * 1e3, mb / both / 1e3))
if world > 1:
  dist.destroy_process_group()
