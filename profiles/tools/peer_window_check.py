"""Two (or more) ranks on ONE GPU (gloo): map rank 0's window, push, read back, time repeated pushes.
torchrun --nproc-per-node 2 profiles/tools/peer_window_check.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist
from genes2genes_b200 import distributed as D

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev_index = int(os.environ.get("LOCAL_RANK", "0")) % torch.cuda.device_count()
torch.cuda.set_device(dev_index)
dev = torch.device("cuda", dev_index)
dist.init_process_group("gloo")
n = 14_812_500
win = D.peer_window(n, dev)
assert win is not None, "window not available"
rec = torch.full((n,), rank + 1, dtype=torch.uint8, device=dev)
dist.barrier()
win.push(rec)
torch.cuda.synchronize()
dist.barrier()
if rank == 0:
    rows = win.rows()
    ok = all(bool((rows[r, :n] == r + 1).all()) for r in range(world))
    print("window rows correct:", ok, tuple(rows.shape), flush=True)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
dist.barrier()
ev0.record()
for _ in range(20):
    win.push(rec)
ev1.record()
torch.cuda.synchronize()
print("rank %d on %s: push of %.1f MB takes %.3f ms" % (rank, dev, n / 1e6, ev0.elapsed_time(ev1) / 20), flush=True)
dist.barrier()
dist.destroy_process_group()
