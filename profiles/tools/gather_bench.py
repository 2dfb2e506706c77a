"""NCCL collectives for the per-step result records (torchrun, one rank per GPU): gather to rank 0 vs all_gather.
    python -m torch.distributed.run --nproc-per-node 8 profiles/tools/gather_bench.py"""
import os, sys
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 2500 * 5925          # one rank's C4 block of records
send = torch.full((n,), rank, dtype=torch.uint8, device=dev)
recv = torch.empty((world, n), dtype=torch.uint8, device=dev)
def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
g = timed(lambda: dist.gather(send, list(recv.unbind(0)) if rank == 0 else None, dst=0))
ag = timed(lambda: dist.all_gather_into_tensor(recv.view(-1), send))
# gather as point-to-point batches
def p2p():
    if rank == 0:
        ops = [dist.P2POp(dist.irecv, recv[r], r) for r in range(1, world)]
    else:
        ops = [dist.P2POp(dist.isend, send, 0)]
    for w in dist.batch_isend_irecv(ops): w.wait()
pp = timed(p2p)
if rank == 0:
    print("world %d, %.1f MB per rank: gather %.3f ms, all_gather %.3f ms, batched send/recv %.3f ms" % (world, n / 1e6, g, ag, pp))
dist.barrier(); dist.destroy_process_group()
