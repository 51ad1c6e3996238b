"""What limits the end-to-end step (ReplicaBatch.step_host: pinned host lattices + chain in, sweeps, lattices + records out)?
Per rank, max over ranks: (a) the full step, (b) the copies alone (0 sweeps), (c) the sweeps alone (device-resident),
(d) the host time to ENQUEUE one step (no synchronisation), for several slice counts.

    python scripts/e2e_limits.py                                     # one GPU
    torchrun --nproc-per-node 8 scripts/e2e_limits.py                 # all ranks at once: shared host memory / PCIe root / CPU
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from cylinder_b200.replicas import ReplicaBatch

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    opts = dist.ProcessGroupNCCL.Options(); opts.is_high_priority_stream = True
    dist.init_process_group("nccl", device_id=dev, pg_options=opts)
B, nsw = 8192, 200
batch = ReplicaBatch(bench.grid_params(B, rank * B), dims=(50, 50), dtype=torch.complex64, device=dev, seed=1, replica0=rank * B)
psi_h, chain_h = batch.psi.cpu().pin_memory(), batch.chain.cpu().pin_memory()
rec = batch.record()
rec_h = torch.empty(rec.shape, dtype=rec.dtype).pin_memory()
out_h = torch.empty_like(psi_h).pin_memory()
h2d = psi_h.numel() * 8 + chain_h.numel() * 8
d2h = out_h.numel() * 8 + rec_h.numel() * 8

def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()

def timed(fn, n=5):
    fn(); sync()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    sync()
    ms = (time.perf_counter() - t0) * 1e3 / n
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms

res = {}
res["sweeps only (device-resident)"] = timed(lambda: batch.run(nsw))
for nch in (4, 16, 64):
    res["full step, %d slices" % nch] = timed(lambda: batch.step_host(psi_h, chain_h, out_h, rec_h, nsw, nchunks=nch))
    res["copies only, %d slices" % nch] = timed(lambda: batch.step_host(psi_h, chain_h, out_h, rec_h, 0, nchunks=nch))
# host enqueue cost of one step (16 slices), measured without waiting for the device
torch.cuda.synchronize()
t0 = time.perf_counter()
batch.step_host(psi_h, chain_h, out_h, rec_h, nsw, nchunks=16, wait=False)
enq = (time.perf_counter() - t0) * 1e3
torch.cuda.synchronize()
if rank == 0:
    print("world %d, %d replicas per rank, %d sweeps per step; per step and rank: %.0f MB host->device, %.0f MB device->host" % (world, B, nsw, h2d / 1e6, d2h / 1e6))
    for k, v in res.items():
        extra = ""
        if k.startswith("copies"):
            extra = "  = %.1f GB/s per rank each way (%.1f GB/s aggregate both ways)" % (max(h2d, d2h) / v / 1e6, (h2d + d2h) * world / v / 1e6)
        print("  %-34s %8.2f ms%s" % (k, v, extra))
    print("  %-34s %8.2f ms (rank 0)" % ("host enqueue of one step, 16 slices", enq))
if world > 1:
    dist.destroy_process_group()
