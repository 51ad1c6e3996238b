"""Time the batched shared-memory kernel alone: python scripts/time_smem.py [B] [sweeps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
DT = torch.complex128 if os.environ.get("CYL_DTYPE") == "f64" else torch.complex64
from cylinder_b200 import _abi
if os.environ.get("CYL_LIB"):
    _abi.LIB_PATH = os.path.abspath(os.environ["CYL_LIB"])      # a variant built by scripts/build_variant.sh
from cylinder_b200.replicas import ReplicaBatch
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
nsw = int(sys.argv[2]) if len(sys.argv) > 2 else 50
dev = torch.device("cuda", 0)
batch = ReplicaBatch(bench.grid_params(B, 0), dims=(50, 50), dtype=DT, device=dev, seed=1)
batch.run(nsw)
torch.cuda.synchronize()
for kw, name in ((dict(amp_moves=False, measure=False), "sweeps only"), (dict(amp_moves=True, measure=False), "sweeps+amp"), (dict(), "sweeps+amp+measure")):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); batch.run(nsw, **kw); batch.run(nsw, **kw); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    print("%s: %.2f ms/launch, %.3e upd/s" % (name, ms, batch.sites * nsw / ms * 1e3))
