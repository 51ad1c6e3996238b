"""Cost of the batched shared-memory kernel per site as a function of the replica's row count Z (T = 50):
uniform batches, one wavenumber of C3's grid each.  python scripts/time_smem_by_z.py [B] [sweeps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from cylinder_b200 import _abi
if os.environ.get("CYL_LIB"):
    _abi.LIB_PATH = os.path.abspath(os.environ["CYL_LIB"])
from cylinder_b200.replicas import ReplicaBatch
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4736
nsw = int(sys.argv[2]) if len(sys.argv) > 2 else 50
dev = torch.device("cuda", 0)
for k in np.linspace(0.5, 1.25, 16):
    p = bench.grid_params(B, 0)
    p["wavenumber"] = np.full(B, k)
    batch = ReplicaBatch(p, dims=(50, 50), dtype=torch.complex64, device=dev, seed=1)
    batch.run(nsw)
    torch.cuda.synchronize()
    out = []
    for kw in (dict(amp_moves=False, measure=False), dict()):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); batch.run(nsw, **kw); batch.run(nsw, **kw); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 2
        out.append(batch.sites * nsw / ms * 1e3)
    print("k %.3f Z %3d: sweeps only %.3e upd/s, full %.3e upd/s" % (k, int(batch.Z[0]) if hasattr(batch, "Z") else -1, out[0], out[1]))
