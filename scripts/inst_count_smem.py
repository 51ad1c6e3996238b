"""Three launches of the batched kernel (sweeps only / + amplitude move / + measurement) for instruction counting under ncu:
   ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum -k regex:sweep_smem python scripts/inst_count_smem.py [B] [sweeps] [Z]
   Z given: every replica has that many rows (T = 50); otherwise the bench's parameter grid."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cylinder_b200 import _abi
if os.environ.get("CYL_LIB"):
    _abi.LIB_PATH = os.path.abspath(os.environ["CYL_LIB"])      # a variant built by scripts/build_variant.sh
from cylinder_b200.replicas import ReplicaBatch
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
nsw = int(sys.argv[2]) if len(sys.argv) > 2 else 20
par = bench.grid_params(B, 0)
if len(sys.argv) > 3:
    par["wavenumber"] = 50.0 / int(sys.argv[3])
batch = ReplicaBatch(par, dims=(50, 50), dtype=torch.complex64, device=torch.device("cuda", 0), seed=1)
batch.run(nsw, amp_moves=False, measure=False)
batch.run(nsw, amp_moves=True, measure=False)
batch.run(nsw)
torch.cuda.synchronize()
print("sites per sweep", batch.sites, "B", B, "sweeps", nsw, "Zmax", batch.Zmax)
