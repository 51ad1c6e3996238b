"""Per-phase cycle accounting of the batched kernel (needs a build with -DCYL_PHASE_TIMING)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cylinder_b200 import _abi
if os.environ.get("CYL_LIB"):
    _abi.LIB_PATH = os.path.abspath(os.environ["CYL_LIB"])
from cylinder_b200.replicas import ReplicaBatch
import bench
dev = torch.device("cuda", 0)
batch = ReplicaBatch(bench.grid_params(8192, 0), dims=(50, 50), dtype=torch.complex64, device=dev, seed=1)
batch.run(50); batch.reset_observables(); batch.run(50)
torch.cuda.synchronize()
ph = batch.obs[:, 10:16].sum(0).cpu().numpy()
names = ["proposal tables (thread 0)", "site moves incl. barriers", "energy + profiles + reduction", "decision", "rebuild", "-"]
tot = ph.sum()
for n, v in zip(names, ph):
    print("%-30s %6.1f%%  %8.0f cycles per CTA-sweep" % (n, 100 * v / tot, v / (8192 * 50)))
