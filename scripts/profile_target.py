"""Small fixed workload for ncu: a few launches of each hot kernel (fp32).  python scripts/profile_target.py [smem|hbm|both]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cylinder_b200 import _abi, ops
from cylinder_b200.replicas import ReplicaBatch
import bench

which = sys.argv[1] if len(sys.argv) > 1 else "both"
dev = torch.device("cuda", 0)
if which in ("smem", "both"):
    B = int(os.environ.get("PROF_B", "2368"))
    batch = ReplicaBatch(bench.grid_params(B, 0), dims=(50, 50), dtype=torch.complex64, device=dev, seed=1)
    for _ in range(3):
        batch.run(10)
    torch.cuda.synchronize()
    print("smem ok", batch.observables()["site_acceptance"].mean().item())
if which in ("hbm", "both"):
    n = int(os.environ.get("PROF_N", "4096"))
    par = ops.make_params(dev, wavenumber=1.0, radius=1.0, alpha=-1.0, u=1.0, C=1.0, n=6.0, temperature=.01)
    psi = torch.zeros(1, n, n, dtype=torch.complex64, device=dev)
    ops.lattice_init(psi, torch.tensor([n], dtype=torch.int32, device=dev), .1, 1, 0)
    ws = ops.HbmWorkspace(n, n, torch.complex64, dev)
    ws.pack(psi[0])
    chain = ops.new_chain(1, dev, amplitude=0.3)[0]
    ws.sweep(par, chain, 1, 0, 4, 0, _abi.SWEEP_ADAPT_SIGMA)
    obs = torch.zeros(16, dtype=torch.float64, device=dev)
    prof = torch.zeros(n, 3, dtype=torch.float64, device=dev)
    ws.sweep(par, chain, 1, 4, 3, 0, 7, obs=obs, prof=prof)            # measurement + profiles + amplitude move every sweep
    torch.cuda.synchronize()
    print("hbm ok", chain[:8].tolist())
