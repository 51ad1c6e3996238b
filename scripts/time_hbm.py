"""Time the HBM sweep kernel alone: python scripts/time_hbm.py [n] [sweeps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
DT = torch.complex128 if os.environ.get("CYL_DTYPE") == "f64" else torch.complex64
from cylinder_b200 import _abi, ops
if os.environ.get("CYL_LIB"):
    _abi.LIB_PATH = os.path.abspath(os.environ["CYL_LIB"])      # a variant built by scripts/build_variant.sh
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nsw = int(sys.argv[2]) if len(sys.argv) > 2 else 50
dev = torch.device("cuda", 0)
par = ops.make_params(dev, wavenumber=1.0, radius=1.0, alpha=-1.0, u=1.0, C=1.0, n=6.0, temperature=.01)
psi = torch.zeros(1, n, n, dtype=DT, device=dev)
ops.lattice_init(psi, torch.tensor([n], dtype=torch.int32, device=dev), .1, 1, 0)
ws = ops.HbmWorkspace(n, n, DT, dev)
ws.pack(psi[0])
chain = ops.new_chain(1, dev, amplitude=0.3)[0]
ws.sweep(par, chain, 1, 0, 10, 0, _abi.SWEEP_ADAPT_SIGMA)
torch.cuda.synchronize()
for flags, name in ((_abi.SWEEP_ADAPT_SIGMA, "sweeps only"), (3, "sweeps+amp"), (7, "sweeps+measure+amp")):
    obs = torch.zeros(16, dtype=torch.float64, device=dev); prof = torch.zeros(n, 3, dtype=torch.float64, device=dev)
    ws.sweep(par, chain, 1, 10, 5, 0, flags, obs=obs, prof=prof)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ws.sweep(par, chain, 1, 15, nsw, 0, flags, obs=obs, prof=prof); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("%d^2 %s: %.1f us/sweep, %.3e upd/s, %.1f%% of 6553.9 GB/s at 24 B/upd; acc %.3f amp acc %.0f/%.0f" % (n, name, 1e3 * ms / nsw, n * n * nsw / ms * 1e3, n * n * nsw * 24 / ms / 1e6 / 6553.9 * 100, (chain[7] / chain[6]).item(), chain[9].item(), chain[8].item()))
