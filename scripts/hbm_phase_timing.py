"""Per-phase cycles of the HBM sweep kernel (needs a build with -DHBM_PHASE_TIMING): thread 0 of every CTA."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cylinder_b200 import _abi, ops
n = 4096
dev = torch.device("cuda", 0)
par = ops.make_params(dev, wavenumber=1.0, radius=1.0, alpha=-1.0, u=1.0, C=1.0, n=6.0, temperature=.01)
psi = torch.zeros(1, n, n, dtype=torch.complex64, device=dev)
ops.lattice_init(psi, torch.tensor([n], dtype=torch.int32, device=dev), .1, 1, 0)
ws = ops.HbmWorkspace(n, n, torch.complex64, dev)
ws.pack(psi[0])
chain = ops.new_chain(1, dev, amplitude=0.3)[0]
ws.sweep(par, chain, 1, 0, 5, 0, _abi.SWEEP_ADAPT_SIGMA)      # counters are zeroed at the start of each call
nsw = 10
ws.sweep(par, chain, 1, 5, nsw, 0, _abi.SWEEP_ADAPT_SIGMA)
torch.cuda.synchronize()
c = ws.scratch[:8].cpu().numpy()
ncta = (n // 32) * (n // 128) * nsw
names = ["prologue (coef, TMA issue)", "TMA wait", "phase 0 main loop", "phase 0 halo-column sites", "phase 0 barrier", "phase 1 + barrier", "write-back"]
tot = sum(c[1:8])
for nm, v in zip(names, c[1:8]):
    print("%-28s %8.0f cycles/CTA  %5.1f%%" % (nm, v / ncta, 100 * v / tot))
print("total %.0f cycles/CTA; CTAs %d" % (tot / ncta, ncta))
