"""Timeline of the three kernels of a C4 sweep with amplitude moves (preparation -> sweep -> decision), from %globaltimer stamps.

The stamps are NOT in the product sources; they are a patch:

    git apply scripts/chain_timing.patch && sh scripts/build_variant.sh chaintime k_sweep_hbm.cu; git apply -R scripts/chain_timing.patch
    python scripts/chain_timing.py            # on a B200; reads build_variants/chaintime.so

Every kernel writes its stamps into a ring of 8 sweeps in the second KB of the scratch header (see the patch for the slots);
printed per sweep, in us relative to the moment the preparation kernel passed its dependency wait."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cylinder_b200 import _abi, ops
_abi.LIB_PATH = os.path.abspath(os.environ.get("CYL_LIB", "build_variants/chaintime.so"))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 3            # 3 = ADAPT | AMP_MOVES, 7 = + MEASURE
dev = torch.device("cuda", 0)
par = ops.make_params(dev, wavenumber=1.0, radius=1.0, alpha=-1.0, u=1.0, C=1.0, n=6.0, temperature=.01)
psi = torch.zeros(1, n, n, dtype=torch.complex64, device=dev)
ops.lattice_init(psi, torch.tensor([n], dtype=torch.int32, device=dev), .1, 1, 0)
ws = ops.HbmWorkspace(n, n, torch.complex64, dev)
ws.pack(psi[0])
chain = ops.new_chain(1, dev, amplitude=0.3)[0]
obs = torch.zeros(16, dtype=torch.float64, device=dev)
prof = torch.zeros(n, 3, dtype=torch.float64, device=dev)
ws.sweep(par, chain, 1, 0, 10, 0, flags, obs=obs, prof=prof)
ws.sweep(par, chain, 1, 10, 40, 0, flags, obs=obs, prof=prof)
torch.cuda.synchronize()
d = ws.scratch.view(torch.uint8)[1024:1536].view(torch.int64).cpu().numpy().reshape(8, 8)
for s in range(43, 49):                                          # the ring holds sweeps 42..49 (slot = sweep & 7)
    r, nxt = d[s & 7], d[(s + 1) & 7]
    t0 = r[6]
    us = lambda t: (t - t0) / 1e3
    print("sweep %d: preparation past its wait 0 | row blocks done %+.2f, header block %+.2f | first tile past its wait %+.2f | last tile done "
          "%+.2f | decision past its wait %+.2f, sums added %+.2f, taken %+.2f | next preparation past its wait %+.2f us"
          % (s, us(r[7]), us(r[2]), us(r[0]), us(r[1]), us(r[3]), us(r[4]), us(r[5]), us(nxt[6])))
