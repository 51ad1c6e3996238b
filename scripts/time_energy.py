"""Time the energy / moment kernels on a 4096 x 4096 lattice: python scripts/time_energy.py [n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cylinder_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda", 0)
par = ops.make_params(dev, wavenumber=1.0, radius=1.0, alpha=-1.0, u=1.0, C=1.0, n=6.0, temperature=.01)
Zt = torch.tensor([n], dtype=torch.int32, device=dev)
for dt in (torch.complex64, torch.complex128):
    psi = torch.zeros(1, n, n, dtype=dt, device=dev)
    ops.lattice_init(psi, Zt, .1, 1, 0)
    ws = ops.HbmWorkspace(n, n, dt, dev)
    ws.pack(psi[0])
    moves = ops.make_moves(dev, ops.MOVE_WAVE, c=.01 + .02j, qz=6.283, qth=12.566)
    def t(fn, reps=10):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return 1e3 * e0.elapsed_time(e1) / reps
    nb = psi.element_size() * n * n
    for name, fn in (("row_moments (packed)", lambda: ops.row_moments(psi, Zt, par)), ("row_moments (tiled layout)", lambda: ws.row_moments(par)),
                     ("move_row_moments (plane wave)", lambda: ops.move_row_moments(psi, Zt, par, moves)),
                     ("pack", lambda: ws.pack(psi[0])), ("unpack", lambda: ws.unpack(psi[0]))):
        us = t(fn)
        print("%s %-30s %8.1f us  %6.0f GB/s of lattice bytes read" % (str(dt)[6:], name, us, nb / us / 1e3))
