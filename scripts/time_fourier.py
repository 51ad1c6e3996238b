"""Throughput of the batched Fourier-model kernels (config 2): python scripts/time_fourier.py [B]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from cylinder_b200 import ops, _abi
if os.environ.get("CYL_LIB"):
    _abi.LIB_PATH = os.path.abspath(os.environ["CYL_LIB"])      # a variant built by scripts/build_variant.sh
from cylinder_b200.fourier_chains import FourierChains
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda", 0)
P = dict(wavenumber=.5, radius=1.0, gamma=1.0, kappa=0.0, intrinsic_curvature=0.0, alpha=-1.0, u=1.0, C=1.0, n=6.0, temperature=.1, amp_bound=.99)
fc = FourierChains(dict(P, wavenumber=np.full(B, .5)), num_field_coeffs=(2, 1), amplitude=0.0, seed=1)
fc.run(20, measure_every=10, fieldsteps_per_ampstep=10)
torch.cuda.synchronize()
n_steps = 50
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); fc.run(n_steps, measure_every=10, fieldsteps_per_ampstep=10); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
moves = B * n_steps * 10 * 11
print("adaptive engine, (nz,nth)=(2,1), %d chains: %.1f ms for %d steps -> %.3e engine moves/s, %.3e coefficient-updates/s (x15)" % (B, ms, n_steps, moves / ms * 1e3, 15 * moves / ms * 1e3))
s = fc.summary()
print("acceptance real %.3f complex %.3f  <|a|> %.4f" % (s["real_acceptance"].mean().item(), s["complex_acceptance"].mean().item(), s["abs_amplitude"].mean().item()))
amps = torch.rand(B, dtype=torch.float64, device=dev) * .6
e0.record()
for _ in range(20):
    E = ops.fourier_energy(fc.params, 2, 1, amps, fc.coeffs)
e1.record(); torch.cuda.synchronize()
print("energy evaluations: %.3e /s (reference: 0.24 ms per calc_field_energy, 47 ms per calc_field_energy_amplitude_change on one CPU core)" % (20 * B / e0.elapsed_time(e1) * 1e3))
