"""Wall time of the drop-in class on the reference's own case (config C1): python scripts/time_lattice_run.py [n_steps] [fieldsteps]"""
import os, sys, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cylinder_b200
n_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 500
fs = int(sys.argv[2]) if len(sys.argv) > 2 else 100
random.seed(12345)
lat = cylinder_b200.LatticeRescaled0thOrder(amplitude=0, wavenumber=1, radius=1, gamma=1, kappa=0, intrinsic_curvature=0, alpha=-1, u=1,
                                            C=1, n=6, temperature=.01, temperature_final=.01, dims=(50, 50), fieldsteps_per_ampstep=fs)
lat.run(20, 50)
torch.cuda.synchronize()
t0 = time.perf_counter()
lat.run(n_steps, 50)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
per_step = lat._sweeps_for(fs * 50)
print("Lattice.run: %d steps x %d sweeps (50x50): %.3f s = %.2f ms per step, %.3e site-updates/s; <|a|> = %.4f, acceptance %.3f"
      % (n_steps, per_step, dt, 1e3 * dt / n_steps, n_steps * per_step * 2500 / dt, lat.amplitude_average, lat.acceptance_counter / max(1, lat.step_counter)))
