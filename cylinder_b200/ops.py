"""Tensor-level operators over the C-ABI (include/cylinder_b200.h).  Everything here takes and returns CUDA
tensors; nothing computes on the host.  The reference-facing classes (lattice.py, system_cylinder.py,
replicas.py, ...) are built on these.
"""
import numpy as np
import torch

from . import _abi
from ._abi import (CHAIN_DOUBLES, NMOM, OBS_DOUBLES, PARAM_DOUBLES, PARAM_FIELDS, SWEEP_ADAPT_SIGMA, SWEEP_AMP_MOVES,
                   SWEEP_MEASURE, VARIANTS, call, ptr, stream)

F64 = torch.float64


def _dev(t):
    return t.device.index if t.device.index is not None else torch.cuda.current_device()


def variant_id(variant):
    if isinstance(variant, str):
        return VARIANTS[variant]
    return int(variant)


def make_params(device, **kw):
    """[B, 12] float64 tensor of CylParams rows from scalars / sequences (broadcast to a common B)."""
    defaults = dict(gamma=1.0, kappa=0.0, intrinsic_curvature=0.0, amp_bound=0.8, amp_target_acceptance=0.3)
    cols = []
    B = 1
    for name in PARAM_FIELDS:
        v = kw.get(name, defaults.get(name))
        if v is None:
            raise ValueError("missing parameter %s" % name)
        v = np.atleast_1d(np.asarray(v, dtype=np.float64))
        B = max(B, v.shape[0])
        cols.append(v)
    out = np.empty((B, PARAM_DOUBLES))
    for c, v in enumerate(cols):
        out[:, c] = v
    return torch.from_numpy(out).to(device)


def lattice_dims(dims, wavenumber):
    """z_len = int(round(dims[0]/k)), th_len = dims[1]    On_lattice_simple.py:41-42"""
    return int(round(dims[0] / wavenumber)), int(dims[1])


def new_chain(B, device, amplitude=0.0, sigma_site=.005, sigma_amp=.005):
    """Fresh chain state [B, 16]: sigma_0 = .005 (On_lattice_simple.py:101, engine sampling_width :81)."""
    ch = torch.zeros(B, CHAIN_DOUBLES, dtype=F64, device=device)
    ch[:, _abi.CH_AMPLITUDE] = torch.as_tensor(amplitude, dtype=F64, device=device)
    ch[:, _abi.CH_SIGMA_SITE] = torch.as_tensor(sigma_site, dtype=F64, device=device)
    ch[:, _abi.CH_SIGMA_AMP] = torch.as_tensor(sigma_amp, dtype=F64, device=device)
    return ch


# ------------------------------------------------------------------------------------------------------
def geometry_eval(k, r, a, z):
    a = a.to(F64).contiguous()
    z = z.to(F64).contiguous()
    out = torch.empty(a.numel(), 6, dtype=F64, device=a.device)
    call("cyl_geometry_eval_f64", _dev(a), float(k), float(r), ptr(a), ptr(z), a.numel(), ptr(out), stream())
    return out


def surface_energy(params, amp, return_terms=False):
    B = params.shape[0]
    amp = amp.to(F64).contiguous()
    E = torch.empty(B, dtype=F64, device=params.device)
    terms = torch.empty(B, 2, dtype=F64, device=params.device) if return_terms else None
    call("cyl_surface_energy_f64", _dev(params), ptr(params), ptr(amp), B, ptr(E), ptr(terms), stream())
    return (E, terms) if return_terms else E


def philox4x32_10(ctr, key):
    n = ctr.shape[0]
    out = torch.empty_like(ctr)
    call("cyl_philox4x32_10", _dev(ctr), ptr(ctr), ptr(key), n, ptr(out), stream())
    return out


def philox2x32_10(ctr, key):
    n = ctr.shape[0]
    out = torch.empty_like(ctr)
    call("cyl_philox2x32_10", _dev(ctr), ptr(ctr), ptr(key), n, ptr(out), stream())
    return out


def _suffix(psi):
    if psi.dtype == torch.complex64:
        return "f32"
    if psi.dtype == torch.complex128:
        return "f64"
    raise TypeError("psi must be complex64 or complex128, got %s" % psi.dtype)


def lattice_init(psi, Z, mag_max=.1, seed=0, replica0=0):
    """psi [B, Zmax, T] complex, Z [B] int32: random_initialize (On_lattice_simple.py:159-182) from Philox."""
    B, Zmax, T = psi.shape
    call("cyl_lattice_init_" + _suffix(psi), _dev(psi), ptr(psi), B, ptr(Z), Zmax, T, float(mag_max), int(seed),
         int(replica0), stream())
    return psi


def replay(psi, params, amp, variant, iz, ith, zre, zim, u, rm_state):
    """Sequential tape replay on psi [Z, T] complex128 (in place).  Returns (accept u8[n], dE[n], sigma[n])."""
    Z, T = psi.shape
    n = iz.numel()
    dev = psi.device
    acc = torch.empty(n, dtype=torch.uint8, device=dev)
    dE = torch.empty(n, dtype=F64, device=dev)
    sig = torch.empty(n, dtype=F64, device=dev)
    call("cyl_replay_f64", _dev(psi), ptr(psi), Z, T, ptr(params), float(amp), variant_id(variant), ptr(iz), ptr(ith),
         ptr(zre), ptr(zim), ptr(u), n, ptr(rm_state), ptr(acc), ptr(dE), ptr(sig), stream())
    return acc, dE, sig


def site_delta(psi, params, amp, variant, iz, ith, zre, zim, sigma, out=None):
    """dE and proposed value of the move psi[iz, ith] += sigma_eff (zre + i zim).  Returns [4] = (dE, new_re, new_im, sqrt_g)."""
    Z, T = psi.shape
    if out is None:
        out = torch.empty(4, dtype=F64, device=psi.device)
    call("cyl_site_delta_f64", _dev(psi), ptr(psi), Z, T, ptr(params), float(amp), variant_id(variant), int(iz), int(ith),
         float(zre), float(zim), float(sigma), ptr(out), stream())
    return out


def row_moments(psi, Z, params):
    """psi [B, Zmax, T] -> S [B, Zmax, 8] float64."""
    B, Zmax, T = psi.shape
    S = torch.empty(B, Zmax, NMOM, dtype=F64, device=psi.device)
    call("cyl_row_moments_" + _suffix(psi), _dev(psi), ptr(psi), B, ptr(Z), Zmax, T, ptr(params), ptr(S), stream())
    return S


def field_energy(S, Z, T, params, amp_eval, amp_ref=None, variant=0):
    B, Zmax, _ = S.shape
    amp_eval = torch.as_tensor(amp_eval, dtype=F64, device=S.device).reshape(-1).contiguous()
    amp_ref = None if amp_ref is None else torch.as_tensor(amp_ref, dtype=F64, device=S.device).reshape(-1).contiguous()
    E = torch.empty(B, dtype=F64, device=S.device)
    call("cyl_field_energy_f64", _dev(S), ptr(S), B, ptr(Z), Zmax, int(T), ptr(params), ptr(amp_eval), ptr(amp_ref),
         variant_id(variant), ptr(E), stream())
    return E


SERIES_NAMES = ("amplitude", "field_energy", "surface_energy", "psi_squared", "abs_psi", "sampling_width", "amplitude_accepted",
                "site_acceptance")                                  # CYL_TS_* of include/cylinder_b200.h


def sweep_smem(psi, Z, params, chain, seed, sweep0, nsweeps, variant=0, flags=SWEEP_ADAPT_SIGMA, replica0=0, obs=None,
               prof=None, series=None):
    """nsweeps checkerboard sweeps of B shared-memory-resident replicas (in place on psi, chain, obs, prof).
    `series` [B, nsweeps, 8] float64 (optional, needs SWEEP_MEASURE): one record per sweep and replica."""
    B, Zmax, T = psi.shape
    if series is None:
        call("cyl_sweep_smem_" + _suffix(psi), _dev(psi), ptr(psi), B, ptr(Z), Zmax, T, ptr(params), ptr(chain), int(seed),
             int(replica0), int(sweep0), int(nsweeps), variant_id(variant), int(flags), ptr(obs), ptr(prof), stream())
    else:
        if tuple(series.shape) != (B, int(nsweeps), len(SERIES_NAMES)) or series.dtype != F64:
            raise ValueError("series must be a float64 tensor of shape [B, nsweeps, %d]" % len(SERIES_NAMES))
        call("cyl_sweep_smem_series_" + _suffix(psi), _dev(psi), ptr(psi), B, ptr(Z), Zmax, T, ptr(params), ptr(chain), int(seed),
             int(replica0), int(sweep0), int(nsweeps), variant_id(variant), int(flags), ptr(obs), ptr(prof), ptr(series), stream())


def new_obs(B, Zmax, device):
    return (torch.zeros(B, OBS_DOUBLES, dtype=F64, device=device), torch.zeros(B, Zmax, 3, dtype=F64, device=device))


# ------------------------------------------------------------------------------------------------------
# one large lattice streamed from HBM
# ------------------------------------------------------------------------------------------------------
def hbm_layout(Z, T, is_f64):
    import ctypes as C
    hr, hc = C.c_int(), C.c_int()
    pitch, buf, scr = C.c_int64(), C.c_int64(), C.c_int64()
    _abi.check(_abi.lib().cyl_hbm_layout(int(Z), int(T), int(bool(is_f64)), C.byref(hr), C.byref(hc), C.byref(pitch),
                                         C.byref(buf), C.byref(scr)), "cyl_hbm_layout")
    return dict(halo_rows=hr.value, halo_cols=hc.value, pitch=pitch.value, buffer_elems=buf.value, scratch_bytes=scr.value)


class HbmWorkspace:
    """Device buffers of the HBM path for one Z x T lattice: ping-pong planes, buffer index, scratch."""

    def __init__(self, Z, T, dtype, device):
        self.Z, self.T, self.dtype, self.device = int(Z), int(T), dtype, device
        self.suffix = "f32" if dtype == torch.complex64 else "f64"
        self.layout = hbm_layout(Z, T, dtype == torch.complex128)
        self.work = torch.zeros(2 * self.layout["buffer_elems"], dtype=dtype, device=device)
        import ctypes
        self.cur = ctypes.c_int32(0)          # host-side index of the buffer that holds the lattice
        self.scratch = torch.zeros((self.layout["scratch_bytes"] + 7) // 8, dtype=torch.int64, device=device)

    def pack(self, psi):
        psi = psi.to(self.dtype).contiguous()
        assert psi.shape == (self.Z, self.T)
        call("cyl_hbm_pack_" + self.suffix, _dev(psi), ptr(psi), self.Z, self.T, ptr(self.work), self.cur, stream())

    def unpack(self, out=None):
        if out is None:
            out = torch.empty(self.Z, self.T, dtype=self.dtype, device=self.device)
        call("cyl_hbm_unpack_" + self.suffix, _dev(out), ptr(self.work), int(self.cur.value), self.Z, self.T, ptr(out), stream())
        return out

    def sweep(self, params, chain, seed, sweep0, nsweeps, variant=0, flags=SWEEP_ADAPT_SIGMA, replica=0, obs=None, prof=None):
        call("cyl_sweep_hbm_" + self.suffix, _dev(self.work), ptr(self.work), self.cur, self.Z, self.T, ptr(params),
             ptr(chain), int(seed), int(replica), int(sweep0), int(nsweeps), variant_id(variant), int(flags),
             ptr(self.scratch), ptr(obs), ptr(prof), stream())

    def row_moments(self, params):
        S = torch.empty(self.Z, NMOM, dtype=F64, device=self.device)
        call("cyl_hbm_row_moments_" + self.suffix, _dev(self.work), ptr(self.work), int(self.cur.value), self.Z, self.T,
             ptr(params), ptr(S), stream())
        return S


# ------------------------------------------------------------------------------------------------------
# Fourier-mode model
# ------------------------------------------------------------------------------------------------------
def fourier_energy(params, nz, nth, amp, coeffs):
    """coeffs [B, 2*nth+1, 2*nz+1] complex128, amp [B] -> E [B]   (system_cylinder2D.py:151-198)."""
    B = params.shape[0]
    amp = amp.to(F64).contiguous()
    coeffs = coeffs.to(torch.complex128).contiguous()
    E = torch.empty(B, dtype=F64, device=params.device)
    call("cyl_fourier_energy_f64", _dev(params), ptr(params), int(nz), int(nth), ptr(amp), ptr(coeffs), B, ptr(E), stream())
    return E


def fourier_surface_energy(params, amp):
    B = params.shape[0]
    amp = amp.to(F64).contiguous()
    E = torch.empty(B, dtype=F64, device=params.device)
    call("cyl_fourier_surface_energy_f64", _dev(params), ptr(params), ptr(amp), B, ptr(E), stream())
    return E


def fourier_tensor_tables(params, nz, nth, amp):
    """The reference's A[d] (d = -4nz..4nz) and diagonal B[i, j, beta] tables for B (parameter row, amplitude) pairs:
    complex128 tensors [B, 8nz+1] and [B, 2nz+1, 2nz+1, 2nth+1] (system_cylinder2D.py:103-147)."""
    B = amp.numel()
    A = torch.empty(B, 8 * nz + 1, dtype=torch.complex128, device=amp.device)
    Bm = torch.empty(B, 2 * nz + 1, 2 * nz + 1, 2 * nth + 1, dtype=torch.complex128, device=amp.device)
    call("cyl_fourier_tensor_tables_f64", _dev(params), ptr(params), int(nz), int(nth), ptr(amp), B, ptr(A), ptr(Bm), stream())
    return A, Bm


# ------------------------------------------------------------------------------------------------------
# whole-lattice proposals (cyl_move_*): shift / plane wave / row twist
# ------------------------------------------------------------------------------------------------------
MOVE_SHIFT, MOVE_WAVE, MOVE_TWIST = 0, 1, 2
_MOVE_DTYPE = np.dtype([("kind", "<i4"), ("r0", "<i4"), ("nr", "<i4"), ("n_rot", "<i4"), ("c_re", "<f8"), ("c_im", "<f8"),
                        ("qz", "<f8"), ("qth", "<f8"), ("phase", "<f8"), ("reserved", "<f8")])          # struct CylMove, 64 bytes


def make_moves(device, kind, B=1, c=0j, qz=0.0, qth=0.0, r0=0, nr=0, n_rot=0, phase=0.0):
    """[B] CylMove records (uint8 tensor [B, 64] on `device`) from scalars / arrays (broadcast)."""
    m = np.zeros(B, dtype=_MOVE_DTYPE)
    c = np.broadcast_to(np.asarray(c, dtype=complex), (B,))
    m["kind"], m["r0"], m["nr"], m["n_rot"] = kind, r0, nr, n_rot
    m["c_re"], m["c_im"], m["qz"], m["qth"], m["phase"] = c.real, c.imag, qz, qth, phase
    return torch.from_numpy(m.view(np.uint8).reshape(B, _MOVE_DTYPE.itemsize).copy()).to(device)


def move_row_moments(psi, Z, params, moves):
    """Row moments [B, Zmax, 8] of the field the moves would produce (psi itself is not touched)."""
    B, Zmax, T = psi.shape
    S = torch.empty(B, Zmax, NMOM, dtype=F64, device=psi.device)
    call("cyl_move_row_moments_" + _suffix(psi), _dev(psi), ptr(psi), B, ptr(Z), Zmax, T, ptr(params), ptr(moves), ptr(S), stream())
    return S


def move_apply(psi, Z, moves, accept=None):
    """Commit the moves in place where accept[b] != 0 (uint8 tensor; None = everywhere)."""
    B, Zmax, T = psi.shape
    call("cyl_move_apply_" + _suffix(psi), _dev(psi), ptr(psi), B, ptr(Z), Zmax, T, ptr(moves), ptr(accept) if accept is not None else None,
         stream())
    return psi
