"""ctypes binding of the C-ABI declared in include/cylinder_b200.h.

The product has NO CPU fallback: if the shared library is missing this module raises at first use, loudly.
Tensors are passed as raw device pointers (tensor.data_ptr()) plus the current torch CUDA stream.
"""
import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libcylinder_b200.so")
HEADER = os.path.join(HERE, "..", "include", "cylinder_b200.h")

ABI_VERSION = 8

# constants mirrored from the header (checked against it in tests/test_abi_cpu.py)
VARIANT_SIMPLE, VARIANT_RESCALED_0TH, VARIANT_RESCALED_1ST, VARIANT_DECOUPLED = 0, 1, 2, 3
VARIANTS = {"simple": VARIANT_SIMPLE, "rescaled_0thorder": VARIANT_RESCALED_0TH, "rescaled_1storder": VARIANT_RESCALED_1ST,
            "decoupled": VARIANT_DECOUPLED}
SWEEP_ADAPT_SIGMA, SWEEP_AMP_MOVES, SWEEP_MEASURE, SWEEP_INDEX_ORDER = 1, 2, 4, 8


def sweep_tile_rows(n):
    """CYL_SWEEP_TILE_ROWS(n): explicit tile height of the HBM path (32, 30, 28, 24), OR-ed into the sweep flags."""
    return (int(n) & 0xFF) << 8


CHAIN_DOUBLES, OBS_DOUBLES, NMOM = 16, 16, 8
CH_AMPLITUDE, CH_SIGMA_SITE, CH_SIGMA_AMP, CH_E_FIELD, CH_E_SURF, CH_STEP_COUNTER = 0, 1, 2, 3, 4, 5
CH_SITE_PROPOSED, CH_SITE_ACCEPTED, CH_AMP_PROPOSED, CH_AMP_ACCEPTED, CH_AMP_COUNTER, CH_FLAGS = 6, 7, 8, 9, 10, 11
OB_COUNT, OB_ABS_A, OB_A, OB_A2, OB_PSI2, OB_ABSPSI, OB_E_FIELD, OB_E_SURF, OB_E_FIELD2, OB_SIGMA = range(10)
PARAM_FIELDS = ("wavenumber", "radius", "gamma", "kappa", "intrinsic_curvature", "alpha", "u", "C", "n",
                "temperature", "amp_bound", "amp_target_acceptance")
PARAM_DOUBLES = len(PARAM_FIELDS)

_vp, _i, _i64, _u64, _u32, _d = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_uint32, C.c_double

# name -> argtypes (restype is int unless listed in _RESTYPES)
_SIGNATURES = {
    "cyl_abi_version": [],
    "cyl_last_error": [],
    "cyl_set_device": [_i],
    "cyl_device_info": [C.POINTER(_i)] * 4,
    "cyl_geometry_eval_f64": [_d, _d, _vp, _vp, _i64, _vp, _vp],
    "cyl_surface_energy_f64": [_vp, _vp, _i, _vp, _vp, _vp],
    "cyl_philox4x32_10": [_vp, _vp, _i64, _vp, _vp],
    "cyl_philox2x32_10": [_vp, _vp, _i64, _vp, _vp],
    "cyl_lattice_init_f32": [_vp, _i, _vp, _i, _i, _d, _u64, _i64, _vp],
    "cyl_lattice_init_f64": [_vp, _i, _vp, _i, _i, _d, _u64, _i64, _vp],
    "cyl_replay_f64": [_vp, _i, _i, _vp, _d, _i, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp],
    "cyl_site_delta_f64": [_vp, _i, _i, _vp, _d, _i, _i, _i, _d, _d, _d, _vp, _vp],
    "cyl_row_moments_f32": [_vp, _i, _vp, _i, _i, _vp, _vp, _vp],
    "cyl_row_moments_f64": [_vp, _i, _vp, _i, _i, _vp, _vp, _vp],
    "cyl_move_row_moments_f32": [_vp, _i, _vp, _i, _i, _vp, _vp, _vp, _vp],
    "cyl_move_row_moments_f64": [_vp, _i, _vp, _i, _i, _vp, _vp, _vp, _vp],
    "cyl_move_apply_f32": [_vp, _i, _vp, _i, _i, _vp, _vp, _vp],
    "cyl_move_apply_f64": [_vp, _i, _vp, _i, _i, _vp, _vp, _vp],
    "cyl_field_energy_f64": [_vp, _i, _vp, _i, _i, _vp, _vp, _vp, _i, _vp, _vp],
    "cyl_sweep_smem_f32": [_vp, _i, _vp, _i, _i, _vp, _vp, _u64, _i64, _u64, _i, _i, _u32, _vp, _vp, _vp],
    "cyl_sweep_smem_f64": [_vp, _i, _vp, _i, _i, _vp, _vp, _u64, _i64, _u64, _i, _i, _u32, _vp, _vp, _vp],
    "cyl_sweep_smem_bytes": [_i, _i, _i],
    "cyl_sweep_smem_series_f32": [_vp, _i, _vp, _i, _i, _vp, _vp, _u64, _i64, _u64, _i, _i, _u32, _vp, _vp, _vp, _vp],
    "cyl_sweep_smem_series_f64": [_vp, _i, _vp, _i, _i, _vp, _vp, _u64, _i64, _u64, _i, _i, _u32, _vp, _vp, _vp, _vp],
    "cyl_hbm_layout": [_i, _i, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)],
    "cyl_hbm_pack_f32": [_vp, _i, _i, _vp, C.POINTER(C.c_int32), _vp],
    "cyl_hbm_pack_f64": [_vp, _i, _i, _vp, C.POINTER(C.c_int32), _vp],
    "cyl_hbm_unpack_f32": [_vp, _i, _i, _i, _vp, _vp],
    "cyl_hbm_unpack_f64": [_vp, _i, _i, _i, _vp, _vp],
    "cyl_sweep_hbm_f32": [_vp, C.POINTER(C.c_int32), _i, _i, _vp, _vp, _u64, _i64, _u64, _i, _i, _u32, _vp, _vp, _vp, _vp],
    "cyl_sweep_hbm_f64": [_vp, C.POINTER(C.c_int32), _i, _i, _vp, _vp, _u64, _i64, _u64, _i, _i, _u32, _vp, _vp, _vp, _vp],
    "cyl_hbm_row_moments_f32": [_vp, _i, _i, _i, _vp, _vp, _vp],
    "cyl_hbm_row_moments_f64": [_vp, _i, _i, _i, _vp, _vp, _vp],
    "cyl_fourier_energy_f64": [_vp, _i, _i, _vp, _vp, _i, _vp, _vp],
    "cyl_fourier_surface_energy_f64": [_vp, _vp, _i, _vp, _vp],
    "cyl_fourier_tensor_tables_f64": [_vp, _i, _i, _vp, _i, _vp, _vp, _vp],
    "cyl_fourier_am_state_doubles": [_i, _i],
    "cyl_fourier_am_steps_f64": [_vp, _i, _i, _vp, _vp, _vp, _i, _u64, _i64, _u64, _i, _i, _i, _d, _d, _d, _i, _vp],
}
_RESTYPES = {"cyl_last_error": C.c_char_p, "cyl_sweep_smem_bytes": C.c_int64}

_lib = None


class CylinderB200Error(RuntimeError):
    pass


def declared_symbols():
    """Every function the header declares (used by the CPU test that the library exports them all)."""
    with open(HEADER) as f:
        text = f.read()
    return sorted(set(re.findall(r"\b(cyl_[a-z0-9_]+)\s*\(", text)))


def lib():
    """Load the shared library once.  Raises if it has not been built -- there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CylinderB200Error(
                "%s is missing: build it with `python -m cylinder_b200.build` (needs nvcc). "
                "cylinder_b200 has no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, argtypes in _SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError if the .so does not export a declared symbol
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, C.c_int)
        v = L.cyl_abi_version()
        if v != ABI_VERSION:
            raise CylinderB200Error("ABI mismatch: library %d, python binding %d -- rebuild" % (v, ABI_VERSION))
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().cyl_last_error()
        raise CylinderB200Error("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))


def _stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise CylinderB200Error("expected a CUDA tensor; cylinder_b200 has no CPU path")
    if not t.is_contiguous():
        raise CylinderB200Error("expected a contiguous tensor")
    return t.data_ptr()


class _CurrentStream:
    """Placeholder for "the current torch stream of the device this call runs on" (resolved inside call())."""


_CUR = _CurrentStream()


def call(name, device, *args):
    """Call `name` on `device` (index of the tensors' GPU): the CUDA device is set for the duration of the call only (torch's
    device guard, so the caller's current device is restored) and the stream handle is the current torch stream OF THAT
    device, not of whatever device happens to be current.  Raises on a non-zero return code."""
    import torch
    L = lib()
    if device is None:
        device = torch.cuda.current_device()
    with torch.cuda.device(int(device)):
        sh = torch.cuda.current_stream(int(device)).cuda_stream
        check(getattr(L, name)(*[sh if a is _CUR else a for a in args]), name)


def stream():
    return _CUR
