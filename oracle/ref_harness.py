"""TEST INFRASTRUCTURE ONLY -- never imported by the product package `cylinder_b200`.

Loads the *unmodified* reference (jklebes/cylinder, mounted read-only at /root/reference)
inside this container so that golden vectors can be generated from it
(`tests/golden/make_golden.py`) and the oracle restatement can be validated against it.
`/root/reference` does not exist on the GPU box: nothing that runs there may import this.

The reference needs three modules that are not installed here:
  * matplotlib / matplotlib.pyplot and seaborn  -- import-only, stubbed with empty modules;
  * metropolisengine (third-party, version string "0.2.19", run.py:19) -- un-vendored.
    The shim below provides only what the lattice hot path calls:
      ctor kwargs             On_lattice_simple.py:80-82
      set_reject_condition    On_lattice_simple.py:83
      metropolis_decision     On_lattice_simple.py:670 (accept rule: predecessor bytecode
                              __pycache__/metropolis_engine.cpython-36.pyc, src lines 226-246, and the
                              five known answers in tests/__pycache__/test_metropolis_engine*.pyc)
      energy / real_params    On_lattice_simple.py:546,562

A `Tape` records the random stream the reference consumes per proposal
 by wrapping random.randrange / random.gauss / random.uniform.
"""
import contextlib
import io
import math
import os
import random
import sys
import types

REFERENCE_ROOT = os.environ.get("CYL_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "surfaces_and_fields"))


class ShimMetropolisEngine:
    """Minimal stand-in for metropolisengine.MetropolisEngine (accept rule only)."""

    def __init__(self, energy_functions=None, initial_complex_params=None, initial_real_params=None,
                 covariance_matrix_complex=None, sampling_width=0.05, temp=0,
                 complex_sample_method=None, **kwargs):
        self.energy_functions = energy_functions
        self.real_params = list(initial_real_params) if initial_real_params is not None else []
        self.complex_params = initial_complex_params
        self.temp = temp
        self.sampling_width = sampling_width
        self.energy = {}
        self.reject_condition = lambda real, cplx: False
        self.decisions = []          # (old, new, accept, u_or_None) for every call

    def set_reject_condition(self, fn):
        self.reject_condition = fn

    def metropolis_decision(self, old_energy, proposed_energy):
        diff = proposed_energy - old_energy
        u = None
        if diff <= 0:
            accept = True
        elif self.temp == 0:
            accept = False
        else:
            u = random.uniform(0, 1)
            accept = u <= math.exp(-diff / self.temp)
        self.decisions.append((old_energy, proposed_energy, accept, u))
        return accept

    def measure_real_system(self):
        pass

    @staticmethod
    def gaussian_complex(sigma=0.1):
        import cmath
        return cmath.rect(random.gauss(0, sigma), random.uniform(0, 2 * math.pi))


def install_stubs():
    """Put the stub modules in sys.modules and the reference root on sys.path."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    shim = types.ModuleType("metropolisengine")
    shim.MetropolisEngine = ShimMetropolisEngine
    shim.__version__ = "shim"
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)
    sys.modules.setdefault("seaborn", types.ModuleType("seaborn"))
    sys.modules["metropolisengine"] = shim
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def load(module):
    """Import surfaces_and_fields.<module> from the reference (stdout silenced)."""
    install_stubs()
    import importlib
    with contextlib.redirect_stdout(io.StringIO()):
        return importlib.import_module("surfaces_and_fields." + module)


def quiet(fn, *a, **k):
    """Call fn with stdout silenced (the reference prints whole lattices on construction)."""
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


class Tape:
    """Records randrange / unit normals / uniforms consumed by the reference.

    random.gauss(mu, sigma) is replaced by mu + gauss(0, 1) * sigma: CPython's gauss returns
    `mu + z * sigma`, so with mu=0, sigma=1 the inner call returns z exactly and the outer
    arithmetic is bit-identical to the unwrapped call.  random.uniform(0, 1) = 0 + (1-0)*random().
    """

    def __init__(self):
        self.randrange = []
        self.normals = []
        self.uniforms = []
        self._orig = None

    def __enter__(self):
        self._orig = (random.randrange, random.gauss, random.uniform)
        o_rr, o_g, o_u = self._orig

        def rr(*a, **k):
            v = o_rr(*a, **k)
            self.randrange.append(v)
            return v

        def g(mu=0.0, sigma=1.0):
            z = o_g(0.0, 1.0)
            self.normals.append(z)
            return mu + z * sigma

        def u(a, b):
            v = o_u(a, b)
            self.uniforms.append(v)
            return v

        random.randrange, random.gauss, random.uniform = rr, g, u
        return self

    def __exit__(self, *exc):
        random.randrange, random.gauss, random.uniform = self._orig
        return False
