"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/cylinder_b200.h declares; the Python mirror of the header constants agrees with the header.  No compute
calls (no GPU here)."""
import ctypes
import re

from cylinder_b200 import _abi


def test_library_exports_every_declared_symbol(built_lib):
    L = ctypes.CDLL(built_lib)
    declared = _abi.declared_symbols()
    assert len(declared) >= 20
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    # the python binding binds exactly the declared set
    assert sorted(_abi._SIGNATURES) == declared


def test_abi_version_and_error_text(built_lib):
    L = _abi.lib()
    assert L.cyl_abi_version() == _abi.ABI_VERSION
    assert isinstance(L.cyl_last_error(), bytes)


def test_header_constants_match_python():
    text = open(_abi.HEADER).read()
    defs = dict(re.findall(r"#define\s+(CYL_[A-Z0-9_]+)\s+\(?(-?\d+)u?\)?", text))
    assert int(defs["CYL_ABI_VERSION"]) == _abi.ABI_VERSION
    assert int(defs["CYL_CHAIN_DOUBLES"]) == _abi.CHAIN_DOUBLES
    assert int(defs["CYL_OBS_DOUBLES"]) == _abi.OBS_DOUBLES
    for name in ("AMPLITUDE", "SIGMA_SITE", "SIGMA_AMP", "E_FIELD", "E_SURF", "STEP_COUNTER", "SITE_PROPOSED",
                 "SITE_ACCEPTED", "AMP_PROPOSED", "AMP_ACCEPTED", "AMP_COUNTER", "FLAGS"):
        assert int(defs["CYL_CH_" + name]) == getattr(_abi, "CH_" + name)
    for name in ("COUNT", "ABS_A", "A", "A2", "PSI2", "ABSPSI", "E_FIELD", "E_SURF", "E_FIELD2", "SIGMA"):
        assert int(defs["CYL_OB_" + name]) == getattr(_abi, "OB_" + name)
    assert int(defs["CYL_SWEEP_ADAPT_SIGMA"]) == _abi.SWEEP_ADAPT_SIGMA
    assert int(defs["CYL_SWEEP_AMP_MOVES"]) == _abi.SWEEP_AMP_MOVES
    assert int(defs["CYL_SWEEP_MEASURE"]) == _abi.SWEEP_MEASURE
    # struct CylParams field order
    body = re.search(r"typedef struct CylParams \{(.*?)\} CylParams;", text, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = [f.strip() for decl in re.findall(r"double\s+([^;]+);", body) for f in decl.split(",")]
    assert tuple(fields) == _abi.PARAM_FIELDS


def test_product_package_never_imports_oracle():
    import os
    root = os.path.dirname(_abi.HERE)
    for dirpath, _, files in os.walk(os.path.join(root, "cylinder_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "/root/reference" not in src or f.endswith((".cu", ".cuh", ".h", ".py")) and "sys.path" not in src
