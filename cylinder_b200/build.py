"""Build libcylinder_b200.so (CUDA kernels + C-ABI) in-tree with nvcc for sm_100a.

    python -m cylinder_b200.build            # incremental
    python -m cylinder_b200.build --force

The shared library lands in cylinder_b200/lib/ (git-ignored, travels to the GPU box with the snapshot).
nvcc cross-compiles without a GPU.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIBNAME = "libcylinder_b200.so"

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
          "-Xptxas", "-v", "--expt-relaxed-constexpr"]
# per-file extra flags
EXTRA = {
    # parity kernel: the reference's operation order, no FMA contraction (numpy/CPython do not fuse)
    "k_replay.cu": ["-fmad=false"],
}


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _newest_header_mtime():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(HERE, "..", "include", "cylinder_b200.h"))
    hs.append(os.path.abspath(__file__))
    return max(os.path.getmtime(h) for h in hs)


def source_sha():
    """sha256 over the kernel sources (csrc/*.cu, *.cuh and the C header): ties profiler counters kept under profiles/ to the
    build they were taken from (bench.py refuses counters of another source state)."""
    import hashlib
    h = hashlib.sha256()
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh")))
    files.append(os.path.join(HERE, "..", "include", "cylinder_b200.h"))
    for f in files:
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def lib_path():
    return os.path.join(LIBDIR, LIBNAME)


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ and link the shared library.  Returns the library path."""
    nvcc = nvcc_path()
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    hdr_m = _newest_header_mtime()
    objs, rebuilt = [], False
    procs = []
    for src in sources():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJDIR, src[:-3] + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_m):
            cmd = [nvcc] + ARCH + COMMON + EXTRA.get(src, []) + os.environ.get("CYL_EXTRA_NVCC_FLAGS", "").split() + ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd))
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
            rebuilt = True
    log = []
    for src, pr in procs:
        out, _ = pr.communicate()
        log.append("== %s\n%s" % (src, out))
        if pr.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed on %s" % src)
    if log:
        with open(os.path.join(OBJDIR, "ptxas.log"), "w") as f:
            f.write("\n".join(log))
        if verbose:
            print("\n".join(log))
    lib = lib_path()
    if rebuilt or not os.path.exists(lib):
        cmd = [nvcc] + ARCH + ["-shared", "-o", lib] + objs + ["-Xcompiler", "-fPIC", "-cudart", "static", "-ldl", "-lpthread"]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    return lib


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv or "--verbose" in sys.argv)
    print(p)
