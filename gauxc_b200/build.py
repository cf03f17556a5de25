"""In-tree build of libgauxc_b200.so (nvcc, sm_100a only) and of the CPU oracle.

nvcc cross-compiles without a GPU, so this runs in the CPU-only build container; the .so
travels to the GPU box with the repo snapshot (it is git-ignored, not gpurun-ignored).
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libgauxc_b200.so")

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fopenmp,-Wall,-Wno-unknown-pragmas",
          "-Xptxas", "-v", "--expt-relaxed-constexpr"]

SOURCES = [
    "host/core.cxx", "host/grid.cxx", "host/load_balancer.cxx", "host/hdf5_io.cxx", "host/c_api.cxx",
    "host/device_integrator.cu",
    "cuda/collocation.cu", "cuda/fused.cu", "cuda/vxc.cu", "cuda/ssf_weights.cu", "cuda/probe.cu", "cuda/lb_screen.cu", "cuda/exc_grad.cu",
]


def _deps_mtime():
    m = 0.0
    for d, _, fs in os.walk(CSRC):
        for f in fs:
            if f.endswith((".hpp", ".cuh", ".h", ".inc")):
                m = max(m, os.path.getmtime(os.path.join(d, f)))
    m = max(m, os.path.getmtime(os.path.join(ROOT, "include", "gauxc_b200.h")))
    return m


def _compile(src, hdr_m, verbose, defs=(), bdir=None):
    obj = os.path.join(bdir or BUILD, src.replace("/", "_") + ".o")
    srcp = os.path.join(CSRC, src)
    if os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(srcp), hdr_m):
        return obj, ""
    cmd = [NVCC] + ARCH + COMMON + ["-D" + d for d in defs] + ["-x", "cu", "-c", srcp, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    log = r.stdout + r.stderr
    with open(obj + ".log", "w") as f:
        f.write(log)
    return obj, log


def build_library(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    if force:
        for f in os.listdir(BUILD):
            os.remove(os.path.join(BUILD, f))
    hdr_m = _deps_mtime()
    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        res = list(ex.map(lambda s: _compile(s, hdr_m, verbose), SOURCES))
    objs = [o for o, _ in res]
    if verbose:
        for _, log in res:
            if log:
                print(log)
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-Xcompiler", "-fopenmp", "-lgomp", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


def build_variant(suffix, defs):
    """Diagnostic build of the library with extra -D macros (timing experiments, e.g. GXB_KNOCKOUT);
    written next to the product library as libgauxc_b200_<suffix>.so and loaded only when
    GAUXC_B200_LIB points at it."""
    bdir = os.path.join(BUILD, suffix)
    os.makedirs(bdir, exist_ok=True)
    hdr_m = _deps_mtime()
    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = [o for o, _ in ex.map(lambda s: _compile(s, hdr_m, False, defs, bdir), SOURCES)]
    out = os.path.join(HERE, "libgauxc_b200_%s.so" % suffix)
    cmd = [NVCC] + ARCH + ["-shared", "-o", out] + objs + ["-Xcompiler", "-fopenmp", "-lgomp", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return out


def build_driver():
    """tools/standalone_driver.cxx -> gauxc_b200/standalone_driver (plain g++ over the C++ facade + C ABI)."""
    src = os.path.join(ROOT, "tools", "standalone_driver.cxx")
    exe = os.path.join(HERE, "standalone_driver")
    hdrs = [os.path.join(ROOT, "include", h) for h in ("gauxc_b200.h", "gauxc_b200.hpp")]
    newest = max(os.path.getmtime(f) for f in [src, LIB] + hdrs)
    if os.path.exists(exe) and os.path.getmtime(exe) > newest:
        return exe
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
           "-L", HERE, "-lgauxc_b200", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("standalone_driver build failed:\n" + r.stdout + r.stderr)
    return exe


def build_oracle(force=False):
    """Compile oracle/ (CPU restatement) and, when /root/reference is present, oracle/_ref."""
    odir = os.path.join(ROOT, "oracle")
    r = subprocess.run(["make", "-C", odir] + (["-B"] if force else []), capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    return os.path.join(odir, "liboracle.so")


if __name__ == "__main__":
    force = "--force" in sys.argv
    print(build_library(force=force, verbose="-v" in sys.argv))
    print(build_driver())
    if "--no-oracle" not in sys.argv:
        print(build_oracle(force=force))
