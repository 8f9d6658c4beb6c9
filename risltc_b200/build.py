"""In-tree builds: librisltc_cuda.so (nvcc, sm_100a) and librisltc_host.so (gcc, C99).

The shared objects are written next to this file so that they travel with the
repository snapshot to the GPU box (they are git-ignored, not gpurun-ignored)."""
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
HOST = PKG / "host"
CUDA_LIB = PKG / "librisltc_cuda.so"
HOST_LIB = PKG / "librisltc_host.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # IEEE arithmetic as written: no a*b+c contraction (parity with the fp32 oracle), exact div / sqrt
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off", "-shared",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target, sources):
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(s).stat().st_mtime > t for s in sources)


def build_cuda(force=False, verbose=False):
    sources = [CSRC / "api.cu", CSRC / "bvh_build.cpp"]
    deps = sources + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list(CSRC.glob("*.inc")) + [PKG.parent / "include" / "risltc_cuda.h"]
    if force or _stale(CUDA_LIB, deps):
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", str(CUDA_LIB)] + [str(s) for s in sources]
        env = {**os.environ, "CC": "gcc", "CXX": "g++"}
        subprocess.check_call(cmd + ["-ccbin", "/usr/bin/g++"], env=env)
    return CUDA_LIB


def build_host(force=False):
    sources = sorted(HOST.glob("*.c"))
    if not sources:
        return None
    deps = sources + list(HOST.glob("*.h")) + [PKG.parent / "include" / "risltc_cuda.h"]
    if force or _stale(HOST_LIB, deps):
        cmd = ["/usr/bin/gcc", "-std=c99", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", "-Wextra",
               "-I", str(PKG.parent / "include"), "-I", str(HOST), "-o", str(HOST_LIB)] + [str(s) for s in sources if s.name != "main.c"] + [
                   "-L", str(PKG), "-lrisltc_cuda", "-Wl,-rpath,$ORIGIN", "-lm"]
        subprocess.check_call(cmd)
        exe = ["/usr/bin/gcc", "-std=c99", "-O2", "-I", str(HOST), "-I", str(PKG.parent / "include"), "-o", str(PKG / "risltc"), str(HOST / "main.c"),
               "-L", str(PKG), "-lrisltc_host", "-lrisltc_cuda", "-Wl,-rpath,$ORIGIN", "-lm"]
        subprocess.check_call(exe)
    return HOST_LIB


def build_all(force=False, verbose=False):
    return build_cuda(force, verbose), build_host(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built", CUDA_LIB, HOST_LIB if HOST_LIB.exists() else "")
