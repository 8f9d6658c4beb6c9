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

NVCC_COMMON = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-O2,-ffp-contract=off"]
# IEEE arithmetic as written, no a*b+c contraction -> plain operators round identically to the fp32 oracle; the
# candidate loop of the fused shading kernel opts out explicitly (fmaf, MUFU approximations; csrc/common.cuh)
NVCC_EXACT = ["-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false"]


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
    units = [(CSRC / "api.cu", NVCC_EXACT), (CSRC / "generic.cu", NVCC_EXACT), (CSRC / "kat.cu", NVCC_EXACT), (CSRC / "winner_cr.cu", NVCC_EXACT), (CSRC / "bvh_gpu.cu", NVCC_EXACT), (CSRC / "bvh_build.cpp", [])]
    deps = [u for u, _ in units] + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list(CSRC.glob("*.inc")) + [PKG.parent / "include" / "risltc_cuda.h"]
    if force or _stale(CUDA_LIB, deps):
        env = {**os.environ, "CC": "gcc", "CXX": "g++"}
        objdir = PKG / "build"
        objdir.mkdir(exist_ok=True)
        procs, objs = [], []
        for src, flags in units:
            obj = objdir / (src.stem + ".o")
            objs.append(str(obj))
            cmd = [_nvcc()] + NVCC_COMMON + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", str(obj), str(src), "-ccbin", "/usr/bin/g++"]
            procs.append((cmd, subprocess.Popen(cmd, env=env)))
        for cmd, p in procs:
            if p.wait() != 0:
                raise subprocess.CalledProcessError(p.returncode, cmd)
        subprocess.check_call([_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(CUDA_LIB)] + objs + ["-ccbin", "/usr/bin/g++"], env=env)
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
