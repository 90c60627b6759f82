"""In-tree build of libogb200.so (nvcc, sm_100a only) and of the test-only host
emulation library.  Used by __graft_entry__.build(); no JIT cache is involved, so the
built .so travels with the repository snapshot to the GPU box."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libogb200.so")
EMU_SRC = os.path.join(ROOT, "tests", "emu", "ogb_emu.cpp")
EMU_LIB = os.path.join(ROOT, "tests", "emu", "libogb_emu.so")

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-fmad=false",                      # no FMA contraction: c(x) rounds like numpy
              "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", "-cudart", "shared"]


def _newer(target, sources):
    if not os.path.isfile(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + \
        [os.path.join(ROOT, "include", "ogb200.h")]


def nvcc_path():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def build_lib(force=False, verbose=False):
    srcs = sources()
    if not force and _newer(LIB, srcs):
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
        ["-I", os.path.join(ROOT, "include"), "-I", CSRC, "-o", LIB,
         os.path.join(CSRC, "ogb_kernels.cu")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


def build_emu(force=False):
    """tests/emu: the same __host__ __device__ arithmetic compiled by g++ (test checker)."""
    srcs = sources() + [EMU_SRC]
    if not force and _newer(EMU_LIB, srcs):
        return EMU_LIB
    cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared",
           "-I", os.path.join(ROOT, "include"), "-o", EMU_LIB, EMU_SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + res.stdout + res.stderr)
    return EMU_LIB


if __name__ == "__main__":
    print(build_lib(force=True, verbose="-v" in sys.argv))
    print(build_emu(force=True))
