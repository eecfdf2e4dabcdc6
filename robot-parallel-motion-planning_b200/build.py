"""Build recipe for libgmt_b200.so (hand-written CUDA for sm_100a + the C ABI of include/gmt_c_api.h).

nvcc cross-compiles without a GPU; the .so is built IN-TREE next to this file so that it travels
to the GPU box with the repository snapshot.  There is no other backend and no fallback: if the
library is missing, importing gmt_planner raises.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgmt_b200.so")
SOURCES = ["gmt_kernels.cu", "gmt_roadmap.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC,-fvisibility=hidden"]


def _deps():
    out = [os.path.abspath(__file__), os.path.join(HERE, "..", "include", "gmt_c_api.h")]
    for f in os.listdir(CSRC):
        if f.endswith((".cu", ".cuh", ".h")):
            out.append(os.path.join(CSRC, f))
    return out


def build(force=False, verbose=False):
    """Compile csrc/*.cu into libgmt_b200.so (skipped when the library is newer than its sources)."""
    if not force and os.path.exists(LIB):
        t = os.path.getmtime(LIB)
        if all(os.path.getmtime(d) <= t for d in _deps()):
            return LIB
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout)
    if verbose:
        print(res.stdout)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
