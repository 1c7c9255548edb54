"""Build inclusivegan_b200/libb200knn.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

nvcc cross-compiles without a GPU, so this runs on the CPU build box; the resulting .so travels to the
GPU box with the repo snapshot.  `python -m inclusivegan_b200.build [--force] [--verbose]`.
"""
import hashlib
import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "libb200knn.so")
_STAMP = os.path.join(_HERE, ".libb200knn.stamp")
SOURCES = ["b200knn.cu"]
# every header under csrc/ takes part in the digest (a forgotten entry here once shipped a stale library)
HEADERS = sorted(f for f in os.listdir(_CSRC) if f.endswith((".cuh", ".h"))) + [os.path.join(_ROOT, "include", "b200knn.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-std=c++17", "-O3", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "-shared",
    "-Xptxas", "-v",
    "--expt-relaxed-constexpr",
]


def _nvcc():
    cand = [os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"]
    for c in cand:
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found (looked at $NVCC, PATH, /usr/local/cuda/bin)")


def _digest():
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        p = f if os.path.isabs(f) else os.path.join(_CSRC, f)
        with open(p, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile if sources changed since the last build; returns the path of the shared library."""
    dig = _digest()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(_STAMP):
        with open(_STAMP) as fh:
            if fh.read().strip() == dig:
                return LIB_PATH
    # the image exports CC/CXX=/opt/gcc/bin/* wrappers; let nvcc use the system host compiler
    host_cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    cmd = [_nvcc()] + NVCC_FLAGS + ["-ccbin", host_cxx, "-o", LIB_PATH] + [os.path.join(_CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = res.stdout
    with open(os.path.join(_HERE, ".build.log"), "w") as fh:
        fh.write(" ".join(cmd) + "\n" + log)
    if verbose or res.returncode != 0:
        sys.stderr.write(log)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed (exit %d); see inclusivegan_b200/.build.log" % res.returncode)
    with open(_STAMP, "w") as fh:
        fh.write(dig)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
