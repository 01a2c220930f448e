"""Compile libmcs_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m montecarlosolvers_b200.build [--force] [--verbose]

The shared object is written next to this file (git-ignored, but it travels to the GPU box with
gpurun).  mcs_exact.cu is additionally compiled with -fmad=false: the validation kernels must not
contract a*b+c (they already route every fp64 op through __dmul_rn/__dadd_rn).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libmcs_b200.so")
OBJ = os.path.join(HERE, "build")
SOURCES = ["mcs_instance.cu", "mcs_piqmc.cu", "mcs_sa.cu", "mcs_svmc.cu", "mcs_refdyn.cu", "mcs_dense.cu", "mcs_cluster.cu", "mcs_exact.cu", "mcs_api.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
          "--expt-relaxed-constexpr"]
EXTRA = {"mcs_exact.cu": ["-fmad=false"]}


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.isfile(c):
            return c
    raise RuntimeError("nvcc not found; libmcs_b200.so cannot be built (there is no CPU fallback)")


def _deps():
    d = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    d.append(os.path.join(os.path.dirname(HERE), "include", "mcs_b200.h"))
    d.append(os.path.abspath(__file__))
    return d


def up_to_date():
    if not os.path.isfile(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(p) <= t for p in _deps() if os.path.isfile(p))


def build(force=False, verbose=False, ptxas_info=False):
    """Build the shared library if any source is newer; returns its path."""
    if not force and up_to_date():
        return OUT
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        cmd = [nvcc] + ARCH + COMMON + EXTRA.get(src, []) + (["-Xptxas", "-v"] if ptxas_info else []) + \
              ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose or ptxas_info:
            sys.stdout.write(out)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libmcs_b200.so")
    cmd = [nvcc] + ARCH + ["-shared", "-o", OUT] + objs
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, ptxas_info="--ptxas" in sys.argv))
