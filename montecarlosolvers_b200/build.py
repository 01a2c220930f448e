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


def _sha(paths, extra=""):
    import hashlib
    h = hashlib.sha256(extra.encode())
    for p in paths:  # names, not absolute paths: the tree is copied to another directory on the GPU box
        with open(p, "rb") as f:
            h.update(os.path.basename(p).encode() + b"\0" + f.read())
    return h.hexdigest()


def _headers():
    hs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    return hs + [os.path.join(os.path.dirname(HERE), "include", "mcs_b200.h")]


def _flags(src):
    return ARCH + COMMON + EXTRA.get(src, [])


def _src_hash(src):
    """Content hash of everything one object file depends on (its source, every header, the flags)."""
    return _sha([os.path.join(CSRC, src)] + _headers(), " ".join(_flags(src)))


def _lib_hash():
    return _sha([], " ".join(_src_hash(s) for s in SOURCES))


def _read(path):
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return None


STAMP = OUT + ".srchash"


def up_to_date():
    """True when the shared object was built from exactly the sources on disk (content hash, not mtimes: the
    snapshot that carries the tree to the GPU box does not keep them)."""
    return os.path.isfile(OUT) and os.path.isdir(CSRC) and _read(STAMP) == _lib_hash()


def build(force=False, verbose=False, ptxas_info=False):
    """Build the shared library if any source changed (per-object: only what changed is recompiled); returns its
    path."""
    if not force and not ptxas_info and up_to_date():
        return OUT
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    # one builder at a time (several ranks of a torchrun launch may get here together): the others wait for the
    # lock and then find the library up to date
    import fcntl
    lock = open(os.path.join(OBJ, ".lock"), "w")
    fcntl.flock(lock, fcntl.LOCK_EX)
    try:
        if not force and not ptxas_info and up_to_date():
            return OUT
        return _build_locked(nvcc, force, verbose, ptxas_info)
    finally:
        fcntl.flock(lock, fcntl.LOCK_UN)
        lock.close()


def _build_locked(nvcc, force, verbose, ptxas_info):
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(obj)
        want = _src_hash(src)
        if not force and not ptxas_info and os.path.isfile(obj) and _read(obj + ".srchash") == want:
            continue
        cmd = [nvcc] + _flags(src) + (["-Xptxas", "-v"] if ptxas_info else []) + \
              ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        procs.append((src, obj, want,
                      subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, obj, want, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose or ptxas_info:
            sys.stdout.write(out)
        if p.returncode == 0:
            with open(obj + ".srchash", "w") as f:
                f.write(want)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libmcs_b200.so")
    tmp = OUT + ".tmp.%d" % os.getpid()
    cmd = [nvcc] + ARCH + ["-shared", "-o", tmp] + objs
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    os.replace(tmp, OUT)  # atomic: a concurrent loader never sees a half-written library
    with open(STAMP, "w") as f:
        f.write(_lib_hash())
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, ptxas_info="--ptxas" in sys.argv))
