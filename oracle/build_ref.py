#!/usr/bin/env python
"""Build the UNMODIFIED reference solvers (Cython) into oracle/_ref/ -- TEST INFRASTRUCTURE ONLY.

The reference (dtoconnor/MonteCarloSolvers) ships four Cython modules under
/root/reference/solvers/{qmc,sa,svmc,tools}.pyx.  They do not build as shipped against
NumPy >= 2 (`np.int_t` was removed from numpy's .pxd) and the shipped .c files are
Cython 0.29 output that cannot compile against CPython 3.12.

Recipe:  read each .pyx where it lies under /root/reference, apply the
one-token mechanical patch `np.int_t -> np.int64_t` (C `long` on Linux *is* int64, so
semantics are unchanged) in a scratch directory under /tmp, cythonize with
language_level=2 (the shipped C was generated with Py2 division semantics,
qmc.c:12 `CYTHON_FUTURE_DIVISION 0`), compile with setuptools' default flags (-O2, NO
-fopenmp: setup.py:10-11,17-18,24-25 have it commented out) and copy ONLY the resulting
shared objects to oracle/_ref/solvers/.  No reference source is copied into the repository;
oracle/_ref/ is git-ignored (it still travels to the GPU box with gpurun).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
load what this produces.
"""
import glob
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MCS_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(HERE, "_ref")
MODULES = ("qmc", "sa", "svmc", "tools")

SETUP_PY = r"""
from setuptools import setup
from setuptools.extension import Extension
from Cython.Build import cythonize
import numpy
exts = [Extension("solvers.%s" % m, ["solvers/%s.pyx" % m], include_dirs=[numpy.get_include()],
                  define_macros=[("NPY_NO_DEPRECATED_API", "NPY_1_7_API_VERSION")])
        for m in @MODULES@]
setup(name="MCS_ref", ext_modules=cythonize(exts, compiler_directives={"language_level": 2}),
      script_args=["build_ext", "--inplace"])
"""


def have_ref():
    return all(os.path.isfile(os.path.join(REF, "solvers", m + ".pyx")) for m in MODULES)


def built():
    return all(glob.glob(os.path.join(OUT, "solvers", m + ".*.so")) for m in MODULES)


def build(force=False, verbose=True):
    """Returns True when oracle/_ref holds importable reference modules."""
    if built() and not force:
        return True
    if not have_ref():
        if verbose:
            print("[build_ref] %s not present; keeping whatever is prebuilt in %s" % (REF, OUT))
        return built()
    tmp = tempfile.mkdtemp(prefix="mcs_ref_build_")
    try:
        os.makedirs(os.path.join(tmp, "solvers"))
        for m in MODULES:
            with open(os.path.join(REF, "solvers", m + ".pyx"), "r", encoding="utf-8") as f:
                src = f.read()
            src = src.replace("np.int_t", "np.int64_t")
            if m == "qmc":
                # Wolff experiments allocate `cluster` as np.intc into an np.int_t buffer
                # (qmc.pyx:685,860,1074,1080,1310,1531) which only ever matched on Windows.
                src = src.replace("dtype=np.intc", "dtype=np.int64")
                # `cimport openmp` needs no OpenMP at build time but keep it harmless.
            with open(os.path.join(tmp, "solvers", m + ".pyx"), "w", encoding="utf-8") as f:
                f.write(src)
        open(os.path.join(tmp, "solvers", "__init__.py"), "w").close()
        with open(os.path.join(tmp, "setup.py"), "w") as f:
            f.write(SETUP_PY.replace("@MODULES@", repr(MODULES)))
        env = dict(os.environ)
        env.setdefault("CFLAGS", "")
        r = subprocess.run([sys.executable, "setup.py"], cwd=tmp, env=env,
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            if verbose:
                print(r.stdout[-4000:])
            return False
        os.makedirs(os.path.join(OUT, "solvers"), exist_ok=True)
        open(os.path.join(OUT, "solvers", "__init__.py"), "w").close()
        for so in glob.glob(os.path.join(tmp, "solvers", "*.so")):
            shutil.copy2(so, os.path.join(OUT, "solvers", os.path.basename(so)))
        if verbose:
            print("[build_ref] built:", sorted(os.path.basename(p) for p in
                                                glob.glob(os.path.join(OUT, "solvers", "*.so"))))
        return built()
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def import_ref():
    """Import the compiled reference package (`solvers`) from oracle/_ref. Returns module or None."""
    if not built():
        return None
    if OUT not in sys.path:
        sys.path.insert(0, OUT)
    import importlib
    try:
        return importlib.import_module("solvers")
    except Exception:
        return None


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("[build_ref] ok" if ok else "[build_ref] FAILED")
    sys.exit(0 if ok else 1)
