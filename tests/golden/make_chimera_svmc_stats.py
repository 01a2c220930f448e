#!/usr/bin/env python
"""Reference statistics of the spin-vector solvers on the BASELINE cfg4 graph (Chimera C16, 2048 rotors, J = +-1).

Protocol (shortened to S schedule steps): theta_0 = pi/2, A = 3 (1 - s), B = s,
s = linspace(1e-3, 1, S), temp = 0.1, mcsteps = 1; read r = one call of svmc.SpinVectorMonteCarlo[TF] after
srand(3000 + r) and np.random.seed(3000 + r); observable = H(A_last, B_last) = B sum J cos cos - A sum sin of the
final angles (oracle.svmc_energy).  Sweeps run through the CPU oracle, which tests/test_oracle_vs_reference.py pins
bit-exactly to the compiled reference; --check-ref N re-runs the first N reads through the compiled reference.

Also a small Noisy (time-dependent couplings) case on a 6x6 torus with fields: 256 reads of NoisySVMC / NoisySVMCTF.

Output: tests/golden/chimera_svmc_ref_stats.json
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import oracle as orc  # noqa: E402
from bench import chimera_instance  # noqa: E402
from tests import instances as inst  # noqa: E402

S = 100


def sched(n=S):
    s = np.linspace(1e-3, 1.0, n)
    return (3.0 * (1 - s)).copy(), s.copy()


def noisy_tables(n=40, seed=11):
    """Time-dependent table for the Noisy solvers: the 6x6 torus couplings (with fields) plus 5 % noise per step."""
    _, nbs = inst.torus(6, seed=7, fields=True)
    rng = np.random.RandomState(seed)
    tabs = np.repeat(nbs[None], n, axis=0).copy()
    tabs[..., 1] *= 1.0 + 0.05 * rng.normal(size=tabs[..., 1].shape)
    return tabs


def _init():
    global NBS, NOISY
    NBS = chimera_instance(16)
    NOISY = noisy_tables()


def one(args):
    r, tf = args
    A, B = sched()
    v = np.full(NBS.shape[0], np.pi / 2)
    np.random.seed(3000 + r)
    fn = orc.SpinVectorMonteCarloTF if tf else orc.SpinVectorMonteCarlo
    fn(A, B, 1, 0.1, v, NBS, rng=3000 + r)
    return orc.svmc_energy(A[-1], B[-1], v, NBS)


def one_noisy(args):
    r, tf = args
    A, B = sched(NOISY.shape[0])
    v = np.full(NOISY.shape[1], np.pi / 2)
    np.random.seed(4000 + r)
    fn = orc.NoisySVMCTF if tf else orc.NoisySVMC
    fn(A, B, 1, 0.1, v, NOISY, rng=4000 + r)
    return orc.svmc_energy(A[-1], B[-1], v, NOISY[-1])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=256)
    ap.add_argument("--procs", type=int, default=os.cpu_count())
    ap.add_argument("--check-ref", type=int, default=2)
    a = ap.parse_args()
    t0 = time.time()
    _init()
    out = {"protocol": __doc__, "S": S, "reps": a.reps, "cells": {}}
    with mp.Pool(a.procs, initializer=_init) as pool:
        for tf in (0, 1):
            out["cells"]["svmc%s_chimera16" % ("_tf" if tf else "")] = pool.map(one, [(r, tf) for r in range(a.reps)])
            out["cells"]["noisy_svmc%s_torus6" % ("_tf" if tf else "")] = pool.map(one_noisy,
                                                                                 [(r, tf) for r in range(a.reps)])
            print("tf=%d done (%.0fs)" % (tf, time.time() - t0), flush=True)
    if a.check_ref:
        import ctypes
        import importlib
        from oracle import build_ref
        build_ref.build(verbose=False)
        if build_ref.import_ref() is not None:
            libc = ctypes.CDLL(None)
            rsv = importlib.import_module("solvers.svmc")
            A, B = sched()
            for r in range(a.check_ref):
                for tf, fn in ((0, rsv.SpinVectorMonteCarlo), (1, rsv.SpinVectorMonteCarloTF)):
                    v = np.full(NBS.shape[0], np.pi / 2)
                    libc.srand(3000 + r)
                    np.random.seed(3000 + r)
                    fn(A, B, 1, 0.1, v, NBS)
                    assert orc.svmc_energy(A[-1], B[-1], v, NBS) == out["cells"]["svmc%s_chimera16" % (
                        "_tf" if tf else "")][r]
            out["checked_against_compiled_reference"] = a.check_ref
            print("compiled-reference spot check OK")
    out["summary"] = {k: {"mean": float(np.mean(v)), "sd": float(np.std(v, ddof=1)), "n": len(v)}
                      for k, v in out["cells"].items()}
    with open(os.path.join(HERE, "chimera_svmc_ref_stats.json"), "w") as f:
        json.dump(out, f, indent=0)
    print(json.dumps(out["summary"], indent=1))


if __name__ == "__main__":
    main()
