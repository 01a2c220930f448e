#!/usr/bin/env python
"""Reference residual-energy statistics at the BASELINE cfg3 shape (80x80, P = 64): the P = 64 cell of tier (c).

Same protocol and seeds as make_santoro_stats.py (the 256 pre-annealed states of santoro_preannealed.npz,
srand(2000 + r) before the call), with P = 64, PT = 1 (temp = 1/64), Gamma 3 -> 1e-8 in tau steps, one sweep per
step, observable = best-slice residual energy per spin.  Sweeps run through the CPU oracle (pinned bit-exactly to
the compiled reference by tests/test_oracle_vs_reference.py); --check-ref N re-runs the first N reps through the
compiled reference itself and demands equality.

Output: tests/golden/santoro_ref_stats_p64.json
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
from tests import instances as inst  # noqa: E402

N, P = 6400, 64


def _init():
    global NBS, EGS
    _, NBS, _, EGS = inst.santoro()


def qa(args):
    r, tau, glob, s = args
    confs = np.tile(s.astype(np.int64), (P, 1)).T.copy(order="F")
    fn = orc.QuantumAnnealGlobal if glob else orc.QuantumAnneal
    fn(np.linspace(3.0, 1e-8, tau), np.ones(tau), 1, 1.0 / P, confs, NBS, 1, rng=2000 + r)
    e = min(orc.ising_energy(np.ascontiguousarray(confs[:, k]), NBS) for k in range(P))
    return (e - EGS) / N


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=256)
    ap.add_argument("--procs", type=int, default=os.cpu_count())
    ap.add_argument("--taus", type=int, nargs="+", default=[60, 146])
    ap.add_argument("--check-ref", type=int, default=1)
    a = ap.parse_args()
    t0 = time.time()
    _init()
    pre = np.load(os.path.join(HERE, "santoro_preannealed.npz"))
    states = np.where(np.unpackbits(pre["packed"], axis=1)[:, :N] > 0, 1, -1).astype(np.int8)[:a.reps]
    out = {"protocol": __doc__, "N": N, "P": P, "reps": a.reps, "cells": {}}
    with mp.Pool(a.procs, initializer=_init) as pool:
        for glob in (0, 1):
            for tau in a.taus:
                res = pool.map(qa, [(r, tau, glob, states[r]) for r in range(a.reps)])
                name = "qmc%s_P%d_tau%d" % ("_global" if glob else "", P, tau)
                out["cells"][name] = res
                print("%s mean=%.5f sd=%.5f (%.0fs)" % (name, np.mean(res), np.std(res, ddof=1), time.time() - t0),
                      flush=True)
    if a.check_ref:
        import ctypes
        import importlib
        from oracle import build_ref
        build_ref.build(verbose=False)
        if build_ref.import_ref() is not None:
            libc = ctypes.CDLL(None)
            rqmc = importlib.import_module("solvers.qmc")
            tau = a.taus[0]
            for r in range(a.check_ref):
                confs = np.tile(states[r].astype(np.int64), (P, 1)).T.copy(order="F")
                libc.srand(2000 + r)
                rqmc.QuantumAnneal(np.linspace(3.0, 1e-8, tau), np.ones(tau), 1, 1.0 / P, confs, NBS, 1)
                e = min(orc.ising_energy(np.ascontiguousarray(confs[:, k]), NBS) for k in range(P))
                assert (e - EGS) / N == out["cells"]["qmc_P%d_tau%d" % (P, tau)][r]
            out["checked_against_compiled_reference"] = a.check_ref
            print("compiled-reference spot check OK")
    out["summary"] = {k: {"mean": float(np.mean(v)), "sd": float(np.std(v, ddof=1)), "n": len(v)}
                      for k, v in out["cells"].items()}
    with open(os.path.join(HERE, "santoro_ref_stats_p64.json"), "w") as f:
        json.dump(out, f, indent=0)
    print(json.dumps(out["summary"], indent=1))


if __name__ == "__main__":
    main()
