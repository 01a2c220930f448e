#!/usr/bin/env python
"""Reference residual-energy statistics on the shipped 80x80 instance (parity tier c).

Protocol = examples/santoro80.py:250-298 with explicit seeds:
  rep r: state = 2*RandomState(r).randint(2, size=6400) - 1
  CA : sa.Anneal(linspace(3.0, 0.0, tau), 1, state, nbs)  after srand(1000+r)           (:258-262)
  pre: sa.Anneal(linspace(3.0, 1.0, 41), 100, state, nbs) after srand(1000+r)           (:284-285)
  QA : confs = tile(pre-annealed state, P); qmc.QuantumAnneal[Global](linspace(3.0, 1e-8, tau),
       ones, 1, 1.0/P, confs, nbs, 1) after srand(2000+r); E = min over slices            (:281-296)
  residual = (E - E_gs)/6400.
The sweeps are executed by the CPU oracle (oracle/mcs_oracle.c), which tests pin bit-exactly
to the compiled reference, so these ARE the reference's numbers for these seeds; with
--check-ref N the first N reps of every cell are re-run through the compiled reference
(oracle/_ref) and must agree exactly.

Outputs: tests/golden/santoro_ref_stats.json  (per-cell mean / sd / n / raw residuals)
         tests/golden/santoro_preannealed.npz (256 pre-annealed states, bit-packed) so that the
         GPU test can start PIQMC from the very same states.
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

N = 6400
P = 20
TAUS = (60, 146, 354, 857)


def _init():
    global NBS, EGS
    _, NBS, _, EGS = inst.santoro()


def _state(r):
    return (2 * np.random.RandomState(r).randint(2, size=N) - 1).astype(np.int64)


def ca(args):
    r, tau = args
    s = _state(r)
    orc.Anneal(np.linspace(3.0, 0.0, tau), 1, s, NBS, rng=1000 + r)
    return (orc.ising_energy(s, NBS) - EGS) / N


def pre(r):
    s = _state(r)
    orc.Anneal(np.linspace(3.0, 1.0, 41), 100, s, NBS, rng=1000 + r)
    return s.astype(np.int8)


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
    ap.add_argument("--procs", type=int, default=6)
    ap.add_argument("--check-ref", type=int, default=2)
    a = ap.parse_args()
    t0 = time.time()
    _init()
    out = {"protocol": __doc__, "N": N, "P": P, "reps": a.reps, "cells": {}}
    with mp.Pool(a.procs, initializer=_init) as pool:
        for tau in TAUS:
            res = pool.map(ca, [(r, tau) for r in range(a.reps)])
            out["cells"]["sa_tau%d" % tau] = res
            print("sa tau=%d mean=%.5f sd=%.5f (%.0fs)" % (tau, np.mean(res), np.std(res, ddof=1), time.time() - t0), flush=True)
        states = pool.map(pre, range(a.reps))
        st = np.stack(states)
        np.savez_compressed(os.path.join(HERE, "santoro_preannealed.npz"),
                            packed=np.packbits(st > 0, axis=1), reps=a.reps)
        res = [(orc.ising_energy(s.astype(np.int64), NBS) - EGS) / N for s in states]
        out["cells"]["preanneal"] = res
        print("pre-anneal mean=%.5f (%.0fs)" % (np.mean(res), time.time() - t0), flush=True)
        for glob in (1, 0):
            for tau in TAUS:
                res = pool.map(qa, [(r, tau, glob, states[r]) for r in range(a.reps)])
                out["cells"]["qmc%s_P%d_tau%d" % ("_global" if glob else "", P, tau)] = res
                print("qmc glob=%d tau=%d mean=%.5f sd=%.5f (%.0fs)" % (
                    glob, tau, np.mean(res), np.std(res, ddof=1), time.time() - t0), flush=True)
    if a.check_ref:
        import ctypes
        import importlib
        from oracle import build_ref
        build_ref.build(verbose=False)
        if build_ref.import_ref() is not None:
            libc = ctypes.CDLL(None)
            rqmc = importlib.import_module("solvers.qmc")
            rsa = importlib.import_module("solvers.sa")
            for r in range(a.check_ref):
                s = _state(r)
                libc.srand(1000 + r)
                rsa.Anneal(np.linspace(3.0, 0.0, 60), 1, s, NBS)
                assert (orc.ising_energy(s, NBS) - EGS) / N == out["cells"]["sa_tau60"][r]
                confs = np.tile(states[r].astype(np.int64), (P, 1)).T.copy(order="F")
                libc.srand(2000 + r)
                rqmc.QuantumAnnealGlobal(np.linspace(3.0, 1e-8, 60), np.ones(60), 1, 1.0 / P, confs, NBS, 1)
                e = min(orc.ising_energy(np.ascontiguousarray(confs[:, k]), NBS) for k in range(P))
                assert (e - EGS) / N == out["cells"]["qmc_global_P20_tau60"][r]
            out["checked_against_compiled_reference"] = a.check_ref
            print("compiled-reference spot check OK")
    summary = {k: {"mean": float(np.mean(v)), "sd": float(np.std(v, ddof=1)), "n": len(v)}
               for k, v in out["cells"].items()}
    out["summary"] = summary
    with open(os.path.join(HERE, "santoro_ref_stats.json"), "w") as f:
        json.dump(out, f, indent=0)
    print(json.dumps(summary, indent=1))


if __name__ == "__main__":
    main()
