#!/usr/bin/env python
"""Residual-energy statistics of the COLOURED-ORDER oracle on the shipped 80x80 instance.

Same protocol, seeds, initial and pre-annealed states as make_santoro_stats.py, same fp64 visit arithmetic
and acceptance rule (oracle/mcs_oracle.c), but sites are visited checkerboard colour by colour and even
Trotter slices before odd ones -- the visiting order of the B200 production kernels.  The GPU kernels are
compared with THESE numbers at the standard-error level (tests/test_gpu_production.py); the difference
between these and santoro_ref_stats.json is the effect of the visiting order alone.
Output: tests/golden/santoro_colored_stats.json
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

N, P = 6400, 20
TAUS = (60, 146, 354, 857)
COLORS = ((np.arange(N) // 80 + np.arange(N) % 80) & 1).astype(np.int32)  # checkerboard (row + col) & 1


def _init():
    global NBS, EGS, PRE
    _, NBS, _, EGS = inst.santoro()
    pre = np.load(os.path.join(HERE, "santoro_preannealed.npz"))
    PRE = np.where(np.unpackbits(pre["packed"], axis=1)[:, :N] > 0, 1, -1).astype(np.int64)


def ca(args):
    r, tau = args
    s = (2 * np.random.RandomState(r).randint(2, size=N) - 1).astype(np.int64)
    orc.AnnealColored(np.linspace(3.0, 0.0, tau), 1, s, NBS, COLORS, rng=1000 + r)
    return (orc.ising_energy(s, NBS) - EGS) / N


def qa(args):
    r, tau, glob = args
    confs = np.tile(PRE[r], (P, 1)).T.copy(order="F")
    orc.QuantumAnnealColored(np.linspace(3.0, 1e-8, tau), np.ones(tau), 1, 1.0 / P, confs, NBS, COLORS,
                             global_moves=bool(glob), rng=2000 + r)
    e = min(orc.ising_energy(np.ascontiguousarray(confs[:, k]), NBS) for k in range(P))
    return (e - EGS) / N


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=256)
    ap.add_argument("--procs", type=int, default=6)
    a = ap.parse_args()
    t0 = time.time()
    out = {"protocol": __doc__, "N": N, "P": P, "reps": a.reps, "cells": {}}
    with mp.Pool(a.procs, initializer=_init) as pool:
        for tau in TAUS:
            res = pool.map(ca, [(r, tau) for r in range(a.reps)])
            out["cells"]["sa_tau%d" % tau] = res
            print("sa colored tau=%d mean=%.5f (%.0fs)" % (tau, np.mean(res), time.time() - t0), flush=True)
        for glob in (1, 0):
            for tau in TAUS:
                res = pool.map(qa, [(r, tau, glob) for r in range(a.reps)])
                out["cells"]["qmc%s_P%d_tau%d" % ("_global" if glob else "", P, tau)] = res
                print("qmc colored glob=%d tau=%d mean=%.5f (%.0fs)" % (glob, tau, np.mean(res), time.time() - t0),
                      flush=True)
    out["summary"] = {k: {"mean": float(np.mean(v)), "sd": float(np.std(v, ddof=1)), "n": len(v)}
                      for k, v in out["cells"].items()}
    with open(os.path.join(HERE, "santoro_colored_stats.json"), "w") as f:
        json.dump(out, f, indent=0)
    print(json.dumps(out["summary"], indent=1))


if __name__ == "__main__":
    main()
