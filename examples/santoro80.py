#!/usr/bin/env python
"""The reference's experiment (examples/santoro80.py: Martonak-Santoro-Tosatti 80x80, classical annealing vs
path-integral quantum annealing, residual energy vs annealing time) on the B200 drop-in.

The reference script is stale (wrong import paths, calls that no longer match the solver signatures,
SURVEY.md section 2 row 8); this is the same protocol (santoro80.py:250-298) written against the current call
surface, with the 45 repetitions of every (tau, P) cell run as ONE batched call.

    python examples/santoro80.py [--taus 60 146 354 857] [--reps 45] [--instance PATH]
"""
import argparse
import os
import sys
import time

import numpy as np
import scipy.sparse as sps

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from montecarlosolvers_b200 import qmc, sa, tools  # noqa: E402


def load(path):
    """(isingJ, E_gs) from the reference's instance file (1-based `i j J`) or the repo's fixture copy of it."""
    if path and os.path.isfile(path):
        d = np.loadtxt(path)
        i, j, v = d[:, 0].astype(int) - 1, d[:, 1].astype(int) - 1, d[:, 2]
        egs = -1.58051667679 * 6400  # santoro_80x80_answer.txt:24
    else:
        fx = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                                  "santoro80.npz"))
        i, j, v, egs = fx["i"], fx["j"], fx["J_file"], float(fx["e_gs_per_spin"]) * 6400
    J = sps.dok_matrix((6400, 6400))
    for a, b, val in zip(i, j, v):
        J[int(a), int(b)] = -1.0 * val  # santoro80.py:244
    return J, egs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--taus", type=int, nargs="+", default=[60, 146, 354, 857])
    ap.add_argument("--reps", type=int, default=45)
    ap.add_argument("--slices", type=int, nargs="+", default=[5, 10, 20, 40])
    ap.add_argument("--instance", default="/root/reference/examples/ising_instances/santoro_80x80.txt")
    a = ap.parse_args()
    J, egs = load(a.instance)
    nbs = tools.GenerateNeighbors(6400, J, 4)
    R, PT = a.reps, 1.0
    start = np.stack([2 * np.random.RandomState(r).randint(2, size=6400) - 1 for r in range(R)]).astype(np.int8)
    print("tau      CA        " + "  ".join("PIQMC P=%-3d" % p for p in a.slices))
    for tau in a.taus:
        t0 = time.time()
        s = start.copy()
        e = sa.Anneal(np.linspace(3.0, 0.0, tau), 1, s, nbs, seed=tau, energies=True)  # santoro80.py:258-262
        row = ["%-7d  %.5f" % (tau, (e.mean() - egs) / 6400)]
        for P in a.slices:
            s = start.copy()
            sa.Anneal(np.linspace(3.0, PT, 41), 100, s, nbs, seed=1000 + tau)           # pre-anneal, :284-285
            confs = np.ascontiguousarray(np.repeat(s[:, :, None], P, axis=2))           # np.tile(state,(P,1)).T, :286
            e = qmc.QuantumAnnealGlobal(np.linspace(3.0, 1e-8, tau), np.ones(tau), 1, PT / P, confs, nbs, 1,
                                        seed=2000 + tau, energies=True)                 # :287-289
            row.append("%.5f    " % ((e.min(axis=1).mean() - egs) / 6400))             # best slice, :290-298
        print("  ".join(row) + "   (%.1f s)" % (time.time() - t0), flush=True)


if __name__ == "__main__":
    main()
