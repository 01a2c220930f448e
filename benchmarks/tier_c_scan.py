#!/usr/bin/env python
"""Tier (c) z-scores of the reference-dynamics mode over several seeds (and, for contrast, the coloured mode).

    python benchmarks/tier_c_scan.py [nseeds] [dynamics]

For every cell of tests/golden/santoro_ref_stats.json (+ the P = 64 cells of santoro_ref_stats_p64.json) and every
seed: 256 anneals from the reference run's initial states, z = (mean_gpu - mean_ref) / combined SEM.  A sampler that
follows the reference's dynamics gives z ~ N(offset_cell, ~0.7) with offset_cell the fixture's own sampling
fluctuation (shared by all seeds), i.e. over cells and seeds mean(z) ~ 0, sd(z) ~ 1, |z| > 2 in ~5 % of the cases.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import montecarlosolvers_b200 as mcs  # noqa: E402
from tests import instances as inst  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")


def main():
    nseeds = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    dyn = sys.argv[2] if len(sys.argv) > 2 else "reference"
    _, nbs, _, e_gs = inst.santoro()
    ref = json.load(open(os.path.join(G, "santoro_ref_stats.json")))["cells"]
    p64 = os.path.join(G, "santoro_ref_stats_p64.json")
    ref64 = json.load(open(p64))["cells"] if os.path.isfile(p64) else {}
    pre = np.load(os.path.join(G, "santoro_preannealed.npz"))
    s_pre = np.where(np.unpackbits(pre["packed"], axis=1)[:, :6400] > 0, 1, -1).astype(np.int8)[:256]
    s_rand = np.stack([inst.random_spins(6400, r) for r in range(256)]).astype(np.int8)
    I = mcs.Instance(nbs)
    zs = {}
    for seed in range(nseeds):
        for tau in (60, 146, 354, 857):
            s = s_rand.copy()
            t0 = time.perf_counter()
            e = mcs.sa.Anneal(np.linspace(3.0, 0.0, tau), 1, s, I, seed=1000 * seed + tau, energies=True, dynamics=dyn)
            dt = time.perf_counter() - t0
            got = (e - e_gs) / 6400
            r = np.asarray(ref["sa_tau%d" % tau])
            z = (got.mean() - r.mean()) / np.sqrt(got.var(ddof=1) / got.size + r.var(ddof=1) / r.size)
            zs.setdefault("sa_tau%d" % tau, []).append(z)
            if seed == 0:
                print("sa tau=%d: %.3g attempts/s (wall, 256 anneals incl. copies)" % (tau, 256 * tau * 6400 / dt))
        for P, cells in ((20, ref), (64, ref64)):
            for name in sorted(k for k in cells if k.startswith("qmc")):
                tau, glob = int(name.split("tau")[1]), "global" in name
                confs = np.ascontiguousarray(np.repeat(s_pre[:, :, None], P, axis=2))
                fn = mcs.qmc.QuantumAnnealGlobal if glob else mcs.qmc.QuantumAnneal
                t0 = time.perf_counter()
                e = fn(np.linspace(3.0, 1e-8, tau), np.ones(tau), 1, 1.0 / P, confs, I, 1, seed=1000 * seed + tau + 7,
                       energies=True, dynamics=dyn)
                dt = time.perf_counter() - t0
                got = (e.min(axis=1) - e_gs) / 6400
                r = np.asarray(cells[name])
                z = (got.mean() - r.mean()) / np.sqrt(got.var(ddof=1) / got.size + r.var(ddof=1) / r.size)
                zs.setdefault(name, []).append(z)
                if seed == 0:
                    print("%s: %.3g attempts/s (wall, 256 anneals incl. copies)" % (name, 256 * tau * P * 6400 / dt))
    allz = []
    for name, z in zs.items():
        print("%-24s z = %s" % (name, " ".join("%+5.2f" % v for v in z)))
        allz += z
    allz = np.array(allz)
    print("dynamics=%s: %d z-scores, mean %+.2f, sd %.2f, |z|>2: %.1f %%, max |z| %.2f" % (
        dyn, allz.size, allz.mean(), allz.std(ddof=1), 100 * np.mean(np.abs(allz) > 2), np.abs(allz).max()))


if __name__ == "__main__":
    main()
