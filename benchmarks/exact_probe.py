#!/usr/bin/env python
"""Throughput of the bit-exact sequential-replay path (exact=True) at the cfg1 / cfg3 shapes.

    python benchmarks/exact_probe.py [R] [sweeps]

Prints one JSON object per shape: attempts/s of mcs_exact_qmc / mcs_exact_sa measured with the wall clock around
the C-ABI call (host buffers in and out: the call is synchronous), minus nothing.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import montecarlosolvers_b200 as mcs  # noqa: E402
from bench import load_instance  # noqa: E402


def main():
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    nbs, _ = load_instance()
    inst = mcs.Instance(nbs)
    N = 6400
    rng = np.random.RandomState(0)
    for P in (20, 64):
        confs = np.repeat((2 * rng.randint(2, size=(R, N, 1)) - 1).astype(np.int8), P, axis=2)
        A, B = np.linspace(3.0, 1e-8, S), np.ones(S)
        for rep in range(2):
            c = confs.copy()
            t0 = time.perf_counter()
            mcs.qmc.QuantumAnneal(A, B, 1, 1.0 / P, c, inst, 1, exact=True, libc_seed=1000)
            dt = time.perf_counter() - t0
        print(json.dumps({"path": "exact qmc.QuantumAnneal", "R": R, "P": P, "sweeps": S, "ms": dt * 1e3,
                          "attempts_per_s": R * S * P * N / dt}), flush=True)
    sv = (2 * rng.randint(2, size=(R, N)) - 1).astype(np.int8)
    sched = np.linspace(3.0, 0.1, 8 * S)
    for rep in range(2):
        c = sv.copy()
        t0 = time.perf_counter()
        mcs.sa.Anneal(sched, 1, c, inst, exact=True, libc_seed=1000)
        dt = time.perf_counter() - t0
    print(json.dumps({"path": "exact sa.Anneal", "R": R, "sweeps": 8 * S, "ms": dt * 1e3,
                      "attempts_per_s": R * 8 * S * N / dt}), flush=True)


if __name__ == "__main__":
    main()
