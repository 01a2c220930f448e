#!/usr/bin/env python
"""End-to-end time of the one-shot C-ABI call mcs_piqmc_anneal on the cfg3 workload (pinned host buffers in and
out), optionally with an explicit window plan: MCS_WINDOWS=1024,1024,1024,1024 python benchmarks/e2e_probe.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montecarlosolvers_b200 as mcs  # noqa: E402
from bench import load_instance  # noqa: E402

nbs, _ = load_instance()
inst = mcs.Instance(nbs)
R, N, P, S = int(os.environ.get("R", "4096")), 6400, 64, int(os.environ.get("S", "1000"))
host = mcs.empty_pinned((R, N, P), np.int8)
s0 = (2 * np.random.RandomState(0).randint(2, size=(R, N, 1)) - 1).astype(np.int8)
e_host = mcs.empty_pinned((R, P), np.float64)
A, B = np.linspace(3, 1e-8, S), np.ones(S)
L = mcs._lib.load()
ts = []
for it in range(3):
    host[...] = s0
    t0 = time.perf_counter()
    mcs._lib.check(L.mcs_piqmc_anneal(inst._h, mcs._lib.dptr(A), mcs._lib.dptr(B), S, 1, 1.0 / P, host.ctypes.data,
                                      R, P, 0, 5 + it, 0, mcs._lib.dptr(e_host)))
    ts.append(time.perf_counter() - t0)
print("windows=%s: %.1f ms per call, %.4g attempts/s end to end (mean best E %.3f)" % (
    os.environ.get("MCS_WINDOWS", "default"), 1e3 * min(ts[1:]), R * S * P * N / min(ts[1:]),
    float(e_host.min(axis=1).mean())))
