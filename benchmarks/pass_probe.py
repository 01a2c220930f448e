#!/usr/bin/env python
"""cfg3-shaped PIQMC pass probe: ms per colour-pass launch and a hash of the final state.

The hash lets a kernel change that must not alter any accept/reject decision (same Philox draws,
same thresholds) be verified bit for bit against the previous build:
    MCS_PIQMC_VARIANT=0 python benchmarks/pass_probe.py ; MCS_PIQMC_VARIANT=1 python benchmarks/pass_probe.py
Environment: R (4096), P (64), S (100 schedule steps), GLOBAL (0/1).
"""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montecarlosolvers_b200 as mcs  # noqa: E402
from bench import load_instance  # noqa: E402

nbs, _ = load_instance()
inst = mcs.Instance(nbs)
R = int(os.environ.get("R", "4096"))
P = int(os.environ.get("P", "64"))
S = int(os.environ.get("S", "100"))
glob = bool(int(os.environ.get("GLOBAL", "0")))
N = inst.nspins
st = mcs.State(inst, mcs._lib.KIND_PIQMC, R, P)
st.init_random(1)
A, B = np.linspace(3.0, 1e-8, S), np.ones(S)
st.piqmc_sweeps(A, B, 1, 1.0 / P, global_moves=glob, seed=7)  # warm-up (also part of the hashed trajectory)
inst.synchronize()
best = 1e30
for it in range(3):
    inst.timer_start()
    st.piqmc_sweeps(A, B, 1, 1.0 / P, global_moves=glob, seed=7, sweep_offset=S * (it + 1))
    best = min(best, inst.timer_stop())
e = st.energies()
h = hashlib.sha256(np.ascontiguousarray(e).tobytes()).hexdigest()[:16]
print("variant=%s R=%d P=%d S=%d global=%d: %.1f us/launch, %.4g attempts/s, mean best-slice E %.6f, hash %s" % (
    os.environ.get("MCS_PIQMC_VARIANT", "-"), R, P, S, glob, 1e3 * best / (2 * S),
    R * S * P * N / (best * 1e-3), float(e.min(axis=1).mean()), h))
