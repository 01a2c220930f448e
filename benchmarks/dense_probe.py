#!/usr/bin/env python
"""Small cfg5-shaped run of the dense path for profiling (ncu -k regex:dense_block)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montecarlosolvers_b200 as mcs  # noqa: E402

n, P, R = 2048, 32, int(os.environ.get("R", "128"))
rng = np.random.default_rng(0)
Jm = np.triu(rng.normal(size=(n, n)) / np.sqrt(n), 1)
full = Jm + Jm.T
nb = np.zeros((n, n - 1, 2))
for i in range(n):
    idx = np.delete(np.arange(n), i)
    nb[i, :, 0] = idx
    nb[i, :, 1] = full[i, idx]
inst = mcs.Instance(nb)
st = mcs.State(inst, mcs._lib.KIND_PIQMC, R, P)
st.init_random(1)
S = int(os.environ.get("S", "20"))
A = np.linspace(3.0, 1.0, S)
st.piqmc_sweeps(A[:2], np.ones(2), 1, 1.0 / P, global_moves=True, seed=2)
inst.synchronize()
best = 1e30
for rep in range(3):
    inst.timer_start()
    st.piqmc_sweeps(A, np.ones(S), 1, 1.0 / P, global_moves=True, seed=2, sweep_offset=2 + rep * S)
    best = min(best, inst.timer_stop() / S)
e = st.energies()
print("R=%d: %.4f ms per sweep, %.3e attempts/s, mean energy %.2f" % (R, best, R * P * n / (best * 1e-3), e.mean()))
