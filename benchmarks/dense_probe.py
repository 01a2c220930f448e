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
A = np.linspace(3.0, 1.0, 2)
st.piqmc_sweeps(A, np.ones(2), 1, 1.0 / P, global_moves=True, seed=2)
inst.synchronize()
inst.timer_start()
st.piqmc_sweeps(A, np.ones(2), 1, 1.0 / P, global_moves=True, seed=2, sweep_offset=2)
print("ms per sweep", inst.timer_stop() / 2)
