#!/usr/bin/env python
"""General-degree (non-LUT) PIQMC / SA kernels on a Chimera C16 graph with local fields (degree 6 + field = 7
planes > 6) and on a degree-12 circulant graph: attempts/s."""
import os
import sys

import numpy as np
import scipy.sparse as sps

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montecarlosolvers_b200 as mcs  # noqa: E402
from benchmarks.bench_configs import chimera  # noqa: E402


def circulant(n, offs, seed=0):
    rng = np.random.default_rng(seed)
    J = sps.dok_matrix((n, n))
    for i in range(n):
        for o in offs:
            j = (i + o) % n
            if (i, j) not in J and (j, i) not in J:
                J[i, j] = float(rng.uniform(-1, 1))
    return mcs.tools.GenerateNeighbors(n, J, 2 * len(offs))


J, _ = chimera(16)
for i in range(J.shape[0]):
    J[i, i] = 0.1 * ((i * 7919) % 13 - 6)
cases = [("chimera C16 + fields", mcs.tools.GenerateNeighbors(J.shape[0], J, 7)),
         ("circulant deg 12, N=2048", circulant(2048, (1, 2, 3, 5, 8, 13)))]
R, P, S = 2048, int(os.environ.get("P", "64")), 20
for name, nbs in cases:
    inst = mcs.Instance(nbs)
    st = mcs.State(inst, mcs._lib.KIND_PIQMC, R, P)
    st.init_random(1)
    A, B = np.linspace(3.0, 1e-8, S), np.ones(S)
    st.piqmc_sweeps(A, B, 1, 1.0 / P, seed=7)
    inst.synchronize()
    inst.timer_start()
    st.piqmc_sweeps(A, B, 1, 1.0 / P, seed=7, sweep_offset=S)
    ms = inst.timer_stop()
    print("PIQMC %s (lut=%s, colours=%d): %.4g attempts/s" % (name, inst.lut_kernels, inst.ncolors,
                                                             R * S * P * inst.nspins / (ms * 1e-3)))
    st.close()
    sa = mcs.State(inst, mcs._lib.KIND_SA, 32768, 1)
    sa.init_random(1)
    sched = np.linspace(3.0, 0.0, 50)
    sa.sa_sweeps(sched, 1, seed=7)
    inst.synchronize()
    inst.timer_start()
    sa.sa_sweeps(sched, 1, seed=7, sweep_offset=50)
    ms = inst.timer_stop()
    print("SA    %s: %.4g attempts/s" % (name, 32768 * 50 * inst.nspins / (ms * 1e-3)))
    sa.close()
