#!/usr/bin/env python
"""Where the end-to-end time of one cfg3 anneal goes: H2D+pack, sweeps, unpack+D2H, energies."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montecarlosolvers_b200 as mcs  # noqa: E402
from bench import load_instance  # noqa: E402

nbs, _ = load_instance()
inst = mcs.Instance(nbs)
R, N, P, S = 4096, 6400, 64, int(os.environ.get("S", "100"))
host = mcs.empty_pinned((R, N, P), np.int8)
host[...] = (2 * np.random.RandomState(0).randint(2, size=(R, N, 1)) - 1).astype(np.int8)
st = mcs.State(inst, mcs._lib.KIND_PIQMC, R, P)
A, B = np.linspace(3, 1e-8, S), np.ones(S)
for it in range(2):
    t = [time.perf_counter()]
    st.upload_spins(host); inst.synchronize(); t.append(time.perf_counter())
    st.piqmc_sweeps(A, B, 1, 1.0 / P, seed=1); inst.synchronize(); t.append(time.perf_counter())
    st.download_spins(host); t.append(time.perf_counter())
    e = st.energies(); t.append(time.perf_counter())
    print("upload+pack %.1f ms | %d sweeps %.1f ms | unpack+download %.1f ms | energies %.1f ms" % (
        1e3 * (t[1] - t[0]), S, 1e3 * (t[2] - t[1]), 1e3 * (t[3] - t[2]), 1e3 * (t[4] - t[3])))
