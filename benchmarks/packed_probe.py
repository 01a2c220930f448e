#!/usr/bin/env python
"""Packed mode (even P <= 20, odd P 3 .. 21): packed words resident in HBM for the duration of a sweep call (default)
against gathering the members in every pass (MCS_PACK_GATHER=1) and against two world lines per word (MCS_NO_PACK=1);
cfg1 shape (80x80, tau = 354)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montecarlosolvers_b200 as mcs  # noqa: E402
from bench import load_instance  # noqa: E402

nbs, _ = load_instance()
inst = mcs.Instance(nbs)
N = inst.nspins
tau = 354
A, B = np.linspace(3.0, 1e-8, tau), np.ones(tau)
for P, glob in ((20, True), (32, True), (10, True)):
    for R in (4096, 512):
        row = {"P": P, "R": R, "global_moves": glob}
        for name, env in (("two_per_word", {"MCS_NO_PACK": "1"}), ("gather", {"MCS_PACK_GATHER": "1"}),
                          ("resident_one_word_threads", {"MCS_PACK_ONE_WORD": "1"}), ("resident", {})):
            os.environ.update(env)
            st = mcs.State(inst, mcs._lib.KIND_PIQMC, R, P)
            st.init_random(1)
            st.piqmc_sweeps(A, B, 1, 1.0 / P, global_moves=glob, seed=3)
            inst.synchronize()
            best = 1e9
            for rep in range(2):
                inst.timer_start()
                st.piqmc_sweeps(A, B, 1, 1.0 / P, global_moves=glob, seed=3)
                best = min(best, inst.timer_stop())
            row[name] = float("%.4g" % (R * tau * P * N / (best * 1e-3)))
            st.close()
            for k in env:
                os.environ.pop(k, None)
        print(json.dumps(row), flush=True)
