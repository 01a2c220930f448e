#!/usr/bin/env python
"""SA colour passes of larger batches: one word per thread on one stream (MCS_SA_WPT=1) against one-warp CTAs that take
all the words of their site and chunk, two chunks on two streams (default); 80x80, 200 temperatures."""
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
S = 200
sched = np.linspace(3.0, 0.0, S)
for R in (2048, 3072, 4096, 8192, 32768):
    row = {"R": R}
    for name, env in (("one_word", {"MCS_SA_WPT": "1"}), ("multi_1stream", {"MCS_STREAMS": "1"}), ("multi", {})):
        os.environ.update(env)
        st = mcs.State(inst, mcs._lib.KIND_SA, R, 1)
        st.init_random(1)
        st.sa_sweeps(sched, 1, seed=3)
        inst.synchronize()
        best = 1e9
        for rep in range(3):
            inst.timer_start()
            st.sa_sweeps(sched, 1, seed=3)
            best = min(best, inst.timer_stop())
        row[name] = float("%.4g" % (R * S * N / (best * 1e-3)))
        st.close()
        for k in env:
            os.environ.pop(k, None)
    print(json.dumps(row), flush=True)
