#!/usr/bin/env python
"""Words per thread of the P = 64 PIQMC pass kernel (wW:n = W-warp CTAs, up to n words per thread): cfg3 shape,
200 schedule steps."""
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
A, B = np.linspace(3.0, 1e-8, S), np.ones(S)
for P in (64,):
    for R in (512, 4096):
        row = {"P": P, "R": R}
        for wpt in ("w4:1", "w4:16", "w1:16", "w1:64", None):
            os.environ.pop("MCS_WPT", None)
            os.environ.pop("MCS_WPT_WARPS", None)
            if wpt and wpt.startswith("w"):
                os.environ["MCS_WPT_WARPS"] = wpt[1:].split(":")[0]
                os.environ["MCS_WPT"] = wpt.split(":")[1]
            elif wpt:
                os.environ["MCS_WPT"] = wpt
            st = mcs.State(inst, mcs._lib.KIND_PIQMC, R, P)
            st.init_random(1)
            st.piqmc_sweeps(A, B, 1, 1.0 / P, seed=3)
            inst.synchronize()
            best = 1e9
            for rep in range(3):
                inst.timer_start()
                st.piqmc_sweeps(A, B, 1, 1.0 / P, seed=3)
                best = min(best, inst.timer_stop())
            row["wpt=%s" % (wpt or "default")] = float("%.4g" % (R * S * P * N / (best * 1e-3)))
            st.close()
        print(json.dumps(row), flush=True)
