#!/usr/bin/env python
"""Fixed-order fp64 energy kernels: chain (MCS_ENERGY_CHAIN=1) and by table (default when rows have at
most four off-diagonal entries): GPU time at several batch sizes, results compared bit for bit."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montecarlosolvers_b200 as mcs  # noqa: E402
from bench import load_instance  # noqa: E402

nbs, _ = load_instance()
inst = mcs.Instance(nbs)
for kind, P in ((mcs._lib.KIND_PIQMC, 64), (mcs._lib.KIND_PIQMC, 20), (mcs._lib.KIND_SA, 1)):
    for R in (1, 32, 512, 1024, 4096):
        st = mcs.State(inst, kind, R, P)
        st.init_random(1)
        res = {}
        for mode in ("chain", "table"):
            if mode != "table":
                os.environ["MCS_ENERGY_CHAIN"] = "1"
            else:
                os.environ.pop("MCS_ENERGY_CHAIN", None)
            e = st.energies()
            best = 1e9
            for rep in range(3):
                inst.timer_start()
                e = st.energies()
                best = min(best, inst.timer_stop())
            res[mode] = (best, e.copy())
        assert np.array_equal(res["chain"][1], res["table"][1])
        print("kind=%d P=%d R=%d: chain %.3f ms, by table %.3f ms (incl. the D2H of the energies)" %
              (kind, P, R, res["chain"][0], res["table"][0]), flush=True)
        st.close()
