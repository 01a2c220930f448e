#!/usr/bin/env python
"""In-kernel cycle split of sa_cluster_kernel (MCS_CLUSTER_PROF=1): item / arrive + tables / CTA barrier / cluster wait."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montecarlosolvers_b200 as mcs  # noqa: E402
from bench import load_instance  # noqa: E402

nbs, _ = load_instance()
inst = mcs.Instance(nbs)
sched = np.linspace(3.0, 0.0, 1000)
os.environ["MCS_CLUSTER_PROF"] = "1"
os.environ["MCS_CLUSTER_VERBOSE"] = "1"
os.environ["MCS_CLUSTER"] = "1"
for R, cs, wc, mcsteps in ((1, 16, 0, 1), (32, 8, 0, 1), (32, 16, 0, 1), (128, 16, 0, 1), (1024, 8, 2, 1), (1024, 8, 3, 1), (1024, 16, 5, 1),
                           (1024, 4, 1, 1), (1024, 16, 5, 4), (2048, 16, 0, 1)):
    os.environ["MCS_CLUSTER_SIZE"] = str(cs)
    os.environ.pop("MCS_CLUSTER_WORDS", None)
    if wc:
        os.environ["MCS_CLUSTER_WORDS"] = str(wc)
    st = mcs.State(inst, mcs._lib.KIND_SA, R, 1)
    st.init_random(1)
    st.sa_sweeps(sched[:1000 // mcsteps], mcsteps, seed=3)
    inst.synchronize()
    inst.timer_start()
    st.sa_sweeps(sched[:1000 // mcsteps], mcsteps, seed=3)
    ms = inst.timer_stop()
    print("R=%d csize=%d words=%d mcsteps=%d: %.3f ms, %.2f us per pass" % (R, cs, wc, mcsteps, ms, ms / 2.0), flush=True)
    st.close()
