#!/usr/bin/env python
"""One short cluster-resident SA schedule (224 restarts, 50 temperatures = 100 colour passes in ONE launch) for the
ncu capture of sa_cluster_kernel (profiles/capture.sh)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montecarlosolvers_b200 as mcs  # noqa: E402
from bench import load_instance  # noqa: E402

nbs, _ = load_instance()
inst = mcs.Instance(nbs)
st = mcs.State(inst, mcs._lib.KIND_SA, 224, 1)
st.init_random(1)
sched = np.linspace(3.0, 0.0, 50)
for rep in range(3):
    inst.timer_start()
    st.sa_sweeps(sched, 1, seed=3)
    ms = inst.timer_stop()
print("224 restarts, 100 colour passes: %.3f ms, %.2f us per pass, %d launch(es)" % (ms, 10 * ms, 1))
