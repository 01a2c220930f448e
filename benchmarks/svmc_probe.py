#!/usr/bin/env python
"""cfg4-shaped SVMC run for profiling (ncu -k regex:svmc_pass)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montecarlosolvers_b200 as mcs  # noqa: E402
from benchmarks.bench_configs import chimera  # noqa: E402

_, cn = chimera(16)
ci = mcs.Instance(cn)
R = int(os.environ.get("R", "2048"))
st = mcs.State(ci, mcs._lib.KIND_SVMC, R, 1)
st.init_random(0)
s = np.linspace(1e-3, 1.0, 200)
st.svmc_sweeps(3.0 * (1 - s), s, 1, 0.1, seed=5)
ci.synchronize()
ci.timer_start()
st.svmc_sweeps(3.0 * (1 - s), s, 1, 0.1, seed=5, sweep_offset=200)
ms = ci.timer_stop()
print("ms", ms, "attempts/s", R * 200 * 2048 / (ms * 1e-3))
