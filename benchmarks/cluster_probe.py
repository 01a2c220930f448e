#!/usr/bin/env python
"""Swendsen-Wang cluster moves (mcs_cluster.cu): time per move on the 80x80 P=64 lattice and on the cfg5 SK shape."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montecarlosolvers_b200 as mcs  # noqa: E402
from bench import load_instance  # noqa: E402

nbs, _ = load_instance()
inst = mcs.Instance(nbs)
for R, P in ((256, 64), (32, 20)):
    st = mcs.State(inst, mcs._lib.KIND_PIQMC, R, P)
    st.init_random(1)
    st.cluster_moves(1.0, 1.0, 1.0 / P, nmoves=2, seed=3)
    inst.synchronize()
    inst.timer_start()
    st.cluster_moves(1.0, 1.0, 1.0 / P, nmoves=10, seed=3, sweep_offset=2)
    ms = inst.timer_stop() / 10
    print("SW 80x80 P=%d R=%d: %.3f ms per move, %.4g space-time sites/s" % (P, R, ms, R * P * inst.nspins / (ms * 1e-3)))
    st.close()
