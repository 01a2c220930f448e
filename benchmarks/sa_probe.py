#!/usr/bin/env python
"""cfg2-shaped SA pass probe (80x80 Santoro couplings): ms per colour-pass launch at R restarts."""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montecarlosolvers_b200 as mcs  # noqa: E402
from bench import load_instance  # noqa: E402

nbs, _ = load_instance()
inst = mcs.Instance(nbs)
S = int(os.environ.get("S", "200"))
for R in [int(x) for x in os.environ.get("R", "1024,32768,131072").split(",")]:
    st = mcs.State(inst, mcs._lib.KIND_SA, R, 1)
    st.init_random(1)
    sched = np.linspace(3.0, 0.0, S)
    st.sa_sweeps(sched, 1, seed=7)
    inst.synchronize()
    best = 1e30
    for it in range(3):
        inst.timer_start()
        st.sa_sweeps(sched, 1, seed=7, sweep_offset=S * (it + 1))
        best = min(best, inst.timer_stop())
    e = st.energies()
    print("SA R=%d S=%d: %.2f us/launch, %.4g attempts/s, mean E %.4f, hash %s" % (
        R, S, 1e3 * best / (2 * S), R * S * inst.nspins / (best * 1e-3), float(e.mean()),
        hashlib.sha256(np.ascontiguousarray(e).tobytes()).hexdigest()[:12]))
    st.close()
