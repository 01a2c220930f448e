#!/usr/bin/env python
"""Latency regime: cfg2 (SA, 1024 restarts) and a single reference-style PIQMC call.  Prints, per case, the host
time spent ENQUEUEING the launches (the call returns before the GPU is done), the GPU time (CUDA events) and the
launch count: enqueue ~ GPU time means the host launch rate is the limit, not the kernels.

    python benchmarks/small_batch_probe.py
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montecarlosolvers_b200 as mcs  # noqa: E402
from bench import load_instance  # noqa: E402

nbs, _ = load_instance()
inst = mcs.Instance(nbs)
N = inst.nspins


def case(name, st, run, attempts):
    run()
    inst.synchronize()
    best = None
    for rep in range(3):
        l0 = inst.launch_count() if hasattr(inst, "launch_count") else 0
        inst.timer_start()
        t0 = time.perf_counter()
        run()
        t_enq = (time.perf_counter() - t0) * 1e3
        ms = inst.timer_stop()
        l1 = inst.launch_count() if hasattr(inst, "launch_count") else 0
        if best is None or ms < best[0]:
            best = (ms, t_enq, l1 - l0)
    print(json.dumps({"case": name, "gpu_ms": best[0], "host_enqueue_ms": best[1], "launches": best[2],
                      "us_per_launch": 1e3 * best[0] / max(1, best[2]), "attempts_per_s": attempts / (best[0] * 1e-3)}),
          flush=True)


for R in (1024, 32, 1):
    st = mcs.State(inst, mcs._lib.KIND_SA, R, 1)
    st.init_random(1)
    sched = np.linspace(3.0, 0.0, 1000)
    case("SA R=%d, 1000 temperatures" % R, st, lambda: st.sa_sweeps(sched, 1, seed=3), R * 1000 * N)
    st.close()
for R, P in ((1, 20), (32, 20), (1, 64)):
    st = mcs.State(inst, mcs._lib.KIND_PIQMC, R, P)
    st.init_random(1)
    A, B = np.linspace(3.0, 1e-8, 354), np.ones(354)
    case("PIQMC-global R=%d P=%d, tau=354" % (R, P), st,
         lambda: st.piqmc_sweeps(A, B, 1, 1.0 / P, global_moves=True, seed=3), R * 354 * P * N)
    st.close()
