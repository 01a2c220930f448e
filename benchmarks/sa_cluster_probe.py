#!/usr/bin/env python
"""Cluster-resident SA schedule (sa_cluster_kernel) against one launch per colour pass, on the 80x80 instance:
GPU time of a 1000-temperature anneal for several batch sizes (default dispatch, forced cluster kernel, forced
multi-launch), and the wall time of the drop-in sa.Anneal call on one configuration.

    python benchmarks/sa_cluster_probe.py
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
sched = np.linspace(3.0, 0.0, 1000)


def run(R, env):
    for k in ("MCS_CLUSTER", "MCS_CLUSTER_SIZE", "MCS_CLUSTER_WORDS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    st = mcs.State(inst, mcs._lib.KIND_SA, R, 1)
    st.init_random(1)
    st.sa_sweeps(sched, 1, seed=3)
    inst.synchronize()
    best = None
    for rep in range(5):
        l0 = inst.launches
        inst.timer_start()
        st.sa_sweeps(sched, 1, seed=3)
        ms = inst.timer_stop()
        nl = inst.launches - l0
        best = ms if best is None else min(best, ms)
    sp = st.download_spins().astype(np.int64)
    h = int((sp * np.arange(1, N + 1)).sum())
    st.close()
    print(json.dumps({"R": R, "env": env, "gpu_ms": round(best, 4), "launches": nl, "us_per_pass": round(best / 2.0, 3),
                      "attempts_per_s": float("%.4g" % (R * 1000 * N / (best * 1e-3))), "state_hash": h}), flush=True)


for R in (1, 32, 128, 224, 448, 640, 896, 1024, 2048):
    run(R, {"MCS_CLUSTER": "0"})
    run(R, {"MCS_CLUSTER": "1"})
    run(R, {})
# the drop-in call on one configuration (host array in, in place), wall clock
for k in ("MCS_CLUSTER", "MCS_CLUSTER_SIZE", "MCS_CLUSTER_WORDS"):
    os.environ.pop(k, None)
conf = (2 * np.random.RandomState(0).randint(2, size=N) - 1).astype(np.int64)
for mode in ("0", None):
    if mode is not None:
        os.environ["MCS_CLUSTER"] = mode
    else:
        os.environ.pop("MCS_CLUSTER", None)
    for warm in range(2):  # (the first cluster launch of a process pays for the function attributes)
        mcs.sa.Anneal(sched, 1, conf.copy(), nbs, seed=1)
    t0 = time.perf_counter()
    for rep in range(5):
        mcs.sa.Anneal(sched, 1, conf.copy(), nbs, seed=1)
    print(json.dumps({"drop_in": "sa.Anneal, one configuration, 1000 temperatures", "MCS_CLUSTER": mode,
                      "wall_ms_per_call": round((time.perf_counter() - t0) / 5 * 1e3, 3)}), flush=True)
