import os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
import montecarlosolvers_b200 as mcs
from bench import load_instance
nbs, _ = load_instance()
inst = mcs.Instance(nbs)
P, S = 20, 400
A, B = np.linspace(3.0, 1e-8, S), np.ones(S)
for R in (1, 32, 128, 512):
    st = mcs.State(inst, mcs._lib.KIND_PIQMC, R, P)
    st.init_random(1)
    st.piqmc_sweeps(A, B, 1, 1.0 / P, global_moves=True, seed=7)
    inst.synchronize()
    inst.timer_start()
    t0 = time.perf_counter()
    st.piqmc_sweeps(A, B, 1, 1.0 / P, global_moves=True, seed=7, sweep_offset=S)
    t1 = time.perf_counter()
    ms = inst.timer_stop()
    print("PIQMC P=20 R=%d: %.2f us/launch device, host enqueue %.2f us/launch" % (R, 1e3 * ms / (2 * S), 1e6 * (t1 - t0) / (2 * S)))
    st.close()
for R in (32, 1024):
    st = mcs.State(inst, mcs._lib.KIND_SA, R, 1)
    st.init_random(1)
    sched = np.linspace(3.0, 0.0, S)
    st.sa_sweeps(sched, 1, seed=7)
    inst.synchronize()
    inst.timer_start()
    t0 = time.perf_counter()
    st.sa_sweeps(sched, 1, seed=7, sweep_offset=S)
    t1 = time.perf_counter()
    ms = inst.timer_stop()
    print("SA R=%d: %.2f us/launch device, host enqueue %.2f us/launch" % (R, 1e3 * ms / (2 * S), 1e6 * (t1 - t0) / (2 * S)))
    st.close()
