#!/usr/bin/env python
"""Throughput of the reference-dynamics mode (resident state, CUDA events): cfg3 / cfg1 / cfg2 shapes.

    python benchmarks/refdyn_probe.py [R] [sweeps]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import montecarlosolvers_b200 as mcs  # noqa: E402
from bench import load_instance  # noqa: E402


def main():
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    nbs, _ = load_instance()
    inst = mcs.Instance(nbs)
    inst.set_dynamics("reference")
    N = 6400
    for P, glob in ((64, False), (20, True)):
        st = mcs.State(inst, mcs._lib.KIND_PIQMC, R, P)
        st.init_random(1)
        A, B = np.linspace(3.0, 1e-8, 1000)[300:300 + S].copy(), np.ones(S)
        st.piqmc_sweeps(A[:1], B[:1], 1, 1.0 / P, global_moves=glob, seed=2)
        inst.synchronize()
        inst.timer_start()
        st.piqmc_sweeps(A, B, 1, 1.0 / P, global_moves=glob, seed=2)
        ms = inst.timer_stop()
        print(json.dumps({"mode": "reference dynamics", "solver": "PIQMC%s" % ("-global" if glob else ""), "R": R,
                          "P": P, "sweeps": S, "ms": ms, "attempts_per_s": R * S * P * N / (ms * 1e-3)}), flush=True)
        st.close()
    st = mcs.State(inst, mcs._lib.KIND_SA, R, 1)
    st.init_random(1)
    sched = np.linspace(3.0, 0.1, 8 * S)
    st.sa_sweeps(sched[:1], 1, seed=2)
    inst.synchronize()
    inst.timer_start()
    st.sa_sweeps(sched, 1, seed=2)
    ms = inst.timer_stop()
    print(json.dumps({"mode": "reference dynamics", "solver": "SA", "R": R, "sweeps": 8 * S, "ms": ms,
                      "attempts_per_s": R * 8 * S * N / (ms * 1e-3)}), flush=True)
    st.close()
    # cfg4 shape: rotors on Chimera C16
    from bench import chimera_instance
    ci = mcs.Instance(chimera_instance(16))
    ci.set_dynamics("reference")
    Rv = min(R, 2048)
    sv = mcs.State(ci, mcs._lib.KIND_SVMC, Rv, 1)
    s = np.linspace(1e-3, 1.0, 1000)[400:400 + 8 * S]
    A, B = (3.0 * (1 - s)).copy(), s.copy()
    for tf in (0, 1):
        sv.init_random(1)
        sv.svmc_sweeps(A[:1], B[:1], 1, 0.1, tf=tf, seed=2)
        ci.synchronize()
        ci.timer_start()
        sv.svmc_sweeps(A, B, 1, 0.1, tf=tf, seed=2)
        ms = ci.timer_stop()
        print(json.dumps({"mode": "reference dynamics", "solver": "SVMC%s" % ("-TF" if tf else ""), "R": Rv,
                          "sweeps": 8 * S, "ms": ms, "attempts_per_s": Rv * 8 * S * 2048 / (ms * 1e-3)}), flush=True)


if __name__ == "__main__":
    main()
