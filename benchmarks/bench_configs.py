#!/usr/bin/env python
"""Secondary measurements for the BASELINE.json configs that are NOT the driver's bench line
(bench.py measures configs[2]).  Prints one JSON object per config.

  cfg1  examples/santoro80.py protocol: 80x80, P=20, QuantumAnnealGlobal, single reference-style call + batch
  cfg2  sa.Anneal, 80x80, 1024 restarts
  cfg4  SVMC on a Chimera C16 graph (2048 rotors), 2048 reads
"""
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sps

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import montecarlosolvers_b200 as mcs  # noqa: E402
from bench import load_instance  # noqa: E402


def chimera(m=16, seed=0):
    """Chimera C_m: m x m cells of K_{4,4}; vertical (left) and horizontal (right) inter-cell couplers."""
    rng = np.random.default_rng(seed)
    n = 8 * m * m
    J = sps.dok_matrix((n, n))

    def q(r, c, side, k):
        return ((r * m + c) * 2 + side) * 4 + k

    for r in range(m):
        for c in range(m):
            for a in range(4):
                for b in range(4):
                    J[q(r, c, 0, a), q(r, c, 1, b)] = float(rng.choice([-1.0, 1.0]))
                if r + 1 < m:
                    J[q(r, c, 0, a), q(r + 1, c, 0, a)] = float(rng.choice([-1.0, 1.0]))
                if c + 1 < m:
                    J[q(r, c, 1, a), q(r, c + 1, 1, a)] = float(rng.choice([-1.0, 1.0]))
    return J, mcs.tools.GenerateNeighbors(n, J, 6)


def timed(inst, fn, reps=3):
    fn()
    inst.synchronize()
    best = 1e30
    for _ in range(reps):
        inst.timer_start()
        fn()
        best = min(best, inst.timer_stop())
    return best


def main():
    out = []
    nbs, name = load_instance()
    inst = mcs.Instance(nbs)
    # ---- cfg1: P = 20, global moves
    P, tau = 20, 354
    A, B = np.linspace(3.0, 1e-8, tau), np.ones(tau)
    for R in (1, 32, 256, 4096):
        st = mcs.State(inst, mcs._lib.KIND_PIQMC, R, P)
        st.init_random(1)
        ms = timed(inst, lambda: st.piqmc_sweeps(A, B, 1, 1.0 / P, global_moves=True, seed=2))
        out.append({"config": "cfg1 PIQMC-global 80x80 P=20 tau=%d" % tau, "replicas": R, "ms": ms,
                    "attempts_per_s": R * tau * P * 6400 / (ms * 1e-3)})
        st.close()
    # single reference-style call through the drop-in, host arrays, wall clock
    conf = np.tile((2 * np.random.RandomState(0).randint(2, size=6400) - 1).astype(np.int64), (P, 1)).T
    mcs.qmc.QuantumAnnealGlobal(A, B, 1, 1.0 / P, conf.copy(), nbs, 1, seed=3)
    t0 = time.perf_counter()
    mcs.qmc.QuantumAnnealGlobal(A, B, 1, 1.0 / P, conf, nbs, 1, seed=3)
    dt = time.perf_counter() - t0
    out.append({"config": "cfg1 single reference-style call qmc.QuantumAnnealGlobal([N,P] int64), wall", "ms": dt * 1e3,
                "attempts_per_s": tau * P * 6400 / dt})
    # ---- cfg2: SA 1024 restarts
    tau = 1000
    sched = np.linspace(3.0, 0.0, tau)
    for R in (1024, 32768):
        st = mcs.State(inst, mcs._lib.KIND_SA, R, 1)
        st.init_random(1)
        ms = timed(inst, lambda: st.sa_sweeps(sched, 1, seed=2))
        out.append({"config": "cfg2 SA 80x80 tau=%d" % tau, "replicas": R, "ms": ms,
                    "attempts_per_s": R * tau * 6400 / (ms * 1e-3)})
        st.close()
    # ---- cfg4: SVMC on Chimera C16
    _, cn = chimera(16)
    ci = mcs.Instance(cn)
    s = np.linspace(1e-3, 1.0, 1000)
    A4, B4 = 3.0 * (1 - s), s
    for tf in (0, 1):
        st = mcs.State(ci, mcs._lib.KIND_SVMC, 2048, 1)
        st.init_random(0)
        ms = timed(ci, lambda: st.svmc_sweeps(A4, B4, 1, 0.1, tf=bool(tf), seed=5))
        out.append({"config": "cfg4 SVMC%s Chimera C16 (N=2048, %d colours), 2048 reads, 1000 sweeps" % (
            "-TF" if tf else "", ci.ncolors), "ms": ms, "attempts_per_s": 2048 * 1000 * 2048 / (ms * 1e-3)})
        st.close()
    # ---- cfg5: dense SK N = 2048, P = 32, R replicas (R P columns), local fields on the tensor cores
    n = 2048
    rng = np.random.default_rng(0)
    Jm = np.triu(rng.normal(size=(n, n)) / np.sqrt(n), 1)
    nb5 = np.zeros((n, n - 1, 2))
    full = Jm + Jm.T
    for i in range(n):
        idx = np.delete(np.arange(n), i)
        nb5[i, :, 0] = idx
        nb5[i, :, 1] = full[i, idx]
    di = mcs.Instance(nb5)
    P5 = 32
    A5, B5 = np.linspace(3.0, 1e-8, 20), np.ones(20)
    for R5 in (128, 512):
        st = mcs.State(di, mcs._lib.KIND_PIQMC, R5, P5)
        st.init_random(1)
        ms = timed(di, lambda: st.piqmc_sweeps(A5, B5, 1, 1.0 / P5, global_moves=True, seed=2), reps=2)
        cols = R5 * P5
        out.append({"config": "cfg5 dense SK N=2048 P=32 PIQMC-global, blocked tensor-core sweeps, 20 sweeps",
                    "replicas": R5, "columns": cols, "ms": ms, "ms_per_sweep": ms / 20,
                    "attempts_per_s": 20.0 * n * cols / (ms * 1e-3),
                    "field_gemm_tflops_bf16x2": 20 * 2 * 2.0 * n * n * cols / (ms * 1e-3) / 1e12})
        st.close()
    di.use_dense(False)
    st = mcs.State(di, mcs._lib.KIND_PIQMC, 128, P5)
    st.init_random(1)
    ms = timed(di, lambda: st.piqmc_sweeps(A5[:2], B5[:2], 1, 1.0 / P5, global_moves=True, seed=2), reps=1)
    out.append({"config": "cfg5 same instance through the general coloured kernel (one colour class per site), 2 sweeps",
                "replicas": 128, "ms": ms, "attempts_per_s": 2.0 * n * 128 * P5 / (ms * 1e-3)})
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
