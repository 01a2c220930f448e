#!/usr/bin/env python
"""Generate the committed golden fixtures from the REFERENCE ITSELF (compiled by
oracle/build_ref.py into oracle/_ref) and from the reference's shipped instance files.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
Outputs (small, committed):
  tests/golden/santoro80.npz       the shipped 80x80 instance (examples/ising_instances/
                                   santoro_80x80.txt, 1-based i j J) + exact ground state
                                   (santoro_80x80_answer.txt:24,36-256)
  tests/golden/traj_*.npz          inputs + outputs of reference calls after srand(s)/np.random.seed(s)
The long statistical fixture (reference residual energies on the Santoro instance) is made by
tests/golden/make_santoro_stats.py.
"""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import build_ref  # noqa: E402
from tests import instances as inst  # noqa: E402

REF = os.environ.get("MCS_REFERENCE_DIR", "/root/reference")
libc = ctypes.CDLL(None)


def santoro():
    d = np.loadtxt(os.path.join(REF, "examples/ising_instances/santoro_80x80.txt"))
    i = d[:, 0].astype(np.int32) - 1
    j = d[:, 1].astype(np.int32) - 1
    val = d[:, 2].astype(np.float64)
    up = []
    e_per_spin = None
    grab = False
    with open(os.path.join(REF, "examples/ising_instances/santoro_80x80_answer.txt")) as f:
        for line in f:
            if "Energy (per spin)" in line:
                e_per_spin = float(line.split(":")[1])
            if "CONFIGURATION_END" in line:
                grab = False
            if grab:
                up.extend(int(t) for t in line.split())
            if "CONFIGURATION_BEGIN" in line:
                grab = True
    gs = -np.ones(6400, dtype=np.int8)
    gs[np.array(up) - 1] = 1
    assert len(up) == 3184
    np.savez_compressed(os.path.join(HERE, "santoro80.npz"), i=i, j=j, J_file=val, ground_state=gs,
                        e_gs_per_spin=np.float64(e_per_spin))
    print("santoro80.npz: %d bonds, E_gs/N = %.11f" % (len(i), e_per_spin))


def trajectories():
    ok = build_ref.build(verbose=False)
    ref = build_ref.import_ref()
    assert ok and ref is not None, "compiled reference unavailable"
    import importlib
    qmc = importlib.import_module("solvers.qmc")
    sa = importlib.import_module("solvers.sa")
    svmc = importlib.import_module("solvers.svmc")
    tools = importlib.import_module("solvers.tools")

    # ---- qmc: torus with fields, P in {2,3,8,20}; local and global; Fortran-strided like the example
    _, nbs = inst.torus(6, seed=3, fields=True)
    n = nbs.shape[0]
    out = {"nbs": nbs}
    A = np.linspace(3.0, 1e-8, 12)
    B = np.linspace(0.2, 1.0, 12)
    out["A"], out["B"], out["mcsteps"] = A, B, 2
    for P in (2, 3, 8, 20):
        s0 = inst.random_spins(n, 5 + P)
        for glob in (0, 1):
            c = np.tile(s0, (P, 1)).T.copy(order="F")
            libc.srand(1000 + P)
            (qmc.QuantumAnnealGlobal if glob else qmc.QuantumAnneal)(A, B, 2, 1.0 / P, c, nbs, 1)
            out["P%d_g%d_in" % (P, glob)] = np.tile(s0, (P, 1)).T.astype(np.int8)
            out["P%d_g%d_out" % (P, glob)] = c.astype(np.int8)
            out["P%d_g%d_next_rand" % (P, glob)] = np.int64(libc.rand())
    np.savez_compressed(os.path.join(HERE, "traj_qmc_torus6.npz"), **out)

    # ---- qmc: irregular padded graph, C-order, independent slices at start
    J, nbs = inst.random_graph(40, 90, seed=2)
    c0 = (2 * np.random.RandomState(1).randint(2, size=(40, 5)) - 1).astype(np.int64)
    c = c0.copy()
    A = np.linspace(2.0, 0.01, 9)
    libc.srand(5)
    qmc.QuantumAnnealGlobal(A, np.ones(9), 3, 0.07, c, nbs, 1)
    k = np.arange(1, 5)
    lut = 0.05 * (np.pi / (5 * np.sin(np.pi * k / 5))) ** 2
    cd = c0.copy()
    libc.srand(6)
    qmc.DissipativeQuantumAnnealGlobal(A, np.ones(9), 2, 0.07, lut, cd, nbs, 1)
    np.savez_compressed(os.path.join(HERE, "traj_qmc_graph40.npz"), nbs=nbs, A=A, B=np.ones(9), mcsteps=3,
                        temp=0.07, conf_in=c0.astype(np.int8), conf_out=c.astype(np.int8), seed=5,
                        lut=lut, diss_out=cd.astype(np.int8), diss_seed=6, diss_mcsteps=2,
                        energies=np.array([tools.ClassicalIsingEnergy(c[:, q], J) for q in range(5)]))

    # ---- sa: schedule ending at T = 0 (santoro80.py:260)
    _, nbs = inst.torus(6, seed=4, fields=True)
    sched = np.linspace(3.0, 0.0, 30)
    s0 = inst.random_spins(36, 8)
    s = s0.copy()
    libc.srand(99)
    sa.Anneal(sched, 3, s, nbs)
    nxt = libc.rand()
    sm = s0.copy()
    libc.srand(4)
    np.random.seed(4)
    sa.AnnealMA(sched, 2, sm, nbs)
    np.savez_compressed(os.path.join(HERE, "traj_sa_torus6.npz"), nbs=nbs, sched=sched, mcsteps=3,
                        s_in=s0.astype(np.int8), s_out=s.astype(np.int8), seed=99, next_rand=np.int64(nxt),
                        ma_out=sm.astype(np.int8), ma_seed=4, ma_mcsteps=2)

    # ---- svmc: plain, TF, Compact, TFCompact
    _, nbs = inst.torus(5, seed=7, fields=True)
    sgrid = np.linspace(1e-3, 1.0, 20)
    A, B = 3.0 * (1 - sgrid), sgrid
    out = {"nbs": nbs, "A": A, "B": B, "mcsteps": 2, "temp": 0.1, "seed": 21}
    for name in ("SpinVectorMonteCarlo", "SpinVectorMonteCarloTF"):
        v = np.full(25, np.pi / 2)
        libc.srand(21)
        np.random.seed(21)
        getattr(svmc, name)(A, B, 2, 0.1, v, nbs)
        out[name] = v
    v = np.full((5, 25), np.pi / 2)
    libc.srand(21)
    np.random.seed(21)
    svmc.SpinVectorMonteCarloCompact(A, B, 2, 0.1, v, nbs)
    out["SpinVectorMonteCarloCompact"] = v
    v = np.full((4, 25), np.pi / 2)
    libc.srand(21)
    svmc.SpinVectorMonteCarloTFCompact(A, B, 2, 0.1, v, nbs)
    out["SpinVectorMonteCarloTFCompact"] = v
    np.savez_compressed(os.path.join(HERE, "traj_svmc_torus5.npz"), **out)
    print("trajectory fixtures written")


WOLFF = ("QuantumAnnealWCL", "DissaptiveQuantumAnnealWCL", "QuantumAnnealWC", "DissipativeQuantumAnnealWC2",
         "DissipativeQuantumAnnealWC3")


def wolff_cases():
    """(key, nbs, P, temp, bath strength, seed, mcsteps) of the Wolff-experiment fixtures"""
    _, t6 = inst.torus(6, seed=1, fields=True)
    _, g30 = inst.random_graph(30, 70, seed=3, fields=True)
    _, t5 = inst.torus(5, seed=2)
    return [("torus6f_P8", t6, 8, 0.5 / 8, 0.1, 5, 7), ("graph30f_P2", g30, 2, 0.1 / 2, 0.02, 7, 7),
            ("torus5_P5", t5, 5, 2.0 / 5, 0.5, 6, 7), ("graph30f_P16", g30, 16, 1.0 / 16, 0.05, 9, 4)]


def wolff():
    """Wolff-cluster experiments (qmc.pyx:612-1621) through the compiled reference (oracle/build_ref.py's dtype patch
    makes them runnable).  A case is kept only if the oracle's replay says the reference stayed inside its
    `cluster` buffer (it does not check; see oracle/mcs_oracle_wolff.c)."""
    ok = build_ref.build(verbose=False)
    ref = build_ref.import_ref()
    assert ok and ref is not None, "compiled reference unavailable"
    import importlib
    from oracle import oracle as orc
    qmc = importlib.import_module("solvers.qmc")
    out = {"A": np.linspace(2.0, 0.1, 6), "B": np.linspace(0.4, 1.0, 6)}
    kept = []
    for key, nbs, P, temp, alpha, seed, mcsteps in wolff_cases():
        n = nbs.shape[0]
        out[key + "_nbs"] = nbs
        lut = alpha * (np.pi / (P * np.sin(np.pi * np.arange(1, P) / P))) ** 2
        c0 = (2 * np.random.RandomState(seed).randint(2, size=(n, P)) - 1).astype(np.int64)
        if seed == 6:
            c0 = np.tile(inst.random_spins(n, 3), (P, 1)).T.copy().astype(np.int64)
        out[key + "_in"] = c0.astype(np.int8)
        out[key + "_par"] = np.array([P, temp, alpha, seed, mcsteps], dtype=np.float64)
        for name in WOLFF:
            diss = "iss" in name
            d = c0.copy()
            getattr(orc, name)(*((out["A"], out["B"], mcsteps, temp) + ((lut,) if diss else ()) + (d, nbs)), rng=seed)
            if orc.last_wolff_overrun:
                print("skipped (reference overruns its buffer):", key, name)
                continue
            c = c0.copy()
            libc.srand(seed)
            args = (out["A"], out["B"], mcsteps, temp) + ((lut,) if diss else ()) + (c, nbs)
            if name.endswith(("WC2", "WC3")):
                args += (1,)
            getattr(qmc, name)(*args)
            out["%s_%s_next_rand" % (key, name)] = np.int64(libc.rand())
            out["%s_%s_out" % (key, name)] = c.astype(np.int8)
            assert np.array_equal(c, d), (key, name)
            kept.append("%s_%s" % (key, name))
    out["cases"] = np.array(kept)
    np.savez_compressed(os.path.join(HERE, "traj_qmc_wolff.npz"), **out)
    print("wolff fixtures written:", len(kept), "cases")


if __name__ == "__main__":
    import sys
    if "wolff" in sys.argv[1:]:
        wolff()
    else:
        santoro()
        trajectories()
        wolff()
