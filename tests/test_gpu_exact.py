"""Parity tiers (a) and (b) on the GPU, through the C ABI: fp64 probes and the sequential-order
validation kernels must reproduce the oracle / the reference's golden trajectories BIT-EXACTLY."""
import os

import numpy as np
import pytest

from oracle import oracle as orc
from tests import instances as inst

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def mcs():
    import montecarlosolvers_b200 as m
    m._lib.require_device()
    return m


def test_probe_delta_e_bit_exact(mcs):
    for (_, nbs), P in ((inst.torus(6, seed=3, fields=True), 8), (inst.random_graph(40, 90, seed=2), 5),
                        (inst.torus(4, seed=1), 2)):
        n = nbs.shape[0]
        R = 3
        c = (2 * np.random.RandomState(7).randint(2, size=(R, n, P)) - 1).astype(np.int64)
        got = mcs.qmc.delta_e(0.8, 0.7, 1.0 / P, c, nbs)
        gotg = mcs.qmc.delta_e_global(0.7, c, nbs)
        for r in range(R):
            want = orc.qmc_delta_e(0.8, 0.7, 1.0 / P, c[r], nbs)
            assert np.array_equal(got[r], want)
            assert np.array_equal(gotg[r], orc.qmc_delta_e_global(0.7, c[r], nbs))
        s = np.ascontiguousarray(c[:, :, 0])
        gs = mcs.sa.delta_e(s, nbs)
        for r in range(R):
            assert np.array_equal(gs[r], orc.sa_delta_e(s[r], nbs))


def test_state_energies_bit_exact_and_roundtrip(mcs):
    _, nbs = inst.random_graph(50, 120, seed=5)
    n = nbs.shape[0]
    for P in (2, 7, 20, 33, 64):
        R = 37
        c = (2 * np.random.RandomState(P).randint(2, size=(R, n, P)) - 1).astype(np.int8)
        I = mcs.Instance(nbs)
        st = mcs.State(I, mcs._lib.KIND_PIQMC, R, P)
        st.upload_spins(c)
        assert np.array_equal(st.download_spins(), c)  # pack -> unpack round trip
        e = st.energies()
        for r in (0, 5, R - 1):
            for k in (0, P - 1):
                assert e[r, k] == orc.ising_energy(c[r, :, k].astype(np.int64), nbs)
        st.close()
    R = 70
    s = (2 * np.random.RandomState(1).randint(2, size=(R, n)) - 1).astype(np.int8)
    st = mcs.State(I, mcs._lib.KIND_SA, R, 1)
    st.upload_spins(s)
    assert np.array_equal(st.download_spins(), s)
    e = st.energies()
    for r in (0, 31, 32, R - 1):
        assert e[r] == orc.ising_energy(s[r].astype(np.int64), nbs)


@pytest.mark.parametrize("case", ["torus19", "torus18_fields", "circulant300_deg6"])
def test_tiled_energy_kernels_bit_exact(mcs, case):
    """The fixed-order energy kernels -- by table (rows with at most four off-diagonal entries: sixteen precomputed
    terms per site, one fp64 addition per site and accumulator) and the chain kernel -- perform the same additions in
    the same order as the oracle: several tiles, a ragged last tile and chunk, rows of 4 / 5 / 7 entries (the
    last: chain kernel only), fields, ragged replica counts."""
    if case == "torus19":
        nbs = inst.torus(19, seed=2)[1]
    elif case == "torus18_fields":
        nbs = inst.torus(18, seed=3, fields=True)[1]
    else:
        nbs = inst.circulant(300, (1, 2, 3), seed=4, fields=True)[1]
    n = nbs.shape[0]
    I = mcs.Instance(nbs)
    for P, R in ((64, 37), (20, 70), (3, 5)):
        c = (2 * np.random.RandomState(P).randint(2, size=(R, n, P)) - 1).astype(np.int8)
        st = mcs.State(I, mcs._lib.KIND_PIQMC, R, P)
        st.upload_spins(c)
        e = st.energies()
        for mode in ("1",):
            os.environ["MCS_ENERGY_CHAIN"] = mode
            try:
                assert np.array_equal(e, st.energies()), mode
            finally:
                os.environ.pop("MCS_ENERGY_CHAIN", None)
        for r in (0, R // 2, R - 1):
            for k in (0, P // 2, P - 1):
                assert e[r, k] == orc.ising_energy(c[r, :, k].astype(np.int64), nbs)
        st.close()
    R = 200
    s = (2 * np.random.RandomState(1).randint(2, size=(R, n)) - 1).astype(np.int8)
    st = mcs.State(I, mcs._lib.KIND_SA, R, 1)
    st.upload_spins(s)
    e = st.energies()
    for mode in ("1",):
        os.environ["MCS_ENERGY_CHAIN"] = mode
        try:
            assert np.array_equal(e, st.energies()), mode
        finally:
            os.environ.pop("MCS_ENERGY_CHAIN", None)
    for r in (0, 31, 32, 127, 128, R - 1):
        assert e[r] == orc.ising_energy(s[r].astype(np.int64), nbs)
    st.close()


def test_exact_qmc_golden(mcs):
    d = np.load(os.path.join(G, "traj_qmc_torus6.npz"))
    for P in (2, 3, 8, 20):
        for glob in (0, 1):
            c = np.asfortranarray(d["P%d_g%d_in" % (P, glob)].astype(np.int64))  # strided like the example
            fn = mcs.qmc.QuantumAnnealGlobal if glob else mcs.qmc.QuantumAnneal
            assert fn(d["A"], d["B"], int(d["mcsteps"]), 1.0 / P, c, d["nbs"], 1, exact=True,
                      libc_seed=1000 + P) is None
            assert np.array_equal(c, d["P%d_g%d_out" % (P, glob)])
    d = np.load(os.path.join(G, "traj_qmc_graph40.npz"))
    c = d["conf_in"].astype(np.int64)
    e = mcs.qmc.QuantumAnnealGlobal(d["A"], d["B"], int(d["mcsteps"]), float(d["temp"]), c, d["nbs"], 1, exact=True,
                                    libc_seed=int(d["seed"]), energies=True)
    assert np.array_equal(c, d["conf_out"])
    assert np.allclose(e, d["energies"], rtol=0, atol=1e-9)
    c = d["conf_in"].astype(np.int64)
    mcs.qmc.DissipativeQuantumAnnealGlobal(d["A"], d["B"], int(d["diss_mcsteps"]), float(d["temp"]), d["lut"], c,
                                           d["nbs"], 1, exact=True, libc_seed=int(d["diss_seed"]))
    assert np.array_equal(c, d["diss_out"])


def test_exact_qmc_batch_vs_oracle(mcs):
    """R replicas at once, replica r seeded srand(50 + r): each equals its own oracle run."""
    _, nbs = inst.torus(6, seed=11, fields=True)
    n, P, R = 36, 6, 9
    A = np.linspace(2.5, 0.05, 7)
    B = np.ones(7)
    c0 = (2 * np.random.RandomState(3).randint(2, size=(R, n, P)) - 1).astype(np.int8)
    c = c0.copy()
    mcs.qmc.QuantumAnnealGlobal(A, B, 2, 1.0 / P, c, nbs, 1, exact=True, libc_seed=50)
    for r in range(R):
        want = c0[r].astype(np.int64)
        orc.QuantumAnnealGlobal(A, B, 2, 1.0 / P, want, nbs, 1, rng=50 + r)
        assert np.array_equal(c[r], want)


def test_exact_sa_golden(mcs):
    d = np.load(os.path.join(G, "traj_sa_torus6.npz"))
    s = d["s_in"].astype(np.int64)
    mcs.sa.Anneal(d["sched"], int(d["mcsteps"]), s, d["nbs"], exact=True, libc_seed=int(d["seed"]))
    assert np.array_equal(s, d["s_out"])
    s = d["s_in"].astype(np.int64)
    np.random.seed(int(d["ma_seed"]))
    mcs.sa.AnnealMA(d["sched"], int(d["ma_mcsteps"]), s, d["nbs"], exact=True, libc_seed=int(d["ma_seed"]))
    assert np.array_equal(s, d["ma_out"])
    s = d["s_in"].astype(np.int64)
    mcs.sa.Anneal_parallel(d["sched"], int(d["mcsteps"]), s, d["nbs"], 6, exact=True, libc_seed=int(d["seed"]))
    assert np.array_equal(s, d["s_out"])


def test_exact_svmc_golden(mcs):
    d = np.load(os.path.join(G, "traj_svmc_torus5.npz"))
    seed = int(d["seed"])
    args = (d["A"], d["B"], int(d["mcsteps"]), float(d["temp"]))
    for name in ("SpinVectorMonteCarlo", "SpinVectorMonteCarloTF"):
        v = np.full(25, np.pi / 2)
        np.random.seed(seed)
        getattr(mcs.svmc, name)(*args, v, d["nbs"], exact=True, libc_seed=seed)
        assert np.array_equal(v, d[name]), name
    v = np.full((5, 25), np.pi / 2)
    np.random.seed(seed)
    mcs.svmc.SpinVectorMonteCarloCompact(*args, v, d["nbs"], exact=True, libc_seed=seed)
    assert np.array_equal(v, d["SpinVectorMonteCarloCompact"])
    v = np.full((4, 25), np.pi / 2)
    mcs.svmc.SpinVectorMonteCarloTFCompact(*args, v, d["nbs"], exact=True, libc_seed=seed)
    assert np.array_equal(v, d["SpinVectorMonteCarloTFCompact"])


def test_exact_santoro_slice_of_reference_protocol(mcs):
    """Full-size instance (80x80, P=20), a short QuantumAnnealGlobal segment: GPU replay == oracle."""
    _, nbs, _, _ = inst.santoro()
    P = 20
    A = np.linspace(3.0, 2.5, 2)
    s0 = inst.random_spins(6400, 0)
    want = np.tile(s0, (P, 1)).T.copy()
    got = want.copy()
    orc.QuantumAnnealGlobal(A, np.ones(2), 1, 1.0 / P, want, nbs, 1, rng=2000)
    mcs.qmc.QuantumAnnealGlobal(A, np.ones(2), 1, 1.0 / P, got, nbs, 1, exact=True, libc_seed=2000)
    assert np.array_equal(got, want)


def test_error_behaviour_matches_reference(mcs):
    _, nbs = inst.torus(4, seed=1)
    c = np.ones((16, 4), dtype=np.int64)
    with pytest.raises(ZeroDivisionError):  # qmc.c:3030-3034
        mcs.qmc.QuantumAnneal(np.ones(2), np.ones(2), 1, 0.0, c, nbs, 1)
    with pytest.raises(ValueError):
        mcs.qmc.QuantumAnneal(np.ones(2), np.ones(2), 1, 0.1, c.astype(np.float64), nbs, 1)
    with pytest.raises(ValueError):
        mcs.qmc.QuantumAnneal(np.ones(2), np.ones(2), 1, 0.1, c, nbs[0], 1)
    with pytest.raises(ValueError):  # P = 1 reads out of bounds in the reference; refused here
        mcs.qmc.QuantumAnneal(np.ones(2), np.ones(2), 1, 0.1, c[:, :1], nbs, 1)
    with pytest.raises(ValueError):
        mcs.sa.Anneal(np.ones(2, dtype=np.float32), 1, c[:, 0].copy(), nbs)


def test_exact_noisy_time_dependent_tables_vs_oracle(mcs):
    """sa.NoisyAnneal (sa.pyx:291-378) and svmc.NoisySVMC[TF] (svmc.pyx:236-448): nbs[step] per schedule step."""
    _, nbs = inst.random_graph(30, 60, seed=6)
    S = 10
    nbs4 = np.stack([nbs * np.array([1.0, 1.0 + 0.1 * t]) for t in range(S)])
    sched = np.linspace(2.0, 0.1, S)
    want = inst.random_spins(30, 3)
    got = want.copy()
    np.random.seed(4)
    orc.NoisyAnneal(sched, 2, want, nbs4, rng=4)
    np.random.seed(4)
    assert mcs.sa.NoisyAnneal(sched, 2, got, nbs4, exact=True, libc_seed=4) is None
    assert np.array_equal(got, want)
    _, nbs = inst.torus(4, seed=7, fields=True)
    s = np.linspace(1e-2, 1.0, 8)
    A, B = 3.0 * (1 - s), s
    nbs4 = np.stack([nbs * np.array([1.0, 1.0 + 0.05 * t]) for t in range(8)])
    for name in ("NoisySVMC", "NoisySVMCTF"):
        want = np.full(16, np.pi / 2)
        got = want.copy()
        np.random.seed(2)
        getattr(orc, name)(A, B, 2, 0.1, want, nbs4, rng=2)
        np.random.seed(2)
        getattr(mcs.svmc, name)(A, B, 2, 0.1, got, nbs4, exact=True, libc_seed=2)
        assert np.array_equal(got, want), name


def test_exact_replay_fed_the_references_own_random_numbers(mcs):
    """North-star tier (b) verbatim: the sequential kernel is FED recorded rand() outputs (here the first
    values of the stream after srand(77), which is what the compiled reference drew when the golden fixture
    was made) and reproduces the reference trajectory, consuming exactly as many values as the reference."""
    d = np.load(os.path.join(G, "traj_qmc_torus6.npz"))
    P = 8
    for glob in (0, 1):
        rng = orc.LibcRand(1000 + P)
        stream = rng.draw(200000)
        c = d["P%d_g%d_in" % (P, glob)].astype(np.int64).copy()
        fn = mcs.qmc.QuantumAnnealGlobal if glob else mcs.qmc.QuantumAnneal
        fn(d["A"], d["B"], int(d["mcsteps"]), 1.0 / P, c, d["nbs"], 1, rand_stream=stream)
        assert np.array_equal(c, d["P%d_g%d_out" % (P, glob)])
        used = int(mcs.qmc.last_rand_consumed()[0])
        # the next value of the stream is what the reference's rand() returned right after the call
        assert stream[used] == int(d["P%d_g%d_next_rand" % (P, glob)])
        with pytest.raises(ValueError):
            fn(d["A"], d["B"], int(d["mcsteps"]), 1.0 / P, c.copy(), d["nbs"], 1, rand_stream=stream[:100])


def test_exact_wolff_experiments_golden(mcs):
    """qmc.pyx:612-1621, the reference's five Wolff-cluster functions: the replay kernel against fixtures made from
    the compiled reference (tests/golden/make_golden.py wolff) -- configurations and rand() draws consumed."""
    from tests.test_oracle_golden import wolff_golden_cases
    n = 0
    for key, name, A, B, mcsteps, temp, lut, nbs, seed, cin, cout, nxt in wolff_golden_cases():
        c = cin.copy()
        args = (A, B, mcsteps, temp) + ((lut,) if lut is not None else ()) + (c, nbs)
        args += (1,) if name.endswith(("WC2", "WC3")) else ()
        assert getattr(mcs.qmc, name)(*args, exact=True, libc_seed=seed) is None
        assert np.array_equal(c, cout), (key, name)
        assert orc.LibcRand(seed).draw(int(mcs.qmc.last_rand_consumed()[0]) + 1)[-1] == nxt, (key, name)
        assert not mcs.qmc.last_wolff_overrun()[0]
        n += 1
    assert n >= 15


@pytest.mark.parametrize("name", ["QuantumAnnealWCL", "DissaptiveQuantumAnnealWCL", "QuantumAnnealWC",
                                  "DissipativeQuantumAnnealWC2", "DissipativeQuantumAnnealWC3"])
def test_exact_wolff_experiments_batch_vs_oracle(mcs, name):
    """A batch of replicas (replica r <-> srand(seed + r)) against the oracle, including replicas in which the
    reference itself would overrun its `cluster` buffer (flag equal to the oracle's)."""
    _, nbs = inst.random_graph(24, 50, seed=8, fields=True)
    n, P, R = nbs.shape[0], 4, 24
    A, B = np.linspace(1.5, 0.2, 5), np.linspace(0.5, 1.0, 5)
    lut = 0.6 * (np.pi / (P * np.sin(np.pi * np.arange(1, P) / P))) ** 2
    diss = "iss" in name
    c0 = (2 * np.random.RandomState(4).randint(2, size=(R, n, P)) - 1).astype(np.int8)
    c0[::3] = c0[::3, :, :1]  # some replicas with aligned world lines (full bath clusters)
    c = c0.copy()
    args = (A, B, 5, 1.2 / P) + ((lut,) if diss else ())
    getattr(mcs.qmc, name)(*(args + (c, nbs)), exact=True, libc_seed=300)
    used, over = mcs.qmc.last_rand_consumed(), mcs.qmc.last_wolff_overrun()
    for r in range(R):
        d = c0[r].astype(np.int64)
        rng = orc.LibcRand(300 + r)
        getattr(orc, name)(*(args + (d, nbs)), rng=rng)
        assert np.array_equal(c[r], d.astype(np.int8)), r
        assert bool(over[r]) == bool(orc.last_wolff_overrun), r
        assert orc.LibcRand(300 + r).draw(int(used[r]) + 1)[-1] == rng.draw(1)[0], r
    assert not np.array_equal(c, c0)
