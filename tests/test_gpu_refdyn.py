"""Reference-dynamics production sweeps (dynamics="reference", csrc/mcs_refdyn.cu) through the C ABI.

The mode must reproduce the reference's update order IN DISTRIBUTION: per (replica, sweep, slice) a fresh random
visiting permutation, strictly sequential visits, slices in order (qmc.pyx:99-143; sa.pyx:73-99), world-line
moves in one more permutation (qmc.pyx:405-438).  Checked four ways:
  1. the dependency-wave execution is bit-identical to ONE thread walking the sites in increasing priority
     order (same kernel, MCS_REFDYN_SEQUENTIAL=1) -- the schedule is a sequential sweep in a random order;
  2. that random order is uniform: every relative order of neighbouring sites is equally likely (chi-square on
     the first-visited site of small cliques, reconstructed from a T = 0 quench);
  3. exact Boltzmann averages by full enumeration (detailed balance of the visit arithmetic);
  4. tier (c): TWO-SIDED |mean_gpu - mean_ref| <= 2 combined standard errors over 256 anneals against the
     reference's own residual energies on santoro_80x80 (tests/golden/santoro_ref_stats.json, all twelve
     cells, plus the P = 64 cells of santoro_ref_stats_p64.json).
"""
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc
from tests import instances as inst
from tests.test_gpu_production import _all_states, _classical_energies, _piqmc_exact

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def mcs():
    import montecarlosolvers_b200 as m
    m._lib.require_device()
    os.environ["MCS_REFDYN_CHECK"] = "1"  # surface the stalled-pass flag as an error
    return m


def _with_sequential(fn):
    os.environ["MCS_REFDYN_SEQUENTIAL"] = "1"
    try:
        return fn()
    finally:
        del os.environ["MCS_REFDYN_SEQUENTIAL"]


@pytest.mark.parametrize("case", ["torus12_P8", "torus12_fields_P5_global", "graph40_fields_P33", "torus6_P64_global",
                                  "torus8_P2"])
def test_piqmc_dependency_waves_equal_a_sequential_sweep_in_priority_order(mcs, case):
    if case == "torus12_P8":
        (_, nbs), P, glob = inst.torus(12, seed=1), 8, False
    elif case == "torus12_fields_P5_global":
        (_, nbs), P, glob = inst.torus(12, seed=2, fields=True), 5, True
    elif case == "graph40_fields_P33":
        (_, nbs), P, glob = inst.random_graph(40, 90, seed=3, fields=True), 33, False
    elif case == "torus6_P64_global":
        (_, nbs), P, glob = inst.torus(6, seed=4), 64, True
    else:
        (_, nbs), P, glob = inst.torus(8, seed=5), 2, False
    n, R, S = nbs.shape[0], 6, 5
    c0 = (2 * np.random.RandomState(7).randint(2, size=(R, n, P)) - 1).astype(np.int8)
    fn = mcs.qmc.QuantumAnnealGlobal if glob else mcs.qmc.QuantumAnneal
    A, B = np.linspace(2.0, 0.3, S), np.linspace(0.5, 1.0, S)

    def run():
        c = c0.copy()
        fn(A, B, 2, 0.7 / P, c, nbs, 1, seed=11, dynamics="reference")
        return c

    waves = run()
    seq = _with_sequential(run)
    assert not np.array_equal(waves, c0)
    assert np.array_equal(waves, seq)
    # a different seed gives a different trajectory; the coloured order gives a different one as well
    c = c0.copy()
    fn(A, B, 2, 0.7 / P, c, nbs, 1, seed=12, dynamics="reference")
    assert not np.array_equal(c, waves)


def test_sa_dependency_waves_equal_a_sequential_sweep_in_priority_order(mcs):
    for (_, nbs) in (inst.torus(10, seed=3, fields=True), inst.random_graph(50, 120, seed=4, fields=True)):
        n = nbs.shape[0]
        s0 = (2 * np.random.RandomState(1).randint(2, size=(37, n)) - 1).astype(np.int8)
        sched = np.linspace(2.5, 0.0, 9)

        def run():
            s = s0.copy()
            mcs.sa.Anneal(sched, 2, s, nbs, seed=21, dynamics="reference")
            return s

        waves = run()
        assert np.array_equal(waves, _with_sequential(run))
        assert not np.array_equal(waves, s0)


@pytest.mark.parametrize("tf", [0, 1])
def test_svmc_dependency_waves_equal_a_sequential_sweep_in_priority_order(mcs, tf):
    """Rotor sweeps (svmc.pyx:83-115) on the same wave machinery: bit-identical to one thread walking the sites in
    priority order; chunked rows (degree > 4), fields, time-dependent tables."""
    for nbs in (inst.torus(9, seed=3, fields=True)[1], inst.random_graph(60, 170, seed=4, fields=True)[1],
                np.repeat(inst.torus(6, seed=5, fields=True)[1][None], 7, axis=0) * np.linspace(0.5, 1.5, 7).reshape(
                    7, 1, 1, 1) ** np.array([0.0, 1.0])):
        noisy = nbs.ndim == 4
        n = nbs.shape[-3]
        R = 1 if noisy else 9
        v0 = np.random.RandomState(2).uniform(0, np.pi, size=(R, n))
        A, B = np.linspace(2.0, 0.2, 7), np.linspace(0.3, 1.0, 7)

        def run(seed=31):
            v = v0.copy()
            if noisy:
                (mcs.svmc.NoisySVMCTF if tf else mcs.svmc.NoisySVMC)(A, B, 2, 0.3, v[0], nbs, seed=seed,
                                                                       dynamics="reference")
            else:
                (mcs.svmc.SpinVectorMonteCarloTFCompact if tf else mcs.svmc.SpinVectorMonteCarloCompact)(
                    A, B, 2, 0.3, v, nbs, seed=seed, dynamics="reference")
            return v

        waves = run()
        assert np.array_equal(waves, _with_sequential(run))
        assert not np.array_equal(waves, v0) and not np.array_equal(waves, run(32))
        assert waves.min() >= 0.0 and waves.max() <= np.pi + 1e-6


def test_results_do_not_depend_on_sharding_or_call_splitting(mcs):
    """Philox counters carry the global replica and sweep numbers: a shard of the batch and a schedule split over
    two calls reproduce the one-call, one-batch result bit for bit."""
    _, nbs = inst.torus(8, seed=9, fields=True)
    I = mcs.Instance(nbs)
    I.set_dynamics("reference")
    P, R = 6, 48
    A, B = np.linspace(2.0, 0.2, 8), np.ones(8)
    st = mcs.State(I, mcs._lib.KIND_PIQMC, R, P)
    st.init_random(5)
    start = st.download_spins()
    st.piqmc_sweeps(A, B, 1, 1.0 / P, global_moves=True, seed=3)
    full = st.download_spins()
    st2 = mcs.State(I, mcs._lib.KIND_PIQMC, 16, P)
    st2.upload_spins(np.ascontiguousarray(start[32:48]))
    st2.piqmc_sweeps(A[:3], B[:3], 1, 1.0 / P, global_moves=True, seed=3, replica_offset=32)
    st2.piqmc_sweeps(A[3:], B[3:], 1, 1.0 / P, global_moves=True, seed=3, replica_offset=32, sweep_offset=3)
    assert np.array_equal(st2.download_spins(), full[32:48])
    # SA: same through the restart axis
    ss = mcs.State(I, mcs._lib.KIND_SA, 64, 1)
    ss.init_random(6)
    s_start = ss.download_spins()
    ss.sa_sweeps(np.linspace(2, 0.1, 6), 2, seed=4)
    s_full = ss.download_spins()
    s2 = mcs.State(I, mcs._lib.KIND_SA, 32, 1)
    s2.upload_spins(np.ascontiguousarray(s_start[32:]))
    s2.sa_sweeps(np.linspace(2, 0.1, 6)[:2], 2, seed=4, replica_offset=32)
    s2.sa_sweeps(np.linspace(2, 0.1, 6)[2:], 2, seed=4, replica_offset=32, sweep_offset=4)
    assert np.array_equal(s2.download_spins(), s_full[32:])


def test_visiting_order_is_a_uniform_random_permutation(mcs):
    """A T = 0 quench of an antiferromagnetic triangle from the all-up state flips exactly the FIRST visited spin
    (dE = -4J < 0; afterwards the two others sit at dE = 0... made positive by a tiny field), so the flipped
    spin reveals which site the permutation put first: each of the three must come first 1/3 of the time.
    Tolerance: chi-square (2 dof) < 18.4 (p = 1e-4) over 32768 restarts x 1 sweep."""
    import scipy.sparse as sps
    J = sps.dok_matrix((3, 3))
    J[0, 1] = J[1, 2] = J[0, 2] = 1.0  # antiferromagnetic in the reference's convention
    for i in range(3):
        J[i, i] = 0.6  # field: E = sum J s s + 0.6 sum s.  up,up,up: flipping one spin: dE = -2(2) - 1.2 < 0;
        # afterwards flipping a second one: dE = -2 s_i (J s_j + J s_k + h) = -2 (1 - 1 + 0.6) < 0 ... see below
    nbs = orc.GenerateNeighbors(3, J, 3)
    # With h = 0.6: state (+,+,+) E = 3 + 1.8; first visited flips (dE = -2*(2+0.6) = -5.2).  Second visited spin
    # sees one up, one down neighbour: dE = -2*(0 + 0.6) = -1.2 -> flips too; third sees two down: dE = -2*(-2+0.6)
    # = +2.8 -> stays.  So the spin that stays UP is the LAST visited: also uniform over the three sites.
    R = 32768
    s = np.ones((R, 3), dtype=np.int8)
    mcs.sa.Anneal(np.zeros(1), 1, s, nbs, seed=77, dynamics="reference")
    assert np.all((s == 1).sum(axis=1) == 1)
    last = np.argmax(s == 1, axis=1)
    counts = np.bincount(last, minlength=3).astype(float)
    chi2 = ((counts - R / 3) ** 2 / (R / 3)).sum()
    assert chi2 < 18.4, (counts, chi2)
    # and successive sweeps / replicas are independent: the last-visited site of sweep 2 is uniform given sweep 1
    s2 = np.ones((R, 3), dtype=np.int8)
    I = mcs.Instance(nbs)
    I.set_dynamics("reference")
    st = mcs.State(I, mcs._lib.KIND_SA, R, 1)
    st.upload_spins(s2)
    st.sa_sweeps(np.zeros(1), 1, seed=77, sweep_offset=1)
    last2 = np.argmax(st.download_spins() == 1, axis=1)
    table = np.zeros((3, 3))
    np.add.at(table, (last, last2), 1)
    chi2 = ((table - R / 9) ** 2 / (R / 9)).sum()
    assert chi2 < 33.7, (table, chi2)  # 8 dof (the margins are not fixed): p = 5e-5


@pytest.mark.parametrize("case", ["ring4_P4", "tri5_fields_P3", "torus_P2_global", "circulant7_P2_global"])
def test_piqmc_reference_dynamics_samples_the_exact_boltzmann_distribution(mcs, case):
    """Tolerance: |GPU mean - exact| <= 4.5 standard errors over 4096 replicas (as for the coloured kernels)."""
    if case == "ring4_P4":
        (_, nbs), P, glob = inst.random_graph(4, 4, seed=1, fields=False), 4, False
    elif case == "tri5_fields_P3":
        import scipy.sparse as sps
        J = sps.dok_matrix((5, 5))
        for (i, j, v) in ((0, 1, 0.9), (1, 2, -0.7), (0, 2, 0.5), (2, 3, 1.1), (3, 4, -0.6), (4, 0, 0.8)):
            J[i, j] = v
        J[1, 1] = 0.4
        J[3, 3] = -0.3
        nbs, P, glob = orc.GenerateNeighbors(5, J, 4), 3, False
    elif case == "circulant7_P2_global":
        (_, nbs), P, glob = inst.circulant(7, (1, 2, 3), seed=3, fields=False), 2, True
    else:
        (_, nbs), P, glob = inst.torus(2, seed=2, fields=True), 2, True
    a, b, temp = 1.1, 0.8, 0.9 / P
    n = nbs.shape[0]
    e_exact, l_exact = _piqmc_exact(nbs, P, a, b, temp)
    I = mcs.Instance(nbs)
    I.set_dynamics("reference")
    R = 4096
    st = mcs.State(I, mcs._lib.KIND_PIQMC, R, P)
    st.init_random(5)
    st.piqmc_sweeps(np.full(150, a), np.full(150, b), 1, temp, global_moves=glob, seed=5)
    es, ls = [], []
    for t in range(40):
        st.piqmc_sweeps(np.full(3, a), np.full(3, b), 1, temp, global_moves=glob, seed=5, sweep_offset=150 + 3 * t)
        c = st.download_spins().astype(np.float64)
        es.append(st.energies().mean(axis=1))
        ls.append((c * np.roll(c, -1, axis=2)).sum(axis=(1, 2)) / (n * P))
    es, ls = np.array(es).mean(axis=0), np.array(ls).mean(axis=0)
    assert abs(es.mean() - e_exact) <= 4.5 * es.std(ddof=1) / np.sqrt(R), (case, es.mean(), e_exact)
    assert abs(ls.mean() - l_exact) <= 4.5 * ls.std(ddof=1) / np.sqrt(R), (case, ls.mean(), l_exact)


def test_sa_reference_dynamics_samples_the_exact_boltzmann_distribution(mcs):
    _, nbs = inst.random_graph(10, 16, seed=4, fields=True)
    T = 1.3
    e_all = _classical_energies(_all_states(10), nbs)
    w = np.exp(-(e_all - e_all.min()) / T)
    w /= w.sum()
    e_exact = float(np.dot(w, e_all))
    R = 4096
    I = mcs.Instance(nbs)
    I.set_dynamics("reference")
    st = mcs.State(I, mcs._lib.KIND_SA, R, 1)
    st.init_random(9)
    st.sa_sweeps(np.full(100, T), 1, seed=9)
    es = []
    for t in range(40):
        st.sa_sweeps(np.full(3, T), 1, seed=9, sweep_offset=100 + 3 * t)
        es.append(st.energies())
    es = np.array(es).mean(axis=0)
    assert abs(es.mean() - e_exact) <= 4.5 * es.std(ddof=1) / np.sqrt(R), (es.mean(), e_exact)


# ---- tier (c): two-sided at the standard-error level against the REFERENCE's own statistics -----------------------
def _ref_stats(name="santoro_ref_stats.json"):
    with open(os.path.join(G, name)) as f:
        return json.load(f)


def _two_sided(name, got, ref_cell, log):
    ref = np.asarray(ref_cell)
    sem = np.sqrt(got.var(ddof=1) / got.size + ref.var(ddof=1) / ref.size)
    msg = "%s: gpu(reference dynamics) %.5f (sd %.5f)  reference %.5f (sd %.5f)  diff %+.2f sem = %+.2f%%" % (
        name, got.mean(), got.std(ddof=1), ref.mean(), ref.std(ddof=1), (got.mean() - ref.mean()) / sem,
        100 * (got.mean() / ref.mean() - 1))
    print(msg)
    log.append(msg)
    assert got.size >= 256 and ref.size >= 256
    assert abs(got.mean() - ref.mean()) <= 2.0 * sem, msg
    assert 0.75 <= got.std(ddof=1) / ref.std(ddof=1) <= 1.33, msg


_LOG = []


@pytest.mark.parametrize("tau", [60, 146, 354, 857])
def test_santoro_sa_residual_energy_two_sided_vs_reference(mcs, tau):
    """CA protocol of santoro80.py:258-262, 256 anneals from the reference run's initial states."""
    _, nbs, _, e_gs = inst.santoro()
    s = np.stack([inst.random_spins(6400, r) for r in range(256)]).astype(np.int8)
    e = mcs.sa.Anneal(np.linspace(3.0, 0.0, tau), 1, s, nbs, seed=4321 + tau, energies=True, dynamics="reference")
    _two_sided("sa_tau%d" % tau, (e - e_gs) / 6400, _ref_stats()["cells"]["sa_tau%d" % tau], _LOG)


@pytest.mark.parametrize("glob", [1, 0])
@pytest.mark.parametrize("tau", [60, 146, 354, 857])
def test_santoro_piqmc_residual_energy_two_sided_vs_reference(mcs, tau, glob):
    """PIQMC protocol of santoro80.py:279-298 (P = 20, PT = 1), from the reference run's 256 pre-annealed states."""
    _, nbs, _, e_gs = inst.santoro()
    P, R = 20, 256
    pre = np.load(os.path.join(G, "santoro_preannealed.npz"))
    s = np.where(np.unpackbits(pre["packed"], axis=1)[:, :6400] > 0, 1, -1).astype(np.int8)[:R]
    confs = np.ascontiguousarray(np.repeat(s[:, :, None], P, axis=2))
    fn = mcs.qmc.QuantumAnnealGlobal if glob else mcs.qmc.QuantumAnneal
    e = fn(np.linspace(3.0, 1e-8, tau), np.ones(tau), 1, 1.0 / P, confs, nbs, 1, seed=199 + tau, energies=True,
           dynamics="reference")
    name = "qmc%s_P20_tau%d" % ("_global" if glob else "", tau)
    _two_sided(name, (e.min(axis=1) - e_gs) / 6400, _ref_stats()["cells"][name], _LOG)


@pytest.mark.parametrize("name", ["qmc_P64_tau60", "qmc_P64_tau146", "qmc_global_P64_tau60", "qmc_global_P64_tau146"])
def test_santoro_piqmc_P64_residual_energy_two_sided_vs_reference(mcs, name):
    """The BASELINE cfg3 shape (P = 64) against tests/golden/santoro_ref_stats_p64.json."""
    ref = _ref_stats("santoro_ref_stats_p64.json")
    if name not in ref["cells"]:
        pytest.skip("cell not in the fixture")
    _, nbs, _, e_gs = inst.santoro()
    P, R = 64, 256
    tau, glob = int(name.split("tau")[1]), "global" in name
    pre = np.load(os.path.join(G, "santoro_preannealed.npz"))
    s = np.where(np.unpackbits(pre["packed"], axis=1)[:, :6400] > 0, 1, -1).astype(np.int8)[:R]
    confs = np.ascontiguousarray(np.repeat(s[:, :, None], P, axis=2))
    fn = mcs.qmc.QuantumAnnealGlobal if glob else mcs.qmc.QuantumAnneal
    e = fn(np.linspace(3.0, 1e-8, tau), np.ones(tau), 1, 1.0 / P, confs, nbs, 1, seed=299 + tau, energies=True,
           dynamics="reference")
    _two_sided(name, (e.min(axis=1) - e_gs) / 6400, ref["cells"][name], _LOG)


def test_zz_write_tier_c_log():
    """Not a check: keeps the measured tier (c) table of this run under gpurun_out/ for profiles/."""
    if not _LOG:
        pytest.skip("no tier (c) cells ran")
    out = os.path.join(os.path.dirname(G), "..", "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "tier_c_refdyn.log"), "w") as f:
        f.write("\n".join(_LOG) + "\n")
