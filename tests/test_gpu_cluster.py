"""Swendsen-Wang cluster moves (GPU union-find).  The reference advertises Swendsen-Wang (README.md:4) but has no
such code (its experimental single-cluster Wolff functions are replayed bit-exactly, tests/test_gpu_exact.py), so
correctness of THIS move = it leaves the exact Boltzmann distribution invariant AND is ergodic on its own:
cluster-only dynamics must reproduce full-enumeration averages -- also with the Ohmic-bath bonds of the Dissipative
solvers (qmc.pyx:268-273)."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests import instances as inst
from tests.test_gpu_production import _piqmc_exact, _all_states, _classical_energies

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mcs():
    import montecarlosolvers_b200 as m
    m._lib.require_device()
    return m


@pytest.mark.parametrize("case", ["ferro_ring_P4", "glass_fields_P3", "torus_P2"])
@pytest.mark.parametrize("mix", ["cluster_only", "cluster_plus_local"])
def test_sw_moves_sample_the_exact_distribution(mcs, case, mix):
    """4096 replicas, tolerance 4.5 standard errors on <E_cl> and the Trotter link correlation."""
    if case == "ferro_ring_P4":
        import scipy.sparse as sps
        J = sps.dok_matrix((4, 4))
        for i in range(4):
            J[i, (i + 1) % 4] = -0.8  # J < 0: ferromagnetic in the reference's convention
        nbs = orc.GenerateNeighbors(4, J, 2)
        P = 4
    elif case == "glass_fields_P3":
        _, nbs = inst.random_graph(5, 7, seed=3, fields=True)
        P = 3
    else:
        _, nbs = inst.torus(2, seed=2, fields=True)
        P = 2
    a, b, temp = 1.0, 0.9, 1.0 / P
    e_exact, l_exact = _piqmc_exact(nbs, P, a, b, temp)
    n = nbs.shape[0]
    R = 4096
    I = mcs.Instance(nbs)
    st = mcs.State(I, mcs._lib.KIND_PIQMC, R, P)
    st.init_random(3)

    def step(t):
        if mix == "cluster_plus_local":
            st.piqmc_sweeps(np.array([a]), np.array([b]), 1, temp, seed=5, sweep_offset=t)
        st.cluster_moves(a, b, temp, 1, seed=5, sweep_offset=t)

    for t in range(200):
        step(t)
    es, ls = [], []
    for t in range(200, 320):
        step(t)
        if t % 3 == 0:
            c = st.download_spins().astype(np.float64)
            es.append(st.energies().mean(axis=1))
            ls.append((c * np.roll(c, -1, axis=2)).sum(axis=(1, 2)) / (n * P))
    es, ls = np.array(es).mean(axis=0), np.array(ls).mean(axis=0)
    assert abs(es.mean() - e_exact) <= 4.5 * es.std(ddof=1) / np.sqrt(R), (es.mean(), e_exact)
    assert abs(ls.mean() - l_exact) <= 4.5 * ls.std(ddof=1) / np.sqrt(R), (ls.mean(), l_exact)


@pytest.mark.parametrize("case", ["ring_P4_bath", "glass_fields_P5_bath", "torus_P2_bath", "pair_P6_antibath"])
@pytest.mark.parametrize("mix", ["cluster_only", "cluster_plus_local"])
def test_sw_moves_with_bath_bonds_sample_the_exact_distribution(mcs, case, mix):
    """The action of DissipativeQuantumAnneal / DissaptiveQuantumAnnealWCL / WC2 / WC3 (qmc.pyx:268-273): bonds
    K_d = lookuptable[d-1] between all pairs of slices of a world line.  4096 replicas, 4.5 standard errors on <E_cl>
    and the Trotter link correlation against full enumeration (even and odd P, P = 2, a negative table)."""
    import scipy.sparse as sps
    b = 0.9
    if case == "ring_P4_bath":
        J = sps.dok_matrix((3, 3))
        for i in range(3):
            J[i, (i + 1) % 3] = -0.7
        nbs, P, alpha = orc.GenerateNeighbors(3, J, 2), 4, 0.4
    elif case == "glass_fields_P5_bath":
        # (b = 0.3: with the full couplings this 3-spin glass has a metastable state that cluster-only dynamics
        # leaves once in ~700 moves -- slow mixing, not a wrong distribution)
        (_, nbs), P, alpha, b = inst.random_graph(3, 3, seed=5, fields=True), 5, 0.6, 0.3
    elif case == "torus_P2_bath":
        (_, nbs), P, alpha = inst.torus(2, seed=2, fields=True), 2, 0.3
    else:
        J = sps.dok_matrix((2, 2))
        J[0, 1] = 0.9
        J[0, 0] = -0.2
        nbs, P, alpha = orc.GenerateNeighbors(2, J, 2), 6, -0.25
    lut = alpha * (np.pi / (P * np.sin(np.pi * np.arange(1, P) / P))) ** 2
    a, temp = 1.0, 1.0 / P
    e_exact, l_exact = _piqmc_exact(nbs, P, a, b, temp, lut=lut)
    n = nbs.shape[0]
    R = 4096
    I = mcs.Instance(nbs)
    st = mcs.State(I, mcs._lib.KIND_PIQMC, R, P)
    st.init_random(3)

    def step(t):
        if mix == "cluster_plus_local":
            st.piqmc_sweeps_dissipative(np.array([a]), np.array([b]), 1, temp, lut, seed=5, sweep_offset=t)
        st.cluster_moves(a, b, temp, 1, seed=5, sweep_offset=t, lookuptable=lut)

    for t in range(200):
        step(t)
    es, ls = [], []
    for t in range(200, 320):
        step(t)
        if t % 3 == 0:
            c = st.download_spins().astype(np.float64)
            es.append(st.energies().mean(axis=1))
            ls.append((c * np.roll(c, -1, axis=2)).sum(axis=(1, 2)) / (n * P))
    es, ls = np.array(es).mean(axis=0), np.array(ls).mean(axis=0)
    assert abs(es.mean() - e_exact) <= 4.5 * es.std(ddof=1) / np.sqrt(R), (es.mean(), e_exact)
    assert abs(ls.mean() - l_exact) <= 4.5 * ls.std(ddof=1) / np.sqrt(R), (ls.mean(), l_exact)
    # the bath must matter in this test: without it the exact averages are different
    e0, l0 = _piqmc_exact(nbs, P, a, b, temp)
    assert abs(l0 - l_exact) > 10 * ls.std(ddof=1) / np.sqrt(R)


def test_sw_bath_table_must_be_symmetric_and_dissipative_names_run(mcs):
    _, nbs = inst.torus(4, seed=2, fields=True)
    P = 6
    I = mcs.Instance(nbs)
    st = mcs.State(I, mcs._lib.KIND_PIQMC, 8, P)
    st.init_random(1)
    with pytest.raises(NotImplementedError):
        st.cluster_moves(1.0, 1.0, 1.0 / P, 1, lookuptable=np.linspace(0.1, 0.5, P - 1))
    lut = 0.2 * (np.pi / (P * np.sin(np.pi * np.arange(1, P) / P))) ** 2
    A, B = np.linspace(2.0, 0.05, 30), np.ones(30)
    conf = (2 * np.random.RandomState(0).randint(2, size=(16, 16, P)) - 1).astype(np.int8)
    c0 = conf.copy()
    for fn, extra in ((mcs.qmc.DissaptiveQuantumAnnealWCL, ()), (mcs.qmc.DissipativeQuantumAnnealWC2, (1,)),
                      (mcs.qmc.DissipativeQuantumAnnealWC3, (1,))):
        c = c0.copy()
        assert fn(A, B, 1, 0.3 / P, lut, c, nbs, *extra, seed=3) is None
        assert not np.array_equal(c, c0)
        e = np.array([[orc.ising_energy(c[r, :, k].astype(np.int64), nbs) for k in range(P)] for r in range(16)])
        e0 = np.array([[orc.ising_energy(c0[r, :, k].astype(np.int64), nbs) for k in range(P)] for r in range(16)])
        assert e.min(axis=1).mean() < e0.min(axis=1).mean() - 5.0


def test_sw_moves_sa_state_and_critical_ferromagnet(mcs):
    """Classical (P = 1) SW on an 8-site ferromagnetic ring with a field, against enumeration; and the
    drop-in entry point QuantumAnnealSW / QuantumAnnealWCL anneals a ferromagnet into its ground state."""
    import scipy.sparse as sps
    J = sps.dok_matrix((8, 8))
    for i in range(8):
        J[i, (i + 1) % 8] = -1.0
    J[0, 0] = 0.3
    nbs = orc.GenerateNeighbors(8, J, 3)
    T = 1.5
    sts = _all_states(8)
    e_all = _classical_energies(sts, nbs)
    w = np.exp(-(e_all - e_all.min()) / T)
    w /= w.sum()
    e_exact = float(np.dot(w, e_all))
    m_exact = float(np.dot(w, sts.mean(axis=1)))
    R = 4096
    I = mcs.Instance(nbs)
    st = mcs.State(I, mcs._lib.KIND_SA, R, 1)
    st.init_random(1)
    st.cluster_moves(0.0, 1.0, T, 100, seed=2)
    es, ms = [], []
    for t in range(40):
        st.cluster_moves(0.0, 1.0, T, 2, seed=2, sweep_offset=100 + 2 * t)
        es.append(st.energies())
        ms.append(st.download_spins().astype(np.float64).mean(axis=1))
    es, ms = np.array(es).mean(axis=0), np.array(ms).mean(axis=0)
    assert abs(es.mean() - e_exact) <= 4.5 * es.std(ddof=1) / np.sqrt(R), (es.mean(), e_exact)
    assert abs(ms.mean() - m_exact) <= 4.5 * ms.std(ddof=1) / np.sqrt(R), (ms.mean(), m_exact)
    # drop-in: 6x6 ferromagnetic torus, every replica ends in one of the two ground states
    Jt = sps.dok_matrix((36, 36))
    for r in range(6):
        for c in range(6):
            i = r * 6 + c
            Jt[i, r * 6 + (c + 1) % 6] = -1.0
            Jt[i, ((r + 1) % 6) * 6 + c] = -1.0
    nt = orc.GenerateNeighbors(36, Jt, 4)
    P = 8
    conf = (2 * np.random.RandomState(0).randint(2, size=(32, 36, P)) - 1).astype(np.int8)
    e = mcs.qmc.QuantumAnnealWCL(np.linspace(2.5, 1e-3, 60), np.ones(60), 1, 0.5 / P, conf, nt, seed=4, energies=True)
    assert np.all(e.min(axis=1) == -72.0)


def test_cluster_labels_are_built_chunk_by_chunk_without_changing_the_move(mcs):
    """The union-find forests take 4 (N P + 1) bytes per replica; batches are labelled in chunks that keep the array
    below 512 MB (80x80, P = 64: 320 replicas at a time instead of 6.7 GB for 4096).  Chunking must not change a
    single flip (Philox counters carry the global replica index): forced 32-replica chunks == one chunk."""
    import os
    _, nbs = inst.torus(8, seed=4, fields=True)
    P, R = 12, 200
    lut = 0.2 * (np.pi / (P * np.sin(np.pi * np.arange(1, P) / P))) ** 2
    I = mcs.Instance(nbs)
    out = []
    for chunk in (None, "32"):
        if chunk:
            os.environ["MCS_CLUSTER_CHUNK"] = chunk
        try:
            st = mcs.State(I, mcs._lib.KIND_PIQMC, R, P)
            st.init_random(2)
            for t in range(5):
                st.cluster_moves(1.0, 0.8, 1.0 / P, 1, seed=9, sweep_offset=t, lookuptable=lut if t % 2 else None)
            out.append(st.download_spins())
            st.close()
        finally:
            os.environ.pop("MCS_CLUSTER_CHUNK", None)
    assert np.array_equal(out[0], out[1])
