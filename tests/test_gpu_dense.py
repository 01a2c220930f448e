"""Dense (SK-like) instances, BASELINE configs[4]: the blocked tensor-core sweep (mcs_dense.cu) against
(i) the general coloured kernel (itself validated by exact enumeration) at equilibrium, (ii) the oracle,
(iii) a zero-temperature quench, where any error in the maintained local fields shows up as an
energy-raising flip or a non-minimal final state."""
import numpy as np
import pytest
import scipy.sparse as sps

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mcs():
    import os
    import montecarlosolvers_b200 as m
    m._lib.require_device()
    os.environ["MCS_DENSE_CHECK"] = "1"  # surface a timed-out wait between overlapping block kernels as an error
    return m


def sk_instance(n, seed=0, fields=False):
    rng = np.random.default_rng(seed)
    J = sps.dok_matrix((n, n))
    for i in range(n):
        for j in range(i + 1, n):
            J[i, j] = rng.normal() / np.sqrt(n)
        if fields and i % 3 == 0:
            J[i, i] = 0.3 * rng.normal()
    return J, orc.GenerateNeighbors(n, J, n - 1 + (1 if fields else 0))


def test_dense_sa_zero_temperature_quench_reaches_a_local_minimum(mcs):
    n, R = 200, 96  # n not a multiple of the 128-site block, R not a multiple of the 64-column tile
    _, nbs = sk_instance(n, seed=1, fields=True)
    I = mcs.Instance(nbs)
    assert I.dense and I.ncolors == n
    st = mcs.State(I, mcs._lib.KIND_SA, R, 1)
    st.init_random(3)
    e_prev = st.energies()
    for t in range(12):
        st.sa_sweeps(np.array([0.0]), 1, seed=5, sweep_offset=t)  # T = 0: only downhill moves (sa.pyx:96-99)
        e = st.energies()
        assert np.all(e <= e_prev + 1e-9), t
        e_prev = e
    s = st.download_spins()
    for r in (0, 31, 32, R - 1):
        de = orc.sa_delta_e(s[r].astype(np.int64), nbs)
        assert de.min() > -1e-4, (r, de.min())  # no single flip lowers the energy (fp32 fields vs fp64 check)
        assert e[r] == orc.ising_energy(s[r].astype(np.int64), nbs)


def test_dense_piqmc_matches_general_kernel_and_oracle_at_equilibrium(mcs):
    """<E_cl> and the Trotter link correlation at fixed (Gamma, T): dense path vs general coloured kernel
    (4.5 combined standard errors over 1024 replicas), and vs the oracle's sequential dynamics."""
    n, P, R = 64, 8, 1024
    _, nbs = sk_instance(n, seed=4)
    a, b, temp = 1.2, 1.0, 0.8 / P

    def run(dense, glob):
        I = mcs.Instance(nbs)
        assert I.dense
        I.use_dense(dense)
        st = mcs.State(I, mcs._lib.KIND_PIQMC, R, P)
        st.init_random(7)
        st.piqmc_sweeps(np.full(60, a), np.full(60, b), 1, temp, global_moves=glob, seed=11)
        es, ls = [], []
        for t in range(10):
            st.piqmc_sweeps(np.full(3, a), np.full(3, b), 1, temp, global_moves=glob, seed=11, sweep_offset=60 + 3 * t)
            c = st.download_spins().astype(np.float64)
            es.append(st.energies().mean(axis=1))
            ls.append((c * np.roll(c, -1, axis=2)).sum(axis=(1, 2)) / (n * P))
        return np.array(es).mean(axis=0), np.array(ls).mean(axis=0)

    for glob in (False, True):
        ed, ld = run(True, glob)
        eg, lg = run(False, glob)
        se = np.sqrt(ed.var(ddof=1) / R + eg.var(ddof=1) / R)
        sl = np.sqrt(ld.var(ddof=1) / R + lg.var(ddof=1) / R)
        assert abs(ed.mean() - eg.mean()) <= 4.5 * se, (glob, ed.mean(), eg.mean(), se)
        assert abs(ld.mean() - lg.mean()) <= 4.5 * sl, (glob, ld.mean(), lg.mean(), sl)
    # oracle sequential dynamics (qmc.pyx:93-143), fewer replicas
    Ro = 48
    eo, lo = [], []
    rng = orc.LibcRand(3)
    for r in range(Ro):
        c = np.tile((2 * np.random.RandomState(r).randint(2, size=n) - 1).astype(np.int64), (P, 1)).T.copy()
        orc.QuantumAnneal(np.full(60, a), np.full(60, b), 1, temp, c, nbs, 1, rng=rng)
        acc_e, acc_l = [], []
        for t in range(10):
            orc.QuantumAnneal(np.full(3, a), np.full(3, b), 1, temp, c, nbs, 1, rng=rng)
            acc_e.append(np.mean([orc.ising_energy(np.ascontiguousarray(c[:, q]), nbs) for q in range(P)]))
            acc_l.append((c * np.roll(c, -1, axis=1)).sum() / (n * P))
        eo.append(np.mean(acc_e))
        lo.append(np.mean(acc_l))
    eo, lo = np.array(eo), np.array(lo)
    ed, ld = run(True, False)
    assert abs(ed.mean() - eo.mean()) <= 4.5 * np.sqrt(ed.var(ddof=1) / R + eo.var(ddof=1) / Ro)
    assert abs(ld.mean() - lo.mean()) <= 4.5 * np.sqrt(ld.var(ddof=1) / R + lo.var(ddof=1) / Ro)


@pytest.mark.parametrize("P", [5, 10, 20, 40, 3, 64])
def test_dense_path_for_any_number_of_slices(mcs, P):
    """The example's Trotter numbers (santoro80.py:250: P = 5, 10, 20, 40) and odd P on a dense instance: a replica
    takes the next power of two (or 64) columns of the tensor-core tile, the extra columns are dead.  <E_cl> and the
    Trotter link correlation at fixed (Gamma, T) against the general coloured kernel, 4.5 combined standard errors
    over 768 replicas, with and without world-line moves."""
    n, R = 64, 768
    _, nbs = sk_instance(n, seed=6, fields=True)
    a, b, temp = 1.1, 1.0, 0.9 / P

    def run(dense, glob):
        I = mcs.Instance(nbs)
        assert I.dense
        I.use_dense(dense)
        st = mcs.State(I, mcs._lib.KIND_PIQMC, R, P)
        st.init_random(7)
        st.piqmc_sweeps(np.full(50, a), np.full(50, b), 1, temp, global_moves=glob, seed=13)
        es, ls = [], []
        for t in range(8):
            st.piqmc_sweeps(np.full(3, a), np.full(3, b), 1, temp, global_moves=glob, seed=13, sweep_offset=50 + 3 * t)
            c = st.download_spins().astype(np.float64)
            es.append(st.energies().mean(axis=1))
            ls.append((c * np.roll(c, -1, axis=2)).sum(axis=(1, 2)) / (n * P))
        return np.array(es).mean(axis=0), np.array(ls).mean(axis=0)

    for glob in (False, True):
        ed, ld = run(True, glob)
        eg, lg = run(False, glob)
        se = np.sqrt(ed.var(ddof=1) / R + eg.var(ddof=1) / R)
        sl = np.sqrt(ld.var(ddof=1) / R + lg.var(ddof=1) / R)
        assert abs(ed.mean() - eg.mean()) <= 4.5 * se, (P, glob, ed.mean(), eg.mean(), se)
        assert abs(ld.mean() - lg.mean()) <= 4.5 * sl, (P, glob, ld.mean(), lg.mean(), sl)


def test_dense_annealing_finds_low_energy_states_cfg5_shape(mcs):
    """cfg5 shape at reduced size: SK N = 256, P = 32, 64 replicas; annealing lowers the energy well below the
    random-state value and close to the SK ground-state density (about -0.76 N for large N)."""
    n, P, R = 256, 32, 64
    _, nbs = sk_instance(n, seed=9)
    c = np.repeat((2 * np.random.RandomState(0).randint(2, size=(R, n, 1)) - 1).astype(np.int8), P, axis=2)
    e = mcs.qmc.QuantumAnnealGlobal(np.linspace(3.0, 1e-8, 200), np.ones(200), 1, 1.0 / P, c, nbs, 1, seed=2,
                                    energies=True)
    best = e.min(axis=1) / n
    assert best.mean() < -0.65, best.mean()
    k = int(np.argmin(e[0]))
    assert abs(e[0, k] - orc.ising_energy(c[0, :, k].astype(np.int64), nbs)) < 1e-9
