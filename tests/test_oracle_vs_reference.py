"""Pin the CPU oracle (oracle/mcs_oracle.c) against the reference itself, compiled here by
oracle/build_ref.py into oracle/_ref.  Bit-exact trajectories given the same libc
rand() / np.random state.  Skipped where oracle/_ref is absent (it travels to the GPU box prebuilt)."""
import ctypes

import numpy as np
import pytest

from oracle import build_ref
from oracle import oracle as orc
from tests import instances as inst

build_ref.build(verbose=False)
ref = build_ref.import_ref()
pytestmark = pytest.mark.skipif(ref is None, reason="compiled reference (oracle/_ref) not available")

libc = ctypes.CDLL(None)


def _ref(mod):
    import importlib
    return importlib.import_module("solvers." + mod)


def test_rand_clone_matches_libc():
    for seed in (1, 42, 123456789, 0):
        libc.srand(seed)
        want = np.array([libc.rand() for _ in range(2000)], dtype=np.int32)
        got = orc.LibcRand(seed).draw(2000)
        assert np.array_equal(want, got)


@pytest.mark.parametrize("P", [2, 3, 8])
@pytest.mark.parametrize("glob", [0, 1])
def test_qmc_bit_exact(P, glob):
    _, nbs = inst.torus(6, seed=3, fields=True)
    n = nbs.shape[0]
    A = np.linspace(3.0, 1e-8, 12)
    B = np.linspace(0.2, 1.0, 12)
    s0 = inst.random_spins(n, 5)
    # Fortran-strided [N,P] view like the example passes (santoro80.py:286)
    c_ref = np.tile(s0, (P, 1)).T.copy(order="F")
    c_orc = c_ref.copy(order="F")
    fn_ref = getattr(_ref("qmc"), "QuantumAnnealGlobal" if glob else "QuantumAnneal")
    fn_orc = orc.QuantumAnnealGlobal if glob else orc.QuantumAnneal
    libc.srand(77)
    fn_ref(A, B, 2, 1.0 / P, c_ref, nbs, 1)
    nxt = libc.rand()
    rng = orc.LibcRand(77)
    fn_orc(A, B, 2, 1.0 / P, c_orc, nbs, 1, rng=rng)
    assert np.array_equal(c_ref, c_orc)
    assert rng.draw(1)[0] == nxt  # same number of rand() draws consumed
    assert not np.array_equal(c_ref, np.tile(s0, (P, 1)).T)


def test_qmc_irregular_graph_c_order():
    _, nbs = inst.random_graph(40, 90, seed=2)
    n = nbs.shape[0]
    P = 5
    A = np.linspace(2.0, 0.01, 9)
    B = np.ones(9)
    c_ref = (2 * np.random.RandomState(1).randint(2, size=(n, P)) - 1).astype(np.int64)
    c_orc = c_ref.copy()
    libc.srand(5)
    _ref("qmc").QuantumAnnealGlobal(A, B, 3, 0.07, c_ref, nbs, 1)
    orc.QuantumAnnealGlobal(A, B, 3, 0.07, c_orc, nbs, 1, rng=5)
    assert np.array_equal(c_ref, c_orc)


@pytest.mark.parametrize("name", ["QuantumAnnealWCL", "DissaptiveQuantumAnnealWCL", "QuantumAnnealWC",
                                  "DissipativeQuantumAnnealWC2", "DissipativeQuantumAnnealWC3"])
def test_wolff_experiments_bit_exact(name):
    """qmc.pyx:612-1621 (runnable through build_ref.py's dtype patch): configurations and the number of rand() draws.
    The oracle runs first: a case in which the reference would write past its unchecked `cluster` buffer is
    undefined behaviour there and is not compared."""
    fn_ref, fn_orc = getattr(_ref("qmc"), name), getattr(orc, name)
    diss, compared = "iss" in name, 0
    for case, (nbs, P) in enumerate([(inst.torus(6, seed=1, fields=True)[1], 8), (inst.torus(5, seed=2)[1], 5),
                                     (inst.random_graph(30, 70, seed=3, fields=True)[1], 2),
                                     (inst.random_graph(20, 40, seed=4, fields=True)[1], 16)]):
        n = nbs.shape[0]
        for q, seed in enumerate((5, 6, 7)):
            c0 = (2 * np.random.RandomState(seed).randint(2, size=(n, P)) - 1).astype(np.int64)
            if seed == 6:  # all slices equal, Fortran order: what the example passes (santoro80.py:286)
                c0 = np.tile(inst.random_spins(n, 3), (P, 1)).T.copy(order="F")
            A, B = np.linspace(2, 0.1, 6), np.linspace(0.4, 1.0, 6)
            temp = [0.5, 2.0, 0.1][q] / P
            lut = (np.pi / (P * np.sin(np.pi * np.arange(1, P) / P))) ** 2 * [0.1, 0.5, 0.02][q]
            c_orc = c0.copy(order="K")
            rng = orc.LibcRand(seed)
            fn_orc(*((A, B, 7, temp) + ((lut,) if diss else ()) + (c_orc, nbs)), rng=rng)
            if orc.last_wolff_overrun:
                continue
            c_ref = c0.copy(order="K")
            libc.srand(seed)
            args = (A, B, 7, temp) + ((lut,) if diss else ()) + (c_ref, nbs)
            fn_ref(*(args + ((1,) if name.endswith(("WC2", "WC3")) else ())))
            assert np.array_equal(c_ref, c_orc), (case, seed)
            assert int(rng.draw(1)[0]) == libc.rand(), (case, seed)
            assert not np.array_equal(c_ref, c0)
            compared += 1
    assert compared >= 8


@pytest.mark.parametrize("glob", [0, 1])
def test_qmc_dissipative_bit_exact(glob):
    _, nbs = inst.torus(5, seed=9, fields=True)
    n = nbs.shape[0]
    P = 6
    k = np.arange(1, P)
    lut = 0.05 * (np.pi / (P * np.sin(np.pi * k / P))) ** 2
    A = np.linspace(3.0, 1e-3, 8)
    B = np.ones(8)
    c_ref = np.tile(inst.random_spins(n, 2), (P, 1)).T.copy()
    c_orc = c_ref.copy()
    name = "DissipativeQuantumAnnealGlobal" if glob else "DissipativeQuantumAnneal"
    libc.srand(11)
    getattr(_ref("qmc"), name)(A, B, 2, 1.0 / P, lut, c_ref, nbs, 1)
    getattr(orc, name)(A, B, 2, 1.0 / P, lut, c_orc, nbs, 1, rng=11)
    assert np.array_equal(c_ref, c_orc)


def test_qmc_zero_teff_raises_like_reference():
    _, nbs = inst.torus(4, seed=1)
    c = np.ones((16, 4), dtype=np.int64)
    with pytest.raises(ZeroDivisionError):
        _ref("qmc").QuantumAnneal(np.ones(2), np.ones(2), 1, 0.0, c.copy(), nbs, 1)
    with pytest.raises(ZeroDivisionError):
        orc.QuantumAnneal(np.ones(2), np.ones(2), 1, 0.0, c.copy(), nbs, 1)


def test_sa_bit_exact_including_zero_temperature():
    _, nbs = inst.torus(6, seed=4, fields=True)
    n = nbs.shape[0]
    sched = np.linspace(3.0, 0.0, 30)  # ends at T = 0 like santoro80.py:260
    s_ref = inst.random_spins(n, 8)
    s_orc = s_ref.copy()
    libc.srand(99)
    _ref("sa").Anneal(sched, 3, s_ref, nbs)
    nxt = libc.rand()
    rng = orc.LibcRand(99)
    orc.Anneal(sched, 3, s_orc, nbs, rng=rng)
    assert np.array_equal(s_ref, s_orc)
    assert rng.draw(1)[0] == nxt
    s_par = inst.random_spins(n, 8)
    libc.srand(99)
    _ref("sa").Anneal_parallel(sched, 3, s_par, nbs, 1)
    assert np.array_equal(s_par, s_orc)


def test_sa_ma_and_noisy_bit_exact():
    _, nbs = inst.random_graph(30, 60, seed=6)
    n = nbs.shape[0]
    sched = np.linspace(2.0, 0.1, 10)
    s_ref = inst.random_spins(n, 3)
    s_orc = s_ref.copy()
    libc.srand(4)
    np.random.seed(4)
    _ref("sa").AnnealMA(sched, 2, s_ref, nbs)
    np.random.seed(4)
    orc.AnnealMA(sched, 2, s_orc, nbs, rng=4)
    assert np.array_equal(s_ref, s_orc)
    nbs4 = np.stack([nbs * np.array([1.0, 1.0 + 0.1 * t]) for t in range(10)])
    s_ref = inst.random_spins(n, 3)
    s_orc = s_ref.copy()
    libc.srand(4)
    np.random.seed(4)
    _ref("sa").NoisyAnneal(sched, 2, s_ref, nbs4)
    np.random.seed(4)
    orc.NoisyAnneal(sched, 2, s_orc, nbs4, rng=4)
    assert np.array_equal(s_ref, s_orc)


@pytest.mark.parametrize("name", ["SpinVectorMonteCarlo", "SpinVectorMonteCarloTF"])
def test_svmc_bit_exact(name):
    _, nbs = inst.torus(5, seed=7, fields=True)
    n = nbs.shape[0]
    s = np.linspace(1e-3, 1.0, 20)
    A, B = 3.0 * (1 - s), s
    v_ref = np.full(n, np.pi / 2)
    v_orc = v_ref.copy()
    libc.srand(21)
    np.random.seed(21)
    getattr(_ref("svmc"), name)(A, B, 2, 0.1, v_ref, nbs)
    np.random.seed(21)
    getattr(orc, name)(A, B, 2, 0.1, v_orc, nbs, rng=21)
    assert np.array_equal(v_ref, v_orc)  # max |dtheta| == 0.0
    assert not np.allclose(v_ref, np.pi / 2)


@pytest.mark.parametrize("name", ["NoisySVMC", "NoisySVMCTF"])
def test_noisy_svmc_bit_exact(name):
    _, nbs = inst.torus(4, seed=7, fields=True)
    n = nbs.shape[0]
    s = np.linspace(1e-2, 1.0, 8)
    A, B = 3.0 * (1 - s), s
    nbs4 = np.stack([nbs * np.array([1.0, 1.0 + 0.05 * t]) for t in range(8)])
    v_ref = np.full(n, np.pi / 2)
    v_orc = v_ref.copy()
    libc.srand(2)
    np.random.seed(2)
    getattr(_ref("svmc"), name)(A, B, 2, 0.1, v_ref, nbs4)
    np.random.seed(2)
    getattr(orc, name)(A, B, 2, 0.1, v_orc, nbs4, rng=2)
    assert np.array_equal(v_ref, v_orc)


def test_svmc_compact_bit_exact():
    _, nbs = inst.random_graph(24, 50, seed=1)
    n = nbs.shape[0]
    s = np.linspace(1e-2, 1.0, 10)
    A, B = 3.0 * (1 - s), s
    v_ref = np.full((5, n), np.pi / 2)
    v_orc = v_ref.copy()
    libc.srand(8)
    np.random.seed(8)
    _ref("svmc").SpinVectorMonteCarloCompact(A, B, 2, 0.2, v_ref, nbs)
    np.random.seed(8)
    orc.SpinVectorMonteCarloCompact(A, B, 2, 0.2, v_orc, nbs, rng=8)
    assert np.array_equal(v_ref, v_orc)
    v_ref = np.full((4, n), np.pi / 2)
    v_orc = v_ref.copy()
    libc.srand(8)
    _ref("svmc").SpinVectorMonteCarloTFCompact(A, B, 2, 0.2, v_ref, nbs)
    nxt = libc.rand()
    rng = orc.LibcRand(8)
    orc.SpinVectorMonteCarloTFCompact(A, B, 2, 0.2, v_orc, nbs, rng=rng)
    assert np.array_equal(v_ref, v_orc)
    assert rng.draw(1)[0] == nxt


def test_generate_neighbors_and_energy_match_tools():
    J, nbs = inst.random_graph(25, 45, seed=12)
    want = np.asarray(_ref("tools").GenerateNeighbors(25, J, nbs.shape[1]))
    assert np.array_equal(want, nbs)
    for seed in range(4):
        s = inst.random_spins(25, seed)
        e_ref = _ref("tools").ClassicalIsingEnergy(s, J)
        e_orc = orc.ising_energy(s, nbs)
        assert abs(e_ref - e_orc) <= 1e-12 * max(1.0, abs(e_ref)) * 45
