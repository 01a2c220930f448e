"""Small seeded Ising instances shared by the tests (built with the oracle's GenerateNeighbors restatement)."""
import numpy as np
import scipy.sparse as sps

from oracle import oracle as orc


def torus(L, seed=0, fields=False, lo=-2.0, hi=2.0):
    """L x L periodic square lattice, J ~ U(lo, hi), optional local fields on the diagonal."""
    rng = np.random.RandomState(seed)
    n = L * L
    J = sps.dok_matrix((n, n))
    for r in range(L):
        for c in range(L):
            i = r * L + c
            for j in (r * L + (c + 1) % L, ((r + 1) % L) * L + c):
                if i != j and (i, j) not in J and (j, i) not in J:
                    J[i, j] = rng.uniform(lo, hi)
    if fields:
        for i in range(n):
            J[i, i] = rng.uniform(-1.0, 1.0)
    maxnb = 4 + (1 if fields else 0)
    if L == 2:
        maxnb = 2 + (1 if fields else 0)
    return J, orc.GenerateNeighbors(n, J, maxnb)


def random_graph(n, nedges, seed=0, fields=True):
    """Irregular sparse graph (rows zero-padded to maxnb, tools.pyx:52-59)."""
    rng = np.random.RandomState(seed)
    J = sps.dok_matrix((n, n))
    deg = np.zeros(n, dtype=int)
    while J.nnz < nedges:
        i, j = rng.randint(n, size=2)
        if i == j or (i, j) in J or (j, i) in J:
            continue
        J[i, j] = rng.normal()
        deg[i] += 1
        deg[j] += 1
    if fields:
        for i in range(0, n, 2):
            J[i, i] = rng.normal()
            deg[i] += 1
    maxnb = int(deg.max())
    return J, orc.GenerateNeighbors(n, J, maxnb)


def circulant(n, offsets=(1, 2, 3), seed=0, fields=True):
    """Ring with couplings to i +- o for every o in offsets (degree 2 len(offsets)), optional fields: a
    high-degree sparse graph for the 7- and 8-plane threshold tables."""
    rng = np.random.RandomState(seed)
    J = sps.dok_matrix((n, n))
    for i in range(n):
        for o in offsets:
            j = (i + o) % n
            if (i, j) not in J and (j, i) not in J:
                J[i, j] = rng.uniform(-1.5, 1.5)
        if fields:
            J[i, i] = rng.uniform(-1.0, 1.0)
    return J, orc.GenerateNeighbors(n, J, 2 * len(offsets) + (1 if fields else 0))


def random_spins(n, seed):
    return (2 * np.random.RandomState(seed).randint(2, size=n) - 1).astype(np.int64)


_SANTORO = {}


def santoro():
    """The reference's shipped 80x80 instance (tests/golden/santoro80.npz, made by make_golden.py),
    with the example's sign flip isingJ[i,j] = -J_file (santoro80.py:242-244).
    Returns (J dok, nbs[6400,4,2], ground_state int64[6400], E_gs)."""
    if not _SANTORO:
        import os
        d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "santoro80.npz"))
        J = sps.dok_matrix((6400, 6400))
        for i, j, v in zip(d["i"], d["j"], d["J_file"]):
            J[int(i), int(j)] = -1.0 * v
        nbs = orc.GenerateNeighbors(6400, J, 4)
        _SANTORO["v"] = (J, nbs, d["ground_state"].astype(np.int64), float(d["e_gs_per_spin"]) * 6400)
    return _SANTORO["v"]
