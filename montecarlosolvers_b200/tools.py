"""Drop-in for the reference's `solvers.tools` (instance format and energy evaluation).

The table format is the hot path's input (reference tools.pyx:28-96) and the energy convention its
output check (tools.pyx:99-118).  These are host-side and O(nnz): the reference's GenerateNeighbors
is an O(nspins * nnz) Python double loop and its ClassicalIsingEnergy densifies J.
"""
import numpy as np
import scipy.sparse as sps

__all__ = ["bits2spins", "spins2bits", "GenerateNeighbors", "ClassicalIsingEnergy"]


def bits2spins(vec):
    """Convert a bitvector to a spinvector (tools.pyx:20-22)."""
    return [-1 if k == 1 else 1 for k in vec]


def spins2bits(vec):
    """Convert a spinvector to a bitvector (tools.pyx:24-26)."""
    return [0 if k == 1 else 1 for k in vec]


def GenerateNeighbors(nspins, J, maxnb, savepath=None):
    """GenerateNeighbors(nspins, J, maxnb, savepath=None)

    neighbours[i] = [[j, J_ij], ...] padded with [0, 0] to `maxnb` rows (a diagonal entry is a
    self entry = local field).  Row order is the DOK key order exactly as the reference produces
    it (tools.pyx:79-92), so fp64 sums over a row are bit-identical."""
    Jd = J.todok() if sps.issparse(J) else sps.dok_matrix(np.asarray(J))
    nbs = np.zeros((int(nspins), int(maxnb), 2))
    fill = np.zeros(int(nspins), dtype=np.int64)
    for (i, j), v in Jd.items():
        if fill[i] >= maxnb or (j != i and fill[j] >= maxnb):
            raise ValueError("spin %d has more than maxnb=%d table entries" % (i if fill[i] >= maxnb else j, maxnb))
        nbs[i, fill[i], 0] = j
        nbs[i, fill[i], 1] = v
        fill[i] += 1
        if j != i:
            nbs[j, fill[j], 0] = i
            nbs[j, fill[j], 1] = v
            fill[j] += 1
    if savepath is not None:
        np.save(savepath, nbs)
    return nbs


def ClassicalIsingEnergy(spins, J):
    """ClassicalIsingEnergy(spins, J): s^T offdiag(J) s + sum_i J_ii s_i (tools.pyx:99-118), sparse."""
    Jc = sps.csr_matrix(J)
    s = np.asarray(spins, dtype=np.float64)
    d = Jc.diagonal()
    off = Jc - sps.diags(d)
    return float(s @ (off @ s) + np.dot(d, s))
