"""Drop-in for the reference's `solvers.svmc` (spin-vector Monte Carlo: O(2) rotors, theta in [0, pi]).

Same names and positional arguments as /root/reference/solvers/svmc.pyx; `svec` (angles, float64) is
mutated in place.  The production kernels give every read its own counter-based stream
(`dynamics="reference"`: the reference's random-permutation sequential visiting order in distribution instead of
colour classes); with
exact=True + libc_seed= the reference's own streams (libc rand() for the shuffles, np.random for
the proposals, ONE randuni array shared by all reads of a Compact call) are replayed bit-exactly.
"""
import numpy as np

from . import _common as C
from . import _lib

__all__ = ["SpinVectorMonteCarlo", "SpinVectorMonteCarloTF", "SpinVectorMonteCarloCompact",
           "SpinVectorMonteCarloTFCompact", "NoisySVMC", "NoisySVMCTF"]


def _run(A_sched, B_sched, mcsteps, temp, svec, nbs, tf, ndim, seed, exact, libc_seed, device, replica_offset,
         rand_driven=False, randuni=None, nbs_ndim=3, dynamics=None):
    A = np.asarray(A_sched)
    B = np.asarray(B_sched)
    for s in (A, B):
        if s.dtype != np.float64:
            raise ValueError("Buffer dtype mismatch, expected 'float64_t' but got '%s'" % s.dtype)
        if s.ndim != 1:
            raise ValueError("Buffer has wrong number of dimensions (expected 1, got %d)" % s.ndim)
    A, B = np.ascontiguousarray(A), np.ascontiguousarray(B)
    if B.size < A.size:
        raise ValueError("B_sched is shorter than A_sched")
    nbs = C.check_nbs(nbs, nbs_ndim)
    if nbs_ndim == 4 and not isinstance(nbs, _lib.Instance) and nbs.shape[0] < A.size:
        raise ValueError("nbs needs one table per schedule step")
    a, need_copy = C.angles_in(svec, ndim, "svec")
    R, N = a.shape
    if _lib.DYNAMICS.get(dynamics) is None:
        raise ValueError("dynamics must be 'colored' or 'reference', got %r" % (dynamics,))
    inst = _lib.instance_for(nbs, device)
    if inst.nspins != N:
        raise ValueError("svec has %d spins but nbs describes %d" % (N, inst.nspins))
    with inst.using(dynamics):  # one call at a time per instance (shared scratch batch and stream)
        _execute(inst, A, B, mcsteps, temp, a, R, N, tf, ndim, seed, exact, libc_seed, replica_offset, rand_driven,
                 randuni)
    if need_copy:
        svec[...] = a if ndim == 2 else a[0]
    return None


def _execute(inst, A, B, mcsteps, temp, a, R, N, tf, ndim, seed, exact, libc_seed, replica_offset, rand_driven,
             randuni):
    L = _lib.load()
    temp = float(np.float32(temp))  # C float in the reference signature (svmc.pyx:24)
    if exact:
        ru = None
        if not rand_driven:
            if randuni is None:
                randuni = np.random.uniform(size=(A.size, int(mcsteps), N, 2))  # svmc.pyx:70
            ru = np.ascontiguousarray(randuni, dtype=np.float64)
        serial = 1 if ndim == 2 else 0
        seeds = C.seeds_u32(libc_seed, R) if not serial else np.full(R, int(libc_seed) & 0xFFFFFFFF, np.uint32)
        _lib.check(L.mcs_exact_svmc(inst._h, _lib.dptr(A), _lib.dptr(B), A.size, int(mcsteps), temp, a.ctypes.data,
                                    R, int(bool(tf)), C.u32p(seeds), _lib.dptr(ru) if ru is not None else None,
                                    serial))
    else:
        _lib.check(L.mcs_svmc_anneal(inst._h, _lib.dptr(A), _lib.dptr(B), A.size, int(mcsteps), temp, a.ctypes.data,
                                     R, int(bool(tf)), _lib.next_seed(seed), int(replica_offset)))


def SpinVectorMonteCarlo(A_sched, B_sched, mcsteps, temp, svec, nbs, *, seed=None, exact=False, libc_seed=None,
                         device=None, randuni=None, dynamics=None):
    """SpinVectorMonteCarlo(A_sched, B_sched, mcsteps, temp, svec, nbs)

    Rotor sweeps (reference svmc.pyx:21-117): propose theta' = pi*u, dE = B*sum_j J_ij (cos theta' -
    cos theta_i) cos theta_j + B*h_i (cos theta' - cos theta_i) + A (sin theta_i - sin theta'),
    Metropolis at temperature `temp`.  Returns None; `svec` [N] is updated in place."""
    return _run(A_sched, B_sched, mcsteps, temp, svec, nbs, False, 1, seed, exact, libc_seed, device, 0,
                randuni=randuni, dynamics=dynamics)


def SpinVectorMonteCarloTF(A_sched, B_sched, mcsteps, temp, svec, nbs, *, seed=None, exact=False, libc_seed=None,
                           device=None, randuni=None, dynamics=None):
    """SpinVectorMonteCarloTF(A_sched, B_sched, mcsteps, temp, svec, nbs)

    Transverse-field-restricted proposals theta' = clamp(theta + min(1, A/B) (2 pi u - pi), 0, pi)
    (reference svmc.pyx:123-229)."""
    return _run(A_sched, B_sched, mcsteps, temp, svec, nbs, True, 1, seed, exact, libc_seed, device, 0,
                randuni=randuni, dynamics=dynamics)


def SpinVectorMonteCarloCompact(A_sched, B_sched, mcsteps, temp, svec, nbs, *, seed=None, exact=False,
                                libc_seed=None, device=None, replica_offset=0, randuni=None, dynamics=None):
    """SpinVectorMonteCarloCompact(A_sched, B_sched, mcsteps, temp, svec, nbs)

    Batched form, `svec` is [numreads, N] (reference svmc.pyx:455-554)."""
    return _run(A_sched, B_sched, mcsteps, temp, svec, nbs, False, 2, seed, exact, libc_seed, device,
                replica_offset, randuni=randuni, dynamics=dynamics)


def SpinVectorMonteCarloTFCompact(A_sched, B_sched, mcsteps, temp, svec, nbs, *, seed=None, exact=False,
                                  libc_seed=None, device=None, replica_offset=0, dynamics=None):
    """SpinVectorMonteCarloTFCompact(A_sched, B_sched, mcsteps, temp, svec, nbs)

    Batched TF form (reference svmc.pyx:561-674; there all uniforms come from libc rand())."""
    return _run(A_sched, B_sched, mcsteps, temp, svec, nbs, True, 2, seed, exact, libc_seed, device,
                replica_offset, rand_driven=True, dynamics=dynamics)


def NoisySVMC(A_sched, B_sched, mcsteps, temp, svec, nbs, *, seed=None, exact=False, libc_seed=None, device=None,
              randuni=None, dynamics=None):
    """NoisySVMC(A_sched, B_sched, mcsteps, temp, svec, nbs)

    SpinVectorMonteCarlo with time-dependent couplings nbs[len(A_sched), nspins, maxnb, 2]: schedule step
    `ifield` uses nbs[ifield] (reference svmc.pyx:236-334)."""
    return _run(A_sched, B_sched, mcsteps, temp, svec, nbs, False, 1, seed, exact, libc_seed, device, 0,
                randuni=randuni, nbs_ndim=4, dynamics=dynamics)


def NoisySVMCTF(A_sched, B_sched, mcsteps, temp, svec, nbs, *, seed=None, exact=False, libc_seed=None, device=None,
                randuni=None, dynamics=None):
    """NoisySVMCTF(A_sched, B_sched, mcsteps, temp, svec, nbs)

    TF-restricted proposals with time-dependent couplings (reference svmc.pyx:340-448)."""
    return _run(A_sched, B_sched, mcsteps, temp, svec, nbs, True, 1, seed, exact, libc_seed, device, 0,
                randuni=randuni, nbs_ndim=4, dynamics=dynamics)
