"""Drop-in for the reference's `solvers.sa` (classical simulated annealing sweeps).

Same names and positional arguments as /root/reference/solvers/sa.pyx; `svec` is mutated in place.
Extensions (keyword only): a leading restart axis on `svec` ([R, N]), `seed=`, `exact=True` +
`libc_seed=` (bit-exact sequential replay), `energies=True`.
"""
import numpy as np

from . import _common as C
from . import _lib

__all__ = ["Anneal", "AnnealMA", "Anneal_parallel", "NoisyAnneal"]


def _run(sched, mcsteps, svec, nbs, seed, exact, libc_seed, device, energies, replica_offset, randuni=None,
         nbs_ndim=3, dynamics=None):
    sched = np.asarray(sched)
    if sched.dtype != np.float64:
        raise ValueError("Buffer dtype mismatch, expected 'float64_t' but got '%s'" % sched.dtype)
    if sched.ndim != 1:
        raise ValueError("Buffer has wrong number of dimensions (expected 1, got %d)" % sched.ndim)
    sched = np.ascontiguousarray(sched)
    nbs = C.check_nbs(nbs, nbs_ndim)
    if nbs_ndim == 4 and not isinstance(nbs, _lib.Instance) and nbs.shape[0] < sched.size:
        raise ValueError("nbs needs one table per schedule step")
    a8, batched, need_copy = C.spins_in(svec, 1, "svec")
    R, N = a8.shape
    if _lib.DYNAMICS.get(dynamics) is None:
        raise ValueError("dynamics must be 'colored' or 'reference', got %r" % (dynamics,))
    inst = _lib.instance_for(nbs, device)
    if inst.nspins != N:
        raise ValueError("svec has %d spins but nbs describes %d" % (N, inst.nspins))
    with inst.using(dynamics):  # one call at a time per instance (shared scratch batch and stream)
        e_out = _execute(inst, sched, mcsteps, a8, R, seed, exact, libc_seed, energies, replica_offset, randuni)
    C.spins_out(svec, a8, batched, need_copy)
    if energies:
        return e_out if batched else e_out[0]
    return None


def _execute(inst, sched, mcsteps, a8, R, seed, exact, libc_seed, energies, replica_offset, randuni):
    L = _lib.load()
    e_out = np.empty(R, dtype=np.float64) if energies else None
    if exact:
        seeds = C.seeds_u32(libc_seed, R)
        ru = None
        if randuni is not None:
            ru = np.ascontiguousarray(randuni, dtype=np.float64)
        _lib.check(L.mcs_exact_sa(inst._h, _lib.dptr(sched), sched.size, int(mcsteps), a8.ctypes.data, R,
                                  C.u32p(seeds), _lib.dptr(ru) if ru is not None else None, None))
        if energies:
            st = _lib.State(inst, _lib.KIND_SA, R, 1)
            st.upload_spins(a8)
            e_out = st.energies()
            st.close()
    else:
        _lib.check(L.mcs_sa_anneal(inst._h, _lib.dptr(sched), sched.size, int(mcsteps), a8.ctypes.data, R,
                                   _lib.next_seed(seed), int(replica_offset),
                                   _lib.dptr(e_out) if energies else None))
    return e_out


def Anneal(sched, mcsteps, svec, nbs, *, seed=None, exact=False, libc_seed=None, device=None, energies=False,
           dynamics=None, replica_offset=0):
    """Anneal(sched, mcsteps, svec, nbs)

    Thermal annealing: for every temperature in `sched`, `mcsteps` Metropolis sweeps over all spins
    (reference sa.pyx:19-101).  A schedule may end at T = 0 (only downhill moves are then accepted).
    Returns None; spins are flipped in place within `svec` ([N] or [R, N])."""
    return _run(sched, mcsteps, svec, nbs, seed, exact, libc_seed, device, energies, replica_offset,
                dynamics=dynamics)


def AnnealMA(sched, mcsteps, svec, nbs, *, seed=None, exact=False, libc_seed=None, device=None, energies=False,
             dynamics=None, replica_offset=0):
    """AnnealMA(sched, mcsteps, svec, nbs)

    Reference sa.pyx:108-193: same sweeps as Anneal with the acceptance uniforms pre-drawn from
    numpy's global generator.  The production path has no such distinction (counter-based RNG);
    with exact=True the uniforms are drawn from np.random exactly as the reference does (:151)."""
    ru = None
    if exact:
        n = svec.shape[-1]
        ru = np.random.uniform(size=(np.asarray(sched).size, int(mcsteps), n, 1))
    return _run(sched, mcsteps, svec, nbs, seed, exact, libc_seed, device, energies, replica_offset, randuni=ru,
                dynamics=dynamics)


def Anneal_parallel(sched, mcsteps, svec, nbs, nthreads=1, *, seed=None, exact=False, libc_seed=None, device=None,
                    dynamics=None, energies=False, replica_offset=0):
    """Anneal_parallel(sched, mcsteps, svec, nbs, nthreads)

    Reference sa.pyx:201-284; identical to Anneal (its OpenMP pragmas are dead at build)."""
    return _run(sched, mcsteps, svec, nbs, seed, exact, libc_seed, device, energies, replica_offset,
                dynamics=dynamics)


def NoisyAnneal(sched, mcsteps, svec, nbs, *, seed=None, exact=False, libc_seed=None, device=None, energies=False,
                dynamics=None, replica_offset=0):
    """NoisyAnneal(sched, mcsteps, svec, nbs)

    Annealing with time-dependent couplings, `nbs` is [len(sched), nspins, maxnb, 2] and temperature step
    `itemp` uses nbs[itemp] (reference sa.pyx:291-378).  All tables are compiled into one device instance
    (one colouring from the union of the steps' graphs, one fp32 coupling row per step).  With exact=True
    the acceptance uniforms are drawn from np.random as the reference does (:336).  energies=True evaluates
    the LAST table."""
    ru = None
    if exact:
        ru = np.random.uniform(size=(np.asarray(sched).size, int(mcsteps), svec.shape[-1], 1))
    return _run(sched, mcsteps, svec, nbs, seed, exact, libc_seed, device, energies, replica_offset, randuni=ru,
                nbs_ndim=4, dynamics=dynamics)


def delta_e(svec, nbs, device=None):
    """fp64 energy difference of flipping each spin (sa.pyx:84-94); parity tier (a)."""
    nbs = C.check_nbs(nbs)
    a8, batched, _ = C.spins_in(svec, 1, "svec")
    R, N = a8.shape
    inst = _lib.instance_for(nbs, device)
    out = np.empty((R, N), dtype=np.float64)
    _lib.check(_lib.load().mcs_probe_sa_delta_e(inst._h, a8.ctypes.data, R, _lib.dptr(out)))
    return out if batched else out[0]
