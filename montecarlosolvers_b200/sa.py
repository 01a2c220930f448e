"""Drop-in for the reference's `solvers.sa` (classical simulated annealing sweeps).

Same names and positional arguments as /root/reference/solvers/sa.pyx; `svec` is mutated in place.
Extensions (keyword only): a leading restart axis on `svec` ([R, N]), `seed=`, `exact=True` +
`libc_seed=` (bit-exact sequential replay), `energies=True`.
"""
import numpy as np

from . import _common as C
from . import _lib

__all__ = ["Anneal", "AnnealMA", "Anneal_parallel", "NoisyAnneal"]


def _run(sched, mcsteps, svec, nbs, seed, exact, libc_seed, device, energies, replica_offset, randuni=None):
    sched = np.asarray(sched)
    if sched.dtype != np.float64:
        raise ValueError("Buffer dtype mismatch, expected 'float64_t' but got '%s'" % sched.dtype)
    if sched.ndim != 1:
        raise ValueError("Buffer has wrong number of dimensions (expected 1, got %d)" % sched.ndim)
    sched = np.ascontiguousarray(sched)
    nbs = C.check_nbs(nbs)
    a8, batched, need_copy = C.spins_in(svec, 1, "svec")
    R, N = a8.shape
    inst = _lib.instance_for(nbs, device)
    if inst.nspins != N:
        raise ValueError("svec has %d spins but nbs describes %d" % (N, inst.nspins))
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
    C.spins_out(svec, a8, batched, need_copy)
    if energies:
        return e_out if batched else e_out[0]
    return None


def Anneal(sched, mcsteps, svec, nbs, *, seed=None, exact=False, libc_seed=None, device=None, energies=False,
           replica_offset=0):
    """Anneal(sched, mcsteps, svec, nbs)

    Thermal annealing: for every temperature in `sched`, `mcsteps` Metropolis sweeps over all spins
    (reference sa.pyx:19-101).  A schedule may end at T = 0 (only downhill moves are then accepted).
    Returns None; spins are flipped in place within `svec` ([N] or [R, N])."""
    return _run(sched, mcsteps, svec, nbs, seed, exact, libc_seed, device, energies, replica_offset)


def AnnealMA(sched, mcsteps, svec, nbs, *, seed=None, exact=False, libc_seed=None, device=None, energies=False,
             replica_offset=0):
    """AnnealMA(sched, mcsteps, svec, nbs)

    Reference sa.pyx:108-193: same sweeps as Anneal with the acceptance uniforms pre-drawn from
    numpy's global generator.  The production path has no such distinction (counter-based RNG);
    with exact=True the uniforms are drawn from np.random exactly as the reference does (:151)."""
    ru = None
    if exact:
        n = svec.shape[-1]
        ru = np.random.uniform(size=(np.asarray(sched).size, int(mcsteps), n, 1))
    return _run(sched, mcsteps, svec, nbs, seed, exact, libc_seed, device, energies, replica_offset, randuni=ru)


def Anneal_parallel(sched, mcsteps, svec, nbs, nthreads=1, *, seed=None, exact=False, libc_seed=None, device=None,
                    energies=False, replica_offset=0):
    """Anneal_parallel(sched, mcsteps, svec, nbs, nthreads)

    Reference sa.pyx:201-284; identical to Anneal (its OpenMP pragmas are dead at build)."""
    return _run(sched, mcsteps, svec, nbs, seed, exact, libc_seed, device, energies, replica_offset)


def NoisyAnneal(sched, mcsteps, svec, nbs, **kw):
    """NoisyAnneal(sched, mcsteps, svec, nbs) -- time-dependent table nbs[sched, nspins, maxnb, 2]
    (reference sa.pyx:291-378): every temperature step is its own compiled instance."""
    sched = np.ascontiguousarray(sched, dtype=np.float64)
    nbs = C.check_nbs(nbs, 4)
    if nbs.shape[0] < sched.size:
        raise ValueError("nbs needs one table per schedule step")
    if kw.get("exact"):
        raise NotImplementedError("NoisyAnneal: exact replay is not implemented in this build")
    seed = _lib.next_seed(kw.pop("seed", None))
    a8, batched, need_copy = C.spins_in(svec, 1, "svec")
    R = a8.shape[0]
    inst0 = _lib.Instance(nbs[0], kw.get("device") or _lib.default_device())
    st = _lib.State(inst0, _lib.KIND_SA, R, 1)
    st.upload_spins(a8)
    cur = a8
    st.close()
    inst0.close()
    for t in range(sched.size):  # one compiled table per step; state round-trips through the host
        inst = _lib.Instance(nbs[t], kw.get("device") or _lib.default_device())
        s = _lib.State(inst, _lib.KIND_SA, R, 1)
        s.upload_spins(cur)
        s.sa_sweeps(sched[t:t + 1], mcsteps, seed=seed, sweep_offset=t * int(mcsteps))
        cur = s.download_spins()
        s.close()
        inst.close()
    a8[...] = cur
    C.spins_out(svec, a8, batched, need_copy)
    return None


def delta_e(svec, nbs, device=None):
    """fp64 energy difference of flipping each spin (sa.pyx:84-94); parity tier (a)."""
    nbs = C.check_nbs(nbs)
    a8, batched, _ = C.spins_in(svec, 1, "svec")
    R, N = a8.shape
    inst = _lib.instance_for(nbs, device)
    out = np.empty((R, N), dtype=np.float64)
    _lib.check(_lib.load().mcs_probe_sa_delta_e(inst._h, a8.ctypes.data, R, _lib.dptr(out)))
    return out if batched else out[0]
