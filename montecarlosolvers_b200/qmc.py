"""Drop-in for the reference's `solvers.qmc` (path-integral quantum Monte Carlo sweeps).

Same names and positional arguments as /root/reference/solvers/qmc.pyx; `confs` is mutated in place
and the functions return None.  Extensions (keyword only): a leading replica axis on `confs`
([R, N, P]: R independent anneals in one call), `seed=` for the counter-based RNG, `exact=True` +
`libc_seed=` for the bit-exact sequential replay of the reference, `energies=True` to get the
final per-slice classical energies back, `dynamics="reference"` for production sweeps that follow the
reference's visiting order in distribution (a fresh random permutation per slice, sequential visits, slices in
order -- qmc.pyx:99-143) instead of the faster checkerboard order.
"""
import numpy as np

from . import _common as C
from . import _lib

__all__ = ["QuantumAnneal", "QuantumAnnealGlobal", "anneal_best_slice", "DissipativeQuantumAnneal", "DissipativeQuantumAnnealGlobal",
           "QuantumAnnealSW", "QuantumAnnealWCL", "QuantumAnnealWC", "DissaptiveQuantumAnnealWCL",
           "DissipativeQuantumAnnealWC2", "DissipativeQuantumAnnealWC3"]


def _run(A_sched, B_sched, mcsteps, temp, confs, nbs, global_moves, lookuptable, seed, exact, libc_seed, device,
         energies, replica_offset, rand_stream=None, dynamics=None):
    _run.last_consumed = None
    A = _lib.f64(A_sched)
    B = _lib.f64(B_sched)
    if A.ndim != 1 or B.ndim != 1:
        raise ValueError("Buffer has wrong number of dimensions (expected 1)")
    if B.size < A.size:
        raise ValueError("B_sched is shorter than A_sched (undefined behaviour in the reference)")
    nbs = C.check_nbs(nbs)
    a8, batched, need_copy = C.spins_in(confs, 2, "confs")
    R, N, P = a8.shape
    if _lib.DYNAMICS.get(dynamics) is None:
        raise ValueError("dynamics must be 'colored' or 'reference', got %r" % (dynamics,))
    inst = _lib.instance_for(nbs, device)
    if inst.nspins != N:
        raise ValueError("confs has %d spins but nbs describes %d" % (N, inst.nspins))
    with inst.using(dynamics):  # one call at a time per instance (shared scratch batch and stream)
        e_out = _execute(inst, A, B, mcsteps, temp, a8, R, N, P, global_moves, lookuptable, seed, exact, libc_seed,
                         energies, replica_offset, rand_stream)
    C.spins_out(confs, a8, batched, need_copy)
    if energies:
        return e_out if batched else e_out[0]
    return None


def _execute(inst, A, B, mcsteps, temp, a8, R, N, P, global_moves, lookuptable, seed, exact, libc_seed, energies,
             replica_offset, rand_stream):
    L = _lib.load()
    temp = float(np.float32(temp))  # C float in the reference signature (qmc.pyx:28)
    e_out = np.empty((R, P), dtype=np.float64) if energies else None
    if lookuptable is not None and not exact:
        lut = _lib.f64(lookuptable)
        if lut.size < P - 1:
            raise ValueError("lookuptable needs P-1 entries")
        st = _lib.State(inst, _lib.KIND_PIQMC, R, P)
        try:
            st.upload_spins(a8)
            st.piqmc_sweeps_dissipative(A, B, mcsteps, temp, lut, global_moves=global_moves,
                                        seed=_lib.next_seed(seed), replica_offset=replica_offset)
            st.download_spins(a8)
            if energies:
                e_out = st.energies()
        finally:
            st.close()
    elif exact:
        lut = None
        if lookuptable is not None:
            lut = _lib.f64(lookuptable)
            if lut.size < P - 1:
                raise ValueError("lookuptable needs P-1 entries")
        consumed = np.zeros(R, dtype=np.int64)
        if rand_stream is not None:
            # recorded libc rand() outputs, int32 [R, n] (or [n] for a single anneal): the reference's own numbers
            rs = np.ascontiguousarray(np.asarray(rand_stream, dtype=np.int32).reshape(R, -1))
            _lib.check(L.mcs_exact_qmc(inst._h, _lib.dptr(A), _lib.dptr(B), A.size, int(mcsteps), temp,
                                       _lib.dptr(lut) if lut is not None else None, a8.ctypes.data, R, P,
                                       int(bool(global_moves)), None, rs.ctypes.data_as(_lib.c_i32p), rs.shape[1],
                                       consumed.ctypes.data_as(_lib.c_i64p)))
            if consumed.max() > rs.shape[1]:
                raise ValueError("rand_stream too short: the replay needed %d values" % consumed.max())
        else:
            seeds = C.seeds_u32(libc_seed, R)
            _lib.check(L.mcs_exact_qmc(inst._h, _lib.dptr(A), _lib.dptr(B), A.size, int(mcsteps), temp,
                                       _lib.dptr(lut) if lut is not None else None, a8.ctypes.data, R, P,
                                       int(bool(global_moves)), C.u32p(seeds), None, 0,
                                       consumed.ctypes.data_as(_lib.c_i64p)))
        _run.last_consumed = consumed
        if energies:
            st = _lib.State(inst, _lib.KIND_PIQMC, R, P) if P <= 64 else None
            if st is None:
                raise NotImplementedError("energies=True needs P <= 64")
            st.upload_spins(a8)
            e_out = st.energies()
            st.close()
    else:
        _lib.check(L.mcs_piqmc_anneal(inst._h, _lib.dptr(A), _lib.dptr(B), A.size, int(mcsteps), temp,
                                      a8.ctypes.data, R, P, int(bool(global_moves)), _lib.next_seed(seed),
                                      int(replica_offset), _lib.dptr(e_out) if energies else None))
    return e_out


def QuantumAnneal(A_sched, B_sched, mcsteps, temp, confs, nbs, nthreads=1, *, seed=None, exact=False, dynamics=None,
                  libc_seed=None, device=None, energies=False, replica_offset=0, rand_stream=None):
    """QuantumAnneal(A_sched, B_sched, mcsteps, temp, confs, nbs, nthreads)

    Path-integral QMC with single-spin flips (reference qmc.pyx:25-143).  H = sum_k (sum_ij J_ij
    s_i^k s_j^k - J_perp sum_i s_i^k s_i^{k+1}), J_perp = -(PT/2) ln tanh(A/(PT)); for every value of
    A_sched, `mcsteps` sweeps over all (spin, slice).  `nthreads` is accepted and ignored (it is
    inert in the reference too: OpenMP is disabled in its setup.py:10-11).
    Returns None; spins are flipped in place within `confs` ([N, P] or [R, N, P])."""
    return _run(A_sched, B_sched, mcsteps, temp, confs, nbs, False, None, seed, exact or rand_stream is not None,
                libc_seed, device, energies, replica_offset, rand_stream=rand_stream, dynamics=dynamics)


def QuantumAnnealGlobal(A_sched, B_sched, mcsteps, temp, confs, nbs, nthreads=1, *, seed=None, exact=False, dynamics=None,
                        libc_seed=None, device=None, energies=False, replica_offset=0, rand_stream=None):
    """QuantumAnnealGlobal(A_sched, B_sched, mcsteps, temp, confs, nbs, nthreads)

    As QuantumAnneal plus one world-line move per spin per sweep (all P slices of a spin flipped
    together, reference qmc.pyx:284-438)."""
    return _run(A_sched, B_sched, mcsteps, temp, confs, nbs, True, None, seed, exact or rand_stream is not None,
                libc_seed, device, energies, replica_offset, rand_stream=rand_stream, dynamics=dynamics)


def DissipativeQuantumAnneal(A_sched, B_sched, mcsteps, temp, lookuptable, confs, nbs, nthreads=1, *, seed=None,
                             exact=False, libc_seed=None, device=None, energies=False, replica_offset=0):
    """DissipativeQuantumAnneal(A_sched, B_sched, mcsteps, temp, lookuptable, confs, nbs, nthreads)

    PIQMC with the Ohmic-bath term sum_{d=1}^{P-1} 2 teff s_k s_{k+d} lookuptable[d-1] (reference
    qmc.pyx:149-278).  The bath couples all slices of a world line, so the production kernel visits the
    slices of a word in order (colour classes and replicas in parallel)."""
    return _run(A_sched, B_sched, mcsteps, temp, confs, nbs, False, lookuptable, seed, exact, libc_seed, device,
                energies, replica_offset)


def DissipativeQuantumAnnealGlobal(A_sched, B_sched, mcsteps, temp, lookuptable, confs, nbs, nthreads=1, *,
                                   seed=None, exact=False, libc_seed=None, device=None, energies=False,
                                   replica_offset=0):
    """DissipativeQuantumAnnealGlobal(A_sched, B_sched, mcsteps, temp, lookuptable, confs, nbs, nthreads)

    Reference qmc.pyx:444-609: bath term + one world-line move per spin per sweep."""
    return _run(A_sched, B_sched, mcsteps, temp, confs, nbs, True, lookuptable, seed, exact, libc_seed, device,
                energies, replica_offset)


def anneal_best_slice(A_sched, B_sched, mcsteps, temp, states, nbs, slices, *, global_moves=True, seed=None,
                      dynamics=None, device=None, replica_offset=0, per_slice_energies=False):
    """The per-anneal protocol of the reference's example (santoro80.py:286-296) in one device round trip:
    confs = tile(state, slices) -> QuantumAnneal[Global] -> E_k = ClassicalIsingEnergy(confs[:, k]) -> best slice.

    `states` is int8/int [R, N] (or [N]): one start configuration per anneal.  Returns (best_energy float64 [R],
    best_slice int32 [R], best_conf int8 [R, N]) and, with per_slice_energies=True, the energies [R, slices] as a
    fourth item.  Host traffic is R N bytes each way instead of the R N P of the drop-in call."""
    A, B = _lib.f64(A_sched), _lib.f64(B_sched)
    if A.ndim != 1 or B.ndim != 1:
        raise ValueError("Buffer has wrong number of dimensions (expected 1)")
    if B.size < A.size:
        raise ValueError("B_sched is shorter than A_sched (undefined behaviour in the reference)")
    nbs = C.check_nbs(nbs)
    a8, batched, _ = C.spins_in(states, 1, "states")
    R, N = a8.shape
    P = int(slices)
    if _lib.DYNAMICS.get(dynamics) is None:
        raise ValueError("dynamics must be 'colored' or 'reference', got %r" % (dynamics,))
    inst = _lib.instance_for(nbs, device)
    if inst.nspins != N:
        raise ValueError("states has %d spins but nbs describes %d" % (N, inst.nspins))
    e_best = np.empty(R, dtype=np.float64)
    k_best = np.empty(R, dtype=np.int32)
    conf = np.empty((R, N), dtype=np.int8)
    e_all = np.empty((R, P), dtype=np.float64) if per_slice_energies else None
    with inst.using(dynamics):
        _lib.check(_lib.load().mcs_piqmc_anneal_best(
            inst._h, _lib.dptr(A), _lib.dptr(B), A.size, int(mcsteps), float(np.float32(temp)), a8.ctypes.data, 1, R, P,
            int(bool(global_moves)), _lib.next_seed(seed), int(replica_offset),
            _lib.dptr(e_all) if per_slice_energies else None, _lib.dptr(e_best),
            k_best.ctypes.data_as(_lib.c_i32p), conf.ctypes.data))
    out = (e_best, k_best, conf) if batched else (e_best[0], k_best[0], conf[0])
    if per_slice_energies:
        out = out + ((e_all if batched else e_all[0]),)
    return out


def last_rand_consumed():
    """Number of libc rand() values each replica of the last exact=True call consumed (int64 [R])."""
    return _run.last_consumed


_run.last_consumed = None


def delta_e(a, b, temp, confs, nbs, device=None):
    """fp64 energy difference of every (spin, slice) visit for frozen configurations (qmc.pyx:112-138);
    parity tier (a).  Returns float64 with the shape of `confs`."""
    nbs = C.check_nbs(nbs)
    a8, batched, _ = C.spins_in(confs, 2, "confs")
    R, N, P = a8.shape
    inst = _lib.instance_for(nbs, device)
    out = np.empty((R, N, P), dtype=np.float64)
    _lib.check(_lib.load().mcs_probe_qmc_delta_e(inst._h, float(a), float(b), float(np.float32(temp)),
                                                 a8.ctypes.data, R, P, _lib.dptr(out)))
    return out if batched else out[0]


def delta_e_global(b, confs, nbs, device=None):
    """fp64 world-line flip energy differences (qmc.pyx:416-431)."""
    nbs = C.check_nbs(nbs)
    a8, batched, _ = C.spins_in(confs, 2, "confs")
    R, N, P = a8.shape
    inst = _lib.instance_for(nbs, device)
    out = np.empty((R, N), dtype=np.float64)
    _lib.check(_lib.load().mcs_probe_qmc_delta_e_global(inst._h, float(b), a8.ctypes.data, R, P, _lib.dptr(out)))
    return out if batched else out[0]


def QuantumAnnealSW(A_sched, B_sched, mcsteps, temp, confs, nbs, nthreads=1, *, cluster_every=1, global_moves=False,
                    lookuptable=None, seed=None, device=None, energies=False, replica_offset=0):
    """QuantumAnnealSW(A_sched, B_sched, mcsteps, temp, confs, nbs, nthreads)

    PIQMC with Swendsen-Wang cluster moves on the (space x Trotter) lattice: for every field value, `mcsteps`
    times { one single-spin sweep; every `cluster_every`-th time one cluster move (GPU union-find) }.
    `lookuptable` [P-1] adds the Ohmic bath of the Dissipative solvers to both (qmc.pyx:268-273).
    The reference advertises cluster updates (README.md:4) but ships only the experimental single-cluster
    Wolff variants of qmc.pyx:612-1621 (replayed bit-exactly by QuantumAnnealWCL & co. with exact=True); this is
    the working cluster move, validated against exact enumeration."""
    A = _lib.f64(A_sched)
    B = _lib.f64(B_sched)
    if B.size < A.size:
        raise ValueError("B_sched is shorter than A_sched")
    nbs = C.check_nbs(nbs)
    a8, batched, need_copy = C.spins_in(confs, 2, "confs")
    R, N, P = a8.shape
    lut = None
    if lookuptable is not None:
        lut = _lib.f64(lookuptable)
        if lut.ndim != 1 or lut.size < P - 1:
            raise ValueError("lookuptable needs slices - 1 entries")
    inst = _lib.instance_for(nbs, device)
    if inst.nspins != N:
        raise ValueError("confs has %d spins but nbs describes %d" % (N, inst.nspins))
    temp = float(np.float32(temp))
    sd = _lib.next_seed(seed)
    st = _lib.State(inst, _lib.KIND_PIQMC, R, P)
    e_out = None
    try:
        st.upload_spins(a8)
        sweep = 0
        for f in range(A.size):
            for step in range(int(mcsteps)):
                if lut is None:
                    st.piqmc_sweeps(A[f:f + 1], B[f:f + 1], 1, temp, global_moves=global_moves, seed=sd,
                                    replica_offset=replica_offset, sweep_offset=sweep)
                else:
                    st.piqmc_sweeps_dissipative(A[f:f + 1], B[f:f + 1], 1, temp, lut, global_moves=global_moves,
                                                seed=sd, replica_offset=replica_offset, sweep_offset=sweep)
                if cluster_every and (sweep + 1) % int(cluster_every) == 0:
                    st.cluster_moves(A[f], B[f], temp, 1, seed=sd, replica_offset=replica_offset, sweep_offset=sweep,
                                     lookuptable=lut)
                sweep += 1
        st.download_spins(a8)
        if energies:
            e_out = st.energies()
    finally:
        st.close()
    C.spins_out(confs, a8, batched, need_copy)
    if energies:
        return e_out if batched else e_out[0]
    return None


WOLFF_VARIANTS = {"QuantumAnnealWCL": 0, "DissaptiveQuantumAnnealWCL": 1, "QuantumAnnealWC": 2,
                  "DissipativeQuantumAnnealWC2": 3, "DissipativeQuantumAnnealWC3": 4}


def _wolff(name, A_sched, B_sched, mcsteps, temp, lookuptable, confs, nbs, exact, libc_seed, kw):
    """The reference's Wolff-cluster experiments (qmc.pyx:612-1621).  exact=True replays the reference function
    itself (sequential single-cluster growth, its rand() stream after srand(libc_seed), replica r <-> libc_seed + r)
    bit for bit.  Otherwise the call runs this library's working cluster dynamics with the same call surface: per
    step one single-spin sweep plus one Swendsen-Wang move on the same bond graph (QuantumAnnealSW)."""
    if not exact:
        if libc_seed is not None:
            raise ValueError("libc_seed= needs exact=True")
        return QuantumAnnealSW(A_sched, B_sched, mcsteps, temp, confs, nbs, 1, lookuptable=lookuptable, **kw)
    if kw.get("energies") or kw.get("cluster_every", 1) != 1 or kw.get("global_moves"):
        raise ValueError("exact=True replays the reference function as it is: no energies / cluster_every / "
                         "global_moves")
    A, B = _lib.f64(A_sched), _lib.f64(B_sched)
    if A.ndim != 1 or B.ndim != 1:
        raise ValueError("Buffer has wrong number of dimensions (expected 1)")
    if B.size < A.size:
        raise ValueError("B_sched is shorter than A_sched (undefined behaviour in the reference)")
    nbs = C.check_nbs(nbs)
    a8, batched, need_copy = C.spins_in(confs, 2, "confs")
    R, N, P = a8.shape
    if P < 2:
        raise ValueError("the Wolff experiments index Trotter slice 1: at least two slices")
    lut = None
    if lookuptable is not None:
        lut = _lib.f64(lookuptable)
        if lut.ndim != 1 or lut.size < P - 1:
            raise ValueError("lookuptable needs slices - 1 entries")
    inst = _lib.instance_for(nbs, kw.get("device"))
    if inst.nspins != N:
        raise ValueError("confs has %d spins but nbs describes %d" % (N, inst.nspins))
    temp = float(np.float32(temp))
    if temp * P == 0 and A.size:
        raise ZeroDivisionError("float division")
    seeds = C.seeds_u32(libc_seed, R)
    consumed = np.zeros(R, dtype=np.int64)
    overrun = np.zeros(R, dtype=np.int32)
    with inst.using(None):
        _lib.check(_lib.load().mcs_exact_qmc_wolff(
            inst._h, WOLFF_VARIANTS[name], _lib.dptr(A), _lib.dptr(B), A.size, int(mcsteps), temp,
            _lib.dptr(lut) if lut is not None else None, a8.ctypes.data, R, P, C.u32p(seeds),
            consumed.ctypes.data_as(_lib.c_i64p), overrun.ctypes.data_as(_lib.c_i32p)))
    _run.last_consumed = consumed
    _wolff.last_overrun = overrun.astype(bool)
    C.spins_out(confs, a8, batched, need_copy)
    return None


_wolff.last_overrun = None


def last_wolff_overrun():
    """bool [R]: replicas of the last exact Wolff call in which the REFERENCE would have written past its unchecked
    `cluster` buffer (qmc.pyx:685, 1310: the un-flipped seed re-joined a full cluster) -- undefined behaviour there,
    well defined here."""
    return _wolff.last_overrun


def QuantumAnnealWCL(A_sched, B_sched, mcsteps, temp, confs, nbs, *, exact=False, libc_seed=None, **kw):
    """QuantumAnnealWCL(A_sched, B_sched, mcsteps, temp, confs, nbs)

    Reference qmc.pyx:620-786 (no `nthreads`): one single-cluster Wolff move per step, bond test on the bond
    energy and the candidate's field.  See _wolff for exact= / the production dynamics."""
    return _wolff("QuantumAnnealWCL", A_sched, B_sched, mcsteps, temp, None, confs, nbs, exact, libc_seed, kw)


def QuantumAnnealWC(A_sched, B_sched, mcsteps, temp, confs, nbs, *, exact=False, libc_seed=None, **kw):
    """QuantumAnnealWC(A_sched, B_sched, mcsteps, temp, confs, nbs)

    Reference qmc.pyx:1006-1225: as WCL with the bond test on the candidate's full energy change."""
    return _wolff("QuantumAnnealWC", A_sched, B_sched, mcsteps, temp, None, confs, nbs, exact, libc_seed, kw)


def DissaptiveQuantumAnnealWCL(A_sched, B_sched, mcsteps, temp, lookuptable, confs, nbs, *, exact=False,
                               libc_seed=None, **kw):
    """DissaptiveQuantumAnnealWCL(A_sched, B_sched, mcsteps, temp, lookuptable, confs, nbs)

    Reference qmc.pyx:792-1000 (the reference's spelling): WCL with Ohmic-bath bonds between all slices of a world
    line."""
    return _wolff("DissaptiveQuantumAnnealWCL", A_sched, B_sched, mcsteps, temp, lookuptable, confs, nbs, exact,
                  libc_seed, kw)


def DissipativeQuantumAnnealWC2(A_sched, B_sched, mcsteps, temp, lookuptable, confs, nbs, nthreads=1, *, exact=False,
                                libc_seed=None, **kw):
    """DissipativeQuantumAnnealWC2(A_sched, B_sched, mcsteps, temp, lookuptable, confs, nbs, nthreads)

    Reference qmc.pyx:1231-1446: per step a local sweep, then one bath-only cluster per spin with a Metropolis
    test on the cluster's remaining energy change."""
    return _wolff("DissipativeQuantumAnnealWC2", A_sched, B_sched, mcsteps, temp, lookuptable, confs, nbs, exact,
                  libc_seed, kw)


def DissipativeQuantumAnnealWC3(A_sched, B_sched, mcsteps, temp, lookuptable, confs, nbs, nthreads=1, *, exact=False,
                                libc_seed=None, **kw):
    """DissipativeQuantumAnnealWC3(A_sched, B_sched, mcsteps, temp, lookuptable, confs, nbs, nthreads)

    Reference qmc.pyx:1452-1621: per step one bath-only cluster per (spin, slice)."""
    return _wolff("DissipativeQuantumAnnealWC3", A_sched, B_sched, mcsteps, temp, lookuptable, confs, nbs, exact,
                  libc_seed, kw)
