"""ctypes front-end of the CPU oracle (oracle/mcs_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Parity status: PINNED against the compiled reference (oracle/_ref, built by
oracle/build_ref.py) and the committed fixtures under tests/golden/.

The functions mirror the reference's Python call surface (positional order of
/root/reference/solvers/{qmc,sa,svmc}.pyx) and mutate the state array in place, with one
addition: `rng`, a `LibcRand` stream standing in for the process-global libc rand() the
reference draws from (calling the reference after `srand(s)` == calling the oracle with
`LibcRand(s)`).  Functions that consume numpy's global generator in the reference
(`AnnealMA`, all of svmc except TFCompact) draw `np.random.uniform(...)` here at the same
point and with the same shape, so `np.random.seed(s)` reproduces the reference.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmcs_oracle.so")
_lib = None

c_dp = ctypes.POINTER(ctypes.c_double)
c_lp = ctypes.POINTER(ctypes.c_int64)
c_ip = ctypes.POINTER(ctypes.c_int32)


def build(force=False):
    """Compile mcs_oracle.c with gcc (oracle/Makefile)."""
    srcs = [os.path.join(_HERE, f) for f in ("mcs_oracle.c", "mcs_oracle_wolff.c")]
    if force or not os.path.isfile(_SO) or any(
            os.path.isfile(src) and os.path.getmtime(src) > os.path.getmtime(_SO) for src in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libmcs_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.mcs_rand_sizeof.restype = ctypes.c_size_t
        L.mcs_oracle_teff.restype = ctypes.c_double
        L.mcs_oracle_teff.argtypes = [ctypes.c_float, ctypes.c_int]
        L.mcs_oracle_ising_energy.restype = ctypes.c_double
        L.mcs_oracle_svmc_energy.restype = ctypes.c_double
        L.mcs_oracle_qmc_anneal.restype = ctypes.c_int
        L.mcs_oracle_qmc_dissipative.restype = ctypes.c_int
        L.mcs_oracle_qmc_wolff.restype = ctypes.c_int
        _lib = L
    return _lib


class LibcRand(object):
    """A private glibc rand() stream: LibcRand(s) behaves like the process after srand(s)."""

    def __init__(self, seed=1):
        L = lib()
        self._buf = ctypes.create_string_buffer(L.mcs_rand_sizeof())
        L.mcs_srand(self._buf, ctypes.c_uint32(int(seed) & 0xFFFFFFFF))

    @property
    def ptr(self):
        return self._buf

    def draw(self, n):
        out = np.empty(int(n), dtype=np.int32)
        lib().mcs_rand_fill(self._buf, out.ctypes.data_as(c_ip), ctypes.c_int64(int(n)))
        return out


GLOBAL_RAND = None  # created lazily; mimics the unseeded process-global stream (== srand(1))


def _rng(rng):
    global GLOBAL_RAND
    if rng is None:
        if GLOBAL_RAND is None:
            GLOBAL_RAND = LibcRand(1)
        return GLOBAL_RAND
    if isinstance(rng, LibcRand):
        return rng
    return LibcRand(int(rng))


def _f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a


def _nbs(nbs, ndim=3):
    nbs = np.ascontiguousarray(np.asarray(nbs), dtype=np.float64)
    if nbs.ndim != ndim:
        raise ValueError("Buffer has wrong number of dimensions (expected %d, got %d)" % (ndim, nbs.ndim))
    return nbs


def _check_spins(a, ndim):
    if not isinstance(a, np.ndarray) or a.dtype != np.int64:
        raise ValueError("Buffer dtype mismatch, expected 'int64_t'")
    if a.ndim != ndim:
        raise ValueError("Buffer has wrong number of dimensions (expected %d, got %d)" % (ndim, a.ndim))


def _estr(a, axis):
    s = a.strides[axis]
    assert s % a.itemsize == 0
    return ctypes.c_int64(s // a.itemsize)


# ----------------------------------------------------------------------------------------------
# qmc
# ----------------------------------------------------------------------------------------------
def _qmc(A_sched, B_sched, mcsteps, temp, confs, nbs, global_moves, rng):
    A = _f64(A_sched)
    B = _f64(B_sched)
    _check_spins(confs, 2)
    nbs = _nbs(nbs)
    rc = lib().mcs_oracle_qmc_anneal(
        A.ctypes.data_as(c_dp), B.ctypes.data_as(c_dp), ctypes.c_int(A.size), ctypes.c_int(int(mcsteps)),
        ctypes.c_float(temp), confs.ctypes.data_as(c_lp), _estr(confs, 0), _estr(confs, 1),
        ctypes.c_int(confs.shape[0]), ctypes.c_int(confs.shape[1]), nbs.ctypes.data_as(c_dp),
        ctypes.c_int(nbs.shape[1]), ctypes.c_int(int(global_moves)), _rng(rng).ptr)
    if rc == -1:
        raise ZeroDivisionError("float division")


def QuantumAnneal(A_sched, B_sched, mcsteps, temp, confs, nbs, nthreads=1, rng=None):
    """qmc.pyx:25-143."""
    _qmc(A_sched, B_sched, mcsteps, temp, confs, nbs, 0, rng)


def QuantumAnnealGlobal(A_sched, B_sched, mcsteps, temp, confs, nbs, nthreads=1, rng=None):
    """qmc.pyx:284-438."""
    _qmc(A_sched, B_sched, mcsteps, temp, confs, nbs, 1, rng)


def _qmc_diss(A_sched, B_sched, mcsteps, temp, lookuptable, confs, nbs, global_moves, rng):
    A = _f64(A_sched)
    B = _f64(B_sched)
    lut = _f64(lookuptable)
    _check_spins(confs, 2)
    nbs = _nbs(nbs)
    rc = lib().mcs_oracle_qmc_dissipative(
        A.ctypes.data_as(c_dp), B.ctypes.data_as(c_dp), ctypes.c_int(A.size), ctypes.c_int(int(mcsteps)),
        ctypes.c_float(temp), lut.ctypes.data_as(c_dp), confs.ctypes.data_as(c_lp), _estr(confs, 0),
        _estr(confs, 1), ctypes.c_int(confs.shape[0]), ctypes.c_int(confs.shape[1]),
        nbs.ctypes.data_as(c_dp), ctypes.c_int(nbs.shape[1]), ctypes.c_int(int(global_moves)),
        _rng(rng).ptr)
    if rc == -1:
        raise ZeroDivisionError("float division")


def DissipativeQuantumAnneal(A_sched, B_sched, mcsteps, temp, lookuptable, confs, nbs, nthreads=1, rng=None):
    """qmc.pyx:149-278."""
    _qmc_diss(A_sched, B_sched, mcsteps, temp, lookuptable, confs, nbs, 0, rng)


def DissipativeQuantumAnnealGlobal(A_sched, B_sched, mcsteps, temp, lookuptable, confs, nbs, nthreads=1,
                                   rng=None):
    """qmc.pyx:444-609."""
    _qmc_diss(A_sched, B_sched, mcsteps, temp, lookuptable, confs, nbs, 1, rng)


WOLFF_VARIANTS = {"QuantumAnnealWCL": 0, "DissaptiveQuantumAnnealWCL": 1, "QuantumAnnealWC": 2,
                  "DissipativeQuantumAnnealWC2": 3, "DissipativeQuantumAnnealWC3": 4}


def _qmc_wolff(variant, A_sched, B_sched, mcsteps, temp, lookuptable, confs, nbs, rng):
    A = _f64(A_sched)
    B = _f64(B_sched)
    lut = _f64(lookuptable) if lookuptable is not None else None
    _check_spins(confs, 2)
    nbs = _nbs(nbs)
    if lut is not None and lut.size < confs.shape[1] - 1:
        raise ValueError("lookuptable needs slices - 1 entries")
    rc = lib().mcs_oracle_qmc_wolff(
        ctypes.c_int(variant), A.ctypes.data_as(c_dp), B.ctypes.data_as(c_dp), ctypes.c_int(A.size),
        ctypes.c_int(int(mcsteps)), ctypes.c_float(temp), lut.ctypes.data_as(c_dp) if lut is not None else None,
        confs.ctypes.data_as(c_lp), _estr(confs, 0), _estr(confs, 1), ctypes.c_int(confs.shape[0]),
        ctypes.c_int(confs.shape[1]), nbs.ctypes.data_as(c_dp), ctypes.c_int(nbs.shape[1]), _rng(rng).ptr)
    if rc == -1:
        raise ZeroDivisionError("float division")
    if rc == -2:
        raise ValueError("the Wolff experiments need at least two Trotter slices")
    global last_wolff_overrun
    last_wolff_overrun = rc == 1


last_wolff_overrun = False  # the last call wrote past the reference's `cluster` buffer (undefined behaviour there)


def QuantumAnnealWCL(A_sched, B_sched, mcsteps, temp, confs, nbs, rng=None):
    """qmc.pyx:620-786."""
    _qmc_wolff(0, A_sched, B_sched, mcsteps, temp, None, confs, nbs, rng)


def DissaptiveQuantumAnnealWCL(A_sched, B_sched, mcsteps, temp, lookuptable, confs, nbs, rng=None):
    """qmc.pyx:792-1000."""
    _qmc_wolff(1, A_sched, B_sched, mcsteps, temp, lookuptable, confs, nbs, rng)


def QuantumAnnealWC(A_sched, B_sched, mcsteps, temp, confs, nbs, rng=None):
    """qmc.pyx:1006-1225."""
    _qmc_wolff(2, A_sched, B_sched, mcsteps, temp, None, confs, nbs, rng)


def DissipativeQuantumAnnealWC2(A_sched, B_sched, mcsteps, temp, lookuptable, confs, nbs, nthreads=1, rng=None):
    """qmc.pyx:1231-1446."""
    _qmc_wolff(3, A_sched, B_sched, mcsteps, temp, lookuptable, confs, nbs, rng)


def DissipativeQuantumAnnealWC3(A_sched, B_sched, mcsteps, temp, lookuptable, confs, nbs, nthreads=1, rng=None):
    """qmc.pyx:1452-1621."""
    _qmc_wolff(4, A_sched, B_sched, mcsteps, temp, lookuptable, confs, nbs, rng)


def qmc_delta_e(a, b, temp, confs, nbs):
    """Energy difference of every (spin, slice) visit for a frozen configuration (qmc.pyx:112-138)."""
    _check_spins(confs, 2)
    nbs = _nbs(nbs)
    out = np.empty(confs.shape, dtype=np.float64)
    lib().mcs_oracle_qmc_delta_e(
        ctypes.c_double(a), ctypes.c_double(b), ctypes.c_float(temp), confs.ctypes.data_as(c_lp),
        _estr(confs, 0), _estr(confs, 1), ctypes.c_int(confs.shape[0]), ctypes.c_int(confs.shape[1]),
        nbs.ctypes.data_as(c_dp), ctypes.c_int(nbs.shape[1]), out.ctypes.data_as(c_dp))
    return out


def qmc_delta_e_global(b, confs, nbs):
    """World-line flip energy differences (qmc.pyx:416-431)."""
    _check_spins(confs, 2)
    nbs = _nbs(nbs)
    out = np.empty(confs.shape[0], dtype=np.float64)
    lib().mcs_oracle_qmc_delta_e_global(
        ctypes.c_double(b), confs.ctypes.data_as(c_lp), _estr(confs, 0), _estr(confs, 1),
        ctypes.c_int(confs.shape[0]), ctypes.c_int(confs.shape[1]), nbs.ctypes.data_as(c_dp),
        ctypes.c_int(nbs.shape[1]), out.ctypes.data_as(c_dp))
    return out


def qmc_coeffs(a, b, temp, slices):
    """(teff, jperp, b_coeff) of one schedule step, qmc.pyx:85,95-96."""
    teff = lib().mcs_oracle_teff(ctypes.c_float(temp), ctypes.c_int(int(slices)))
    jp = ctypes.c_double()
    bc = ctypes.c_double()
    lib().mcs_oracle_qmc_coeffs(ctypes.c_double(a), ctypes.c_double(b), ctypes.c_double(teff),
                                ctypes.byref(jp), ctypes.byref(bc))
    return teff, jp.value, bc.value


# ----------------------------------------------------------------------------------------------
# sa
# ----------------------------------------------------------------------------------------------
def _sa(sched, mcsteps, svec, nbs, randuni, rng, noisy=False):
    sched = _f64(sched)
    _check_spins(svec, 1)
    nbs = _nbs(nbs, 4 if noisy else 3)
    maxnb = nbs.shape[-2]
    step_stride = nbs.shape[1] * nbs.shape[2] * 2 if noisy else 0
    ru = None
    if randuni is not None:
        ru = np.ascontiguousarray(randuni, dtype=np.float64)
    lib().mcs_oracle_sa_anneal(
        sched.ctypes.data_as(c_dp), ctypes.c_int(sched.size), ctypes.c_int(int(mcsteps)),
        svec.ctypes.data_as(c_lp), _estr(svec, 0), ctypes.c_int(svec.shape[0]),
        nbs.ctypes.data_as(c_dp), ctypes.c_int(maxnb), ctypes.c_int64(step_stride),
        ru.ctypes.data_as(c_dp) if ru is not None else None, _rng(rng).ptr)


def Anneal(sched, mcsteps, svec, nbs, rng=None):
    """sa.pyx:19-101."""
    _sa(sched, mcsteps, svec, nbs, None, rng)


def Anneal_parallel(sched, mcsteps, svec, nbs, nthreads=1, rng=None):
    """sa.pyx:201-284 (identical to Anneal when built without OpenMP)."""
    _sa(sched, mcsteps, svec, nbs, None, rng)


def AnnealMA(sched, mcsteps, svec, nbs, rng=None):
    """sa.pyx:108-193: acceptance uniforms pre-drawn from the global numpy generator (:151)."""
    sched = _f64(sched)
    randuni = np.random.uniform(size=(sched.size, int(mcsteps), svec.shape[0], 1))
    _sa(sched, mcsteps, svec, nbs, randuni, rng)


def NoisyAnneal(sched, mcsteps, svec, nbs, rng=None):
    """sa.pyx:291-378: time-dependent 4-D table nbs[sched, nspins, maxnb, 2]."""
    sched = _f64(sched)
    randuni = np.random.uniform(size=(sched.size, int(mcsteps), svec.shape[0], 1))
    _sa(sched, mcsteps, svec, nbs, randuni, rng, noisy=True)


def sa_delta_e(svec, nbs):
    _check_spins(svec, 1)
    nbs = _nbs(nbs)
    out = np.empty(svec.shape[0], dtype=np.float64)
    lib().mcs_oracle_sa_delta_e(svec.ctypes.data_as(c_lp), _estr(svec, 0), ctypes.c_int(svec.shape[0]),
                                nbs.ctypes.data_as(c_dp), ctypes.c_int(nbs.shape[1]),
                                out.ctypes.data_as(c_dp))
    return out


def ising_energy(svec, nbs):
    """Fixed-order fp64 classical energy in the reference's convention (tools.pyx:99-118)."""
    svec = np.asarray(svec)
    if svec.dtype != np.int64:
        svec = svec.astype(np.int64)
    nbs = _nbs(nbs)
    return lib().mcs_oracle_ising_energy(svec.ctypes.data_as(c_lp), _estr(svec, 0),
                                         ctypes.c_int(svec.shape[0]), nbs.ctypes.data_as(c_dp),
                                         ctypes.c_int(nbs.shape[1]))


# ----------------------------------------------------------------------------------------------
# svmc
# ----------------------------------------------------------------------------------------------
def _svmc(A_sched, B_sched, mcsteps, temp, svec, nbs, tf, rng, noisy=False, randuni=None):
    A = _f64(A_sched)
    B = _f64(B_sched)
    if not isinstance(svec, np.ndarray) or svec.dtype != np.float64:
        raise ValueError("Buffer dtype mismatch, expected 'float64_t'")
    nbs = _nbs(nbs, 4 if noisy else 3)
    maxnb = nbs.shape[-2]
    step_stride = nbs.shape[1] * nbs.shape[2] * 2 if noisy else 0
    if svec.ndim == 1:
        numreads, nspins, rs, ss = 1, svec.shape[0], 0, svec.strides[0] // 8
    else:
        numreads, nspins, rs, ss = svec.shape[0], svec.shape[1], svec.strides[0] // 8, svec.strides[1] // 8
    if randuni is None:
        randuni = np.random.uniform(size=(A.size, int(mcsteps), nspins, 2))  # svmc.pyx:70
    randuni = np.ascontiguousarray(randuni, dtype=np.float64)
    lib().mcs_oracle_svmc(
        A.ctypes.data_as(c_dp), B.ctypes.data_as(c_dp), ctypes.c_int(A.size), ctypes.c_int(int(mcsteps)),
        ctypes.c_float(temp), svec.ctypes.data_as(c_dp), ctypes.c_int64(rs), ctypes.c_int64(ss),
        ctypes.c_int(numreads), ctypes.c_int(nspins), nbs.ctypes.data_as(c_dp), ctypes.c_int(maxnb),
        ctypes.c_int64(step_stride), randuni.ctypes.data_as(c_dp), ctypes.c_int(int(tf)), _rng(rng).ptr)


def SpinVectorMonteCarlo(A_sched, B_sched, mcsteps, temp, svec, nbs, rng=None, randuni=None):
    """svmc.pyx:21-117."""
    assert svec.ndim == 1
    _svmc(A_sched, B_sched, mcsteps, temp, svec, nbs, 0, rng, randuni=randuni)


def SpinVectorMonteCarloTF(A_sched, B_sched, mcsteps, temp, svec, nbs, rng=None, randuni=None):
    """svmc.pyx:123-229."""
    assert svec.ndim == 1
    _svmc(A_sched, B_sched, mcsteps, temp, svec, nbs, 1, rng, randuni=randuni)


def NoisySVMC(A_sched, B_sched, mcsteps, temp, svec, nbs, rng=None, randuni=None):
    """svmc.pyx:236-334."""
    assert svec.ndim == 1
    _svmc(A_sched, B_sched, mcsteps, temp, svec, nbs, 0, rng, noisy=True, randuni=randuni)


def NoisySVMCTF(A_sched, B_sched, mcsteps, temp, svec, nbs, rng=None, randuni=None):
    """svmc.pyx:340-448."""
    assert svec.ndim == 1
    _svmc(A_sched, B_sched, mcsteps, temp, svec, nbs, 1, rng, noisy=True, randuni=randuni)


def SpinVectorMonteCarloCompact(A_sched, B_sched, mcsteps, temp, svec, nbs, rng=None, randuni=None):
    """svmc.pyx:455-554 (svec is [numreads, nspins]; one randuni shared by all reads)."""
    assert svec.ndim == 2
    _svmc(A_sched, B_sched, mcsteps, temp, svec, nbs, 0, rng, randuni=randuni)


def SpinVectorMonteCarloTFCompact(A_sched, B_sched, mcsteps, temp, svec, nbs, rng=None):
    """svmc.pyx:561-674 (all uniforms from rand())."""
    A = _f64(A_sched)
    B = _f64(B_sched)
    if not isinstance(svec, np.ndarray) or svec.dtype != np.float64 or svec.ndim != 2:
        raise ValueError("Buffer dtype mismatch, expected 'float64_t'")
    nbs = _nbs(nbs)
    lib().mcs_oracle_svmc_tf_compact(
        A.ctypes.data_as(c_dp), B.ctypes.data_as(c_dp), ctypes.c_int(A.size), ctypes.c_int(int(mcsteps)),
        ctypes.c_float(temp), svec.ctypes.data_as(c_dp), ctypes.c_int64(svec.strides[0] // 8),
        ctypes.c_int64(svec.strides[1] // 8), ctypes.c_int(svec.shape[0]), ctypes.c_int(svec.shape[1]),
        nbs.ctypes.data_as(c_dp), ctypes.c_int(nbs.shape[1]), _rng(rng).ptr)


def svmc_energy(a, b, svec, nbs):
    svec = np.ascontiguousarray(svec, dtype=np.float64)
    nbs = _nbs(nbs)
    return lib().mcs_oracle_svmc_energy(ctypes.c_double(a), ctypes.c_double(b), svec.ctypes.data_as(c_dp),
                                        ctypes.c_int64(1), ctypes.c_int(svec.shape[0]),
                                        nbs.ctypes.data_as(c_dp), ctypes.c_int(nbs.shape[1]))


# ----------------------------------------------------------------------------------------------
# tools (instance format) -- restated in numpy, tools.pyx:28-96 / 99-118
# ----------------------------------------------------------------------------------------------
def GenerateNeighbors(nspins, J, maxnb, savepath=None):
    """tools.pyx:28-96 restated: same row order (DOK key iteration order), O(nnz) instead of O(N*nnz)."""
    J = J.todok()
    nbs = np.zeros((nspins, maxnb, 2))
    fill = np.zeros(nspins, dtype=np.int64)
    for (i, j) in J.keys():
        v = J[i, j]
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
    """tools.pyx:99-118 verbatim semantics (dense)."""
    J = np.asarray(J.todense())
    d = np.diag(np.diag(J))
    np.fill_diagonal(J, 0.0)
    return np.dot(spins, np.dot(J, spins)) + np.sum(np.dot(d, spins))


# ----------------------------------------------------------------------------------------------
# coloured-order variants (not in the reference; see the comment in mcs_oracle.c)
# ----------------------------------------------------------------------------------------------
def _color_args(colors):
    colors = np.asarray(colors, dtype=np.int32)
    order = np.argsort(colors, kind="stable").astype(np.int32)
    nc = int(colors.max()) + 1
    start = np.zeros(nc + 1, dtype=np.int32)
    start[1:] = np.cumsum(np.bincount(colors, minlength=nc))
    return order, start, nc


def QuantumAnnealColored(A_sched, B_sched, mcsteps, temp, confs, nbs, colors, global_moves=False, rng=None):
    """qmc.pyx:93-143 visit arithmetic in the B200 kernels' visiting order (colour classes, slice parity)."""
    A = _f64(A_sched)
    B = _f64(B_sched)
    _check_spins(confs, 2)
    nbs = _nbs(nbs)
    order, start, nc = _color_args(colors)
    L = lib()
    L.mcs_oracle_qmc_anneal_colored.restype = ctypes.c_int
    rc = L.mcs_oracle_qmc_anneal_colored(
        A.ctypes.data_as(c_dp), B.ctypes.data_as(c_dp), ctypes.c_int(A.size), ctypes.c_int(int(mcsteps)),
        ctypes.c_float(temp), confs.ctypes.data_as(c_lp), _estr(confs, 0), _estr(confs, 1),
        ctypes.c_int(confs.shape[0]), ctypes.c_int(confs.shape[1]), nbs.ctypes.data_as(c_dp),
        ctypes.c_int(nbs.shape[1]), ctypes.c_int(int(bool(global_moves))), order.ctypes.data_as(c_ip),
        start.ctypes.data_as(c_ip), ctypes.c_int(nc), _rng(rng).ptr)
    if rc == -1:
        raise ZeroDivisionError("float division")


def AnnealColored(sched, mcsteps, svec, nbs, colors, rng=None):
    """sa.pyx:84-99 visit arithmetic in the B200 kernels' visiting order (colour class by colour class)."""
    sched = _f64(sched)
    _check_spins(svec, 1)
    nbs = _nbs(nbs)
    order, start, nc = _color_args(colors)
    lib().mcs_oracle_sa_anneal_colored(
        sched.ctypes.data_as(c_dp), ctypes.c_int(sched.size), ctypes.c_int(int(mcsteps)),
        svec.ctypes.data_as(c_lp), _estr(svec, 0), ctypes.c_int(svec.shape[0]), nbs.ctypes.data_as(c_dp),
        ctypes.c_int(nbs.shape[1]), order.ctypes.data_as(c_ip), start.ctypes.data_as(c_ip), ctypes.c_int(nc),
        _rng(rng).ptr)
