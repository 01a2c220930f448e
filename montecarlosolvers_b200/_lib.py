"""ctypes binding of libmcs_b200.so (the C ABI declared in include/mcs_b200.h).

There is no CPU fallback: if the shared library is missing (and cannot be built with nvcc) the
import of any compute entry point raises, and every compute call fails with MCS_ENODEVICE when no
CUDA device is visible.
"""
import contextlib
import ctypes
import hashlib
import os
import threading
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("MCS_B200_LIB") or os.path.join(_HERE, "libmcs_b200.so")  # override: kernel experiments

MCS_OK, MCS_EINVAL, MCS_ENODEVICE, MCS_EZERODIV, MCS_EUNSUPPORTED, MCS_ENOMEM = 0, -1, -2, -3, -4, -5
KIND_PIQMC, KIND_SA, KIND_SVMC = 1, 2, 3
DYNAMICS = {"colored": 0, "coloured": 0, "checkerboard": 0, 0: 0, None: 0, "reference": 1, 1: 1}

c_i64 = ctypes.c_int64
c_u64 = ctypes.c_uint64
c_dp = ctypes.POINTER(ctypes.c_double)
c_i8p = ctypes.POINTER(ctypes.c_int8)
c_i32p = ctypes.POINTER(ctypes.c_int32)
c_u32p = ctypes.POINTER(ctypes.c_uint32)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_vp = ctypes.c_void_p

# name -> (restype, argtypes); mirrors include/mcs_b200.h one to one
SIGNATURES = {
    "mcs_abi_version": (ctypes.c_int, []),
    "mcs_last_error": (ctypes.c_char_p, []),
    "mcs_device_count": (ctypes.c_int, []),
    "mcs_host_alloc": (c_vp, [ctypes.c_size_t]),
    "mcs_host_free": (None, [c_vp]),
    "mcs_instance_create": (ctypes.c_int, [c_dp, c_i64, c_i64, ctypes.c_int, ctypes.POINTER(c_vp)]),
    "mcs_instance_create_steps": (ctypes.c_int, [c_dp, c_i64, c_i64, c_i64, ctypes.c_int, ctypes.POINTER(c_vp)]),
    "mcs_instance_destroy": (None, [c_vp]),
    "mcs_instance_info": (ctypes.c_int, [c_vp, c_i64p]),
    "mcs_instance_colors": (ctypes.c_int, [c_vp, c_i32p]),
    "mcs_instance_set_dense": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "mcs_instance_set_dynamics": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "mcs_timer_start": (ctypes.c_int, [c_vp]),
    "mcs_timer_stop": (ctypes.c_int, [c_vp, c_dp]),
    "mcs_synchronize": (ctypes.c_int, [c_vp]),
    "mcs_instance_trim": (ctypes.c_int, [c_vp]),
    "mcs_launch_count": (c_i64, [c_vp]),
    "mcs_state_create": (ctypes.c_int, [c_vp, ctypes.c_int, c_i64, c_i64, ctypes.POINTER(c_vp)]),
    "mcs_state_destroy": (None, [c_vp]),
    "mcs_state_upload_spins": (ctypes.c_int, [c_vp, c_vp]),
    "mcs_state_download_spins": (ctypes.c_int, [c_vp, c_vp]),
    "mcs_state_upload_angles": (ctypes.c_int, [c_vp, c_vp]),
    "mcs_state_download_angles": (ctypes.c_int, [c_vp, c_vp]),
    "mcs_state_init_random": (ctypes.c_int, [c_vp, c_u64, c_u64]),
    "mcs_state_energies": (ctypes.c_int, [c_vp, c_dp]),
    "mcs_state_best": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, ctypes.c_int]),
    "mcs_state_svmc_energies": (ctypes.c_int, [c_vp, ctypes.c_double, ctypes.c_double, c_dp]),
    "mcs_piqmc_sweeps": (ctypes.c_int, [c_vp, c_dp, c_dp, c_i64, ctypes.c_int, ctypes.c_float, ctypes.c_int,
                                        c_u64, c_u64, c_u64]),
    "mcs_piqmc_sweeps_dissipative": (ctypes.c_int, [c_vp, c_dp, c_dp, c_i64, ctypes.c_int, ctypes.c_float, c_dp,
                                                    ctypes.c_int, c_u64, c_u64, c_u64]),
    "mcs_sa_sweeps": (ctypes.c_int, [c_vp, c_dp, c_i64, ctypes.c_int, c_u64, c_u64, c_u64]),
    "mcs_svmc_sweeps": (ctypes.c_int, [c_vp, c_dp, c_dp, c_i64, ctypes.c_int, ctypes.c_float, ctypes.c_int,
                                       c_u64, c_u64, c_u64]),
    "mcs_cluster_moves": (ctypes.c_int, [c_vp, ctypes.c_double, ctypes.c_double, ctypes.c_float, ctypes.c_int, c_u64,
                                         c_u64, c_u64]),
    "mcs_cluster_moves_dissipative": (ctypes.c_int, [c_vp, ctypes.c_double, ctypes.c_double, ctypes.c_float, c_dp,
                                                     ctypes.c_int, c_u64, c_u64, c_u64]),
    "mcs_piqmc_anneal": (ctypes.c_int, [c_vp, c_dp, c_dp, c_i64, ctypes.c_int, ctypes.c_float, c_vp, c_i64, c_i64,
                                        ctypes.c_int, c_u64, c_u64, c_dp]),
    "mcs_piqmc_anneal_best": (ctypes.c_int, [c_vp, c_dp, c_dp, c_i64, ctypes.c_int, ctypes.c_float, c_vp, ctypes.c_int,
                                             c_i64, c_i64, ctypes.c_int, c_u64, c_u64, c_dp, c_dp, c_i32p, c_vp]),
    "mcs_sa_anneal": (ctypes.c_int, [c_vp, c_dp, c_i64, ctypes.c_int, c_vp, c_i64, c_u64, c_u64, c_dp]),
    "mcs_svmc_anneal": (ctypes.c_int, [c_vp, c_dp, c_dp, c_i64, ctypes.c_int, ctypes.c_float, c_vp, c_i64,
                                       ctypes.c_int, c_u64, c_u64]),
    "mcs_exact_qmc": (ctypes.c_int, [c_vp, c_dp, c_dp, c_i64, ctypes.c_int, ctypes.c_float, c_dp, c_vp, c_i64, c_i64,
                                     ctypes.c_int, c_u32p, c_i32p, c_i64, c_i64p]),
    "mcs_exact_qmc_wolff": (ctypes.c_int, [c_vp, ctypes.c_int, c_dp, c_dp, c_i64, ctypes.c_int, ctypes.c_float, c_dp,
                                           c_vp, c_i64, c_i64, c_u32p, c_i64p, c_i32p]),
    "mcs_exact_sa": (ctypes.c_int, [c_vp, c_dp, c_i64, ctypes.c_int, c_vp, c_i64, c_u32p, c_dp, c_i64p]),
    "mcs_exact_svmc": (ctypes.c_int, [c_vp, c_dp, c_dp, c_i64, ctypes.c_int, ctypes.c_float, c_vp, c_i64,
                                      ctypes.c_int, c_u32p, c_dp, ctypes.c_int]),
    "mcs_probe_qmc_delta_e": (ctypes.c_int, [c_vp, ctypes.c_double, ctypes.c_double, ctypes.c_float, c_vp, c_i64,
                                             c_i64, c_dp]),
    "mcs_probe_qmc_delta_e_global": (ctypes.c_int, [c_vp, ctypes.c_double, c_vp, c_i64, c_i64, c_dp]),
    "mcs_probe_sa_delta_e": (ctypes.c_int, [c_vp, c_vp, c_i64, c_dp]),
}

_lib = None
_lock = threading.Lock()


class McsError(RuntimeError):
    pass


def load():
    """Load (building first if the sources are newer and nvcc is present) the shared library."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        from . import build as _build
        if not os.path.isfile(SO_PATH):
            _build.build()  # raises when nvcc is absent: no CPU fallback
        elif not _build.up_to_date() and os.path.isdir(_build.CSRC):
            try:  # a source is newer than the binary: rebuild where nvcc exists, else say so and load what is there
                _build.build()
            except RuntimeError:
                import warnings
                warnings.warn("libmcs_b200.so is older than its sources and nvcc is not available to rebuild it")
        L = ctypes.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.mcs_abi_version() != 1:
            raise McsError("libmcs_b200.so ABI version mismatch")
        _lib = L
        return _lib


def check(rc):
    """Map a status code to the Python exception the reference would raise."""
    if rc == MCS_OK:
        return
    msg = load().mcs_last_error().decode("utf-8", "replace")
    if rc == MCS_EZERODIV:
        raise ZeroDivisionError(msg or "float division")
    if rc == MCS_EINVAL:
        raise ValueError(msg)
    if rc == MCS_EUNSUPPORTED:
        raise NotImplementedError(msg)
    if rc == MCS_ENOMEM:
        raise MemoryError(msg)
    raise McsError(msg)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def dptr(a):
    return a.ctypes.data_as(c_dp)


def device_count():
    return load().mcs_device_count()


def require_device():
    if device_count() < 1:
        raise McsError("montecarlosolvers_b200: no CUDA device visible; this package has no CPU fallback")


# ---------------------------------------------------------------------------------------------
# pinned host arrays
# ---------------------------------------------------------------------------------------------
class _Pinned(object):
    def __init__(self, nbytes):
        self.ptr = load().mcs_host_alloc(nbytes)
        if not self.ptr:
            raise McsError(load().mcs_last_error().decode())
        self.nbytes = nbytes

    def __del__(self):
        try:
            if self.ptr:
                load().mcs_host_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


def empty_pinned(shape, dtype):
    """numpy array backed by page-locked host memory (cudaHostAlloc): full-speed H2D / D2H."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    owner = _Pinned(max(n, 1))
    buf = (ctypes.c_char * max(n, 1)).from_address(owner.ptr)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    arr = arr.view(_PinnedArray)
    arr._owner = owner
    return arr


class _PinnedArray(np.ndarray):
    _owner = None

    def __array_finalize__(self, obj):
        if obj is not None:
            self._owner = getattr(obj, "_owner", None)


# ---------------------------------------------------------------------------------------------
# compiled instances and resident replica batches
# ---------------------------------------------------------------------------------------------
class Instance(object):
    """A neighbour table (tools.GenerateNeighbors format) compiled and resident on one GPU."""

    def __init__(self, nbs, device=0):
        nbs = np.ascontiguousarray(np.asarray(nbs), dtype=np.float64)
        if nbs.ndim not in (3, 4):
            raise ValueError("Buffer has wrong number of dimensions (expected 3, got %d)" % nbs.ndim)
        if nbs.shape[-1] != 2:
            raise ValueError("nbs must be [nspins, maxnb, 2] (or [nsteps, nspins, maxnb, 2] for the Noisy solvers)")
        self.nsteps = int(nbs.shape[0]) if nbs.ndim == 4 else 1
        self.nspins, self.maxnb = int(nbs.shape[-3]), int(nbs.shape[-2])
        self.device = int(device)
        h = c_vp()
        check(load().mcs_instance_create_steps(dptr(nbs), self.nsteps, self.nspins, self.maxnb, self.device,
                                               ctypes.byref(h)))
        self._h = h
        self._states = weakref.WeakSet()
        info = (c_i64 * 8)()
        check(load().mcs_instance_info(self._h, info))
        self.ncolors, self.maxdeg = int(info[2]), int(info[3])
        self.has_field, self.lut_kernels = bool(info[4]), bool(info[6])
        self.dense = bool(int(info[7]) >> 32)
        self.dynamics = "colored"
        # One-shot calls share this instance's scratch batch, staging buffer and stream, and ctypes releases the
        # GIL during the C call: concurrent Python threads annealing the same problem serialise on this lock.
        self.lock = threading.RLock()

    def set_dynamics(self, dynamics):
        """"colored" (default, fastest) or "reference" (the reference's random-permutation sequential order in
        distribution, include/mcs_b200.h: mcs_instance_set_dynamics)."""
        mode = DYNAMICS.get(dynamics)
        if mode is None:
            raise ValueError("dynamics must be 'colored' or 'reference', got %r" % (dynamics,))
        check(load().mcs_instance_set_dynamics(self._h, mode))
        self.dynamics = "reference" if mode else "colored"

    @contextlib.contextmanager
    def using(self, dynamics=None):
        """Serialise the one-shot calls on this instance and run them with the given dynamics."""
        with self.lock:
            self.set_dynamics(dynamics)
            try:
                yield self
            finally:
                self.set_dynamics(None)

    def colors(self):
        out = np.empty(self.nspins, dtype=np.int32)
        check(load().mcs_instance_colors(self._h, out.ctypes.data_as(c_i32p)))
        return out

    def use_dense(self, enable=True):
        """Dense instances: switch between the blocked tensor-core sweeps and the general coloured kernels."""
        check(load().mcs_instance_set_dense(self._h, int(bool(enable))))
        self.dense = bool(enable)

    def timer_start(self):
        check(load().mcs_timer_start(self._h))

    def timer_stop(self):
        ms = ctypes.c_double()
        check(load().mcs_timer_stop(self._h, ctypes.byref(ms)))
        return ms.value

    def synchronize(self):
        check(load().mcs_synchronize(self._h))

    def trim(self):
        """Release the device batch the one-shot calls keep cached between calls."""
        check(load().mcs_instance_trim(self._h))

    @property
    def launches(self):
        return int(load().mcs_launch_count(self._h))

    def close(self):
        if getattr(self, "_h", None):
            for st in list(self._states):
                st.close()
            load().mcs_instance_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class State(object):
    """A replica batch resident in HBM (PIQMC world lines, SA restarts or SVMC rotors)."""

    def __init__(self, inst, kind, R, P=1):
        self.inst, self.kind, self.R, self.P = inst, kind, int(R), int(P)
        h = c_vp()
        check(load().mcs_state_create(inst._h, kind, self.R, self.P, ctypes.byref(h)))
        self._h = h
        inst._states.add(self)

    def _spin_shape(self):
        return (self.R, self.inst.nspins, self.P) if self.kind == KIND_PIQMC else (self.R, self.inst.nspins)

    def upload_spins(self, a):
        a = np.asarray(a)
        assert a.dtype == np.int8 and a.flags.c_contiguous and a.shape == self._spin_shape(), (a.dtype, a.shape)
        check(load().mcs_state_upload_spins(self._h, a.ctypes.data))

    def download_spins(self, out=None):
        if out is None:
            out = np.empty(self._spin_shape(), dtype=np.int8)
        assert out.dtype == np.int8 and out.flags.c_contiguous and out.shape == self._spin_shape()
        check(load().mcs_state_download_spins(self._h, out.ctypes.data))
        return out

    def upload_angles(self, a):
        a = np.asarray(a)
        assert a.dtype == np.float64 and a.flags.c_contiguous and a.shape == (self.R, self.inst.nspins)
        check(load().mcs_state_upload_angles(self._h, a.ctypes.data))

    def download_angles(self, out=None):
        if out is None:
            out = np.empty((self.R, self.inst.nspins), dtype=np.float64)
        check(load().mcs_state_download_angles(self._h, out.ctypes.data))
        return out

    def init_random(self, seed, replica_offset=0):
        check(load().mcs_state_init_random(self._h, int(seed) & (2 ** 64 - 1), int(replica_offset)))

    def energies(self):
        out = np.empty((self.R, self.P) if self.kind == KIND_PIQMC else (self.R,), dtype=np.float64)
        check(load().mcs_state_energies(self._h, dptr(out)))
        return out

    def best(self, conf=True):
        """(best-slice energy float64 [R], best slice int32 [R], that slice's spins int8 [R, N] or None), evaluated
        on the device (santoro80.py:290-296 without downloading the world lines)."""
        e = np.empty(self.R, dtype=np.float64)
        k = np.empty(self.R, dtype=np.int32)
        c = np.empty((self.R, self.inst.nspins), dtype=np.int8) if conf else None
        check(load().mcs_state_best(self._h, e.ctypes.data, k.ctypes.data, c.ctypes.data if conf else None, 0))
        return e, k, c

    def best_into(self, energy_ptr, slice_ptr, conf_ptr):
        """As best(), into DEVICE buffers given by address (e.g. torch.Tensor.data_ptr()); asynchronous on the
        instance's stream -- call inst.synchronize() before another stream reads them."""
        check(load().mcs_state_best(self._h, energy_ptr or None, slice_ptr or None, conf_ptr or None, 1))

    def svmc_energies(self, a, b):
        out = np.empty(self.R, dtype=np.float64)
        check(load().mcs_state_svmc_energies(self._h, float(a), float(b), dptr(out)))
        return out

    def piqmc_sweeps(self, A, B, mcsteps, temp, global_moves=False, seed=0, replica_offset=0, sweep_offset=0):
        A, B = f64(A), f64(B)
        if B.size < A.size:
            raise ValueError("B_sched shorter than A_sched")
        check(load().mcs_piqmc_sweeps(self._h, dptr(A), dptr(B), A.size, int(mcsteps), float(temp),
                                      int(bool(global_moves)), int(seed) & (2 ** 64 - 1), int(replica_offset),
                                      int(sweep_offset)))

    def piqmc_sweeps_dissipative(self, A, B, mcsteps, temp, lookuptable, global_moves=False, seed=0,
                                 replica_offset=0, sweep_offset=0):
        A, B, lut = f64(A), f64(B), f64(lookuptable)
        if B.size < A.size:
            raise ValueError("B_sched shorter than A_sched")
        if lut.size < self.P - 1:
            raise ValueError("lookuptable needs P-1 entries")
        check(load().mcs_piqmc_sweeps_dissipative(self._h, dptr(A), dptr(B), A.size, int(mcsteps), float(temp),
                                                  dptr(lut), int(bool(global_moves)), int(seed) & (2 ** 64 - 1),
                                                  int(replica_offset), int(sweep_offset)))

    def cluster_moves(self, a, b, temp, nmoves=1, seed=0, replica_offset=0, sweep_offset=0, lookuptable=None):
        """Swendsen-Wang moves at transverse field a, longitudinal coefficient b, temperature temp (SA: a, b unused);
        lookuptable [P-1]: with the Ohmic-bath bonds of the Dissipative solvers (qmc.pyx:268-273)."""
        if lookuptable is None:
            check(load().mcs_cluster_moves(self._h, float(a), float(b), float(temp), int(nmoves),
                                           int(seed) & (2 ** 64 - 1), int(replica_offset), int(sweep_offset)))
            return
        lut = f64(lookuptable)
        if lut.size < self.P - 1:
            raise ValueError("lookuptable needs P-1 entries")
        check(load().mcs_cluster_moves_dissipative(self._h, float(a), float(b), float(temp), dptr(lut), int(nmoves),
                                                   int(seed) & (2 ** 64 - 1), int(replica_offset),
                                                   int(sweep_offset)))

    def sa_sweeps(self, sched, mcsteps, seed=0, replica_offset=0, sweep_offset=0):
        sched = f64(sched)
        check(load().mcs_sa_sweeps(self._h, dptr(sched), sched.size, int(mcsteps), int(seed) & (2 ** 64 - 1),
                                   int(replica_offset), int(sweep_offset)))

    def svmc_sweeps(self, A, B, mcsteps, temp, tf=False, seed=0, replica_offset=0, sweep_offset=0):
        A, B = f64(A), f64(B)
        if B.size < A.size:
            raise ValueError("B_sched shorter than A_sched")
        check(load().mcs_svmc_sweeps(self._h, dptr(A), dptr(B), A.size, int(mcsteps), float(temp), int(bool(tf)),
                                     int(seed) & (2 ** 64 - 1), int(replica_offset), int(sweep_offset)))

    def close(self):
        if getattr(self, "_h", None):
            load().mcs_state_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------
# instance cache for the drop-in calls (the reference re-reads `nbs` on every call)
# ---------------------------------------------------------------------------------------------
_cache = {}
_CACHE_MAX = 8


def instance_for(nbs, device=None):
    if isinstance(nbs, Instance):
        return nbs
    if device is None:
        device = default_device()
    nbs = np.ascontiguousarray(np.asarray(nbs), dtype=np.float64)
    key = (nbs.shape, int(device), hashlib.blake2b(nbs.view(np.uint8).reshape(-1), digest_size=16).digest())
    inst = _cache.get(key)
    if inst is None:
        if len(_cache) >= _CACHE_MAX:
            _cache.pop(next(iter(_cache)))
        inst = Instance(nbs, device)
        _cache[key] = inst
    return inst


def default_device():
    """LOCAL_RANK under torchrun (one process per GPU), else 0."""
    try:
        return int(os.environ.get("MCS_DEVICE", os.environ.get("LOCAL_RANK", "0"))) % max(device_count(), 1)
    except ValueError:
        return 0


# Seeds for calls that do not pass `seed=`: the reference draws from a process-global stream that is
# never seeded (identical every process start, advancing from call to call); mirror that behaviour.
_seed_state = [0x243F6A8885A308D3]


def next_seed(seed=None):
    if seed is not None:
        return int(seed) & (2 ** 64 - 1)
    _seed_state[0] = (_seed_state[0] * 6364136223846793005 + 1442695040888963407) & (2 ** 64 - 1)
    return _seed_state[0]


def reseed(seed):
    """Reset the package-global seed sequence (analogue of srand())."""
    _seed_state[0] = int(seed) & (2 ** 64 - 1)
