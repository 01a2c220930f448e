"""Argument handling shared by the drop-in modules: the reference's buffer rules,
made batch-aware.  State arrays are mutated in place and the functions return None."""
import ctypes

import numpy as np

from . import _lib

_INT_KINDS = "iu"


def check_nbs(nbs, ndim=3):
    if isinstance(nbs, _lib.Instance):
        return nbs
    nbs = np.asarray(nbs)
    if nbs.dtype != np.float64:
        raise ValueError("Buffer dtype mismatch, expected 'float64_t' but got '%s'" % nbs.dtype)
    if nbs.ndim != ndim:
        raise ValueError("Buffer has wrong number of dimensions (expected %d, got %d)" % (ndim, nbs.ndim))
    if nbs.shape[-1] != 2:
        raise ValueError("nbs must end in (neighbour index, coupling) pairs")
    return nbs


def spins_in(arr, single_ndim, name):
    """(int8 C-contiguous batch [R, ...], batched?) from a reference-style spin array.

    The reference insists on C `long`; this accepts any integer dtype and any
    strides (the example passes a Fortran-ordered view, santoro80.py:286)."""
    if not isinstance(arr, np.ndarray):
        raise TypeError("Argument '%s' has incorrect type (expected numpy.ndarray, got %s)" % (name, type(arr).__name__))
    if arr.dtype.kind not in _INT_KINDS:
        raise ValueError("Buffer dtype mismatch, expected an integer spin array but got '%s'" % arr.dtype)
    if arr.ndim not in (single_ndim, single_ndim + 1):
        raise ValueError("Buffer has wrong number of dimensions (expected %d, got %d)" % (single_ndim, arr.ndim))
    batched = arr.ndim == single_ndim + 1
    a8 = np.ascontiguousarray(arr if batched else arr[None], dtype=np.int8)
    if a8 is arr or (batched and np.shares_memory(a8, arr)):
        return a8, batched, False  # zero-copy: already int8 contiguous, results land in place
    return a8, batched, True


def spins_out(arr, a8, batched, need_copy):
    if need_copy:
        arr[...] = a8 if batched else a8[0]


def angles_in(arr, single_ndim, name):
    if not isinstance(arr, np.ndarray):
        raise TypeError("Argument '%s' has incorrect type (expected numpy.ndarray, got %s)" % (name, type(arr).__name__))
    if arr.dtype != np.float64:
        raise ValueError("Buffer dtype mismatch, expected 'float64_t' but got '%s'" % arr.dtype)
    if arr.ndim != single_ndim:
        raise ValueError("Buffer has wrong number of dimensions (expected %d, got %d)" % (single_ndim, arr.ndim))
    a = np.ascontiguousarray(arr if arr.ndim == 2 else arr[None])
    return a, (a is not arr and not np.shares_memory(a, arr))


def seeds_u32(libc_seed, R):
    """Per-replica srand() seeds for the exact kernels: scalar s -> s, s+1, ... ; array -> as given."""
    if libc_seed is None:
        raise ValueError("exact=True replays the reference's libc rand() stream: pass libc_seed= (the value the "
                         "reference run would give to srand() right before the call)")
    s = np.asarray(libc_seed)
    if s.ndim == 0:
        s = (int(s) + np.arange(R, dtype=np.int64))
    s = np.ascontiguousarray(s.astype(np.int64) & 0xFFFFFFFF, dtype=np.uint32)
    if s.shape != (R,):
        raise ValueError("libc_seed must be a scalar or one seed per replica")
    return s


def u32p(a):
    return a.ctypes.data_as(_lib.c_u32p)
