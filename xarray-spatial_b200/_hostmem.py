"""Pinned (page-locked) host memory for the numpy-backed path.

cudaHostAlloc is slow (~0.3 s / GiB), so freed blocks are cached by size and recycled; a
numpy array handed to the user owns its block through a finalizer.
"""
import ctypes
import threading
import weakref

import numpy as np

from . import _lib

_lock = threading.Lock()
_free = {}          # nbytes -> [ptr, ...]
_cached_bytes = 0
MAX_CACHED_BYTES = 64 << 30


def _release(ptr, nbytes):
    global _cached_bytes
    with _lock:
        if _cached_bytes + nbytes <= MAX_CACHED_BYTES:
            _free.setdefault(nbytes, []).append(ptr)
            _cached_bytes += nbytes
            return
    _lib.lib().xrs_host_free(ctypes.c_void_p(ptr))


def empty(shape, dtype):
    """np.empty(shape, dtype) backed by pinned memory (falls back to pageable for 0 bytes)."""
    global _cached_bytes
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    if n == 0:
        return np.empty(shape, dtype)
    nbytes = (n + 4095) // 4096 * 4096
    ptr = None
    with _lock:
        lst = _free.get(nbytes)
        if lst:
            ptr = lst.pop()
            _cached_bytes -= nbytes
    if ptr is None:
        p = ctypes.c_void_p()
        _lib.call("xrs_host_alloc", ctypes.byref(p), nbytes)
        ptr = p.value
    buf = (ctypes.c_char * nbytes).from_address(ptr)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    weakref.finalize(buf, _release, ptr, nbytes)
    return arr


def pinned_copy(a):
    out = empty(a.shape, a.dtype)
    np.copyto(out, a)
    return out


def trim():
    """Free every cached block."""
    global _cached_bytes
    with _lock:
        blocks = [(p, n) for n, lst in _free.items() for p in lst]
        _free.clear()
        _cached_bytes = 0
    for p, _ in blocks:
        _lib.lib().xrs_host_free(ctypes.c_void_p(p))
