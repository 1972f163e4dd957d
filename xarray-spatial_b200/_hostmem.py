"""Pinned (page-locked) host memory for the numpy-backed path.

cudaHostAlloc is slow (~0.3 s / GiB), so freed blocks are cached by size and recycled; a
numpy array handed to the user owns its block through a finalizer.

The cache is bounded: page-locked memory is taken away from the rest of the system, so the idle
blocks are capped at XRS_B200_PINNED_CACHE_GB (default: a quarter of physical RAM, at most
16 GiB), the least recently released blocks are freed first when the cap is exceeded, and an
allocation that fails trims the whole cache and retries once.
"""
import collections
import ctypes
import os
import threading
import weakref

import numpy as np

from . import _lib


def _default_cap():
    env = os.environ.get("XRS_B200_PINNED_CACHE_GB")
    if env is not None:
        return max(0, int(float(env) * (1 << 30)))
    try:
        total = os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES")
    except (ValueError, OSError, AttributeError):
        total = 16 << 30
    return int(min(16 << 30, total // 4))


_lock = threading.Lock()
_free = collections.OrderedDict()   # (release order) id -> (ptr, nbytes); oldest first
_by_size = {}                        # nbytes -> [id, ...]
_cached_bytes = 0
_next_id = 0
MAX_CACHED_BYTES = _default_cap()


def _free_block(ptr):
    _lib.lib().xrs_host_free(ctypes.c_void_p(ptr))


def _release(ptr, nbytes):
    """Finalizer of a handed-out array: keep the block for reuse, evicting the least recently
    released blocks while the idle cache exceeds its cap."""
    global _cached_bytes, _next_id
    evict = []
    with _lock:
        if nbytes > MAX_CACHED_BYTES:
            evict.append(ptr)
        else:
            _next_id += 1
            _free[_next_id] = (ptr, nbytes)
            _by_size.setdefault(nbytes, []).append(_next_id)
            _cached_bytes += nbytes
            while _cached_bytes > MAX_CACHED_BYTES and _free:
                bid, (p, n) = _free.popitem(last=False)
                _by_size[n].remove(bid)
                _cached_bytes -= n
                evict.append(p)
    for p in evict:
        _free_block(p)


def _take(nbytes):
    global _cached_bytes
    with _lock:
        ids = _by_size.get(nbytes)
        if ids:
            bid = ids.pop()                 # most recently released block of that size
            ptr, n = _free.pop(bid)
            _cached_bytes -= n
            return ptr
    return None


def empty(shape, dtype):
    """np.empty(shape, dtype) backed by pinned memory (falls back to pageable for 0 bytes)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    if n == 0:
        return np.empty(shape, dtype)
    nbytes = (n + 4095) // 4096 * 4096
    ptr = _take(nbytes)
    if ptr is None:
        p = ctypes.c_void_p()
        try:
            _lib.call("xrs_host_alloc", ctypes.byref(p), nbytes)
        except Exception:
            trim()                          # give the idle blocks back and try once more
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


def cached_bytes():
    return _cached_bytes


def trim():
    """Free every cached block."""
    global _cached_bytes
    with _lock:
        blocks = [p for p, _ in _free.values()]
        _free.clear()
        _by_size.clear()
        _cached_bytes = 0
    for p in blocks:
        _free_block(p)
