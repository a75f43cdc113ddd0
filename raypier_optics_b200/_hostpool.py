"""Page-locked host buffers for the arrays the drop-in hands back (every generation of
``trace_rays``, capture-plane output, downloaded collections).

Why: a generation of 1e6 gausslets is 668 MB.  Into a fresh ``numpy.empty`` the D2H copy pays twice --
every 4 KB page of the new mapping faults in on first touch (measured: ~1-2 GB/s on one thread) and a
pageable destination makes the driver bounce the copy through its own staging buffer.  Into page-locked
memory the same copy runs at PCIe speed (~55 GB/s).  Page-locking is itself as slow as a first touch, so
the blocks are POOLED: the numpy array that wraps a block keeps a small owner object alive, and when the
last view of the array dies the block goes back to the free list for the next trace (an interactive
model re-traces the same scene over and over: ``raypier/tracer.py`` of the reference calls
``trace_rays`` on every trait change).

This is memory management, not a compute path: when page-locking is refused (ulimit, cgroup) or the
caps below are reached the arrays are plain ``numpy.empty`` and everything else is unchanged.

Environment:
  RPX_PINNED_RESULTS=0     plain numpy arrays always
  RPX_PINNED_POOL_GB       idle blocks kept for reuse (default: min(8 GB, 5 % of MemTotal))
  RPX_PINNED_MAX_GB        page-locked bytes handed out + idle (default: 25 % of MemTotal)
"""
import ctypes as C
import os
import threading

import numpy as np

MIN_BYTES = 4 << 20  # smaller results are not worth a page-locked block


def _mem_total_bytes():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemTotal:"):
                return int(line.split()[1]) * 1024
    except Exception:
        pass
    return 64 << 30


def _size_class(nbytes):
    """Round up to 1/8-octave steps so that nearly equal generations share blocks (<= 12.5 % slack)."""
    step = 1 << max(int(nbytes - 1).bit_length() - 4, 12)
    return (nbytes + step - 1) // step * step


class _Block(object):
    """Owner of one page-locked block; dies with the last numpy view of it."""
    __slots__ = ("pool", "ptr", "cap")

    def __init__(self, pool, ptr, cap):
        self.pool, self.ptr, self.cap = pool, ptr, cap

    def __del__(self):
        try:
            self.pool._release(self.ptr, self.cap)
        except Exception:  # interpreter shutdown: the process's page-locked memory goes with it
            pass


class HostPool(object):
    def __init__(self, lib):
        self._L = lib
        # re-entrant: a garbage-collection pass that starts inside _take (any allocation can trigger one) may finalise a
        # _Block of an unreachable cycle on the same thread, which calls _release; it only appends to _free
        self._lock = threading.RLock()
        self._free = []  # (cap, ptr), idle blocks
        self._idle = 0   # bytes in _free
        self._out = 0    # bytes handed out
        total = _mem_total_bytes()
        gb = float(1 << 30)
        self.enabled = os.environ.get("RPX_PINNED_RESULTS", "1") != "0"
        self.idle_cap = int(float(os.environ.get("RPX_PINNED_POOL_GB", min(8.0, 0.05 * total / gb))) * gb)
        self.max_bytes = int(float(os.environ.get("RPX_PINNED_MAX_GB", 0.25 * total / gb)) * gb)
        self.hits = self.misses = self.fallbacks = 0

    # -- internals ----------------------------------------------------------------
    def _take(self, nbytes):
        cap = _size_class(nbytes)
        with self._lock:
            best = -1
            for k, (c, _p) in enumerate(self._free):
                if c >= nbytes and c <= cap + (cap >> 1) and (best < 0 or c < self._free[best][0]):
                    best = k
            if best >= 0:
                c, p = self._free.pop(best)
                self._idle -= c
                self._out += c
                self.hits += 1
                return p, c
            if self._out + self._idle + cap > self.max_bytes:
                # make room from the idle blocks first
                while self._free and self._out + self._idle + cap > self.max_bytes:
                    c, p = self._free.pop()
                    self._idle -= c
                    self._L.rpx_host_free(p)
                if self._out + cap > self.max_bytes:
                    self.fallbacks += 1
                    return None, 0
            self._out += cap  # reserved before the (slow) allocation outside the lock
        p = self._L.rpx_host_alloc(cap)
        if not p:
            with self._lock:
                self._out -= cap
                self.fallbacks += 1
            return None, 0
        self.misses += 1
        return p, cap

    def _release(self, ptr, cap):
        with self._lock:
            self._out -= cap
            if self._idle + cap <= self.idle_cap:
                self._free.append((cap, ptr))
                self._idle += cap
                return
        self._L.rpx_host_free(ptr)

    # -- API ----------------------------------------------------------------------
    def empty(self, n, dtype):
        """``numpy.empty(n, dtype)``, on a pooled page-locked block when that pays."""
        dtype = np.dtype(dtype)
        n = int(n)
        nbytes = n * dtype.itemsize
        if not self.enabled or nbytes < MIN_BYTES:
            return np.empty(n, dtype=dtype)
        ptr, cap = self._take(nbytes)
        if not ptr:
            return np.empty(n, dtype=dtype)
        buf = (C.c_char * nbytes).from_address(ptr)
        buf._rpx_owner = _Block(self, ptr, cap)  # numpy keeps `buf` as the base of every view
        return np.frombuffer(buf, dtype=dtype, count=n)

    def trim(self):
        """Give the idle blocks back to the OS."""
        with self._lock:
            free, self._free, self._idle = self._free, [], 0
        for _c, p in free:
            self._L.rpx_host_free(p)

    def stats(self):
        with self._lock:
            return {"idle_bytes": self._idle, "out_bytes": self._out, "hits": self.hits, "misses": self.misses,
                    "fallbacks": self.fallbacks}


_pool = None
_pool_lock = threading.Lock()


def get_pool(lib=None):
    """The process-wide pool (page-locked memory is not tied to a device context).  ``lib``: the loaded librpx
    (``_lib.load()``); None loads it on first use."""
    global _pool
    with _pool_lock:
        if _pool is None:
            if lib is None:
                from ._lib import load
                lib = load()
            _pool = HostPool(lib)
        return _pool
