"""thread pool over the maps of a batch (ctypes releases the GIL inside liboracle.so; maps are independent)"""
import os
from concurrent.futures import ThreadPoolExecutor

_pool = None


def pool():
    global _pool
    if _pool is None:
        _pool = ThreadPoolExecutor(max(1, int(os.environ.get("GENRE_ORACLE_THREADS", len(os.sched_getaffinity(0))))))
    return _pool


def per_map(fn, n):
    return list(pool().map(fn, range(n)))
