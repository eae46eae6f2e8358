"""The std::unordered_set<int64_t> iteration-order restatement used by the device sampler (csrc/uset_emul.h) against
the real libstdc++ container (oracle/csrc/oracle_host.cpp) -- CPU only, no compute kernels involved."""
import ctypes

import numpy as np

from graphslim_b200 import _lib
from oracle import hostlib


def _orders(keys):
    keys = np.ascontiguousarray(keys, dtype=np.int64)
    a, b = np.empty(32, np.int64), np.empty(32, np.int64)
    i64p = ctypes.POINTER(ctypes.c_int64)
    na = hostlib.lib().oracle_uset_order(keys.ctypes.data_as(i64p), keys.size, a.ctypes.data_as(i64p))
    nb = _lib.load().gs_uset_emul_order(keys.ctypes.data, keys.size, b.ctypes.data)
    return a[:na].tolist(), b[:nb].tolist()


def test_sequential_keys_all_sizes():
    for n in range(0, 16):                      # the `deg <= fanout` branch inserts 0..deg-1
        real, emu = _orders(np.arange(n))
        assert real == emu, (n, real, emu)


def test_random_floyd_like_sequences():
    rng = np.random.default_rng(0)
    for trial in range(20000):
        k = int(rng.integers(1, 16))
        deg = int(rng.integers(k + 1, 5000 if trial % 3 else 40))
        keys = []
        for j in range(deg - k, deg):           # Robert-Floyd: insert r, or j when r is already present
            r = int(rng.integers(0, j))
            keys.append(j if r in keys else r)
        real, emu = _orders(keys)
        assert real == emu, (keys, real, emu)


def test_duplicates_and_collisions():
    rng = np.random.default_rng(1)
    for _ in range(5000):
        n = int(rng.integers(1, 16))
        base = rng.integers(0, 60, n) * int(rng.choice([1, 13, 29]))     # many same-bucket keys
        keys = list(dict.fromkeys(base.tolist()))[:15]
        keys = keys + keys[:2]                                           # repeated inserts are ignored
        real, emu = _orders(keys)
        assert real == emu, (keys, real, emu)
