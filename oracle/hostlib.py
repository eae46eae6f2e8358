"""ORACLE (test infrastructure, not product code).

ctypes loader for ``oracle/csrc/oracle_host.cpp`` plus the glue that keeps the
C++ mt19937 restatement in lock-step with torch's default CPU generator
(``torch.get_rng_state`` / ``torch.set_rng_state``).

Reference behaviour restated: torch_sparse ``sample_adj`` consumes
``torch::randint(0, j, {1})`` per Floyd draw, i.e. one mt19937 word of the
*global* torch CPU generator (call sites: graphslim/dataset/loader.py:216-223).
"""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "csrc", "oracle_host.cpp")
_OUT_DIR = os.path.join(_HERE, "_build")
_OUT = os.path.join(_OUT_DIR, "liboracle_host.so")

_lib = None

# torch CPUGeneratorImpl serialised state (5056 bytes):
#   u64 seed | i32 left | i32 seeded | u64 next | u64 state[624] | 3*f64 | i32 | pad | f32 | bool | pad
_OFF_LEFT, _OFF_NEXT, _OFF_STATE = 8, 16, 24


def build(force=False):
    if not force and os.path.exists(_OUT) and os.path.getmtime(_OUT) >= os.path.getmtime(_SRC):
        return _OUT
    os.makedirs(_OUT_DIR, exist_ok=True)
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", _OUT, _SRC]
    subprocess.check_call(cmd)
    return _OUT


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_OUT):
            build()
        L = ctypes.CDLL(_OUT)
        i64p = ctypes.POINTER(ctypes.c_int64)
        u32p = ctypes.POINTER(ctypes.c_uint32)
        i32p = ctypes.POINTER(ctypes.c_int32)
        f32p = ctypes.POINTER(ctypes.c_float)
        L.oracle_mt_randint.argtypes = [u32p, i32p, i32p, i64p, ctypes.c_int64, i64p]
        L.oracle_mt_randint.restype = None
        L.oracle_sample_adj.argtypes = [i64p, i64p, i64p, ctypes.c_int64, ctypes.c_int64, u32p, i32p, i32p,
                                        i64p, i64p, i64p, i64p]
        L.oracle_sample_adj.restype = ctypes.c_int64
        L.oracle_spmm_csr_f32.argtypes = [ctypes.c_int64, i64p, i64p, f32p, f32p, ctypes.c_int64, f32p]
        L.oracle_spmm_csr_f32.restype = None
        L.oracle_uset_order.argtypes = [i64p, ctypes.c_int64, i64p]
        L.oracle_uset_order.restype = ctypes.c_int64
        _lib = L
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


class TorchMt:
    """Checks out torch's default CPU generator state for C++ use and writes it back."""

    def __enter__(self):
        raw = torch.get_rng_state().numpy().copy()
        self.raw = raw
        self.left = np.frombuffer(raw[_OFF_LEFT:_OFF_LEFT + 4].tobytes(), dtype=np.int32).copy()
        self.next = np.frombuffer(raw[_OFF_NEXT:_OFF_NEXT + 8].tobytes(), dtype=np.uint64).astype(np.int32)
        st64 = np.frombuffer(raw[_OFF_STATE:_OFF_STATE + 624 * 8].tobytes(), dtype=np.uint64)
        self.state = st64.astype(np.uint32)
        return self

    def __exit__(self, *exc):
        raw = self.raw
        raw[_OFF_LEFT:_OFF_LEFT + 4] = np.frombuffer(self.left.astype(np.int32).tobytes(), dtype=np.uint8)
        raw[_OFF_NEXT:_OFF_NEXT + 8] = np.frombuffer(self.next.astype(np.uint64).tobytes(), dtype=np.uint8)
        raw[_OFF_STATE:_OFF_STATE + 624 * 8] = np.frombuffer(self.state.astype(np.uint64).tobytes(), dtype=np.uint8)
        torch.set_rng_state(torch.from_numpy(raw))
        return False


def mt_randint(high):
    """``[torch.randint(0, h, (1,)) for h in high]`` through the C++ restatement (advances torch's RNG)."""
    high = np.ascontiguousarray(high, dtype=np.int64)
    out = np.empty_like(high)
    with TorchMt() as g:
        lib().oracle_mt_randint(_p(g.state, ctypes.c_uint32), _p(g.left, ctypes.c_int32), _p(g.next, ctypes.c_int32),
                                _p(high, ctypes.c_int64), high.size, _p(out, ctypes.c_int64))
    return out


def sample_adj(rowptr, col, idx, k):
    """One hop of torch_sparse ``sample_adj(rowptr, col, idx, k, replace=False)``.

    Returns (out_rowptr, out_col, n_id, e_id) as int64 numpy arrays.
    """
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    col = np.ascontiguousarray(col, dtype=np.int64)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    n = idx.size
    assert k >= 0, "oracle restates the replace=False, k>=0 path only"
    out_rowptr = np.empty(n + 1, dtype=np.int64)
    out_col = np.empty(max(n * k, 1), dtype=np.int64)
    out_eid = np.empty(max(n * k, 1), dtype=np.int64)
    out_nid = np.empty(n + n * k, dtype=np.int64)
    with TorchMt() as g:
        nn = lib().oracle_sample_adj(_p(rowptr, ctypes.c_int64), _p(col, ctypes.c_int64), _p(idx, ctypes.c_int64),
                                     n, k, _p(g.state, ctypes.c_uint32), _p(g.left, ctypes.c_int32),
                                     _p(g.next, ctypes.c_int32), _p(out_rowptr, ctypes.c_int64),
                                     _p(out_col, ctypes.c_int64), _p(out_eid, ctypes.c_int64),
                                     _p(out_nid, ctypes.c_int64))
    e = int(out_rowptr[n])
    return out_rowptr, out_col[:e].copy(), out_nid[:nn].copy(), out_eid[:e].copy()


def spmm_csr(rowptr, col, val, x):
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    col = np.ascontiguousarray(col, dtype=np.int64)
    val = np.ascontiguousarray(val, dtype=np.float32)
    x = np.ascontiguousarray(x, dtype=np.float32)
    n = rowptr.size - 1
    y = np.empty((n, x.shape[1]), dtype=np.float32)
    lib().oracle_spmm_csr_f32(n, _p(rowptr, ctypes.c_int64), _p(col, ctypes.c_int64), _p(val, ctypes.c_float),
                              _p(x, ctypes.c_float), x.shape[1], _p(y, ctypes.c_float))
    return y
