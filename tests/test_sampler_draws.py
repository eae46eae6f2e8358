"""Class batches drawn by the C restatement of numpy's legacy shuffle (gs_np_legacy_class_batches) are bit-identical to
``np.random.permutation(members)[:batch]`` per class in order (graphslim/dataset/loader.py:222) and leave numpy's global
generator at the same position -- for ragged, empty and single-member classes and across the 624-word block boundary."""
import threading

import numpy as np
import pytest

from graphslim_b200 import _lib
from graphslim_b200.sampler import draw_class_batches


def numpy_batches(members, batch):
    parts = [np.random.permutation(m)[:batch].astype(np.int64) for m in members]
    off = np.zeros(len(members) + 1, dtype=np.int64)
    off[1:] = np.cumsum([p.size for p in parts])
    return np.concatenate(parts) if parts else np.zeros(0, np.int64), off


@pytest.mark.parametrize("trial", range(12))
def test_native_draws_equal_numpy(trial):
    lib = _lib.load()
    rng = np.random.default_rng(trial)
    n_class = int(rng.integers(1, 12))
    sizes = [int(rng.integers(0, 5000)) if trial % 4 else int(rng.integers(0, 3)) for _ in range(n_class)]
    members = [np.sort(rng.choice(200000, s, replace=False)).astype(np.int64) for s in sizes]
    np.random.seed(int(rng.integers(0, 2 ** 31)))
    np.random.randint(0, 10, size=int(rng.integers(0, 2000)))         # a random position, block boundary included
    start = np.random.get_state()
    cache = {}
    for step in range(3):                                              # consecutive steps keep the stream aligned
        state = np.random.get_state()
        ref, ref_off = numpy_batches(members, 256)
        probe_ref = np.random.randint(0, 2 ** 31 - 1, size=4)
        np.random.set_state(state)
        got, got_off = draw_class_batches(lib, members, 256, np.int64, cache)
        probe_got = np.random.randint(0, 2 ** 31 - 1, size=4)
        assert np.array_equal(got, ref) and np.array_equal(got_off, ref_off)
        assert np.array_equal(probe_got, probe_ref)
    assert start[2] <= 624


def test_native_draws_from_a_worker_thread_and_int32():
    lib = _lib.load()
    members = [np.arange(1000, dtype=np.int64) + 1000 * c for c in range(7)]
    np.random.seed(3)
    ref, ref_off = numpy_batches(members, 256)
    np.random.seed(3)
    box = {}
    t = threading.Thread(target=lambda: box.update(r=draw_class_batches(lib, members, 256, np.int32, {})))
    t.start()
    t.join()
    got, got_off = box["r"]
    assert got.dtype == np.int32 and np.array_equal(got, ref) and np.array_equal(got_off, ref_off)


def test_rejects_a_foreign_bit_generator(monkeypatch):
    lib = _lib.load()
    monkeypatch.setattr(np.random, "get_state", lambda: ("PCG64", None, 0, 0, 0.0))
    with pytest.raises(RuntimeError):
        draw_class_batches(lib, [np.arange(5)], 2, np.int64, {})
