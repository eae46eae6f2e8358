"""Class-batch neighbour sampling for one outer step (host C++ in csrc/host_sampler.cpp).

Replaces ``TransAndInd.retrieve_class_sampler`` (graphslim/dataset/loader.py:187-224) for all classes
at once.  The class batch is ``np.random.permutation(members)[:256]`` (numpy's global legacy generator,
as in the reference); neighbour draws come from torch's default CPU generator whose mt19937 state is
checked out, advanced by the C++ sampler and written back, so both streams interleave exactly like the
reference's.  Output arrays are packed into one pinned buffer and moved to HBM with a single copy.
"""
import ctypes
import queue
import threading
import time

import numpy as np
import torch

from . import _lib
from .ops import Csr

_OFF_LEFT, _OFF_NEXT, _OFF_STATE = 8, 16, 24      # torch CPUGeneratorImpl serialised layout


def fanouts(dataset, nlayers):
    """loader.py:197-210."""
    if nlayers == 1:
        return [15]
    if nlayers == 2:
        return [15, 8] if dataset in ("reddit", "flickr") else [10, 5]
    if nlayers in (3, 4, 5):
        return {3: [15, 10, 5], 4: [15, 10, 5, 5], 5: [15, 10, 5, 5, 5]}[nlayers]
    raise ValueError(f"nlayers={nlayers} has no fan-out schedule in the reference")


class _TorchMt:
    def __enter__(self):
        raw = torch.get_rng_state().numpy().copy()
        self.raw = raw
        self.left = raw[_OFF_LEFT:_OFF_LEFT + 4].view(np.int32).copy()
        self.next = raw[_OFF_NEXT:_OFF_NEXT + 8].view(np.uint64).astype(np.int32)
        self.state = raw[_OFF_STATE:_OFF_STATE + 624 * 8].view(np.uint64).astype(np.uint32)
        return self

    def __exit__(self, *exc):
        raw = self.raw
        raw[_OFF_LEFT:_OFF_LEFT + 4] = self.left.view(np.uint8)
        raw[_OFF_NEXT:_OFF_NEXT + 8] = self.next.astype(np.uint64).view(np.uint8)
        raw[_OFF_STATE:_OFF_STATE + 624 * 8] = self.state.astype(np.uint64).view(np.uint8)
        torch.set_rng_state(torch.from_numpy(raw))
        return False


class Block:
    """One hop of the batched sampled structure: rows = level h nodes, cols = level h+1 nodes."""

    def __init__(self, csr, csr_t, gcol):
        self.csr, self.csr_t, self._gcol = csr, csr_t, gcol

    def with_global_cols(self):
        if self._gcol is None:
            raise ValueError("only the outermost hop carries global column ids")
        return Csr(self.csr.rowptr, self._gcol, self.csr.val, self.csr.n_rows, -1)


class RealBatch:
    pass


class ClassSampler:
    def __init__(self, adj_rowptr, adj_col, adj_val, members, dataset, nlayers, device, batch=256, align=64):
        """adj_*: CSR of the normalised graph on the host (int64 rowptr, int32 col, fp32 val);
        members[c]: node ids of class c (global ids in 'trans', train-local in 'ind')."""
        self.lib = _lib.load()
        self.rowptr = np.ascontiguousarray(adj_rowptr, dtype=np.int64)
        self.col = np.ascontiguousarray(adj_col, dtype=np.int32)
        self.val = np.ascontiguousarray(adj_val, dtype=np.float32)
        self.n = self.rowptr.size - 1
        self.members = [np.ascontiguousarray(m) for m in members]
        self.n_class = len(members)
        self.fan = np.asarray(fanouts(dataset, nlayers), dtype=np.int32)
        self.nh = len(self.fan)
        self.batch = batch
        self.device = torch.device(device)
        self.handle = self.lib.gs_sampler_create(self.n, self.rowptr.ctypes.data, self.col.ctypes.data,
                                                 self.val.ctypes.data, self.nh, self.fan.ctypes.data)
        if not self.handle:
            raise _lib.GraphSlimLibraryError("gs_sampler_create failed")
        self.align = int(align)
        self.lib.gs_sampler_set_align(self.handle, self.align)
        # worst-case packed size (every class segment padded to `align` rows at every level)
        pad = lambda x: (x + self.align - 1) // self.align * self.align
        rows = pad(min(batch, max(len(m) for m in self.members))) * self.n_class
        cap, lvl = 0, rows
        cap += 4 * (self.nh + 1) * (self.n_class + 1) + 64
        for k in self.fan:
            nnz = lvl * int(k)
            cap += 4 * (lvl + 1) + 7 * 4 * nnz + 4 * (lvl + nnz + self.align * self.n_class + 1) + 8 * 16
            lvl = lvl + nnz + self.align * self.n_class
        cap += 4 * lvl + 4 * 4 * rows + 4 * (self.nh + 1) * self.n_class + 16 * 8
        self.cap = int(cap)
        # ring of pinned staging buffers: a buffer is reused only after its H2D copy has completed
        # (3 slots: one being filled by the prefetch worker, one queued, one being copied / consumed)
        self.ring = [torch.empty(self.cap, dtype=torch.uint8, pin_memory=self.device.type == "cuda")
                     for _ in range(3)]
        self.ring_events = [None] * len(self.ring)
        self.ring_pos = 0
        self.pinned = self.ring[0]
        self.desc = np.empty(64, dtype=np.int64)
        self.labels = None
        self.bytes_moved = 0              # host->device bytes of sampled blocks so far
        self.stats = dict(steps=0, draw_batches_ms=0.0, serial_hops_ms=0.0, parallel_last_hop_ms=0.0, pack_ms=0.0,
                          stage1_total_ms=0.0, stage2_total_ms=0.0)
        self._group_cache = {}

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.gs_sampler_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def set_labels(self, labels):
        """int32 label per graph node; the sampler then emits the labels of the target rows."""
        self.labels = np.ascontiguousarray(labels, dtype=np.int32)
        assert self.labels.size == self.n
        self.lib.gs_sampler_set_labels(self.handle, self.labels.ctypes.data)

    def draw_batches(self):
        """np.random.permutation(class members)[:256] per class, in class order (loader.py:222)."""
        parts = [np.random.permutation(m)[:self.batch].astype(np.int64) for m in self.members]
        off = np.zeros(self.n_class + 1, dtype=np.int64)
        off[1:] = np.cumsum([p.size for p in parts])
        return np.concatenate(parts), off

    def begin_host(self, materialise=None):
        """Stage 1 (owns numpy's and torch's global generators): class batches + the serial hops.  Returns a job."""
        t0 = time.perf_counter()
        batch, off = self.draw_batches()
        t_draw = time.perf_counter() - t0
        mat = None if materialise is None else np.ascontiguousarray(materialise, dtype=np.uint8)
        with _TorchMt() as g:
            job = self.lib.gs_sampler_begin_step(self.handle, self.n_class, batch.ctypes.data, off.ctypes.data,
                                                 mat.ctypes.data if mat is not None else None, g.state.ctypes.data,
                                                 g.left.ctypes.data, g.next.ctypes.data)
        if not job:
            raise _lib.GraphSlimLibraryError("gs_sampler_begin_step failed (class batch out of range?)")
        st = self.stats
        st["draw_batches_ms"] += t_draw * 1e3
        st["stage1_total_ms"] += (time.perf_counter() - t0) * 1e3
        return job

    def finish_host(self, job):
        """Stage 2: parallel last hop + packing into the next ring slot.  Returns (slot, bytes_used, desc copy).
        No CUDA work is issued here; the slot is reused only after its previous H2D copy has completed."""
        t0 = time.perf_counter()
        self.ring_pos = (self.ring_pos + 1) % len(self.ring)
        slot = self.ring_pos
        if self.ring_events[slot] is not None:
            self.ring_events[slot].synchronize()
        self.pinned = self.ring[slot]
        desc = np.empty(64, dtype=np.int64)
        used = self.lib.gs_sampler_finish_step(self.handle, job, self.pinned.data_ptr(), self.cap, desc.ctypes.data)
        if used < 0:
            raise _lib.GraphSlimLibraryError(f"gs_sampler_finish_step failed with code {used}")
        st = self.stats
        st["steps"] += 1
        st["serial_hops_ms"] += desc[40] / 1e3
        st["parallel_last_hop_ms"] += desc[41] / 1e3
        st["pack_ms"] += desc[42] / 1e3
        st["stage2_total_ms"] += (time.perf_counter() - t0) * 1e3
        return slot, int(used), desc

    def sample_host(self, materialise=None):
        """Both stages back to back; returns (bytes_used, desc).  The packed arrays are in self.pinned."""
        slot, used, desc = self.finish_host(self.begin_host(materialise))
        return used, desc

    # ---- views --------------------------------------------------------------------------------
    @staticmethod
    def _view(buf, off, count, dtype):
        nbytes = count * 4
        return buf[off:off + nbytes].view(dtype)

    def unpack(self, buf, desc, materialise=None):
        """Build a RealBatch of tensor views over `buf` (host or device copy of the packed bytes)."""
        nh, nc = self.nh, self.n_class
        rb = RealBatch()
        counts = [int(desc[2 + l]) for l in range(nh + 1)]
        segs = self._view(buf, int(desc[8]), (nh + 1) * (nc + 1), torch.int32).view(nh + 1, nc + 1)
        rb.counts = counts
        rb.nid = self._view(buf, int(desc[9]), counts[nh], torch.int32)
        rb.tcls = self._view(buf, int(desc[10]), counts[0], torch.int32)
        rb.inv_b = self._view(buf, int(desc[11]), counts[0], torch.float32)
        rb.target_ids = self._view(buf, int(desc[12]), counts[0], torch.int32)
        rb.labels = self._view(buf, int(desc[13]), counts[0], torch.int32) if desc[13] >= 0 else None
        rb.aligned = self.align % 64 == 0      # class segments start on tensor-core tile boundaries
        rb.cnt = self._view(buf, int(desc[14]), (nh + 1) * nc, torch.int32).view(nh + 1, nc)   # unpadded class sizes
        blocks = []
        for h in range(nh):
            d = desc[16 + 8 * h: 24 + 8 * h]
            nnz = int(d[0])
            rows, cols = counts[h], counts[h + 1]
            csr = Csr(self._view(buf, int(d[1]), rows + 1, torch.int32), self._view(buf, int(d[2]), nnz, torch.int32),
                      self._view(buf, int(d[3]), nnz, torch.float32), rows, cols)
            csr_t = Csr(self._view(buf, int(d[4]), cols + 1, torch.int32),
                        self._view(buf, int(d[5]), nnz, torch.int32),
                        self._view(buf, int(d[6]), nnz, torch.float32), cols, rows)
            gcol = self._view(buf, int(d[7]), nnz, torch.int32) if d[7] >= 0 else None
            blocks.append(Block(csr, csr_t, gcol))
        rb.blocks = blocks
        rb.blocks_fwd = blocks[::-1]          # application order: outermost hop first (adjs[::-1] in PyG)
        # groups = materialised classes; per-level row segments restricted to them
        key = (None if materialise is None else tuple(int(m) for m in materialise), str(buf.device))
        cached = self._group_cache.get(key)
        if cached is None:          # constant across steps: build once (a fresh torch.tensor(...) would sync the stream)
            keep = np.arange(nc) if materialise is None else np.nonzero(np.asarray(materialise))[0]
            cached = (torch.tensor(keep, dtype=torch.int32, device=buf.device),
                      torch.arange(len(keep), dtype=torch.int32, device=buf.device),
                      torch.tensor(np.concatenate([keep, [nc]]), dtype=torch.int64, device=buf.device),
                      len(keep) == nc)
            self._group_cache[key] = cached
        rb.class_ids, rb.out_block, idx, all_classes = cached      # out_block: class-column block per group
        if all_classes:
            rb.seg = [segs[l] for l in range(nh + 1)]
        else:
            # segments of skipped classes are empty, so dropping them keeps the offsets contiguous
            rb.seg = [segs[l][idx].contiguous() for l in range(nh + 1)]
        return rb

    def upload(self, slot, used, desc, materialise=None):
        """One H2D copy of the packed bytes of ring slot `slot`, then device views."""
        src = self.ring[slot]
        if self.device.type == "cuda":
            dev = torch.empty(used, dtype=torch.uint8, device=self.device)
            dev.copy_(src[:used], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self.ring_events[slot] = ev
        else:
            dev = src[:used].clone()
        rb = self.unpack(dev, desc, materialise)
        rb.h2d_bytes = used
        self.bytes_moved += used
        return rb

    def sample(self, materialise=None):
        """Full step: host sampling, one H2D copy of the packed bytes, device views."""
        used, desc = self.sample_host(materialise)
        return self.upload(self.ring_pos, used, desc, materialise)

    def prefetch(self, n_steps, materialise=None):
        """Samples the next `n_steps` outer steps on a worker thread (the C++ sampler releases the GIL), at most one
        step ahead of the consumer, so host sampling overlaps the GPU work of the previous step.  Both random streams
        are consumed in exactly the order a synchronous loop would consume them; the caller must not touch numpy's /
        torch's global generators until `join()`."""
        return _Prefetcher(self, n_steps, materialise)


class _Prefetcher:
    """Two worker threads: stage 1 (random streams, serial hops) feeds stage 2 (parallel last hop + packing); each is
    at most one step ahead of its consumer.  The C++ calls release the GIL."""

    def __init__(self, sampler, n_steps, materialise):
        self.s, self.mat = sampler, materialise
        self.q1 = queue.Queue(maxsize=1)
        self.q = queue.Queue(maxsize=1)
        self.stop = False
        self.t1 = threading.Thread(target=self._stage1, args=(n_steps,), daemon=True)
        self.t2 = threading.Thread(target=self._stage2, args=(n_steps,), daemon=True)
        self.t1.start()
        self.t2.start()

    def _put(self, q, item):
        while not self.stop:
            try:
                q.put(item, timeout=0.05)
                return True
            except queue.Full:
                continue
        return False

    def _get(self, q):
        while not self.stop:
            try:
                return q.get(timeout=0.05)
            except queue.Empty:
                continue
        return None

    def _stage1(self, n_steps):
        try:
            for _ in range(n_steps):
                if self.stop:
                    return
                self._put(self.q1, self.s.begin_host(self.mat))
        except BaseException as exc:
            self._put(self.q1, exc)

    def _stage2(self, n_steps):
        try:
            for _ in range(n_steps):
                job = self._get(self.q1)
                if job is None:
                    return
                if isinstance(job, BaseException):
                    raise job
                self._put(self.q, self.s.finish_host(job))
        except BaseException as exc:           # surfaced by next()
            self._put(self.q, exc)

    def next(self):
        item = self.q.get()
        if isinstance(item, BaseException):
            raise item
        slot, used, desc = item
        return self.s.upload(slot, used, desc, self.mat)

    def join(self):
        """Normal end of an epoch: every step has been consumed.  After an error in the consumer the workers are told
        to stop (the random streams are then left mid-epoch, as they would be after the same error in a serial loop;
        a job already begun is leaked rather than finished)."""
        if self.t1.is_alive() or self.t2.is_alive():
            if not (self.q.empty() and self.q1.empty()):
                self.stop = True
        self.t1.join(timeout=60)
        self.t2.join(timeout=60)
        self.stop = True
