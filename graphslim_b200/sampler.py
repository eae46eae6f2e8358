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


def draw_class_batches(lib, members, batch, out_dtype, cache):
    """``np.random.permutation(members of class c)[:batch]`` for every class in order (dataset/loader.py:222) on numpy's
    global legacy generator -- drawn by ``gs_np_legacy_class_batches`` (csrc/host_sampler.cpp: the frozen legacy stream
    restated, bit-exact incl. the generator position; tests/test_sampler_draws.py), because numpy's own shuffle keeps the
    interpreter lock busy for ~2 ms per outer step at the Reddit shape and the thread issuing the kernels then waits for
    it.  Returns (concatenated batches, offsets) like the numpy formulation did."""
    if "cat" not in cache:
        off = np.zeros(len(members) + 1, dtype=np.int64)
        off[1:] = np.cumsum([m.size for m in members])
        cache["off"] = off
        cache["cat"] = (np.ascontiguousarray(np.concatenate(members), dtype=np.int64) if off[-1]
                        else np.zeros(1, dtype=np.int64))
        cache["cap"] = int(sum(min(int(m.size), int(batch)) for m in members))
    state = np.random.get_state()
    if state[0] != "MT19937":
        raise RuntimeError("numpy's global generator is not the legacy MT19937 stream")
    key = np.ascontiguousarray(state[1], dtype=np.uint32).copy()
    pos = np.array([int(state[2])], dtype=np.int32)
    out = np.empty(max(cache["cap"], 1), dtype=np.int32)
    out_off = np.zeros(len(members) + 1, dtype=np.int32)
    rc = lib.gs_np_legacy_class_batches(key.ctypes.data, pos.ctypes.data, len(members), cache["cat"].ctypes.data,
                                        cache["off"].ctypes.data, int(batch), out.ctypes.data, out_off.ctypes.data)
    if rc != 0:
        raise _lib.GraphSlimLibraryError(f"gs_np_legacy_class_batches failed ({rc})")
    np.random.set_state((state[0], key, int(pos[0]), state[3], state[4]))
    return out[:out_off[-1]].astype(out_dtype, copy=False), out_off.astype(out_dtype, copy=False)


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


class _Unpack:
    """Tensor views over the packed bytes of one step (needs self.nh, self.n_class, self.align, self._group_cache)."""

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


class ClassSampler(_Unpack):
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
        return draw_class_batches(self.lib, self.members, self.batch, np.int64, self.__dict__.setdefault("_draw", {}))

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
        if self.t1.is_alive() or self.t2.is_alive():
            # a worker that is still running holds (and rewrites) numpy's / torch's global generator state: carrying on
            # would silently break the bit-exact stream contract
            raise RuntimeError("sampler prefetch worker did not finish within 60 s; the random streams are in an "
                               "undefined position")


# =================================================================================================================
# Device-side sampler (csrc/device_sampler.cu): same blocks, same random streams, no host sampling / H2D of blocks
# =================================================================================================================
class DeviceClassSampler(_Unpack):
    """Drop-in for ClassSampler on a CUDA device.  numpy's generator still draws the class batches on the host (a
    sequential Fisher-Yates per class, loader.py:222); torch's CPU generator state is checked out to the device at the
    start of an epoch, advanced there by the sampling kernels, and written back when the epoch ends, so both streams
    are consumed exactly as the reference consumes them."""

    RING = 4            # blocks of steps i (being consumed), i+1, i+2 (sampling ahead) + the slot being recycled
    DEPTH = 2           # steps sampled ahead of the consumer: collect() then never waits for the side stream

    def __init__(self, adj_csr, members, dataset, nlayers, device, labels=None, batch=256, align=64):
        """adj_csr: ops.Csr of the normalised graph in HBM (int32 rowptr/col, fp32 val); labels: int32 device tensor."""
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.GraphSlimLibraryError("DeviceClassSampler needs a CUDA device")
        self.adj = adj_csr
        self.n = int(adj_csr.n_rows)
        self.members = [np.ascontiguousarray(m) for m in members]
        for m in self.members:
            if m.size and (int(m.min()) < 0 or int(m.max()) >= self.n):
                raise ValueError("class member ids out of range")
        self.n_class = len(members)
        self.fan = np.asarray(fanouts(dataset, nlayers), dtype=np.int32)
        self.nh = len(self.fan)
        self.batch = int(batch)
        self.align = int(align)
        self.labels = None if labels is None else labels.to(self.device, torch.int32).contiguous()
        self.max_batch = int(min(batch, max(len(m) for m in self.members)))
        # high priority: the sampler's short serial kernels are scheduled ahead of the main stream's wide ones
        self.side = torch.cuda.Stream(self.device, priority=-1)
        with torch.cuda.device(self.device):
            self.handle = self.lib.gs_dsampler_create(
                self.n, adj_csr.rowptr.data_ptr(), adj_csr.col.data_ptr(), adj_csr.val.data_ptr(),
                0 if self.labels is None else self.labels.data_ptr(), self.nh, self.fan.ctypes.data, self.n_class,
                self.max_batch, self.align, self.side.cuda_stream)
        if not self.handle:
            raise _lib.GraphSlimLibraryError("gs_dsampler_create failed: " +
                                             self.lib.gs_last_error().decode("utf-8", "replace"))
        self.cap = int(self.lib.gs_dsampler_out_capacity(self.handle))
        R = self.RING
        self.ring = [torch.empty(self.cap, dtype=torch.uint8, device=self.device) for _ in range(R)]
        self.desc_dev = [torch.empty(64, dtype=torch.int64, device=self.device) for _ in range(R)]
        self.desc_host = [torch.empty(64, dtype=torch.int64).pin_memory() for _ in range(R)]
        nb = self.n_class * self.max_batch + self.n_class + 1
        self.batch_host = [torch.empty(nb, dtype=torch.int32).pin_memory() for _ in range(R)]
        self.batch_dev = [torch.empty(nb, dtype=torch.int32, device=self.device) for _ in range(R)]
        self.done = [None] * R                 # side-stream event: slot's blocks and desc are ready
        self.consumed = [None] * R             # main-stream event: slot's previous blocks are no longer needed
        self.slot = -1
        self._mat_cache = {}
        self._group_cache = {}
        self.bytes_moved = 0
        self.stats = dict(steps=0, draw_batches_ms=0.0, launch_ms=0.0, wait_ready_ms=0.0)

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                torch.cuda.synchronize(self.device)
                self.lib.gs_dsampler_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def set_labels(self, labels):
        raise RuntimeError("pass labels to the constructor (the device sampler binds them at creation)")

    # ---- random streams --------------------------------------------------------------------------
    def draw_batches(self):
        """np.random.permutation(class members)[:256] per class, in class order (loader.py:222)."""
        return draw_class_batches(self.lib, self.members, self.batch, np.int32, self.__dict__.setdefault("_draw", {}))

    def checkout_rng(self):
        """torch's CPU generator -> device.  Nothing may draw from it until `checkin_rng`."""
        with _TorchMt() as g:
            _lib.check(self.lib.gs_dsampler_set_rng(self.handle, g.state.ctypes.data, int(g.left[0]), int(g.next[0]),
                                                    self.side.cuda_stream), "gs_dsampler_set_rng")

    def checkin_rng(self):
        with _TorchMt() as g:
            _lib.check(self.lib.gs_dsampler_get_rng(self.handle, g.state.ctypes.data, g.left.ctypes.data,
                                                    g.next.ctypes.data, self.side.cuda_stream), "gs_dsampler_get_rng")

    # ---- one step ----------------------------------------------------------------------------------
    def _materialise_dev(self, materialise):
        if materialise is None:
            return None
        key = tuple(int(m) for m in materialise)
        t = self._mat_cache.get(key)
        if t is None:
            t = torch.tensor(key, dtype=torch.uint8, device=self.device)
            self._mat_cache[key] = t
        return t

    def launch(self, batch, off, materialise=None):
        """Queues the sampling of one step on the side stream; returns the ring slot."""
        t0 = time.perf_counter()
        self.slot = slot = (self.slot + 1) % self.RING
        if self.done[slot] is not None:
            self.done[slot].synchronize()             # the pinned staging of this slot is free again
        nb = batch.size
        hb = self.batch_host[slot]
        hb[:nb] = torch.from_numpy(batch)
        hb[nb:nb + off.size] = torch.from_numpy(off)
        mat = self._materialise_dev(materialise)
        with torch.cuda.stream(self.side):
            if self.consumed[slot] is not None:
                self.side.wait_event(self.consumed[slot])
            db = self.batch_dev[slot]
            db[:nb + off.size].copy_(hb[:nb + off.size], non_blocking=True)
            _lib.check(self.lib.gs_dsampler_sample_step(
                self.handle, self.n_class, db.data_ptr(), db.data_ptr() + 4 * nb, 0 if mat is None else mat.data_ptr(),
                self.max_batch, self.ring[slot].data_ptr(), self.cap, self.desc_dev[slot].data_ptr(),
                self.side.cuda_stream), "gs_dsampler_sample_step")
            self.desc_host[slot].copy_(self.desc_dev[slot], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.side)
        self.done[slot] = ev
        self.bytes_moved += 4 * (nb + off.size) + 512
        self.stats["launch_ms"] += (time.perf_counter() - t0) * 1e3
        return slot

    def collect(self, slot, materialise=None):
        """Waits for the slot's step, makes the current stream depend on it, returns the RealBatch of device views."""
        t0 = time.perf_counter()
        self.done[slot].synchronize()
        desc = self.desc_host[slot].numpy().copy()
        if desc[63] != 0:
            raise _lib.GraphSlimLibraryError(f"device sampler: packed output does not fit ({int(desc[63])})")
        torch.cuda.current_stream(self.device).wait_event(self.done[slot])
        rb = self.unpack(self.ring[slot], desc, materialise)
        rb.h2d_bytes = 0
        rb.draws = int(desc[60])
        self.stats["steps"] += 1
        self.stats["wait_ready_ms"] += (time.perf_counter() - t0) * 1e3
        return rb

    def release(self, slot):
        """Call once every consumer of the slot's views has been queued on the current stream."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.consumed[slot] = ev

    def sample(self, materialise=None):
        """Synchronous step (tests, non-prefetched runs): both random streams advance exactly once."""
        batch, off = self.draw_batches()
        self.checkout_rng()
        slot = self.launch(batch, off, materialise)
        rb = self.collect(slot, materialise)
        self.checkin_rng()
        return rb

    def prefetch(self, n_steps, materialise=None):
        return _DevicePrefetcher(self, n_steps, materialise)


class _DevicePrefetcher:
    """Sampling of steps i+1 .. i+DEPTH runs on the side stream while step i is consumed.  A worker thread draws the class
    batches (numpy releases the GIL inside the shuffle); the kernels are queued from the consumer's thread."""

    def __init__(self, sampler, n_steps, materialise):
        self.s, self.mat, self.n = sampler, materialise, n_steps
        self.q = queue.Queue(maxsize=2)
        self.stop = False
        self.t = threading.Thread(target=self._draw, daemon=True)
        self.t.start()
        self.s.checkout_rng()
        self.i = 0
        self.prev_slot = None
        self.slots = {}
        for k in range(min(self.s.DEPTH, n_steps)):
            self._launch(k)

    def _draw(self):
        try:
            for _ in range(self.n):
                t0 = time.perf_counter()
                item = self.s.draw_batches()
                self.s.stats["draw_batches_ms"] += (time.perf_counter() - t0) * 1e3
                while not self.stop:
                    try:
                        self.q.put(item, timeout=0.05)
                        break
                    except queue.Full:
                        continue
                if self.stop:
                    return
        except BaseException as exc:
            self.q.put(exc)

    def _launch(self, i):
        item = self.q.get()
        if isinstance(item, BaseException):
            raise item
        self.slots[i] = self.s.launch(item[0], item[1], self.mat)

    def next(self):
        i = self.i
        if self.prev_slot is not None:
            self.s.release(self.prev_slot)       # everything that read step i-1's blocks is queued by now
        if i + self.s.DEPTH < self.n:
            self._launch(i + self.s.DEPTH)
        slot = self.slots.pop(i)
        rb = self.s.collect(slot, self.mat)
        self.prev_slot = slot
        self.i += 1
        return rb

    def join(self):
        if self.i < self.n:
            self.stop = True                     # consumer failed mid-epoch: streams stay where they are
        self.t.join(timeout=60)
        self.stop = True
        if self.t.is_alive():
            # the draw thread still owns the generators: checking the device stream back in over it would race
            raise RuntimeError("device sampler draw thread did not finish within 60 s; the random streams are in an "
                               "undefined position")
        if self.prev_slot is not None:
            self.s.release(self.prev_slot)
            self.prev_slot = None
        self.s.checkin_rng()
