"""Kernel namespace: thin torch-tensor wrappers over the C ABI (one C call per method).

PyTorch only supplies device memory and the stream here.  Every method launches hand-written sm_100a
kernels from libgraphslim_b200.so on the current CUDA stream; CPU tensors are rejected.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib


class Csr:
    """CSR matrix in HBM: int32 rowptr/col, fp32 val."""

    __slots__ = ("rowptr", "col", "val", "n_rows", "n_cols", "chunks")

    def __init__(self, rowptr, col, val, n_rows, n_cols, chunks=None):
        self.rowptr, self.col, self.val = rowptr, col, val
        self.n_rows, self.n_cols = int(n_rows), int(n_cols)
        # optional long-row splitting: (chunk_row, chunk_beg, chunk_end, long_thr) from graph_utils.build_row_chunks
        self.chunks = chunks


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _f32(t, name="tensor"):
    if t.dtype != torch.float32 or not t.is_cuda:
        raise TypeError(f"{name}: expected a CUDA float32 tensor, got {t.dtype} on {t.device}")
    return t


def _mat(t, name):
    # hot path of every wrapper (the host launch rate bounds the small-shape and the multi-GPU runs): one pass, no
    # helper calls
    if t.dtype is not torch.float32 or not t.is_cuda:
        raise TypeError(f"{name}: expected a CUDA float32 tensor, got {t.dtype} on {t.device}")
    shp = t.shape
    if len(shp) != 2:
        raise ValueError(f"{name}: need a 2-D matrix with unit column stride, got shape {tuple(shp)}")
    st = t.stride()
    cols = shp[1]
    if cols > 1 and st[1] != 1:
        raise ValueError(f"{name}: need a 2-D matrix with unit column stride, got shape {tuple(shp)} stride {st}")
    ld = st[0] if shp[0] > 1 else (st[0] if st[0] > cols else cols)
    return ld if ld > cols else cols


METRIC_ID = {"ours": 0, "mse": 1, "cos": 2}

# gs_spmm_set_tuning is process-global library state: the record of what is currently set lives next to the library
# handle (module level, lock-guarded), not per CudaOps instance -- one instance is created per reducer and the tests /
# bench create more, so a per-instance cache could disagree with the library about the active tuning.
import threading
_SPMM_TUNE_LOCK = threading.Lock()
_SPMM_TUNE = {"state": (0, 0)}


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
if _raw_stream is None:                      # older torch: the public (slower) route
    def _raw_stream(index):
        return torch.cuda.current_stream(index).cuda_stream


class _Timed:
    """``with K.timed(tag):`` brackets the enclosed launches with CUDA events when timing is switched on."""

    def __init__(self, K, tag):
        self.K, self.tag = K, tag

    def __enter__(self):
        self.on = getattr(self.K, "_timers", None) is not None
        if self.on:
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record(torch.cuda.current_stream(self.K.device))

    def __exit__(self, *exc):
        if self.on:
            self.b.record(torch.cuda.current_stream(self.K.device))
            self.K._timers.setdefault(self.tag, []).append((self.a, self.b))
        return False


class CudaOps:
    """All device work of the GCond path.  ``precision``: 0 exact fp32 SIMT, 1 tcgen05 3xBF16, 2 tcgen05 BF16."""

    def __init__(self, device, precision=0):
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.GraphSlimLibraryError("graphslim_b200 runs on CUDA devices only (no CPU fallback); "
                                             f"got device {device!r}")
        self.precision = int(precision)
        self._dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self._ws_bytes = {}

    # -- plumbing ------------------------------------------------------------------------------
    @property
    def stream(self):
        # raw handle of torch's current stream on this device (what torch.cuda.current_stream(dev).cuda_stream returns,
        # without building the Stream object: ~6x cheaper, and it is read once per launch)
        return _raw_stream(self._dev_index)

    def empty(self, *shape, dtype=torch.float32):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    def zeros(self, *shape, dtype=torch.float32):
        return torch.zeros(*shape, dtype=dtype, device=self.device)

    def launches(self):
        return int(self.lib.gs_launch_count())

    # -- per-kernel device timing (bench.py roofline): CUDA events on the launching stream ---------
    def start_timing(self):
        self._timers = {}

    def stop_timing(self):
        """Synchronises and returns {tag: (launches, total_ms)}."""
        timers, self._timers = getattr(self, "_timers", None) or {}, None
        torch.cuda.synchronize(self.device)
        return {tag: (len(ev), sum(a.elapsed_time(b) for a, b in ev)) for tag, ev in timers.items()}

    def timed(self, tag):
        return _Timed(self, tag)

    # -- dense ---------------------------------------------------------------------------------
    def gemm(self, A, B, ta=False, tb=False, out=None, alpha=1.0, beta=0.0, precision=None, bias=None, relu=False,
             mask=None):
        """out = epi(alpha op(A) op(B) + beta out); epi adds `bias` per column, applies ReLU, and zeroes the entries
        whose `mask` (same shape as out) is not positive -- all inside the product's store."""
        lda, ldb = _mat(A, "A"), _mat(B, "B")
        M, K = (A.shape[1], A.shape[0]) if ta else (A.shape[0], A.shape[1])
        K2, N = (B.shape[1], B.shape[0]) if tb else (B.shape[0], B.shape[1])
        if K != K2:
            raise ValueError(f"gemm inner dimensions differ: {K} vs {K2}")
        if out is None:
            if beta != 0.0:
                raise ValueError("beta != 0 needs an output to accumulate into")
            out = self.empty(M, N)
        ldc = _mat(out, "out")
        if out.shape != (M, N):
            raise ValueError(f"gemm output shape {tuple(out.shape)} != {(M, N)}")
        prec = self.precision if precision is None else precision
        ws, ws_bytes = None, 0
        if prec:
            key = (M, N, K, prec)
            ws_bytes = self._ws_bytes.get(key)
            if ws_bytes is None:
                ws_bytes = self._ws_bytes[key] = int(self.lib.gs_gemm_workspace_bytes(M, N, K, prec))
            if ws_bytes:
                ws = self._gemm_workspace(ws_bytes)
        if bias is None and mask is None and not relu:
            _lib.check(self.lib.gs_gemm_f32(int(ta), int(tb), M, N, K, alpha, _ptr(A), lda, _ptr(B), ldb, beta,
                                            _ptr(out), ldc, prec, _ptr(ws), ws_bytes, self.stream), "gs_gemm_f32")
            return out
        ldm = 0
        if mask is not None:
            ldm = _mat(mask, "mask")
            if mask.shape != (M, N):
                raise ValueError(f"gemm mask shape {tuple(mask.shape)} != {(M, N)}")
        if bias is not None:
            _f32(bias, "bias")
            if bias.numel() != N or not bias.is_contiguous():
                raise ValueError("gemm bias must be a contiguous vector of N elements")
        _lib.check(self.lib.gs_gemm_epi_f32(int(ta), int(tb), M, N, K, alpha, _ptr(A), lda, _ptr(B), ldb, beta,
                                            _ptr(out), ldc, _ptr(bias), int(bool(relu)), _ptr(mask), ldm, prec,
                                            _ptr(ws), ws_bytes, self.stream), "gs_gemm_epi_f32")
        return out

    def _gemm_workspace(self, nbytes):
        """Scratch for the BF16 tile image of op(B); one buffer reused by every call on this stream (stream order
        makes reuse safe: the next pack kernel runs after the previous GEMM has finished reading)."""
        ws = getattr(self, "_ws", None)
        if ws is None or ws.numel() < nbytes:
            if ws is not None:
                # captured CUDA graphs (inner loop, matching segment) hold this buffer's address: keep it alive
                self.__dict__.setdefault("_ws_retired", []).append(ws)
            ws = torch.empty(max(nbytes, 1 << 22), dtype=torch.uint8, device=self.device)
            self._ws = ws
        return ws

    def _zeroed(self, out, rows, cols):
        """A zeroed (rows x cols) result: fresh, or the caller's fixed buffer cleared in place (the matching graph's input
        buffers: writing there directly saves a device-to-device copy and an allocation per gradient and step)."""
        if out is None:
            return self.zeros(rows, cols)
        if out.shape != (rows, cols) or not out.is_contiguous() or out.dtype is not torch.float32:
            raise ValueError(f"out must be a contiguous float32 ({rows} x {cols}) tensor, got {tuple(out.shape)}")
        return out.zero_()

    def gemm_grouped_tn(self, A, B, seg, out_block, nblk, aligned=False, precision=None, out=None):
        """out[:, out_block[g]*N:(out_block[g]+1)*N] = A[seg[g]:seg[g+1]]^T @ B[seg[g]:seg[g+1]]; other blocks zero.
        aligned=True promises 64-row-aligned segments (the sampler's padding), which lets the product run on tcgen05."""
        lda, ldb = _mat(A, "A"), _mat(B, "B")
        M, N, Kt = A.shape[1], B.shape[1], A.shape[0]
        G = seg.numel() - 1
        prec = (self.precision if precision is None else precision) if aligned else 0
        if prec and self._mn_ok(A, lda, B, ldb, M, N, prec):
            # both operands consumed row-major through TMA as MN-major UMMA operands (csrc/grouped_tn.cu)
            out = self._zeroed(out, M, nblk * N)
            _lib.check(self.lib.gs_gemm_grouped_mn_f32(G, _ptr(seg), _ptr(out_block), M, N, Kt, _ptr(A), lda, _ptr(B),
                                                       ldb, _ptr(out), nblk * N, prec, self.stream),
                       "gs_gemm_grouped_mn_f32")
            return out
        ws, ws_bytes = None, 0
        if prec:
            ws_bytes = int(self.lib.gs_gemm_workspace_bytes(M, N, Kt, prec))
            if ws_bytes:
                ws = self._gemm_workspace(ws_bytes)
        if out is not None or G < nblk or prec:
            out = self._zeroed(out, M, nblk * N)
        else:
            out = self.empty(M, nblk * N)
        _lib.check(self.lib.gs_gemm_grouped_tn_f32(G, _ptr(seg), _ptr(out_block), M, N, Kt, _ptr(A), lda, _ptr(B), ldb,
                                                   _ptr(out), nblk * N, prec, _ptr(ws), ws_bytes, self.stream),
                   "gs_gemm_grouped_tn_f32")
        return out

    def _mn_ok(self, A, lda, B, ldb, M, N, prec):
        return (getattr(self, "grouped_mn", True) and lda % 4 == 0 and ldb % 4 == 0 and A.data_ptr() % 16 == 0
                and B.data_ptr() % 16 == 0 and (N * 4) % 16 == 0
                and bool(self.lib.gs_gemm_grouped_mn_supported(M, N, prec)))

    def mlp_bwd_grouped_supported(self, X, H1, dU, aligned):
        prec = self.precision
        return (aligned and prec in (1, 2) and getattr(self, "grouped_mn", True) and X.shape[1] in (128, 256)
                and H1.shape[1] == 256 and dU.shape[1] % 4 == 0 and 4 <= dU.shape[1] <= 64
                and all(t.stride(0) % 4 == 0 and t.data_ptr() % 16 == 0 for t in (X, H1, dU)))

    def mlp_bwd_grouped(self, X, H1, dU, W2, seg, out_block, nblk, out=None):
        """(gW1, gb1) of the hidden layer on the real side for all classes: dA1 = (dU W2^T) . [H1 > 0] is generated on
        chip, gW1 block g = X[rows g]^T dA1[rows g], gb1 block g = column sums of dA1[rows g]  (csrc/grouped_tn.cu)."""
        ldx, ldh, ldu, ldw = _mat(X, "X"), _mat(H1, "H1"), _mat(dU, "dU"), _mat(W2, "W2")
        M, N, Cw, Kt = X.shape[1], H1.shape[1], dU.shape[1], X.shape[0]
        G = seg.numel() - 1
        gW1 = self._zeroed(None if out is None else out[0], M, nblk * N)
        gb1 = self._zeroed(None if out is None else out[1], 1, nblk * N)
        _lib.check(self.lib.gs_mlp_bwd_grouped_f32(G, _ptr(seg), _ptr(out_block), M, N, Cw, Kt, _ptr(X), ldx, _ptr(H1),
                                                   ldh, _ptr(dU), ldu, _ptr(W2), ldw, _ptr(gW1), nblk * N, _ptr(gb1),
                                                   self.precision, self.stream), "gs_mlp_bwd_grouped_f32")
        return gW1, gb1

    def segment_colsum(self, X, seg, out_block, nblk, out=None):
        """(1 x nblk*cols): block out_block[g] holds the column sums of rows seg[g]..seg[g+1] of X; other blocks zero."""
        ldx = _mat(X, "X")
        cols = X.shape[1]
        G = seg.numel() - 1
        out = self._zeroed(out, 1, nblk * cols)
        _lib.check(self.lib.gs_segment_colsum_f32(G, _ptr(seg), _ptr(out_block), cols, _ptr(X), ldx, X.shape[0],
                                                  _ptr(out), self.stream), "gs_segment_colsum_f32")
        return out

    def colsum(self, X):
        """(1 x cols) column sums of X (bias gradients: 1^T X)."""
        rows = X.shape[0]
        cache = getattr(self, "_colsum_seg", None)
        if cache is None:
            cache = self._colsum_seg = {}
        if rows not in cache:
            cache[rows] = (torch.tensor([0, rows], dtype=torch.int32, device=self.device),
                           torch.zeros(1, dtype=torch.int32, device=self.device))
        seg, ob = cache[rows]
        return self.segment_colsum(X, seg, ob, 1)

    # -- sparse --------------------------------------------------------------------------------
    def spmm(self, csr, X, out=None, accumulate=False, tile_cols=None):
        ldx = _mat(X, "X")
        F = X.shape[1]
        n_chunks, thr, cr, cb, ce = 0, 0, None, None, None
        if csr.chunks is not None:
            cr, cb, ce, thr = csr.chunks
            n_chunks = cr.numel()
        if out is None:
            out = self.empty(csr.n_rows, F)
        ldy = _mat(out, "out")
        # dense graph + wide rows: 64-float4 column tiles with 8 gathers in flight (1.2x at 492 nnz/row, F = 602; see
        # profiles/r1_spmm_sweep_v3.json).  The library cannot see nnz without a device read, the wrapper can.
        tile = self.spmm_tile_cols(csr, X) if tile_cols is None else int(tile_cols)
        if tile and tile < F:
            # X does not fit the L2: sweep it in L2-resident column slices (gs_spmm_csr_tiled_f32)
            _lib.check(self.lib.gs_spmm_csr_tiled_f32(csr.n_rows, _ptr(csr.rowptr), _ptr(csr.col), _ptr(csr.val),
                                                      _ptr(X), ldx, F, _ptr(out), ldy, int(accumulate), n_chunks,
                                                      int(thr), _ptr(cr), _ptr(cb), _ptr(ce), tile, self.stream),
                       "gs_spmm_csr_tiled_f32")
            return out
        with _SPMM_TUNE_LOCK:
            tune = _SPMM_TUNE["state"]
            if tune != "manual":
                want = (8, 2) if (F > 256 and csr.col.numel() >= 48 * csr.n_rows) else (0, 0)
                if want != tune:
                    _lib.check(self.lib.gs_spmm_set_tuning(1, want[0], 0, 0, 0, want[1]), "gs_spmm_set_tuning")
                    _SPMM_TUNE["state"] = want
            _lib.check(self.lib.gs_spmm_csr_f32(csr.n_rows, _ptr(csr.rowptr), _ptr(csr.col), _ptr(csr.val), _ptr(X),
                                                ldx, F, _ptr(out), ldy, int(accumulate), n_chunks, int(thr), _ptr(cr),
                                                _ptr(cb), _ptr(ce), self.stream), "gs_spmm_csr_f32")
        return out

    # bytes of a gathered column slice that may live in the 126 MB L2 next to the streamed (col, val) / Y traffic
    SPMM_L2_BUDGET = int(os.environ.get("GS_SPMM_L2_BUDGET", 64 << 20))

    def spmm_tile_cols(self, csr, X):
        """Width (floats, multiple of 32) of the L2-resident column slices the wide SpMM should sweep, or 0 for the
        untiled kernel: tiling pays when X exceeds the L2, a >= 128-byte slice of every source row still fits, and the
        graph is dense enough that re-gathered feature rows (4 F B / nnz) outweigh the re-read indices (8 B / nnz / slice)."""
        F = X.shape[1]
        n_src = max(int(csr.n_cols), 1)
        if X.stride(0) % 4 or X.data_ptr() % 16 or F < 64 or n_src * F * 4 <= 2 * self.SPMM_L2_BUDGET:
            return 0
        tile = (self.SPMM_L2_BUDGET // (n_src * 4)) // 32 * 32
        if tile < 32 or tile >= F or csr.col.numel() < 16 * csr.n_rows:
            return 0
        return int(tile)

    def spmm_set_tuning(self, impl=1, unr=0, group=0, flags=0, wpb=0, max_nv=0):
        """Kernel generation / gathers in flight / rows per warp (v2) / cache hints / warps per CTA and widest column
        tile (v1) of the wide SpMM (include/graphslim_b200.h); 0 = the library's automatic choice.  An explicit call
        switches off the density heuristic of `spmm` until `spmm_auto_tuning()`."""
        with _SPMM_TUNE_LOCK:
            _lib.check(self.lib.gs_spmm_set_tuning(int(impl), int(unr), int(group), int(flags), int(wpb), int(max_nv)),
                       "gs_spmm_set_tuning")
            _SPMM_TUNE["state"] = "manual"

    def spmm_auto_tuning(self):
        self.spmm_set_tuning()
        with _SPMM_TUNE_LOCK:
            _SPMM_TUNE["state"] = (0, 0)

    def spmm_scatter(self, csr, dY, out):
        """out[col[e],:] += val[e]*dY[row(e),:] (atomics); `out` must be pre-initialised."""
        ldy, ldx = _mat(dY, "dY"), _mat(out, "out")
        _lib.check(self.lib.gs_spmm_csr_scatter_f32(csr.n_rows, _ptr(csr.rowptr), _ptr(csr.col), _ptr(csr.val),
                                                    _ptr(dY), ldy, dY.shape[1], _ptr(out), ldx, self.stream),
                   "gs_spmm_csr_scatter_f32")
        return out

    def gather_rows(self, X, idx):
        ldx = _mat(X, "X")
        out = self.empty(idx.numel(), X.shape[1])
        _lib.check(self.lib.gs_gather_rows_f32(idx.numel(), _ptr(idx), _ptr(X), ldx, X.shape[1], _ptr(out),
                                               X.shape[1], self.stream), "gs_gather_rows_f32")
        return out

    def csr_gcn_norm(self, rowptr, col, a, r64):
        out = torch.empty_like(a)
        _lib.check(self.lib.gs_csr_gcn_norm_f64(rowptr.numel() - 1, _ptr(rowptr), _ptr(col), _ptr(a), _ptr(r64),
                                                _ptr(out), self.stream), "gs_csr_gcn_norm_f64")
        return out

    # -- condense-model glue -------------------------------------------------------------------
    def bias_act(self, Z, bias, relu):
        ldz = _mat(Z, "Z")
        _lib.check(self.lib.gs_bias_act_f32(Z.shape[0], Z.shape[1], _ptr(Z), ldz, _ptr(bias), int(relu), self.stream),
                   "gs_bias_act_f32")
        return Z

    def relu_mask(self, D, H, groups=1):
        """D viewed as (rows, groups, cols) is zeroed where H (rows, cols) <= 0; in place."""
        ldh = _mat(H, "H")
        rows, cols = H.shape
        if not D.is_contiguous() or D.numel() != rows * groups * cols:
            raise ValueError("relu_mask: D must be contiguous with rows*groups*cols elements")
        _lib.check(self.lib.gs_relu_mask_f32(rows, groups, cols, _ptr(D), _ptr(H), ldh, self.stream),
                   "gs_relu_mask_f32")
        return D

    def softmax_residual(self, Z, labels, row_scale, want_nll=False):
        ldz = _mat(Z, "Z")
        rows, C = Z.shape
        S, R = self.empty(rows, C), self.empty(rows, C)
        nll = self.empty(rows) if want_nll else None
        _lib.check(self.lib.gs_softmax_residual_f32(rows, C, _ptr(Z), ldz, _ptr(labels), _ptr(row_scale), _ptr(S),
                                                    _ptr(R), _ptr(nll), self.stream), "gs_softmax_residual_f32")
        return (S, R, nll) if want_nll else (S, R)

    def expand_class_blocks(self, R, blk, nblk):
        rows, C = R.shape
        E = self.empty(rows, nblk * C)
        _lib.check(self.lib.gs_expand_class_blocks_f32(rows, C, nblk, _ptr(R.contiguous()), _ptr(blk), _ptr(E),
                                                       self.stream), "gs_expand_class_blocks_f32")
        return E

    def pick_class_blocks(self, Zf, blk, nblk):
        rows = Zf.shape[0]
        C = Zf.shape[1] // nblk
        Q = self.empty(rows, C)
        _lib.check(self.lib.gs_pick_class_blocks_f32(rows, C, nblk, _ptr(Zf.contiguous()), _ptr(blk), _ptr(Q),
                                                     self.stream), "gs_pick_class_blocks_f32")
        return Q

    def softmax_jvp(self, S, Q, row_scale):
        rows, C = S.shape
        dZ = self.empty(rows, C)
        _lib.check(self.lib.gs_softmax_jvp_f32(rows, C, _ptr(S), _ptr(Q), _ptr(row_scale), _ptr(dZ), self.stream),
                   "gs_softmax_jvp_f32")
        return dZ

    # -- gradient matching -----------------------------------------------------------------------
    def match(self, gs_list, gr_list, widths, is_bias, coeff, metric, loss_accum):
        """condensation/utils.py:12-106 on class-column gradients.

        gs_list/gr_list: per parameter a (rows x n_class*width) matrix; returns dLoss/dgs in the same
        layout and adds the loss into the device scalar ``loss_accum``.
        """
        n_class = coeff.numel()
        cols = [g.shape[1] for g in gs_list]
        off = np.concatenate([[0], np.cumsum(cols)]).astype(np.int32)
        total = int(off[-1])
        stats = self.empty(4, total)
        for p, (a, b) in enumerate(zip(gs_list, gr_list)):
            ld = _mat(a, "gs")
            if _mat(b, "gr") != ld:
                raise ValueError("match: gs/gr leading dimensions differ")
            _lib.check(self.lib.gs_match_col_stats_f32(a.shape[0], a.shape[1], _ptr(a), _ptr(b), ld,
                                                       stats.data_ptr() + 4 * int(off[p]), total, self.stream),
                       "gs_match_col_stats_f32")
        key = (tuple(cols), tuple(widths), tuple(is_bias))
        cache = getattr(self, "_match_desc", None)
        if cache is None or cache[0] != key:
            dev = lambda x: torch.tensor(np.asarray(x, dtype=np.int32), device=self.device)
            cache = (key, dev(off), dev(widths), dev([int(b) for b in is_bias]))
            self._match_desc = cache
        alpha, beta, closs = self.empty(total), self.empty(total), self.empty(n_class)
        _lib.check(self.lib.gs_match_finalize_f32(METRIC_ID[metric], len(gs_list), _ptr(cache[1]), _ptr(cache[2]),
                                                  _ptr(cache[3]), n_class, _ptr(coeff), _ptr(stats), total,
                                                  _ptr(alpha), _ptr(beta), _ptr(closs), _ptr(loss_accum),
                                                  self.stream), "gs_match_finalize_f32")
        out = []
        for p, (a, b) in enumerate(zip(gs_list, gr_list)):
            G = torch.empty_like(a)
            ld = _mat(a, "gs")
            _lib.check(self.lib.gs_match_apply_f32(a.shape[0], a.shape[1], _ptr(a), _ptr(b), ld,
                                                   alpha.data_ptr() + 4 * int(off[p]),
                                                   beta.data_ptr() + 4 * int(off[p]), _ptr(G), self.stream),
                       "gs_match_apply_f32")
            out.append(G)
        return out

    # -- dense normalisation ---------------------------------------------------------------------
    def dense_gcn_norm(self, A, out=None):
        n = A.shape[0]
        Ahat, r = (self.empty(n, n) if out is None else out), self.empty(n)
        if Ahat.shape != (n, n) or not Ahat.is_contiguous():
            raise ValueError("dense_gcn_norm: out must be a contiguous (n x n) tensor")
        _lib.check(self.lib.gs_dense_gcn_norm_fwd_f32(n, _ptr(A), _ptr(Ahat), _ptr(r), self.stream),
                   "gs_dense_gcn_norm_fwd_f32")
        return Ahat, r

    def dense_gcn_norm_bwd(self, dAhat, Ahat, r):
        n = Ahat.shape[0]
        dA, work = self.empty(n, n), self.empty(2 * n)
        _lib.check(self.lib.gs_dense_gcn_norm_bwd_f32(n, _ptr(dAhat), _ptr(Ahat), _ptr(r), _ptr(dA), _ptr(work),
                                                      self.stream), "gs_dense_gcn_norm_bwd_f32")
        return dA

    # -- PGE ---------------------------------------------------------------------------------------
    def _work(self, n):
        return torch.empty(n, dtype=torch.float64, device=self.device)

    def pge_l1_stats(self, Pa, Pb, chunk_off, eps=1e-5):
        n, h = Pa.shape
        nch = chunk_off.numel() - 1
        mean, rstd = self.empty(nch, h), self.empty(nch, h)
        _lib.check(self.lib.gs_pge_l1_stats_f32(n, h, _ptr(Pa), _ptr(Pb), nch, _ptr(chunk_off), eps, _ptr(mean),
                                                _ptr(rstd), _ptr(self._work(2 * nch * h)), self.stream),
                   "gs_pge_l1_stats_f32")
        return mean, rstd

    def pge_l1_stats_closed(self, Pa, Pb, eps=1e-5):
        """Unchunked BN1 statistics from the column statistics of Pa and Pb; returns (mean 1xh, rstd 1xh, col_mean 2xh)."""
        n, h = Pa.shape
        mean, rstd, cm = self.empty(1, h), self.empty(1, h), self.empty(2, h)
        _lib.check(self.lib.gs_pge_l1_stats_closed_f32(n, h, _ptr(Pa), _ptr(Pb), eps, _ptr(mean), _ptr(rstd), _ptr(cm),
                                                       self.stream), "gs_pge_l1_stats_closed_f32")
        return mean, rstd, cm

    def pge_bn1_bwd_closed(self, dH1, Pa, Pb, mean, rstd, gamma, beta, col_mean):
        """Unchunked BN1+ReLU backward in one pass over dH1; returns (dPa, dPb, dgamma1, dbeta1)."""
        n, h = Pa.shape
        dPa, dPb, dg, db = self.empty(n, h), self.empty(n, h), self.empty(h), self.empty(h)
        nbytes = 16 * h + 8 * n * h
        work = torch.empty((nbytes + 7) // 8, dtype=torch.float64, device=self.device)
        _lib.check(self.lib.gs_pge_bn1_bwd_closed_f32(n, h, _ptr(dH1), _ptr(Pa), _ptr(Pb), _ptr(mean), _ptr(rstd),
                                                      _ptr(gamma), _ptr(beta), _ptr(col_mean), _ptr(dPa), _ptr(dPb),
                                                      _ptr(dg), _ptr(db), _ptr(work), work.numel() * 8, self.stream),
                   "gs_pge_bn1_bwd_closed_f32")
        return dPa, dPb, dg, db

    # ---- row-sharded PGE pieces (include/graphslim_b200.h, "row-sharded PGE") ----------------------------------
    def pge_l1_expand_rows(self, Pa, Pb_rows, off_rows, mean, rstd, gamma, beta):
        """H1 of the pair rows (i, j), i over the rows of `Pb_rows` (this rank's slice), j over all rows of Pa."""
        n, h = Pa.shape
        n_i = Pb_rows.shape[0]
        H1 = self.empty(n_i * n, h)
        _lib.check(self.lib.gs_pge_l1_expand_rows_f32(n_i, n, h, _ptr(Pa), _ptr(Pb_rows), _ptr(off_rows), _ptr(mean),
                                                      _ptr(rstd), _ptr(gamma), _ptr(beta), _ptr(H1), self.stream),
                   "gs_pge_l1_expand_rows_f32")
        return H1

    def col_stats_partial(self, Y, off_rows):
        """float64 [sum(y - y[0]) | sum((y - y[0])^2)] over the rows of Y (2h)."""
        rows, h = Y.shape
        work = self._work(2 * h)
        _lib.check(self.lib.gs_col_stats_partial_f64(rows, h, _ptr(Y), _ptr(off_rows), _ptr(work), self.stream),
                   "gs_col_stats_partial_f64")
        return work

    def col_stats_combine(self, parts, counts, eps=1e-5):
        """parts (world, 3h) float64 = per-rank [S1 | S2 | shift row]; counts (world,) int64 -> mean, rstd (1, h)."""
        world, h3 = parts.shape
        h = h3 // 3
        mean, rstd = self.empty(1, h), self.empty(1, h)
        _lib.check(self.lib.gs_col_stats_combine_f32(world, h, _ptr(parts), _ptr(counts), eps, _ptr(mean), _ptr(rstd),
                                                     self.stream), "gs_col_stats_combine_f32")
        return mean, rstd

    def pge_bn1_bwd_pass_rows(self, dH1_rows, Pa, Pb, i_first, n_i, mean, rstd, gamma, beta):
        """Linear reductions of dH1 over this rank's pair rows; returns the work buffer (float64 storage:
        [t1, t2 (2h doubles) | Ga (n x h floats) | Gb (n x h floats)]) to be summed over the ranks."""
        n, h = Pa.shape
        nbytes = int(self.lib.gs_pge_bn1_bwd_work_bytes(n, h))
        work = torch.empty((nbytes + 7) // 8, dtype=torch.float64, device=self.device)
        _lib.check(self.lib.gs_pge_bn1_bwd_pass_rows_f32(int(n_i), int(i_first), n, h, _ptr(dH1_rows), _ptr(Pa), _ptr(Pb),
                                                         _ptr(mean), _ptr(rstd), _ptr(gamma), _ptr(beta), _ptr(work),
                                                         work.numel() * 8, self.stream), "gs_pge_bn1_bwd_pass_rows_f32")
        return work

    def pge_bn1_bwd_final(self, Pa, Pb, rstd, gamma, col_mean, work):
        n, h = Pa.shape
        dPa, dPb, dg, db = self.empty(n, h), self.empty(n, h), self.empty(h), self.empty(h)
        _lib.check(self.lib.gs_pge_bn1_bwd_final_f32(n, h, _ptr(Pa), _ptr(Pb), _ptr(rstd), _ptr(gamma), _ptr(col_mean),
                                                     _ptr(work), _ptr(dPa), _ptr(dPb), _ptr(dg), _ptr(db), self.stream),
                   "gs_pge_bn1_bwd_final_f32")
        return dPa, dPb, dg, db

    # ---- fused layer-2 pipeline (csrc/pge_fused.cu): H1 / dY2 / dH1 never reach HBM ---------------------------
    def pge_fused_supported(self, h, nchunks):
        """The fused tcgen05 pipeline covers the tensor-core precisions, unchunked BatchNorm and h in {128, 256}."""
        return (self.precision in (1, 2) and int(h) in (128, 256) and int(nchunks) == 1
                and getattr(self, "pge_fused", True))

    def _pge_fused_ws(self, h):
        nbytes = int(self.lib.gs_pge_fused_workspace_bytes(int(h), self.precision))
        ws = getattr(self, "_pge_ws", None)
        if ws is None or ws.numel() < nbytes:
            ws = self._pge_ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return ws

    def pge_fused_l2_fwd(self, Pa, Pb, i_first, n_i, mean1, rstd1, gamma1, beta1, W2):
        """Y2 rows (i, j), i in [i_first, i_first + n_i): relu(bn1(Pa[j] + Pb[i])) W2^T, and stats = [sum y | sum y^2]
        (2h doubles) over those rows."""
        n, h = Pa.shape
        _mat(Pa, "Pa"), _mat(Pb, "Pb")
        ldw = _mat(W2, "W2")
        Y2 = self.empty(int(n_i) * n, h)
        stats = torch.empty(2 * h, dtype=torch.float64, device=self.device)
        ws = self._pge_fused_ws(h)
        _lib.check(self.lib.gs_pge_fused_l2_fwd_f32(n, int(n_i), int(i_first), h, _ptr(Pa), _ptr(Pb), _ptr(mean1),
                                                    _ptr(rstd1), _ptr(gamma1), _ptr(beta1), _ptr(W2), ldw, _ptr(Y2),
                                                    _ptr(stats), self.precision, _ptr(ws), ws.numel(), self.stream),
                   "gs_pge_fused_l2_fwd_f32")
        return Y2, stats

    def pge_stats_finalize(self, stats, count, eps=1e-5):
        h = stats.numel() // 2
        mean, rstd = self.empty(1, h), self.empty(1, h)
        _lib.check(self.lib.gs_pge_stats_finalize_f32(h, _ptr(stats), float(count), eps, _ptr(mean), _ptr(rstd),
                                                      self.stream), "gs_pge_stats_finalize_f32")
        return mean, rstd

    def pge_bn1_work(self, n, h):
        """Zeroed [t1 | t2 (2h doubles) | Ga (n x h floats) | Gb (n x h floats)] buffer (float64 storage)."""
        nbytes = int(self.lib.gs_pge_bn1_bwd_work_bytes(n, h))
        return torch.zeros((nbytes + 7) // 8, dtype=torch.float64, device=self.device)

    def pge_fused_l2_bwd_dx(self, Pa, Pb, i_first, n_i, bn1, W2, Y2, dE, bn2, w3, s1, s2, count, work=None,
                            store=False):
        """dH1 = dY2 W2 with dY2 computed on the fly from Y2 / dE.  store=True returns dH1 (n_i*n x h, unmasked);
        otherwise the masked tile sums are accumulated into the Ga / Gb parts of `work` (pge_bn1_work)."""
        n, h = Pa.shape
        ldw = _mat(W2, "W2")
        ws = self._pge_fused_ws(h)
        dH1 = Ga = Gb = None
        if store:
            dH1 = self.empty(int(n_i) * n, h)
        else:
            base = work.data_ptr() + 16 * h
            Ga, Gb = base, base + 4 * n * h
        _lib.check(self.lib.gs_pge_fused_l2_bwd_dx_f32(n, int(n_i), int(i_first), h, _ptr(Pa), _ptr(Pb), _ptr(bn1[0]),
                                                       _ptr(bn1[1]), _ptr(bn1[2]), _ptr(bn1[3]), _ptr(W2), ldw,
                                                       _ptr(Y2), _ptr(dE), _ptr(bn2[0]), _ptr(bn2[1]), _ptr(bn2[2]),
                                                       _ptr(bn2[3]), _ptr(w3), _ptr(s1), _ptr(s2), float(count),
                                                       Ga or 0, Gb or 0, _ptr(dH1), self.precision, _ptr(ws),
                                                       ws.numel(), self.stream), "gs_pge_fused_l2_bwd_dx_f32")
        return dH1 if store else work

    def pge_fused_l2_bwd_dw(self, Pa, Pb, i_first, n_i, bn1, Y2, dE, bn2, w3, s1, s2, count):
        """dW2 (h x h) = dY2^T H1 over the slice's pair rows, both operands produced on chip."""
        n, h = Pa.shape
        dW2 = self.empty(h, h)
        _lib.check(self.lib.gs_pge_fused_l2_bwd_dw_f32(n, int(n_i), int(i_first), h, _ptr(Pa), _ptr(Pb), _ptr(bn1[0]),
                                                       _ptr(bn1[1]), _ptr(bn1[2]), _ptr(bn1[3]), _ptr(Y2), _ptr(dE),
                                                       _ptr(bn2[0]), _ptr(bn2[1]), _ptr(bn2[2]), _ptr(bn2[3]),
                                                       _ptr(w3), _ptr(s1), _ptr(s2), float(count), _ptr(dW2),
                                                       self.precision, self.stream), "gs_pge_fused_l2_bwd_dw_f32")
        return dW2

    def pge_bn1_tsum(self, Pa, Pb, col_mean, rstd1, work):
        n, h = Pa.shape
        _lib.check(self.lib.gs_pge_bn1_tsum_f64(n, h, _ptr(Pa), _ptr(Pb), _ptr(col_mean), _ptr(rstd1), _ptr(work),
                                                self.stream), "gs_pge_bn1_tsum_f64")
        return work

    def pge_l1_expand(self, Pa, Pb, chunk_off, mean, rstd, gamma, beta):
        n, h = Pa.shape
        H1 = self.empty(n * n, h)
        _lib.check(self.lib.gs_pge_l1_expand_f32(n, h, _ptr(Pa), _ptr(Pb), chunk_off.numel() - 1, _ptr(chunk_off),
                                                 _ptr(mean), _ptr(rstd), _ptr(gamma), _ptr(beta), _ptr(H1),
                                                 self.stream), "gs_pge_l1_expand_f32")
        return H1

    def col_stats_chunked(self, Y, chunk_off, eps=1e-5):
        rows, h = Y.shape
        nch = chunk_off.numel() - 1
        mean, rstd = self.empty(nch, h), self.empty(nch, h)
        _lib.check(self.lib.gs_col_stats_chunked_f32(rows, h, _ptr(Y), nch, _ptr(chunk_off), eps, _ptr(mean),
                                                     _ptr(rstd), _ptr(self._work(2 * nch * h)), self.stream),
                   "gs_col_stats_chunked_f32")
        return mean, rstd

    def pge_l3(self, Y2, chunk_off, mean, rstd, gamma, beta, w3, b3):
        rows, h = Y2.shape
        E = self.empty(rows)
        _lib.check(self.lib.gs_pge_l3_f32(rows, h, _ptr(Y2), chunk_off.numel() - 1, _ptr(chunk_off), _ptr(mean),
                                          _ptr(rstd), _ptr(gamma), _ptr(beta), _ptr(w3), _ptr(b3), _ptr(E),
                                          self.stream), "gs_pge_l3_f32")
        return E

    def pge_symm_sigmoid(self, E, n):
        A = self.empty(n, n)
        _lib.check(self.lib.gs_pge_symm_sigmoid_f32(n, _ptr(E), _ptr(A), self.stream), "gs_pge_symm_sigmoid_f32")
        return A

    def pge_symm_sigmoid_bwd(self, dA, A):
        n = A.shape[0]
        dE = self.empty(n * n)
        _lib.check(self.lib.gs_pge_symm_sigmoid_bwd_f32(n, _ptr(dA), _ptr(A), _ptr(dE), self.stream),
                   "gs_pge_symm_sigmoid_bwd_f32")
        return dE

    def pge_l3_bwd_stats(self, Y2, dE, chunk_off, mean, rstd, gamma, beta, w3):
        rows, h = Y2.shape
        nch = chunk_off.numel() - 1
        s1, s2 = self.empty(nch, h), self.empty(nch, h)
        dw3, db3 = self.zeros(h), self.zeros(1)
        _lib.check(self.lib.gs_pge_l3_bwd_stats_f32(rows, h, _ptr(Y2), _ptr(dE), nch, _ptr(chunk_off), _ptr(mean),
                                                    _ptr(rstd), _ptr(gamma), _ptr(beta), _ptr(w3), _ptr(s1), _ptr(s2),
                                                    _ptr(dw3), _ptr(db3), _ptr(self._work(2 * nch * h + h + 1)),
                                                    self.stream), "gs_pge_l3_bwd_stats_f32")
        return s1, s2, dw3, db3

    def pge_bn2_bwd_apply(self, Y2, dE, chunk_off, mean, rstd, gamma, beta, w3, s1, s2):
        rows, h = Y2.shape
        dY2 = self.empty(rows, h)
        _lib.check(self.lib.gs_pge_bn2_bwd_apply_f32(rows, h, _ptr(Y2), _ptr(dE), chunk_off.numel() - 1,
                                                     _ptr(chunk_off), _ptr(mean), _ptr(rstd), _ptr(gamma), _ptr(beta),
                                                     _ptr(w3), _ptr(s1), _ptr(s2), _ptr(dY2), self.stream),
                   "gs_pge_bn2_bwd_apply_f32")
        return dY2

    def pge_bn1_bwd_stats(self, dH1, Pa, Pb, chunk_off, mean, rstd, gamma, beta):
        n, h = Pa.shape
        nch = chunk_off.numel() - 1
        s1, s2 = self.empty(nch, h), self.empty(nch, h)
        _lib.check(self.lib.gs_pge_bn1_bwd_stats_f32(n, h, _ptr(dH1), _ptr(Pa), _ptr(Pb), nch, _ptr(chunk_off),
                                                     _ptr(mean), _ptr(rstd), _ptr(gamma), _ptr(beta), _ptr(s1),
                                                     _ptr(s2), _ptr(self._work(2 * nch * h)), self.stream),
                   "gs_pge_bn1_bwd_stats_f32")
        return s1, s2

    def pge_bn1_bwd_reduce(self, dH1, Pa, Pb, chunk_off, mean, rstd, gamma, beta, s1, s2):
        n, h = Pa.shape
        dPa, dPb = self.empty(n, h), self.empty(n, h)
        _lib.check(self.lib.gs_pge_bn1_bwd_reduce_f32(n, h, _ptr(dH1), _ptr(Pa), _ptr(Pb), chunk_off.numel() - 1,
                                                      _ptr(chunk_off), _ptr(mean), _ptr(rstd), _ptr(gamma),
                                                      _ptr(beta), _ptr(s1), _ptr(s2), _ptr(dPa), _ptr(dPb),
                                                      self.stream), "gs_pge_bn1_bwd_reduce_f32")
        return dPa, dPb

    # -- optimiser ---------------------------------------------------------------------------------
    def adam_step(self, p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
        if not (p.is_contiguous() and g.is_contiguous()):
            raise ValueError("adam_step needs contiguous tensors")
        _lib.check(self.lib.gs_adam_step_f32(p.numel(), _ptr(p), _ptr(g), _ptr(m), _ptr(v), step, lr, beta1, beta2,
                                             eps, self.stream), "gs_adam_step_f32")

    def adam_table(self, steps, lr, beta1=0.9, beta2=0.999):
        """(steps x 2) device table of Adam's step-dependent scalars for steps 1..steps (see gs_adam_table_f32)."""
        host = torch.empty(max(int(steps), 1), 2, dtype=torch.float32)
        _lib.check(self.lib.gs_adam_table_f32(int(steps), lr, beta1, beta2, host.data_ptr()), "gs_adam_table_f32")
        return host.to(self.device)

    def adam_step_table(self, p, g, m, v, table, step_dev, beta1=0.9, beta2=0.999, eps=1e-8):
        if not (p.is_contiguous() and g.is_contiguous()):
            raise ValueError("adam_step_table needs contiguous tensors")
        _lib.check(self.lib.gs_adam_step_table_f32(p.numel(), _ptr(p), _ptr(g), _ptr(m), _ptr(v), _ptr(table),
                                                   _ptr(step_dev), beta1, beta2, eps, self.stream),
                   "gs_adam_step_table_f32")

    def counter_add(self, counter, inc=1):
        _lib.check(self.lib.gs_counter_add_i32(_ptr(counter), int(inc), self.stream), "gs_counter_add_i32")

    def axpby(self, a, x, b, y):
        _lib.check(self.lib.gs_axpby_f32(y.numel(), a, _ptr(x), b, _ptr(y), self.stream), "gs_axpby_f32")
        return y
