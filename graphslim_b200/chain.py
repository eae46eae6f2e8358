"""Record a routine of small dense kernel calls once, then run it as ONE persistent kernel (csrc/chain.cu).

`ChainRecorder` stands in for `CudaOps` while an engine routine (the condense-model training step of the inner loop,
graphslim/condensation/gcond.py:63-72) runs a single time: every call allocates its result and appends a `gs_chain_op`
descriptor instead of launching.  `ChainProgram.run()` then executes the whole list with `gs_chain_run_f32`.  Only the
calls the training steps make are understood; anything else raises and the caller keeps its CUDA-graph path.
"""
import struct

import torch

from . import _lib

GEMM, SOFTMAX_RESIDUAL, COLSUM, ADAM_TABLE, COUNTER_ADD, FILL = range(6)
_OP = struct.Struct("<8i4q4f8Q")           # include/graphslim_b200.h: gs_chain_op (144 bytes)


class ChainUnsupported(RuntimeError):
    pass


def _mat(t, name):
    """Leading dimension of a row-major fp32 matrix operand (same rules as ops._mat, minus the device check: the
    recorder also runs on CPU tensors in the host-logic tests)."""
    if t.dtype is not torch.float32 or t.dim() != 2:
        raise TypeError(f"{name}: expected a 2-D float32 matrix, got {t.dtype} {tuple(t.shape)}")
    st, cols = t.stride(), t.shape[1]
    if cols > 1 and st[1] != 1:
        raise ValueError(f"{name}: need unit column stride, got shape {tuple(t.shape)} stride {st}")
    ld = st[0] if t.shape[0] > 1 else (st[0] if st[0] > cols else cols)
    return int(ld if ld > cols else cols)


def _span(t):
    """Byte range [lo, hi) a (possibly strided 2-D) tensor can touch."""
    if t is None:
        return None
    lo = t.data_ptr()
    if t.numel() == 0:
        return (lo, lo)
    last = sum((s - 1) * st for s, st in zip(t.shape, t.stride()))
    return (lo, lo + (last + 1) * t.element_size())


def _overlap(a, b):
    return a is not None and b is not None and a[0] < b[1] and b[0] < a[1]


class ChainRecorder:
    def __init__(self, K):
        self.K, self.device = K, K.device
        self.precision = K.precision
        self.ops, self.keep = [], []

    # -- allocation (results of recorded operations live as long as the program) ---------------------------------
    def empty(self, *shape, dtype=torch.float32):
        t = torch.empty(*shape, dtype=dtype, device=self.device)
        self.keep.append(t)
        return t

    def zeros(self, *shape, dtype=torch.float32):
        if dtype != torch.float32:
            raise ChainUnsupported("zeros of a non-fp32 tensor")
        t = self.empty(*shape)
        self._add(dict(kind=FILL, C=t, lda=t.numel(), alpha=0.0), reads=[], writes=[t])
        return t

    def _add(self, f, reads, writes):
        for t in reads + writes:
            if t is not None:
                if t.device != torch.device(self.device) and t.device.type != "cuda":
                    raise ChainUnsupported("operand not on the device")
                self.keep.append(t)
        self.ops.append((f, [_span(t) for t in reads if t is not None], [_span(t) for t in writes if t is not None]))

    # -- the CudaOps calls a training step makes ------------------------------------------------------------------
    def gemm(self, A, B, ta=False, tb=False, out=None, alpha=1.0, beta=0.0, precision=None, bias=None, relu=False,
             mask=None):
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
        ldm = 0
        if mask is not None:
            ldm = _mat(mask, "mask")
            if mask.shape != (M, N):
                raise ValueError("gemm mask shape")
        if bias is not None:
            if bias.dtype is not torch.float32:
                raise TypeError("gemm bias must be float32")
            if bias.numel() != N or not bias.is_contiguous():
                raise ValueError("gemm bias must be a contiguous vector of N elements")
        self._add(dict(kind=GEMM, ta=int(ta), tb=int(tb), relu=int(bool(relu)), M=M, N=N, K=K, lda=lda, ldb=ldb, ldc=ldc,
                       ldmask=ldm, alpha=float(alpha), beta=float(beta), A=A, B=B, C=out, bias=bias, mask=mask),
                  reads=[A, B, bias, mask] + ([out] if beta != 0.0 else []), writes=[out])
        return out

    def softmax_residual(self, Z, labels, row_scale, want_nll=False):
        if want_nll:
            raise ChainUnsupported("softmax_residual(want_nll=True)")
        ldz = _mat(Z, "Z")
        rows, C = Z.shape
        S, R = self.empty(rows, C), self.empty(rows, C)
        self._add(dict(kind=SOFTMAX_RESIDUAL, M=rows, N=C, lda=ldz, A=Z, B=labels, bias=row_scale, C=S, p5=R),
                  reads=[Z, labels, row_scale], writes=[S, R])
        return S, R

    def colsum(self, X):
        ldx = _mat(X, "X")
        out = self.empty(1, X.shape[1])
        self._add(dict(kind=COLSUM, M=X.shape[0], N=X.shape[1], lda=ldx, A=X, C=out), reads=[X], writes=[out])
        return out

    def adam_step_table(self, p, g, m, v, table, step_dev, beta1=0.9, beta2=0.999, eps=1e-8):
        if not (p.is_contiguous() and g.is_contiguous()):
            raise ValueError("adam_step_table needs contiguous tensors")
        self._add(dict(kind=ADAM_TABLE, lda=p.numel(), alpha=float(1.0 - beta1), beta=float(beta2),
                       f0=float(1.0 - beta2), f1=float(eps), C=p, A=g, p5=m, p6=v, B=table, p7=step_dev),
                  reads=[g, table, step_dev, p, m, v], writes=[p, m, v])

    def counter_add(self, counter, inc=1):
        self._add(dict(kind=COUNTER_ADD, M=int(inc), C=counter), reads=[counter], writes=[counter])

    def __getattr__(self, name):                       # any other CudaOps call: this routine cannot be recorded
        raise ChainUnsupported(f"CudaOps.{name} is not recordable")

    # -- finish ---------------------------------------------------------------------------------------------------
    def program(self):
        if not self.ops:
            raise ChainUnsupported("nothing recorded")
        blob = bytearray()
        live_r, live_w = [], []                          # spans touched since the last grid barrier
        n_sync = 0
        for i, (f, reads, writes) in enumerate(self.ops):
            sync = 0
            if i > 0:
                raw = any(_overlap(r, w) for r in reads for w in live_w)
                waw = any(_overlap(w2, w) for w2 in writes for w in live_w)
                war = any(_overlap(w2, r) for w2 in writes for r in live_r)
                sync = int(raw or waw or war)
            if sync:
                live_r, live_w = [], []
                n_sync += 1
            live_r += reads
            live_w += writes
            ptr = lambda k: 0 if f.get(k) is None else f[k].data_ptr()
            blob += _OP.pack(f["kind"], f.get("ta", 0), f.get("tb", 0), f.get("relu", 0),
                             f.get("M", 0), f.get("N", 0), f.get("K", 0), sync,
                             f.get("lda", 0), f.get("ldb", 0), f.get("ldc", 0), f.get("ldmask", 0),
                             f.get("alpha", 0.0), f.get("beta", 0.0), f.get("f0", 0.0), f.get("f1", 0.0),
                             ptr("A"), ptr("B"), ptr("C"), ptr("bias"), ptr("mask"), ptr("p5"), ptr("p6"), ptr("p7"))
        dev = torch.frombuffer(blob, dtype=torch.uint8).clone().to(self.device)
        return ChainProgram(self.K, dev, len(self.ops), n_sync, self.keep)


class ChainProgram:
    def __init__(self, K, ops_dev, n_ops, n_sync, keep):
        self.K, self.ops_dev, self.n_ops, self.n_sync, self.keep = K, ops_dev, n_ops, n_sync, keep

    def run(self):
        _lib.check(self.K.lib.gs_chain_run_f32(self.ops_dev.data_ptr(), self.n_ops, self.K.stream), "gs_chain_run_f32")


def record(K, owners, routine):
    """Run `routine()` once with every object in `owners` (things holding a `.K`) pointed at a recorder; returns the
    ChainProgram.  Nothing is launched while recording."""
    rec = ChainRecorder(K)
    saved = [o.K for o in owners]
    for o in owners:
        o.K = rec
    try:
        routine()
    finally:
        for o, k in zip(owners, saved):
            o.K = k
    return rec.program()
