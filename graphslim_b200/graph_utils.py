"""Host helpers for the full-graph SpMM: work splitting of power-law rows."""
import numpy as np


def build_row_chunks(rowptr, max_nnz=512):
    """Split every row into work items of at most ``max_nnz`` non-zeros (int32 arrays: row, begin, end).

    Long rows of a power-law graph would otherwise serialise on one warp; the SpMM kernel accumulates the
    items of a row with atomics (gs_spmm_csr_f32, chunk_* arguments).
    """
    rowptr = np.asarray(rowptr, dtype=np.int64)
    deg = rowptr[1:] - rowptr[:-1]
    n_items = np.maximum((deg + max_nnz - 1) // max_nnz, 1)
    rows = np.repeat(np.arange(deg.size, dtype=np.int64), n_items)
    first = np.concatenate([[0], np.cumsum(n_items)[:-1]])
    k = np.arange(rows.size, dtype=np.int64) - np.repeat(first, n_items)
    beg = rowptr[rows] + k * max_nnz
    end = np.minimum(beg + max_nnz, rowptr[rows + 1])
    return rows.astype(np.int32), beg.astype(np.int32), end.astype(np.int32)
