"""Host helpers for the full-graph SpMM: work splitting of power-law rows."""
import numpy as np


def build_row_chunks(rowptr, max_nnz=64):
    """Slices of the rows with more than ``max_nnz`` non-zeros: int32 arrays (row, begin, end) and the threshold.

    A 1e4-1e5-degree hub of a power-law graph would otherwise serialise on one warp; the SpMM kernel clears those rows
    and accumulates their slices with atomics (gs_spmm_csr_f32, chunk_* arguments).  Every other row stays a plain
    warp-per-row item with a direct store.
    """
    rowptr = np.asarray(rowptr, dtype=np.int64)
    deg = rowptr[1:] - rowptr[:-1]
    long_rows = np.nonzero(deg > max_nnz)[0]
    n_items = (deg[long_rows] + max_nnz - 1) // max_nnz
    rows = np.repeat(long_rows, n_items)
    first = np.concatenate([[0], np.cumsum(n_items)[:-1]]) if long_rows.size else np.zeros(0, dtype=np.int64)
    k = np.arange(rows.size, dtype=np.int64) - np.repeat(first, n_items)
    beg = rowptr[rows] + k * max_nnz
    end = np.minimum(beg + max_nnz, rowptr[rows + 1])
    return rows.astype(np.int32), beg.astype(np.int32), end.astype(np.int32), int(max_nnz)


def chunks_to_device(chunks, device):
    """(row, beg, end, thr) numpy -> the tuple Csr.chunks expects (None when no row is long)."""
    import torch
    rows, beg, end, thr = chunks
    if rows.size == 0:
        return None
    return (torch.from_numpy(rows).to(device), torch.from_numpy(beg).to(device), torch.from_numpy(end).to(device), thr)
