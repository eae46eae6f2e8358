"""ctypes binding of libgraphslim_b200.so (include/graphslim_b200.h).

The library is built in-tree by ``graphslim_b200/build.py``.  There is no fallback: if the shared
object is missing or a call fails, the caller gets an exception.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgraphslim_b200.so")

c_int, c_i32, c_i64, c_f32, c_f64, c_vp = (ctypes.c_int, ctypes.c_int32, ctypes.c_int64, ctypes.c_float,
                                           ctypes.c_double, ctypes.c_void_p)

# name -> (restype, argtypes); pointers are passed as raw addresses (c_void_p)
SIGNATURES = {
    "gs_version": (c_int, []),
    "gs_launch_count": (c_i64, []),
    "gs_launch_count_reset": (None, []),
    "gs_last_error": (ctypes.c_char_p, []),
    "gs_spmm_csr_f32": (c_int, [c_i32, c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_vp, c_i64, c_int, c_i32, c_i32, c_vp,
                                c_vp, c_vp, c_vp]),
    "gs_spmm_csr_tiled_f32": (c_int, [c_i32, c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_vp, c_i64, c_int, c_i32, c_i32, c_vp,
                                      c_vp, c_vp, c_i32, c_vp]),
    "gs_spmm_set_tuning": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "gs_spmm_csr_scatter_f32": (c_int, [c_i32, c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_vp, c_i64, c_vp]),
    "gs_gather_rows_f32": (c_int, [c_i32, c_vp, c_vp, c_i64, c_i32, c_vp, c_i64, c_vp]),
    "gs_csr_gcn_norm_f64": (c_int, [c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "gs_gemm_workspace_bytes": (c_i64, [c_i32, c_i32, c_i32, c_int]),
    "gs_gemm_f32": (c_int, [c_int, c_int, c_i32, c_i32, c_i32, c_f32, c_vp, c_i64, c_vp, c_i64, c_f32, c_vp, c_i64,
                            c_int, c_vp, c_i64, c_vp]),
    "gs_gemm_epi_f32": (c_int, [c_int, c_int, c_i32, c_i32, c_i32, c_f32, c_vp, c_i64, c_vp, c_i64, c_f32, c_vp, c_i64,
                                c_vp, c_int, c_vp, c_i64, c_int, c_vp, c_i64, c_vp]),
    "gs_gemm_grouped_tn_f32": (c_int, [c_i32, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64,
                                       c_int, c_vp, c_i64, c_vp]),
    "gs_gemm_grouped_mn_supported": (c_int, [c_i32, c_i32, c_int]),
    "gs_gemm_grouped_mn_f32": (c_int, [c_i32, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64,
                                       c_int, c_vp]),
    "gs_mlp_bwd_grouped_f32": (c_int, [c_i32, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_i64, c_vp, c_i64, c_vp,
                                       c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_int, c_vp]),
    "gs_segment_colsum_f32": (c_int, [c_i32, c_vp, c_vp, c_i32, c_vp, c_i64, c_i32, c_vp, c_vp]),
    "gs_bias_act_f32": (c_int, [c_i32, c_i32, c_vp, c_i64, c_vp, c_int, c_vp]),
    "gs_relu_mask_f32": (c_int, [c_i32, c_i32, c_i32, c_vp, c_vp, c_i64, c_vp]),
    "gs_softmax_residual_f32": (c_int, [c_i32, c_i32, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "gs_expand_class_blocks_f32": (c_int, [c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "gs_pick_class_blocks_f32": (c_int, [c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "gs_softmax_jvp_f32": (c_int, [c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "gs_match_col_stats_f32": (c_int, [c_i32, c_i32, c_vp, c_vp, c_i64, c_vp, c_i64, c_vp]),
    "gs_match_finalize_f32": (c_int, [c_int, c_i32, c_vp, c_vp, c_vp, c_i32, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp,
                                      c_vp, c_vp]),
    "gs_match_apply_f32": (c_int, [c_i32, c_i32, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "gs_dense_gcn_norm_fwd_f32": (c_int, [c_i32, c_vp, c_vp, c_vp, c_vp]),
    "gs_dense_gcn_norm_bwd_f32": (c_int, [c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "gs_pge_l1_stats_f32": (c_int, [c_i32, c_i32, c_vp, c_vp, c_i32, c_vp, c_f32, c_vp, c_vp, c_vp, c_vp]),
    "gs_pge_l1_stats_closed_f32": (c_int, [c_i32, c_i32, c_vp, c_vp, c_f32, c_vp, c_vp, c_vp, c_vp]),
    "gs_pge_bn1_bwd_closed_f32": (c_int, [c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                                          c_vp, c_vp, c_vp, c_i64, c_vp]),
    "gs_pge_l1_expand_f32": (c_int, [c_i32, c_i32, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "gs_col_stats_chunked_f32": (c_int, [c_i64, c_i32, c_vp, c_i32, c_vp, c_f32, c_vp, c_vp, c_vp, c_vp]),
    "gs_pge_l3_f32": (c_int, [c_i64, c_i32, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "gs_pge_symm_sigmoid_f32": (c_int, [c_i32, c_vp, c_vp, c_vp]),
    "gs_pge_symm_sigmoid_bwd_f32": (c_int, [c_i32, c_vp, c_vp, c_vp, c_vp]),
    "gs_pge_l3_bwd_stats_f32": (c_int, [c_i64, c_i32, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                                        c_vp, c_vp, c_vp, c_vp, c_vp]),
    "gs_pge_bn2_bwd_apply_f32": (c_int, [c_i64, c_i32, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                                         c_vp, c_vp, c_vp]),
    "gs_pge_bn1_bwd_stats_f32": (c_int, [c_i32, c_i32, c_vp, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                                         c_vp, c_vp, c_vp]),
    "gs_pge_bn1_bwd_reduce_f32": (c_int, [c_i32, c_i32, c_vp, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                                          c_vp, c_vp, c_vp, c_vp]),
    "gs_pge_l1_expand_rows_f32": (c_int, [c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "gs_col_stats_partial_f64": (c_int, [c_i64, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "gs_col_stats_combine_f32": (c_int, [c_i32, c_i32, c_vp, c_vp, c_f32, c_vp, c_vp, c_vp]),
    "gs_pge_bn1_bwd_work_bytes": (c_i64, [c_i32, c_i32]),
    "gs_pge_bn1_bwd_pass_rows_f32": (c_int, [c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                                             c_i64, c_vp]),
    "gs_pge_bn1_bwd_final_f32": (c_int, [c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "gs_pge_fused_workspace_bytes": (c_i64, [c_i32, c_int]),
    "gs_pge_fused_l2_fwd_f32": (c_int, [c_i32, c_i32, c_i32, c_i32] + [c_vp] * 7 + [c_i64, c_vp, c_vp, c_int, c_vp, c_i64,
                                                                               c_vp]),
    "gs_pge_stats_finalize_f32": (c_int, [c_i32, c_vp, c_f64, c_f32, c_vp, c_vp, c_vp]),
    "gs_pge_fused_l2_bwd_dx_f32": (c_int, [c_i32, c_i32, c_i32, c_i32] + [c_vp] * 7 + [c_i64] + [c_vp] * 9 +
                                   [c_f64, c_vp, c_vp, c_vp, c_int, c_vp, c_i64, c_vp]),
    "gs_pge_fused_l2_bwd_dw_f32": (c_int, [c_i32, c_i32, c_i32, c_i32] + [c_vp] * 15 + [c_f64, c_vp, c_int, c_vp]),
    "gs_pge_bn1_tsum_f64": (c_int, [c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "gs_adam_step_f32": (c_int, [c_i64, c_vp, c_vp, c_vp, c_vp, c_i32, c_f64, c_f64, c_f64, c_f64, c_vp]),
    "gs_adam_table_f32": (c_int, [c_i32, c_f64, c_f64, c_f64, c_vp]),
    "gs_adam_step_table_f32": (c_int, [c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_f64, c_f64, c_f64, c_vp]),
    "gs_counter_add_i32": (c_int, [c_vp, c_i32, c_vp]),
    "gs_axpby_f32": (c_int, [c_i64, c_f32, c_vp, c_f32, c_vp, c_vp]),
    "gs_chain_run_f32": (c_int, [c_vp, c_i32, c_vp]),
    "gs_np_legacy_class_batches": (c_int, [c_vp, c_vp, c_i32, c_vp, c_vp, c_i32, c_vp, c_vp]),
    "gs_sampler_create": (c_vp, [c_i32, c_vp, c_vp, c_vp, c_i32, c_vp]),
    "gs_sampler_destroy": (None, [c_vp]),
    "gs_sampler_set_labels": (None, [c_vp, c_vp]),
    "gs_sampler_set_threads": (None, [c_vp, c_i32]),
    "gs_sampler_set_align": (None, [c_vp, c_i32]),
    "gs_sampler_begin_step": (c_vp, [c_vp, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "gs_sampler_finish_step": (c_i64, [c_vp, c_vp, c_vp, c_i64, c_vp]),
    "gs_sampler_sample_step": (c_i64, [c_vp, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "gs_dsampler_create": (c_vp, [c_i32, c_vp, c_vp, c_vp, c_vp, c_i32, c_vp, c_i32, c_i32, c_i32, c_vp]),
    "gs_dsampler_destroy": (None, [c_vp]),
    "gs_dsampler_out_capacity": (c_i64, [c_vp]),
    "gs_dsampler_scratch_bytes": (c_i64, [c_vp]),
    "gs_dsampler_set_rng": (c_int, [c_vp, c_vp, c_i32, c_i32, c_vp]),
    "gs_dsampler_get_rng": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp]),
    "gs_dsampler_sample_step": (c_int, [c_vp, c_i32, c_vp, c_vp, c_vp, c_i32, c_vp, c_i64, c_vp, c_vp]),
    "gs_uset_emul_order": (c_i64, [c_vp, c_i64, c_vp]),
    "gs_dsampler_debug_counters": (c_int, [c_vp, c_vp, c_vp]),
}

_lib = None


class GraphSlimLibraryError(RuntimeError):
    pass


def load():
    """Load the shared library (raises if it has not been built: there is no other compute path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GraphSlimLibraryError(
            f"{LIB_PATH} is missing. Build it with `python -m graphslim_b200.build` (needs nvcc); "
            "graphslim_b200 has no fallback compute path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)        # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().gs_last_error().decode("utf-8", "replace")
        raise GraphSlimLibraryError(f"{what} failed with code {rc}: {msg}")
