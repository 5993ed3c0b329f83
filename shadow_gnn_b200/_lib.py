"""ctypes binding of libshadow_b200.so (C ABI: include/shadow_b200.h).

There is no CPU fallback: if the CUDA library is missing this module raises at import time, and every
compute entry point of the library fails with SHADOW_ECUDA when no device is visible.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SHADOW_B200_LIB selects another build of the same library (kernel-variant exploration); there is still no CPU fallback
LIB_PATH = os.environ.get("SHADOW_B200_LIB") or os.path.join(_HERE, "libshadow_b200.so")


class ShadowError(RuntimeError):
    pass


class SamplerCfg(C.Structure):
    """shadow_sampler_cfg"""
    _fields_ = [("method", C.c_int32), ("num_roots", C.c_int32), ("depth", C.c_int32), ("budget", C.c_int32),
                ("k", C.c_int32), ("threshold", C.c_float), ("add_self_edge", C.c_int32),
                ("include_target_conn", C.c_int32), ("return_target_only", C.c_int32), ("aug", C.c_int32),
                ("fixed_mode", C.c_int32), ("rng_mode", C.c_int32)]


class BatchInfo(C.Structure):
    """shadow_batch_info"""
    _fields_ = [("num_subg", C.c_int32), ("num_roots", C.c_int32), ("total_nodes", C.c_int64),
                ("total_edges", C.c_int64), ("has_csr", C.c_int32), ("has_hop", C.c_int32), ("has_ppr", C.c_int32),
                ("has_drnl", C.c_int32), ("rand_draws", C.c_int64)]


class LinearBranch(C.Structure):
    """shadow_linear_branch"""
    _fields_ = [(n, C.c_void_p) for n in ("X", "W", "bias", "scale", "offset", "Z", "out", "mean", "rstd", "W_lo")]


METHOD = {"khop": 0, "ppr": 1, "ppr_st": 2, "nodeIID": 3}
AUG = {"hops": 1, "pprs": 2, "drnls": 4}
RNG_GLIBC, RNG_PHILOX = 0, 1
(F_NODE_PTR, F_EDGE_PTR, F_ROWPTR, F_INDICES, F_ORIG_NODE, F_ORIG_EDGE, F_TARGET, F_PPR, F_HOP, F_DRNL,
 F_NUM_TARGET, F_ROW_SPAN, F_EDGE_SPAN, F_INDICES_RAW, F_ORIG_EDGE_RAW) = range(15)

# every symbol include/shadow_b200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "shadow_last_error", "shadow_version",
    "shadow_sampler_create", "shadow_sampler_create_dev", "shadow_sampler_destroy", "shadow_sampler_set_stream",
    "shadow_sampler_num_nodes", "shadow_sampler_num_edges", "shadow_sampler_num_nodes_target",
    "shadow_sampler_get_idx_root", "shadow_sampler_set_num_per_batch", "shadow_sampler_shuffle_targets",
    "shadow_sampler_shuffle_targets_dev", "shadow_sampler_reseed", "shadow_sampler_drop_full_graph_info",
    "shadow_sampler_set_ppr_tables", "shadow_sampler_preproc_ppr_approximate", "shadow_sampler_get_ppr_row",
    "shadow_sampler_sample", "shadow_sampler_batch_info", "shadow_sampler_batch_field_dev",
    "shadow_sampler_batch_field_host", "shadow_sampler_last_redo_count", "shadow_sampler_last_sym", "shadow_sampler_last_kernel_ms", "shadow_sampler_last_sequence_ms", "shadow_gather_rows_f32", "shadow_load_batch",
    "shadow_edge_vals_fill", "shadow_edge_vals_dropedge", "shadow_edge_vals_row_normalize", "shadow_edge_vals_sym_normalize",
    "shadow_spmm_csr_fwd_f32", "shadow_spmm_csr_bwd_f32", "shadow_act_norm_fwd_f32", "shadow_act_norm_bwd_f32", "shadow_act_norm_bwd_pair_f32", "shadow_act_norm_bwd_pair_nofinish_f32", "shadow_act_norm_bwd_pair_nparts",
    "shadow_gat_fwd_f32", "shadow_gat_bwd_f32", "shadow_segment_pool_fwd_f32", "shadow_segment_pool_bwd_f32",
    "shadow_adam_clip_step_f32", "shadow_p2p_alloc", "shadow_p2p_open", "shadow_p2p_close", "shadow_p2p_free", "shadow_p2p_zero_grad_f32",
    "shadow_p2p_adam_clip_step_f32", "shadow_gemm_tf32x3_f32", "shadow_gemm_tf32x3_pair_f32",
    "shadow_linear_umma_fwd_f32", "shadow_linear_umma_dgrad_f32", "shadow_linear_umma_wgrad_f32", "shadow_linear_tc_f32", "shadow_tf32_split_f32", "shadow_tf32_split_transpose_f32", "shadow_wgrad_tc_f32", "shadow_wgrad_tc_scratch_floats",
    "shadow_gat_supported", "shadow_gat_scratch_floats", "shadow_gat_logits_fwd_f32", "shadow_gat_agg_fwd_f32", "shadow_gat_agg_bwd_f32",
    "shadow_gat_headnorm_fwd_f32", "shadow_gat_headnorm_bwd_f32", "shadow_gat_pre_bwd_f32", "shadow_colsum_finish_f32",
]

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(shadow_gnn_b200/csrc/build.sh).  shadow_gnn_b200 has no CPU fallback.")

lib = C.CDLL(LIB_PATH)
_vp, _i, _u32, _u64, _i64, _f, _cp = C.c_void_p, C.c_int, C.c_uint32, C.c_uint64, C.c_int64, C.c_float, C.c_char_p

lib.shadow_last_error.restype = _cp
lib.shadow_sampler_create.argtypes = [_vp, _vp, _u32, _u32, _cp, _cp, _i, _i, _i, _i, _i, C.POINTER(_vp)]
lib.shadow_sampler_create_dev.argtypes = [_vp, _vp, _u32, _u32, _i, _i, _i, _i, _i, C.POINTER(_vp)]
lib.shadow_sampler_destroy.argtypes = [_vp]
lib.shadow_sampler_set_stream.argtypes = [_vp, _vp]
for _n in ("num_nodes", "num_edges", "num_nodes_target", "get_idx_root"):
    getattr(lib, f"shadow_sampler_{_n}").argtypes = [_vp]
    getattr(lib, f"shadow_sampler_{_n}").restype = _u32
lib.shadow_sampler_set_num_per_batch.argtypes = [_vp, _i]
lib.shadow_sampler_shuffle_targets.argtypes = [_vp, _vp, _u32]
lib.shadow_sampler_shuffle_targets_dev.argtypes = [_vp, _vp, _u32]
lib.shadow_sampler_reseed.argtypes = [_vp, _i]
lib.shadow_sampler_drop_full_graph_info.argtypes = [_vp]
lib.shadow_sampler_set_ppr_tables.argtypes = [_vp, _vp, _vp, _vp]
lib.shadow_sampler_preproc_ppr_approximate.argtypes = [_vp, _vp, _u64, _i, _f, _f, _cp, _cp]
lib.shadow_sampler_get_ppr_row.argtypes = [_vp, _u32, _u32, _vp, _vp, C.POINTER(_u32)]
lib.shadow_sampler_sample.argtypes = [_vp, C.POINTER(SamplerCfg), _i]
lib.shadow_sampler_batch_info.argtypes = [_vp, _i, C.POINTER(BatchInfo)]
lib.shadow_sampler_batch_field_dev.argtypes = [_vp, _i, _i, C.POINTER(_vp), C.POINTER(_i64)]
lib.shadow_sampler_batch_field_host.argtypes = [_vp, _i, _i, _vp, _i64]
lib.shadow_sampler_last_redo_count.argtypes = [_vp]
lib.shadow_sampler_last_redo_count.restype = _i64
lib.shadow_sampler_last_sym.argtypes = [_vp]
lib.shadow_sampler_last_kernel_ms.argtypes = [_vp]
lib.shadow_sampler_last_kernel_ms.restype = C.c_float
lib.shadow_sampler_last_sequence_ms.argtypes = [_vp]
lib.shadow_sampler_last_sequence_ms.restype = C.c_float
lib.shadow_gather_rows_f32.argtypes = [_vp, _i64, C.c_int32, _vp, _i64, _vp, _vp]
_i32 = C.c_int32
lib.shadow_load_batch.argtypes = [_vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _i32, _vp, _vp, _i32, _vp, _vp, _i32, _vp, _vp]
lib.shadow_edge_vals_fill.argtypes = [_vp, _i32, _vp, _f, _vp]
lib.shadow_edge_vals_dropedge.argtypes = [_vp, _vp, _i32, _f, _u32, _vp, _vp, _vp]
lib.shadow_edge_vals_row_normalize.argtypes = [_vp, _i32, _i32, _vp, _vp]
lib.shadow_edge_vals_sym_normalize.argtypes = [_vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp]
lib.shadow_spmm_csr_fwd_f32.argtypes = [_vp, _vp, _i32, _vp, _vp, _vp, _i32, _i32, _f, _vp]
lib.shadow_spmm_csr_bwd_f32.argtypes = [_vp, _vp, _i32, _vp, _vp, _vp, _i32, _i32, _vp]
lib.shadow_act_norm_fwd_f32.argtypes = [_vp, _i32, _vp, _vp, _vp, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]
lib.shadow_act_norm_bwd_f32.argtypes = [_vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]
lib.shadow_gat_fwd_f32.argtypes = [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp]
lib.shadow_gat_bwd_f32.argtypes = [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp]
lib.shadow_segment_pool_fwd_f32.argtypes = [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp]
lib.shadow_segment_pool_bwd_f32.argtypes = [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp]
lib.shadow_gemm_tf32x3_f32.argtypes = [_vp, _i32, _i32, _vp, _i32, _i32, _vp, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _vp]
lib.shadow_gemm_tf32x3_pair_f32.argtypes = [_vp, _vp, _i32, _i32, _vp, _vp, _i32, _i32, _vp, _vp, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]
lib.shadow_linear_umma_fwd_f32.argtypes = [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i32, _i32, _i32, _vp]
lib.shadow_linear_umma_dgrad_f32.argtypes = [_vp, _i64, _vp, _i64, _vp, _i64, _i32, _i32, _i32, _vp]
lib.shadow_linear_umma_wgrad_f32.argtypes = [_vp, _i64, _vp, _i64, _vp, _i32, _i32, _i32, _i32, _vp]
lib.shadow_act_norm_bwd_pair_f32.argtypes = [_vp, _i32, _vp, _vp, _i32] + [_vp] * 8 + [_i32] + [_vp] * 6 + [_i32, _i32, _i32, _i32, _vp, _i64, _vp]
lib.shadow_gat_supported.argtypes = [_i32, _i32]
lib.shadow_gat_scratch_floats.argtypes = [_i32, _i32]
lib.shadow_gat_scratch_floats.restype = _i64
lib.shadow_gat_logits_fwd_f32.argtypes = [_vp] * 7 + [_i32, _i32, _i32, _vp]
lib.shadow_gat_agg_fwd_f32.argtypes = [_vp, _vp, _i32] + [_vp] * 7 + [_i32, _i32, _i32, _vp]
lib.shadow_gat_agg_bwd_f32.argtypes = [_vp, _vp, _i32] + [_vp] * 11 + [_i32, _i32, _i32, _vp]
lib.shadow_gat_headnorm_fwd_f32.argtypes = [_vp] * 7 + [_i32, _i32, _i32, _i32, _vp]
lib.shadow_gat_headnorm_bwd_f32.argtypes = [_vp] * 10 + [_i32, _i32, _i32, _i32, _vp, _i64, _vp]
lib.shadow_gat_pre_bwd_f32.argtypes = [_vp] * 14 + [_i32, _i32, _i32, _i32, _vp, _i64, _vp]
lib.shadow_colsum_finish_f32.argtypes = [_vp, _i32, _i32] + [_vp] * 7
lib.shadow_act_norm_bwd_pair_nofinish_f32.argtypes = [_vp, _i32, _vp, _vp, _i32] + [_vp] * 8 + [_i32] + [_i32, _i32, _i32, _i32, _vp, _i64, _vp]
lib.shadow_act_norm_bwd_pair_nparts.argtypes = [_i32]
lib.shadow_act_norm_bwd_pair_nparts.restype = _i32
lib.shadow_wgrad_tc_scratch_floats.argtypes = [_i32, _i32, _i32]
lib.shadow_wgrad_tc_scratch_floats.restype = _i64
lib.shadow_wgrad_tc_f32.argtypes = [_vp] * 8 + [_i32, _i32, _i32, _vp]
lib.shadow_tf32_split_f32.argtypes = [_vp, _i64, _vp, _vp, _vp]
lib.shadow_tf32_split_transpose_f32.argtypes = [_vp, _vp, _i32, _vp, _vp, _vp]
lib.shadow_linear_tc_f32.argtypes = [C.POINTER(LinearBranch), _i32, _i64, _i64, _i64, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _vp]
lib.shadow_adam_clip_step_f32.argtypes = [_vp, _vp, _vp, _vp, _i64, _f, _f, _f, _f, _f, _f, _vp, _vp, _vp]


def check(rc):
    if rc != 0:
        msg = lib.shadow_last_error().decode(errors="replace")
        if rc == -1:
            raise ValueError(msg)
        raise ShadowError(f"libshadow_b200 error {rc}: {msg}")
    return rc

lib.shadow_p2p_alloc.argtypes = [_i64, C.POINTER(_vp), _vp]
lib.shadow_p2p_open.argtypes = [_vp, C.POINTER(_vp)]
lib.shadow_p2p_close.argtypes = [_vp]
lib.shadow_p2p_free.argtypes = [_vp]
lib.shadow_p2p_zero_grad_f32.argtypes = [_vp, _vp, _i32, _i32, _i64, _vp, _vp]
lib.shadow_p2p_adam_clip_step_f32.argtypes = [_vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _i64, _f, _f, _f, _f, _f, _f, _vp, _vp, _vp, _vp]
