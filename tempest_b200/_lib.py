"""ctypes binding of libtempest_b200.so (the C ABI declared in include/tempest_b200.h).

The library is the product: there is no CPU or torch fallback.  ``load()`` raises
``RuntimeError`` when the shared object is missing or CUDA is unavailable.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TEMPEST_B200_LIB") or os.path.join(_HERE, "lib", "libtempest_b200.so")

WIDE_DEFAULT = 1   # n_dim > 16: warp-cooperative runtime-d step kernel (tape parity at d = 50 / 100; 4.9x faster at d = 50)

c_i32, c_i64, c_u32, c_u64, c_f64 = C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_double
PTR = C.c_void_p
SIZE = C.c_size_t


class TbTape(C.Structure):
    _fields_ = [("gamma", PTR), ("acc_u", PTR), ("z", PTR), ("z_off", PTR), ("z_cnt", PTR),
                ("steps", c_i32)]


class TbXgpu(C.Structure):
    _fields_ = [("rank", c_i32), ("world", c_i32), ("seq", c_u64), ("peer", PTR * 8)]


class TbCdfX(C.Structure):
    _fields_ = [("rank", c_i32), ("world", c_i32), ("seq", c_u64), ("peer", PTR * 8)]


class TbXcoll(C.Structure):
    _fields_ = [("rank", c_i32), ("world", c_i32), ("seq", c_u64), ("cap_bytes", c_u64), ("ticket", PTR), ("peer", PTR * 8)]


class TbMcmcParams(C.Structure):
    _fields_ = [
        ("n_dim", c_i32), ("n_modes", c_i32), ("sampler", c_i32), ("rng_mode", c_i32),
        ("like_id", c_i32), ("prior_id", c_i32), ("n_steps", c_i32), ("n_max", c_i32),
        ("defer_update", c_i32), ("reserved", c_i32),
        ("beta", c_f64), ("seed", c_u64), ("iteration", c_u64), ("slot_offset", c_i64),
        ("n_global", c_i64), ("like_params", PTR), ("prior_params", PTR), ("mode_mean", PTR),
        ("mode_chol", PTR), ("mode_inv", PTR), ("mode_dof", PTR), ("bc_kind", PTR),
        ("xgpu", C.POINTER(TbXgpu)),
    ]


# name -> (restype, argtypes); every symbol declared in include/tempest_b200.h
SIGNATURES = {
    "tb_version": (c_i32, []),
    "tb_sm_count": (c_i32, []),
    "tb_mixture_build": (c_i32, [PTR, PTR, c_i64, PTR, PTR, PTR, c_i32, PTR]),
    "tb_mixture_append": (c_i32, [PTR, PTR, c_i64, c_i64, PTR, PTR, PTR, c_i32, PTR]),
    "tb_probe_workspace_bytes": (SIZE, []),
    "tb_probe": (c_i32, [PTR, PTR, c_i64, c_f64, PTR, PTR, PTR]),
    "tb_weights": (c_i32, [PTR, PTR, c_i64, c_f64, PTR, PTR, PTR]),
    "tb_log_weights": (c_i32, [PTR, PTR, c_i64, c_f64, PTR, PTR, PTR]),
    "tb_next_beta_workspace_bytes": (SIZE, []),
    "tb_next_beta": (c_i32, [PTR, PTR, c_i64, c_f64, c_f64, c_i32, PTR, PTR, PTR, c_i32, PTR]),
    "tb_cdf_workspace_bytes": (SIZE, [c_i64]),
    "tb_cdf_exact": (c_i32, [PTR, c_i64, PTR, PTR, PTR]),
    "tb_cdf_sequential": (c_i32, [PTR, c_i64, PTR, PTR]),
    "tb_cdf_set_chain": (None, [c_i32]),
    "tb_cdf_chain_diag_ptr": (PTR, [PTR, c_i64]),
    "tb_cdf_tile_cap": (c_i64, [c_i64, c_i32]),
    "tb_cdf_x_workspace_bytes": (SIZE, [c_i64]),
    "tb_cdf_x_table_bytes": (SIZE, [c_i64]),
    "tb_cdf_status_ptr": (PTR, [PTR]),
    "tb_cdf_total_ptr": (PTR, [PTR, c_i64]),
    "tb_cdf_exact_x": (c_i32, [PTR, c_i64, PTR, c_i32, c_i64, c_i64, PTR, PTR, C.POINTER(TbCdfX), PTR]),
    "tb_cdf_search_x": (c_i32, [PTR, c_i64, PTR, PTR, c_i64, C.POINTER(TbCdfX), PTR, c_i64, c_i32, c_f64, PTR, PTR,
                                PTR]),
    "tb_search_right": (c_i32, [PTR, c_i64, PTR, c_i64, PTR, PTR]),
    "tb_search_guide_bytes": (SIZE, [c_i32]),
    "tb_search_right_sharded_guided": (c_i32, [PTR, c_i64, PTR, PTR, PTR, c_i32, c_f64, PTR, c_i64, PTR, c_i32, PTR, PTR]),
    "tb_search_right_guided": (c_i32, [PTR, c_i64, PTR, c_i64, PTR, c_i32, PTR, PTR]),
    "tb_systematic": (c_i32, [PTR, c_i64, c_f64, c_i64, PTR, PTR, PTR]),
    "tb_gather_rows": (c_i32, [PTR, PTR, c_i32, PTR, c_i64, PTR, PTR, PTR]),
    "tb_gmm_block_doubles": (c_i64, [c_i32, c_i32]),
    "tb_gmm_offsets": (c_i32, [c_i32, c_i32, PTR]),
    "tb_gmm_workspace_bytes": (SIZE, []),
    "tb_kpp_prob": (c_i32, [PTR, PTR, PTR, c_i64, c_i32, PTR, c_i32, PTR, PTR]),
    "tb_kpp_pick": (c_i32, [PTR, c_i64, c_f64, PTR, PTR, c_i32, PTR, PTR, PTR]),
    "tb_gmm_init": (c_i32, [PTR, PTR, PTR, c_i64, c_i32, c_i32, PTR, PTR, PTR, PTR, PTR, c_f64, PTR]),
    "tb_gmm_em": (c_i32, [PTR, PTR, PTR, c_i64, c_i32, c_i32, PTR, PTR, PTR, PTR, c_f64, c_f64, c_i32, c_i32, PTR]),
    "tb_gmm_bound": (c_i32, [PTR, PTR, PTR, c_i64, c_i32, c_i32, PTR, PTR, PTR]),
    "tb_gmm_prepare": (c_i32, [PTR, c_i32, c_i32, c_f64, c_f64, PTR]),
    "tb_gmm_predict": (c_i32, [PTR, PTR, c_i64, c_i32, c_i32, PTR, PTR, PTR, c_i32, PTR, PTR]),
    "tb_col_minmax_workspace_bytes": (SIZE, [c_i32]),
    "tb_col_minmax": (c_i32, [PTR, PTR, c_i64, c_i32, PTR, PTR, PTR, PTR]),
    "tb_gather_normalised": (c_i32, [PTR, PTR, c_i64, c_i32, PTR, PTR, PTR, PTR]),
    "tb_take": (c_i32, [PTR, PTR, c_i64, PTR, PTR]),
    "tb_split_workspace_bytes": (SIZE, [c_i64]),
    "tb_split_by_label": (c_i32, [PTR, PTR, c_i64, PTR, PTR, PTR, PTR, PTR]),
    "tb_mcmc_propose": (c_i32, [c_i64, C.POINTER(TbMcmcParams), C.POINTER(TbTape), PTR, PTR, PTR, PTR, PTR, PTR, PTR,
                                PTR, PTR]),
    "tb_mcmc_accept": (c_i32, [c_i64, C.POINTER(TbMcmcParams), C.POINTER(TbTape), PTR, PTR, PTR, PTR, PTR, PTR, PTR,
                               PTR, PTR, PTR]),
    "tb_set_mcmc_wide": (c_i32, [c_i32]),
    "tb_set_mcmc_kone": (c_i32, [c_i32]),
    "tb_xrows_buffer_bytes": (SIZE, [c_i64, c_i32]),
    "tb_xrows_offset": (c_i64, [c_i64, c_i32, c_i32]),
    "tb_xrows_scatter": (c_i32, [PTR, PTR, c_i32, PTR, c_i64, c_i64, C.POINTER(TbXcoll), PTR, PTR]),
    "tb_vv_regularise": (c_i32, [PTR, PTR, c_i32, PTR, PTR, PTR, PTR]),
    "tb_vv_finish": (c_i32, [PTR, PTR, PTR, PTR, PTR, PTR]),
    "tb_xcoll_buffer_bytes": (SIZE, [SIZE]),
    "tb_xcoll_allreduce_sum": (c_i32, [PTR, c_i64, c_i32, C.POINTER(TbXcoll), PTR, PTR]),
    "tb_xcoll_allgather": (c_i32, [PTR, c_i64, PTR, C.POINTER(TbXcoll), PTR, PTR]),
    "tb_xgpu_bench_workspace_bytes": (SIZE, [c_i32]),
    "tb_xgpu_bench": (c_i32, [C.POINTER(TbXgpu), c_i32, c_i32, PTR, PTR, PTR]),
    "tb_fp64_peak_flops": (c_i64, [c_i32]),
    "tb_fp64_peak_run": (c_i32, [c_i32, PTR, PTR]),
    "tb_debug_variates": (c_i32, [c_u64, c_u64, c_i64, c_i64, c_i32, c_i32, c_f64, c_i32, c_i32, PTR, PTR, PTR, PTR]),
    "tb_moments_workspace_bytes": (SIZE, [c_i32]),
    "tb_weighted_moments": (c_i32, [PTR, PTR, c_i64, c_i32, PTR, PTR, PTR, PTR]),
    "tb_mahalanobis_cv": (c_i32, [PTR, PTR, c_i64, c_i32, PTR, PTR, PTR, PTR, PTR]),
    "tb_chol_inv": (c_i32, [PTR, c_i32, c_i32, PTR, PTR, PTR, PTR, PTR]),
    "tb_student_sigma": (c_i32, [PTR, c_i32, c_f64, PTR, PTR]),
    "tb_median_pairs": (c_i32, [PTR, c_i32, PTR, PTR]),
    "tb_add_trace_reg": (c_i32, [PTR, c_i32, c_f64, PTR]),
    "tb_reduce_workspace_bytes": (SIZE, []),
    "tb_normalize_inplace": (c_i32, [PTR, c_i64, PTR, PTR, PTR]),
    "tb_binade_hist": (c_i32, [PTR, c_i64, PTR, PTR, PTR, PTR]),
    "tb_subbin_hist": (c_i32, [PTR, c_i64, c_i32, PTR, PTR, PTR, PTR]),
    "tb_masked_sums": (c_i32, [PTR, c_i64, c_f64, PTR, PTR, PTR]),
    "tb_compact_workspace_bytes": (SIZE, [c_i64]),
    "tb_compact_ge": (c_i32, [PTR, c_i64, c_f64, c_f64, PTR, PTR, PTR, PTR, PTR]),
    "tb_select_workspace_bytes": (SIZE, [c_i32, c_i32]),
    "tb_select_ranks": (c_i32, [PTR, PTR, c_i64, c_i64, c_i32, PTR, PTR, c_i32, PTR, PTR, PTR]),
    "tb_select_pair_workspace_bytes": (SIZE, [c_i32]),
    "tb_select_pair": (c_i32, [PTR, PTR, c_i64, c_i64, c_i32, PTR, c_i64, c_i32, PTR, PTR, PTR]),
    "tb_unit_median_workspace_bytes": (SIZE, [c_i32]),
    "tb_bucket_offsets": (c_i32, [c_i32, PTR]),
    "tb_bucket_stage": (c_i32, [PTR, PTR, PTR, c_i64, c_i32, c_i64, c_i32, c_f64, c_f64, c_i32, c_i32, PTR, PTR, PTR, PTR]),
    "tb_bucket_merge": (c_i32, [PTR, PTR, PTR, c_i32, c_i32, c_i32, PTR, PTR]),
    "tb_unit_median_pair": (c_i32, [PTR, PTR, PTR, c_i64, c_i32, c_i64, PTR, PTR, PTR, PTR]),
    "tb_bucket_select_pair": (c_i32, [PTR, c_i64, c_i64, c_i32, c_f64, c_f64, c_i32, PTR, PTR, PTR, PTR]),
    "tb_count_indices": (c_i32, [PTR, c_i64, PTR, c_i64, PTR]),
    "tb_counted_moments": (c_i32, [PTR, PTR, PTR, c_i64, c_i32, c_f64, PTR, PTR, PTR, PTR]),
    "tb_prior_draw": (c_i32, [c_i64, C.POINTER(TbMcmcParams), PTR, PTR, PTR, PTR, PTR]),
    "tb_transform": (c_i32, [PTR, c_i64, C.POINTER(TbMcmcParams), PTR, PTR, PTR]),
    "tb_mcmc_workspace_bytes": (SIZE, [c_i64, c_i32]),
    "tb_mcmc_ctrl_doubles": (SIZE, [c_i32]),
    "tb_mcmc_begin": (c_i32, [c_i64, C.POINTER(TbMcmcParams), PTR, PTR, PTR, PTR, PTR, PTR]),
    "tb_mcmc_steps": (c_i32, [c_i64, C.POINTER(TbMcmcParams), C.POINTER(TbTape), PTR, PTR, PTR, PTR,
                              PTR, PTR, c_i32, PTR]),
    "tb_philox_uniform": (c_i32, [c_u64, c_u64, c_u32, c_i64, c_i64, PTR, PTR]),
    "tb_set_mcmc_generic": (c_i32, [c_i32]),
    "tb_search_right_sharded": (c_i32, [PTR, c_i64, PTR, PTR, PTR, c_i32, c_f64, PTR, c_i64, PTR, PTR]),
    "tb_scale_inplace": (c_i32, [PTR, c_i64, c_f64, PTR]),
    "tb_scale_inplace_dev": (c_i32, [PTR, c_i64, PTR, PTR]),
    "tb_select_stage": (c_i32, [PTR, PTR, c_i64, c_i64, c_i32, PTR, PTR, c_i32, PTR, PTR, c_i32, c_i32, PTR]),
    "tb_select_hist_offset": (SIZE, [c_i32, c_i32]),
    "tb_moments_partial": (c_i32, [PTR, PTR, PTR, PTR, c_i64, c_i32, c_f64, c_i32, c_i32, PTR, PTR, PTR, PTR]),
    "tb_mcmc_update": (c_i32, [C.POINTER(TbMcmcParams), PTR, PTR]),
    "tb_xgpu_buffer_bytes": (SIZE, []),
    "tb_next_beta_x": (c_i32, [PTR, PTR, c_i64, c_f64, c_f64, c_i32, PTR, PTR, PTR, c_i32, C.POINTER(TbXgpu), PTR]),
}

_lib: Optional[C.CDLL] = None


class TbError(RuntimeError):
    pass


def load_library(path: str = LIB_PATH) -> C.CDLL:
    """dlopen the library and attach signatures (no CUDA call is made)."""
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: build it with `python -m tempest_b200.build` "
            "(there is no CPU fallback for the Persistent Sampling kernels)")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
        per_call = KERNELS_PER_CALL.get(name)
        if per_call is not None:
            setattr(lib, name, _counted(fn, per_call))
    # TEMPEST_B200_WIDE=0/1 selects the runtime-dimension step kernel for n_dim > 16 (see tb_mcmc_wide.cu)
    lib.tb_set_mcmc_wide(int(os.environ.get("TEMPEST_B200_WIDE", str(WIDE_DEFAULT))))
    return lib


# kernels each entry point enqueues (csrc/*.cu); used for the `gpu_launches` claim of bench.py
KERNELS_PER_CALL = {
    "tb_mixture_build": 1, "tb_mixture_append": 1, "tb_probe": 1, "tb_weights": 1, "tb_log_weights": 1,
    "tb_next_beta": 1, "tb_cdf_exact": 7, "tb_cdf_exact_x": 6, "tb_cdf_search_x": 1, "tb_cdf_sequential": 1, "tb_search_right": 1, "tb_systematic": 1,
    "tb_gather_rows": 1, "tb_weighted_moments": 2, "tb_mahalanobis_cv": 1, "tb_chol_inv": 1,
    "tb_student_sigma": 1, "tb_median_pairs": 1, "tb_add_trace_reg": 1, "tb_normalize_inplace": 2,
    "tb_binade_hist": 1, "tb_subbin_hist": 1, "tb_masked_sums": 1, "tb_compact_ge": 3, "tb_select_ranks": 14, "tb_count_indices": 1,
    "tb_counted_moments": 2, "tb_prior_draw": 1, "tb_transform": 1, "tb_mcmc_begin": 2,
    "tb_mcmc_steps": 1, "tb_debug_variates": 1, "tb_fp64_peak_run": 1, "tb_xrows_scatter": 2, "tb_vv_regularise": 1, "tb_vv_finish": 1, "tb_xcoll_allreduce_sum": 2, "tb_xcoll_allgather": 2, "tb_philox_uniform": 1, "tb_search_right_sharded": 1,
    "tb_scale_inplace": 1, "tb_scale_inplace_dev": 1, "tb_select_stage": 1, "tb_select_pair": 11, "tb_unit_median_pair": 4, "tb_bucket_select_pair": 4, "tb_bucket_stage": 1, "tb_bucket_merge": 1, "tb_next_beta_x": 1, "tb_moments_partial": 2, "tb_mcmc_update": 1,
    "tb_search_right_guided": 2, "tb_search_right_sharded_guided": 2, "tb_mcmc_propose": 1, "tb_mcmc_accept": 1, "tb_kpp_prob": 1, "tb_kpp_pick": 1, "tb_gmm_init": lambda args: 4 + 2 * int(args[5]),
    "tb_gmm_em": lambda args: int(args[13]) * (3 + 2 * int(args[5])), "tb_gmm_bound": 1, "tb_gmm_prepare": 1,
    "tb_gmm_predict": 1, "tb_col_minmax": 1, "tb_gather_normalised": 1, "tb_take": 1, "tb_split_by_label": 3,
}
launch_count = 0


def _counted(fn, per_call):
    def wrapper(*args):
        global launch_count
        launch_count += per_call(args) if callable(per_call) else per_call
        return fn(*args)

    wrapper.__name__ = getattr(fn, "__name__", "tb_fn")
    return wrapper


def load() -> C.CDLL:
    """Library handle for compute calls: requires a CUDA device."""
    global _lib
    if _lib is None:
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("tempest_b200 needs a CUDA device (sm_100a); no CPU fallback exists")
        _lib = load_library()
    return _lib


def check(code: int, what: str = "") -> None:
    if code != 0:
        raise TbError(f"libtempest_b200: {what} failed with code {code}")
