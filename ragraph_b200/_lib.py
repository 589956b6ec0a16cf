"""ctypes binding of libragraph_b200.so (the C ABI declared in include/ragraph_b200.h).

There is NO CPU fallback: if the library is missing this module raises at first use, and every
entry point takes device pointers only.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libragraph_b200.so")

RAG_OK = 0
ABI_VERSION = 2
RAG_MAX_K = 128
SIM_FP32, SIM_TF32, SIM_BF16, SIM_BF16_REFINE, SIM_F16, SIM_F16_REFINE = 0, 1, 2, 3, 4, 5
FMT_BF16, FMT_F16 = 0, 1
SIM_DOT = 1
SIM_WIDE_LISTS = 2
EPI_ROWNORM, EPI_BIAS, EPI_RELU, EPI_PRELU, EPI_BLEND, EPI_ACCUM = 1, 2, 4, 8, 16, 32
REDUCE_SUM, REDUCE_MEAN = 0, 1
ACT_NONE, ACT_ELU = 0, 1
SCORES_RAW, SCORES_SOFTMAX, SCORES_LOG_SOFTMAX = 0, 1, 2

_p, _i64, _i32, _u32, _f32, _sz = C.c_void_p, C.c_int64, C.c_int32, C.c_uint32, C.c_float, C.c_size_t

# name -> (restype, argtypes); mirrors include/ragraph_b200.h one to one
SIGNATURES = {
    "rag_abi_version": (C.c_int, []),
    "rag_last_error": (C.c_char_p, []),
    "rag_status_string": (C.c_char_p, [C.c_int]),
    "rag_launch_count": (_i64, []),
    "rag_sim_mode_supported": (C.c_int, [_i32, _i32, _i32]),
    "rag_row_inv_norm_f32": (C.c_int, [_p, _i64, _i32, _f32, _p, _p]),
    "rag_rows_normalize_f32": (C.c_int, [_p, _i64, _i32, _f32, _p, _p]),
    "rag_rows_to_bf16": (C.c_int, [_p, _i64, _i32, _i32, _f32, _p, _i32, _p]),
    "rag_rows_to_shadow16": (C.c_int, [_p, _i64, _i32, _i32, _i32, _f32, _p, _i32, _p, _p, _p]),
    "rag_tf32_shadow_dpad": (_i32, [_i32]),
    "rag_rows_to_tf32": (C.c_int, [_p, _i64, _i32, _i32, _f32, _p, _i32, _p]),
    "rag_cosine_similarity_workspace": (_sz, [_i64, _i64]),
    "rag_cosine_similarity_f32": (C.c_int, [_p, _i64, _p, _i64, _i32, _u32, _p, _p, _sz, _p]),
    "rag_cosine_topk_workspace": (_sz, [_i64, _i64, _i32, _i32, _i32]),
    "rag_cosine_topk_f32": (C.c_int, [_p, _i64, _p, _p, _p, _p, _i64, _i32, _i32, _i32, _u32, _i64, _p, _p, _p, _sz, _p]),
    "rag_cosine_topk_stat_offsets": (C.c_int, [_i64, _i64, _i32, _i32, _i32, _p]),
    "rag_tc_set_option": (C.c_int, [C.c_char_p, _i32]),
    "rag_cosine_topk_plan": (C.c_int, [_i64, _i64, _i32, _i32, _i32, _u32, _i32, _p]),
    "rag_retrieve_small_supported": (C.c_int, [_i64, _i64, _i32, _i32]),
    "rag_retrieve_small_workspace": (_sz, [_i64, _i64, _i32, _i32]),
    "rag_retrieve_small_f32": (C.c_int, [_p, _i64, _p, _p, _i64, _i32, _i32, _u32, _p, _i64, _p, _i64, _p, _p, _p, _p, _p, _sz, _p]),
    "rag_topk_masked_f32": (C.c_int, [_p, _i64, _p, _p, _i64, _i32, _i32, _u32, _p, _p, _i64, _p, _p, _p, _sz, _p]),
    "rag_topk_masked_tc_f32": (C.c_int, [_p, _i64, _p, _p, _p, _p, _i64, _i32, _i32, _i32, _u32, C.c_float, _p, _p, _i64, _p, _p,
                                        _p, _sz, _p]),
    "rag_cosine2_topk_workspace": (_sz, [_i64, _i64, _i32, _i32, _i32]),
    "rag_cosine2_topk_f32": (C.c_int, [_p, _p, _i32, _f32, _p, _p, _i32, _f32, _i64, _i64, _i32, _p, _p, _p, _sz, _p]),
    "rag_topk_merge": (C.c_int, [_p, _p, _i32, _i64, _i32, _i32, _p, _p, _p]),
    "rag_xchg_layout": (C.c_int, [_i64, _i32, _i32, _i64, _i64, _p]),
    "rag_sharded_finish": (C.c_int, [_p, _p, _i64, _i32, _i32, _i32, _p, _i64, _i32, _p, _i64, _p, _i64, _i64, _i64,
                                     C.c_uint64, _p, _p, _p]),
    "rag_gather_rows": (C.c_int, [_p, _i64, _i64, _p, _i64, _i64, _i64, _p, _p]),
    "rag_gather_oob_count": (_i64, []),
    "rag_gather_reduce_f32": (C.c_int, [_p, _i64, _i32, _p, _i64, _i32, _i32, _p, _f32, _p, _p]),
    "rag_csr_spmm_f32": (C.c_int, [_p, _i32, _p, _p, _i64, _i64, _i64, _p, _i32, _u32, _p, _p, _p, _f32, _p, _p, _p]),
    "rag_coo_count_rows": (C.c_int, [_p, _i64, _i64, _p, _p]),
    "rag_coo_fill_csr": (C.c_int, [_p, _p, _i64, _i64, _p, _p, _p, _p, _p]),
    "rag_dense_count_rows": (C.c_int, [_p, _i64, _i64, _p, _p]),
    "rag_dense_fill_csr": (C.c_int, [_p, _i64, _i64, _p, _p, _p, _p]),
    "rag_scatter_softmax_workspace": (_sz, [_i64]),
    "rag_scatter_softmax_f32": (C.c_int, [_p, _p, _i64, _i64, _f32, _f32, _p, _p, _f32, _f32, _p, _p, _sz, _p]),
    "rag_negative_sample": (C.c_int, [_p, _i64, _i32, _p, _p, _i64, _i64, C.c_uint64, _p, _p]),
    "rag_prompt_act_f32": (C.c_int, [_p, _i64, _i32, _p, _i32, _p, _p]),
    "rag_prototype_scores_f32": (C.c_int, [_p, _i64, _i32, _p, _i32, _p, _i32, _f32, _i32, _p, _p]),
}

_lock = threading.Lock()
_lib = None


class RagError(RuntimeError):
    """A C-ABI call returned a negative RAG_E* status."""


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"ragraph_b200: {LIB_PATH} is missing -- build it with `python -m ragraph_b200.build` "
                    "(nvcc, sm_100a). There is no CPU or PyTorch fallback for the retrieval/propagation ops.")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
                fn.restype, fn.argtypes = res, args
            if lib.rag_abi_version() != ABI_VERSION:
                raise RuntimeError(f"ragraph_b200: ABI version {lib.rag_abi_version()} != {ABI_VERSION}")
            _lib = lib
    return _lib


def check(status: int, what: str) -> None:
    if status != RAG_OK:
        lib = load()
        raise RagError(f"{what}: {lib.rag_status_string(status).decode()}: {lib.rag_last_error().decode()}")


PLAN_KERNELS = {0: "fp32", 1: "ss", 2: "ts", 3: "two_pass"}


def cosine_topk_plan(Q: int, N: int, d: int, k: int, mode: int, flags: int = 0, has_mask: bool = False) -> dict:
    """Which kernel would serve cosine_topk(Q, N, d, k, mode) and with what geometry (rag_cosine_topk_plan: host arithmetic
    only, works without a GPU -- 148 SMs assumed)."""
    out = (C.c_int32 * 8)()
    check(load().rag_cosine_topk_plan(Q, N, d, k, mode, flags, int(has_mask), out), "cosine_topk_plan")
    keys = ("kernel", "q_tiles", "key_splits", "tiles_per_cta", "list_len", "sweep_ctas", "prepass_tiles", "stages")
    r = dict(zip(keys, list(out)))
    r["kernel"] = PLAN_KERNELS[r["kernel"]]
    return r


def tc_set_option(name: str, value: int) -> None:
    """process-wide tuning / test hook of the tensor-core retrieval path (see rag_tc_set_option); value < 0 = default"""
    check(load().rag_tc_set_option(name.encode(), int(value)), "tc_set_option")


def launch_count() -> int:
    return int(load().rag_launch_count())
