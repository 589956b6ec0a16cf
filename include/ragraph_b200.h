/* ragraph_b200.h -- C ABI of libragraph_b200.so (B200 / sm_100a).
 *
 * Drop-in boundary for RAGraph's retrieve -> gather -> propagate hot path.  The reference
 * (Artessay/RAGraph, pure Python/PyTorch) has no FFI; the interface each entry point
 * replaces is the torch call sequence cited beside it (paths relative to the reference
 * root).  INTEGRATION.md shows the ctypes / torch custom-op binding a maintainer adds.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; all tensors are
 *     row-major and contiguous; float pointers 16-byte aligned where noted.
 *   - the caller owns every buffer (inputs, outputs, workspace); the library never
 *     allocates or frees device memory, keeps no pointer after the call returns and never
 *     synchronises the stream.  All work is enqueued on `stream` (a cudaStream_t).
 *   - return value: RAG_OK (0) or a negative RAG_E* code; rag_last_error() gives a
 *     thread-local human readable message for the last failing call on this thread.
 *   - re-entrant: no global mutable state except that thread-local string and a
 *     per-process cache of device attributes.
 */
#ifndef RAGRAPH_B200_H_
#define RAGRAPH_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RAG_ABI_VERSION 2

#if defined(__GNUC__)
#define RAG_API __attribute__((visibility("default")))
#else
#define RAG_API
#endif

typedef void* rag_stream_t; /* cudaStream_t */

enum rag_status {
  RAG_OK = 0,
  RAG_EINVAL = -1,       /* bad argument (null pointer, negative size, k > N ...) */
  RAG_EALIGN = -2,       /* pointer not aligned as documented */
  RAG_EUNSUPPORTED = -3, /* shape outside the compiled range (k > RAG_MAX_K, d % 8 ...) */
  RAG_ECUDA = -4,        /* a CUDA runtime / driver call failed */
  RAG_EWORKSPACE = -5    /* workspace_bytes smaller than rag_*_workspace() asked for */
};

/* largest k handled by the fused similarity+top-k kernels */
#define RAG_MAX_K 128

/* similarity modes of rag_cosine_topk_f32 */
enum rag_sim_mode {
  RAG_SIM_FP32 = 0,       /* fp32 FMA on CUDA cores, exact-order oracle-grade path */
  RAG_SIM_TF32 = 1,       /* tcgen05 tf32 x tf32 -> fp32 in TMEM, raw approximate scores (|err| <= 2^-10 for unit
                             vectors); d <= 64 with k <= 26, d <= 128 with k <= 10 */
  RAG_SIM_BF16 = 2,       /* tcgen05 bf16 x bf16 -> fp32 in TMEM, raw approximate scores */
  RAG_SIM_BF16_REFINE = 3, /* bf16 tensor-core filter + fp32 re-score + certificate; results
                              equal RAG_SIM_FP32 (see RAG_SIM_F16_REFINE) */
  RAG_SIM_F16 = 4,         /* tcgen05 fp16 x fp16 -> fp32 in TMEM, raw approximate scores (8x tighter than bf16:
                              unit vectors need no exponent range) */
  RAG_SIM_F16_REFINE = 5   /* fp16 tensor-core filter + fp32 re-score + certificate; results equal RAG_SIM_FP32.
                              Rows the first pass cannot certify (more near-ties than its candidate lists hold:
                              clustered / duplicated libraries) take a SECOND tensor-core pass that collects every key
                              above (exact k-th score - error bound); only rows that overflow that too fall back to
                              the fp32 kernel.  The library's default exact mode. */
};
/* element formats of the 16-bit key shadow (rag_rows_to_shadow16) */
enum rag_shadow_fmt { RAG_FMT_BF16 = 0, RAG_FMT_F16 = 1 };
/* flags */
#define RAG_SIM_DOT 1u /* skip L2 normalisation: plain dot product (edge eval top-k,
                          RAGraph_edge/utils/metrics.py:102-117) */

#define RAG_SIM_WIDE_LISTS 2u /* tensor-core modes, d <= 128: 32-entry candidate lists per (row, key split) instead of 16 for
                                 k <= 10 -- a clustered library then certifies in the first pass (about 12 % more time on a
                                 Gaussian one); ignored where the shape has no such instantiation */

/* epilogues of rag_csr_spmm_f32 (bit flags, applied in this order) */
#define RAG_EPI_ROWNORM 1u /* divide row i by sum_j val[i,j]        (Propagation.py:15-16) */
#define RAG_EPI_BIAS 2u    /* + bias[F]                              (layers/gcn.py:37-38)  */
#define RAG_EPI_RELU 4u    /* max(x, 0)                              (Propagation.py:25)    */
#define RAG_EPI_PRELU 8u   /* x >= 0 ? x : alpha[0] * x              (layers/gcn.py:40)     */
#define RAG_EPI_BLEND 16u  /* (1 - blend_w) * y + blend_w * blend_in (RAGraph.py:53)        */
#define RAG_EPI_ACCUM 32u  /* y + accum_in  (sum of LightGCN layers, modules/RAGraph.py:327)*/

/* reduce ops of rag_gather_reduce_f32 */
enum rag_reduce_op { RAG_REDUCE_SUM = 0, RAG_REDUCE_MEAN = 1 };

RAG_API int rag_abi_version(void);
RAG_API const char* rag_last_error(void);
RAG_API const char* rag_status_string(int status);
/* number of kernels this library has launched in this process (all threads) */
RAG_API int64_t rag_launch_count(void);
/* 1 if rag_cosine_topk_f32 implements `mode` for embedding dim d and top-k k in this build, else 0 */
RAG_API int rag_sim_mode_supported(int32_t mode, int32_t d, int32_t k);

/* ---- a1/K1: F.normalize pieces (SimilarityFunctions.py:8,11) -------------------------- */
/* out_inv_norm[r] = 1 / max(||x[r,:]||_2, eps) */
RAG_API int rag_row_inv_norm_f32(const float* x, int64_t rows, int32_t d, float eps, float* out_inv_norm,
                         rag_stream_t stream);
/* out[r,:] = x[r,:] / max(||x[r,:]||_2, eps): F.normalize(x, p=2, dim=-1) as applied to library keys at insert
 * (ToyGraphBase.py:109, RAGraph_graph/.../ToyGraphBase.py:112).  out may alias x. */
RAG_API int rag_rows_normalize_f32(const float* x, int64_t rows, int32_t d, float eps, float* out,
                           rag_stream_t stream);
/* bf16 shadow of the key matrix for the tensor-core filter: out[r, 0:d] =
 * bf16_rn(x[r,:] * (normalize ? 1/max(||x[r]||,eps) : 1)), columns d..d_pad-1 zero filled.
 * d_pad % 64 == 0, out is [rows, d_pad] bf16 (uint16 storage), 16-byte aligned. */
RAG_API int rag_rows_to_bf16(const float* x, int64_t rows, int32_t d, int32_t normalize, float eps,
                     uint16_t* out, int32_t d_pad, rag_stream_t stream);

/* 16-bit shadow (fmt = RAG_FMT_BF16 | RAG_FMT_F16) with the rounding-error norms the exactness certificate of the
 * *_REFINE modes uses: err_rows[r] (nullable) = || rn16(xhat_r) - xhat_r ||_2 and *err_max (nullable, DEVICE float the
 * caller zero-initialises; updated with atomicMax so a library can be converted in pieces) = the largest of them.
 * By Cauchy-Schwarz |s_16 - s| <= err_query + err_key for unit vectors -- a proven bound about half the element-wise
 * worst case.  fp16 subnormals are flushed to zero before the error is measured. */
RAG_API int rag_rows_to_shadow16(const float* x, int64_t rows, int32_t d, int32_t fmt, int32_t normalize, float eps,
                         uint16_t* out, int32_t d_pad, float* err_rows, float* err_max, rag_stream_t stream);

/* tf32 shadow of the key matrix for RAG_SIM_TF32: out[r, 0:d] = tf32_rn(x[r,:] * (normalize ? 1/max(||x[r]||,eps) : 1))
 * stored as fp32 words (low 13 mantissa bits zero), columns d..d_pad-1 zero filled; d_pad = rag_tf32_shadow_dpad(d)
 * (32, 64 or 128; 0 if d > 128), out is [rows, d_pad] fp32, 16-byte aligned. */
RAG_API int32_t rag_tf32_shadow_dpad(int32_t d);
RAG_API int rag_rows_to_tf32(const float* x, int64_t rows, int32_t d, int32_t normalize, float eps, float* out,
                     int32_t d_pad, rag_stream_t stream);

/* ---- a1/K2: materialised similarity (API compatibility) ------------------------------- */
/* out[Q,N] = cosine (or dot with RAG_SIM_DOT) similarity, fp32.
 * Replaces SimilarityFunctions.calculate_cosine_similarity (SimilarityFunctions.py:6-16). */
RAG_API size_t rag_cosine_similarity_workspace(int64_t Q, int64_t N);
RAG_API int rag_cosine_similarity_f32(const float* q, int64_t Q, const float* keys, int64_t N, int32_t d,
                              uint32_t flags, float* out, void* workspace, size_t workspace_bytes,
                              rag_stream_t stream);

/* ---- a2/K1+K2+K3: fused similarity + top-k (scores never reach HBM) -------------------- */
/* Replaces calculate_cosine_similarity + torch.topk(largest, sorted)
 * (ToyGraphBase.py:53,67; RAGraph_edge/modules/RAGraph.py:303,311).
 *   q[Q,d], keys[N,d] fp32.  key_inv_norm[N] nullable (computed into workspace if null and
 *   cosine).  keys_shadow nullable: REQUIRED for the tensor-core modes -- the 16-bit shadow [N, round_up(d,64)] made
 *   by rag_rows_to_shadow16 (bf16 for the BF16 modes, fp16 for the F16 modes), the tf32 shadow
 *   [N, rag_tf32_shadow_dpad(d)] fp32 made by rag_rows_to_tf32 for RAG_SIM_TF32 (all L2-normalised).
 *   shadow_err nullable: DEVICE float = err_max of rag_rows_to_shadow16 for this shadow; the *_REFINE certificate
 *   then uses the measured bound, else the element-wise worst case (2^-8 bf16, 2^-11 fp16 per operand).
 *   out_scores[Q,k] fp32 descending; out_idx[Q,k] int64 = idx_offset + local row; order is
 *   deterministic: score desc, index asc.  Requires 1 <= k <= min(N, RAG_MAX_K). */
RAG_API size_t rag_cosine_topk_workspace(int64_t Q, int64_t N, int32_t d, int32_t k, int32_t mode);
RAG_API int rag_cosine_topk_f32(const float* q, int64_t Q, const float* keys, const float* key_inv_norm,
                        const void* keys_shadow, const float* shadow_err, int64_t N, int32_t d, int32_t k,
                        int32_t mode, uint32_t flags, int64_t idx_offset, float* out_scores, int64_t* out_idx,
                        void* workspace, size_t workspace_bytes, rag_stream_t stream);
/* Byte offsets, inside the caller's workspace of the same (Q, N, d, k, mode), of three consecutive int32 counters the
 * *_REFINE modes leave behind: offsets_out[0] = rows the first pass could not certify (they take the second tensor-core pass;
 * libraries below 32 768 keys, or "pass2" = 0, send them straight to the fp32 kernel), offsets_out[1] = rows that
 * fell back to the fp32 kernel, offsets_out[2] = rows whose certificate would fail under an 8x larger error bound (what a
 * bf16 filter would have: lets a caller running fp16 decide whether bf16 would do).  Two diagnostic words follow at
 * offsets_out[0] + 12 and + 16: the worker CTAs counted by the cross-split threshold sweep (0 = the sweep did not run) and
 * the (lane, 32-score chunk) hits the query-stationary filter queued.  Valid after the call's work has finished on the
 * stream; all 0 for other modes. */
RAG_API int rag_cosine_topk_stat_offsets(int64_t Q, int64_t N, int32_t d, int32_t k, int32_t mode, size_t* offsets_out);
/* Process-wide tuning / test hooks of the tensor-core path, read from the environment ONCE at load (RAG_TC_VARIANT,
 * RAG_TC_PREPASS, RAG_TC_PREPASS_MIN_TILES, RAG_TC_PREPASS_DIV, RAG_TC_KP) and settable here: name without the RAG_TC_
 * prefix in lower case ("variant": 0 auto / 1 ss / 2 ts; "prepass": 0/1; "prepass_min_tiles"; "prepass_div"; "prepass_max_tiles"; "kp": 0 auto /
 * 16 / 32; "pass2": 0/1; "gshare": 0/1 = cross-split threshold sharing by sweeping CTAs on the SMs the (query tile, key
 * split) grid leaves idle, "gshare_ctas": how many at most; "twopass": 0/1 = short key streams in two passes (group maxima, then a collect pass)
 * instead of list warm-up, "twopass_max_tiles": up to this many key tiles per CTA).  value < 0 restores the default.  Returns RAG_EINVAL for an unknown name. */
RAG_API int rag_tc_set_option(const char* name, int32_t value);
/* Introspection (pure host arithmetic; assumes 148 SMs without a device): which kernel would serve
 * rag_cosine_topk_f32(Q, N, d, k, mode, flags) / rag_topk_masked_tc_f32 (has_mask = 1), and with what geometry.
 * out[0] = kernel: 0 = the fp32 CUDA-core kernel (mode RAG_SIM_FP32 or a shape outside the tensor-core range), 1 = SS (query
 * tile in shared memory), 2 = TS (query tile in tensor memory, threshold pre-pass), 3 = two-pass mode (group maxima + collect
 * pass on the TS kernel); out[1] = query tiles, out[2] = key splits, out[3] = key tiles per CTA, out[4] = candidate-list
 * length k', out[5] = sweeping CTAs of the cross-split threshold sharing, out[6] = pre-pass tiles per CTA, out[7] = pipeline
 * stages.  out = int32[8]. */
RAG_API int rag_cosine_topk_plan(int64_t Q, int64_t N, int32_t d, int32_t k, int32_t mode, uint32_t flags, int32_t has_mask,
                         int32_t* out);

/* Small-problem retrieve in ONE launch (the reference's real call sites: RAGraph_graph/ragraph_utils/ToyGraphBase.py:56-87 --
 * one pooled query against a few hundred library rows; the few-shot and noise branches): F.normalize + similarity + top-k +
 * values[idx] + labels[idx] for Q <= 64, N <= 65536, k <= 16, Q*d <= 16384 (rag_retrieve_small_supported).
 * key_inv_norm nullable (norms then formed on the fly); values / labels nullable together with their outputs; rows are
 * copied bit exact (row sizes multiples of 4 bytes).  out_values [Q,k,value_row_bytes], out_labels [Q,k,label_row_bytes].
 * The workspace (rag_retrieve_small_workspace bytes, 256-byte aligned) must be ZERO-FILLED once by the caller and may then
 * be reused by successive calls on the same stream (the kernel re-arms its ticket); scores as the fp32 kernel (<= 1e-6). */
RAG_API int rag_retrieve_small_supported(int64_t Q, int64_t N, int32_t d, int32_t k);
RAG_API size_t rag_retrieve_small_workspace(int64_t Q, int64_t N, int32_t d, int32_t k);
RAG_API int rag_retrieve_small_f32(const float* q, int64_t Q, const float* keys, const float* key_inv_norm, int64_t N,
                           int32_t d, int32_t k, uint32_t flags, const void* values, int64_t value_row_bytes,
                           const void* labels, int64_t label_row_bytes, float* out_scores, int64_t* out_idx,
                           void* out_values, void* out_labels, void* workspace, size_t workspace_bytes,
                           rag_stream_t stream);

/* Top-k with per-query exclusion lists, fp32 path: query row r never returns the key indices
 * mask_col[mask_rowptr[r] .. mask_rowptr[r+1]) (global indices, i.e. including idx_offset; any order).  With
 * RAG_SIM_DOT this is the edge variant's evaluation ranking -- rating = U @ I^T on the GPU, history items set to
 * -inf, torch.topk(max(k)) on the CPU (RAGraph_edge/utils/metrics.py:48-53, 96-118) -- as ONE launch without the
 * [B, n_items] rating matrix or its device-to-host copy.  Rows with fewer than k admissible keys are padded with
 * index -1 / score -FLT_MAX.  Workspace: rag_cosine_topk_workspace(Q, N, d, k, RAG_SIM_FP32). */
RAG_API int rag_topk_masked_f32(const float* q, int64_t Q, const float* keys, const float* key_inv_norm, int64_t N,
                        int32_t d, int32_t k, uint32_t flags, const int64_t* mask_rowptr, const int64_t* mask_col,
                        int64_t idx_offset, float* out_scores, int64_t* out_idx, void* workspace,
                        size_t workspace_bytes, rag_stream_t stream);

/* The same ranking on the tensor cores (exact: 16-bit filter + fp32 refine + certificate, mode = RAG_SIM_F16_REFINE or
 * RAG_SIM_BF16_REFINE; d <= 128 with k <= 128, d <= 256 with k <= 10).  The filter ignores the exclusion lists; the refine
 * passes drop excluded candidates, a row whose candidate lists hold fewer than k admissible keys fails its certificate and
 * takes the second pass / the fp32 kernel, so results equal rag_topk_masked_f32's (ties within 1e-6 aside).
 * RAG_SIM_DOT: keys_shadow = rag_rows_to_shadow16(keys * key_scale, normalize = 0) with key_scale = 1 / max_j |keys[j]| (row
 * norms <= 1; shadow_err = that shadow's err_max), key_inv_norm unused (may be NULL); returned scores are q . k.
 * Without RAG_SIM_DOT (cosine): key_scale = 0, shadow / key_inv_norm as for rag_cosine_topk_f32.
 * Workspace: rag_cosine_topk_workspace(Q, N, d, k, mode). */
RAG_API int rag_topk_masked_tc_f32(const float* q, int64_t Q, const float* keys, const float* key_inv_norm,
                           const void* keys_shadow, const float* shadow_err, int64_t N, int32_t d, int32_t k,
                           int32_t mode, uint32_t flags, float key_scale, const int64_t* mask_rowptr,
                           const int64_t* mask_col, int64_t idx_offset, float* out_scores, int64_t* out_idx,
                           void* workspace, size_t workspace_bytes, rag_stream_t stream);

/* weighted two-metric variant (RAGraph_node_fewshot/ragraph_utils/ToyGraphBase.py:47-79):
 * score = w_a * cos(qa, ka) + w_b * cos(qb, kb), fp32 path only. */
RAG_API size_t rag_cosine2_topk_workspace(int64_t Q, int64_t N, int32_t da, int32_t db, int32_t k);
RAG_API int rag_cosine2_topk_f32(const float* qa, const float* ka, int32_t da, float w_a, const float* qb,
                         const float* kb, int32_t db, float w_b, int64_t Q, int64_t N, int32_t k,
                         float* out_scores, int64_t* out_idx, void* workspace,
                         size_t workspace_bytes, rag_stream_t stream);

/* ---- C1: merge of per-shard candidates after the NCCL all-gather ----------------------- */
/* scores/idx [R, Q, k_in] -> out [Q, k_out]; order score desc, index asc; k_out <= R*k_in,
 * R*k_in <= 4096. */
RAG_API int rag_topk_merge(const float* scores, const int64_t* idx, int32_t R, int64_t Q, int32_t k_in,
                   int32_t k_out, float* out_scores, int64_t* out_idx, rag_stream_t stream);

/* ---- e: sharded retrieval finish as ONE kernel over NVLink peer memory ------------------- */
/* New functionality (the reference is single-GPU): replaces "all-gather candidates -> merge ->
 * owners gather -> all-gather rows -> select" by peer-memory stores + two flag waits.
 * Every rank owns one block of a SYMMETRIC allocation (same size everywhere, every block
 * mapped into every process, e.g. torch.distributed._symmetric_memory); peers_dev is a DEVICE
 * array of `world` pointers, entry r = rank r's block as mapped in this process.  The block
 * must be zero-filled once (and all ranks synchronised) before the first call.
 * rag_xchg_layout fills offsets_out[8] = { total bytes, flags offset, result-A offset, result-A
 * bytes per parity, result-B offset, result-B bytes per parity, candidate-score offset,
 * candidate-index offset } for blocks sized for (Q_max, k_max, world, row bytes).
 * rag_sharded_finish: local_scores/local_idx [Q,k] = this rank's fused top-k with GLOBAL row
 * indices; table_a/table_b = this rank's rows [lo,hi) of the value / label tables (table_b
 * nullable); step = 1, 2, 3, ... identical on all ranks.  Outputs: out_scores/out_idx [Q,k]
 * (merged, identical on all ranks, order score desc / index asc) in ordinary memory, and
 * the gathered rows [Q,k,row_bytes] inside this rank's block at result offset + (step & 1) *
 * bytes per parity (valid until the next-but-one call).  Row copies are bit exact.
 * All ranks must call with the same Q, k, step; a missing peer traps after ~10 s. */
RAG_API int rag_xchg_layout(int64_t Q_max, int32_t k_max, int32_t world, int64_t row_bytes_a,
                    int64_t row_bytes_b, size_t* offsets_out);
RAG_API int rag_sharded_finish(const float* local_scores, const int64_t* local_idx, int64_t Q, int32_t k,
                       int32_t world, int32_t rank, void* const* peers_dev, int64_t Q_max,
                       int32_t k_max, const void* table_a, int64_t row_bytes_a, const void* table_b,
                       int64_t row_bytes_b, int64_t lo, int64_t hi, uint64_t step, float* out_scores,
                       int64_t* out_idx, rag_stream_t stream);

/* ---- a3/K5: gathers (bit exact) ------------------------------------------------------- */
/* out[m, :] = table[idx[m], :] for m < M; rows are row_bytes long (any dtype).  Negative
 * indices wrap once (torch semantics); an index still out of range writes zeros and raises
 * the sticky device flag readable with rag_gather_oob_count().
 * Replaces resource_values[topk_indices] / resource_labels[topk_indices]
 * (ToyGraphBase.py:70-71, modules/RAGraph.py:314).
 * owner_lo/owner_hi: only rows with owner_lo <= idx < owner_hi are fetched (from
 * table[idx - owner_lo]); other output rows are left untouched -- the sharded
 * "owners gather" step.  Pass 0, N for the plain gather. */
RAG_API int rag_gather_rows(const void* table, int64_t N, int64_t row_bytes, const int64_t* idx, int64_t M,
                    int64_t owner_lo, int64_t owner_hi, void* out, rag_stream_t stream);
RAG_API int64_t rag_gather_oob_count(void); /* host-side read of the sticky flag (synchronises) */

/* out[q,:] = reduce_j table[idx[q,j],:] (sum or mean over k), optionally blended:
 * out = (1-blend_w)*blend_in + blend_w*reduce when blend_in != NULL.
 * Replaces values[idx].sum(1) / .mean(1) + the convex blend
 * (RAGraph_node/RAGraph.py:49,53; modules/RAGraph.py:322,328).  Summation order j=0..k-1. */
RAG_API int rag_gather_reduce_f32(const float* table, int64_t N, int32_t d, const int64_t* idx, int64_t Q,
                          int32_t k, int32_t op, const float* blend_in, float blend_w, float* out,
                          rag_stream_t stream);

/* ---- a4/a5/a6/K6-K9: CSR SpMM with fused epilogues ------------------------------------ */
/* Y[n_rows,F] = epilogue( sum_j val[j] * X[col[j], :] ), rowptr int64[n_rows+1] or
 * int32[n_rows+1] (ptr_is_64), col int32[nnz], val fp32[nnz] nullable (= all ones).
 * Replaces torch.matmul(adj_normalized, x)+relu (Propagation.py:15-25), torch.mm(adj, XW)
 * +bias+PReLU (layers/gcn.py:32-40) and gather*w -> scatter_add_ (modules/RAGraph.py:232-240).
 * X and Y must not alias.  F % 4 == 0 uses the 128-bit path; any F >= 1 is supported. */
RAG_API int rag_csr_spmm_f32(const void* rowptr, int32_t ptr_is_64, const int32_t* col, const float* val,
                     int64_t n_rows, int64_t n_src, int64_t nnz, const float* X, int32_t F,
                     uint32_t epilogue, const float* bias, const float* alpha, const float* blend_in,
                     float blend_w, const float* accum_in, float* Y, rag_stream_t stream);

/* CSR construction on the device (no host round trip).
 * COO (edges[E,2] int64: [:,0]=src, [:,1]=dst as in modules/RAGraph.py:22-24) -> CSR by dst.
 * Entries inside a row land in atomic-cursor order (like the reference's scatter_add_, the
 * fp32 sum order is then not fixed run to run); the host layer offers a stable-sorted build
 * when bit-reproducibility matters.
 * Step 1: rag_coo_count_rows -> counts int32[n_rows] (zero-initialised by the call);
 * step 2: caller makes rowptr = exclusive cumsum(counts) (int64[n_rows+1]);
 * step 3: rag_coo_fill_csr writes col/val; cursor int32[n_rows] is scratch it zeroes. */
RAG_API int rag_coo_count_rows(const int64_t* edges, int64_t E, int64_t n_rows, int32_t* counts,
                       rag_stream_t stream);
RAG_API int rag_coo_fill_csr(const int64_t* edges, const float* w, int64_t E, int64_t n_rows,
                     const int64_t* rowptr, int32_t* cursor, int32_t* col, float* val,
                     rag_stream_t stream);
/* dense adjacency [n,n] -> per-row nonzero counts, then fill (Propagation.py / gcn.py take
 * dense block-diagonal adj). */
RAG_API int rag_dense_count_rows(const float* adj, int64_t n_rows, int64_t n_cols, int32_t* counts,
                         rag_stream_t stream);
RAG_API int rag_dense_fill_csr(const float* adj, int64_t n_rows, int64_t n_cols, const int64_t* rowptr,
                       int32_t* col, float* val, rag_stream_t stream);

/* ---- K10 / 8f-4: per-destination softmax over edge scalars ------------------------------ */
/* out[e] = mix_a * base[e] + mix_b * softmax_{e' : index[e'] == index[e]}((src[e] - lo) / span), COO order.
 * Replaces the min-max rescale + torch_scatter.scatter_softmax(edge_times, dst, dim_size) of
 * RAGraph_edge/modules/RAGraph.py:250-263 and, with base = edge_norm and mix_a = mix_b = 0.5, the mix of :267.
 * base nullable (then out = mix_b * softmax).  lo = 0, span = 1 gives torch_scatter.scatter_softmax itself.
 * range_dev nullable: DEVICE float[2] = { min(src), max_step }; when given, lo = range_dev[0] and span =
 * range_dev[1] - range_dev[0] are read on the device (the reference's tensor-valued min/max, no host sync).
 * Entries whose index is outside [0, n_groups) get softmax 0.  The group sums use fp32 atomics: like the
 * reference's kernel the summation order is not fixed (parity ~1e-6 relative). */
RAG_API size_t rag_scatter_softmax_workspace(int64_t n_groups);
RAG_API int rag_scatter_softmax_f32(const float* src, const int64_t* index, int64_t E, int64_t n_groups, float lo,
                            float span, const float* range_dev, const float* base, float mix_a, float mix_b, float* out,
                            void* workspace, size_t workspace_bytes, rag_stream_t stream);

/* ---- 8f-3: edge-variant negative sampling on the device ---------------------------------- */
/* out[m * n_neg + j] = an item drawn uniformly from [0, num_items) that is NOT in user users[m]'s history row
 * hist_items[hist_rowptr[u] .. hist_rowptr[u+1]) (int64, SORTED ascending within the row) -- the distribution of the
 * rejection loop in get_train_batch (RAGraph_edge/utils/dataloader.py:140-152: np.random.randint until the item is not
 * in train_user_set[user]); counter-based RNG keyed by (seed, slot, attempt): the same seed gives the same draws, the
 * numpy stream of the reference is not reproduced.  -1 if no admissible item turned up in 16 384 attempts. */
RAG_API int rag_negative_sample(const int64_t* users, int64_t M, int32_t n_neg, const int64_t* hist_rowptr,
                        const int64_t* hist_items, int64_t num_users, int64_t num_items, uint64_t seed, int64_t* out,
                        rag_stream_t stream);

/* ---- a9: downstream prompt + class-prototype scores (downprompt.py) ---------------------- */
/* out[n,d] = act(w[d] * x[n,d]); act 0 = identity (RAGraph_graph/downprompt.py:197-209), 1 = ELU
 * (RAGraph_node/downprompt.py:118-130).  out may alias x. */
RAG_API int rag_prompt_act_f32(const float* x, int64_t n, int32_t d, const float* w, int32_t act, float* out,
                       rag_stream_t stream);
/* out[n,C]: cosine of every (optionally prompted: act(w * x), w nullable) row against the C <= 32 class
 * prototypes proto[C,d], torch.cosine_similarity semantics (each norm clamped at eps, default 1e-8), then
 * mode 0 = raw scores, 1 = softmax over classes (RAGraph_node/downprompt.py:36-46), 2 = log_softmax
 * (RAGraph_graph/downprompt.py:41-56 `predict`).  Replaces the reference's per-(row, class) Python loop. */
RAG_API int rag_prototype_scores_f32(const float* x, int64_t n, int32_t d, const float* w, int32_t act,
                             const float* proto, int32_t C, float eps, int32_t mode, float* out,
                             rag_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RAGRAPH_B200_H_ */
