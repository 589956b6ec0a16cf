// C-ABI dispatch of the fused similarity + top-k entry points.
#include "common.cuh"

namespace rag {
// topk_f32.cu
size_t topk_f32_workspace(int64_t Q, int64_t N, int d, int k);
int topk_f32_run(const float* q, int64_t Q, const float* keys, const float* key_inv_norm, int64_t N, int d, int k,
                 uint32_t flags, int64_t idx_offset, float* out_scores, int64_t* out_idx, void* ws, size_t ws_bytes,
                 cudaStream_t s, const int64_t* mask_rowptr = nullptr, const int64_t* mask_col = nullptr);
// topk_tc.cu (tcgen05 filter + fp32 refine)
size_t topk_tc_workspace(int64_t Q, int64_t N, int d, int k, int mode);
bool topk_tc_available(int d, int k);
bool topk_tc_tf32_available(int d, int k);
int topk_tc_tf32_dpad(int d);
int topk_tc_run(const float* q, int64_t Q, const float* keys, const float* key_inv_norm, const void* keys_shadow,
                const float* shadow_err, int64_t N, int d, int k, int mode, uint32_t flags, int64_t idx_offset,
                float* out_scores, int64_t* out_idx, void* ws, size_t ws_bytes, cudaStream_t s,
                const int64_t* mask_rowptr = nullptr, const int64_t* mask_col = nullptr, float key_scale = 0.f);
void topk_tc_stat_offsets(int64_t Q, int64_t N, int d, int k, int mode, size_t* out);

// out[r, :] = [ wa * xa[r]/max(|xa[r]|,eps)  (padded to da4) | wb * xb[r]/max(|xb[r]|,eps) (padded to db4) ]
__global__ void __launch_bounds__(256) concat_normalized_kernel(const float* __restrict__ xa, int da, float wa,
                                                                const float* __restrict__ xb, int db, float wb,
                                                                int64_t rows, int da4, int db4,
                                                                float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < rows; r += nwarps) {
    float sa = 0.f, sb = 0.f;
    for (int c = lane; c < da; c += 32) { float v = __ldg(xa + r * da + c); sa = fmaf(v, v, sa); }
    for (int c = lane; c < db; c += 32) { float v = __ldg(xb + r * db + c); sb = fmaf(v, v, sb); }
    sa = wa / fmaxf(sqrtf(warp_sum(sa)), 1e-12f);
    sb = wb / fmaxf(sqrtf(warp_sum(sb)), 1e-12f);
    float* o = out + r * (da4 + db4);
    for (int c = lane; c < da4; c += 32) o[c] = c < da ? __ldg(xa + r * da + c) * sa : 0.f;
    for (int c = lane; c < db4; c += 32) o[da4 + c] = c < db ? __ldg(xb + r * db + c) * sb : 0.f;
  }
}
}  // namespace rag

static int check_topk_args(const char* fn, const float* q, int64_t Q, const float* keys, int64_t N, int32_t d,
                           int32_t k, float* out_scores, int64_t* out_idx) {
  RAG_REQUIRE(Q >= 0 && N >= 0 && d >= 1, RAG_EINVAL, "%s: Q=%lld N=%lld d=%d", fn, (long long)Q, (long long)N, d);
  RAG_REQUIRE(k >= 1 && (int64_t)k <= N, RAG_EINVAL, "%s: k=%d must satisfy 1 <= k <= N=%lld", fn, k, (long long)N);
  RAG_REQUIRE(k <= RAG_MAX_K, RAG_EUNSUPPORTED, "%s: k=%d exceeds RAG_MAX_K=%d of the fused kernels", fn, k,
              RAG_MAX_K);
  RAG_REQUIRE(N <= 0x7fffffffLL, RAG_EUNSUPPORTED, "%s: N=%lld per shard exceeds 2^31-1", fn, (long long)N);
  if (Q == 0) return RAG_OK;
  RAG_REQUIRE(q && keys && out_scores && out_idx, RAG_EINVAL, "%s: null pointer", fn);
  RAG_REQUIRE(rag::aligned16(q) && rag::aligned16(keys), RAG_EALIGN, "%s: q/keys must be 16-byte aligned", fn);
  return RAG_OK;
}

extern "C" int rag_sim_mode_supported(int32_t mode, int32_t d, int32_t k) {
  if (mode == RAG_SIM_FP32) return (d >= 1 && k >= 1 && k <= RAG_MAX_K) ? 1 : 0;
  if (mode == RAG_SIM_BF16 || mode == RAG_SIM_BF16_REFINE || mode == RAG_SIM_F16 || mode == RAG_SIM_F16_REFINE)
    return rag::topk_tc_available(d, k) ? 1 : 0;
  if (mode == RAG_SIM_TF32) return rag::topk_tc_tf32_available(d, k) ? 1 : 0;
  return 0;
}

extern "C" int32_t rag_tf32_shadow_dpad(int32_t d) { return (d >= 1 && d <= 128) ? rag::topk_tc_tf32_dpad(d) : 0; }

extern "C" size_t rag_cosine_topk_workspace(int64_t Q, int64_t N, int32_t d, int32_t k, int32_t mode) {
  if (Q <= 0 || N <= 0 || d < 1 || k < 1) return 256;
  size_t f32 = rag::topk_f32_workspace(Q, N, d, k);
  if (mode == RAG_SIM_FP32) return f32;
  return rag::topk_tc_workspace(Q, N, d, k, mode);
}

extern "C" int rag_cosine_topk_stat_offsets(int64_t Q, int64_t N, int32_t d, int32_t k, int32_t mode, size_t* offsets_out) {
  RAG_REQUIRE(offsets_out, RAG_EINVAL, "cosine_topk_stat_offsets: null pointer");
  rag::topk_tc_stat_offsets(Q, N, d, k, mode, offsets_out);
  return RAG_OK;
}

extern "C" int rag_cosine_topk_f32(const float* q, int64_t Q, const float* keys, const float* key_inv_norm,
                                   const void* keys_shadow, const float* shadow_err, int64_t N, int32_t d, int32_t k,
                                   int32_t mode, uint32_t flags, int64_t idx_offset, float* out_scores, int64_t* out_idx,
                                   void* workspace, size_t workspace_bytes, rag_stream_t stream) {
  int st = check_topk_args("cosine_topk", q, Q, keys, N, d, k, out_scores, out_idx);
  if (st || Q == 0) return st;
  cudaStream_t s = (cudaStream_t)stream;
  switch (mode) {
    case RAG_SIM_FP32:
      return rag::topk_f32_run(q, Q, keys, key_inv_norm, N, d, k, flags, idx_offset, out_scores, out_idx, workspace,
                               workspace_bytes, s);
    case RAG_SIM_TF32:
    case RAG_SIM_BF16:
    case RAG_SIM_BF16_REFINE:
    case RAG_SIM_F16:
    case RAG_SIM_F16_REFINE:
      RAG_REQUIRE(keys_shadow, RAG_EINVAL, "cosine_topk: mode %d needs the key shadow (rag_rows_to_shadow16 / rag_rows_to_tf32)", mode);
      return rag::topk_tc_run(q, Q, keys, key_inv_norm, keys_shadow, shadow_err, N, d, k, mode, flags, idx_offset,
                              out_scores, out_idx, workspace, workspace_bytes, s);
    default:
      return rag::fail(RAG_EINVAL, "cosine_topk: unknown mode %d", mode);
  }
}

extern "C" int rag_topk_masked_f32(const float* q, int64_t Q, const float* keys, const float* key_inv_norm, int64_t N,
                                   int32_t d, int32_t k, uint32_t flags, const int64_t* mask_rowptr,
                                   const int64_t* mask_col, int64_t idx_offset, float* out_scores, int64_t* out_idx,
                                   void* workspace, size_t workspace_bytes, rag_stream_t stream) {
  int st = check_topk_args("topk_masked", q, Q, keys, N, d, k, out_scores, out_idx);
  if (st || Q == 0) return st;
  RAG_REQUIRE((mask_rowptr == nullptr) == (mask_col == nullptr) || mask_rowptr, RAG_EINVAL,
              "topk_masked: mask_col without mask_rowptr");
  return rag::topk_f32_run(q, Q, keys, key_inv_norm, N, d, k, flags, idx_offset, out_scores, out_idx, workspace,
                           workspace_bytes, (cudaStream_t)stream, mask_rowptr, mask_rowptr ? mask_col : nullptr);
}

extern "C" int rag_topk_masked_tc_f32(const float* q, int64_t Q, const float* keys, const float* key_inv_norm,
                                      const void* keys_shadow, const float* shadow_err, int64_t N, int32_t d, int32_t k,
                                      int32_t mode, uint32_t flags, float key_scale, const int64_t* mask_rowptr,
                                      const int64_t* mask_col, int64_t idx_offset, float* out_scores, int64_t* out_idx,
                                      void* workspace, size_t workspace_bytes, rag_stream_t stream) {
  int st = check_topk_args("topk_masked_tc", q, Q, keys, N, d, k, out_scores, out_idx);
  if (st || Q == 0) return st;
  RAG_REQUIRE(mode == RAG_SIM_BF16_REFINE || mode == RAG_SIM_F16_REFINE, RAG_EUNSUPPORTED,
              "topk_masked_tc: mode %d (exact tensor-core modes only: RAG_SIM_BF16_REFINE / RAG_SIM_F16_REFINE)", mode);
  RAG_REQUIRE(keys_shadow, RAG_EINVAL, "topk_masked_tc: needs the key shadow (rag_rows_to_shadow16)");
  RAG_REQUIRE((mask_rowptr == nullptr) == (mask_col == nullptr) || mask_rowptr, RAG_EINVAL,
              "topk_masked_tc: mask_col without mask_rowptr");
  return rag::topk_tc_run(q, Q, keys, key_inv_norm, keys_shadow, shadow_err, N, d, k, mode, flags, idx_offset, out_scores,
                          out_idx, workspace, workspace_bytes, (cudaStream_t)stream, mask_rowptr,
                          mask_rowptr ? mask_col : nullptr, key_scale);
}

extern "C" size_t rag_cosine2_topk_workspace(int64_t Q, int64_t N, int32_t da, int32_t db, int32_t k) {
  if (Q <= 0 || N <= 0 || da < 1 || db < 1 || k < 1) return 256;
  const int dc = (da + 3) / 4 * 4 + (db + 3) / 4 * 4;
  return rag::align_up((size_t)Q * dc * 4, 256) + rag::align_up((size_t)N * dc * 4, 256) +
         rag::topk_f32_workspace(Q, N, dc, k);
}

extern "C" int rag_cosine2_topk_f32(const float* qa, const float* ka, int32_t da, float w_a, const float* qb,
                                    const float* kb, int32_t db, float w_b, int64_t Q, int64_t N, int32_t k,
                                    float* out_scores, int64_t* out_idx, void* workspace, size_t workspace_bytes,
                                    rag_stream_t stream) {
  RAG_REQUIRE(da >= 1 && db >= 1, RAG_EINVAL, "cosine2_topk: da=%d db=%d", da, db);
  const int da4 = (da + 3) / 4 * 4, db4 = (db + 3) / 4 * 4, dc = da4 + db4;
  int st = check_topk_args("cosine2_topk", qa, Q, ka, N, dc, k, out_scores, out_idx);
  if (st || Q == 0) return st;
  RAG_REQUIRE(qb && kb, RAG_EINVAL, "cosine2_topk: null pointer");
  RAG_REQUIRE(workspace_bytes >= rag_cosine2_topk_workspace(Q, N, da, db, k), RAG_EWORKSPACE,
              "cosine2_topk: workspace %zu < %zu bytes", workspace_bytes, rag_cosine2_topk_workspace(Q, N, da, db, k));
  RAG_REQUIRE(workspace && (reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, RAG_EALIGN,
              "cosine2_topk: workspace must be 256-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  unsigned char* w = static_cast<unsigned char*>(workspace);
  float* qc = reinterpret_cast<float*>(w);
  const size_t off_k = rag::align_up((size_t)Q * dc * 4, 256);
  float* kc = reinterpret_cast<float*>(w + off_k);
  const size_t off_ws = off_k + rag::align_up((size_t)N * dc * 4, 256);
  auto grid = [](int64_t rows) {
    int64_t b = (rows + 7) / 8, cap = (int64_t)rag::sm_count() * 16;
    return (unsigned)(b > cap ? cap : (b < 1 ? 1 : b));
  };
  // weights ride on the query side only, so any sign of w_a / w_b is exact
  rag::concat_normalized_kernel<<<grid(Q), 256, 0, s>>>(qa, da, w_a, qb, db, w_b, Q, da4, db4, qc);
  RAG_LAUNCH_OK("concat_normalized_kernel(q)");
  rag::concat_normalized_kernel<<<grid(N), 256, 0, s>>>(ka, da, 1.0f, kb, db, 1.0f, N, da4, db4, kc);
  RAG_LAUNCH_OK("concat_normalized_kernel(keys)");
  return rag::topk_f32_run(qc, Q, kc, nullptr, N, dc, k, RAG_SIM_DOT, 0, out_scores, out_idx, w + off_ws,
                           workspace_bytes - off_ws, s);
}
