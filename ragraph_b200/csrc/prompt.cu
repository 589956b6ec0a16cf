// Downstream-prompt pieces of the label/feature fusion:
//   prompt_act_kernel        ELU(weight * emb) / weight * emb      RAGraph_node/downprompt.py:118-130,
//                                                                   RAGraph_graph/downprompt.py:197-209
//   prototype_scores_kernel  per-row cosine against the C class prototypes + (log-)softmax, the reference's
//                            Python loop of torch.cosine_similarity(...).item() per (row, class)
//                                                                   RAGraph_node/downprompt.py:36-44,
//                                                                   RAGraph_graph/downprompt.py:41-56 (predict)
// Both are HBM-bound streaming kernels: every embedding row is read once (128-bit loads when the row pitch allows),
// the prototypes live in shared memory for the CTA's lifetime, and the [n, d] prompted copy is never written when
// the caller only needs the class scores (the prompt is applied on the fly inside prototype_scores_kernel).
#include <cfloat>
#include "common.cuh"

namespace rag {

constexpr int PR_THREADS = 256;
constexpr int PR_MAX_CLASSES = 32;

__device__ __forceinline__ float prompt_act(float v, int act) { return act == 1 ? (v > 0.f ? v : expm1f(v)) : v; }

__global__ void __launch_bounds__(PR_THREADS) prompt_act_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                 int64_t n, int d, int act, float* __restrict__ out,
                                                                 int vec4) {
  const int64_t total = n * (int64_t)d;
  if (vec4) {
    const int d4 = d >> 2;
    const int64_t total4 = total >> 2;
    for (int64_t i = (int64_t)blockIdx.x * PR_THREADS + threadIdx.x; i < total4; i += (int64_t)gridDim.x * PR_THREADS) {
      const int c = (int)(i % d4);
      float4 v = __ldcs(reinterpret_cast<const float4*>(x) + i);
      const float4 ww = __ldg(reinterpret_cast<const float4*>(w) + c);
      v.x = prompt_act(v.x * ww.x, act); v.y = prompt_act(v.y * ww.y, act);
      v.z = prompt_act(v.z * ww.z, act); v.w = prompt_act(v.w * ww.w, act);
      reinterpret_cast<float4*>(out)[i] = v;
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * PR_THREADS + threadIdx.x; i < total; i += (int64_t)gridDim.x * PR_THREADS)
      out[i] = prompt_act(__ldcs(x + i) * __ldg(w + (int)(i % d)), act);
  }
}

// One warp per embedding row.  CMAX is the compiled class bound (accumulators stay in registers).
template <int CMAX>
__global__ void __launch_bounds__(PR_THREADS) prototype_scores_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                       const float* __restrict__ proto, int64_t n, int d,
                                                                       int C, int act, float eps, int mode,
                                                                       float* __restrict__ out) {
  extern __shared__ __align__(16) float sm[];
  float* sp = sm;                 // [C, d] prototypes
  float* sw = sm + (size_t)C * d; // [d] prompt weights (ones when w == nullptr)
  float* spn = sw + d;            // [C] prototype norms
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = PR_THREADS / 32;
  for (int i = threadIdx.x; i < C * d; i += PR_THREADS) sp[i] = __ldg(proto + i);
  for (int i = threadIdx.x; i < d; i += PR_THREADS) sw[i] = w ? __ldg(w + i) : 1.f;
  __syncthreads();
  for (int c = warp; c < C; c += wpb) {
    float s = 0.f;
    for (int j = lane; j < d; j += 32) s = fmaf(sp[(size_t)c * d + j], sp[(size_t)c * d + j], s);
    s = warp_sum(s);
    if (lane == 0) spn[c] = fmaxf(sqrtf(s), eps);
  }
  __syncthreads();

  for (int64_t row = (int64_t)blockIdx.x * wpb + warp; row < n; row += (int64_t)gridDim.x * wpb) {
    const float* xr = x + row * (int64_t)d;
    float dot[CMAX];
#pragma unroll
    for (int c = 0; c < CMAX; ++c) dot[c] = 0.f;
    float nx = 0.f;
    for (int j = lane; j < d; j += 32) {
      const float v = prompt_act(__ldcs(xr + j) * sw[j], act);
      nx = fmaf(v, v, nx);
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) dot[c] = fmaf(v, sp[(size_t)c * d + j], dot[c]);
    }
    nx = fmaxf(sqrtf(warp_sum(nx)), eps);
    float mine = -FLT_MAX;        // lane c keeps the score of class c
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      if (c < C) {
        const float s = warp_sum(dot[c]);
        if (lane == c) mine = (s / nx) / spn[c];
      }
    }
    if (mode != 0) {
      float mx = mine;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      const float e = lane < C ? expf(mine - mx) : 0.f;
      const float den = warp_sum(e);
      mine = mode == 1 ? e / den : (mine - mx) - logf(den);
    }
    if (lane < C) out[row * (int64_t)C + lane] = mine;
  }
}

}  // namespace rag

extern "C" int rag_prompt_act_f32(const float* x, int64_t n, int32_t d, const float* w, int32_t act, float* out,
                                  rag_stream_t stream) {
  using namespace rag;
  RAG_REQUIRE(n >= 0 && d >= 1 && (act == 0 || act == 1), RAG_EINVAL, "prompt_act: n=%lld d=%d act=%d", (long long)n, d, act);
  if (n == 0) return RAG_OK;
  RAG_REQUIRE(x && w && out, RAG_EINVAL, "prompt_act: null pointer");
  const int vec4 = (d % 4 == 0) && aligned16(x) && aligned16(w) && aligned16(out);
  const int64_t work = (n * (int64_t)d) / (vec4 ? 4 : 1);
  const int64_t want = (work + PR_THREADS - 1) / PR_THREADS;
  const int grid = (int)(want < (int64_t)sm_count() * 8 ? (want > 0 ? want : 1) : (int64_t)sm_count() * 8);
  prompt_act_kernel<<<grid, PR_THREADS, 0, (cudaStream_t)stream>>>(x, w, n, d, act, out, vec4);
  RAG_LAUNCH_OK("prompt_act_kernel");
  return RAG_OK;
}

extern "C" int rag_prototype_scores_f32(const float* x, int64_t n, int32_t d, const float* w, int32_t act,
                                        const float* proto, int32_t C, float eps, int32_t mode, float* out,
                                        rag_stream_t stream) {
  using namespace rag;
  RAG_REQUIRE(n >= 0 && d >= 1 && C >= 1 && (act == 0 || act == 1) && mode >= 0 && mode <= 2, RAG_EINVAL,
              "prototype_scores: n=%lld d=%d C=%d act=%d mode=%d", (long long)n, d, C, act, mode);
  RAG_REQUIRE(C <= PR_MAX_CLASSES, RAG_EUNSUPPORTED, "prototype_scores: C=%d > %d classes", C, PR_MAX_CLASSES);
  if (n == 0) return RAG_OK;
  RAG_REQUIRE(x && proto && out, RAG_EINVAL, "prototype_scores: null pointer");
  const size_t smem = ((size_t)C * d + d + C) * sizeof(float);
  RAG_REQUIRE(smem <= (size_t)max_smem_optin(), RAG_EUNSUPPORTED,
              "prototype_scores: C*d = %d*%d prototypes do not fit shared memory", C, d);
  const int wpb = PR_THREADS / 32;
  const int64_t want = (n + wpb - 1) / wpb;
  const int grid = (int)(want < (int64_t)sm_count() * 8 ? want : (int64_t)sm_count() * 8);
  cudaError_t e;
#define RAG_PS_LAUNCH(CM)                                                                                              \
  do {                                                                                                                 \
    if (smem > 48 * 1024) {                                                                                            \
      e = cudaFuncSetAttribute(prototype_scores_kernel<CM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
      if (e != cudaSuccess) return cuda_fail(e, "prototype_scores: cudaFuncSetAttribute");                             \
    }                                                                                                                  \
    prototype_scores_kernel<CM><<<grid, PR_THREADS, smem, (cudaStream_t)stream>>>(x, w, proto, n, d, C, act, eps, mode, out); \
  } while (0)
  if (C <= 4) RAG_PS_LAUNCH(4);
  else if (C <= 8) RAG_PS_LAUNCH(8);
  else RAG_PS_LAUNCH(32);
#undef RAG_PS_LAUNCH
  RAG_LAUNCH_OK("prototype_scores_kernel");
  return RAG_OK;
}
