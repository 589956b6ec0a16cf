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

// exp(v) - 1 for v <= 0 without a branch (libdevice expm1f branches per element and a warp sees both signs): a degree-7
// Taylor polynomial above -1/4 (truncation < 4e-10), ex2.approx below it (|result| >= 0.22, so the absolute error of
// ~2e-7 stays under 1e-6 relative); both are computed and one is selected.
__device__ __forceinline__ float expm1_nonpos(float v) {
  const float p = v * fmaf(v, fmaf(v, fmaf(v, fmaf(v, fmaf(v, fmaf(v, 1.f / 5040.f, 1.f / 720.f), 1.f / 120.f), 1.f / 24.f),
                                          1.f / 6.f), 0.5f), 1.f);
  const float e = __expf(v) - 1.0f;
  return v > -0.25f ? p : e;
}
// act 1 = ELU (alpha 1): v > 0 ? v : exp(v) - 1.  `act` is uniform over the launch.
__device__ __forceinline__ float prompt_act(float v, int act) {
  if (act != 1) return v;
  const float neg = expm1_nonpos(fminf(v, 0.f));
  return v > 0.f ? v : neg;
}

__global__ void __launch_bounds__(PR_THREADS) prompt_act_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                 int64_t n, int d, int act, float* __restrict__ out,
                                                                 int vec4) {
  const int64_t total = n * (int64_t)d;
  // the column of element i is i mod d; a 64-bit modulo per element costs more than the load it indexes, so every
  // thread keeps a running column that advances by (grid stride mod d) per iteration
  if (vec4) {
    const int d4 = d >> 2;
    const int64_t total4 = total >> 2;
    const int64_t stride = (int64_t)gridDim.x * PR_THREADS;
    const int step = (int)(stride % d4);
    int64_t i = (int64_t)blockIdx.x * PR_THREADS + threadIdx.x;
    int c = (int)(i % d4);
    for (; i < total4; i += stride) {
      float4 v = __ldcs(reinterpret_cast<const float4*>(x) + i);
      const float4 ww = __ldg(reinterpret_cast<const float4*>(w) + c);
      v.x = prompt_act(v.x * ww.x, act); v.y = prompt_act(v.y * ww.y, act);
      v.z = prompt_act(v.z * ww.z, act); v.w = prompt_act(v.w * ww.w, act);
      __stcs(reinterpret_cast<float4*>(out) + i, v);
      c += step;
      if (c >= d4) c -= d4;
    }
  } else {
    const int64_t stride = (int64_t)gridDim.x * PR_THREADS;
    const int step = (int)(stride % d);
    int64_t i = (int64_t)blockIdx.x * PR_THREADS + threadIdx.x;
    int c = (int)(i % d);
    for (; i < total; i += stride) {
      out[i] = prompt_act(__ldcs(x + i) * __ldg(w + c), act);
      c += step;
      if (c >= d) c -= d;
    }
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

// Vector path (d % 4 == 0, C <= 8): 8 lanes per row, every 8-lane group works on TWO rows at a time so that each
// prototype float4 read from shared memory feeds 8 FMAs (shared-memory traffic per row = C/2 KB against 1 KB of HBM at
// d = 256, inside the 128 B/clk port at full HBM rate); the 8-lane reductions are 3 shuffle steps per value.  A warp
// covers 8 rows per iteration with 128-bit loads: 4 rows x 128 contiguous bytes per instruction.
template <int CMAX>
__global__ void __launch_bounds__(PR_THREADS) prototype_scores_vec_kernel(const float4* __restrict__ x, const float* __restrict__ w,
                                                                           const float* __restrict__ proto, int64_t n, int d4,
                                                                           int C, int act, float eps, int mode,
                                                                           float* __restrict__ out) {
  extern __shared__ __align__(16) float sm[];
  float4* sp = reinterpret_cast<float4*>(sm);            // [C, d4]
  float4* sw = sp + (size_t)C * d4;                      // [d4]
  float* spn = reinterpret_cast<float*>(sw + d4);        // [C]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = PR_THREADS / 32;
  const int d = d4 * 4;
  for (int i = threadIdx.x; i < C * d; i += PR_THREADS) sm[i] = __ldg(proto + i);
  for (int i = threadIdx.x; i < d; i += PR_THREADS) reinterpret_cast<float*>(sw)[i] = w ? __ldg(w + i) : 1.f;
  __syncthreads();
  for (int c = warp; c < C; c += wpb) {
    float s = 0.f;
    for (int j = lane; j < d; j += 32) s = fmaf(sm[(size_t)c * d + j], sm[(size_t)c * d + j], s);
    s = warp_sum(s);
    if (lane == 0) spn[c] = 1.0f / fmaxf(sqrtf(s), eps);       // reciprocal of the clamped prototype norm
  }
  __syncthreads();

  const int sub = lane & 7, grp = lane >> 3;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t base = ((int64_t)blockIdx.x * wpb + warp) * 8; base < n; base += (int64_t)gridDim.x * wpb * 8) {
    const int64_t r0 = base + 2 * grp, r1 = r0 + 1;
    const bool v0 = r0 < n, v1 = r1 < n;
    const float4* x0 = x + r0 * d4;
    const float4* x1 = x + r1 * d4;
    float dot0[CMAX], dot1[CMAX];
#pragma unroll
    for (int c = 0; c < CMAX; ++c) { dot0[c] = 0.f; dot1[c] = 0.f; }
    float n0 = 0.f, n1 = 0.f;
#pragma unroll 2
    for (int j = sub; j < d4; j += 8) {
      float4 a = v0 ? __ldcs(x0 + j) : zero;
      float4 b = v1 ? __ldcs(x1 + j) : zero;
      const float4 ww = sw[j];
      a.x = prompt_act(a.x * ww.x, act); a.y = prompt_act(a.y * ww.y, act);
      a.z = prompt_act(a.z * ww.z, act); a.w = prompt_act(a.w * ww.w, act);
      b.x = prompt_act(b.x * ww.x, act); b.y = prompt_act(b.y * ww.y, act);
      b.z = prompt_act(b.z * ww.z, act); b.w = prompt_act(b.w * ww.w, act);
      n0 = fmaf(a.x, a.x, fmaf(a.y, a.y, fmaf(a.z, a.z, fmaf(a.w, a.w, n0))));
      n1 = fmaf(b.x, b.x, fmaf(b.y, b.y, fmaf(b.z, b.z, fmaf(b.w, b.w, n1))));
#pragma unroll
      for (int c = 0; c < CMAX; ++c) {
        if (c < C) {
          const float4 p = sp[(size_t)c * d4 + j];
          dot0[c] = fmaf(a.x, p.x, fmaf(a.y, p.y, fmaf(a.z, p.z, fmaf(a.w, p.w, dot0[c]))));
          dot1[c] = fmaf(b.x, p.x, fmaf(b.y, p.y, fmaf(b.z, p.z, fmaf(b.w, p.w, dot1[c]))));
        }
      }
    }
    // 8-lane group reductions: afterwards every lane of the group holds the totals of its two rows
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      n0 += __shfl_xor_sync(0xffffffffu, n0, o);
      n1 += __shfl_xor_sync(0xffffffffu, n1, o);
#pragma unroll
      for (int c = 0; c < CMAX; ++c) {            // unconditional: classes >= C hold zeros (no predicated collectives)
        dot0[c] += __shfl_xor_sync(0xffffffffu, dot0[c], o);
        dot1[c] += __shfl_xor_sync(0xffffffffu, dot1[c], o);
      }
    }
    // lane `sub` of the group owns class `sub`: one scale, then max / sum over the group's 8 lanes
    const float inv0 = 1.0f / fmaxf(sqrtf(n0), eps), inv1 = 1.0f / fmaxf(sqrtf(n1), eps);
    float mine0 = -FLT_MAX, mine1 = -FLT_MAX;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      if (c < C && c == sub) { mine0 = dot0[c] * inv0 * spn[c]; mine1 = dot1[c] * inv1 * spn[c]; }
    }
    if (mode != 0) {
      float m0 = mine0, m1 = mine1;
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, o));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, o));
      }
      const float e0 = sub < C ? expf(mine0 - m0) : 0.f, e1 = sub < C ? expf(mine1 - m1) : 0.f;
      float den0 = e0, den1 = e1;
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        den0 += __shfl_xor_sync(0xffffffffu, den0, o);
        den1 += __shfl_xor_sync(0xffffffffu, den1, o);
      }
      mine0 = mode == 1 ? e0 / den0 : (mine0 - m0) - logf(den0);
      mine1 = mode == 1 ? e1 / den1 : (mine1 - m1) - logf(den1);
    }
    if (sub < C) {
      if (v0) out[r0 * (int64_t)C + sub] = mine0;
      if (v1) out[r1 * (int64_t)C + sub] = mine1;
    }
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
  cudaError_t e;
  if (C <= 8 && d % 4 == 0 && aligned16(x)) {
    const int64_t want = (n + wpb * 8 - 1) / (wpb * 8);
    // persistent grid = exactly the CTAs that are resident at once (a partial second wave would run at a fraction of
    // the machine for as long as the first)
#define RAG_PSV_LAUNCH(CM)                                                                                                 \
  do {                                                                                                                     \
    if (smem > 48 * 1024) {                                                                                                \
      e = cudaFuncSetAttribute(prototype_scores_vec_kernel<CM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
      if (e != cudaSuccess) return cuda_fail(e, "prototype_scores: cudaFuncSetAttribute");                                 \
    }                                                                                                                      \
    int per_sm = 0;                                                                                                        \
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, prototype_scores_vec_kernel<CM>, PR_THREADS, smem);         \
    if (e != cudaSuccess || per_sm < 1) per_sm = 1;                                                                        \
    const int64_t cap = (int64_t)sm_count() * per_sm;                                                                      \
    const int grid = (int)(want < cap ? want : cap);                                                                       \
    prototype_scores_vec_kernel<CM><<<grid, PR_THREADS, smem, (cudaStream_t)stream>>>(                                     \
        reinterpret_cast<const float4*>(x), w, proto, n, d / 4, C, act, eps, mode, out);                                   \
  } while (0)
    if (C <= 4) RAG_PSV_LAUNCH(4);
    else RAG_PSV_LAUNCH(8);
#undef RAG_PSV_LAUNCH
    RAG_LAUNCH_OK("prototype_scores_vec_kernel");
    return RAG_OK;
  }
  const int64_t want = (n + wpb - 1) / wpb;
  const int grid = (int)(want < (int64_t)sm_count() * 8 ? want : (int64_t)sm_count() * 8);
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
