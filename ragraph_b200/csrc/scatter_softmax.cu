// Per-destination softmax over edge scalars: the edge variant's relative time
// encoding, torch_scatter.scatter_softmax(edge_times, dst, dim_size) in RAGraph_edge/modules/RAGraph.py:250-263,
// fused with the min-max rescale in front of it (:254-257) and the convex mix with the bi-normalised edge weight
// behind it (:267, edge_norm * 1/2 + time_norm * 1/2).
//
//   out[e] = mix_a * base[e] + mix_b * exp(t[e] - max_g) / sum_{e' in g} exp(t[e'] - max_g),  g = index[e],
//   t[e]   = (src[e] - lo) / span          (lo = min(src), span = max_step - lo as in :257; identity with lo = 0, span = 1;
//                                           range_dev = device {min, max_step} overrides both: no host sync for the min/max)
//
// Works in COO order like the reference (the output is aligned with the caller's edge list, so it can be handed to
// the CSR builder as the edge weight).  Three streaming passes over E scalars -- group max (ordered-int atomicMax),
// group sum of exponentials (atomicAdd; the fp32 sum order is not fixed, exactly like torch_scatter's own kernel),
// normalise + mix -- with the two per-group arrays (n_groups * 8 bytes) living in L2.
#include <cfloat>
#include "common.cuh"

namespace rag {

constexpr int SS_THREADS = 256;

// order-preserving float <-> int map so that atomicMax on int orders floats
__device__ __forceinline__ int float_to_ordered(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void __launch_bounds__(SS_THREADS) ss_init_kernel(int* gmax, float* gsum, int64_t n) {
  const int lowest = float_to_ordered(-FLT_MAX);
  for (int64_t i = (int64_t)blockIdx.x * SS_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * SS_THREADS) {
    gmax[i] = lowest;
    gsum[i] = 0.f;
  }
}

__global__ void __launch_bounds__(SS_THREADS) ss_max_kernel(const float* __restrict__ src, const int64_t* __restrict__ index,
                                                             int64_t E, int64_t n, float lo, float span, const float* __restrict__ range_dev, int* gmax) {
  if (range_dev) { lo = __ldg(range_dev); span = __ldg(range_dev + 1) - lo; }
  for (int64_t e = (int64_t)blockIdx.x * SS_THREADS + threadIdx.x; e < E; e += (int64_t)gridDim.x * SS_THREADS) {
    const int64_t g = __ldg(index + e);
    if (g < 0 || g >= n) continue;
    atomicMax(gmax + g, float_to_ordered((__ldg(src + e) - lo) / span));
  }
}

__global__ void __launch_bounds__(SS_THREADS) ss_sum_kernel(const float* __restrict__ src, const int64_t* __restrict__ index,
                                                             int64_t E, int64_t n, float lo, float span,
                                                             const float* __restrict__ range_dev,
                                                             const int* __restrict__ gmax, float* gsum) {
  if (range_dev) { lo = __ldg(range_dev); span = __ldg(range_dev + 1) - lo; }
  for (int64_t e = (int64_t)blockIdx.x * SS_THREADS + threadIdx.x; e < E; e += (int64_t)gridDim.x * SS_THREADS) {
    const int64_t g = __ldg(index + e);
    if (g < 0 || g >= n) continue;
    atomicAdd(gsum + g, expf((__ldg(src + e) - lo) / span - ordered_to_float(__ldg(gmax + g))));
  }
}

__global__ void __launch_bounds__(SS_THREADS) ss_norm_kernel(const float* __restrict__ src, const int64_t* __restrict__ index,
                                                              int64_t E, int64_t n, float lo, float span,
                                                              const float* __restrict__ range_dev,
                                                              const int* __restrict__ gmax, const float* __restrict__ gsum,
                                                              const float* __restrict__ base, float mix_a, float mix_b,
                                                              float* __restrict__ out) {
  if (range_dev) { lo = __ldg(range_dev); span = __ldg(range_dev + 1) - lo; }
  for (int64_t e = (int64_t)blockIdx.x * SS_THREADS + threadIdx.x; e < E; e += (int64_t)gridDim.x * SS_THREADS) {
    const int64_t g = __ldg(index + e);
    float sm = 0.f;
    if (g >= 0 && g < n) sm = expf((__ldg(src + e) - lo) / span - ordered_to_float(__ldg(gmax + g))) / __ldg(gsum + g);
    out[e] = base ? fmaf(mix_a, __ldg(base + e), mix_b * sm) : mix_b * sm;
  }
}

static int grid_for(int64_t work) {
  const int64_t want = (work + SS_THREADS - 1) / SS_THREADS;
  const int64_t cap = (int64_t)sm_count() * 8;
  return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace rag

extern "C" size_t rag_scatter_softmax_workspace(int64_t n_groups) {
  return n_groups > 0 ? rag::align_up((size_t)n_groups * 4, 256) * 2 : 256;
}

extern "C" int rag_scatter_softmax_f32(const float* src, const int64_t* index, int64_t E, int64_t n_groups, float lo,
                                       float span, const float* range_dev, const float* base, float mix_a, float mix_b, float* out,
                                       void* workspace, size_t workspace_bytes, rag_stream_t stream) {
  using namespace rag;
  RAG_REQUIRE(E >= 0 && n_groups >= 0, RAG_EINVAL, "scatter_softmax: E=%lld n_groups=%lld", (long long)E, (long long)n_groups);
  if (E == 0) return RAG_OK;
  RAG_REQUIRE(src && index && out && workspace, RAG_EINVAL, "scatter_softmax: null pointer");
  RAG_REQUIRE(n_groups >= 1, RAG_EINVAL, "scatter_softmax: %lld edges but no groups", (long long)E);
  RAG_REQUIRE(workspace_bytes >= rag_scatter_softmax_workspace(n_groups), RAG_EWORKSPACE,
              "scatter_softmax: workspace %zu < %zu bytes", workspace_bytes, rag_scatter_softmax_workspace(n_groups));
  RAG_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 3u) == 0, RAG_EALIGN, "scatter_softmax: workspace not 4-byte aligned");
  int* gmax = static_cast<int*>(workspace);
  float* gsum = reinterpret_cast<float*>(static_cast<unsigned char*>(workspace) + align_up((size_t)n_groups * 4, 256));
  cudaStream_t s = (cudaStream_t)stream;
  ss_init_kernel<<<grid_for(n_groups), SS_THREADS, 0, s>>>(gmax, gsum, n_groups);
  RAG_LAUNCH_OK("ss_init_kernel");
  ss_max_kernel<<<grid_for(E), SS_THREADS, 0, s>>>(src, index, E, n_groups, lo, span, range_dev, gmax);
  RAG_LAUNCH_OK("ss_max_kernel");
  ss_sum_kernel<<<grid_for(E), SS_THREADS, 0, s>>>(src, index, E, n_groups, lo, span, range_dev, gmax, gsum);
  RAG_LAUNCH_OK("ss_sum_kernel");
  ss_norm_kernel<<<grid_for(E), SS_THREADS, 0, s>>>(src, index, E, n_groups, lo, span, range_dev, gmax, gsum, base, mix_a, mix_b, out);
  RAG_LAUNCH_OK("ss_norm_kernel");
  return RAG_OK;
}
