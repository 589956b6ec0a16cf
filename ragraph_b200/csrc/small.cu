// Small-problem retrieve in ONE launch: normalise + scores + top-k + value/label gathers.
//
// The reference's real call sites are tiny next to the benchmark shapes: the graph variant retrieves for ONE pooled query
// against a library of a few hundred rows (RAGraph_graph/ragraph_utils/ToyGraphBase.py:56-87, k = 3), the few-shot and
// noise branches a handful of rows.  There the reference's six torch launches (two normalisations, matmul, topk, two
// index gathers) are pure launch latency, and so was this library's own four-launch sequence.  This kernel does the
// whole of ToyGraphBase.retrieve for Q <= 64 queries and N <= 65 536 keys:
//   * every CTA stages the L2-normalised queries in shared memory (fp32, the same q * inv_norm the fp32 kernel forms);
//   * each warp walks a contiguous key range: the lanes read one key row (128-bit loads), form its norm on the fly when no
//     inverse norms are given, reduce the Q dot products across the warp, and keep a sorted top-k list per query in shared
//     memory (score desc, index asc -- the library's deterministic order);
//   * warp lists go to a global scratch area; the LAST CTA to finish (atomic ticket + __threadfence) merges them per query,
//     writes scores / indices and copies the winners' value and label rows (bit exact) -- then re-arms the ticket, so the
//     scratch area needs zeroing only once, when the caller allocates it.
// fp32 FMA throughout: results match the fp32 kernel within summation-order rounding (<= 1e-6).
#include <cfloat>
#include "common.cuh"

namespace rag {

constexpr int SM_THREADS = 256;
constexpr int SM_WARPS = SM_THREADS / 32;
constexpr int SM_MAX_Q = 64;
constexpr int SM_MAX_K = 16;
constexpr int SM_MAX_QD = 16384;            // floats of staged queries (64 KB)
constexpr int64_t SM_MAX_N = 65536;

struct SmallArgs {
  const float* q; int Q; const float* keys; const float* key_inv_norm; int N; int d; int k; int dot;
  const unsigned char* values; int64_t vbytes; const unsigned char* labels; int64_t lbytes;
  float* out_scores; int64_t* out_idx; unsigned char* out_values; unsigned char* out_labels;
  unsigned int* ticket; float* part_s; int32_t* part_i;      // scratch: [n_lists][Q][k]
  int keys_per_warp;
};

__device__ __forceinline__ void copy_row(unsigned char* dst, const unsigned char* src, int64_t nbytes, int lane) {
  if (((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src) | (uintptr_t)nbytes) & 15u) == 0) {
    for (int64_t o = (int64_t)lane * 16; o < nbytes; o += 512) *reinterpret_cast<uint4*>(dst + o) = __ldg(reinterpret_cast<const uint4*>(src + o));
  } else {
    for (int64_t o = (int64_t)lane * 4; o < nbytes; o += 128) *reinterpret_cast<uint32_t*>(dst + o) = __ldg(reinterpret_cast<const uint32_t*>(src + o));
  }
}

__global__ void __launch_bounds__(SM_THREADS) retrieve_small_kernel(const SmallArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sq = reinterpret_cast<float*>(smem_raw);                                   // [Q][d] normalised queries
  float* ls = sq + (size_t)a.Q * a.d;                                               // [warps][Q][k]
  int32_t* li = reinterpret_cast<int32_t*>(ls + (size_t)SM_WARPS * a.Q * a.k);      // [warps][Q][k]
  __shared__ unsigned int s_last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Q = a.Q, d = a.d, k = a.k;

  // ---- stage the normalised queries -----------------------------------------------------------------------------
  for (int r = warp; r < Q; r += SM_WARPS) {
    float ss = 0.f;
    for (int c = lane; c < d; c += 32) { const float v = __ldg(a.q + (size_t)r * d + c); ss = fmaf(v, v, ss); }
    const float inv = a.dot ? 1.0f : 1.0f / fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
    for (int c = lane; c < d; c += 32) sq[(size_t)r * d + c] = __ldg(a.q + (size_t)r * d + c) * inv;
  }
  float* my_s = ls + (size_t)warp * Q * k;
  int32_t* my_i = li + (size_t)warp * Q * k;
  for (int p = lane; p < Q * k; p += 32) { my_s[p] = -FLT_MAX; my_i[p] = INT32_MAX; }
  __syncthreads();

  // ---- this warp's key range ---------------------------------------------------------------------------------------
  const int gw = blockIdx.x * SM_WARPS + warp;
  const int j0 = gw * a.keys_per_warp, j1 = min(a.N, j0 + a.keys_per_warp);
  const bool vec = (d & 127) == 0;                                  // every lane owns whole float4 columns
  for (int j = j0; j < j1; ++j) {
    const float* kr = a.keys + (size_t)j * d;
    float kinv = 1.0f;
    if (!a.dot) {
      if (a.key_inv_norm) kinv = __ldg(a.key_inv_norm + j);
      else {
        float ss = 0.f;
        for (int c = lane; c < d; c += 32) { const float v = __ldg(kr + c); ss = fmaf(v, v, ss); }
        kinv = 1.0f / fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
      }
    }
    for (int r = 0; r < Q; ++r) {
      const float* qr = sq + (size_t)r * d;
      float dot = 0.f;
      if (vec) {
        for (int c = lane * 4; c < d; c += 128) {
          const float4 kv = __ldg(reinterpret_cast<const float4*>(kr + c));
          const float4 qv = *reinterpret_cast<const float4*>(qr + c);
          dot = fmaf(kv.x, qv.x, dot); dot = fmaf(kv.y, qv.y, dot); dot = fmaf(kv.z, qv.z, dot); dot = fmaf(kv.w, qv.w, dot);
        }
      } else {
        for (int c = lane; c < d; c += 32) dot = fmaf(__ldg(kr + c), qr[c], dot);
      }
      const float s = warp_sum(dot) * kinv;
      float* rs = my_s + r * k;
      int32_t* ri = my_i + r * k;
      if (ranks_before(s, (int64_t)j, rs[k - 1], (int64_t)ri[k - 1])) warp_sorted_insert<int32_t>(rs, ri, k, s, j, lane);
    }
  }
  __syncwarp();
  // ---- publish this warp's lists, take a ticket ----------------------------------------------------------------------
  for (int p = lane; p < Q * k; p += 32) {
    a.part_s[(size_t)gw * Q * k + p] = my_s[p];
    a.part_i[(size_t)gw * Q * k + p] = my_i[p];
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();

  // ---- last CTA: merge per query, write results, gather the winners' rows --------------------------------------------
  const int n_lists = gridDim.x * SM_WARPS;
  int64_t* fi = reinterpret_cast<int64_t*>(smem_raw);                 // [warps][k] (the query staging area is free now)
  float* fv = reinterpret_cast<float*>(fi + (size_t)SM_WARPS * k);    // [warps][k]
  for (int r = warp; r < Q; r += SM_WARPS) {
    int64_t* mi = fi + (size_t)warp * k;
    float* mv = fv + (size_t)warp * k;
    for (int p = lane; p < k; p += 32) { mv[p] = -FLT_MAX; mi[p] = INT64_MAX; }
    __syncwarp();
    const int total = n_lists * k;
    for (int c0 = 0; c0 < total; c0 += 32) {
      const int c = c0 + lane;
      float s = -FLT_MAX; int64_t j = -1;
      if (c < total) {
        const int l = c / k, p = c - l * k;
        const size_t o = ((size_t)l * Q + r) * k + p;
        s = __ldcg(a.part_s + o);
        const int32_t jj = __ldcg(a.part_i + o);
        j = (jj == INT32_MAX) ? -1 : (int64_t)jj;
      }
      unsigned m = __ballot_sync(0xffffffffu, j >= 0 && ranks_before(s, j, mv[k - 1], mi[k - 1]));
      while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const float sv = __shfl_sync(0xffffffffu, s, src);
        const int64_t jv = __shfl_sync(0xffffffffu, j, src);
        if (ranks_before(sv, jv, mv[k - 1], mi[k - 1])) warp_sorted_insert<int64_t>(mv, mi, k, sv, jv, lane);
      }
    }
    for (int p = lane; p < k; p += 32) {
      a.out_scores[(size_t)r * k + p] = mv[p];
      a.out_idx[(size_t)r * k + p] = (mi[p] == INT64_MAX) ? (int64_t)-1 : mi[p];
    }
    for (int p = 0; p < k; ++p) {
      const int64_t g = mi[p];
      if (g < 0 || g == INT64_MAX) continue;
      if (a.values) copy_row(a.out_values + ((size_t)r * k + p) * a.vbytes, a.values + g * a.vbytes, a.vbytes, lane);
      if (a.labels) copy_row(a.out_labels + ((size_t)r * k + p) * a.lbytes, a.labels + g * a.lbytes, a.lbytes, lane);
    }
    __syncwarp();
  }
  if (threadIdx.x == 0) *a.ticket = 0u;                               // re-armed for the next call on this stream
}

static int small_plan(int64_t N, int* keys_per_warp, int* ctas) {
  int64_t warps = (N + 31) / 32;                                      // >= 32 keys per warp
  const int64_t cap = (int64_t)sm_count() * SM_WARPS * 2;
  if (warps > cap) warps = cap;
  if (warps < 1) warps = 1;
  int kpw = (int)((N + warps - 1) / warps);
  if (kpw < 1) kpw = 1;
  warps = (N + kpw - 1) / kpw;
  *keys_per_warp = kpw;
  *ctas = (int)((warps + SM_WARPS - 1) / SM_WARPS);
  return RAG_OK;
}

}  // namespace rag

extern "C" int rag_retrieve_small_supported(int64_t Q, int64_t N, int32_t d, int32_t k) {
  using namespace rag;
  return (Q >= 1 && Q <= SM_MAX_Q && N >= 1 && N <= SM_MAX_N && d >= 1 && k >= 1 && k <= SM_MAX_K && k <= N &&
          Q * (int64_t)d <= SM_MAX_QD) ? 1 : 0;
}

extern "C" size_t rag_retrieve_small_workspace(int64_t Q, int64_t N, int32_t d, int32_t k) {
  if (!rag_retrieve_small_supported(Q, N, d, k)) return 256;
  int kpw, ctas;
  rag::small_plan(N, &kpw, &ctas);
  return 256 + 2 * rag::align_up((size_t)ctas * rag::SM_WARPS * Q * k * 4, 256);
}

extern "C" int rag_retrieve_small_f32(const float* q, int64_t Q, const float* keys, const float* key_inv_norm, int64_t N,
                                      int32_t d, int32_t k, uint32_t flags, const void* values, int64_t value_row_bytes,
                                      const void* labels, int64_t label_row_bytes, float* out_scores, int64_t* out_idx,
                                      void* out_values, void* out_labels, void* workspace, size_t workspace_bytes,
                                      rag_stream_t stream) {
  using namespace rag;
  RAG_REQUIRE(rag_retrieve_small_supported(Q, N, d, k), RAG_EUNSUPPORTED,
              "retrieve_small: Q=%lld N=%lld d=%d k=%d outside Q <= %d, N <= %lld, k <= %d, Q*d <= %d", (long long)Q,
              (long long)N, d, k, SM_MAX_Q, (long long)SM_MAX_N, SM_MAX_K, SM_MAX_QD);
  RAG_REQUIRE(q && keys && out_scores && out_idx, RAG_EINVAL, "retrieve_small: null pointer");
  RAG_REQUIRE(aligned16(q) && aligned16(keys), RAG_EALIGN, "retrieve_small: q/keys must be 16-byte aligned");
  RAG_REQUIRE((values == nullptr) == (out_values == nullptr) && (labels == nullptr) == (out_labels == nullptr), RAG_EINVAL,
              "retrieve_small: a table and its output go together");
  RAG_REQUIRE(value_row_bytes % 4 == 0 && label_row_bytes % 4 == 0 && value_row_bytes >= 0 && label_row_bytes >= 0,
              RAG_EUNSUPPORTED, "retrieve_small: row sizes must be multiples of 4 bytes");
  RAG_REQUIRE(workspace_bytes >= rag_retrieve_small_workspace(Q, N, d, k), RAG_EWORKSPACE, "retrieve_small: workspace %zu < %zu bytes",
              workspace_bytes, rag_retrieve_small_workspace(Q, N, d, k));
  RAG_REQUIRE(workspace && (reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, RAG_EALIGN, "retrieve_small: workspace must be 256-byte aligned");
  SmallArgs a{};
  int ctas;
  small_plan(N, &a.keys_per_warp, &ctas);
  a.q = q; a.Q = (int)Q; a.keys = keys; a.key_inv_norm = key_inv_norm; a.N = (int)N; a.d = d; a.k = k;
  a.dot = (flags & RAG_SIM_DOT) ? 1 : 0;
  a.values = static_cast<const unsigned char*>(values); a.vbytes = value_row_bytes;
  a.labels = static_cast<const unsigned char*>(labels); a.lbytes = label_row_bytes;
  a.out_scores = out_scores; a.out_idx = out_idx;
  a.out_values = static_cast<unsigned char*>(out_values); a.out_labels = static_cast<unsigned char*>(out_labels);
  unsigned char* w = static_cast<unsigned char*>(workspace);
  a.ticket = reinterpret_cast<unsigned int*>(w);
  const size_t part = align_up((size_t)ctas * SM_WARPS * Q * k * 4, 256);
  a.part_s = reinterpret_cast<float*>(w + 256);
  a.part_i = reinterpret_cast<int32_t*>(w + 256 + part);
  size_t smem = (size_t)Q * d * 4 + (size_t)SM_WARPS * Q * k * 8;
  const size_t merge = (size_t)SM_WARPS * k * 12;
  if (smem < merge) smem = merge;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(retrieve_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(retrieve_small_kernel)");
  }
  retrieve_small_kernel<<<(unsigned)ctas, SM_THREADS, smem, (cudaStream_t)stream>>>(a);
  RAG_LAUNCH_OK("retrieve_small_kernel");
  return RAG_OK;
}
