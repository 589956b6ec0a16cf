// K5: bit-exact row gathers (resource_values[topk_indices], ToyGraphBase.py:70-71) and the
// fused gather + reduce-over-k (+ convex blend) the callers always do next
// (RAGraph_node/RAGraph.py:48-53, RAGraph_edge/modules/RAGraph.py:322-328).
#include "common.cuh"

namespace rag {

__device__ unsigned long long g_gather_oob = 0ull;

// One thread moves UNROLL chunks of sizeof(VecT) bytes; consecutive threads move consecutive
// chunks of the same output row, so stores are fully coalesced and loads are coalesced per row.
template <typename VecT, int UNROLL>
__global__ void __launch_bounds__(256) gather_rows_kernel(const VecT* __restrict__ table, int64_t N,
                                                          int64_t chunks_per_row,
                                                          const int64_t* __restrict__ idx, int64_t M,
                                                          int64_t owner_lo, int64_t owner_hi,
                                                          VecT* __restrict__ out) {
  const int64_t total = M * chunks_per_row;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; t0 < total; t0 += stride * UNROLL) {
    VecT v[UNROLL];
    bool live[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      int64_t t = t0 + u * stride;
      live[u] = false;
      if (t < total) {
        int64_t m = t / chunks_per_row, c = t - m * chunks_per_row;
        int64_t j = __ldg(idx + m);
        if (j < 0) j += N;                       // torch wraps negative indices once
        if (j < 0 || j >= N) {                   // torch device-asserts; we flag and write zeros
          if (c == 0) atomicAdd(&g_gather_oob, 1ull);
          v[u] = VecT{};
          live[u] = (owner_lo == 0 && owner_hi == N);
        } else if (j >= owner_lo && j < owner_hi) {
          v[u] = __ldg(table + (j - owner_lo) * chunks_per_row + c);
          live[u] = true;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
      if (live[u]) out[t0 + u * stride] = v[u];
  }
}

// One lane-group of LANES lanes owns one query row; each lane owns VPL float4 columns.
// k rows are summed in index order j = 0..k-1 (deterministic), 4 loads in flight per lane.
template <int LANES, int VPL>
__global__ void __launch_bounds__(256) gather_reduce_kernel(const float4* __restrict__ table, int64_t N,
                                                            const int64_t* __restrict__ idx, int64_t Q,
                                                            int k, int op, const float4* __restrict__ blend_in,
                                                            float blend_w, float4* __restrict__ out) {
  constexpr int F4 = LANES * VPL;
  const int sub = threadIdx.x % LANES;
  const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
  const int64_t ngroups = (int64_t)gridDim.x * blockDim.x / LANES;
  for (int64_t q = group; q < Q; q += ngroups) {
    float4 acc[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int64_t* iq = idx + q * k;
    for (int j0 = 0; j0 < k; j0 += 4) {
      float4 x[4][VPL];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        int64_t j = (j0 + u < k) ? __ldg(iq + j0 + u) : -1;
        if (j < 0 && j0 + u < k) j += N;
        bool ok = (j0 + u < k) && j >= 0 && j < N;
        if ((j0 + u < k) && !ok && sub == 0) atomicAdd(&g_gather_oob, 1ull);
#pragma unroll
        for (int v = 0; v < VPL; ++v)
          x[u][v] = ok ? __ldg(table + j * F4 + sub + v * LANES) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          acc[v].x += x[u][v].x; acc[v].y += x[u][v].y; acc[v].z += x[u][v].z; acc[v].w += x[u][v].w;
        }
    }
    const float kf = (float)k;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      float4 r = acc[v];
      if (op == RAG_REDUCE_MEAN) { r.x /= kf; r.y /= kf; r.z /= kf; r.w /= kf; }
      if (blend_in) {
        float4 b = __ldg(blend_in + q * F4 + sub + v * LANES);
        const float wb = 1.0f - blend_w;
        r.x = b.x * wb + r.x * blend_w; r.y = b.y * wb + r.y * blend_w;
        r.z = b.z * wb + r.z * blend_w; r.w = b.w * wb + r.w * blend_w;
      }
      out[q * F4 + sub + v * LANES] = r;
    }
  }
}

// any d: scalar columns, one warp per query row
__global__ void __launch_bounds__(256) gather_reduce_generic_kernel(const float* __restrict__ table, int64_t N,
                                                                    int d, const int64_t* __restrict__ idx,
                                                                    int64_t Q, int k, int op,
                                                                    const float* __restrict__ blend_in,
                                                                    float blend_w, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t q = warp; q < Q; q += nwarps) {
    for (int c = lane; c < d; c += 32) {
      float acc = 0.f;
      for (int j = 0; j < k; ++j) {
        int64_t r = __ldg(idx + q * k + j);
        if (r < 0) r += N;
        if (r < 0 || r >= N) { if (c == 0) atomicAdd(&g_gather_oob, 1ull); continue; }
        acc += __ldg(table + r * d + c);
      }
      if (op == RAG_REDUCE_MEAN) acc /= (float)k;
      if (blend_in) acc = __ldg(blend_in + q * d + c) * (1.0f - blend_w) + acc * blend_w;
      out[q * d + c] = acc;
    }
  }
}

template <typename VecT>
static int launch_gather(const void* table, int64_t N, int64_t row_bytes, const int64_t* idx, int64_t M,
                         int64_t lo, int64_t hi, void* out, cudaStream_t s) {
  constexpr int UNROLL = 4;
  const int64_t cpr = row_bytes / (int64_t)sizeof(VecT);
  const int64_t total = M * cpr;
  int64_t blocks = (total + 256 * UNROLL - 1) / (256 * UNROLL);
  const int64_t cap = (int64_t)sm_count() * 64;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  gather_rows_kernel<VecT, UNROLL><<<(unsigned)blocks, 256, 0, s>>>(
      reinterpret_cast<const VecT*>(table), N, cpr, idx, M, lo, hi, reinterpret_cast<VecT*>(out));
  RAG_LAUNCH_OK("gather_rows_kernel");
  return RAG_OK;
}

}  // namespace rag

extern "C" int rag_gather_rows(const void* table, int64_t N, int64_t row_bytes, const int64_t* idx, int64_t M,
                               int64_t owner_lo, int64_t owner_hi, void* out, rag_stream_t stream) {
  RAG_REQUIRE(N >= 0 && row_bytes >= 1 && M >= 0, RAG_EINVAL, "gather_rows: N=%lld row_bytes=%lld M=%lld",
              (long long)N, (long long)row_bytes, (long long)M);
  RAG_REQUIRE(owner_lo >= 0 && owner_hi >= owner_lo && owner_hi <= N, RAG_EINVAL,
              "gather_rows: owner range [%lld,%lld) outside [0,%lld)", (long long)owner_lo,
              (long long)owner_hi, (long long)N);
  if (M == 0) return RAG_OK;
  RAG_REQUIRE(table && idx && out, RAG_EINVAL, "gather_rows: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  const uintptr_t a = reinterpret_cast<uintptr_t>(table) | reinterpret_cast<uintptr_t>(out) | (uintptr_t)row_bytes;
  if ((a & 15u) == 0) return rag::launch_gather<uint4>(table, N, row_bytes, idx, M, owner_lo, owner_hi, out, s);
  if ((a & 7u) == 0) return rag::launch_gather<uint2>(table, N, row_bytes, idx, M, owner_lo, owner_hi, out, s);
  if ((a & 3u) == 0) return rag::launch_gather<uint32_t>(table, N, row_bytes, idx, M, owner_lo, owner_hi, out, s);
  return rag::launch_gather<uint8_t>(table, N, row_bytes, idx, M, owner_lo, owner_hi, out, s);
}

extern "C" int64_t rag_gather_oob_count(void) {
  unsigned long long v = 0;
  if (cudaMemcpyFromSymbol(&v, rag::g_gather_oob, sizeof(v)) != cudaSuccess) return -1;
  return (int64_t)v;
}

extern "C" int rag_gather_reduce_f32(const float* table, int64_t N, int32_t d, const int64_t* idx, int64_t Q,
                                     int32_t k, int32_t op, const float* blend_in, float blend_w, float* out,
                                     rag_stream_t stream) {
  RAG_REQUIRE(N >= 0 && d >= 1 && Q >= 0 && k >= 1, RAG_EINVAL, "gather_reduce: N=%lld d=%d Q=%lld k=%d",
              (long long)N, d, (long long)Q, k);
  RAG_REQUIRE(op == RAG_REDUCE_SUM || op == RAG_REDUCE_MEAN, RAG_EINVAL, "gather_reduce: op=%d", op);
  if (Q == 0) return RAG_OK;
  RAG_REQUIRE(table && idx && out, RAG_EINVAL, "gather_reduce: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  const bool vec = (d % 4 == 0) && rag::aligned16(table) && rag::aligned16(out) &&
                   (!blend_in || rag::aligned16(blend_in));
  const int64_t cap = (int64_t)rag::sm_count() * 32;
  auto grid = [&](int lanes) {
    int64_t b = (Q * lanes + 255) / 256;
    if (b > cap) b = cap;
    return (unsigned)(b < 1 ? 1 : b);
  };
#define RAG_GR(LANES, VPL)                                                                           \
  rag::gather_reduce_kernel<LANES, VPL><<<grid(LANES), 256, 0, s>>>(                                 \
      reinterpret_cast<const float4*>(table), N, idx, Q, k, op, reinterpret_cast<const float4*>(blend_in), \
      blend_w, reinterpret_cast<float4*>(out))
  if (vec && d == 512) RAG_GR(32, 4);
  else if (vec && d == 256) RAG_GR(32, 2);
  else if (vec && d == 128) RAG_GR(32, 1);
  else if (vec && d == 64) RAG_GR(16, 1);
  else if (vec && d == 32) RAG_GR(8, 1);
  else if (vec && d == 16) RAG_GR(4, 1);
  else
    rag::gather_reduce_generic_kernel<<<grid(32), 256, 0, s>>>(table, N, d, idx, Q, k, op, blend_in, blend_w, out);
#undef RAG_GR
  RAG_LAUNCH_OK("gather_reduce_kernel");
  return RAG_OK;
}
