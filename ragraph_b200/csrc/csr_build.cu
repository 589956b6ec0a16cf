// CSR construction on the device: from the COO edge list of the edge variant
// (RAGraph_edge/modules/RAGraph.py:22-24: edges[E,2], [:,0]=src, [:,1]=dst) and from the dense
// block-diagonal adjacency the node/graph variants pass around
// (RAGraph_node/ragraph_utils/utility.py:66-69 -> Propagation.py / layers/gcn.py).
#include "common.cuh"

namespace rag {

__global__ void coo_count_kernel(const int64_t* __restrict__ edges, int64_t E, int64_t n_rows,
                                 int32_t* __restrict__ counts) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t dst = __ldg(edges + 2 * e + 1);
    if (dst >= 0 && dst < n_rows) atomicAdd(counts + dst, 1);
  }
}

__global__ void coo_fill_kernel(const int64_t* __restrict__ edges, const float* __restrict__ w, int64_t E,
                                int64_t n_rows, const int64_t* __restrict__ rowptr, int32_t* __restrict__ cursor,
                                int32_t* __restrict__ col, float* __restrict__ val) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
    const longlong2 sd = __ldg(reinterpret_cast<const longlong2*>(edges) + e);
    if (sd.y < 0 || sd.y >= n_rows) continue;
    const int64_t slot = __ldg(rowptr + sd.y) + atomicAdd(cursor + sd.y, 1);
    col[slot] = (int32_t)sd.x;
    if (val) val[slot] = w ? __ldg(w + e) : 1.0f;
  }
}

// one warp per dense row: ballot-compacted scan keeps column order (deterministic)
__global__ void dense_count_kernel(const float* __restrict__ adj, int64_t n_rows, int64_t n_cols,
                                   int32_t* __restrict__ counts) {
  const int lane = threadIdx.x & 31;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < n_rows; r += nwarps) {
    int c = 0;
    for (int64_t j = lane; j < n_cols; j += 32) c += (__ldg(adj + r * n_cols + j) != 0.f);
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) counts[r] = c;
  }
}

__global__ void dense_fill_kernel(const float* __restrict__ adj, int64_t n_rows, int64_t n_cols,
                                  const int64_t* __restrict__ rowptr, int32_t* __restrict__ col,
                                  float* __restrict__ val) {
  const int lane = threadIdx.x & 31;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < n_rows; r += nwarps) {
    int64_t base = __ldg(rowptr + r);
    for (int64_t j0 = 0; j0 < n_cols; j0 += 32) {
      const int64_t j = j0 + lane;
      const float a = j < n_cols ? __ldg(adj + r * n_cols + j) : 0.f;
      const unsigned m = __ballot_sync(0xffffffffu, a != 0.f);
      if (a != 0.f) {
        const int64_t slot = base + __popc(m & ((1u << lane) - 1u));
        col[slot] = (int32_t)j;
        val[slot] = a;
      }
      base += __popc(m);
    }
  }
}

static unsigned grid_for(int64_t work_items, int per_block) {
  int64_t b = (work_items + per_block - 1) / per_block;
  const int64_t cap = (int64_t)sm_count() * 32;
  if (b > cap) b = cap;
  return (unsigned)(b < 1 ? 1 : b);
}
}  // namespace rag

extern "C" int rag_coo_count_rows(const int64_t* edges, int64_t E, int64_t n_rows, int32_t* counts,
                                  rag_stream_t stream) {
  RAG_REQUIRE(E >= 0 && n_rows >= 0, RAG_EINVAL, "coo_count_rows: E=%lld n_rows=%lld", (long long)E, (long long)n_rows);
  if (n_rows == 0) return RAG_OK;
  RAG_REQUIRE(counts && (E == 0 || edges), RAG_EINVAL, "coo_count_rows: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(counts, 0, (size_t)n_rows * 4, s);
  if (e != cudaSuccess) return rag::cuda_fail(e, "cudaMemsetAsync(counts)");
  if (E == 0) return RAG_OK;
  rag::coo_count_kernel<<<rag::grid_for(E, 256), 256, 0, s>>>(edges, E, n_rows, counts);
  RAG_LAUNCH_OK("coo_count_kernel");
  return RAG_OK;
}

extern "C" int rag_coo_fill_csr(const int64_t* edges, const float* w, int64_t E, int64_t n_rows,
                                const int64_t* rowptr, int32_t* cursor, int32_t* col, float* val,
                                rag_stream_t stream) {
  RAG_REQUIRE(E >= 0 && n_rows >= 0, RAG_EINVAL, "coo_fill_csr: E=%lld n_rows=%lld", (long long)E, (long long)n_rows);
  if (E == 0 || n_rows == 0) return RAG_OK;
  RAG_REQUIRE(edges && rowptr && cursor && col, RAG_EINVAL, "coo_fill_csr: null pointer");
  RAG_REQUIRE(rag::aligned16(edges), RAG_EALIGN, "coo_fill_csr: edges must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(cursor, 0, (size_t)n_rows * 4, s);
  if (e != cudaSuccess) return rag::cuda_fail(e, "cudaMemsetAsync(cursor)");
  rag::coo_fill_kernel<<<rag::grid_for(E, 256), 256, 0, s>>>(edges, w, E, n_rows, rowptr, cursor, col, val);
  RAG_LAUNCH_OK("coo_fill_kernel");
  return RAG_OK;
}

extern "C" int rag_dense_count_rows(const float* adj, int64_t n_rows, int64_t n_cols, int32_t* counts,
                                    rag_stream_t stream) {
  RAG_REQUIRE(n_rows >= 0 && n_cols >= 0, RAG_EINVAL, "dense_count_rows: %lld x %lld", (long long)n_rows, (long long)n_cols);
  if (n_rows == 0) return RAG_OK;
  RAG_REQUIRE(adj && counts, RAG_EINVAL, "dense_count_rows: null pointer");
  rag::dense_count_kernel<<<rag::grid_for(n_rows, 8), 256, 0, (cudaStream_t)stream>>>(adj, n_rows, n_cols, counts);
  RAG_LAUNCH_OK("dense_count_kernel");
  return RAG_OK;
}

extern "C" int rag_dense_fill_csr(const float* adj, int64_t n_rows, int64_t n_cols, const int64_t* rowptr,
                                  int32_t* col, float* val, rag_stream_t stream) {
  RAG_REQUIRE(n_rows >= 0 && n_cols >= 0, RAG_EINVAL, "dense_fill_csr: %lld x %lld", (long long)n_rows, (long long)n_cols);
  if (n_rows == 0 || n_cols == 0) return RAG_OK;
  RAG_REQUIRE(adj && rowptr && col && val, RAG_EINVAL, "dense_fill_csr: null pointer");
  rag::dense_fill_kernel<<<rag::grid_for(n_rows, 8), 256, 0, (cudaStream_t)stream>>>(adj, n_rows, n_cols, rowptr, col, val);
  RAG_LAUNCH_OK("dense_fill_kernel");
  return RAG_OK;
}
