// K6-K9: CSR SpMM  Y = epilogue(A . X)  -- the message-passing propagation.
//
// Replaces, with one kernel:
//   * Propagation.aggregate_k_hop_features  (RAGraph_node/ragraph_utils/Propagation.py:15-25):
//     dense adj / rowsum, matmul, relu            -> RAG_EPI_ROWNORM | RAG_EPI_RELU
//   * GCN.forward  (RAGraph_node/layers/gcn.py:32-40): adj @ (XW) + b, PReLU
//                                                 -> RAG_EPI_BIAS | RAG_EPI_PRELU
//   * edge _agg  (RAGraph_edge/modules/RAGraph.py:232-240): gather * w -> scatter_add_
//                                                 -> no epilogue (CSR grouped by dst)
//   * the convex blend / layer sum of RAGraph.forward -> RAG_EPI_BLEND / RAG_EPI_ACCUM
//
// Layout and schedule (HBM-bound: algorithmic bytes = nnz*(8 + 4F) + n*(4F + 4/8)):
//   * one warp owns one output row; a lane-group of LANES lanes spans the F/4 float4 columns,
//     each lane keeping VPL float4 accumulators (F=256 -> 32 lanes x 2; F=64 -> two
//     16-lane groups working on alternate nonzeros);
//   * the (col, val) stream of the row is staged through a per-warp shared-memory ring with
//     cp.async (LDGSTS), one 32-entry tile ahead of the consumer, so the dependent
//     col -> X[col] chain never stalls on the index load;
//   * 4 nonzeros (x VPL float4) are in flight per lane before the FMAs retire;
//   * rows are handed out dynamically in blocks of ROWS_PER_GRAB to persistent CTAs
//     (grid = SMs x resident CTAs); rows longer than LONG_ROW nonzeros are processed by all
//     8 warps of the CTA and reduced through shared memory in warp order, so a 17k-degree
//     hub costs ~2k nonzeros of latency instead of 17k.  No atomics on Y: results are
//     deterministic.
#include "common.cuh"

namespace rag {

constexpr int SPMM_THREADS = 256;
constexpr int SPMM_WARPS = SPMM_THREADS / 32;
constexpr int ROWS_PER_GRAB = 64;   // rows per dynamic work item (8 per warp) on large graphs; see SpmmArgs::rows_per_grab
constexpr int LONG_ROW = 1024;      // rows with more nonzeros are split across the CTA
constexpr int STAGE = 32;           // (col,val) entries per cp.async tile

__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

struct SpmmArgs {
  const void* rowptr; int ptr_is_64;
  const int32_t* col; const float* val;
  int64_t n_rows; int64_t n_src;
  const float* X; int F;
  uint32_t epi; const float* bias; const float* alpha; const float* blend_in; float blend_w;
  const float* accum_in; float* Y;
  unsigned long long* work_counter;   // dynamic row-block scheduler
  // rows per work item: ROWS_PER_GRAB on large graphs; small graphs (the reference's real ones: a few thousand rows of ~5
  // non-zeros, where a row is three dependent memory round trips) get down to one row per warp, so that every SM has work
  // and no warp walks eight latency chains back to back (measured, Cora-shaped graph n = 2 708: 24 us -> see DESIGN 3.4)
  int rows_per_grab;
};

__device__ __forceinline__ int64_t load_ptr(const void* rowptr, int is64, int64_t i) {
  return is64 ? __ldg(reinterpret_cast<const int64_t*>(rowptr) + i)
              : (int64_t)__ldg(reinterpret_cast<const int32_t*>(rowptr) + i);
}

__device__ __forceinline__ float apply_epi(float y, float inv_deg_num, bool rownorm, uint32_t epi, float b,
                                           float alpha, float blend, float blend_w, float accum) {
  if (rownorm) y = y / inv_deg_num;                       // (sum a_ij x_j) / deg_i ; 0/0 -> NaN as torch
  if (epi & RAG_EPI_BIAS) y += b;
  if (epi & RAG_EPI_RELU) y = (y > 0.f || y != y) ? y : 0.f;    // torch.relu keeps NaN (fmaxf would turn a 0/0 row into 0)
  if (epi & RAG_EPI_PRELU) y = y >= 0.f ? y : alpha * y;
  if (epi & RAG_EPI_BLEND) y = y * (1.0f - blend_w) + blend * blend_w;
  if (epi & RAG_EPI_ACCUM) y += accum;
  return y;
}

// Accumulate nonzeros [beg, end) of one row into acc[] (this lane's float4 columns).
// ring: per-warp smem ring of 2 x STAGE (col,val) pairs.
template <int LANES, int VPL>
__device__ __forceinline__ void accumulate_range(const SpmmArgs& a, int64_t beg, int64_t end, int lane,
                                                 int32_t* ring_col, float* ring_val, float4 (&acc)[VPL],
                                                 float& valsum) {
  constexpr int GROUPS = 32 / LANES;            // nonzeros processed per warp step
  constexpr int F4 = LANES * VPL;
  const int sub = lane % LANES;
  const int grp = lane / LANES;
  const float4* X4 = reinterpret_cast<const float4*>(a.X);
  const bool has_val = a.val != nullptr;

  auto stage_tile = [&](int64_t base, int buf) {
    int64_t j = base + lane;
    if (j < end) {
      cp_async4(ring_col + buf * STAGE + lane, a.col + j);
      if (has_val) cp_async4(ring_val + buf * STAGE + lane, a.val + j);
    }
    cp_async_commit();
  };

  if (beg >= end) return;
  stage_tile(beg, 0);
  int buf = 0;
  for (int64_t base = beg; base < end; base += STAGE, buf ^= 1) {
    if (base + STAGE < end) { stage_tile(base + STAGE, buf ^ 1); cp_async_wait<1>(); }
    else cp_async_wait<0>();
    __syncwarp();
    const int cnt = (int)min((int64_t)STAGE, end - base);
    const int32_t* rc = ring_col + buf * STAGE;
    const float* rv = ring_val + buf * STAGE;
    // each group walks entries grp, grp+GROUPS, ...; 4 entries in flight per lane
    for (int e0 = 0; e0 < cnt; e0 += 4 * GROUPS) {
      float4 x[4][VPL];
      float w[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = e0 + u * GROUPS + grp;
        const bool ok = e < cnt;
        const int32_t c = ok ? rc[e] : 0;
        w[u] = ok ? (has_val ? rv[e] : 1.0f) : 0.f;
#pragma unroll
        for (int v = 0; v < VPL; ++v)
          x[u][v] = ok ? __ldg(X4 + (int64_t)c * F4 + sub + v * LANES) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        valsum += w[u];
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          acc[v].x = fmaf(w[u], x[u][v].x, acc[v].x); acc[v].y = fmaf(w[u], x[u][v].y, acc[v].y);
          acc[v].z = fmaf(w[u], x[u][v].z, acc[v].z); acc[v].w = fmaf(w[u], x[u][v].w, acc[v].w);
        }
      }
    }
    __syncwarp();   // ring slot `buf` may be overwritten by the stage issued next iteration
  }
}

// fold the GROUPS partial sums (lane-groups worked on alternate nonzeros) into group 0
template <int LANES, int VPL>
__device__ __forceinline__ void fold_groups(float4 (&acc)[VPL], float& valsum) {
#pragma unroll
  for (int o = 16; o >= LANES; o >>= 1) {
    valsum += __shfl_xor_sync(0xffffffffu, valsum, o);
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      acc[v].x += __shfl_xor_sync(0xffffffffu, acc[v].x, o);
      acc[v].y += __shfl_xor_sync(0xffffffffu, acc[v].y, o);
      acc[v].z += __shfl_xor_sync(0xffffffffu, acc[v].z, o);
      acc[v].w += __shfl_xor_sync(0xffffffffu, acc[v].w, o);
    }
  }
}

template <int LANES, int VPL>
__device__ __forceinline__ void store_row(const SpmmArgs& a, int64_t row, int lane, const float4 (&acc)[VPL],
                                          float valsum) {
  constexpr int F4 = LANES * VPL;
  if (lane >= LANES) return;
  const bool rownorm = (a.epi & RAG_EPI_ROWNORM) != 0;
  const float alpha = (a.epi & RAG_EPI_PRELU) ? __ldg(a.alpha) : 0.f;
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int c4 = lane + v * LANES;
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f), bl = b, ac = b;
    if (a.epi & RAG_EPI_BIAS) b = __ldg(reinterpret_cast<const float4*>(a.bias) + c4);
    if (a.epi & RAG_EPI_BLEND) bl = __ldg(reinterpret_cast<const float4*>(a.blend_in) + row * F4 + c4);
    if (a.epi & RAG_EPI_ACCUM) ac = __ldg(reinterpret_cast<const float4*>(a.accum_in) + row * F4 + c4);
    float4 y;
    y.x = apply_epi(acc[v].x, valsum, rownorm, a.epi, b.x, alpha, bl.x, a.blend_w, ac.x);
    y.y = apply_epi(acc[v].y, valsum, rownorm, a.epi, b.y, alpha, bl.y, a.blend_w, ac.y);
    y.z = apply_epi(acc[v].z, valsum, rownorm, a.epi, b.z, alpha, bl.z, a.blend_w, ac.z);
    y.w = apply_epi(acc[v].w, valsum, rownorm, a.epi, b.w, alpha, bl.w, a.blend_w, ac.w);
    reinterpret_cast<float4*>(a.Y)[row * F4 + c4] = y;
  }
}

// MINB (resident CTAs per SM the register budget is cut for): F = 256 (VPL = 2) needs 79 registers; squeezed into the 64 of
// four CTAs per SM it spilled 24-32 bytes and ran 6 % slower than three spill-free ones (measured at cfg5: 10.58 vs 9.94 ms).
template <int LANES, int VPL, int MINB = (VPL < 2 ? 4 : (VPL == 2 ? 3 : 2))>
__global__ void __launch_bounds__(SPMM_THREADS, MINB)
csr_spmm_kernel(const SpmmArgs a) {
  constexpr int F4 = LANES * VPL;
  __shared__ int32_t s_col[SPMM_WARPS][2 * STAGE];
  __shared__ float s_val[SPMM_WARPS][2 * STAGE];
  __shared__ float4 s_part[SPMM_WARPS][F4];        // long-row partial sums (F=256: 8 KB)
  __shared__ float s_vsum[SPMM_WARPS];
  __shared__ long long s_block;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_blocks = (a.n_rows + a.rows_per_grab - 1) / a.rows_per_grab;

  for (;;) {
    __syncthreads();                               // s_block / s_part reuse
    if (threadIdx.x == 0) s_block = (long long)atomicAdd(a.work_counter, 1ull);
    __syncthreads();
    const int64_t blk = s_block;
    if (blk >= n_blocks) break;
    const int64_t r0 = blk * a.rows_per_grab;
    const int64_t r1 = min(r0 + (int64_t)a.rows_per_grab, a.n_rows);

    // pass 1: short rows, one warp each (rows interleaved across the 8 warps)
    bool any_long = false;
    for (int64_t row = r0 + warp; row < r1; row += SPMM_WARPS) {
      const int64_t beg = load_ptr(a.rowptr, a.ptr_is_64, row), end = load_ptr(a.rowptr, a.ptr_is_64, row + 1);
      if (end - beg > LONG_ROW) { any_long = true; continue; }
      float4 acc[VPL];
#pragma unroll
      for (int v = 0; v < VPL; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
      float valsum = 0.f;
      accumulate_range<LANES, VPL>(a, beg, end, lane, s_col[warp], s_val[warp], acc, valsum);
      fold_groups<LANES, VPL>(acc, valsum);
      store_row<LANES, VPL>(a, row, lane, acc, valsum);
    }
    // pass 2: long rows, the whole CTA per row
    if (__syncthreads_or(any_long)) {
      for (int64_t row = r0; row < r1; ++row) {
        const int64_t beg = load_ptr(a.rowptr, a.ptr_is_64, row), end = load_ptr(a.rowptr, a.ptr_is_64, row + 1);
        if (end - beg <= LONG_ROW) continue;
        // contiguous slices, multiples of STAGE so the cp.async tiles stay aligned to the slice
        int64_t per = ((end - beg + SPMM_WARPS - 1) / SPMM_WARPS + STAGE - 1) / STAGE * STAGE;
        const int64_t b = min(beg + warp * per, end), e = min(b + per, end);
        float4 acc[VPL];
#pragma unroll
        for (int v = 0; v < VPL; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        float valsum = 0.f;
        accumulate_range<LANES, VPL>(a, b, e, lane, s_col[warp], s_val[warp], acc, valsum);
        fold_groups<LANES, VPL>(acc, valsum);
        if (lane < LANES) {
#pragma unroll
          for (int v = 0; v < VPL; ++v) s_part[warp][lane + v * LANES] = acc[v];
        }
        if (lane == 0) s_vsum[warp] = valsum;
        __syncthreads();
        if (warp == 0) {
          float4 tot[VPL];
          float vs = 0.f;
#pragma unroll
          for (int v = 0; v < VPL; ++v) tot[v] = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int w = 0; w < SPMM_WARPS; ++w) {      // fixed order -> deterministic
            vs += s_vsum[w];
            if (lane < LANES) {
#pragma unroll
              for (int v = 0; v < VPL; ++v) {
                float4 p = s_part[w][lane + v * LANES];
                tot[v].x += p.x; tot[v].y += p.y; tot[v].z += p.z; tot[v].w += p.w;
              }
            }
          }
          store_row<LANES, VPL>(a, row, lane, tot, vs);
        }
        __syncthreads();
      }
    }
  }
}

// Any F (not a multiple of 4, or unaligned): one warp per row, columns in tiles of 32.
__global__ void __launch_bounds__(256) csr_spmm_generic_kernel(const SpmmArgs a) {
  const int lane = threadIdx.x & 31;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const bool rownorm = (a.epi & RAG_EPI_ROWNORM) != 0;
  const float alpha = (a.epi & RAG_EPI_PRELU) ? __ldg(a.alpha) : 0.f;
  for (int64_t row = warp; row < a.n_rows; row += nwarps) {
    const int64_t beg = load_ptr(a.rowptr, a.ptr_is_64, row), end = load_ptr(a.rowptr, a.ptr_is_64, row + 1);
    for (int c0 = 0; c0 < a.F; c0 += 32) {
      const int c = c0 + lane;
      float acc = 0.f, valsum = 0.f;
      for (int64_t j = beg; j < end; ++j) {
        const float w = a.val ? __ldg(a.val + j) : 1.0f;
        valsum += w;
        if (c < a.F) acc = fmaf(w, __ldg(a.X + (int64_t)__ldg(a.col + j) * a.F + c), acc);
      }
      if (c < a.F) {
        const int64_t o = row * a.F + c;
        a.Y[o] = apply_epi(acc, valsum, rownorm, a.epi, (a.epi & RAG_EPI_BIAS) ? __ldg(a.bias + c) : 0.f, alpha,
                           (a.epi & RAG_EPI_BLEND) ? __ldg(a.blend_in + o) : 0.f, a.blend_w,
                           (a.epi & RAG_EPI_ACCUM) ? __ldg(a.accum_in + o) : 0.f);
      }
    }
  }
}

// scheduler counters: a small ring so back-to-back launches on one stream never share one
__device__ unsigned long long g_spmm_counters[64];
static unsigned g_spmm_next = 0;

template <int LANES, int VPL, int MINB = (VPL < 2 ? 4 : (VPL == 2 ? 3 : 2))>
static int launch_spmm(SpmmArgs a, cudaStream_t s) {
  unsigned long long* base = nullptr;
  cudaError_t e = cudaGetSymbolAddress(reinterpret_cast<void**>(&base), g_spmm_counters);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetSymbolAddress(g_spmm_counters)");
  a.work_counter = base + (g_spmm_next++ & 63u);
  e = cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned long long), s);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(work_counter)");
  const int resident = MINB;
  int64_t grid = (int64_t)sm_count() * resident;
  a.rows_per_grab = ROWS_PER_GRAB;
  while (a.rows_per_grab > SPMM_WARPS && (a.n_rows + a.rows_per_grab - 1) / a.rows_per_grab < 2 * grid) a.rows_per_grab /= 2;
  int64_t n_blocks = (a.n_rows + a.rows_per_grab - 1) / a.rows_per_grab;
  if (grid > n_blocks) grid = n_blocks;
  csr_spmm_kernel<LANES, VPL, MINB><<<(unsigned)grid, SPMM_THREADS, 0, s>>>(a);
  RAG_LAUNCH_OK("csr_spmm_kernel");
  return RAG_OK;
}

}  // namespace rag

extern "C" int rag_csr_spmm_f32(const void* rowptr, int32_t ptr_is_64, const int32_t* col, const float* val,
                                int64_t n_rows, int64_t n_src, int64_t nnz, const float* X, int32_t F,
                                uint32_t epilogue, const float* bias, const float* alpha, const float* blend_in,
                                float blend_w, const float* accum_in, float* Y, rag_stream_t stream) {
  using namespace rag;
  RAG_REQUIRE(n_rows >= 0 && n_src >= 0 && nnz >= 0 && F >= 1, RAG_EINVAL,
              "csr_spmm: n_rows=%lld n_src=%lld nnz=%lld F=%d", (long long)n_rows, (long long)n_src,
              (long long)nnz, F);
  RAG_REQUIRE(n_src <= 0x7fffffffLL, RAG_EUNSUPPORTED, "csr_spmm: n_src=%lld exceeds int32 column indices",
              (long long)n_src);
  if (n_rows == 0) return RAG_OK;
  RAG_REQUIRE(rowptr && Y && (nnz == 0 || (col && X)), RAG_EINVAL, "csr_spmm: null pointer");
  RAG_REQUIRE(X != Y, RAG_EINVAL, "csr_spmm: X and Y must not alias");
  RAG_REQUIRE(!(epilogue & RAG_EPI_BIAS) || bias, RAG_EINVAL, "csr_spmm: RAG_EPI_BIAS without bias");
  RAG_REQUIRE(!(epilogue & RAG_EPI_PRELU) || alpha, RAG_EINVAL, "csr_spmm: RAG_EPI_PRELU without alpha");
  RAG_REQUIRE(!(epilogue & RAG_EPI_BLEND) || blend_in, RAG_EINVAL, "csr_spmm: RAG_EPI_BLEND without blend_in");
  RAG_REQUIRE(!(epilogue & RAG_EPI_ACCUM) || accum_in, RAG_EINVAL, "csr_spmm: RAG_EPI_ACCUM without accum_in");
  SpmmArgs a{rowptr, ptr_is_64, col, val, n_rows, n_src, X, F, epilogue, bias, alpha, blend_in, blend_w,
             accum_in, Y, nullptr, ROWS_PER_GRAB};
  cudaStream_t s = (cudaStream_t)stream;
  const bool vec = (F % 4 == 0) && aligned16(X) && aligned16(Y) && (!bias || aligned16(bias)) &&
                   (!blend_in || aligned16(blend_in)) && (!accum_in || aligned16(accum_in));
  if (vec) {
    switch (F) {
      case 512: return launch_spmm<32, 4>(a, s);
      case 256: return launch_spmm<32, 2>(a, s);
      case 128: return launch_spmm<32, 1>(a, s);
      case 64: return launch_spmm<16, 1>(a, s);
      case 32: return launch_spmm<8, 1>(a, s);
      case 16: return launch_spmm<4, 1>(a, s);
      default: break;
    }
  }
  int64_t blocks = (n_rows + 7) / 8;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  csr_spmm_generic_kernel<<<(unsigned)blocks, 256, 0, s>>>(a);
  RAG_LAUNCH_OK("csr_spmm_generic_kernel");
  return RAG_OK;
}
