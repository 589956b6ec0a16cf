// Sharded retrieval "finish" as ONE kernel over NVLink peer memory: the step after every
// rank's fused similarity + top-k over its key rows.  It replaces, per step, two NCCL all-gathers of the
// candidates, the merge kernel, two owner gathers, two more all-gathers of [R,Q,k,row] bytes and the select by
// owner with
//   phase 1  push this rank's [Q,k] (score, global index) candidates into slot `rank` of every peer's
//            candidate block (plain stores to peer-mapped addresses: NVLink 5 / NVSwitch),
//   flag     release-add on every peer's arrival counter, acquire-spin on ours,
//   phase 2  merge R*k -> k per query row (score desc, index asc -- identical on every rank),
//   phase 3  every merged winner that lives in THIS rank's rows is read once from the local value / label
//            shard and stored straight into the result block of every peer (bit-exact row copies),
//   flag     second counter: when it reaches the target the local result blocks are complete.
// NVLink traffic per rank: Q*k*12 B * R (candidates) + Q*k*row_bytes (each owned row to R peers), i.e. 1/R of what
// the all-gather formulation moves; no [R,Q,k,d] staging and no zero fill.
//
// Buffers live in a symmetric allocation (same layout on every rank; torch.distributed._symmetric_memory on the
// host side).  Candidate and result blocks are double buffered by step parity: a rank can be at most one step
// ahead of a peer, because finishing step s needs every peer's phase-3 signal of step s.  Counters are monotonic
// (target = step * CTAs per launch), so nothing is ever reset.  Every spin has a clock64 timeout that traps: a
// missing peer fails loudly instead of hanging the GPU.
#include <cfloat>
#include "common.cuh"

namespace rag {

constexpr int XC_CTAS = 128;
constexpr int XC_THREADS = 256;
constexpr unsigned long long XC_TIMEOUT_CYCLES = 20000000000ull;   // ~10 s

struct XchgLayout {
  size_t off_flags, off_cand_s, off_cand_i, off_out_a, off_out_b, total;
  size_t cand_s_par, cand_i_par, out_a_par, out_b_par;   // bytes per parity block
};

static XchgLayout xchg_layout(int64_t Qm, int km, int world, int64_t rba, int64_t rbb) {
  XchgLayout l{};
  size_t off = 0;
  l.off_flags = off; off += align_up((size_t)2 * world * 8, 256);
  l.cand_s_par = align_up((size_t)world * Qm * km * 4, 256);
  l.cand_i_par = align_up((size_t)world * Qm * km * 8, 256);
  l.out_a_par = align_up((size_t)Qm * km * rba, 256);
  l.out_b_par = align_up((size_t)Qm * km * rbb, 256);
  l.off_cand_s = off; off += 2 * l.cand_s_par;
  l.off_cand_i = off; off += 2 * l.cand_i_par;
  l.off_out_a = off; off += 2 * l.out_a_par;
  l.off_out_b = off; off += 2 * l.out_b_par;
  l.total = off;
  return l;
}

struct XchgArgs {
  const float* local_s; const int64_t* local_i;
  int64_t Q; int k; int world; int rank;
  unsigned char* const* peers;            // device array [world]
  size_t off_flags, off_cand_s, off_cand_i, off_out_a, off_out_b;   // parity already applied to the last four
  const unsigned char* table_a; int64_t rba;
  const unsigned char* table_b; int64_t rbb;
  int64_t lo, hi;
  unsigned long long target;
  float* out_s; int64_t* out_i;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("red.release.sys.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// all threads' earlier stores (to any peer) are ordered before the counter increments seen by the peers
__device__ __forceinline__ void xc_signal(const XchgArgs& a, int slot) {
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < a.world) {
    unsigned long long* f = reinterpret_cast<unsigned long long*>(a.peers[threadIdx.x] + a.off_flags);
    red_release_sys(f + (size_t)slot * a.world + a.rank, 1ull);
  }
}
__device__ __forceinline__ void xc_wait(const XchgArgs& a, int slot) {
  if ((int)threadIdx.x < a.world) {
    const unsigned long long* f =
        reinterpret_cast<const unsigned long long*>(a.peers[a.rank] + a.off_flags) + (size_t)slot * a.world + threadIdx.x;
    const unsigned long long t0 = clock64();
    while (ld_acquire_sys(f) < a.target) {
      if (clock64() - t0 > XC_TIMEOUT_CYCLES) __trap();
      __nanosleep(64);
    }
  }
  __syncthreads();
}

// copy one row (rb bytes) from src to the same offset of every peer's block; all 32 lanes call
__device__ __forceinline__ void xc_row_to_peers(const XchgArgs& a, const unsigned char* src, size_t dst_off, int64_t rb,
                                                bool vec16, int lane) {
  if (vec16) {
    for (int64_t o = (int64_t)lane * 16; o < rb; o += 32 * 16) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + o));
      for (int p = 0; p < a.world; ++p) *reinterpret_cast<uint4*>(a.peers[p] + dst_off + o) = v;
    }
  } else {
    for (int64_t o = (int64_t)lane * 4; o < rb; o += 32 * 4) {
      const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(src + o));
      for (int p = 0; p < a.world; ++p) *reinterpret_cast<uint32_t*>(a.peers[p] + dst_off + o) = v;
    }
  }
}

__global__ void __launch_bounds__(XC_THREADS) sharded_finish_kernel(const XchgArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = XC_THREADS / 32;
  int64_t* li = reinterpret_cast<int64_t*>(smem_raw) + (size_t)warp * a.k;
  float* lv = reinterpret_cast<float*>(reinterpret_cast<int64_t*>(smem_raw) + (size_t)wpb * a.k) + (size_t)warp * a.k;
  const int64_t QK = a.Q * a.k;

  // ---- phase 1: push my candidates into slot `rank` of every peer --------------------------------
  for (int64_t e = (int64_t)blockIdx.x * XC_THREADS + threadIdx.x; e < QK; e += (int64_t)gridDim.x * XC_THREADS) {
    const float s = __ldg(a.local_s + e);
    const int64_t j = __ldg(a.local_i + e);
    for (int p = 0; p < a.world; ++p) {
      reinterpret_cast<float*>(a.peers[p] + a.off_cand_s)[(size_t)a.rank * QK + e] = s;
      reinterpret_cast<int64_t*>(a.peers[p] + a.off_cand_i)[(size_t)a.rank * QK + e] = j;
    }
  }
  xc_signal(a, 0);
  xc_wait(a, 0);

  // ---- phase 2 + 3: merge, then ship the rows this rank owns ---------------------------------------
  const float* cs = reinterpret_cast<const float*>(a.peers[a.rank] + a.off_cand_s);
  const int64_t* ci = reinterpret_cast<const int64_t*>(a.peers[a.rank] + a.off_cand_i);
  const bool vec_a = a.table_a && (a.rba % 16 == 0) && ((reinterpret_cast<uintptr_t>(a.table_a) & 15u) == 0);
  const bool vec_b = a.table_b && (a.rbb % 16 == 0) && ((reinterpret_cast<uintptr_t>(a.table_b) & 15u) == 0);
  const int total = a.world * a.k;
  for (int64_t q = (int64_t)blockIdx.x * wpb + warp; q < a.Q; q += (int64_t)gridDim.x * wpb) {
    for (int p = lane; p < a.k; p += 32) { lv[p] = -FLT_MAX; li[p] = INT64_MAX; }
    __syncwarp();
    for (int c0 = 0; c0 < total; c0 += 32) {
      const int c = c0 + lane;
      float s = -FLT_MAX; int64_t j = -1;
      if (c < total) {
        const int r = c / a.k, p = c - r * a.k;
        const size_t o = (size_t)r * QK + (size_t)q * a.k + p;
        s = __ldcg(cs + o); j = __ldcg(ci + o);              // written by peers: read at L2, never from a stale L1 line
      }
      unsigned m = __ballot_sync(0xffffffffu, j >= 0 && ranks_before(s, j, lv[a.k - 1], li[a.k - 1]));
      while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const float sv = __shfl_sync(0xffffffffu, s, src);
        const int64_t jv = __shfl_sync(0xffffffffu, j, src);
        if (ranks_before(sv, jv, lv[a.k - 1], li[a.k - 1])) warp_sorted_insert<int64_t>(lv, li, a.k, sv, jv, lane);
      }
    }
    for (int p = lane; p < a.k; p += 32) {
      a.out_s[q * a.k + p] = lv[p];
      a.out_i[q * a.k + p] = (li[p] == INT64_MAX) ? (int64_t)-1 : li[p];
    }
    for (int p = 0; p < a.k; ++p) {
      const int64_t g = li[p];
      if (g >= a.lo && g < a.hi) {
        const size_t slot = (size_t)q * a.k + p;
        if (a.table_a) xc_row_to_peers(a, a.table_a + (g - a.lo) * a.rba, a.off_out_a + slot * a.rba, a.rba, vec_a, lane);
        if (a.table_b) xc_row_to_peers(a, a.table_b + (g - a.lo) * a.rbb, a.off_out_b + slot * a.rbb, a.rbb, vec_b, lane);
      }
    }
    __syncwarp();
  }
  xc_signal(a, 1);
  xc_wait(a, 1);
}

}  // namespace rag

extern "C" int rag_xchg_layout(int64_t Q_max, int32_t k_max, int32_t world, int64_t row_bytes_a, int64_t row_bytes_b,
                               size_t* offsets_out) {
  RAG_REQUIRE(Q_max >= 1 && k_max >= 1 && k_max <= RAG_MAX_K && world >= 1 && world <= 32 && row_bytes_a >= 0 &&
                  row_bytes_b >= 0 && offsets_out,
              RAG_EINVAL, "xchg_layout: Q_max=%lld k_max=%d world=%d", (long long)Q_max, k_max, world);
  const rag::XchgLayout l = rag::xchg_layout(Q_max, k_max, world, row_bytes_a, row_bytes_b);
  offsets_out[0] = l.total;
  offsets_out[1] = l.off_flags;
  offsets_out[2] = l.off_out_a; offsets_out[3] = l.out_a_par;
  offsets_out[4] = l.off_out_b; offsets_out[5] = l.out_b_par;
  offsets_out[6] = l.off_cand_s; offsets_out[7] = l.off_cand_i;
  return RAG_OK;
}

extern "C" int rag_sharded_finish(const float* local_scores, const int64_t* local_idx, int64_t Q, int32_t k,
                                  int32_t world, int32_t rank, void* const* peers_dev, int64_t Q_max, int32_t k_max,
                                  const void* table_a, int64_t row_bytes_a, const void* table_b, int64_t row_bytes_b,
                                  int64_t lo, int64_t hi, uint64_t step, float* out_scores, int64_t* out_idx,
                                  rag_stream_t stream) {
  using namespace rag;
  RAG_REQUIRE(world >= 1 && world <= 32 && rank >= 0 && rank < world, RAG_EINVAL, "sharded_finish: world=%d rank=%d", world, rank);
  RAG_REQUIRE(Q >= 0 && Q <= Q_max && k >= 1 && k <= k_max && k_max <= RAG_MAX_K, RAG_EINVAL,
              "sharded_finish: Q=%lld (max %lld) k=%d (max %d)", (long long)Q, (long long)Q_max, k, k_max);
  RAG_REQUIRE(step >= 1, RAG_EINVAL, "sharded_finish: step counts from 1 and must increase by 1 per call on every rank");
  RAG_REQUIRE(row_bytes_a % 4 == 0 && row_bytes_b % 4 == 0 && row_bytes_a >= 0 && row_bytes_b >= 0, RAG_EUNSUPPORTED,
              "sharded_finish: row sizes must be multiples of 4 bytes (%lld, %lld)", (long long)row_bytes_a, (long long)row_bytes_b);
  RAG_REQUIRE(peers_dev && local_scores && local_idx && out_scores && out_idx, RAG_EINVAL, "sharded_finish: null pointer");
  RAG_REQUIRE(lo <= hi, RAG_EINVAL, "sharded_finish: lo > hi");
  const XchgLayout l = xchg_layout(Q_max, k_max, world, row_bytes_a, row_bytes_b);
  const int par = (int)(step & 1ull);
  XchgArgs a{};
  a.local_s = local_scores; a.local_i = local_idx; a.Q = Q; a.k = k; a.world = world; a.rank = rank;
  a.peers = reinterpret_cast<unsigned char* const*>(peers_dev);
  a.off_flags = l.off_flags;
  a.off_cand_s = l.off_cand_s + par * l.cand_s_par; a.off_cand_i = l.off_cand_i + par * l.cand_i_par;
  a.off_out_a = l.off_out_a + par * l.out_a_par; a.off_out_b = l.off_out_b + par * l.out_b_par;
  a.table_a = static_cast<const unsigned char*>(table_a); a.rba = row_bytes_a;
  a.table_b = static_cast<const unsigned char*>(table_b); a.rbb = row_bytes_b;
  a.lo = lo; a.hi = hi;
  a.target = (unsigned long long)step * XC_CTAS;
  a.out_s = out_scores; a.out_i = out_idx;
  const size_t smem = (size_t)(XC_THREADS / 32) * k * 12;
  // The kernel's CTAs wait for one another (and for the peers) inside the launch: they must all be resident at once.
  // A cooperative launch makes the driver guarantee that -- or fail the launch -- instead of leaving it to "128 CTAs
  // surely fit 148 SMs", which a kernel running beside it on another stream could turn into a 10 s spin-and-trap.
  void* kargs[] = {const_cast<XchgArgs*>(&a)};
  cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(sharded_finish_kernel), dim3(XC_CTAS), dim3(XC_THREADS),
                                              kargs, smem, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchCooperativeKernel(sharded_finish_kernel)");
  RAG_LAUNCH_OK("sharded_finish_kernel");
  return RAG_OK;
}
