// K1+K2+K3 on the 5th-gen tensor cores: bf16 similarity filter (tcgen05.mma, accumulators in TMEM,
// operands staged by TMA) with a fused per-row top-k' selection, followed by an exact fp32 re-score of
// the surviving candidates with a correctness certificate (RAG_SIM_BF16_REFINE), or by a plain merge of
// the approximate scores (RAG_SIM_BF16).
//
// Replaces SimilarityFunctions.calculate_cosine_similarity + torch.topk for large libraries
// (RAGraph_node/ragraph_utils/SimilarityFunctions.py:6-16, ToyGraphBase.py:53,67;
//  RAGraph_edge/modules/RAGraph.py:298-311).
//
// Filter kernel (one CTA per (query tile, key split); 10 warps, 1 CTA/SM):
//   * a CTA owns 256 query rows (two 128-row blocks) whose normalised bf16 image stays in shared memory
//     for the whole kernel (TMA, SWIZZLE_128B, K-major);
//   * warp 8 (one elected lane) streams the split's keys, 128 keys x 64 K (16 KB) per pipeline stage;
//   * warp 9 (one elected lane) issues tcgen05.mma.cta_group::1.kind::f16 M=128 N=128 K=16 into one of two
//     TMEM accumulator buffers (2 row blocks x 128 columns each -> all 512 columns), tcgen05.commit frees
//     smem stages and publishes finished accumulators through mbarriers;
//   * warps 0-7 drain the other buffer: tcgen05.ld 32x32b.x32 gives each thread 32 scores of ITS query
//     row; a 3-input-max tree compares them with the row's running k'-th best held in a register, and only
//     on a hit does the thread insert into its row's sorted list in shared memory.  Scores never reach HBM.
//   CTAs that share a key split run side by side (blockIdx = split * q_tiles + q_tile), so a key tile is
//   fetched from HBM once and served to the other query tiles from L2.
//
// Exactness (mode 3): for unit vectors |s_bf16 - s| <= EPS (2 roundings of 2^-9 each, Cauchy-Schwarz).
// Every key NOT in a split's list has s_bf16 <= t_split (the list's last score), hence s <= t_split + EPS.
// The refine kernel re-scores all S*k' candidates in fp32, keeps the best k (score desc, index asc) and
// certifies the row iff its k-th exact score > max_split t_split + EPS.  Uncertified rows (ties or dense
// clusters at the boundary) are recomputed by the fp32 kernel, so results always equal RAG_SIM_FP32.
#include <cuda.h>
#include <algorithm>
#include <cfloat>
#include <cstdlib>
#include <cstring>
#include <cuda_bf16.h>
#include "common.cuh"

namespace rag {

// topk_f32.cu: fp32 path restricted to a device-side list of rows
int topk_f32_run_rows(const float* q, int64_t Q, const float* keys, const float* key_inv_norm, const float* q_inv_norm,
                      int64_t N, int d, int k, int64_t idx_offset, const int32_t* row_map, const int32_t* n_rows_dev,
                      float* out_scores, int64_t* out_idx, void* ws, size_t ws_bytes, cudaStream_t s,
                      const int64_t* mask_rowptr = nullptr, const int64_t* mask_col = nullptr);
size_t topk_f32_rows_workspace(int64_t Q, int64_t N, int d, int k);

constexpr int TC_ROWS = 256;          // query rows per CTA (2 row blocks of 128)
constexpr int TC_BN = 128;            // keys per tile
constexpr int TC_BOX_BYTES = 128 * 128;   // one TMA box: 128 rows x 64 bf16
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_THREADS = (TC_EPI_WARPS + 2) * 32;     // SS kernel: 8 epilogue warps + TMA + MMA
constexpr int TS_THREADS = (TC_EPI_WARPS + 3) * 32;     // TS kernel: 8 epilogue warps + TMA + 2 MMA issuers
constexpr int TC_PQ = 4;               // per-row pending-candidate queue depth (drained after the TMEM hand-back)
// Score error bound of the 16-bit filter, per query row: eps_row = err_q[row] + err_k + TC_SLACK, where err_q / err_k are
// the MEASURED rounding-error norms of the operand images (rows_to_16_kernel; Cauchy-Schwarz) or, without them, the
// element-wise worst case u per operand (u = 2^-8 bf16, 2^-11 fp16).  TC_SLACK covers what the norms do not: the fp32
// accumulation inside the tensor core (<= d products of magnitude <= 1, truncating adds: < 2^-16 for d <= 256), the
// (1 + 2^-7) factor of ||kh||, and the few-ulp difference between the refine kernel's fp32 score and the ideal one.
constexpr float TC_SLACK = 3.0517578125e-05f;           // 2^-15
constexpr float TC_U_BF16 = 0.00390625f;                // 2^-8
constexpr float TC_U_F16 = 0.00048828125f;              // 2^-11
constexpr int64_t TC_PASS2_MIN_KEYS = 32768;             // smaller libraries skip the second tensor-core pass
constexpr int TC_SPILL_MAX = 1024;                      // pass-2 candidates per row (second tensor-core pass)
constexpr unsigned long long TC_TIMEOUT_CYCLES = 20000000000ull;   // ~10 s: trap instead of hanging the GPU

// profiling trace: built only with -DRAG_TC_TRACE_BUILD=1 (python -m ragraph_b200.build --trace); the production
// kernel carries no trace registers.  CTA 0 stamps clock64 at pipeline events of a 512-tile window (RAG_TC_TRACE=1)
#ifndef RAG_TC_TRACE_BUILD
#define RAG_TC_TRACE_BUILD 0
#endif
#define TC_TRACE_ON(a) (RAG_TC_TRACE_BUILD && (a).trace)
__device__ unsigned long long g_tc_trace[4 * 512];
__device__ unsigned int g_tc_trace2[2 * 512];
__device__ unsigned int g_tc_trace3[16 * 256];  // MMA-thread micro-trace (TS kernel): 16 clock stamps per tile, first 256 tiles of the window   // per tile: drains by warp 0, cycles from tmem_full to tmem_empty arrive

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t addr, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.b32 %0, 1, 0, P;\n\t}"
               : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  return done != 0;
}
__device__ __forceinline__ bool mbar_test(uint32_t addr, uint32_t parity) {     // non-blocking probe
  uint32_t done;
  asm volatile("{\n\t.reg .pred P;\n\tmbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.b32 %0, 1, 0, P;\n\t}"
               : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  if (mbar_try(addr, parity)) return;                    // common case: no clock read on the critical path
  const unsigned long long t0 = clock64();
  while (!mbar_try(addr, parity)) {
    if (clock64() - t0 > TC_TIMEOUT_CYCLES) __trap();     // a broken pipeline must fail loudly, not hang
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accum) : "memory");
}
// kind::tf32: fp32 words in shared memory, the tensor core reads the upper 19 bits; K = 8 per instruction (32 bytes, the
// same K-step in bytes as kind::f16, so descriptors, swizzle and stage geometry are shared with the bf16 path)
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                 "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                 "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr));
}
// wait for the outstanding tcgen05.ld and tie the destination registers to the wait, so that no consumer of v[]
// can be scheduled above it when another chunk's load is issued in between (software pipelining)
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&v)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
               :: "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// wait for the outstanding tcgen05.ld and tie the destination registers to the wait, so that no consumer of
// v[] can be scheduled above it when another chunk's load is issued in between (software pipelining)
__device__ __forceinline__ void tmem_ld_wait_regs(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                 "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                 "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
               :: "memory");
}
// no instruction: makes the compiler treat v[] as redefined here, so consumers stay below the preceding volatile wait
__device__ __forceinline__ void tie_regs(uint32_t (&v)[32]) {
  asm volatile(""
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                 "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                 "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
               :: "memory");
}
// v[j] for a per-lane dynamic j: 31 selects (registers cannot be indexed dynamically)
__device__ __forceinline__ float select32(const uint32_t (&v)[32], int j) {
  uint32_t a[16], b[8], c[4], d[2];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = (j & 1) ? v[2 * i + 1] : v[2 * i];
#pragma unroll
  for (int i = 0; i < 8; ++i) b[i] = (j & 2) ? a[2 * i + 1] : a[2 * i];
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i] = (j & 4) ? b[2 * i + 1] : b[2 * i];
#pragma unroll
  for (int i = 0; i < 2; ++i) d[i] = (j & 8) ? c[2 * i + 1] : c[2 * i];
  return __uint_as_float((j & 16) ? d[1] : d[0]);
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (8-row x 128-byte atoms, SBO = 1024 B, LBO unused,
// descriptor version 1, layout type 2); `addr` is the 1024-aligned box base (+32 B per K=16 step).
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr) {
  const uint32_t lo = (addr >> 4) & 0x3FFFu;
  const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  return ((uint64_t)hi << 32) | lo;
}
// instruction descriptor: D=f32, A=B=bf16 (format code 1; fp16 = 0), both K-major, N=128, M=128
constexpr uint32_t TC_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((TC_BN >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t TC_IDESC_F16 = (1u << 4) | (0u << 7) | (0u << 10) | ((TC_BN >> 3) << 17) | ((128u >> 4) << 24);
// same shape with A = B = tf32 (format code 2 in bits [7,10) and [10,13))
constexpr uint32_t TC_IDESC_TF32 = (1u << 4) | (2u << 7) | (2u << 10) | ((TC_BN >> 3) << 17) | ((128u >> 4) << 24);

// Candidate list of one (query row, key split): KP unsorted (score, index) slots in shared memory, entry p of
// row r at ls[p*256 + r] (conflict free across a warp), plus meta[r] = slots in use | (position of the
// current minimum << 8).  A new candidate (s > threshold) fills a free slot or overwrites the minimum, then
// the minimum is found again by one pass over the KP scores.  Returns the new threshold: the list minimum once
// all KP slots are in use, -inf before.
// NOTE (measured): this kernel leaves ~1 KB of L1 (226 KB of the SM's 228 KB are shared memory), so ANY local
// memory traffic -- register spills, ABI saves around a non-inlined call -- costs an L2 round trip (~2000 cycles
// per slow-path visit in the first versions).  Everything below is written as small rolled loops over shared
// memory so the kernel has a zero-byte stack frame.
template <int KP>
__device__ __forceinline__ float tc_list_push(float* ls, int32_t* li, int32_t* meta, float s, int32_t idx) {
  const int m = *meta;
  int cnt = m & 0xff;
  const int pos = (cnt < KP) ? cnt : (m >> 8);
  ls[pos * TC_ROWS] = s;
  li[pos * TC_ROWS] = idx;
  if (cnt < KP) {
    ++cnt;
    if (cnt < KP) { *meta = cnt; return -INFINITY; }
  }
  // new minimum: KP independent loads issued back to back, then a depth-log2(KP) tournament -- dependent
  // load->compare->branch chains cost ~50 cycles per element on a lone warp (measured), this costs ~150 in all
  float v[KP];
  int ps[KP];
#pragma unroll
  for (int p = 0; p < KP; ++p) { v[p] = ls[p * TC_ROWS]; ps[p] = p; }
#pragma unroll
  for (int w = KP / 2; w >= 1; w >>= 1) {
#pragma unroll
    for (int p = 0; p < w; ++p) {
      const bool hi = v[p + w] < v[p];
      v[p] = hi ? v[p + w] : v[p];
      ps[p] = hi ? ps[p + w] : ps[p];
    }
  }
  *meta = KP | (ps[0] << 8);
  return v[0];
}

struct TcArgs {
  int64_t Q; int64_t N;            // rows in use
  int n_qtiles; int n_splits; int tiles_per_split; int n_tiles;
  int kp;                          // list length per (row, split)
  float* part_s; int32_t* part_i;  // [n_splits][Q][kp]
  // threshold pre-pass (TS kernel): premax = 1 runs only the first pre_tiles tiles of every split and records, per query
  // row, the maximum score of pre_groups consecutive tile groups into gmax[(split * pre_groups + g) * Q + row]; the main
  // pass then starts every row at thr0[row] (a proven lower bound of its final kp-th best score) instead of -inf
  int premax; int pre_tiles; int pre_groups;
  float* gmax; const float* thr0;
  float* overflow;                 // TS kernel: [CTAs][8 warps][32 lanes][128] scores parked when a warp's hit queue overflows
  uint32_t idesc;                  // tcgen05 instruction descriptor (operand format bf16 / fp16 / tf32 is a run-time field)
  // ---- second pass (TS kernel, `collect`): the rows refine could not certify, as a device-side list ----
  // row_map[i] = query row of compact row i, *n_rows_dev = how many; the CTA derives (query tile, key split) geometry from
  // that count on the device (no host sync), CTAs beyond it exit at once.  thr0 is then indexed by COMPACT row and taken
  // as is (thr_exact): a fixed threshold = exact k-th score - error bound.  Every score above it is a candidate: a full
  // append buffer is spilled to spill_{s,i}[compact row][spill_cap] (slot from an atomic per-row counter) instead of
  // being compacted against a top-k' list, so the pass returns ALL keys that can still belong to the exact top k.
  const int32_t* row_map; const int32_t* n_rows_dev;
  int collect; int thr_exact;
  float* spill_s; int32_t* spill_i; int32_t* spill_cnt; int spill_cap;
  // ---- cross-split threshold sharing (TS kernel, main pass) ----
  // The key splits of a query row only need what can reach the row's GLOBAL top k, but a CTA sees one split.  Every
  // candidate a CTA accepts is mirrored into pool[(row * n_splits + split) * pool_slots + slot] (slot = its position in the
  // append buffer: one key per slot, distinct keys in distinct slots at all times); the CTAs past the workers
  // (blockIdx >= n_qtiles * n_splits: the SMs the (query tile, split) grid leaves idle) sweep the pool, take the k-th
  // largest published score of a row -- a lower bound of the row's final k-th best 16-bit score, whatever subset of the
  // candidates is visible -- and publish gthr[row] = that - g_margin(row); the workers fold gthr into their thresholds
  // every few tiles.  A word of 0 means "nothing yet" (the prologue clears pool and gthr).  part_thr[split][row] = the
  // threshold a CTA ended with (refine: every key the split did not list scores <= it).
  uint32_t* pool; uint32_t* gthr; float* part_thr; int32_t* done;
  int pool_slots; int n_mergers; int g_k; int g_dbg;   // g_dbg (experiments): 1 = sweeping CTAs leave at once, 2 = workers ignore gthr, 3 = workers do not publish
  const float* g_qerr; const float* g_kerr; float g_eps_fixed; int g_exact;
  uint32_t* hit_count;             // nullable diagnostic: (lane, chunk) hits queued by all epilogue warps of the launch
  int trace;                       // RAG_TC_DEBUG=3 or RAG_TC_TRACE=1: CTA 0 stamps clock64 at pipeline events (g_tc_trace)
  int trace_t0;                    // RAG_TC_TRACE_T0: first traced tile (window of 512)
  int no_tma;                      // RAG_TC_NOTMA=1 (experiment): the producer arrives without loading keys (garbage operands)
  int swap_halves;                 // RAG_TS_SWAP=1 (diagnostic): swap the bf16 halves of every A word stored to TMEM
  int debug;                       // RAG_TC_DEBUG (profiling experiments only): 1 = epilogue skips TMEM reads, 2 = reads but skips the filter
};

struct __align__(8) TcBarriers {
  uint64_t a_full;
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t tmem_full[2][2];    // [buffer][row block]: the two 128-row accumulators of a tile are handed over separately
  uint64_t tmem_empty[2][2];
  uint32_t tmem_base;
};

// KH = K chunks of 128 bytes (bf16: d_pad = 64*KH; tf32: d_pad = 32*KH); NSTAGE = pipeline depth in 16 KB boxes;
// KP = list length; TF32 = operands are fp32 words consumed as tf32 (RAG_SIM_TF32), else bf16
template <int KH, int NSTAGE, int KP, bool TF32 = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
cosine_topk_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                      const TcArgs a) {
  extern __shared__ unsigned char smem_dyn[];
  // carve: [A: 2*KH boxes][B: NSTAGE boxes][lists: KP x 256 x (f32 + i32)][barriers]
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* sA = base;
  unsigned char* sB = sA + 2 * KH * TC_BOX_BYTES;
  float* list_s = reinterpret_cast<float*>(sB + NSTAGE * TC_BOX_BYTES);    // [KP][256]
  int32_t* list_i = reinterpret_cast<int32_t*>(list_s + KP * TC_ROWS);     // [KP][256]
  int32_t* list_meta = list_i + KP * TC_ROWS;                              // [256]
  float* pq_s = reinterpret_cast<float*>(list_meta + TC_ROWS);             // [TC_PQ][256] pending scores
  int32_t* pq_i = reinterpret_cast<int32_t*>(pq_s + TC_PQ * TC_ROWS);      // [TC_PQ][256] pending indices
  TcBarriers* bars = reinterpret_cast<TcBarriers*>(pq_i + TC_PQ * TC_ROWS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qtile = blockIdx.x % a.n_qtiles;
  const int split = blockIdx.x / a.n_qtiles;
  const int tile0 = split * a.tiles_per_split;
  const int tile1 = min(tile0 + a.tiles_per_split, a.n_tiles);
  const int n_my_tiles = max(tile1 - tile0, 0);
  constexpr int BOX_ELEMS = TF32 ? 32 : 64;                 // elements per 128-byte box row (TMA K coordinate step)
  const uint32_t idesc = a.idesc;
  auto mma = [idesc](uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accum) {
    if constexpr (TF32) tc_mma_tf32(tmem_d, da, db, idesc, accum);
    else tc_mma_bf16(tmem_d, da, db, idesc, accum);
  };

  // ---- one-time setup ---------------------------------------------------------------------------
  if (warp == TC_EPI_WARPS && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_q)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_k)) : "memory");
    mbar_init(&bars->a_full, 1);
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
    for (int b = 0; b < 2; ++b)
      for (int rb = 0; rb < 2; ++rb) { mbar_init(&bars->tmem_full[b][rb], 1); mbar_init(&bars->tmem_empty[b][rb], TC_EPI_WARPS / 2); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == TC_EPI_WARPS + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < KP * TC_ROWS; i += TC_THREADS) { list_s[i] = -INFINITY; list_i[i] = -1; }
  for (int i = threadIdx.x; i < TC_ROWS; i += TC_THREADS) list_meta[i] = 0;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == TC_EPI_WARPS) {
    // =============================== TMA producer ===============================================
    if (elect_one()) {
      mbar_expect_tx(&bars->a_full, 2 * KH * TC_BOX_BYTES);
      for (int rb = 0; rb < 2; ++rb)
        for (int kh = 0; kh < KH; ++kh)
          tma_load_2d(sA + (rb * KH + kh) * TC_BOX_BYTES, &map_q, kh * BOX_ELEMS, qtile * TC_ROWS + rb * 128, &bars->a_full);
      int s = 0; uint32_t ph = 0;
      for (int t = 0; t < n_my_tiles; ++t) {
        for (int kh = 0; kh < KH; ++kh) {
          mbar_wait(&bars->empty[s], ph ^ 1);
          mbar_expect_tx(&bars->full[s], TC_BOX_BYTES);
          tma_load_2d(sB + s * TC_BOX_BYTES, &map_k, kh * BOX_ELEMS, (tile0 + t) * TC_BN, &bars->full[s]);
          if (++s == NSTAGE) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == TC_EPI_WARPS + 1) {
    // =============================== MMA issuer =================================================
    if (elect_one()) {
      mbar_wait(&bars->a_full, 0);
      tc_fence_after();
      const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);
      int s = 0; uint32_t ph = 0;
      for (int t = 0; t < n_my_tiles; ++t) {
        const int b = t & 1;
        const uint32_t use_parity = ((uint32_t)t >> 1) & 1u;
        if constexpr (KH <= 2) {
          // Row block 0 first, then row block 1, each published on its own barrier: the epilogue of block 0 starts
          // half a tile earlier and each accumulator has 1.5 tile times (not 1) to be drained before its reuse.
          // (A tile's KH stages stay occupied until both row blocks have consumed them: needs NSTAGE >= 2*KH.)
          static_assert(NSTAGE >= 2 * KH, "row-block-major MMA order holds a tile's stages for the whole tile");
#pragma unroll
          for (int rb = 0; rb < 2; ++rb) {
            mbar_wait(&bars->tmem_empty[b][rb], use_parity ^ 1u);
            tc_fence_after();
            if (TC_TRACE_ON(a) && blockIdx.x == 0 && (unsigned)(t - a.trace_t0) < 512u && rb == 0) g_tc_trace[4 * (t - a.trace_t0) + 0] = clock64();
            int sk = s; uint32_t phk = ph;
            for (int kh = 0; kh < KH; ++kh) {
              if (rb == 0) { mbar_wait(&bars->full[sk], phk); tc_fence_after(); }
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4) {
                const uint64_t da = umma_desc(a_addr + (rb * KH + kh) * TC_BOX_BYTES + k4 * 32);
                const uint64_t db = umma_desc(b_addr + sk * TC_BOX_BYTES + k4 * 32);
                mma(tmem_base + (uint32_t)((b * 2 + rb) * TC_BN), da, db, (kh | k4) != 0 ? 1u : 0u);
              }
              if (++sk == NSTAGE) { sk = 0; phk ^= 1; }
            }
            tc_commit(&bars->tmem_full[b][rb]);             // this row block's accumulator is complete
          }
          for (int kh = 0; kh < KH; ++kh) {                 // smem stages reusable once all 2*KH*4 MMAs retire
            tc_commit(&bars->empty[s]);
            if (++s == NSTAGE) { s = 0; ph ^= 1; }
          }
        } else {
          // d = 256: the resident query tile leaves room for 3 stages only, so stages are released K-half by K-half
          // (both row blocks consume a stage back to back); the MMA time per tile doubles, the epilogue has slack.
          mbar_wait(&bars->tmem_empty[b][0], use_parity ^ 1u);
          mbar_wait(&bars->tmem_empty[b][1], use_parity ^ 1u);
          tc_fence_after();
          for (int kh = 0; kh < KH; ++kh) {
            mbar_wait(&bars->full[s], ph);
            tc_fence_after();
#pragma unroll
            for (int rb = 0; rb < 2; ++rb) {
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4) {
                const uint64_t da = umma_desc(a_addr + (rb * KH + kh) * TC_BOX_BYTES + k4 * 32);
                const uint64_t db = umma_desc(b_addr + s * TC_BOX_BYTES + k4 * 32);
                mma(tmem_base + (uint32_t)((b * 2 + rb) * TC_BN), da, db, (kh | k4) != 0 ? 1u : 0u);
              }
            }
            tc_commit(&bars->empty[s]);
            if (++s == NSTAGE) { s = 0; ph ^= 1; }
          }
          tc_commit(&bars->tmem_full[b][0]);
          tc_commit(&bars->tmem_full[b][1]);
        }
        if (TC_TRACE_ON(a) && blockIdx.x == 0 && (unsigned)(t - a.trace_t0) < 512u) g_tc_trace[4 * (t - a.trace_t0) + 1] = clock64();
      }
    }
  } else if (warp < TC_EPI_WARPS) {
    // =============================== epilogue: fused top-k' =====================================
    const int quarter = warp & 3, rb = warp >> 2;
    const int row = rb * 128 + quarter * 32 + lane;         // this thread's query row inside the CTA tile
    float* my_s = list_s + row;                             // entry p at my_s[p * 256] (conflict free)
    int32_t* my_i = list_i + row;
    int32_t* my_meta = list_meta + row;
    float* my_pqs = pq_s + row;                             // pending entry q at my_pqs[q * 256]
    int32_t* my_pqi = pq_i + row;
    int npend = 0;
    float thr = -INFINITY;
    unsigned int n_drain = 0;
    const int64_t key_base = (int64_t)tile0 * TC_BN;
    for (int t = 0; t < n_my_tiles; ++t) {
      const int b = t & 1;
      mbar_wait(&bars->tmem_full[b][rb], ((uint32_t)t >> 1) & 1u);
      tc_fence_after();
      if (TC_TRACE_ON(a) && blockIdx.x == 0 && (unsigned)(t - a.trace_t0) < 512u && threadIdx.x == 0) g_tc_trace[4 * (t - a.trace_t0) + 2] = clock64();
      const int64_t tile_key0 = key_base + (int64_t)t * TC_BN;
      const int n_valid = (int)min((int64_t)TC_BN, a.N - tile_key0);
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((b * 2 + rb) * TC_BN);
      // Deferred half of the slow path: fold this row's pending candidates into its list (runs after the TMEM
      // buffer went back to the MMA warp, or when the 4-deep queue is full).
      auto flush = [&]() {
        for (int q = 0; q < npend; ++q) {
          const float sc = my_pqs[q * TC_ROWS];
          if (sc > thr) thr = tc_list_push<KP>(my_s, my_i, my_meta, sc, my_pqi[q * TC_ROWS]);
        }
        npend = 0;
      };
      // One 32-column chunk: max tree vs the row threshold (registers only).  On a hit the lane extracts ITS
      // candidates (bitmask + 31-select, no memory) and appends them to its pending queue: ~150 cycles on the
      // critical path instead of a full list update.
      auto filter = [&](const uint32_t (&v)[32], int col0) {
        float m1[11];
#pragma unroll
        for (int j = 0; j < 10; ++j)
          m1[j] = max3(__uint_as_float(v[3 * j]), __uint_as_float(v[3 * j + 1]), __uint_as_float(v[3 * j + 2]));
        m1[10] = fmaxf(__uint_as_float(v[30]), __uint_as_float(v[31]));
        const float m2a = max3(m1[0], m1[1], m1[2]), m2b = max3(m1[3], m1[4], m1[5]);
        const float m2c = max3(m1[6], m1[7], m1[8]), m2d = fmaxf(m1[9], m1[10]);
        if (fmaxf(max3(m2a, m2b, m2c), m2d) > thr) {        // rare once the list has warmed up
          uint32_t mask = 0;
#pragma unroll
          for (int j = 0; j < 32; ++j) mask |= (__uint_as_float(v[j]) > thr) ? (1u << j) : 0u;
          const int room = n_valid - col0;                  // columns >= n_valid are TMA zero fill past the library end
          if (room < 32) mask &= (room <= 0) ? 0u : ((1u << room) - 1u);
          while (mask) {
            const int j = __ffs(mask) - 1;
            mask &= mask - 1;
            if (npend == TC_PQ) { ++n_drain; flush(); }
            my_pqs[npend * TC_ROWS] = select32(v, j);
            my_pqi[npend * TC_ROWS] = (int32_t)(tile_key0 + col0 + j);
            ++npend;
          }
        }
      };
      uint32_t va[32], vb[32];
      const bool do_filter = (a.debug == 0);
      if (a.debug == 1) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->tmem_empty[b][rb]);
      } else {
        // software pipeline over the 4 chunks of the tile: the load of chunk c+1 is in flight while chunk c is
        // filtered; the TMEM buffer goes back to the MMA warp as soon as the last chunk is in registers
        tmem_ld32(taddr, va);
        tmem_ld_wait_regs(va);
#pragma unroll 1
        for (int c = 0; c < TC_BN / 32; c += 2) {
          tmem_ld32(taddr + (c + 1) * 32, vb);
          if (do_filter) filter(va, c * 32);
          tmem_ld_wait_regs(vb);
          if (c + 2 < TC_BN / 32) {
            tmem_ld32(taddr + (c + 2) * 32, va);
          } else {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->tmem_empty[b][rb]);
          }
          if (do_filter) filter(vb, (c + 1) * 32);
          if (c + 2 < TC_BN / 32) tmem_ld_wait_regs(va);
        }
        if (npend) { ++n_drain; flush(); }
      }
      if (TC_TRACE_ON(a) && blockIdx.x == 0 && (unsigned)(t - a.trace_t0) < 512u && threadIdx.x == 0) { g_tc_trace[4 * (t - a.trace_t0) + 3] = clock64(); g_tc_trace2[2 * (t - a.trace_t0)] = n_drain; n_drain = 0; }
    }
    // ---- publish this split's list --------------------------------------------------------------
    const int64_t grow = (int64_t)qtile * TC_ROWS + row;
    if (grow < a.Q) {
      float* ps = a.part_s + ((int64_t)split * a.Q + grow) * KP;
      int32_t* pi = a.part_i + ((int64_t)split * a.Q + grow) * KP;
#pragma unroll
      for (int p = 0; p < KP; ++p) { ps[p] = my_s[p * TC_ROWS]; pi[p] = my_i[p * TC_ROWS]; }
    }
  }

  // ---- teardown ---------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == TC_EPI_WARPS + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// =====================================================================================================
// Query-stationary variant: the query tile lives in TENSOR MEMORY as the A operand (tcgen05.mma "TS" form).
//
// The SS kernel above re-reads the 256-row query tile from shared memory for every key tile: per 128-key tile
// that is 64 KB of A reads + 64 KB of B reads + 32 KB of TMA writes = 160 KB against the 1024 cycles the tensor
// pipe needs, i.e. 156 B/clk of a 128 B/clk shared-memory port -- the MMA-only ceiling measured at 0.81 of the
// bf16 peak.  Retrieval has a STATIONARY operand (a CTA's queries never change), so here the epilogue warps copy
// their query rows once into TMEM (tcgen05.st, lane = row, 32-bit column j = bf16 elements 2j, 2j+1) and every
// MMA reads A from there: shared-memory traffic drops to 96 KB per tile (94 B/clk), and the 64 KB of shared
// memory the query tile occupied becomes pipeline depth (11 stages).
// TMEM budget (512 columns): A = 2 row blocks x d_pad/2 columns at the top, accumulators = a ring of
// NBUF = (512 - d_pad) / 128 buffers of 128 columns; "use" u = 2*tile + row_block takes buffer u % NBUF.
// Accumulator hand-off barriers are per (buffer, row block): with an odd ring a buffer alternates between the two
// row blocks, and a parity wait is only safe when ONE party consumes every completion of a barrier in order (a
// waiter that arrives a full phase early would otherwise see the stale parity of the previous-but-one completion --
// possible once two issuer warps run skewed).
struct __align__(8) TsBarriers {
  uint64_t a_full;
  uint64_t full[12];
  uint64_t empty[12];
  uint64_t tmem_full[3][2];
  uint64_t tmem_empty[3][2];
  uint32_t tmem_base;
};
constexpr int TS_BAR_BYTES = 512;

__device__ __forceinline__ void tc_mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
               "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                 "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
               : "memory");
}

// ---- selection state of the TS kernel ---------------------------------------------------------------
// Per (query row, key split): a SORTED list of the KP best (score, key) seen so far plus an append buffer of KP pending
// candidates, both in shared memory with entry p of row r at [p * TS_LSTRIDE + r] (stride 257: a whole warp reading ONE
// row is bank-conflict free).
//  * Fast path, registers only: each lane (= query row) reduces its 128 scores of the tile to four chunk maxima,
//    compares with its threshold (thr = the row's KP-th best at its last compaction, or the pre-pass bound) and the
//    warp votes once per tile.
//  * A hit ((lane, 32-score chunk) with a maximum above thr) costs the tile loop only a copy of the chunk into the
//    warp's hit QUEUE (static register indices, independent stores).  A lone warp runs dependent code at ~6 cycles per
//    instruction (measured), so anything serial done here delays the warp's next TMEM drain and, through the
//    accumulator ring, the tensor pipe.
//  * Queued hits are served when the warp would otherwise wait for its next accumulator: all 32 lanes take one score
//    each, compare with the owner row's current threshold, and the candidates append themselves at ballot-derived
//    slots -- no per-lane serial scan, no dynamic register indexing.
//  * When a buffer fills, the warp compacts that row together: every lane ranks one (KP = 16) or two (KP = 32) of the
//    2*KP entries against all others (rank = number of entries that precede it; ties broken by position, so ranks are
//    a permutation), entries ranked below KP are stored back at slot = rank and the new threshold is the entry ranked
//    KP-1.  A frozen threshold admits candidates at the rate it had when it was set, so a row is compacted about once
//    per doubling of the keys seen, instead of one replace-min list update per candidate.
constexpr int TS_LSTRIDE = TC_ROWS + 1;
// append-buffer length for a list of KP entries.  Measured (B200, 100 M x 128, profiles/r2_ab_fmt_kp_100m.jsonl): a 16-entry
// buffer under 32-entry lists buys three more pipeline stages (7 instead of 4) but doubles the compactions -- 107.6 ms vs
// 85.8 ms on Gaussian keys, 121 vs 93.9 ms clustered: selection work, not pipeline depth, is what wide lists cost.
__host__ __device__ constexpr int ts_kpb(int kp) { return kp; }
constexpr int TS_QN = 16;                 // hit-queue entries per epilogue warp
constexpr int TS_QSTRIDE = 36;            // 32 scores + first key index + (owner lane | valid columns << 8), 16 B aligned
constexpr int TS_QUEUE_BYTES = TC_EPI_WARPS * TS_QN * TS_QSTRIDE * 4;

template <int KP>
__device__ __forceinline__ float ts_compact_row(float* ls, int32_t* li, float* ps, int32_t* pi, int row, int npend, int lane) {
  // Every lane reads all 2*KP scores of the row straight from shared memory (broadcast loads: independent, pipelined)
  // instead of pulling them one shuffle at a time -- the shuffle version spent ~60 cycles per entry on latency.
  constexpr unsigned FULL = 0xffffffffu;
  if constexpr (KP == 16) {
    const bool is_list = lane < 16;
    const int slot = lane & 15;
    const bool valid = is_list || slot < npend;
    const float s = valid ? (is_list ? ls : ps)[slot * TS_LSTRIDE + row] : -INFINITY;
    const int32_t id = valid ? (is_list ? li : pi)[slot * TS_LSTRIDE + row] : -1;
    int rank = 0;
#pragma unroll 8
    for (int m = 0; m < 16; ++m) {                          // positions 0..15: the sorted list
      const float a = ls[m * TS_LSTRIDE + row];
      rank += (a > s || (a == s && m < lane)) ? 1 : 0;
    }
#pragma unroll 8
    for (int m = 0; m < 16; ++m) {                          // positions 16..31: the append buffer
      const float a = (m < npend) ? ps[m * TS_LSTRIDE + row] : -INFINITY;
      rank += (a > s || (a == s && 16 + m < lane)) ? 1 : 0;
    }
    __syncwarp();                                           // all reads of the row happen before any write
    if (rank < 16) { ls[rank * TS_LSTRIDE + row] = s; li[rank * TS_LSTRIDE + row] = id; }
    __syncwarp();
    return ls[15 * TS_LSTRIDE + row];                       // the new KP-th best
  } else {
    static_assert(KP == 32, "list lengths: 16 or 32");
    const float s0 = ls[lane * TS_LSTRIDE + row];                       // list entry `lane`     (position lane)
    const int32_t i0 = li[lane * TS_LSTRIDE + row];
    const bool v1 = lane < npend;
    const float s1 = v1 ? ps[lane * TS_LSTRIDE + row] : -INFINITY;      // pending entry `lane`  (position 32 + lane)
    const int32_t i1 = v1 ? pi[lane * TS_LSTRIDE + row] : -1;
    int r0 = 0, r1 = 0;
#pragma unroll 8
    for (int m = 0; m < 32; ++m) {
      const float a0 = ls[m * TS_LSTRIDE + row], a1 = (m < npend) ? ps[m * TS_LSTRIDE + row] : -INFINITY;
      r0 += ((a0 > s0 || (a0 == s0 && m < lane)) ? 1 : 0) + ((a1 > s0) ? 1 : 0);
      r1 += ((a0 >= s1) ? 1 : 0) + ((a1 > s1 || (a1 == s1 && m < lane)) ? 1 : 0);
    }
    __syncwarp();
    if (r0 < 32) { ls[r0 * TS_LSTRIDE + row] = s0; li[r0 * TS_LSTRIDE + row] = i0; }
    if (r1 < 32) { ls[r1 * TS_LSTRIDE + row] = s1; li[r1 * TS_LSTRIDE + row] = i1; }
    __syncwarp();
    return ls[31 * TS_LSTRIDE + row];
  }
}

// Second pass (TcArgs::collect): the threshold is fixed, so a full append buffer is not compacted against a list but
// written out -- all `np` pending (score, key) pairs of CTA row `row` go to the row's global spill area at the slots an
// atomic per-row counter hands out (key splits of the same row append concurrently).  Entries past spill_cap are dropped;
// the counter keeps counting, which is how the refine pass sees the overflow.
__device__ __forceinline__ void ts_spill_row(const float* ps, const int32_t* pi, int row, int np, int lane, const TcArgs& a,
                                             int64_t crow) {
  int base = 0;
  if (lane == 0) base = atomicAdd(a.spill_cnt + crow, np);
  base = __shfl_sync(0xffffffffu, base, 0);
  for (int p = lane; p < np; p += 32) {
    const int o = base + p;
    if (o < a.spill_cap) {
      a.spill_s[crow * a.spill_cap + o] = ps[p * TS_LSTRIDE + row];
      a.spill_i[crow * a.spill_cap + o] = pi[p * TS_LSTRIDE + row];
    }
  }
  __syncwarp();
}

// Serve one queued hit (all 32 lanes): lane j takes score j of the 32-score chunk, compares it with the owner row's
// current threshold, and the candidates append themselves at ballot-derived slots of the owner's buffer; a full buffer
// is compacted on the spot.  Returns the owner's new (threshold, pending count); `owner` = the owner lane.
struct TsServed { float thr; int np; int owner; float pmin; };
// order-preserving float <-> int map (an involution), so that redux.sync's integer minimum is a float minimum
__device__ __forceinline__ int f32_ordered(float f) { const int i = __float_as_int(f); return i ^ ((i >> 31) & 0x7fffffff); }
__device__ __forceinline__ float f32_unordered(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }
template <int KP>
__device__ __forceinline__ TsServed ts_serve_entry(const float* ev, int key0, int meta, float* ls, int32_t* li, float* ps,
                                                   int32_t* pi, int wrow0, float thr, int npend, float pmin, int lane,
                                                   const TcArgs& a, int64_t crow0, int split, float gmargin) {
  constexpr int KPB = ts_kpb(KP);
  const int L = meta & 31, nv = meta >> 8;
  const float val = ev[lane];
  float thr_l = __shfl_sync(0xffffffffu, thr, L);
  int np = __shfl_sync(0xffffffffu, npend, L);
  float pm = __shfl_sync(0xffffffffu, pmin, L);             // smallest score in the owner row's append buffer
  const float mg = __shfl_sync(0xffffffffu, gmargin, L);    // see below: what the shared bound keeps under the k-th best
  bool cand = val > thr_l && lane < nv;                     // columns >= nv are TMA zero fill past the library end
  const int orow = wrow0 + L;
  while (true) {
    const unsigned cm = __ballot_sync(0xffffffffu, cand);
    if (cm == 0) break;
    const int pos = np + __popc(cm & ((1u << lane) - 1u));
    const bool fit = cand && pos < KPB;
    if (fit) {
      ps[pos * TS_LSTRIDE + orow] = val;
      pi[pos * TS_LSTRIDE + orow] = key0 + lane;
      // TcArgs::pool.  The shared bound is (k-th largest published score) - margin and thr_l >= it, so a score <= thr_l +
      // margin cannot raise that k-th largest: only the others are worth a global store (dense clusters: very few)
      if (a.pool && a.g_dbg != 3 && val > thr_l + mg) a.pool[((crow0 + orow) * a.n_splits + split) * KPB + pos] = __float_as_uint(val);
    }
    pm = fminf(pm, f32_unordered(__reduce_min_sync(0xffffffffu, fit ? f32_ordered(val) : 0x7fffffff)));
    np = min(np + __popc(cm), KPB);
    cand = cand && !fit;
    if (__ballot_sync(0xffffffffu, cand) == 0) break;
    __syncwarp();                                           // buffer full with candidates left: compact the row, go on
    if (a.collect) ts_spill_row(ps, pi, orow, KPB, lane, a, crow0 + orow);
    else thr_l = fmaxf(thr_l, ts_compact_row<KP>(ls, li, ps, pi, orow, KPB, lane));   // (a shared bound may sit above the list)
    np = 0;
    pm = INFINITY;
    cand = cand && val > thr_l;
  }
  // Tighten the threshold WITHOUT compacting: the (KP - np) best list entries together with the np pending candidates are
  // KP distinct keys, every one of them >= min(ls[KP-1-np], pm) -- so the row's KP-th best is at least that.  A frozen
  // threshold admits ~KP candidates per doubling of the keys seen; this one rises with every candidate and roughly halves the
  // candidates a row takes over a stream (each costs this warp-serial serve).  An unfilled list holds -inf: no change.
  if (!a.collect && np > 0) {
    const float lsv = (np < KP) ? ls[(KP - 1 - np) * TS_LSTRIDE + orow] : INFINITY;
    thr_l = fmaxf(thr_l, fminf(lsv, pm));
  }
  return TsServed{thr_l, np, L, pm};
}
__device__ __forceinline__ float chunk_max32(const uint32_t (&v)[32]) {
  float m1[11];
#pragma unroll
  for (int j = 0; j < 10; ++j)
    m1[j] = max3(__uint_as_float(v[3 * j]), __uint_as_float(v[3 * j + 1]), __uint_as_float(v[3 * j + 2]));
  m1[10] = fmaxf(__uint_as_float(v[30]), __uint_as_float(v[31]));
  const float m2a = max3(m1[0], m1[1], m1[2]), m2b = max3(m1[3], m1[4], m1[5]);
  const float m2c = max3(m1[6], m1[7], m1[8]), m2d = fmaxf(m1[9], m1[10]);
  return fmaxf(max3(m2a, m2b, m2c), m2d);
}
// one queue entry: 8 x 128-bit stores of the chunk (static register indices) + one 64-bit header store
__device__ __forceinline__ void queue_put(float* entry, const uint32_t (&v)[32], int key0, int meta) {
  uint4* e4 = reinterpret_cast<uint4*>(entry);
#pragma unroll
  for (int j = 0; j < 8; ++j) e4[j] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  *reinterpret_cast<int2*>(entry + 32) = make_int2(key0, meta);
}

// ---- cross-split threshold sharing: the sweep run by the CTAs past the worker grid (TcArgs::pool) -------------------
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.b32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_relaxed_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// One warp per query row and sweep: the row's n_splits * pool_slots published words (<= 32 * NREG) go to registers, the
// k-th largest is found by removing the maximum (one instance: equal scores of distinct keys count separately) k times.
// A lone warp runs this dependent code at 6-15 cycles per instruction, and the sweep rate is what the workers feel (measured,
// 12.5 M x 128: 1 / 2 / 4 sweeping CTAs -> 716 k / 543 k / 452 k hits, 9.80 / 9.63 / 9.07 ms), so a warp works on TWO rows
// at a time (two independent dependency chains) and skips a row whose words are unchanged since its last visit (XOR
// checksum kept in shared memory; late in the stream almost every row).
// The workers never wait for this: a late, stale or missing gthr only costs them hits, never correctness, and the loop
// ends when every worker CTA has counted itself in *done (or at once, when the workers already finished).
template <int NREG>
__device__ void ts_merger_loop(const TcArgs& a, int n_workers, uint32_t* sums, int sums_per_warp) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int words = a.n_splits * a.pool_slots;
  const int64_t first = (int64_t)(blockIdx.x - n_workers) * nw + warp, stride = (int64_t)a.n_mergers * nw;
  const float kerr = a.g_kerr ? __ldg(a.g_kerr) : 0.f;
  uint32_t* my_sums = sums + (size_t)warp * sums_per_warp;
  for (int i = lane; i < sums_per_warp; i += 32) my_sums[i] = 0u;      // (an all-empty row XORs to 0: nothing to do for it)
  __syncwarp();
  const unsigned long long t_begin = clock64();
  // The launch ends with its last CTA, so a sweep must not outlive the workers: the worker count is polled at every row pair
  // (measured: with ONE sweeping CTA a full sweep takes ~1 ms, and checking only between sweeps added half of that to every call).
  bool finished = false;
  while (!finished) {
    int n = 0;
    for (int64_t row = first; row < a.Q; row += 2 * stride, n += 2) {
      if ((int)ld_relaxed_u32(reinterpret_cast<const uint32_t*>(a.done)) >= n_workers) { finished = true; break; }
      const int64_t rows[2] = {row, row + stride};
      float v[2][NREG];
      bool live[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const bool in = rows[r] < a.Q;
        const uint32_t* pr = a.pool + rows[r] * words;
        uint32_t x = 0u;
#pragma unroll
        for (int t = 0; t < NREG; ++t) {
          const int w = lane + 32 * t;
          const uint32_t bits = (in && w < words) ? ld_relaxed_u32(pr + w) : 0u;
          v[r][t] = bits ? __uint_as_float(bits) : -INFINITY;
          x ^= bits * (uint32_t)(2 * w + 1);                       // position-dependent, so that a moved value counts
        }
        x = __reduce_xor_sync(0xffffffffu, x);
        live[r] = in;
        if (in && n + r < sums_per_warp) {
          live[r] = my_sums[n + r] != x;
          __syncwarp();
          if (lane == 0) my_sums[n + r] = x;
        }
      }
      if (!live[0] && !live[1]) continue;
      float kth[2] = {-INFINITY, -INFINITY};
      for (int it = 0; it < a.g_k; ++it) {
        float m[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          m[r] = v[r][0];
#pragma unroll
          for (int t = 1; t < NREG; ++t) m[r] = fmaxf(m[r], v[r][t]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          m[0] = fmaxf(m[0], __shfl_xor_sync(0xffffffffu, m[0], o));
          m[1] = fmaxf(m[1], __shfl_xor_sync(0xffffffffu, m[1], o));
        }
        kth[0] = m[0]; kth[1] = m[1];
        if (m[0] == -INFINITY && m[1] == -INFINITY) break;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          int mine = -1;
#pragma unroll
          for (int t = NREG - 1; t >= 0; --t) mine = (v[r][t] == m[r] && m[r] > -INFINITY) ? t : mine;
          const unsigned who = __ballot_sync(0xffffffffu, mine >= 0);
          if (who != 0u && lane == __ffs(who) - 1) {
#pragma unroll
            for (int t = 0; t < NREG; ++t) if (t == mine) v[r][t] = -INFINITY;
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        if (lane == 0 && live[r] && kth[r] > -INFINITY) {
          // exact modes: refine needs every key whose 16-bit score is within 2 eps of the k-th best (TcArgs::pool); raw
          // modes: one notch below, so that keys tying with the k-th best are still taken
          float g;
          if (a.g_exact) g = kth[r] - (2.0f * ((a.g_qerr ? __ldg(a.g_qerr + rows[r]) : 0.f) + kerr + a.g_eps_fixed) + 1e-6f);
          else g = (kth[r] > 0.f) ? kth[r] * (1.0f - 1e-6f) : kth[r] * (1.0f + 1e-6f) - 1e-30f;
          const uint32_t old = a.gthr[rows[r]];                   // only this warp ever writes the word
          if (g != 0.f && (old == 0u || g > __uint_as_float(old))) st_relaxed_u32(a.gthr + rows[r], __float_as_uint(g));
        }
      }
    }
    if (first >= a.Q && (int)ld_relaxed_u32(reinterpret_cast<const uint32_t*>(a.done)) >= n_workers) finished = true;   // (a warp without rows)
    // (a sweep that has run for ~10 s -- a debugger or sanitizer slowing the launch down 100x -- simply leaves: the workers
    //  never depend on it, so giving up costs hits, not correctness; a trap here would kill a healthy launch)
    if (clock64() - t_begin > TC_TIMEOUT_CYCLES) break;
  }
}

// Epilogue: a warp pulls its whole 32 x 128 accumulator slice into registers (4 x tcgen05.ld.x32, one wait) and hands
// the TMEM buffer back BEFORE it filters, so the MMA issuers never wait on selection work.
// PRE = true: threshold pre-pass instantiation (group maxima only, see TcArgs::premax)
template <int KH, int NSTAGE, int KP, bool PRE>
__global__ void __launch_bounds__(TS_THREADS, 1)
cosine_topk_ts_kernel(const __grid_constant__ CUtensorMap map_k, const uint16_t* __restrict__ q_bf, const TcArgs a) {
  constexpr int A_COLS = 2 * KH * 32;                 // 32-bit TMEM columns of the resident query tile
  constexpr int NBUF = (512 - A_COLS) / TC_BN;        // accumulator ring: 3 (d <= 128) or 2 (d <= 256)
  constexpr uint32_t A_COL0 = NBUF * TC_BN;
  static_assert(NBUF >= 2 && NBUF <= 3, "accumulator ring");
  // Two issuer warps.  Ring of 3 (d <= 128): they take alternate key tiles.  Ring of 2 (d > 128: one buffer per row
  // block): consecutive tiles of a row block are serialised through ONE barrier pair, which only a single in-order party
  // may follow by parity -- so there the issuers split by ROW BLOCK (issuer r issues row block r of every tile and is the
  // only producer of tmem_full[r] / consumer of tmem_empty[r]); a key stage is released when BOTH have committed their
  // MMAs on it (empty barriers count 2).  Measured (r2_variant_ab.jsonl, 10 M x 256, epilogue disabled): one issuer has 32
  // MMAs + 6 commits + 6 waits per 2048-cycle tile to get through and ran the pipe at 1 439 TFLOP/s against 1 631 at d = 128.
  constexpr int NISS = 2;
  constexpr bool ROWSPLIT = (NBUF == 2);
  static_assert(NSTAGE >= 2 * KH && NSTAGE <= 12, "row-block-major MMA order holds a tile's stages for the whole tile");
  // Offsets are taken from the extern array itself (no integer round trip), so every list / queue access below stays
  // in the shared address space (LDS/STS, not generic LD/ST).  The dynamic window starts 1 KB-aligned when the kernel
  // has no static shared memory; checked once below.
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  // carve: [B: NSTAGE boxes][sorted lists: KP x 257 x (f32 + i32)][append buffers: same][warp hit queues][barriers]
  unsigned char* sB = smem_dyn;
  float* list_s = reinterpret_cast<float*>(smem_dyn + NSTAGE * TC_BOX_BYTES);
  int32_t* list_i = reinterpret_cast<int32_t*>(list_s + KP * TS_LSTRIDE);
  constexpr int KPB = ts_kpb(KP);
  float* pq_s = reinterpret_cast<float*>(list_i + KP * TS_LSTRIDE);
  int32_t* pq_i = reinterpret_cast<int32_t*>(pq_s + KPB * TS_LSTRIDE);
  float* queue = reinterpret_cast<float*>(pq_i + KPB * TS_LSTRIDE);       // [8 warps][TS_QN][TS_QSTRIDE]
  constexpr int BAR_OFF = (NSTAGE * TC_BOX_BYTES + 2 * (KP + KPB) * TS_LSTRIDE * 4 + TS_QUEUE_BYTES + 15) / 16 * 16;
  TsBarriers* bars = reinterpret_cast<TsBarriers*>(smem_dyn + BAR_OFF);
  if (threadIdx.x == 0 && (smem_u32(smem_dyn) & 1023u) != 0) __trap();     // TMA SWIZZLE_128B boxes need 1 KB alignment

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // (query tile, key split) geometry: from the host plan, or -- second pass -- from the device-side row count
  int n_qtiles = a.n_qtiles, tiles_per_split = a.tiles_per_split;
  int64_t Qn = a.Q;
  if (!PRE && a.n_rows_dev) {
    const int n = *a.n_rows_dev;                            // written by the refine kernel earlier on this stream
    if (n <= 0) return;                                     // (uniform: the whole CTA leaves before any barrier / TMEM)
    n_qtiles = (n + TC_ROWS - 1) / TC_ROWS;
    int S = (int)gridDim.x / n_qtiles;                      // the launch has >= max(#SMs, worst-case query tiles) CTAs
    S = max(1, min(S, a.n_tiles));
    tiles_per_split = (a.n_tiles + S - 1) / S;
    const int n_splits = (a.n_tiles + tiles_per_split - 1) / tiles_per_split;
    if ((int)blockIdx.x >= n_qtiles * n_splits) return;
    Qn = n;
  }
  if constexpr (!PRE) {
    if (a.n_mergers > 0 && (int)blockIdx.x >= a.n_qtiles * a.n_splits) {     // (uniform per CTA, before any barrier / TMEM)
      if (a.g_dbg == 1) return;
      uint32_t* sums = reinterpret_cast<uint32_t*>(smem_dyn);                 // row checksums: all of the dynamic window
      const int spw = (int)((size_t)(NSTAGE * TC_BOX_BYTES) / 4 / (TS_THREADS / 32));
      if (a.n_splits * a.pool_slots <= 160) ts_merger_loop<5>(a, a.n_qtiles * a.n_splits, sums, spw);
      else ts_merger_loop<16>(a, a.n_qtiles * a.n_splits, sums, spw);
      return;
    }
  }
  const int qtile = blockIdx.x % n_qtiles;
  const int split = blockIdx.x / n_qtiles;
  const int tile0 = split * tiles_per_split;
  const int tile1 = min(tile0 + tiles_per_split, a.n_tiles);
  const int n_my_tiles = PRE ? min(max(tile1 - tile0, 0), a.pre_tiles) : max(tile1 - tile0, 0);

  // ---- one-time setup ---------------------------------------------------------------------------
  if (warp == TC_EPI_WARPS && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_k)) : "memory");
    mbar_init(&bars->a_full, TC_EPI_WARPS);
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], ROWSPLIT ? 2 : 1); }
    for (int b = 0; b < NBUF; ++b)
      for (int r = 0; r < 2; ++r) { mbar_init(&bars->tmem_full[b][r], 1); mbar_init(&bars->tmem_empty[b][r], TC_EPI_WARPS / 2); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == TC_EPI_WARPS + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < KP * TS_LSTRIDE; i += TS_THREADS) { list_s[i] = -INFINITY; list_i[i] = -1; }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == TC_EPI_WARPS) {
    // =============================== TMA producer: the key stream ===============================
    if (elect_one()) {
      int s = 0; uint32_t ph = 0;
      for (int t = 0; t < n_my_tiles; ++t) {
        for (int kh = 0; kh < KH; ++kh) {
          mbar_wait(&bars->empty[s], ph ^ 1);
          if (a.no_tma) { mbar_arrive(&bars->full[s]); }
          else {
            mbar_expect_tx(&bars->full[s], TC_BOX_BYTES);
            tma_load_2d(sB + s * TC_BOX_BYTES, &map_k, kh * 64, (tile0 + t) * TC_BN, &bars->full[s]);
          }
          if (++s == NSTAGE) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp >= TC_EPI_WARPS + 1) {
    // =============================== MMA issuers (A from TMEM) ==================================
    // The issuing thread's own latencies (four mbarrier waits, fences and four commits per tile, ~1300 cycles
    // measured) exceed the 1024 cycles the tensor pipe needs for a d=128 tile, so ONE issuer leaves the pipe idle a
    // quarter of the time.  Two issuer warps take alternate key tiles (mw = 0: even, 1: odd); each walks the stage
    // ring and the accumulator ring in steps of two tiles.  tcgen05.commit tracks the MMAs of the executing thread,
    // so every barrier is still completed by exactly one commit.
    const int mw = warp - (TC_EPI_WARPS + 1);
    if (mw < NISS && elect_one()) {
      mbar_wait(&bars->a_full, 0);
      tc_fence_after();
      const uint32_t b_addr = smem_u32(sB);
      const uint32_t idesc = a.idesc;
      int s = ROWSPLIT ? 0 : mw * KH; uint32_t ph = 0;
      while (s >= NSTAGE) { s -= NSTAGE; ph ^= 1; }
      for (int t = ROWSPLIT ? 0 : mw; t < n_my_tiles; t += ROWSPLIT ? 1 : NISS) {
        const bool tr3 = TC_TRACE_ON(a) && blockIdx.x == 0 && (unsigned)(t - a.trace_t0) < 256u;
        unsigned int* t3 = g_tc_trace3 + 16 * (tr3 ? (t - a.trace_t0) : 0);
        if (tr3) t3[0] = (unsigned int)clock64();
        // The tile's key boxes land long before its accumulators are free: take those waits FIRST, so that the critical
        // hand-off (epilogue releases a buffer -> first MMA of the next use enters the pipe) is one barrier wake-up plus
        // the MMA issue, not three wake-ups.
        {
          int sk = s; uint32_t phk = ph;
          for (int kh = 0; kh < KH; ++kh) {
            mbar_wait(&bars->full[sk], phk);
            if (++sk == NSTAGE) { sk = 0; phk ^= 1; }
          }
          if (tr3) t3[2] = (unsigned int)clock64();
        }
        for (int rb = ROWSPLIT ? mw : 0; rb < (ROWSPLIT ? mw + 1 : 2); ++rb) {
          // use u = 2t + rb takes buffer u % NBUF; its previous user was use u - NBUF (row block (u - NBUF) & 1,
          // that row block's k'-th visit of the buffer): wait until its epilogue has drained it
          const uint32_t u = 2u * (uint32_t)t + (uint32_t)rb;
          const uint32_t buf = (NBUF == 2) ? (u & 1u) : (u % 3u);
          if (u >= (uint32_t)NBUF) {
            const uint32_t up = u - NBUF, tp = up >> 1;
            mbar_wait(&bars->tmem_empty[buf][up & 1u], ((NBUF == 2) ? tp : tp / 3u) & 1u);
          }
          tc_fence_after();
          if (tr3) t3[rb == 0 ? 1 : 7] = (unsigned int)clock64();
          if (TC_TRACE_ON(a) && blockIdx.x == 0 && (unsigned)(t - a.trace_t0) < 512u && rb == 0) g_tc_trace[4 * (t - a.trace_t0) + 0] = clock64();
          int sk = s;
          for (int kh = 0; kh < KH; ++kh) {
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              const uint64_t db = umma_desc(b_addr + sk * TC_BOX_BYTES + k4 * 32);
              tc_mma_bf16_ts(tmem_base + (uint32_t)(buf * TC_BN), tmem_base + A_COL0 + (uint32_t)((rb * KH + kh) * 32 + k4 * 8),
                             db, idesc, (kh | k4) != 0 ? 1u : 0u);
            }
            if (tr3 && rb == 0 && kh < 2) t3[3 + 2 * kh] = (unsigned int)clock64();
            if (++sk == NSTAGE) sk = 0;
          }
          if (tr3 && rb == 1) t3[8] = (unsigned int)clock64();
          tc_commit(&bars->tmem_full[buf][rb]);             // this row block's accumulator is complete
          if (tr3) t3[rb == 0 ? 6 : 9] = (unsigned int)clock64();
        }
        for (int kh = 0; kh < KH; ++kh) {                   // smem stages reusable once all 2*KH*4 MMAs retire
          tc_commit(&bars->empty[s]);
          if (++s == NSTAGE) { s = 0; ph ^= 1; }
        }
        if (tr3) t3[10] = (unsigned int)clock64();
        if (TC_TRACE_ON(a) && blockIdx.x == 0 && (unsigned)(t - a.trace_t0) < 512u) g_tc_trace[4 * (t - a.trace_t0) + 1] = clock64();
        // hop over the other issuer's tile: KH stages
        if (!ROWSPLIT) {
          s += KH;
          if (s >= NSTAGE) { s -= NSTAGE; ph ^= 1; }
        }
      }
    }
  } else if (warp < TC_EPI_WARPS) {
    // =============================== epilogue warps ==============================================
    const int quarter = warp & 3, rb = warp >> 2;
    const int row = rb * 128 + quarter * 32 + lane;         // this thread's query row inside the CTA tile
    const int64_t grow = (int64_t)qtile * TC_ROWS + row;
    // ---- one-time: this thread's normalised bf16 query row -> TMEM (the stationary A operand) ----
    {
      const int64_t srow = (!PRE && a.row_map && grow < Qn) ? (int64_t)a.row_map[grow] : grow;   // second pass: compact rows
      const uint4* src = reinterpret_cast<const uint4*>(q_bf + srow * (int64_t)(KH * 64));
      const uint32_t a_taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + A_COL0 + (uint32_t)(rb * KH * 32);
#pragma unroll 1
      for (int c = 0; c < KH * 2; ++c) {                    // 16 columns = 32 bf16 per step
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint4 x = (grow < Qn) ? __ldg(src + c * 4 + i) : make_uint4(0u, 0u, 0u, 0u);
          w[4 * i] = x.x; w[4 * i + 1] = x.y; w[4 * i + 2] = x.z; w[4 * i + 3] = x.w;
        }
        if (a.swap_halves) {
#pragma unroll
          for (int i = 0; i < 16; ++i) w[i] = __byte_perm(w[i], 0u, 0x1032);
        }
        tmem_st16(a_taddr + c * 16, w);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->a_full);
    }
    // ---- fused top-k' ------------------------------------------------------------------------------
    const int wrow0 = rb * 128 + quarter * 32;              // first row of this warp (lane L owns row wrow0 + L)
    int npend = 0;                                          // candidates in this row's append buffer
    // KP-th best score at the last compaction of this row; starts at the pre-pass bound (one notch lower, so that keys
    // tying with the bound are still taken) or -inf
    float thr = (!PRE && a.thr0 && grow < Qn) ? a.thr0[grow] : -INFINITY;
    if (thr > -INFINITY && !a.thr_exact) thr = (thr > 0.f) ? thr * (1.0f - 1e-6f) : thr * (1.0f + 1e-6f) - 1e-30f;
    if (grow >= Qn) thr = INFINITY;                         // padding rows of the last query tile never produce a hit
    if (a.debug == 4) thr = INFINITY;                       // RAG_DIAG experiment: filter runs, no hit is ever taken
    const int64_t crow0 = (int64_t)qtile * TC_ROWS;         // compact (second pass) / plain global row of CTA row 0
    // pre-pass state: running maximum of the current tile group
    float pmin = INFINITY;                                  // smallest score among this row's pending candidates
    float gmargin = 0.f;                                    // cross-split sharing: k-th best minus this = the shared bound
    float gm = -INFINITY;
    int gi = 0;
    const int g_tiles = PRE ? (n_my_tiles + a.pre_groups - 1) / a.pre_groups : 0;
    int g_end = g_tiles;
    const int key_base = tile0 * TC_BN;                     // 32-bit: a shard holds < 2^31 keys (checked by the C ABI)
    // ---- this warp's hit queue: (lane, chunk) hits wait here until the warp has nothing better to do ----
    float* q_mine = queue + warp * (TS_QN * TS_QSTRIDE);
    unsigned qhead = 0, qtail = 0;                          // warp-uniform counters; slot = counter % TS_QN
    // all lanes: serve the oldest queued hit (ts_serve_entry)
    auto process_one = [&]() {
      const float* ev = q_mine + (qhead % TS_QN) * TS_QSTRIDE;
      const TsServed r = ts_serve_entry<KP>(ev, __float_as_int(ev[32]), __float_as_int(ev[33]), list_s, list_i, pq_s, pq_i, wrow0,
                                            thr, npend, pmin, lane, a, crow0, split, gmargin);
      if (lane == r.owner) { npend = r.np; thr = r.thr; pmin = r.pmin; }
      ++qhead;
      __syncwarp();
    };
    // cross-split bound (TcArgs::pool): every 16th tile the value read 16 tiles earlier is folded in and the next read is
    // issued -- ONE predictable branch per tile and no wait on the load.  (Measured at 100 M keys: a read every 4th tile
    // consumed one tile later cost the epilogue warps 1.9 % of the step -- every instruction in this loop is on the
    // critical path of a lone warp.)  Padding rows keep thr = +inf.
    uint32_t gbits = 0u;
    const bool g_on = !PRE && a.gthr != nullptr && grow < Qn && a.g_dbg != 2;
    if (g_on && a.g_exact) gmargin = 2.0f * ((a.g_qerr ? __ldg(a.g_qerr + grow) : 0.f) + (a.g_kerr ? __ldg(a.g_kerr) : 0.f) + a.g_eps_fixed) + 1e-6f;
    for (int t = 0; t <= n_my_tiles; ++t) {
      if ((t & 15) == 0 && g_on) {
        if (gbits) thr = fmaxf(thr, __uint_as_float(gbits));
        gbits = ld_relaxed_u32(a.gthr + grow);
      }
      // use u = 2t + rb -> buffer u % NBUF; this row block's (t / 3)-th visit of it (ring of 3), t-th (ring of 2)
      const uint32_t u = 2u * (uint32_t)t + (uint32_t)rb;
      const uint32_t buf = (NBUF == 2) ? (u & 1u) : (u % 3u);
      const uint32_t full_par = ((NBUF == 2) ? (uint32_t)t : (uint32_t)t / 3u) & 1u;
      // idle time is selection time: while the next accumulator is not ready, work off queued hits; the extra pass
      // t == n_my_tiles drains what is left (one call site keeps the kernel small)
      if (qhead != qtail) {
        // Keep half of the queue free before the accumulator is pulled into registers: a tile's hits then always fit, and
        // the parking path below (~2 300 cycles per hit: 16 KB to global memory and back) stays the rare case.  Measured
        // (profiles/r2_ts_trace_12m5_before.txt): on a 10 851-tile stream the queue sat at 14-16 of 16 entries for thousands of
        // tiles -- every hit overflowed, the warp fell further behind and never found the idle time that drains the queue.
        while (qtail - qhead > (unsigned)(TS_QN / 2)) process_one();
        const uint32_t full_addr = smem_u32(&bars->tmem_full[buf][rb]);
        while (qhead != qtail && (t == n_my_tiles || !__any_sync(0xffffffffu, mbar_test(full_addr, full_par)))) process_one();
      }
      if (t == n_my_tiles) break;
      mbar_wait(&bars->tmem_full[buf][rb], full_par);
      tc_fence_after();
      if (TC_TRACE_ON(a) && blockIdx.x == 0 && (unsigned)(t - a.trace_t0) < 512u && threadIdx.x == 0) g_tc_trace[4 * (t - a.trace_t0) + 2] = clock64();
      const int tile_key0 = key_base + t * TC_BN;
      const int n_valid = min(TC_BN, (int)a.N - tile_key0);
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * TC_BN);
      bool hit = false;
      if (a.debug == 1) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->tmem_empty[buf][rb]);
      } else {
        // The warp pulls its whole 32 x 128 slice into registers and hands the TMEM buffer back BEFORE it looks at a
        // single score: the MMA issuers wait for this release, and every cycle between "accumulator complete" and
        // "buffer free" comes out of the one tile of slack the 3-deep accumulator ring provides.
        uint32_t va[32], vb[32], vc[32], vd[32];
        tmem_ld32(taddr, va); tmem_ld32(taddr + 32, vb); tmem_ld32(taddr + 64, vc); tmem_ld32(taddr + 96, vd);
        tmem_ld_wait();
        tie_regs(va); tie_regs(vb); tie_regs(vc); tie_regs(vd);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->tmem_empty[buf][rb]);
        if constexpr (PRE) {
          // pre-pass: group maxima only (no lists, no slow path); a tile that runs past the library end is skipped
          const float c0 = chunk_max32(va), c1 = chunk_max32(vb), c2 = chunk_max32(vc), c3 = chunk_max32(vd);
          if (n_valid == TC_BN) gm = fmaxf(gm, fmaxf(fmaxf(c0, c1), fmaxf(c2, c3)));
          if (t + 1 == g_end || t + 1 == n_my_tiles) {
            if (grow < a.Q) a.gmax[((int64_t)split * a.pre_groups + gi) * a.Q + grow] = gm;
            gm = -INFINITY; ++gi; g_end += g_tiles;
          }
        } else if (a.debug == 0 || a.debug == 4) {
          // fast path (registers only): four chunk maxima against the row threshold, ONE warp vote per tile
          const float c0 = chunk_max32(va), c1 = chunk_max32(vb), c2 = chunk_max32(vc), c3 = chunk_max32(vd);
          const bool h0 = c0 > thr, h1 = c1 > thr, h2 = c2 > thr, h3 = c3 > thr;
          hit = h0 || h1 || h2 || h3;
          if (__any_sync(0xffffffffu, hit)) {
            // A hit costs the tile loop only the copy of its 32-score chunk into the queue (8 x 128-bit stores with static
            // register indices + a header) at the slot its ballot rank says; everything serial happens in process_one.
            const unsigned b0 = __ballot_sync(0xffffffffu, h0), b1 = __ballot_sync(0xffffffffu, h1);
            const unsigned b2 = __ballot_sync(0xffffffffu, h2), b3 = __ballot_sync(0xffffffffu, h3);
            const int n0 = __popc(b0), n1 = __popc(b1), n2 = __popc(b2), n3 = __popc(b3);
            const int nv3 = max(n_valid - 96, 0);           // only the library's last tile can be short
            if (__builtin_expect(qtail - qhead + (unsigned)(n0 + n1 + n2 + n3) <= (unsigned)TS_QN, 1)) {
              const unsigned lt = (1u << lane) - 1u;
              if (h0) queue_put(q_mine + ((qtail + __popc(b0 & lt)) % TS_QN) * TS_QSTRIDE, va, tile_key0, lane | (min(32, n_valid) << 8));
              if (h1) queue_put(q_mine + ((qtail + n0 + __popc(b1 & lt)) % TS_QN) * TS_QSTRIDE, vb, tile_key0 + 32, lane | (max(min(32, n_valid - 32), 0) << 8));
              if (h2) queue_put(q_mine + ((qtail + n0 + n1 + __popc(b2 & lt)) % TS_QN) * TS_QSTRIDE, vc, tile_key0 + 64, lane | (max(min(32, n_valid - 64), 0) << 8));
              if (h3) queue_put(q_mine + ((qtail + n0 + n1 + n2 + __popc(b3 & lt)) % TS_QN) * TS_QSTRIDE, vd, tile_key0 + 96, lane | (min(32, nv3) << 8));
              qtail += n0 + n1 + n2 + n3;
              __syncwarp();
            } else {
              // Queue overflow (dense phases only: no pre-pass bound yet, tiny or adversarial libraries).  Every lane
              // parks its four chunks in this warp's private global-memory area -- after that no score register is live
              // and the hits are served straight from the area, one (lane, chunk) at a time.
              float* ov = a.overflow + ((size_t)blockIdx.x * TC_EPI_WARPS + warp) * (32 * TC_BN);
              uint4* mine = reinterpret_cast<uint4*>(ov + lane * TC_BN);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                mine[j] = make_uint4(va[4 * j], va[4 * j + 1], va[4 * j + 2], va[4 * j + 3]);
                mine[8 + j] = make_uint4(vb[4 * j], vb[4 * j + 1], vb[4 * j + 2], vb[4 * j + 3]);
                mine[16 + j] = make_uint4(vc[4 * j], vc[4 * j + 1], vc[4 * j + 2], vc[4 * j + 3]);
                mine[24 + j] = make_uint4(vd[4 * j], vd[4 * j + 1], vd[4 * j + 2], vd[4 * j + 3]);
              }
              __syncwarp();
#pragma unroll 1
              for (int c = 0; c < 4; ++c) {
                unsigned m = (c == 0) ? b0 : (c == 1) ? b1 : (c == 2) ? b2 : b3;
                const int nvc = max(min(32, n_valid - c * 32), 0);
                while (m) {
                  const int L = __ffs(m) - 1;
                  m &= m - 1;
                  const TsServed r = ts_serve_entry<KP>(ov + L * TC_BN + c * 32, tile_key0 + c * 32, L | (nvc << 8), list_s, list_i,
                                                        pq_s, pq_i, wrow0, thr, npend, pmin, lane, a, crow0, split, gmargin);
                  if (lane == r.owner) { npend = r.np; thr = r.thr; pmin = r.pmin; }
                  __syncwarp();
                }
              }
            }
          }
        }
      }
      if (TC_TRACE_ON(a) && blockIdx.x == 0 && (unsigned)(t - a.trace_t0) < 512u && warp == 0) {
        const unsigned hb = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) {
          g_tc_trace[4 * (t - a.trace_t0) + 3] = clock64();
          g_tc_trace2[2 * (t - a.trace_t0)] = qtail - qhead; g_tc_trace2[2 * (t - a.trace_t0) + 1] = __popc(hb);
        }
      }
    }
    if constexpr (PRE) {
      for (; gi < a.pre_groups; ++gi)
        if (grow < a.Q) a.gmax[((int64_t)split * a.pre_groups + gi) * a.Q + grow] = -INFINITY;
    }
    // ---- fold what is still pending, publish this split's list (sorted: score desc) ---------------------
    if constexpr (!PRE) {
      unsigned todo = __ballot_sync(0xffffffffu, npend > 0);
      while (todo) {
        const int L = __ffs(todo) - 1;
        todo &= todo - 1;
        const int np = __shfl_sync(0xffffffffu, npend, L);
        if (a.collect) ts_spill_row(pq_s, pq_i, wrow0 + L, np, lane, a, crow0 + wrow0 + L);
        else ts_compact_row<KP>(list_s, list_i, pq_s, pq_i, wrow0 + L, np, lane);
      }
    }
    if (!PRE && a.hit_count && lane == 0) atomicAdd(a.hit_count, qtail);
    if (!PRE && !a.collect && grow < a.Q) {
      float* ps = a.part_s + ((int64_t)split * a.Q + grow) * KP;
      int32_t* pi = a.part_i + ((int64_t)split * a.Q + grow) * KP;
#pragma unroll
      for (int p = 0; p < KP; ++p) { ps[p] = list_s[p * TS_LSTRIDE + row]; pi[p] = list_i[p * TS_LSTRIDE + row]; }
      if (a.part_thr) a.part_thr[(int64_t)split * a.Q + grow] = thr;
    }
  }

  // ---- teardown ---------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == TC_EPI_WARPS + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
  if (!PRE && a.n_mergers > 0 && threadIdx.x == 0) atomicAdd(a.done, 1);   // lets the sweeping CTAs go (TcArgs::pool)
}

// ---- pre-pass threshold: thr0[row] = kp-th largest of the row's G group maxima ---------------------------
// Each group maximum is the bf16 score of a distinct key of this shard, so at least kp keys score >= thr0: the final
// kp-th best score of the row (over the whole shard, hence over the union of the split lists) is >= thr0.
// margin = 1 (two-pass mode, below): the result is lowered by 2 eps_row + 1e-6, which makes it a COLLECT threshold -- at least kp
// distinct keys score >= kth in 16 bits, so the kp-th best exact score is >= kth - eps and every key that can be among the
// exact top kp scores > kth - 2 eps in 16 bits.
__global__ void __launch_bounds__(256) sample_threshold_kernel(const float* __restrict__ gmax, int G, int64_t Q, int kp,
                                                               float* __restrict__ thr0, int margin = 0,
                                                               const float* __restrict__ qerr = nullptr,
                                                               const float* __restrict__ kerr_max = nullptr,
                                                               float eps_fixed = 0.f) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= Q) return;
  float v[8];                                                 // G <= 256
#pragma unroll
  for (int t = 0; t < 8; ++t) { const int g = lane + 32 * t; v[t] = (g < G) ? __ldg(gmax + (int64_t)g * Q + row) : -INFINITY; }
  float kth = -INFINITY;
  for (int it = 0; it < kp; ++it) {                           // remove the current maximum (one instance) kp times
    float m = v[0];
#pragma unroll
    for (int t = 1; t < 8; ++t) m = fmaxf(m, v[t]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    kth = m;
    if (m == -INFINITY) break;
    int mine = -1;
#pragma unroll
    for (int t = 7; t >= 0; --t) mine = (v[t] == m) ? t : mine;
    const unsigned who = __ballot_sync(0xffffffffu, mine >= 0);
    if (lane == __ffs(who) - 1) {
#pragma unroll
      for (int t = 0; t < 8; ++t) if (t == mine) v[t] = -INFINITY;
    }
  }
  if (margin && kth > -INFINITY)
    kth -= 2.0f * ((qerr ? __ldg(qerr + row) : 0.f) + (kerr_max ? __ldg(kerr_max) : 0.f) + eps_fixed) + 1e-6f;
  if (lane == 0) thr0[row] = kth;
}

// ---- refine: exact fp32 re-score of the candidates + certificate ------------------------------------
struct RefineArgs {
  const float* q; const float* keys; const float* q_inv_norm; const float* key_inv_norm;
  int64_t Q; int64_t N; int d; int k; int kp; int n_splits;
  const float* part_s; const int32_t* part_i;
  int exact;                         // 1: fp32 re-score + certificate, 0: keep bf16 scores (RAG_SIM_BF16)
  int64_t idx_offset;
  float* out_scores; int64_t* out_idx;
  int32_t* fb_rows; int32_t* fb_count;   // uncertified rows (exact only)
  const float* thr0;                     // nullable: pre-pass bound; keys never listed score <= thr0[row] (16-bit score)
  const float* part_thr;                 // nullable: [n_splits][Q] threshold each split ended with (cross-split sharing)
  // Exclusion lists (nullable): query row r never returns the GLOBAL key indices mask_col[mask_rowptr[r] .. mask_rowptr[r+1])
  // (the evaluation ranking's history mask, RAGraph_edge/utils/metrics.py:48-53,110-117).  The filter knows nothing about
  // them: excluded candidates are dropped here, the certificate still bounds every key outside the lists, and a row whose
  // lists hold fewer than k admissible candidates simply fails it and takes the second pass.
  const int64_t* mask_rowptr; const int64_t* mask_col;
  // Dot-product ranking (RAG_SIM_DOT): the caller's shadow holds keys * key_scale (row norms <= 1), the queries are
  // normalised like in the cosine modes, so filter and certificate work in units of q_hat . (key_scale * k); key_scale > 0
  // replaces the per-key inverse norms in the exact re-score and the returned scores are scaled back to q . k.
  float key_scale;
  // error bound of the 16-bit scores, per row: qerr[row] (nullable) + *kerr_max (nullable) + eps_fixed
  const float* qerr; const float* kerr_max; float eps_fixed;
  float* thr2;                           // per uncertified row (same slot as fb_rows): threshold of the second pass
  // diagnostics for the caller's format policy: rows whose certificate margin would not survive an error bound
  // loose_mult times larger (fp16 -> bf16 is 8x) are counted in *loose_count
  float loose_mult; int32_t* loose_count;
};

constexpr int REFINE_SEL_CAP = 64;        // candidates compacted per round of the re-score (uint16 slots per warp)
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// One warp per query row.
//   pass 1: stage the row's S*kp approximate scores in shared memory and take every full split list's minimum
//           (t_split; lists are unsorted, kp in {16, 32} so a list is a lane segment of the warp);
//   pass 2: tau = k-th largest approximate score over all candidates (max-below-previous iteration);
//   pass 3: only candidates with s_bf16 >= tau - 2 EPS can be among the exact top k (the k candidates that define
//           tau have exact scores >= tau - EPS, everything below the cut has exact score < tau - EPS), so only
//           those are re-scored in fp32 from the master keys -- typically 15-30 of the 144+ candidates;
//   certificate: k-th exact score > max_split t_split + EPS, else the row goes to the second-pass list together with
//           thr2 = k-th exact score - EPS - 1e-6: every key that can still enter the exact top k scores above it.
__global__ void __launch_bounds__(256) refine_kernel(const RefineArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int total = a.n_splits * a.kp;
  int64_t* li = reinterpret_cast<int64_t*>(smem_raw) + (size_t)warp * a.k;
  float* lv = reinterpret_cast<float*>(reinterpret_cast<int64_t*>(smem_raw) + (size_t)wpb * a.k) + (size_t)warp * a.k;
  float* ap = reinterpret_cast<float*>(reinterpret_cast<int64_t*>(smem_raw) + (size_t)wpb * a.k) + (size_t)wpb * a.k +
              (size_t)warp * total;
  uint16_t* sel = reinterpret_cast<uint16_t*>(reinterpret_cast<float*>(reinterpret_cast<int64_t*>(smem_raw) + (size_t)wpb * a.k) +
                                              (size_t)wpb * a.k + (size_t)wpb * total) + (size_t)warp * REFINE_SEL_CAP;
  const int nd = (a.d + 31) >> 5;                              // <= 8 (tensor-core shapes have d <= 256)
  const float kerr = a.kerr_max ? __ldg(a.kerr_max) : 0.f;
  for (int64_t row = (int64_t)blockIdx.x * wpb + warp; row < a.Q; row += (int64_t)gridDim.x * wpb) {
    const float eps = (a.qerr ? __ldg(a.qerr + row) : 0.f) + kerr + a.eps_fixed;
    for (int p = lane; p < a.k; p += 32) { lv[p] = -FLT_MAX; li[p] = INT64_MAX; }
    // ---- pass 1 ------------------------------------------------------------------------------------
    float tmax = -INFINITY;
    for (int c0 = 0; c0 < total; c0 += 32) {
      const int c = c0 + lane;
      const bool in = c < total;
      const int sp = c / a.kp, p = c - sp * a.kp;
      const size_t o = ((size_t)sp * a.Q + row) * a.kp + p;
      const int32_t j = in ? __ldg(a.part_i + o) : -1;
      const float s = in ? __ldg(a.part_s + o) : 0.f;
      const bool valid = j >= 0;
      // history mask: an excluded candidate keeps its place in the list statistics below (it bounds what the split did
      // not list) but never becomes a result.  All 32 lanes scan the row's exclusion list together.
      bool excluded = false;
      if (a.mask_rowptr) {
        const int64_t mlo = __ldg(a.mask_rowptr + row), mhi = __ldg(a.mask_rowptr + row + 1);
        const int64_t gj = (int64_t)j + a.idx_offset;
        for (int64_t m0 = mlo; m0 < mhi; m0 += 32) {
          const int64_t mv = (m0 + lane < mhi) ? __ldg(a.mask_col + m0 + lane) : (int64_t)-1;
#pragma unroll 8
          for (int t = 0; t < 32; ++t) excluded |= (__shfl_sync(0xffffffffu, mv, t) == gj);
        }
      }
      if (in) ap[c] = (valid && !excluded) ? s : -INFINITY;
      float mn = valid ? s : INFINITY;
      int full = valid ? 1 : 0;
      for (int w = 1; w < a.kp; w <<= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, w));
        full &= __shfl_xor_sync(0xffffffffu, full, w);
      }
      if (in && full) tmax = fmaxf(tmax, mn);
    }
    tmax = warp_max(tmax);
    if (a.thr0) tmax = fmaxf(tmax, __ldg(a.thr0 + row));
    // (diagnostic for the format policy: what would bound the row under a LARGER error bound -- the shared threshold below
    // sits 2 eps under the k-th best by construction and would move with eps, the list minima and the pre-pass bound would not)
    const float tmax_lists = tmax;
    if (a.part_thr) {
      float tp = -INFINITY;
      for (int sp = lane; sp < a.n_splits; sp += 32) tp = fmaxf(tp, __ldg(a.part_thr + (size_t)sp * a.Q + row));
      tmax = fmaxf(tmax, warp_max(tp));
    }
    __syncwarp();
    // ---- pass 2 ------------------------------------------------------------------------------------
    float prev = INFINITY;
    int cnt = 0;
    while (cnt < a.k) {
      float m = -INFINITY;
      for (int c = lane; c < total; c += 32) { const float v = ap[c]; if (v < prev) m = fmaxf(m, v); }
      m = warp_max(m);
      if (m == -INFINITY) break;                               // fewer than k valid candidates
      int n = 0;
      for (int c = lane; c < total; c += 32) n += (ap[c] == m) ? 1 : 0;
      cnt += __reduce_add_sync(0xffffffffu, n);
      prev = m;
    }
    const float cut = (cnt >= a.k) ? (a.exact ? prev - 2.0f * eps : prev) : -INFINITY;
    // ---- pass 3 ------------------------------------------------------------------------------------
    float qreg[8];
    const float qinv = a.q_inv_norm ? __ldg(a.q_inv_norm + row) : 1.0f;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int e = lane + 32 * t;
      qreg[t] = (a.exact && t < nd && e < a.d) ? __ldg(a.q + row * a.d + e) * qinv : 0.f;
    }
    // The selected candidates (typically 5-30 of the S * kp) are compacted into a short list first and re-scored four at a
    // time from there: every batch is a chain of three dependent memory round trips (index -> key row -> inverse norm), and
    // walking the candidate chunks one ballot at a time made that chain once per chunk holding a candidate (measured at the
    // reference's Cora shape: 13 splits -> up to 7 chains, ~20 of the op's 190 us) instead of once per four candidates.
    for (int c0 = 0; c0 < total;) {
      int n_sel = 0;
      for (; c0 < total && n_sel <= REFINE_SEL_CAP - 32; c0 += 32) {
        const int c = c0 + lane;
        const float mine = (c < total) ? ap[c] : -INFINITY;
        const bool pick = mine >= cut && mine > -INFINITY;
        const unsigned m = __ballot_sync(0xffffffffu, pick);
        if (pick) sel[n_sel + __popc(m & ((1u << lane) - 1u))] = (uint16_t)c;
        n_sel += __popc(m);
      }
      __syncwarp();
      for (int b0 = 0; b0 < n_sel; b0 += 4) {
        float sc[4]; int64_t id[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          sc[u] = -FLT_MAX; id[u] = -1;
          if (b0 + u < n_sel) {
            const int cc = sel[b0 + u];
            const int sp = cc / a.kp, p = cc - sp * a.kp;
            const int32_t j = __ldg(a.part_i + ((size_t)sp * a.Q + row) * a.kp + p);
            id[u] = j;
            if (a.exact) {
              const float* kr = a.keys + (int64_t)j * a.d;
              float dot = 0.f;
#pragma unroll
              for (int t = 0; t < 8; ++t) {
                const int e = lane + 32 * t;
                if (t < nd && e < a.d) dot = fmaf(qreg[t], __ldg(kr + e), dot);
              }
              sc[u] = dot;
            } else {
              sc[u] = ap[cc];
            }
          }
        }
        if (a.exact) {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (id[u] >= 0)
              sc[u] = warp_sum(sc[u]) * (a.key_scale > 0.f ? a.key_scale : (a.key_inv_norm ? __ldg(a.key_inv_norm + id[u]) : 1.0f));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (id[u] >= 0 && ranks_before(sc[u], id[u], lv[a.k - 1], li[a.k - 1]))
            warp_sorted_insert<int64_t>(lv, li, a.k, sc[u], id[u], lane);
      }
      __syncwarp();                                             // sel is refilled by the next round
    }
    const float kth = lv[a.k - 1];
    // dot-product ranking: back from q_hat . (key_scale * k) to q . k (a zero query row scores 0 everywhere)
    const float out_scale = (a.exact && a.key_scale > 0.f) ? 1.0f / (qinv * a.key_scale) : 1.0f;
    const bool have_k = li[a.k - 1] != INT64_MAX;
    const bool certified = !a.exact || (have_k && kth > tmax + eps);
    for (int p = lane; p < a.k; p += 32) {
      a.out_scores[row * a.k + p] = (li[p] == INT64_MAX) ? lv[p] : lv[p] * out_scale;
      a.out_idx[row * a.k + p] = (li[p] == INT64_MAX) ? (int64_t)-1 : li[p] + a.idx_offset;
    }
    if (a.exact && a.loose_count && lane == 0 && !(have_k && kth > tmax_lists + a.loose_mult * eps)) atomicAdd(a.loose_count, 1);
    if (!certified && lane == 0) {
      const int slot = atomicAdd(a.fb_count, 1);
      a.fb_rows[slot] = (int32_t)row;
      a.thr2[slot] = have_k ? kth - eps - 1e-6f : -INFINITY;
    }
    __syncwarp();
  }
}


// ---- refine of the second pass: exact fp32 re-score of EVERYTHING the collect pass returned ----------------
// One warp per uncertified row (compact slot).  The collect pass delivered every key whose 16-bit score exceeds
// thr2 = (k-th exact score of the first refine) - eps - 1e-6, so a key it did not deliver has an exact score below that
// k-th score: the best k of the delivered set are the exact top k -- unless the row's spill area overflowed, in which
// case the row goes to the fp32 kernel's list (dense ties beyond TC_SPILL_MAX: adversarial libraries only).
struct Refine2Args {
  const float* q; const float* keys; const float* q_inv_norm; const float* key_inv_norm;
  int d; int k; int64_t idx_offset;
  const int32_t* rows; const int32_t* n_rows_dev;            // both NULL (two-pass mode): every row 0 .. n_rows - 1, slot = row
  int n_rows;
  const float* spill_s; const int32_t* spill_i; const int32_t* spill_cnt; int spill_cap;
  float* out_scores; int64_t* out_idx;
  int32_t* fb_rows; int32_t* fb_count;
  const int64_t* mask_rowptr; const int64_t* mask_col; float key_scale;      // as in RefineArgs
  // two-pass mode: a row whose spill area overflowed is not sent to the fp32 kernel but RETRIED -- the exact k-th score
  // of the candidates that did fit is a lower bound of the row's true k-th best, so (row, that - eps) goes to the list
  // the standard second pass consumes (NULL: overflow -> fb_rows as usual)
  int32_t* retry_rows; float* retry_thr; int32_t* retry_count;
  const float* qerr; const float* kerr_max; float eps_fixed;
};

__global__ void __launch_bounds__(256) refine2_kernel(const Refine2Args a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  int64_t* li = reinterpret_cast<int64_t*>(smem_raw) + (size_t)warp * a.k;
  float* lv = reinterpret_cast<float*>(reinterpret_cast<int64_t*>(smem_raw) + (size_t)wpb * a.k) + (size_t)warp * a.k;
  const int n = a.n_rows_dev ? *a.n_rows_dev : a.n_rows;
  const int nd = (a.d + 31) >> 5;
  for (int slot = blockIdx.x * wpb + warp; slot < n; slot += gridDim.x * wpb) {
    const int64_t row = a.rows ? (int64_t)a.rows[slot] : (int64_t)slot;
    const int cnt_all = a.spill_cnt[slot];
    const bool over = cnt_all > a.spill_cap;
    if (over && !a.retry_rows) {
      if (lane == 0) a.fb_rows[atomicAdd(a.fb_count, 1)] = (int32_t)row;
      continue;
    }
    const int cnt = over ? a.spill_cap : cnt_all;
    for (int p = lane; p < a.k; p += 32) { lv[p] = -FLT_MAX; li[p] = INT64_MAX; }
    __syncwarp();
    float qreg[8];
    const float qinv = a.q_inv_norm ? __ldg(a.q_inv_norm + row) : 1.0f;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int e = lane + 32 * t;
      qreg[t] = (t < nd && e < a.d) ? __ldg(a.q + row * a.d + e) * qinv : 0.f;
    }
    const int32_t* ci = a.spill_i + (size_t)slot * a.spill_cap;
    for (int c0 = 0; c0 < cnt; c0 += 4) {
      float sc[4]; int64_t id[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        id[u] = (c0 + u < cnt) ? (int64_t)ci[c0 + u] : -1;     // plain loads: written by the collect pass on this stream
        if (a.mask_rowptr && id[u] >= 0) {                     // history mask (RefineArgs::mask_rowptr)
          const int64_t mlo = __ldg(a.mask_rowptr + row), mhi = __ldg(a.mask_rowptr + row + 1), gj = id[u] + a.idx_offset;
          bool ex = false;
          for (int64_t m0 = mlo; m0 < mhi && !ex; m0 += 32)
            ex = __any_sync(0xffffffffu, m0 + lane < mhi && __ldg(a.mask_col + m0 + lane) == gj);
          if (ex) id[u] = -1;
        }
        float dot = 0.f;
        if (id[u] >= 0) {
          const float* kr = a.keys + id[u] * a.d;
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const int e = lane + 32 * t;
            if (t < nd && e < a.d) dot = fmaf(qreg[t], __ldg(kr + e), dot);
          }
        }
        sc[u] = dot;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (id[u] >= 0)
          sc[u] = warp_sum(sc[u]) * (a.key_scale > 0.f ? a.key_scale : (a.key_inv_norm ? __ldg(a.key_inv_norm + id[u]) : 1.0f));
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (id[u] >= 0 && ranks_before(sc[u], id[u], lv[a.k - 1], li[a.k - 1]))
          warp_sorted_insert<int64_t>(lv, li, a.k, sc[u], id[u], lane);
    }
    if (over) {                                              // (retry_rows != NULL: see Refine2Args)
      if (lane == 0) {
        const int s2 = atomicAdd(a.retry_count, 1);
        const float eps = (a.qerr ? __ldg(a.qerr + row) : 0.f) + (a.kerr_max ? __ldg(a.kerr_max) : 0.f) + a.eps_fixed;
        a.retry_rows[s2] = (int32_t)row;
        a.retry_thr[s2] = (li[a.k - 1] != INT64_MAX) ? lv[a.k - 1] - eps - 1e-6f : -INFINITY;
      }
      __syncwarp();
      continue;
    }
    const float out_scale = a.key_scale > 0.f ? 1.0f / (qinv * a.key_scale) : 1.0f;
    for (int p = lane; p < a.k; p += 32) {
      a.out_scores[row * a.k + p] = (li[p] == INT64_MAX) ? lv[p] : lv[p] * out_scale;
      a.out_idx[row * a.k + p] = (li[p] == INT64_MAX) ? (int64_t)-1 : li[p] + a.idx_offset;
    }
    __syncwarp();
  }
}

// ---- host side --------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<PFN_encodeTiled>(p);
  }();
  return fn;
}

// K-major [rows, d_pad] operand image, boxes of box_rows x 128 bytes (64 bf16 or 32 fp32/tf32 elements), SWIZZLE_128B
static int make_map_bf16(CUtensorMap* map, const void* ptr, int64_t rows, int d_pad, int box_rows, bool f32 = false) {
  PFN_encodeTiled enc = get_encode();
  RAG_REQUIRE(enc, RAG_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)d_pad, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)d_pad * (f32 ? 4 : 2)};
  cuuint32_t box[2] = {f32 ? 32u : 64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RAG_REQUIRE(r == CUDA_SUCCESS, RAG_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return RAG_OK;
}

struct TcPlan {
  int kh, kp, nstage, n_qtiles, n_splits, tiles_per_split, n_tiles, d_pad;
  size_t smem;
  size_t off_qbf, off_qinv, off_qerr, off_ps, off_pi, off_fb, off_zero, off_thr2, off_fb2, off_gmax, off_thr0, off_ovf,
      off_spill_s, off_spill_i, off_f32, total;
  int64_t n_zero_share;               // ... + gthr + candidate pool, when the call runs the cross-split sweep
  int64_t n_zero;                     // 32-bit words cleared by the prologue: {pass-2 row count, fp32 row count, pad} + spill counters
  int spill_cap;
  int pre_tiles, pre_groups;          // threshold pre-pass (0 tiles = off)
  // cross-split threshold sharing (TcArgs::pool): sweeping CTAs on the SMs the worker grid leaves idle (0 = off)
  int n_mergers; size_t off_gthr, off_pool, off_pthr;
};
constexpr int TC_ZERO_HDR = 64;       // words in front of the per-row spill counters
constexpr int TWOPASS_RESERVE_TILES = 8192;
constexpr int64_t TWOPASS_MAX_Q = 262144;                 // (its group maxima take up to 1 KB of workspace per query row)
// measured crossover against the one-pass SS kernel (profiles/r2_twopass_ab.jsonl): d <= 128 still 1.2x at 868 tiles per CTA
// (4 096 x 1 M keys), d = 256 (twice the MMA time per tile, so the second pass costs more) even at ~220
// k > 16 (32-entry lists, costlier warm-up): 1.3x at 1 202 tiles (4 096 x 2 M x 128, k = 100), 0.73x at 3 473
constexpr int TWOPASS_TILES_D128 = 1024, TWOPASS_TILES_D256 = 192, TWOPASS_TILES_WIDEK = 1536;   // largest "twopass_max_tiles" the workspace is laid out for

// Process-wide tuning / test hooks (rag_tc_set_option); the environment is read ONCE, when the first call needs them.
struct TcOptions {
  int variant = 0;                    // 0 auto, 1 = SS kernel wherever it is instantiated, 2 = TS kernel
  int prepass = 1;
  int prepass_min_tiles = 1024;
  int prepass_div = 64;
  int prepass_max_tiles = 192;        // pre-pass length cap per CTA (key tiles)
  int kp = 0;                         // 0 auto, 16 / 32 = candidate-list length per (row, split) where the shape allows
  int pass2 = 1;                      // 0: uncertified rows go straight to the fp32 kernel (the round-1 behaviour)
  int gshare = 1;                     // cross-split threshold sharing on the idle SMs (TcArgs::pool); 0 = off
  int gshare_ctas = 4;                // at most this many sweeping CTAs
  int gshare_dbg = 0;                 // experiments (TcArgs::g_dbg)
  int twopass = 1;                    // short streams: maxima pass + collect pass instead of list warm-up (topk_tc_run)
  int twopass_max_tiles = 0;          // ... up to this many key tiles per CTA (0 = by d: TWOPASS_TILES_D128 / _D256)
};
static TcOptions tc_env_defaults() {
  TcOptions o;
  if (const char* e = getenv("RAG_TC_VARIANT")) o.variant = (e[0] == 's' && e[1] == 's') ? 1 : ((e[0] == 't' && e[1] == 's') ? 2 : 0);
  if (const char* e = getenv("RAG_TC_PREPASS")) o.prepass = (e[0] == '0') ? 0 : 1;
  if (const char* e = getenv("RAG_TC_PREPASS_MIN_TILES")) { if (atoi(e) >= 64) o.prepass_min_tiles = atoi(e); }
  if (const char* e = getenv("RAG_TC_PREPASS_DIV")) { if (atoi(e) >= 4) o.prepass_div = atoi(e); }
  if (const char* e = getenv("RAG_TC_KP")) { if (atoi(e) == 16 || atoi(e) == 32) o.kp = atoi(e); }
  if (const char* e = getenv("RAG_TC_PASS2")) o.pass2 = (e[0] == '0') ? 0 : 1;
  if (const char* e = getenv("RAG_TC_GSHARE")) o.gshare = (e[0] == '0') ? 0 : 1;
  if (const char* e = getenv("RAG_TC_TWOPASS")) o.twopass = (e[0] == '0') ? 0 : 1;
  return o;
}
static TcOptions& tc_opts() {
  static TcOptions o = tc_env_defaults();
  return o;
}

constexpr int TC_MAX_K_WIDE_FWD = 128;
// SS kernel shapes (query tile resident in shared memory)
static bool tc_shape_ok_ss(int d, int k) {
  if (d < 1 || k < 1) return false;
  if (d <= 128) return k <= TC_MAX_K_WIDE_FWD;
  if (d > 192 && d <= 256) return k <= 10;        // 128 KB of resident queries leave room for 16-entry lists only
  return false;
}
// TS kernel shapes (query tile resident in tensor memory): d <= 256; k' = 16 slots per (row, split) for k <= 10,
// 32 for k <= 26 (d <= 128 only: 128 KB of lists + append buffers leave 4 pipeline stages, d > 128 needs 8)
// k in (26, 128] (d <= 128; the edge variant's vanilla phase retrieves 50, RAGraph_edge/modules/RAGraph.py:36-38): no
// longer lists -- MORE KEY SPLITS.  The exact top k lies in the union of the per-split top-32 lists unless one split holds
// more than 32 of them, and the refine certificate (k-th exact score > every full list's minimum + eps) detects exactly
// that case and sends the row through the second pass; the planner makes the lists of a row hold >= 4k entries together
// (n_splits >= 4k / 32, CTAs run in waves when that exceeds the SM count), so on spread-out keys a split sees k / n_splits
// of the winners on average and the certificate margin is the gap between the k-th and the (32 n_splits)-th best score.
constexpr int TC_MAX_K_WIDE = TC_MAX_K_WIDE_FWD;
bool tc_shape_ok(int d, int k) { return d >= 1 && d <= 256 && k >= 1 && (k <= 10 || (k <= TC_MAX_K_WIDE && d <= 128)); }
// tf32 (SS kernel only): an fp32 operand row is twice as wide as a bf16 one, so d <= 64 has the bf16 d <= 128 budget
// (k <= 26) and d <= 128 the bf16 d = 256 budget (128 KB of resident queries, 16-entry lists, k <= 10)
bool tc_shape_ok_tf32(int d, int k) { return d >= 1 && k >= 1 && ((d <= 64 && k <= 26) || (d <= 128 && k <= 10)); }

// Which filter kernel runs.  Round 1 measured the query-stationary TS kernel ahead only from ~8 192 key tiles per CTA (its
// selection is built for the sparse steady state; short streams are one long warm-up, where the SS kernel's lane-parallel
// list updates are 2-3x faster).  With the threshold pre-pass and the cross-split sharing of round 2 the TS kernel wins from
// the first shape that has a pre-pass (profiles/r2_variant_ab_mid_streams.jsonl, Q = 4 096: 1.5 M x 128 keys 1.87 vs 2.30 ms,
// 2 M x 64 1.73 vs 2.71, 2 M x 256 3.12 vs 3.12, 4 M x 256 6.32 vs 6.43, 8 M x 128 6.21 vs 6.59), and below that the two-pass
// mode (topk_tc_run) replaces list warm-up altogether.  Default: TS from 1 024 tiles per CTA and for the shapes SS does not
// instantiate; SS in between (d > 128: 192-1 024 tiles).  rag_tc_set_option("variant", 1|2) forces one (A/B, tests).
constexpr int TS_MIN_TILES_PER_SPLIT = 1024;
static bool tc_use_ts(int d, int k, int tiles_per_split) {
  const bool ss_ok = tc_shape_ok_ss(d, k);
  const int v = tc_opts().variant;
  if (v == 1) return !ss_ok;
  if (v == 2) return true;
  return !ss_ok || tiles_per_split >= TS_MIN_TILES_PER_SPLIT;
}

// kp_req: candidate-list length asked for by the call (RAG_SIM_WIDE_LISTS -> 32; 0 = by k); the process-wide option wins
static TcPlan tc_plan(int64_t Q, int64_t N, int d, int k, bool ts, bool tf32 = false, int kp_req = 0) {
  const TcOptions& o = tc_opts();
  TcPlan p{};
  if (tf32) {
    p.d_pad = d <= 32 ? 32 : (d <= 64 ? 64 : 128);   // fp32 elements; the shadow is padded to the SS instantiations 1, 2, 4
    p.kh = p.d_pad / 32;
  } else {
    p.d_pad = (d + 63) / 64 * 64;
    p.kh = p.d_pad / 64;                     // 1..4
    if (!ts && p.kh == 3) p.kh = 4;          // SS instantiations: 1, 2, 4
  }
  p.kp = (k <= 10) ? 16 : 32;
  // wider lists certify more rows of a clustered library in the first pass (d <= 128: the shared-memory budget)
  if (!tf32 && d <= 128 && (o.kp == 32 || (o.kp == 0 && kp_req == 32))) p.kp = 32;
  const int kp_layout = (!tf32 && d <= 128) ? 32 : p.kp;      // workspace areas are sized for the widest lists of the shape
  // shared memory: (SS: A 2*KH boxes +) NSTAGE boxes + lists + barriers + 1 KB alignment slack  <= 227 KB
  // SS: lists + meta + pending queues; TS: sorted lists + append buffers (stride 257) + 16 B alignment slack
  const int list_bytes = ts ? (p.kp + ts_kpb(p.kp)) * TS_LSTRIDE * 8 + TS_QUEUE_BYTES + 16 : p.kp * TC_ROWS * 8 + TC_ROWS * 4 + TC_PQ * TC_ROWS * 8;
  const int a_boxes = ts ? 0 : 2 * p.kh;
  const int bar_bytes = ts ? TS_BAR_BYTES : 256;
  const int slack = ts ? 0 : 1024;           // SS aligns its window by hand; TS declares the 1 KB alignment
  const int budget = 232448 - slack - bar_bytes - list_bytes - a_boxes * TC_BOX_BYTES;
  p.nstage = budget / TC_BOX_BYTES;
  const int cap = ts ? (p.kp == 16 ? 9 : 4) : 8;
  if (p.nstage > cap) p.nstage = cap;
  if (p.nstage < 2) p.nstage = 2;
  p.smem = slack + (size_t)(a_boxes + p.nstage) * TC_BOX_BYTES + list_bytes + bar_bytes;
  p.n_qtiles = (int)((Q + TC_ROWS - 1) / TC_ROWS);
  p.n_tiles = (int)((N + TC_BN - 1) / TC_BN);
  int s = sm_count() / p.n_qtiles;
  if (s < 1) s = 1;
  if (k > 26 && s * p.kp < 4 * k) s = (4 * k + p.kp - 1) / p.kp;     // wide k: the lists of a row hold >= 4k entries together
  if (s > p.n_tiles) s = p.n_tiles;
  p.tiles_per_split = (p.n_tiles + s - 1) / s;
  p.n_splits = (p.n_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  int64_t sc = ((int64_t)1 << 26) / (Q > 0 ? Q : 1);
  p.spill_cap = (int)(sc > TC_SPILL_MAX ? TC_SPILL_MAX : (sc < 64 ? 64 : sc));
  // cross-split sharing: needs idle SMs next to the worker grid, a long stream, and a row's published words in the sweep's
  // registers (<= 512); the zero block is then [header | spill counters Q | gthr Qpad | pool Qpad * S * kp].  The workspace
  // reserve depends on the shape only (not on options), so offsets are the same for every call of a shape.
  const int64_t Qpad = (int64_t)p.n_qtiles * TC_ROWS;
  const bool share_shape = ts && !tf32 && p.n_splits > 1 && p.n_splits * 16 <= 512 && sm_count() > p.n_qtiles * p.n_splits;
  p.n_mergers = 0;
  if (share_shape && o.gshare && p.n_splits * p.kp <= 512 && p.tiles_per_split >= o.prepass_min_tiles)
    p.n_mergers = std::min(o.gshare_ctas, sm_count() - p.n_qtiles * p.n_splits);
  p.n_zero = TC_ZERO_HDR + Q;
  p.n_zero_share = TC_ZERO_HDR + Q + Qpad + Qpad * p.n_splits * p.kp;
  size_t off = 0;
  p.off_qbf = off; off += align_up((size_t)p.n_qtiles * TC_ROWS * p.d_pad * (tf32 ? 4 : 2), 256);
  p.off_qinv = off; off += align_up((size_t)Q * 4, 256);
  p.off_qerr = off; off += align_up((size_t)Q * 4, 256);
  p.off_ps = off; off += align_up((size_t)p.n_splits * Q * kp_layout * 4, 256);
  p.off_pi = off; off += align_up((size_t)p.n_splits * Q * kp_layout * 4, 256);
  p.off_fb = off; off += align_up((size_t)Q * 4, 256);
  p.off_zero = off; off += align_up((size_t)(TC_ZERO_HDR + Q + (share_shape ? Qpad + Qpad * p.n_splits * kp_layout : 0)) * 4, 256);
  p.off_gthr = p.off_zero + (size_t)(TC_ZERO_HDR + Q) * 4;
  p.off_pool = p.off_gthr + (size_t)Qpad * 4;
  p.off_pthr = off; off += share_shape ? align_up((size_t)p.n_splits * Q * 4, 256) : 0;
  p.off_thr2 = off; off += align_up((size_t)Q * 4, 256);
  p.off_fb2 = off; off += align_up((size_t)Q * 4, 256);
  // threshold pre-pass (TS kernel): 1/64 of every split, worthwhile once a split has >= 1024 tiles; G = n_splits * groups
  // group maxima per row, G >= 2 * kp so that the kp-th largest of them sits near the middle of their distribution
  p.pre_tiles = 0; p.pre_groups = 0;
  if (ts && p.tiles_per_split >= o.prepass_min_tiles) {
    int g = (2 * p.kp + p.n_splits - 1) / p.n_splits;
    if (g < 1) g = 1;
    // 1/64 of the stream, but no more than prepass_max_tiles: with the cross-split sharing tightening the bound within the
    // first few hundred microseconds, what a longer sample buys no longer pays for its tiles (measured, 100 M x 128 keys:
    // 1 356 / 339 / 169 pre-pass tiles per CTA -> 72.99 / 72.56 / 72.33 ms; 12.5 M: 169 tiles beat 85 and 42)
    const int pre = std::min(p.tiles_per_split / o.prepass_div, o.prepass_max_tiles);
    if ((int64_t)g * p.n_splits <= 256 && g <= pre) { p.pre_tiles = pre; p.pre_groups = g; }
  }
  // (two-pass mode groups the WHOLE stream into up to 256 groups per row; only short streams qualify: TWOPASS_RESERVE_TILES)
  size_t gmax_groups = (size_t)((2 * kp_layout + p.n_splits - 1) / p.n_splits) * p.n_splits;
  if (ts && !tf32 && p.tiles_per_split <= TWOPASS_RESERVE_TILES && Q <= TWOPASS_MAX_Q)
    gmax_groups = std::max(gmax_groups, (size_t)std::min(p.tiles_per_split, 256 / std::max(p.n_splits, 1)) * p.n_splits);
  p.off_gmax = off; off += align_up(gmax_groups * Q * 4, 256);
  p.off_thr0 = off; off += align_up((size_t)Q * 4, 256);
  // hit-queue overflow area: one slice per CTA of the largest TS launch (the second pass runs max(#SMs, query tiles) CTAs)
  const int64_t ts_ctas = std::max<int64_t>((int64_t)p.n_qtiles * p.n_splits, std::max<int64_t>(sm_count(), p.n_qtiles));
  p.off_ovf = off; off += ts ? align_up((size_t)ts_ctas * TC_EPI_WARPS * 32 * TC_BN * 4, 256) : 0;
  p.off_spill_s = off; off += (ts && !tf32) ? align_up((size_t)Q * p.spill_cap * 4, 256) : 0;
  p.off_spill_i = off; off += (ts && !tf32) ? align_up((size_t)Q * p.spill_cap * 4, 256) : 0;
  p.off_f32 = off; off += topk_f32_rows_workspace(Q, N, d, k);
  p.total = off;
  return p;
}

bool topk_tc_available(int d, int k) { return tc_shape_ok(d, k); }
bool topk_tc_tf32_available(int d, int k) { return tc_shape_ok_tf32(d, k); }
int topk_tc_tf32_dpad(int d) { return d <= 32 ? 32 : (d <= 64 ? 64 : 128); }

size_t topk_tc_workspace(int64_t Q, int64_t N, int d, int k, int mode) {
  if (mode == RAG_SIM_TF32) return tc_shape_ok_tf32(d, k) ? tc_plan(Q, N, d, k, false, true).total : 256;
  if (!tc_shape_ok(d, k)) return 256;
  return tc_plan(Q, N, d, k, true).total;   // offsets do not depend on the variant (the TS layout is the superset)
}

void topk_tc_stat_offsets(int64_t Q, int64_t N, int d, int k, int mode, size_t* out) {
  out[0] = out[1] = out[2] = 0;
  if ((mode != RAG_SIM_BF16_REFINE && mode != RAG_SIM_F16_REFINE) || !tc_shape_ok(d, k) || Q <= 0 || N <= 0) return;
  const TcPlan p = tc_plan(Q, N, d, k, true);
  out[0] = p.off_zero;
  out[1] = p.off_zero + 4;
  out[2] = p.off_zero + 8;
}

template <int KH, int NSTAGE, int KP, bool TF32 = false>
static int launch_filter(const CUtensorMap& mq, const CUtensorMap& mk, const TcArgs& a, const TcPlan& p, cudaStream_t s) {
  auto kern = cosine_topk_tc_kernel<KH, NSTAGE, KP, TF32>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(cosine_topk_tc_kernel)");
  kern<<<(unsigned)(p.n_qtiles * p.n_splits), TC_THREADS, p.smem, s>>>(mq, mk, a);
  RAG_LAUNCH_OK("cosine_topk_tc_kernel");
  return RAG_OK;
}

template <int KH, int NSTAGE, int KP, bool PRE>
static int launch_filter_ts(const CUtensorMap& mk, const uint16_t* q_bf, const TcArgs& a, const TcPlan& p, unsigned grid,
                            cudaStream_t s) {
  static_assert(sizeof(TsBarriers) <= TS_BAR_BYTES, "barrier block");
  auto kern = cosine_topk_ts_kernel<KH, NSTAGE, KP, PRE>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(cosine_topk_ts_kernel)");
  kern<<<grid, TS_THREADS, p.smem, s>>>(mk, q_bf, a);
  RAG_LAUNCH_OK("cosine_topk_ts_kernel");
  return RAG_OK;
}

// one TS launch of plan p: the pre-pass instantiation when ta.premax, else the main / collect kernel
static int run_ts(const CUtensorMap& mk, const uint16_t* q_bf, const TcArgs& ta, const TcPlan& p, unsigned grid, cudaStream_t s) {
#define RAG_TS_CASE(KH_, NS_, KP_) \
  if (p.kh == KH_ && p.nstage == NS_ && p.kp == KP_) \
    return ta.premax ? launch_filter_ts<KH_, NS_, KP_, true>(mk, q_bf, ta, p, grid, s) : launch_filter_ts<KH_, NS_, KP_, false>(mk, q_bf, ta, p, grid, s);
  RAG_TS_CASE(1, 9, 16) RAG_TS_CASE(2, 9, 16) RAG_TS_CASE(3, 9, 16) RAG_TS_CASE(4, 9, 16)
  RAG_TS_CASE(1, 4, 32) RAG_TS_CASE(2, 4, 32)
#undef RAG_TS_CASE
  return fail(RAG_EUNSUPPORTED, "cosine_topk: no tensor-core instantiation for kh=%d nstage=%d kp=%d", p.kh, p.nstage, p.kp);
}

int rows_to_16_launch(const float* x, int64_t rows, int d, int fmt, int normalize, float eps, uint16_t* out, int d_pad,
                      float* inv_out, float* err_rows, float* err_max, uint32_t* zero_words, int64_t n_zero, cudaStream_t s);

// Which kernel serves a call -- one place, shared by topk_tc_run and rag_cosine_topk_plan (introspection, CPU tests).
struct TcDispatch {
  TcPlan pts, p;
  bool two_pass, ts;
};
static TcDispatch tc_dispatch(int64_t Q, int64_t N, int d, int k, int mode, uint32_t flags, bool has_mask) {
  const bool tf32 = (mode == RAG_SIM_TF32);
  const bool exact = (mode == RAG_SIM_BF16_REFINE || mode == RAG_SIM_F16_REFINE);
  const bool dot = (flags & RAG_SIM_DOT) != 0;
  const int kp_req = (flags & RAG_SIM_WIDE_LISTS) ? 32 : 0;
  TcDispatch D{};
  D.pts = tc_plan(Q, N, d, k, true, false, kp_req);
  const TcPlan& pts = D.pts;
  // two-pass mode for short streams (topk_tc_run): exact modes, no exclusion lists (a group maximum may be an excluded key), at
  // least 2k of the <= 256 group maxima per row (k <= 128), automatic kernel choice (a forced variant runs its own selection)
  const TcOptions& opt = tc_opts();
  const int tp_tiles = opt.twopass_max_tiles > 0 ? opt.twopass_max_tiles
                                                 : (d > 128 ? TWOPASS_TILES_D256 : (k > 16 ? TWOPASS_TILES_WIDEK : TWOPASS_TILES_D128));
  D.two_pass = exact && !tf32 && !has_mask && !dot && k <= TC_MAX_K_WIDE && opt.twopass && opt.variant == 0 &&
               pts.tiles_per_split <= std::min(tp_tiles, TWOPASS_RESERVE_TILES) && Q <= TWOPASS_MAX_Q &&
               std::min(pts.tiles_per_split, 256 / std::max(pts.n_splits, 1)) * pts.n_splits >= 2 * k && N >= 4 * TC_BN;
  D.ts = !tf32 && (D.two_pass || tc_use_ts(d, k, pts.tiles_per_split));
  D.p = tf32 ? tc_plan(Q, N, d, k, false, true) : (D.ts ? pts : tc_plan(Q, N, d, k, false, false, kp_req));
  return D;
}

// mask_rowptr / mask_col (nullable): per-row exclusion lists of GLOBAL key indices; key_scale > 0 (with RAG_SIM_DOT): dot-product
// ranking over a shadow of keys * key_scale (RefineArgs::key_scale).  Both need an exact (refine) mode.
int topk_tc_run(const float* q, int64_t Q, const float* keys, const float* key_inv_norm, const void* keys_shadow,
                const float* shadow_err, int64_t N, int d, int k, int mode, uint32_t flags, int64_t idx_offset,
                float* out_scores, int64_t* out_idx, void* ws, size_t ws_bytes, cudaStream_t s,
                const int64_t* mask_rowptr, const int64_t* mask_col, float key_scale) {
  const bool tf32 = (mode == RAG_SIM_TF32);
  const bool f16 = (mode == RAG_SIM_F16 || mode == RAG_SIM_F16_REFINE);
  const bool exact = (mode == RAG_SIM_BF16_REFINE || mode == RAG_SIM_F16_REFINE);
  const bool dot = (flags & RAG_SIM_DOT) != 0;
  RAG_REQUIRE(!dot || (exact && key_scale > 0.f), RAG_EUNSUPPORTED,
              "cosine_topk: RAG_SIM_DOT on the tensor cores needs an exact mode and key_scale > 0 (rag_topk_masked_tc_f32)");
  RAG_REQUIRE(dot || key_scale == 0.f, RAG_EINVAL, "cosine_topk: key_scale without RAG_SIM_DOT");
  RAG_REQUIRE(!mask_rowptr || (exact && mask_col), RAG_EUNSUPPORTED, "cosine_topk: exclusion lists need an exact mode");
  RAG_REQUIRE(!(flags & ~(RAG_SIM_DOT | RAG_SIM_WIDE_LISTS)), RAG_EINVAL, "cosine_topk: unknown flag bits 0x%x", flags);
  if (tf32)
    RAG_REQUIRE(tc_shape_ok_tf32(d, k), RAG_EUNSUPPORTED,
                "cosine_topk: the tf32 mode covers d <= 64 with k <= 26 and d <= 128 with k <= 10 (d=%d k=%d)", d, k);
  else
    RAG_REQUIRE(tc_shape_ok(d, k), RAG_EUNSUPPORTED,
                "cosine_topk: tensor-core modes cover d <= 128 with k <= 128 and d <= 256 with k <= 10 (d=%d k=%d)", d, k);
  RAG_REQUIRE(key_inv_norm || dot, RAG_EINVAL, "cosine_topk: the tensor-core modes need key_inv_norm (rag_row_inv_norm_f32)");
  RAG_REQUIRE(aligned16(keys_shadow), RAG_EALIGN, "cosine_topk: the key shadow must be 16-byte aligned");
  const TcDispatch D = tc_dispatch(Q, N, d, k, mode, flags, mask_rowptr != nullptr);
  const TcPlan& pts = D.pts;                                    // TS plan: workspace layout + the second pass
  const TcPlan& p = D.p;                                        // plan of the kernel that runs
  const TcOptions& opt = tc_opts();
  const bool two_pass = D.two_pass, ts = D.ts;
  const size_t need = tf32 ? p.total : pts.total;
  RAG_REQUIRE(ws_bytes >= need, RAG_EWORKSPACE, "cosine_topk: workspace %zu < %zu bytes", ws_bytes, need);
  RAG_REQUIRE(ws && (reinterpret_cast<uintptr_t>(ws) & 255u) == 0, RAG_EALIGN, "cosine_topk: workspace must be 256-byte aligned");
  RAG_REQUIRE(p.smem <= (size_t)max_smem_optin() && pts.smem <= (size_t)max_smem_optin(), RAG_EUNSUPPORTED,
              "cosine_topk: needs %zu bytes of shared memory", p.smem);
  // every offset below comes from the layout plan L: the TS plan for the 16-bit modes (its list areas are at least as
  // large as the SS kernel's: same kp, n_splits), the tf32 plan otherwise
  const TcPlan& L = tf32 ? p : pts;
  unsigned char* w = static_cast<unsigned char*>(ws);
  uint16_t* q_bf = reinterpret_cast<uint16_t*>(w + L.off_qbf);
  float* qinv = reinterpret_cast<float*>(w + L.off_qinv);
  float* qerr = reinterpret_cast<float*>(w + L.off_qerr);
  int32_t* zero = reinterpret_cast<int32_t*>(w + L.off_zero);
  int32_t* fb_count = zero;            // rows for the second pass
  int32_t* fb2_count = zero + 1;       // rows for the fp32 kernel
  int32_t* spill_cnt = zero + TC_ZERO_HDR;
  const int d_pad_keys = tf32 ? topk_tc_tf32_dpad(d) : (d + 63) / 64 * 64;        // layout of the caller's shadow
  RAG_REQUIRE(p.d_pad == d_pad_keys, RAG_EUNSUPPORTED, "internal: d_pad mismatch");

  int st;
  if (tf32) {
    st = rag_rows_to_tf32(q, Q, d, 1, 1e-12f, reinterpret_cast<float*>(q_bf), p.d_pad, s);
    if (st) return st;
    st = rag_row_inv_norm_f32(q, Q, d, 1e-12f, qinv, s);
    if (st) return st;
    cudaError_t e = cudaMemsetAsync(zero, 0, 8, s);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(fb_count)");
  } else {
    // ONE prologue launch: normalised 16-bit query image, inverse norms, rounding-error norms, device counters cleared
    st = rows_to_16_launch(q, Q, d, f16 ? RAG_FMT_F16 : RAG_FMT_BF16, 1, 1e-12f, q_bf, p.d_pad, qinv, qerr, nullptr,
                           reinterpret_cast<uint32_t*>(zero), (ts && p.n_mergers > 0) ? p.n_zero_share : L.n_zero, s);
    if (st) return st;
  }

  CUtensorMap mq, mk;
  st = make_map_bf16(&mk, keys_shadow, N, p.d_pad, TC_BN, tf32);
  if (st) return st;

  TcArgs a{};
  a.Q = Q; a.N = N; a.n_qtiles = p.n_qtiles; a.n_splits = p.n_splits; a.tiles_per_split = p.tiles_per_split;
  a.n_tiles = p.n_tiles; a.kp = p.kp;
  a.idesc = tf32 ? TC_IDESC_TF32 : (f16 ? TC_IDESC_F16 : TC_IDESC);
  // Diagnostic switches (tools/ only: MMA-only ceilings, pipeline trace, operand experiments).  They change or void the
  // results, so they are honoured only in a process started with RAG_DIAG=1 (read once); production calls never look.
  static const bool diag = [] { const char* e = getenv("RAG_DIAG"); return e && atoi(e) != 0; }();
  if (diag) {
    { const char* dbg = getenv("RAG_TC_DEBUG"); a.debug = dbg ? atoi(dbg) : 0; }
    { const char* tr = getenv("RAG_TC_TRACE"); a.trace = (tr && atoi(tr)) ? 1 : 0; if (a.debug == 3) { a.debug = 0; a.trace = 1; } }
    { const char* t0 = getenv("RAG_TC_TRACE_T0"); a.trace_t0 = t0 ? atoi(t0) : 0; }
    { const char* nt = getenv("RAG_TC_NOTMA"); a.no_tma = nt ? atoi(nt) : 0; }
    { const char* sw = getenv("RAG_TS_SWAP"); a.swap_halves = sw ? atoi(sw) : 0; }
  }
  a.part_s = reinterpret_cast<float*>(w + L.off_ps);
  a.part_i = reinterpret_cast<int32_t*>(w + L.off_pi);
  a.overflow = reinterpret_cast<float*>(w + L.off_ovf);
  if (two_pass) {
    // ---- short streams: maxima pass + collect pass ----------------------------------------------------------------------
    // A CTA that sees a few dozen key tiles spends its time WARMING UP lists (every score is a candidate until a row's slots
    // are full; measured at the reference's Cora shape, 2 708 x 10 832 x 256: 95 of 160 us).  Two passes over the same
    // stream need no lists at all: (1) the pre-pass instantiation over the WHOLE stream keeps one maximum per (row, tile
    // group) -- registers only; (2) the k-th largest of a row's G group maxima, each the 16-bit score of a distinct key,
    // minus 2 eps is a collect threshold above which every member of the exact top k lies; (3) the collect instantiation
    // rescans with that fixed threshold and returns the handful of keys above it; (4) refine2 re-scores exactly what came
    // back.  Overflowing rows (> spill_cap keys above the bound: clusters) go to the fp32 kernel as always.
    const unsigned grid = (unsigned)(pts.n_qtiles * pts.n_splits);
    int g = std::min(pts.tiles_per_split, 256 / pts.n_splits);
    TcArgs pre = a;
    pre.premax = 1; pre.pre_tiles = pts.tiles_per_split; pre.pre_groups = g;
    pre.gmax = reinterpret_cast<float*>(w + L.off_gmax);
    pre.trace = 0;
    st = run_ts(mk, q_bf, pre, pts, grid, s);
    if (st) return st;
    float* thr = reinterpret_cast<float*>(w + L.off_thr2);
    sample_threshold_kernel<<<(unsigned)((Q + 7) / 8), 256, 0, s>>>(pre.gmax, g * pts.n_splits, Q, k, thr, 1, qerr, shadow_err,
                                                                   TC_SLACK + (shadow_err ? 0.f : (f16 ? TC_U_F16 : TC_U_BF16)));
    RAG_LAUNCH_OK("sample_threshold_kernel");
    TcArgs c = a;
    c.thr0 = thr; c.thr_exact = 1; c.collect = 1; c.premax = 0; c.trace = 0;
    c.spill_s = reinterpret_cast<float*>(w + L.off_spill_s);
    c.spill_i = reinterpret_cast<int32_t*>(w + L.off_spill_i);
    c.spill_cnt = spill_cnt; c.spill_cap = pts.spill_cap;
    c.kp = pts.kp;
    st = run_ts(mk, q_bf, c, pts, grid, s);
    if (st) return st;
    Refine2Args r2{};
    r2.q = q; r2.keys = keys; r2.q_inv_norm = qinv; r2.key_inv_norm = key_inv_norm;
    r2.d = d; r2.k = k; r2.idx_offset = idx_offset;
    r2.rows = nullptr; r2.n_rows_dev = nullptr; r2.n_rows = (int)Q;
    r2.spill_s = c.spill_s; r2.spill_i = c.spill_i; r2.spill_cnt = spill_cnt; r2.spill_cap = pts.spill_cap;
    r2.out_scores = out_scores; r2.out_idx = out_idx;
    // overflowing rows (clusters: more than spill_cap keys above the group-maximum bound) go to the list of the standard
    // second pass with a threshold from their exact scores -- or, in small libraries, straight to the fp32 kernel
    const bool retry = opt.pass2 && N >= TC_PASS2_MIN_KEYS;
    r2.fb_rows = reinterpret_cast<int32_t*>(w + L.off_fb); r2.fb_count = fb_count;
    if (retry) {
      r2.retry_rows = reinterpret_cast<int32_t*>(w + L.off_fb); r2.retry_thr = thr; r2.retry_count = fb_count;
      r2.qerr = qerr; r2.kerr_max = shadow_err; r2.eps_fixed = TC_SLACK + (shadow_err ? 0.f : (f16 ? TC_U_F16 : TC_U_BF16));
    }
    int64_t b2 = (Q + 7) / 8;
    const int64_t cap2 = (int64_t)sm_count() * 8;
    if (b2 > cap2) b2 = cap2;
    refine2_kernel<<<(unsigned)b2, 256, (size_t)8 * k * 12, s>>>(r2);
    RAG_LAUNCH_OK("refine2_kernel");
    if (retry) {                                              // the second pass indexes the spill areas by list slot
      cudaError_t e = cudaMemsetAsync(spill_cnt, 0, (size_t)Q * 4, s);
      if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(spill_cnt)");
    }
  }
  if (two_pass) {
    // (front end done: the shared tail below picks up the retry list)
  } else if (ts) {
    const unsigned grid = (unsigned)(p.n_qtiles * p.n_splits);
    if (p.pre_tiles > 0 && tc_opts().prepass && a.debug == 0) {
      // threshold pre-pass over the first 1/64 of every split (group maxima only), then the per-row bound
      TcArgs pre = a;
      pre.premax = 1; pre.pre_tiles = p.pre_tiles; pre.pre_groups = p.pre_groups;
      pre.gmax = reinterpret_cast<float*>(w + L.off_gmax);
      pre.trace = 0;
      st = run_ts(mk, q_bf, pre, p, grid, s);
      if (st) return st;
      float* thr0 = reinterpret_cast<float*>(w + L.off_thr0);
      sample_threshold_kernel<<<(unsigned)((Q + 7) / 8), 256, 0, s>>>(pre.gmax, p.pre_groups * p.n_splits, Q, p.kp, thr0);
      RAG_LAUNCH_OK("sample_threshold_kernel");
      a.thr0 = thr0;
    }
    unsigned grid_main = grid;
    a.hit_count = reinterpret_cast<uint32_t*>(zero + 4);
    // (exclusion lists: the k-th best ADMISSIBLE key may sit far below the k-th best key the sweep would share -- no sweep)
    if (p.n_mergers > 0 && a.debug == 0 && !mask_rowptr) {
      // cross-split threshold sharing: p.n_mergers extra CTAs sweep the published candidates (TcArgs::pool)
      a.pool = reinterpret_cast<uint32_t*>(w + L.off_pool);
      a.gthr = reinterpret_cast<uint32_t*>(w + L.off_gthr);
      a.part_thr = reinterpret_cast<float*>(w + L.off_pthr);
      a.done = zero + 3;
      a.pool_slots = ts_kpb(p.kp); a.n_mergers = p.n_mergers; a.g_k = k;
      a.g_qerr = qerr; a.g_kerr = shadow_err; a.g_eps_fixed = TC_SLACK + (shadow_err ? 0.f : (f16 ? TC_U_F16 : TC_U_BF16));
      a.g_exact = exact ? 1 : 0; a.g_dbg = tc_opts().gshare_dbg;
      grid_main = grid + (unsigned)p.n_mergers;
    }
    st = run_ts(mk, q_bf, a, p, grid_main, s);
  } else {
    st = make_map_bf16(&mq, q_bf, Q, p.d_pad, 128, tf32);
    if (st) return st;
#define RAG_TC_CASE(KH_, NS_, KP_) \
  if (p.kh == KH_ && p.nstage == NS_ && p.kp == KP_) \
    st = tf32 ? launch_filter<KH_, NS_, KP_, true>(mq, mk, a, p, s) : launch_filter<KH_, NS_, KP_, false>(mq, mk, a, p, s); \
  else
    RAG_TC_CASE(1, 8, 16) RAG_TC_CASE(1, 7, 32) RAG_TC_CASE(2, 7, 16) RAG_TC_CASE(2, 5, 32)
    RAG_TC_CASE(4, 3, 16)
    return fail(RAG_EUNSUPPORTED, "cosine_topk: no tensor-core instantiation for kh=%d nstage=%d kp=%d", p.kh, p.nstage, p.kp);
#undef RAG_TC_CASE
  }
  if (st) return st;

  RefineArgs r{};
  r.q = q; r.keys = keys; r.q_inv_norm = qinv; r.key_inv_norm = key_inv_norm;
  r.Q = Q; r.N = N; r.d = d; r.k = k; r.kp = p.kp; r.n_splits = p.n_splits;
  r.part_s = a.part_s; r.part_i = a.part_i; r.exact = exact ? 1 : 0;
  r.idx_offset = idx_offset; r.out_scores = out_scores; r.out_idx = out_idx;
  r.fb_rows = reinterpret_cast<int32_t*>(w + L.off_fb);
  r.fb_count = fb_count;
  r.thr0 = a.thr0;
  r.part_thr = a.part_thr;
  r.mask_rowptr = mask_rowptr; r.mask_col = mask_col; r.key_scale = dot ? key_scale : 0.f;
  r.thr2 = reinterpret_cast<float*>(w + L.off_thr2);
  const float u = f16 ? TC_U_F16 : TC_U_BF16;
  r.qerr = tf32 ? nullptr : qerr;
  r.kerr_max = shadow_err;
  r.eps_fixed = TC_SLACK + (shadow_err ? 0.f : u);
  r.loose_mult = 8.0f; r.loose_count = zero + 2;
  const int total_c = p.n_splits * p.kp;
  const int wpb = (total_c <= 1024) ? 8 : 2;                    // per-warp smem: k*12 + total*4 bytes (< 48 KB per CTA)
  int64_t blocks = (Q + wpb - 1) / wpb;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  RAG_REQUIRE(total_c <= 65535, RAG_EUNSUPPORTED, "cosine_topk: %d candidates per row", total_c);
  if (!two_pass) {                                              // (two-pass mode: its own refine already ran, r only carries the lists)
    refine_kernel<<<(unsigned)blocks, wpb * 32, (size_t)wpb * (k * 12 + total_c * 4 + REFINE_SEL_CAP * 2), s>>>(r);
    RAG_LAUNCH_OK("refine_kernel");
  }
  if (!exact) return RAG_OK;

  const int32_t* fp32_rows = r.fb_rows;
  const int32_t* fp32_count = fb_count;
  // (libraries below TC_PASS2_MIN_KEYS: the fp32 kernel rescans the few uncertified rows faster than two more launches cost
  //  -- the reference's own shapes are launch-latency problems; the first counter then still holds the uncertified rows)
  if (tc_opts().pass2 && a.debug == 0 && N >= TC_PASS2_MIN_KEYS) {
    // ---- second pass: the uncertified rows (device-side list, usually empty: the CTAs read the count and leave) rescan
    //      the shard on the tensor cores with the fixed threshold thr2 and return EVERY key above it ----
    TcArgs c = a;
    c.thr0 = r.thr2; c.thr_exact = 1; c.collect = 1; c.premax = 0; c.trace = 0;
    c.hit_count = nullptr; c.pool = nullptr; c.gthr = nullptr; c.part_thr = nullptr; c.done = nullptr; c.n_mergers = 0;
    c.row_map = r.fb_rows; c.n_rows_dev = fb_count;
    c.spill_s = reinterpret_cast<float*>(w + L.off_spill_s);
    c.spill_i = reinterpret_cast<int32_t*>(w + L.off_spill_i);
    c.spill_cnt = spill_cnt; c.spill_cap = pts.spill_cap;
    c.kp = pts.kp;
    const unsigned grid2 = (unsigned)std::max(sm_count(), pts.n_qtiles);
    st = run_ts(mk, q_bf, c, pts, grid2, s);
    if (st) return st;
    Refine2Args r2{};
    r2.q = q; r2.keys = keys; r2.q_inv_norm = qinv; r2.key_inv_norm = key_inv_norm;
    r2.d = d; r2.k = k; r2.idx_offset = idx_offset;
    r2.rows = r.fb_rows; r2.n_rows_dev = fb_count;
    r2.spill_s = c.spill_s; r2.spill_i = c.spill_i; r2.spill_cnt = spill_cnt; r2.spill_cap = pts.spill_cap;
    r2.out_scores = out_scores; r2.out_idx = out_idx;
    r2.fb_rows = reinterpret_cast<int32_t*>(w + L.off_fb2); r2.fb_count = fb2_count;
    r2.mask_rowptr = mask_rowptr; r2.mask_col = mask_col; r2.key_scale = r.key_scale;
    int64_t b2 = (Q + 7) / 8;
    const int64_t cap2 = (int64_t)sm_count() * 8;
    if (b2 > cap2) b2 = cap2;
    refine2_kernel<<<(unsigned)b2, 256, (size_t)8 * k * 12, s>>>(r2);
    RAG_LAUNCH_OK("refine2_kernel");
    fp32_rows = r2.fb_rows;
    fp32_count = fb2_count;
  }
  // rows still open (spill overflow: more than spill_cap near-ties) are recomputed by the fp32 kernel
  // (dot-product ranking: the fp32 kernel scores q . k directly, no norms)
  return topk_f32_run_rows(q, Q, keys, dot ? nullptr : key_inv_norm, dot ? nullptr : qinv, N, d, k, idx_offset, fp32_rows,
                           fp32_count, out_scores, out_idx, w + L.off_f32, ws_bytes - L.off_f32, s, mask_rowptr, mask_col);
}

}  // namespace rag

// Introspection: which kernel would serve a call, and with what geometry.  Pure host arithmetic (148 SMs are assumed without
// a device), so the dispatch rules are testable on a CPU-only box.  out[0] = kernel (0 = not a tensor-core shape: the fp32
// kernel; 1 = SS, query tile in shared memory; 2 = TS, query tile in tensor memory; 3 = two-pass mode on the TS kernel),
// out[1] = query tiles, out[2] = key splits, out[3] = key tiles per CTA, out[4] = list length k', out[5] = sweeping CTAs of the
// cross-split threshold sharing, out[6] = tiles of the threshold pre-pass per CTA, out[7] = pipeline stages.
extern "C" RAG_API int rag_cosine_topk_plan(int64_t Q, int64_t N, int32_t d, int32_t k, int32_t mode, uint32_t flags,
                                            int32_t has_mask, int32_t* out) {
  using namespace rag;
  RAG_REQUIRE(out, RAG_EINVAL, "cosine_topk_plan: null pointer");
  for (int i = 0; i < 8; ++i) out[i] = 0;
  if (Q <= 0 || N <= 0 || mode == RAG_SIM_FP32) return RAG_OK;
  const bool tf32 = (mode == RAG_SIM_TF32);
  if (tf32 ? !tc_shape_ok_tf32(d, k) : !tc_shape_ok(d, k)) return RAG_OK;
  const TcDispatch D = tc_dispatch(Q, N, d, k, mode, flags, has_mask != 0);
  const bool sweep = D.ts && !D.two_pass && !has_mask && D.p.n_mergers > 0;
  out[0] = D.two_pass ? 3 : (D.ts ? 2 : 1);
  out[1] = D.p.n_qtiles; out[2] = D.p.n_splits; out[3] = D.p.tiles_per_split; out[4] = D.p.kp;
  out[5] = sweep ? D.p.n_mergers : 0;
  out[6] = (D.ts && !D.two_pass && tc_opts().prepass) ? D.p.pre_tiles : 0;
  out[7] = D.p.nstage;
  return RAG_OK;
}

// diagnostics: copy the pipeline trace (4 x 512 clock64 stamps + 2 x 512 uint32 + 16 x 256 uint32) to the host (36 KB)
extern "C" RAG_API int rag_tc_trace_read(unsigned long long* host_out) {
  cudaError_t e = cudaMemcpyFromSymbol(host_out, rag::g_tc_trace, sizeof(unsigned long long) * 4 * 512);
  if (e == cudaSuccess) e = cudaMemcpyFromSymbol(host_out + 4 * 512, rag::g_tc_trace2, sizeof(unsigned int) * 2 * 512);
  if (e == cudaSuccess) e = cudaMemcpyFromSymbol(host_out + 4 * 512 + 512, rag::g_tc_trace3, sizeof(unsigned int) * 16 * 256);
  return e == cudaSuccess ? RAG_OK : rag::cuda_fail(e, "rag_tc_trace_read");
}

extern "C" RAG_API int rag_tc_set_option(const char* name, int32_t value) {
  if (!name) return rag::fail(RAG_EINVAL, "tc_set_option: null name");
  rag::TcOptions& o = rag::tc_opts();
  const rag::TcOptions dflt;
  if (!strcmp(name, "variant")) o.variant = (value >= 0 && value <= 2) ? value : dflt.variant;
  else if (!strcmp(name, "prepass")) o.prepass = value < 0 ? dflt.prepass : (value != 0);
  else if (!strcmp(name, "prepass_min_tiles")) o.prepass_min_tiles = value >= 64 ? value : dflt.prepass_min_tiles;
  else if (!strcmp(name, "prepass_div")) o.prepass_div = value >= 4 ? value : dflt.prepass_div;
  else if (!strcmp(name, "prepass_max_tiles")) o.prepass_max_tiles = value >= 16 ? value : dflt.prepass_max_tiles;
  else if (!strcmp(name, "kp")) o.kp = (value == 16 || value == 32) ? value : dflt.kp;
  else if (!strcmp(name, "pass2")) o.pass2 = value < 0 ? dflt.pass2 : (value != 0);
  else if (!strcmp(name, "gshare")) o.gshare = value < 0 ? dflt.gshare : (value != 0);
  else if (!strcmp(name, "twopass")) o.twopass = value < 0 ? dflt.twopass : (value != 0);
  else if (!strcmp(name, "twopass_max_tiles")) o.twopass_max_tiles = value >= 1 ? value : dflt.twopass_max_tiles;
  else if (!strcmp(name, "gshare_dbg")) o.gshare_dbg = value >= 0 ? value : 0;
  else if (!strcmp(name, "gshare_ctas")) o.gshare_ctas = (value >= 1 && value <= 16) ? value : dflt.gshare_ctas;
  else return rag::fail(RAG_EINVAL, "tc_set_option: unknown option '%s'", name);
  return RAG_OK;
}
