// K1+K2+K3 on CUDA cores (fp32 FMA): fused L2-normalised similarity + per-row top-k, plus the
// materialising variant kept for API compatibility and the candidate-merge kernel used after
// key-split partials and after the multi-GPU all-gather.
//
// Replaces SimilarityFunctions.calculate_cosine_similarity + torch.topk
// (RAGraph_node/ragraph_utils/SimilarityFunctions.py:6-16, ToyGraphBase.py:53,67;
//  RAGraph_edge/modules/RAGraph.py:303,311).  This is the RAG_SIM_FP32 mode: the exact-order
// fp32 path that defines parity and re-computes rows the tensor-core filter cannot certify.
//
// Tiling: a CTA owns BM=64 queries and streams its key range in BN=128-key tiles.  The
// normalised query block lives transposed in shared memory for the whole kernel
// (Qs[d][64]); key tiles are staged transposed in KC=32-column chunks (Bs[32][128+4]),
// register-prefetched one chunk ahead.  Each thread accumulates a 4x8 micro-tile.  Scores
// are scaled by the key inverse norm and compared with the row's running k-th best
// (threshold in shared memory): only when some thread of the CTA sees a candidate is the
// tile spilled to shared memory and the per-row sorted lists updated by warp-cooperative
// insertion.  Scores never reach HBM.
#include <cfloat>
#include "common.cuh"

namespace rag {

constexpr int TK_BM = 64, TK_BN = 128, TK_KC = 32, TK_THREADS = 256;
constexpr int TK_LDB = TK_BN + 4;
constexpr int TK_LDS = TK_BN + 1;      // score tile leading dim (conflict-free row scans)

struct TopkArgs {
  const float* q; int64_t Q;
  const float* keys; const float* key_inv_norm;   // key_inv_norm == nullptr -> dot product
  const float* q_inv_norm;                        // nullptr -> dot product
  int64_t N; int d; int k;
  int64_t keys_per_split; int n_splits;
  int64_t idx_offset;
  float* out_scores; int64_t* out_idx;            // [n_splits, Q, k] (partials) or [Q, k]
  float* dense_out;                               // MATERIALIZE: [Q, N]
  const int32_t* row_map;                         // optional: compact row r -> query row row_map[r]
  const int32_t* n_rows_dev;                      // optional: device-side count of rows in row_map
  // optional exclusion lists (edge evaluation: a user's history items never rank, utils/metrics.py:48-53,113-115):
  // query row r must not return the (global) key indices mask_col[mask_rowptr[r] .. mask_rowptr[r+1])
  const int64_t* mask_rowptr; const int64_t* mask_col;
};

template <bool MATERIALIZE>
__global__ void __launch_bounds__(TK_THREADS, 2) cosine_topk_f32_kernel(const TopkArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int d = a.d, k = a.k;
  const int dk = (d + TK_KC - 1) / TK_KC * TK_KC;             // d rounded up to the chunk
  float* Qs = reinterpret_cast<float*>(smem_raw);             // [dk][BM]
  float* Bs = Qs + (size_t)dk * TK_BM;                        // [2][KC][LDB]
  float* Ss = Bs + 2 * TK_KC * TK_LDB;                        // [BM][LDS] score tile
  float* thr = Ss + TK_BM * TK_LDS;                           // [BM]
  float* lvals = thr + TK_BM;                                 // [BM][k]
  int32_t* lidx = reinterpret_cast<int32_t*>(lvals + (size_t)TK_BM * k);   // [BM][k]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = tid & 15, ty = tid >> 4;                     // 16 x 16 threads: 8 cols x 4 rows each
  const int64_t q0 = (int64_t)blockIdx.x * TK_BM;
  const int split = blockIdx.y;
  const int64_t key_lo = (int64_t)split * a.keys_per_split;
  const int64_t key_hi = min(key_lo + a.keys_per_split, a.N);
  const bool vec4 = (d & 3) == 0;
  // restricted-row mode (rows the tensor-core certificate rejected): the row count lives on the device
  const int64_t Qeff = a.n_rows_dev ? (int64_t)__ldg(a.n_rows_dev) : a.Q;
  if (q0 >= Qeff) return;

  // ---- stage the normalised query block, transposed ------------------------------------
  for (int i = tid; i < dk * TK_BM; i += TK_THREADS) {
    const int r = i / dk, c = i - r * dk;                     // coalesced over c (query row)
    float v = 0.f;
    if (q0 + r < Qeff && c < d) {
      const int64_t qr = a.row_map ? (int64_t)__ldg(a.row_map + q0 + r) : q0 + r;
      v = __ldg(a.q + qr * d + c);
      if (a.q_inv_norm) v *= __ldg(a.q_inv_norm + qr);
    }
    Qs[(size_t)c * TK_BM + r] = v;
  }
  if (!MATERIALIZE) {
    for (int i = tid; i < TK_BM * k; i += TK_THREADS) { lvals[i] = -FLT_MAX; lidx[i] = 0x7fffffff; }
    if (tid < TK_BM) thr[tid] = -FLT_MAX;
  }
  __syncthreads();

  const int n_chunks = dk / TK_KC;
  // loader mapping: thread -> (key n = tid/8 + 32*p, float4 column c4 = tid%8), p = 0..3
  const int ld_c4 = tid & 7, ld_n = tid >> 3;

  for (int64_t n0 = key_lo; n0 < key_hi; n0 += TK_BN) {
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float4 pre[4];
    auto load_chunk = [&](int ch) {
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int64_t n = n0 + ld_n + 32 * p;
        const int c = ch * TK_KC + ld_c4 * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n < key_hi) {
          const float* src = a.keys + n * d + c;
          if (vec4) { if (c < d) v = __ldg(reinterpret_cast<const float4*>(src)); }
          else {
            if (c < d) v.x = __ldg(src);
            if (c + 1 < d) v.y = __ldg(src + 1);
            if (c + 2 < d) v.z = __ldg(src + 2);
            if (c + 3 < d) v.w = __ldg(src + 3);
          }
        }
        pre[p] = v;
      }
    };
    auto store_chunk = [&](int buf) {
      float* B = Bs + buf * TK_KC * TK_LDB;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int n = ld_n + 32 * p, c = ld_c4 * 4;
        B[(c + 0) * TK_LDB + n] = pre[p].x; B[(c + 1) * TK_LDB + n] = pre[p].y;
        B[(c + 2) * TK_LDB + n] = pre[p].z; B[(c + 3) * TK_LDB + n] = pre[p].w;
      }
    };

    load_chunk(0);
    store_chunk(0);
    __syncthreads();
    for (int ch = 0; ch < n_chunks; ++ch) {
      const int buf = ch & 1;
      if (ch + 1 < n_chunks) load_chunk(ch + 1);
      const float* B = Bs + buf * TK_KC * TK_LDB;
      const float* A = Qs + (size_t)ch * TK_KC * TK_BM;
#pragma unroll 8
      for (int kk = 0; kk < TK_KC; ++kk) {
        const float4 av = *reinterpret_cast<const float4*>(A + kk * TK_BM + ty * 4);
        const float4 b0 = *reinterpret_cast<const float4*>(B + kk * TK_LDB + tx * 8);
        const float4 b1 = *reinterpret_cast<const float4*>(B + kk * TK_LDB + tx * 8 + 4);
        const float ar[4] = {av.x, av.y, av.z, av.w};
        const float br[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
      }
      if (ch + 1 < n_chunks) store_chunk(buf ^ 1);
      __syncthreads();
    }

    // ---- epilogue: scale by key inverse norms, filter against the running thresholds ----
    float kin[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int64_t n = n0 + tx * 8 + j;
      kin[j] = (a.key_inv_norm && n < key_hi) ? __ldg(a.key_inv_norm + n) : 1.0f;
    }
    if (MATERIALIZE) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t r = q0 + ty * 4 + i;
        if (r >= a.Q) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int64_t n = n0 + tx * 8 + j;
          if (n < key_hi) a.dense_out[r * a.N + n] = acc[i][j] * kin[j];
        }
      }
      continue;
    }
    bool cand = false;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float t = thr[ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int64_t n = n0 + tx * 8 + j;
        float s = acc[i][j] * kin[j];
        if (n >= key_hi) s = -FLT_MAX;            // padding never qualifies
        acc[i][j] = s;
        cand |= (s > t);
      }
    }
    if (__syncthreads_or(cand)) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) Ss[(ty * 4 + i) * TK_LDS + tx * 8 + j] = acc[i][j];
      __syncthreads();
      // warp w owns rows w*8 .. w*8+7; keys are visited in ascending index so a later key
      // that only ties the k-th best never displaces it (order: score desc, index asc)
      for (int rr = 0; rr < 8; ++rr) {
        const int row = warp * 8 + rr;
        float* rv = lvals + (size_t)row * k;
        int32_t* ri = lidx + (size_t)row * k;
        float t = thr[row];
        for (int c0 = 0; c0 < TK_BN; c0 += 32) {
          const float s = Ss[row * TK_LDS + c0 + lane];
          unsigned m = __ballot_sync(0xffffffffu, s > t);
          while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const float sv = __shfl_sync(0xffffffffu, s, src);
            if (sv > t) {                         // t may have risen since the ballot
              bool excluded = false;
              if (a.mask_rowptr && q0 + row < Qeff) {         // only candidates pay for the membership test
                const int64_t qr = a.row_map ? (int64_t)__ldg(a.row_map + q0 + row) : q0 + row;
                const int64_t key = a.idx_offset + n0 + c0 + src;
                const int64_t mlo = __ldg(a.mask_rowptr + qr), mhi = __ldg(a.mask_rowptr + qr + 1);
                for (int64_t m0 = mlo; m0 < mhi && !excluded; m0 += 32)
                  excluded = __any_sync(0xffffffffu, m0 + lane < mhi && __ldg(a.mask_col + m0 + lane) == key);
              }
              if (!excluded) {
                warp_sorted_insert<int32_t>(rv, ri, k, sv, (int32_t)(n0 - key_lo + c0 + src), lane);
                t = rv[k - 1];
              }
            }
          }
        }
        __syncwarp();                             // every lane has read thr[row] (racecheck: intra-warp WAR)
        if (lane == 0) thr[row] = t;
      }
      __syncthreads();
    }
  }
  if (MATERIALIZE) return;

  // ---- write this split's sorted candidates -------------------------------------------
  __syncthreads();
  for (int i = tid; i < TK_BM * k; i += TK_THREADS) {
    const int r = i / k, p = i - r * k;
    if (q0 + r >= Qeff) continue;
    int64_t orow = q0 + r;
    if (a.n_splits == 1 && a.row_map) orow = __ldg(a.row_map + q0 + r);   // direct write to the final row
    const size_t o = ((size_t)split * a.Q + orow) * k + p;
    const float v = lvals[i];
    const int32_t li = lidx[i];
    a.out_scores[o] = v;
    a.out_idx[o] = (li == 0x7fffffff) ? (int64_t)-1 : a.idx_offset + key_lo + li;
  }
}

static size_t tk_smem_host(int d, int k, bool materialize) {
  const int dk = (d + TK_KC - 1) / TK_KC * TK_KC;
  size_t b = (size_t)dk * TK_BM * 4 + 2 * TK_KC * TK_LDB * 4;
  if (!materialize) b += (size_t)TK_BM * TK_LDS * 4 + TK_BM * 4 + (size_t)TK_BM * k * 8;
  return b;
}

// ---- merge: [R, Q, k_in] candidates -> [Q, k_out], order score desc / index asc ---------
// one warp per query row; list in shared memory.  Entries with idx < 0 are padding.
__global__ void __launch_bounds__(256) topk_merge_kernel(const float* __restrict__ scores,
                                                         const int64_t* __restrict__ idx, int R, int64_t Q,
                                                         int k_in, int k_out, float* __restrict__ out_scores,
                                                         int64_t* __restrict__ out_idx,
                                                         const int32_t* __restrict__ row_map,
                                                         const int32_t* __restrict__ n_rows_dev) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  int64_t* li = reinterpret_cast<int64_t*>(smem_raw) + (size_t)warp * k_out;
  float* lv = reinterpret_cast<float*>(reinterpret_cast<int64_t*>(smem_raw) + (size_t)wpb * k_out) +
              (size_t)warp * k_out;
  const int64_t Qeff = n_rows_dev ? (int64_t)__ldg(n_rows_dev) : Q;
  for (int64_t q = (int64_t)blockIdx.x * wpb + warp; q < Qeff; q += (int64_t)gridDim.x * wpb) {
    for (int p = lane; p < k_out; p += 32) { lv[p] = -FLT_MAX; li[p] = INT64_MAX; }
    __syncwarp();
    const int total = R * k_in;
    for (int c0 = 0; c0 < total; c0 += 32) {
      const int c = c0 + lane;
      float s = -FLT_MAX; int64_t j = -1;
      if (c < total) {
        const int r = c / k_in, p = c - r * k_in;
        const size_t o = ((size_t)r * Q + q) * k_in + p;
        s = __ldg(scores + o); j = __ldg(idx + o);
      }
      const float ts = lv[k_out - 1]; const int64_t ti = li[k_out - 1];
      unsigned m = __ballot_sync(0xffffffffu, j >= 0 && ranks_before(s, j, ts, ti));
      while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const float sv = __shfl_sync(0xffffffffu, s, src);
        const int64_t jv = __shfl_sync(0xffffffffu, j, src);
        if (ranks_before(sv, jv, lv[k_out - 1], li[k_out - 1]))
          warp_sorted_insert<int64_t>(lv, li, k_out, sv, jv, lane);
      }
    }
    const int64_t orow = row_map ? (int64_t)__ldg(row_map + q) : q;
    for (int p = lane; p < k_out; p += 32) {
      out_scores[orow * k_out + p] = lv[p];
      out_idx[orow * k_out + p] = (li[p] == INT64_MAX) ? (int64_t)-1 : li[p];
    }
    __syncwarp();
  }
}

static int launch_merge(const float* scores, const int64_t* idx, int R, int64_t Q, int k_in, int k_out,
                        float* out_scores, int64_t* out_idx, cudaStream_t s, const int32_t* row_map = nullptr,
                        const int32_t* n_rows_dev = nullptr) {
  const int wpb = 8;
  int64_t blocks = (Q + wpb - 1) / wpb;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  const size_t smem = (size_t)wpb * k_out * 12;
  topk_merge_kernel<<<(unsigned)blocks, wpb * 32, smem, s>>>(scores, idx, R, Q, k_in, k_out, out_scores, out_idx,
                                                                row_map, n_rows_dev);
  RAG_LAUNCH_OK("topk_merge_kernel");
  return RAG_OK;
}

// workspace carving shared by the query and the launch
struct TkPlan {
  int n_splits; int64_t keys_per_split;
  size_t off_qinv, off_kinv, off_ps, off_pi, total;
};
static TkPlan tk_plan(int64_t Q, int64_t N, int d, int k, bool need_kinv) {
  TkPlan p{};
  const int64_t qtiles = (Q + TK_BM - 1) / TK_BM;
  const int64_t want = 2LL * sm_count();                       // >= 2 waves of CTAs
  int64_t s = (want + qtiles - 1) / qtiles;
  const int64_t max_s = (N + 4 * TK_BN - 1) / (4 * TK_BN);     // >= 4 key tiles per split
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  if (s > 256) s = 256;
  int64_t per = (N + s - 1) / s;
  per = (per + TK_BN - 1) / TK_BN * TK_BN;
  s = (N + per - 1) / per;
  if (s < 1) s = 1;
  p.n_splits = (int)s; p.keys_per_split = per;
  size_t off = 0;
  p.off_qinv = off; off += align_up((size_t)Q * 4, 256);
  p.off_kinv = off; if (need_kinv) off += align_up((size_t)N * 4, 256);
  p.off_ps = off; off += align_up((size_t)s * Q * k * 4, 256);
  p.off_pi = off; off += align_up((size_t)s * Q * k * 8, 256);
  p.total = off;
  return p;
}

size_t topk_f32_workspace(int64_t Q, int64_t N, int d, int k) { return tk_plan(Q, N, d, k, true).total; }

static int topk_f32_core(const float* q, int64_t Q, const float* keys, const float* key_inv_norm,
                         const float* q_inv_norm, int64_t N, int d, int k, int64_t idx_offset, const TkPlan& p,
                         unsigned char* w, const int32_t* row_map, const int32_t* n_rows_dev, float* out_scores,
                         int64_t* out_idx, cudaStream_t s, const int64_t* mask_rowptr = nullptr,
                         const int64_t* mask_col = nullptr) {
  TopkArgs a{};
  a.mask_rowptr = mask_rowptr; a.mask_col = mask_col;
  a.q = q; a.Q = Q; a.keys = keys; a.key_inv_norm = key_inv_norm; a.q_inv_norm = q_inv_norm;
  a.N = N; a.d = d; a.k = k; a.keys_per_split = p.keys_per_split; a.n_splits = p.n_splits;
  a.idx_offset = idx_offset; a.row_map = row_map; a.n_rows_dev = n_rows_dev;
  const bool direct = p.n_splits == 1;
  a.out_scores = direct ? out_scores : reinterpret_cast<float*>(w + p.off_ps);
  a.out_idx = direct ? out_idx : reinterpret_cast<int64_t*>(w + p.off_pi);
  const size_t smem = tk_smem_host(d, k, false);
  RAG_REQUIRE(smem <= (size_t)max_smem_optin(), RAG_EUNSUPPORTED,
              "cosine_topk(fp32): d=%d k=%d needs %zu bytes of shared memory (> %d)", d, k, smem, max_smem_optin());
  cudaError_t e = cudaFuncSetAttribute(cosine_topk_f32_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(cosine_topk_f32_kernel)");
  dim3 grid((unsigned)((Q + TK_BM - 1) / TK_BM), (unsigned)p.n_splits);
  cosine_topk_f32_kernel<false><<<grid, TK_THREADS, smem, s>>>(a);
  RAG_LAUNCH_OK("cosine_topk_f32_kernel");
  if (!direct)
    return launch_merge(a.out_scores, a.out_idx, p.n_splits, Q, k, k, out_scores, out_idx, s, row_map, n_rows_dev);
  return RAG_OK;
}

int topk_f32_run(const float* q, int64_t Q, const float* keys, const float* key_inv_norm, int64_t N, int d, int k,
                 uint32_t flags, int64_t idx_offset, float* out_scores, int64_t* out_idx, void* ws, size_t ws_bytes,
                 cudaStream_t s, const int64_t* mask_rowptr, const int64_t* mask_col) {
  const bool dot = (flags & RAG_SIM_DOT) != 0;
  const bool need_kinv = !dot && key_inv_norm == nullptr;
  TkPlan p = tk_plan(Q, N, d, k, need_kinv);
  RAG_REQUIRE(ws_bytes >= p.total, RAG_EWORKSPACE, "cosine_topk: workspace %zu < %zu bytes", ws_bytes, p.total);
  RAG_REQUIRE(ws && (reinterpret_cast<uintptr_t>(ws) & 255u) == 0, RAG_EALIGN,
              "cosine_topk: workspace must be 256-byte aligned");
  unsigned char* w = static_cast<unsigned char*>(ws);
  float* qinv = reinterpret_cast<float*>(w + p.off_qinv);
  float* kinv = reinterpret_cast<float*>(w + p.off_kinv);
  if (!dot) {
    int st = rag_row_inv_norm_f32(q, Q, d, 1e-12f, qinv, s);
    if (st) return st;
    if (need_kinv) {
      st = rag_row_inv_norm_f32(keys, N, d, 1e-12f, kinv, s);
      if (st) return st;
    }
  }
  return topk_f32_core(q, Q, keys, dot ? nullptr : (need_kinv ? kinv : key_inv_norm), dot ? nullptr : qinv, N, d, k,
                       idx_offset, p, w, nullptr, nullptr, out_scores, out_idx, s, mask_rowptr, mask_col);
}

// fp32 path over a device-side list of rows (the tensor-core refine's uncertified rows).  The grid covers the
// worst case (all Q rows); CTAs beyond the device-side count exit at once.
size_t topk_f32_rows_workspace(int64_t Q, int64_t N, int d, int k) { return tk_plan(Q, N, d, k, false).total; }

int topk_f32_run_rows(const float* q, int64_t Q, const float* keys, const float* key_inv_norm, const float* q_inv_norm,
                      int64_t N, int d, int k, int64_t idx_offset, const int32_t* row_map, const int32_t* n_rows_dev,
                      float* out_scores, int64_t* out_idx, void* ws, size_t ws_bytes, cudaStream_t s,
                      const int64_t* mask_rowptr, const int64_t* mask_col) {
  TkPlan p = tk_plan(Q, N, d, k, false);
  RAG_REQUIRE(ws_bytes >= p.total, RAG_EWORKSPACE, "cosine_topk(rows): workspace %zu < %zu bytes", ws_bytes, p.total);
  return topk_f32_core(q, Q, keys, key_inv_norm, q_inv_norm, N, d, k, idx_offset, p, static_cast<unsigned char*>(ws),
                       row_map, n_rows_dev, out_scores, out_idx, s, mask_rowptr, mask_col);
}

}  // namespace rag

extern "C" int rag_topk_merge(const float* scores, const int64_t* idx, int32_t R, int64_t Q, int32_t k_in,
                              int32_t k_out, float* out_scores, int64_t* out_idx, rag_stream_t stream) {
  RAG_REQUIRE(R >= 1 && Q >= 0 && k_in >= 1 && k_out >= 1, RAG_EINVAL, "topk_merge: R=%d Q=%lld k_in=%d k_out=%d",
              R, (long long)Q, k_in, k_out);
  RAG_REQUIRE(k_out <= RAG_MAX_K && (int64_t)R * k_in <= 65536 && k_out <= R * k_in, RAG_EUNSUPPORTED,
              "topk_merge: k_out=%d (max %d, <= R*k_in=%d)", k_out, RAG_MAX_K, R * k_in);
  if (Q == 0) return RAG_OK;
  RAG_REQUIRE(scores && idx && out_scores && out_idx, RAG_EINVAL, "topk_merge: null pointer");
  return rag::launch_merge(scores, idx, R, Q, k_in, k_out, out_scores, out_idx, (cudaStream_t)stream);
}

extern "C" size_t rag_cosine_similarity_workspace(int64_t Q, int64_t N) {
  return rag::align_up((size_t)(Q > 0 ? Q : 0) * 4, 256) + rag::align_up((size_t)(N > 0 ? N : 0) * 4, 256);
}

extern "C" int rag_cosine_similarity_f32(const float* q, int64_t Q, const float* keys, int64_t N, int32_t d,
                                         uint32_t flags, float* out, void* workspace, size_t workspace_bytes,
                                         rag_stream_t stream) {
  using namespace rag;
  RAG_REQUIRE(Q >= 0 && N >= 0 && d >= 1, RAG_EINVAL, "cosine_similarity: Q=%lld N=%lld d=%d", (long long)Q,
              (long long)N, d);
  if (Q == 0 || N == 0) return RAG_OK;
  RAG_REQUIRE(q && keys && out, RAG_EINVAL, "cosine_similarity: null pointer");
  RAG_REQUIRE(aligned16(q) && aligned16(keys), RAG_EALIGN, "cosine_similarity: inputs must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  TopkArgs a{};
  a.q = q; a.Q = Q; a.keys = keys; a.N = N; a.d = d; a.k = 0; a.dense_out = out;
  if (!(flags & RAG_SIM_DOT)) {
    RAG_REQUIRE(workspace_bytes >= rag_cosine_similarity_workspace(Q, N), RAG_EWORKSPACE,
                "cosine_similarity: workspace %zu < %zu bytes", workspace_bytes, rag_cosine_similarity_workspace(Q, N));
    RAG_REQUIRE(workspace && (reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, RAG_EALIGN,
                "cosine_similarity: workspace must be 256-byte aligned");
    float* qinv = static_cast<float*>(workspace);
    float* kinv = reinterpret_cast<float*>(static_cast<unsigned char*>(workspace) + align_up((size_t)Q * 4, 256));
    int st = rag_row_inv_norm_f32(q, Q, d, 1e-12f, qinv, s);
    if (!st) st = rag_row_inv_norm_f32(keys, N, d, 1e-12f, kinv, s);
    if (st) return st;
    a.q_inv_norm = qinv; a.key_inv_norm = kinv;
  }
  const size_t smem = tk_smem_host(d, 0, true);
  RAG_REQUIRE(smem <= (size_t)max_smem_optin(), RAG_EUNSUPPORTED, "cosine_similarity: d=%d too large", d);
  cudaError_t e = cudaFuncSetAttribute(cosine_topk_f32_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(cosine_similarity)");
  // split keys over blockIdx.y so a small query batch still fills the machine
  const int64_t qtiles = (Q + TK_BM - 1) / TK_BM;
  int64_t splits = (2LL * sm_count() + qtiles - 1) / qtiles;
  const int64_t max_s = (N + TK_BN - 1) / TK_BN;
  if (splits > max_s) splits = max_s;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  a.keys_per_split = ((N + splits - 1) / splits + TK_BN - 1) / TK_BN * TK_BN;
  splits = (N + a.keys_per_split - 1) / a.keys_per_split;
  a.n_splits = (int)splits;
  dim3 grid((unsigned)qtiles, (unsigned)splits);
  cosine_topk_f32_kernel<true><<<grid, TK_THREADS, smem, s>>>(a);
  RAG_LAUNCH_OK("cosine_similarity kernel");
  return RAG_OK;
}
