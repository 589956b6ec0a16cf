// Edge-variant negative sampling on the device.
//
// Replaces get_train_batch's per-pair Python rejection loop (RAGraph_edge/utils/dataloader.py:140-162): for every
// (user, positive item) pair of the batch draw n items uniformly from [0, num_items) until one is NOT in the user's
// training history -- `np.random.randint` + a Python set lookup per draw, i.e. tens of thousands of interpreter steps
// per 4 096-pair batch on the host, with the GPU idle.  Here one thread owns one (pair, j) slot: counter-based random
// numbers (a 64-bit mix of (seed, slot, attempt): no state, no host RNG), an unbiased range reduction (128-bit multiply,
// with the rejection step that removes the modulo bias), and a binary search of the user's SORTED history row.
// The result has exactly the reference's distribution -- uniform over the items outside the user's history -- but not its
// numpy random stream (that one lives in host memory); tests check the distribution, the exclusion and the determinism.
#include "common.cuh"

namespace rag {

__device__ __forceinline__ uint64_t mix64(uint64_t x) {          // splitmix64 finaliser
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

constexpr int NS_MAX_ATTEMPTS = 1 << 14;

__global__ void __launch_bounds__(256) negative_sample_kernel(const int64_t* __restrict__ users, int64_t M, int n_neg,
                                                              const int64_t* __restrict__ hist_rowptr,
                                                              const int64_t* __restrict__ hist_items, int64_t num_users,
                                                              int64_t num_items, uint64_t seed, int64_t* __restrict__ out) {
  const int64_t total = M * n_neg;
  for (int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; slot < total; slot += (int64_t)gridDim.x * blockDim.x) {
    const int64_t u = __ldg(users + slot / n_neg);
    int64_t lo = 0, hi = 0;
    if (u >= 0 && u < num_users) { lo = __ldg(hist_rowptr + u); hi = __ldg(hist_rowptr + u + 1); }
    int64_t pick = -1;
    const uint64_t base = mix64(seed ^ mix64((uint64_t)slot));
    const uint64_t n = (uint64_t)num_items;
    const uint64_t reject_below = (0 - n) % n;                   // 2^64 mod n: draws whose low product word is below it are biased
    for (int attempt = 0; attempt < NS_MAX_ATTEMPTS; ++attempt) {
      const uint64_t r = mix64(base + (uint64_t)attempt * 0xD1342543DE82EF95ull);
      const uint64_t lo64 = r * n;
      if (lo64 < reject_below) continue;                         // Lemire's unbiased bounded integer
      const int64_t cand = (int64_t)__umul64hi(r, n);
      int64_t a = lo, b = hi;                                    // binary search in the sorted history row
      while (a < b) {
        const int64_t m = (a + b) >> 1;
        if (__ldg(hist_items + m) < cand) a = m + 1; else b = m;
      }
      if (a < hi && __ldg(hist_items + a) == cand) continue;     // in the user's history: draw again
      pick = cand;
      break;
    }
    out[slot] = pick;                                            // -1: the history covers (nearly) every item
  }
}

}  // namespace rag

extern "C" int rag_negative_sample(const int64_t* users, int64_t M, int32_t n_neg, const int64_t* hist_rowptr,
                                   const int64_t* hist_items, int64_t num_users, int64_t num_items, uint64_t seed,
                                   int64_t* out, rag_stream_t stream) {
  RAG_REQUIRE(M >= 0 && n_neg >= 1 && num_users >= 0 && num_items >= 1, RAG_EINVAL,
              "negative_sample: M=%lld n_neg=%d num_users=%lld num_items=%lld", (long long)M, n_neg, (long long)num_users,
              (long long)num_items);
  if (M == 0) return RAG_OK;
  RAG_REQUIRE(users && hist_rowptr && out, RAG_EINVAL, "negative_sample: null pointer");
  int64_t blocks = (M * n_neg + 255) / 256;
  const int64_t cap = (int64_t)rag::sm_count() * 8;
  if (blocks > cap) blocks = cap;
  rag::negative_sample_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(users, M, n_neg, hist_rowptr, hist_items,
                                                                                   num_users, num_items, seed, out);
  RAG_LAUNCH_OK("negative_sample_kernel");
  return RAG_OK;
}
