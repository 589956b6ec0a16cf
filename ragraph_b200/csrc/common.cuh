// Shared host/device helpers for libragraph_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/ragraph_b200.h"

namespace rag {

// thread-local error string + status helpers (api.cu)
int fail(int status, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
void count_launch(int n = 1);
int sm_count();
int max_smem_optin();

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

#define RAG_REQUIRE(cond, status, ...)                 \
  do {                                                 \
    if (!(cond)) return ::rag::fail(status, __VA_ARGS__); \
  } while (0)

#define RAG_LAUNCH_OK(what)                                          \
  do {                                                               \
    cudaError_t e__ = cudaGetLastError();                            \
    if (e__ != cudaSuccess) return ::rag::cuda_fail(e__, what);      \
    ::rag::count_launch();                                           \
  } while (0)

// ---- device helpers -------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// "a ranks before b" in the library's deterministic order: score desc, index asc.
__device__ __forceinline__ bool ranks_before(float sa, int64_t ia, float sb, int64_t ib) {
  return sa > sb || (sa == sb && ia < ib);
}

// Warp-cooperative insertion of (s, j) into a list sorted by ranks_before, length k, held in
// shared memory (vals[k], idx[k]).  The last element drops out.  All 32 lanes must call.
template <typename IdxT>
__device__ __forceinline__ void warp_sorted_insert(float* vals, IdxT* idx, int k, float s, IdxT j,
                                                   int lane) {
  int pos = 0;
  float v[(RAG_MAX_K + 31) / 32];
  IdxT ix[(RAG_MAX_K + 31) / 32];
  bool better[(RAG_MAX_K + 31) / 32];
#pragma unroll
  for (int t = 0; t < (RAG_MAX_K + 31) / 32; ++t) {
    int p = lane + 32 * t;
    better[t] = false;
    if (32 * t < k) {
      if (p < k) {
        v[t] = vals[p];
        ix[t] = idx[p];
        better[t] = ranks_before(v[t], (int64_t)ix[t], s, (int64_t)j);
      }
      pos += __popc(__ballot_sync(0xffffffffu, better[t]));
    }
  }
  __syncwarp();
#pragma unroll
  for (int t = 0; t < (RAG_MAX_K + 31) / 32; ++t) {
    int p = lane + 32 * t;
    if (32 * t < k && p < k && !better[t] && p + 1 < k) {
      vals[p + 1] = v[t];
      idx[p + 1] = ix[t];
    }
  }
  if (lane == 0 && pos < k) {
    vals[pos] = s;
    idx[pos] = j;
  }
  __syncwarp();
}

}  // namespace rag
