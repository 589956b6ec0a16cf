// K1: the pieces of F.normalize (SimilarityFunctions.py:8,11): inverse row norms with the
// 1e-12 clamp, and the normalised bf16 shadow of the key matrix for the tensor-core filter.
#include <cuda_bf16.h>
#include "common.cuh"

namespace rag {

__device__ __forceinline__ float row_sumsq(const float* __restrict__ x, int d, int lane) {
  float ss = 0.f;
  if ((d & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) & 15u) == 0)) {
    const float4* x4 = reinterpret_cast<const float4*>(x);
    for (int c = lane; c < (d >> 2); c += 32) {
      float4 v = __ldg(x4 + c);
      ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss);
      ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
    }
  } else {
    for (int c = lane; c < d; c += 32) { float v = __ldg(x + c); ss = fmaf(v, v, ss); }
  }
  return warp_sum(ss);
}

__global__ void __launch_bounds__(256) row_inv_norm_kernel(const float* __restrict__ x, int64_t rows,
                                                           int d, float eps, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = warp; r < rows; r += nwarps) {
    float ss = row_sumsq(x + r * d, d, lane);
    if (lane == 0) out[r] = 1.0f / fmaxf(sqrtf(ss), eps);
  }
}

__global__ void __launch_bounds__(256) rows_to_bf16_kernel(const float* __restrict__ x, int64_t rows,
                                                           int d, int normalize, float eps,
                                                           __nv_bfloat16* __restrict__ out, int d_pad) {
  const int lane = threadIdx.x & 31;
  int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = warp; r < rows; r += nwarps) {
    const float* xr = x + r * d;
    float inv = 1.0f;
    if (normalize) inv = 1.0f / fmaxf(sqrtf(row_sumsq(xr, d, lane)), eps);
    __nv_bfloat16* o = out + r * d_pad;
    // two elements per lane per step -> 4-byte stores
    for (int c = 2 * lane; c < d_pad; c += 64) {
      float a = c < d ? __ldg(xr + c) * inv : 0.f;
      float b = c + 1 < d ? __ldg(xr + c + 1) * inv : 0.f;
      *reinterpret_cast<__nv_bfloat162*>(o + c) = __floats2bfloat162_rn(a, b);
    }
  }
}

// F.normalize(x, p=2, dim=-1): out[r,:] = x[r,:] / max(||x[r,:]||, eps)  (library insert, ToyGraphBase.py:109)
__global__ void __launch_bounds__(256) rows_normalize_kernel(const float* __restrict__ x, int64_t rows, int d, float eps,
                                                             float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = warp; r < rows; r += nwarps) {
    const float* xr = x + r * d;
    const float nrm = fmaxf(sqrtf(row_sumsq(xr, d, lane)), eps);
    float* o = out + r * d;
    if ((d & 3) == 0 && ((reinterpret_cast<uintptr_t>(xr) | reinterpret_cast<uintptr_t>(o)) & 15u) == 0) {
      for (int c = lane; c < (d >> 2); c += 32) {
        float4 v = __ldg(reinterpret_cast<const float4*>(xr) + c);
        v.x /= nrm; v.y /= nrm; v.z /= nrm; v.w /= nrm;      // true division, like the reference
        reinterpret_cast<float4*>(o)[c] = v;
      }
    } else {
      for (int c = lane; c < d; c += 32) o[c] = __ldg(xr + c) / nrm;
    }
  }
}

// tf32 shadow: fp32 words whose low 13 mantissa bits are already zero (cvt.rna = round to nearest, ties away), so the
// tensor core's truncation of kind::tf32 operands is exact and the per-element error is 2^-11 instead of 2^-10
__global__ void __launch_bounds__(256) rows_to_tf32_kernel(const float* __restrict__ x, int64_t rows, int d,
                                                           int normalize, float eps, float* __restrict__ out,
                                                           int d_pad) {
  const int lane = threadIdx.x & 31;
  int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = warp; r < rows; r += nwarps) {
    const float* xr = x + r * d;
    float inv = 1.0f;
    if (normalize) inv = 1.0f / fmaxf(sqrtf(row_sumsq(xr, d, lane)), eps);
    float* o = out + r * d_pad;
    for (int c = lane; c < d_pad; c += 32) {
      const float v = c < d ? __ldg(xr + c) * inv : 0.f;
      uint32_t u;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
      o[c] = __uint_as_float(u);
    }
  }
}

}  // namespace rag

extern "C" int rag_rows_normalize_f32(const float* x, int64_t rows, int32_t d, float eps, float* out, rag_stream_t stream) {
  RAG_REQUIRE(rows >= 0 && d >= 1, RAG_EINVAL, "rows_normalize: rows=%lld d=%d", (long long)rows, d);
  if (rows == 0) return RAG_OK;
  RAG_REQUIRE(x && out, RAG_EINVAL, "rows_normalize: null pointer");
  int64_t blocks = (rows + 7) / 8;
  int64_t cap = (int64_t)rag::sm_count() * 16;
  if (blocks > cap) blocks = cap;
  rag::rows_normalize_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, rows, d, eps, out);
  RAG_LAUNCH_OK("rows_normalize_kernel");
  return RAG_OK;
}

extern "C" int rag_rows_to_tf32(const float* x, int64_t rows, int32_t d, int32_t normalize, float eps, float* out,
                                int32_t d_pad, rag_stream_t stream) {
  RAG_REQUIRE(rows >= 0 && d >= 1 && d_pad >= d && d_pad % 32 == 0, RAG_EINVAL,
              "rows_to_tf32: rows=%lld d=%d d_pad=%d (d_pad must be a multiple of 32, >= d)", (long long)rows, d, d_pad);
  if (rows == 0) return RAG_OK;
  RAG_REQUIRE(x && out, RAG_EINVAL, "rows_to_tf32: null pointer");
  RAG_REQUIRE(rag::aligned16(out), RAG_EALIGN, "rows_to_tf32: out not 16-byte aligned");
  int64_t blocks = (rows + 7) / 8;
  int64_t cap = (int64_t)rag::sm_count() * 16;
  if (blocks > cap) blocks = cap;
  rag::rows_to_tf32_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, rows, d, normalize, eps, out, d_pad);
  RAG_LAUNCH_OK("rows_to_tf32_kernel");
  return RAG_OK;
}

extern "C" int rag_row_inv_norm_f32(const float* x, int64_t rows, int32_t d, float eps, float* out,
                                    rag_stream_t stream) {
  RAG_REQUIRE(rows >= 0 && d >= 1, RAG_EINVAL, "row_inv_norm: rows=%lld d=%d", (long long)rows, d);
  if (rows == 0) return RAG_OK;
  RAG_REQUIRE(x && out, RAG_EINVAL, "row_inv_norm: null pointer");
  int64_t blocks = (rows + 7) / 8;
  int64_t cap = (int64_t)rag::sm_count() * 16;
  if (blocks > cap) blocks = cap;
  rag::row_inv_norm_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, rows, d, eps, out);
  RAG_LAUNCH_OK("row_inv_norm_kernel");
  return RAG_OK;
}

extern "C" int rag_rows_to_bf16(const float* x, int64_t rows, int32_t d, int32_t normalize, float eps,
                                uint16_t* out, int32_t d_pad, rag_stream_t stream) {
  RAG_REQUIRE(rows >= 0 && d >= 1 && d_pad >= d && d_pad % 64 == 0, RAG_EINVAL,
              "rows_to_bf16: rows=%lld d=%d d_pad=%d (d_pad must be a multiple of 64, >= d)",
              (long long)rows, d, d_pad);
  if (rows == 0) return RAG_OK;
  RAG_REQUIRE(x && out, RAG_EINVAL, "rows_to_bf16: null pointer");
  RAG_REQUIRE(rag::aligned16(out), RAG_EALIGN, "rows_to_bf16: out not 16-byte aligned");
  int64_t blocks = (rows + 7) / 8;
  int64_t cap = (int64_t)rag::sm_count() * 16;
  if (blocks > cap) blocks = cap;
  rag::rows_to_bf16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      x, rows, d, normalize, eps, reinterpret_cast<__nv_bfloat16*>(out), d_pad);
  RAG_LAUNCH_OK("rows_to_bf16_kernel");
  return RAG_OK;
}
