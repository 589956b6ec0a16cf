// K1: the pieces of F.normalize (SimilarityFunctions.py:8,11): inverse row norms with the
// 1e-12 clamp, and the normalised bf16 shadow of the key matrix for the tensor-core filter.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
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

// 16-bit shadow (bf16 or fp16) of a row matrix for the tensor-core filter, with the quantities the exactness
// certificate needs: besides out[r, :] = rn16(x[r, :] * inv) it can emit
//   inv_out[r]  = 1 / max(||x[r]||, eps)                      (the query prologue wants it anyway),
//   err_rows[r] = || rn16(xhat_r) - xhat_r ||_2               (xhat = the fp32 normalised row the shadow was rounded from),
//   *err_max    = max_r err_rows[r]                            (atomicMax on the bit pattern: non-negative floats order like
//                                                               unsigned integers; the caller zero-initialises it),
// and it clears `n_zero` 32-bit words (device-side counters of the refine passes) so the caller needs no memset launch.
// The rounding-error NORMS give a proven score bound by Cauchy-Schwarz: for a query row q and a key row k,
//   | qh . kh - qhat . khat |  <=  ||qh - qhat|| * ||kh|| + ||qhat|| * ||kh - khat||  <=  (err_q + err_k) * (1 + 2^-7),
// about half of the element-wise worst case (2 u, u = 2^-8 bf16 / 2^-11 fp16), and it holds whatever the tensor core does
// with subnormal fp16 inputs because those are flushed HERE, before the error is measured.
template <bool F16>
__global__ void __launch_bounds__(256) rows_to_16_kernel(const float* __restrict__ x, int64_t rows, int d, int normalize,
                                                         float eps, uint16_t* __restrict__ out, int d_pad,
                                                         float* __restrict__ inv_out, float* __restrict__ err_rows,
                                                         float* __restrict__ err_max, uint32_t* __restrict__ zero_words,
                                                         int64_t n_zero) {
  const int lane = threadIdx.x & 31;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = tid; i < n_zero; i += nthr) zero_words[i] = 0u;
  int64_t warp = tid >> 5;
  const int64_t nwarps = nthr >> 5;
  float emax = 0.f;
  for (int64_t r = warp; r < rows; r += nwarps) {
    const float* xr = x + r * d;
    float inv = 1.0f;
    if (normalize || inv_out) {
      const float i2 = 1.0f / fmaxf(sqrtf(row_sumsq(xr, d, lane)), eps);
      if (inv_out && lane == 0) inv_out[r] = i2;
      if (normalize) inv = i2;
    }
    uint16_t* o = out + r * d_pad;
    float e2 = 0.f;
    // two elements per lane per step -> 4-byte stores
    for (int c = 2 * lane; c < d_pad; c += 64) {
      const float a = c < d ? __ldg(xr + c) * inv : 0.f;
      const float b = c + 1 < d ? __ldg(xr + c + 1) * inv : 0.f;
      float ra, rb;
      uint32_t packed;
      if constexpr (F16) {
        __half ha = __float2half_rn(a), hb = __float2half_rn(b);
        // flush subnormal halves (|h| < 2^-14): the bound must not depend on how the MMA treats them
        if (fabsf(__half2float(ha)) < 6.103515625e-05f) ha = __float2half_rn(0.f);
        if (fabsf(__half2float(hb)) < 6.103515625e-05f) hb = __float2half_rn(0.f);
        ra = __half2float(ha); rb = __half2float(hb);
        packed = (uint32_t)__half_as_ushort(ha) | ((uint32_t)__half_as_ushort(hb) << 16);
      } else {
        const __nv_bfloat16 ha = __float2bfloat16_rn(a), hb = __float2bfloat16_rn(b);
        ra = __bfloat162float(ha); rb = __bfloat162float(hb);
        packed = (uint32_t)__bfloat16_as_ushort(ha) | ((uint32_t)__bfloat16_as_ushort(hb) << 16);
      }
      *reinterpret_cast<uint32_t*>(o + c) = packed;
      const float da = ra - a, db = rb - b;                   // exact in fp32 (Sterbenz / small exponent gap)
      e2 = fmaf(da, da, e2); e2 = fmaf(db, db, e2);
    }
    if (err_rows || err_max) {
      const float e = sqrtf(warp_sum(e2)) * 1.0001f;          // the fp32 evaluation of the norm itself
      if (err_rows && lane == 0) err_rows[r] = e;
      emax = fmaxf(emax, e);
    }
  }
  if (err_max && lane == 0 && emax > 0.f) atomicMax(reinterpret_cast<unsigned int*>(err_max), __float_as_uint(emax));
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

namespace rag {
int rows_to_16_launch(const float* x, int64_t rows, int d, int fmt, int normalize, float eps, uint16_t* out, int d_pad,
                      float* inv_out, float* err_rows, float* err_max, uint32_t* zero_words, int64_t n_zero, cudaStream_t s) {
  int64_t blocks = (rows + 7) / 8;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (fmt == RAG_FMT_F16)
    rows_to_16_kernel<true><<<(unsigned)blocks, 256, 0, s>>>(x, rows, d, normalize, eps, out, d_pad, inv_out, err_rows, err_max,
                                                             zero_words, n_zero);
  else
    rows_to_16_kernel<false><<<(unsigned)blocks, 256, 0, s>>>(x, rows, d, normalize, eps, out, d_pad, inv_out, err_rows, err_max,
                                                              zero_words, n_zero);
  RAG_LAUNCH_OK("rows_to_16_kernel");
  return RAG_OK;
}
}  // namespace rag

extern "C" int rag_rows_to_shadow16(const float* x, int64_t rows, int32_t d, int32_t fmt, int32_t normalize, float eps,
                                    uint16_t* out, int32_t d_pad, float* err_rows, float* err_max, rag_stream_t stream) {
  RAG_REQUIRE(rows >= 0 && d >= 1 && d_pad >= d && d_pad % 64 == 0, RAG_EINVAL,
              "rows_to_shadow16: rows=%lld d=%d d_pad=%d (d_pad must be a multiple of 64, >= d)", (long long)rows, d, d_pad);
  RAG_REQUIRE(fmt == RAG_FMT_BF16 || fmt == RAG_FMT_F16, RAG_EINVAL, "rows_to_shadow16: fmt=%d", fmt);
  if (rows == 0) return RAG_OK;
  RAG_REQUIRE(x && out, RAG_EINVAL, "rows_to_shadow16: null pointer");
  RAG_REQUIRE(rag::aligned16(out), RAG_EALIGN, "rows_to_shadow16: out not 16-byte aligned");
  return rag::rows_to_16_launch(x, rows, d, fmt, normalize, eps, out, d_pad, nullptr, err_rows, err_max, nullptr, 0,
                                (cudaStream_t)stream);
}

extern "C" int rag_rows_to_bf16(const float* x, int64_t rows, int32_t d, int32_t normalize, float eps,
                                uint16_t* out, int32_t d_pad, rag_stream_t stream) {
  return rag_rows_to_shadow16(x, rows, d, RAG_FMT_BF16, normalize, eps, out, d_pad, nullptr, nullptr, stream);
}
