// Device helpers shared by the talker kernels and the frame megakernel: warp reductions, streaming loads, and the
// dequant-in-registers lane primitives of the GEMV (MLX affine 4/8-bit, bf16, f16, f32 weight rows).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "kernels.h"

namespace q3 {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float load_as_f32(const void* p, size_t i, int dt) {
  if (dt == Q3TTS_F32) return reinterpret_cast<const float*>(p)[i];
  if (dt == Q3TTS_F16) return __half2float(reinterpret_cast<const __half*>(p)[i]);
  return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {  // weights are read once: keep them out of L1
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }


enum WFmt { W_Q4 = 0, W_Q8 = 1, W_BF16 = 2, W_F16 = 3, W_F32 = 4 };
template <int FMT> struct FmtTraits;
template <> struct FmtTraits<W_Q4> { static constexpr int VPL = 32; };   // values per lane per 16-byte load
template <> struct FmtTraits<W_Q8> { static constexpr int VPL = 16; };
template <> struct FmtTraits<W_BF16> { static constexpr int VPL = 8; };
template <> struct FmtTraits<W_F16> { static constexpr int VPL = 8; };
template <> struct FmtTraits<W_F32> { static constexpr int VPL = 4; };

// Integer code -> float without I2F (quarter-rate pipe): OR the code into the mantissa of 2^23, subtract 2^23.
__device__ __forceinline__ float code_to_f32(uint32_t code) { return __uint_as_float(0x4B000000u | code) - 8388608.0f; }

// Expand this lane's 16-byte weight load into VPL floats (integer codes for the packed formats; the group scale / bias are
// applied to the finished dot product).  Done ONCE per weight load and reused for every activation row of the M tile.
template <int FMT>
__device__ __forceinline__ void lane_expand(const uint4& w, float (&o)[FmtTraits<FMT>::VPL]) {
  const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
  if constexpr (FMT == W_Q4) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
      for (int n = 0; n < 8; ++n) o[i * 8 + n] = code_to_f32((ww[i] >> (4 * n)) & 0xF);
    }
  } else if constexpr (FMT == W_Q8) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      // __byte_perm places byte k of the word into byte 0 of 0x4B0000xx in one PRMT
      o[i * 4 + 0] = __uint_as_float(__byte_perm(ww[i], 0x4B000000u, 0x7440)) - 8388608.0f;
      o[i * 4 + 1] = __uint_as_float(__byte_perm(ww[i], 0x4B000000u, 0x7441)) - 8388608.0f;
      o[i * 4 + 2] = __uint_as_float(__byte_perm(ww[i], 0x4B000000u, 0x7442)) - 8388608.0f;
      o[i * 4 + 3] = __uint_as_float(__byte_perm(ww[i], 0x4B000000u, 0x7443)) - 8388608.0f;
    }
  } else if constexpr (FMT == W_BF16) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      o[2 * i] = __uint_as_float(ww[i] << 16);
      o[2 * i + 1] = __uint_as_float(ww[i] & 0xFFFF0000u);
    }
  } else if constexpr (FMT == W_F16) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&ww[i]));
      o[2 * i] = a.x;
      o[2 * i + 1] = a.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = __uint_as_float(ww[i]);
  }
}
// 4-bit codes as m = 1 + q/16 (the nibble shifted into the top mantissa bits of 1.0f: one shift + one LOP3 per value, no
// FADD).  sum(m_i x_i) = sum(x_i) + sum(q_i x_i) / 16; the caller removes the sum(x_i) term with the staged per-lane sums.
__device__ __forceinline__ void lane_expand_q4_unit(const uint4& w, float (&o)[32]) {
  const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const uint32_t sh = (19 - 4 * n) >= 0 ? (ww[i] << ((19 - 4 * n) & 31)) : (ww[i] >> ((4 * n - 19) & 31));
      o[i * 8 + n] = __uint_as_float((sh & 0x00780000u) | 0x3F800000u);
    }
  }
}
// dot of the expanded weights with this lane's VPL activations (VPL/4 float4 from lane-major smem, stride 32 float4)
template <int FMT>
__device__ __forceinline__ float lane_dot(const float (&w)[FmtTraits<FMT>::VPL], const float4* __restrict__ xs) {
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < FmtTraits<FMT>::VPL / 4; ++j) {
    const float4 x = xs[j * 32];
    acc = fmaf(w[4 * j], x.x, acc);
    acc = fmaf(w[4 * j + 1], x.y, acc);
    acc = fmaf(w[4 * j + 2], x.z, acc);
    acc = fmaf(w[4 * j + 3], x.w, acc);
  }
  return acc;
}


}  // namespace q3
