// Shared epilogue arithmetic of the tcgen05 TDNN kernels: accumulator -> bias -> ReLU -> BatchNorm
// offset -> 16-bit, two columns per instruction.
//
//   y = relu(acc + b) + o  ==  max(acc + (b + o), o)
//
// so with bo = b + o precomputed per column the fp32 part is ONE packed add (add.rn.f32x2, SASS
// FADD2), and because rounding to 16 bits is monotonic, max(round16(t), round16(o)) ==
// round16(max(t, o)): the ReLU moves behind the conversion and becomes one packed 16-bit max
// (HMNMX2).  Per column pair: FADD2 + F2FP + HMNMX2 instead of 4 FADD + 2 FMNMX + F2FP, which is
// what the short-K layers (two K blocks per tile) are bound by.  The BatchNorm scale is already
// folded into the weights and `b` (engine.cu), only its offset is left here.
//
// Parameter staging in shared memory, per N tile of kGemmBlockN columns:
//   float    bo[kGemmBlockN]       b + o
//   uint16_t oh[kGemmBlockN]       o rounded to the operand type
#ifndef CVX_KERNELS_EPILOGUE_CUH_
#define CVX_KERNELS_EPILOGUE_CUH_

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdint>

#include "kernels.h"

namespace cvx {
namespace epi {

template <bool kBF16>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  if constexpr (kBF16) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  } else {
    uint32_t r;  // saturating: a finite fp32 never becomes an fp16 infinity
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
  }
}

template <bool kBF16>
__device__ __forceinline__ uint32_t max16x2(uint32_t a, uint32_t b) {
  uint32_t r;
  if constexpr (kBF16) asm("max.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  else asm("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}

// One column's parameters into the staging area: kOutF32 kernels keep {b, o} as float2 (their
// epilogue stays in fp32), the 16-bit kernels the bo / oh form above.
template <bool kBF16, bool kOutF32>
__device__ __forceinline__ void store_post(uint8_t* smem_post, int j, float2 p) {
  if constexpr (kOutF32) {
    reinterpret_cast<float2*>(smem_post)[j] = p;
  } else {
    reinterpret_cast<float*>(smem_post)[j] = p.x + p.y;
    reinterpret_cast<uint16_t*>(smem_post + kGemmBlockN * 4)[j] = static_cast<uint16_t>(pack2<kBF16>(p.y, 0.f) & 0xffffu);
  }
}

__device__ __forceinline__ float2 load_post(const float2* __restrict__ post, int col, int n_pad) {
  return col < n_pad ? __ldg(&post[col]) : make_float2(0.f, 0.f);
}

// 8 consecutive columns, `col` (multiple of 8) counted from the start of the N tile: r[0..7] are this
// lane's accumulators; pk receives the 8 results as 16-bit pairs.
template <bool kBF16>
__device__ __forceinline__ void pack8(const uint32_t* r, const uint8_t* smem_post, int col, bool relu, uint32_t (&pk)[4]) {
  const float* bo = reinterpret_cast<const float*>(smem_post) + col;
  const float4 b0 = *reinterpret_cast<const float4*>(bo), b1 = *reinterpret_cast<const float4*>(bo + 4);
  const uint4 o = *reinterpret_cast<const uint4*>(smem_post + kGemmBlockN * 4 + col * 2);
  const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
  const uint32_t oo[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    uint64_t acc, add, sum;
    asm("mov.b64 %0, {%1, %2};" : "=l"(acc) : "r"(r[2 * e]), "r"(r[2 * e + 1]));
    asm("mov.b64 %0, {%1, %2};" : "=l"(add) : "f"(bb[2 * e]), "f"(bb[2 * e + 1]));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(sum) : "l"(acc), "l"(add));
    float t0, t1;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(t0), "=f"(t1) : "l"(sum));
    const uint32_t h = pack2<kBF16>(t0, t1);
    pk[e] = relu ? max16x2<kBF16>(h, oo[e]) : h;
  }
}

}  // namespace epi
}  // namespace cvx
#endif  // CVX_KERNELS_EPILOGUE_CUH_
