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
#include "ptx.cuh"

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


// One 128-byte-wide output chunk (64 16-bit or 32 fp32 columns) of this warp's 32 accumulator rows:
// TMEM -> bias / ReLU / BatchNorm offset -> this warp's 4 KB staging buffer (128B-swizzled like the
// tensor map) -> TMA store at (n0 + chunk columns, row0).  Shared by the CTA-pair and tap-halo
// kernels.  `t_row`: TMEM address of the warp's lane quarter and accumulator stage; `smem_post`: the
// tile's staged column parameters (store_post); n_cols: width of the N tile (multiple of 32 -- a
// tile of odd width in 32-column units ends in a half chunk that goes through the un-swizzled
// 32-column map).
template <bool kBF16, bool kOutF32>
__device__ __forceinline__ void store_chunk(uint32_t t_row, int c, int n_cols, int n0, int row0, int lane, uint8_t* buf,
                                            const uint8_t* smem_post, bool relu, const CUtensorMap* c_map,
                                            const CUtensorMap* c_map_half) {
  constexpr int kChunkCols = kOutF32 ? 32 : 64;
  if (ptx::elect_one_sync()) ptx::tma_store_wait_read<0>();  // this warp's previous store is done with `buf`
  __syncwarp();
  if constexpr (kOutF32) {
    const float2* post = reinterpret_cast<const float2*>(smem_post);
    uint32_t r[32];
    ptx::tmem_ld_32x32b_x32(t_row + c * 32, r);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; e += 2) {
        const float4 p = *reinterpret_cast<const float4*>(&post[c * 32 + j + e]);
        v[e] = __uint_as_float(r[j + e]) + p.x;
        v[e + 1] = __uint_as_float(r[j + e + 1]) + p.z;
        if (relu) {
          v[e] = fmaxf(v[e], 0.f);
          v[e + 1] = fmaxf(v[e + 1], 0.f);
        }
        v[e] += p.y;
        v[e + 1] += p.w;
      }
      const int unit = j >> 2;  // 16-byte unit (4 floats) inside the 128-byte row
      *reinterpret_cast<float4*>(buf + lane * 128 + ((unit ^ (lane & 7)) << 4)) = make_float4(v[0], v[1], v[2], v[3]);
    }
  } else {
    const bool half_chunk = (n_cols - c * 64) < 64;
    // both 32-column halves of the chunk are requested before the single wait
    uint32_t r0[32], r1[32];
    ptx::tmem_ld_32x32b_x32(t_row + c * 64, r0);
    if (!half_chunk) ptx::tmem_ld_32x32b_x32(t_row + c * 64 + 32, r1);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h == 1 && half_chunk) break;
      const uint32_t* r = h ? r1 : r0;
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint32_t pk[4];
        pack8<kBF16>(&r[j], smem_post, c * 64 + h * 32 + j, relu, pk);
        const int unit = h * 4 + (j >> 3);  // 16-byte unit inside the 128-byte row
        uint8_t* dst = half_chunk ? buf + lane * 64 + (unit << 4) : buf + lane * 128 + ((unit ^ (lane & 7)) << 4);
        *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
    }
    if (half_chunk) {
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (ptx::elect_one_sync()) {
        ptx::tma_store_2d(c_map_half, buf, n0 + c * 64, row0);
        ptx::tma_store_commit();
      }
      return;
    }
  }
  ptx::fence_proxy_async_smem();
  __syncwarp();
  if (ptx::elect_one_sync()) {
    ptx::tma_store_2d(c_map, buf, n0 + c * kChunkCols, row0);
    ptx::tma_store_commit();
  }
}

}  // namespace epi
}  // namespace cvx
#endif  // CVX_KERNELS_EPILOGUE_CUH_
