// Input staging: fp32 feature rows -> operand rows of the first TDNN layer(s).
//
// The first layer's splice (e.g. Append(input@-2..input@2), 5 x 23 = 115 columns,
// egs/sre/v2/local/nnet3/xvector/run_xvector_new.sh:96) is too narrow for 64-column
// K blocks per tap, so it is gathered once into a compact [rows, 128] operand
// (fp32 -> fp16/bf16 conversion fused); pad columns are written as zeros.
// One thread per (row, 8-column group); 16-byte stores, coalesced along the row.
#include "kernels.h"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <type_traits>

namespace cvx {
namespace {

template <typename T>
__device__ __forceinline__ T to_elem(float v);
template <>
__device__ __forceinline__ __half to_elem<__half>(float v) {
  return __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
}
template <>
__device__ __forceinline__ __nv_bfloat16 to_elem<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}
template <>
__device__ __forceinline__ float to_elem<float>(float v) {
  return v;
}

// two fp32 values -> one 32-bit word of two 16-bit operands (fp16: saturating, like to_elem)
template <typename T>
__device__ __forceinline__ uint32_t pack_pair(float a, float b) {
  if constexpr (sizeof(T) == 4) {
    return 0;  // unused for fp32
  } else if constexpr (std::is_same<T, __half>::value) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
  } else {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
}

// colmap[c] for output column c: bit 31 = valid, [28,31) stream, [16,28) time offset + 2048,
// [0,16) source column.  Built once per staging buffer on the host (engine.cu).
template <typename T>
__global__ void __launch_bounds__(256) stage_input_kernel(const __grid_constant__ StageArgs args) {
  constexpr int kVec = 16 / sizeof(T);
  extern __shared__ uint32_t s_colmap[];
  for (int c = threadIdx.x; c < args.ld_out; c += blockDim.x) s_colmap[c] = args.colmap[c];
  __syncthreads();
  const int groups = args.ld_out / kVec;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t row = idx / groups;
  if (row >= args.rows) return;
  const int c0 = static_cast<int>(idx % groups) * kVec;
  alignas(16) T vals[kVec];
#pragma unroll
  for (int e = 0; e < kVec; ++e) {
    const uint32_t m = s_colmap[c0 + e];
    float v = 0.f;
    if (m & 0x80000000u) {
      const int stream = (m >> 28) & 7;
      const int64_t sr = row + static_cast<int>((m >> 16) & 0xFFF) - 2048;
      if (sr >= args.src_row_min && sr < args.src_row_max)
        v = __ldg(args.feats[stream] + sr * args.feat_dim[stream] + (m & 0xFFFF));
    }
    vals[e] = to_elem<T>(v);
  }
  *reinterpret_cast<uint4*>(static_cast<T*>(args.out) + row * args.ld_out + c0) = *reinterpret_cast<uint4*>(vals);
}

// Tiled variant for the common case (one stream, small offsets): a CTA first copies the
// 128 + span source rows it needs into shared memory with fully coalesced loads (each feature
// row is read from global memory once instead of once per tap), then gathers from there.
constexpr int kStageTileRows = 128;

template <typename T>
__global__ void __launch_bounds__(256) stage_input_tiled_kernel(const __grid_constant__ StageArgs args) {
  constexpr int kVec = 16 / sizeof(T);
  extern __shared__ uint32_t s_mem[];
  // s_src[c]: index of output column c's source inside the tile's feature rows, relative to the
  // output row's own first element ((offset - off_min) * dim + column), or -1 for a padding column
  int* s_src = reinterpret_cast<int*>(s_mem);
  float* s_feat = reinterpret_cast<float*>(s_mem + args.ld_out);
  const int dim = args.feat_dim[args.tile_stream];
  const float* __restrict__ src = args.feats[args.tile_stream];
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * kStageTileRows;
  for (int c = threadIdx.x; c < args.ld_out; c += blockDim.x) {
    const uint32_t m = args.colmap[c];
    s_src[c] = (m & 0x80000000u) ? (static_cast<int>((m >> 16) & 0xFFF) - 2048 - args.off_min) * dim + static_cast<int>(m & 0xFFFF)
                                 : -1;
  }
  const int n_src = (kStageTileRows + args.off_max - args.off_min) * dim;
  const int64_t first = (r0 + args.off_min) * dim;
  const int64_t lo = args.src_row_min * dim, hi = args.src_row_max * dim;
  for (int i = threadIdx.x; i < n_src; i += blockDim.x) {
    const int64_t g = first + i;
    s_feat[i] = (g >= lo && g < hi) ? __ldg(src + g) : 0.f;
  }
  __syncthreads();
  // A warp covers 8 rows x 4 consecutive column groups: with 16-bit outputs (8 columns per group) the 32
  // shared-memory reads of one instruction then fall into 32 different banks for any odd feature
  // dimension (rows 23 floats apart give 8 distinct residues mod 8, groups add multiples of 8), and a
  // row's four 16-byte stores are one contiguous 64-byte segment.
  const int groups = args.ld_out / kVec;  // multiple of 8
  const int quads = groups >> 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int item = warp; item < (kStageTileRows / 8) * quads; item += blockDim.x >> 5) {
    const int row = (item / quads) * 8 + (lane >> 2);
    if (r0 + row >= args.rows) continue;
    const int c0 = ((item % quads) * 4 + (lane & 3)) * kVec;
    const float* base = s_feat + row * dim;
    alignas(16) int idx[kVec];
#pragma unroll
    for (int e = 0; e < kVec; e += 4) *reinterpret_cast<int4*>(&idx[e]) = *reinterpret_cast<const int4*>(&s_src[c0 + e]);
    float v[kVec];
#pragma unroll
    for (int e = 0; e < kVec; ++e) v[e] = idx[e] >= 0 ? base[idx[e]] : 0.f;
    uint4 packed;
    if constexpr (sizeof(T) == 4) {
      packed = make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
    } else {
      uint32_t w[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) w[e] = pack_pair<T>(v[2 * e], v[2 * e + 1]);  // one saturating convert per column pair
      packed = make_uint4(w[0], w[1], w[2], w[3]);
    }
    *reinterpret_cast<uint4*>(static_cast<T*>(args.out) + (r0 + row) * args.ld_out + c0) = packed;
  }
}

}  // namespace

cudaError_t LaunchStageInput(const StageArgs& args, cudaStream_t stream) {
  if (args.rows <= 0) return cudaSuccess;
  if (args.tile_stream >= 0) {
    const unsigned tgrid = static_cast<unsigned>((args.rows + kStageTileRows - 1) / kStageTileRows);
    const size_t tsmem = static_cast<size_t>(args.ld_out) * 4 +
                         static_cast<size_t>(kStageTileRows + args.off_max - args.off_min) * args.feat_dim[args.tile_stream] * 4;
    if (args.elem == kF16) stage_input_tiled_kernel<__half><<<tgrid, 256, tsmem, stream>>>(args);
    else if (args.elem == kBF16) stage_input_tiled_kernel<__nv_bfloat16><<<tgrid, 256, tsmem, stream>>>(args);
    else stage_input_tiled_kernel<float><<<tgrid, 256, tsmem, stream>>>(args);
    return cudaGetLastError();
  }
  const int vec = 16 / static_cast<int>(ElemSize(args.elem));
  const int64_t total = args.rows * (args.ld_out / vec);
  const unsigned grid = static_cast<unsigned>((total + 255) / 256);
  const size_t smem = static_cast<size_t>(args.ld_out) * sizeof(uint32_t);
  if (args.elem == kF16) stage_input_kernel<__half><<<grid, 256, smem, stream>>>(args);
  else if (args.elem == kBF16) stage_input_kernel<__nv_bfloat16><<<grid, 256, smem, stream>>>(args);
  else stage_input_kernel<float><<<grid, 256, smem, stream>>>(args);
  return cudaGetLastError();
}

}  // namespace cvx
