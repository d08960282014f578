// K2 -- ragged mean + stddev statistics pooling (HBM-bound streaming reduction).
//
// Replaces upstream Kaldi's StatisticsExtractionComponent (materialises
// [N, 1+2D] = count, x, x^2) + StatisticsPoolingComponent (AddRowRanges,
// DivRowsVec, AddMatMatElements, ApplyFloor(1e-10), ApplyPow(0.5)), configured by
// `stats-layer ... config=mean+stddev(0:1:1:10000)`
// (egs/sre/v2/local/nnet3/xvector/run_xvector_new.sh:106).
//
// Each chunk's valid rows [begin, end) are cut into slabs of <= 256 rows; one CTA
// reduces (slab, 32 x 16-byte column group): every activation is read exactly once
// with 128-bit loads, a warp reads 512 contiguous bytes of a row, the 8 warps of
// the CTA take interleaved rows.  Per column the CTA forms fp32 sums S1 = sum x,
// S2 = sum x^2 over its <= 256 rows and emits the slab's (mean, M2 = S2 - S1^2/n); a
// second tiny kernel merges a chunk's slabs with the Chan/Welford pairwise update
// (count, mean, M2) and writes [mean | sqrt(max(M2/n, floor))] -- cancellation is
// confined to one slab, where upstream sums x^2 over the whole chunk in fp32.
// No tensor cores: there is no contraction here.
#include "kernels.h"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace cvx {
namespace {

constexpr int kPoolWarps = 8;
constexpr int kPoolUnroll = 8;

// acc(s1, s2) += (x, x*x) for the 16-byte vector `u` of element type T.  fp16 / bf16
// use sm_100a's mixed-precision FHADD / FHFMA (add.rn.f32.f16, fma.rn.f32.f16): the
// 16-bit operand is consumed directly, accumulation is fp32 -- 2 instructions per
// element and no conversion, which is what keeps this kernel HBM- and not issue-bound.
template <typename T>
struct Vec16;
template <>
struct Vec16<__half> {
  static constexpr int kN = 8;
  __device__ __forceinline__ static void accumulate(const uint4& u, float (&s1)[8], float (&s2)[8]) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %4;\n\t"
          "add.rn.f32.f16 %0, lo, %0;\n\tadd.rn.f32.f16 %1, hi, %1;\n\t"
          "fma.rn.f32.f16 %2, lo, lo, %2;\n\tfma.rn.f32.f16 %3, hi, hi, %3;\n\t}"
          : "+f"(s1[2 * i]), "+f"(s1[2 * i + 1]), "+f"(s2[2 * i]), "+f"(s2[2 * i + 1])
          : "r"(w[i]));
  }
};
template <>
struct Vec16<__nv_bfloat16> {
  static constexpr int kN = 8;
  __device__ __forceinline__ static void accumulate(const uint4& u, float (&s1)[8], float (&s2)[8]) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %4;\n\t"
          "add.rn.f32.bf16 %0, lo, %0;\n\tadd.rn.f32.bf16 %1, hi, %1;\n\t"
          "fma.rn.f32.bf16 %2, lo, lo, %2;\n\tfma.rn.f32.bf16 %3, hi, hi, %3;\n\t}"
          : "+f"(s1[2 * i]), "+f"(s1[2 * i + 1]), "+f"(s2[2 * i]), "+f"(s2[2 * i + 1])
          : "r"(w[i]));
  }
};
template <>
struct Vec16<float> {
  static constexpr int kN = 4;
  __device__ __forceinline__ static void accumulate(const uint4& u, float (&s1)[4], float (&s2)[4]) {
    const float x[4] = {__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w)};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      s1[i] += x[i];
      s2[i] = fmaf(x[i], x[i], s2[i]);
    }
  }
};

__device__ __forceinline__ uint4 ld_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

template <typename T>
__global__ void __launch_bounds__(kPoolWarps * 32, 4) stats_pool_partial_kernel(const T* __restrict__ act, int ld,
                                                                                 const PoolSlab* __restrict__ slabs,
                                                                                 float* __restrict__ partials) {
  constexpr int kN = Vec16<T>::kN;
  constexpr int kCols = 32 * kN;  // columns per CTA
  __shared__ float red[kPoolWarps][2][kCols];

  const PoolSlab slab = slabs[blockIdx.x];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col0 = blockIdx.y * kCols;
  const int col = col0 + lane * kN;
  const bool active = col < ld;
  const T* base = act + slab.row_begin * ld + (active ? col : 0);

  // The fp32 validation path sums x - c (c = the slab's first row) so that it tracks the
  // fp64 oracle to ~1e-6 even for near-constant columns; fp16/bf16 inputs have 11/8-bit
  // mantissas, their squares are exact in fp32 and the plain sums are accurate enough.
  constexpr bool kShift = sizeof(T) == 4;
  float s1[kN], s2[kN], c[kN];
#pragma unroll
  for (int i = 0; i < kN; ++i) s1[i] = s2[i] = c[i] = 0.f;
  if (active) {
    if constexpr (kShift) {
      const uint4 u0 = ld_stream(base);
      c[0] = __uint_as_float(u0.x), c[1] = __uint_as_float(u0.y), c[2] = __uint_as_float(u0.z), c[3] = __uint_as_float(u0.w);
    }
    for (int r = warp; r < slab.nrows; r += kPoolWarps * kPoolUnroll) {
      uint4 v[kPoolUnroll];
#pragma unroll
      for (int u = 0; u < kPoolUnroll; ++u) {
        const int rr = r + u * kPoolWarps;
        v[u] = (rr < slab.nrows) ? ld_stream(base + static_cast<int64_t>(rr) * ld) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < kPoolUnroll; ++u) {
        if constexpr (kShift) {
          if (r + u * kPoolWarps < slab.nrows) {
            const float x[4] = {__uint_as_float(v[u].x), __uint_as_float(v[u].y), __uint_as_float(v[u].z), __uint_as_float(v[u].w)};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float d = x[i] - c[i];
              s1[i] += d;
              s2[i] = fmaf(d, d, s2[i]);
            }
          }
        } else {
          Vec16<T>::accumulate(v[u], s1, s2);  // zero rows add nothing
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kN; ++i) {
    red[warp][0][lane * kN + i] = s1[i];
    red[warp][1][lane * kN + i] = s2[i];
  }
  __syncthreads();
  for (int j = threadIdx.x; j < kCols; j += kPoolWarps * 32) {
    if (col0 + j >= ld) break;
    float t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int w = 0; w < kPoolWarps; ++w) {
      t1 += red[w][0][j];
      t2 += red[w][1][j];
    }
    const float mean = t1 / static_cast<float>(slab.nrows);  // of x (or of x - c)
    float cj = 0.f;
    if constexpr (kShift) cj = reinterpret_cast<const float*>(act)[slab.row_begin * ld + col0 + j];
    float* out = partials + static_cast<int64_t>(blockIdx.x) * 2 * ld + col0 + j;
    out[0] = cj + mean;
    out[ld] = fmaxf(t2 - t1 * mean, 0.f);  // M2 = sum (x - mean)^2 over this slab (<= 256 rows, fp32 sums)
  }
}

__global__ void __launch_bounds__(256) stats_pool_finalize_kernel(const float* __restrict__ partials, int ld,
                                                                   const PoolSlab* __restrict__ slabs,
                                                                   const int32_t* __restrict__ chunk_slab_off,
                                                                   int dim, int output_stddevs, float floor_,
                                                                   float* __restrict__ stats, int ld_stats) {
  const int chunk = blockIdx.x;
  const int j = blockIdx.y * blockDim.x + threadIdx.x;
  if (j >= dim) return;
  const int s0 = chunk_slab_off[chunk], s1 = chunk_slab_off[chunk + 1];
  float n = 0.f, mean = 0.f, m2 = 0.f;
  for (int s = s0; s < s1; ++s) {
    const float nb = static_cast<float>(slabs[s].nrows);
    const float mb = partials[static_cast<int64_t>(s) * 2 * ld + j];
    const float m2b = partials[static_cast<int64_t>(s) * 2 * ld + ld + j];
    const float tot = n + nb;
    const float delta = mb - mean;
    mean += delta * (nb / tot);
    m2 += m2b + delta * delta * (n * nb / tot);
    n = tot;
  }
  float* out = stats + static_cast<int64_t>(chunk) * ld_stats;
  out[j] = mean;
  if (output_stddevs) out[dim + j] = sqrtf(fmaxf(m2 / n, floor_));
}

// Finalize for the fused path: slots hold plain sums (sum x, sum x^2) over n <= 32 valid rows.
// CTA = (chunk, 256 / kLanes columns) x kLanes slot lanes: each lane walks a contiguous part of the
// chunk's slots, adding 8 slots (<= 256 rows, like a stand-alone slab; independent loads) before
// each Chan merge; the lanes are then merged in order through shared memory (deterministic).
// kLanes = 4 for long chunks (hundreds of slots), 1 for short segments (a handful of slots).
template <int kLanes>
__global__ void __launch_bounds__(256) stats_pool_finalize_fused_kernel(const float* __restrict__ partials, int total_slots,
                                                                         const int32_t* __restrict__ slot_begin,
                                                                         const int32_t* __restrict__ slot_n, int dim,
                                                                         int output_stddevs, float floor_,
                                                                         float* __restrict__ stats, int ld_stats) {
  constexpr int kCols = 256 / kLanes;
  __shared__ float sh[3][kLanes][kCols];
  const int chunk = blockIdx.x;
  const int cl = threadIdx.x % kCols, part = threadIdx.x / kCols;
  const int j = blockIdx.y * kCols + cl;
  const int s0 = slot_begin[chunk], s1 = slot_begin[chunk + 1];
  const int per = (s1 - s0 + kLanes - 1) / kLanes;
  const int a = s0 + part * per, b = min(s1, a + per);
  float n = 0.f, mean = 0.f, m2 = 0.f;
  if (j < dim) {
    // column j of slot s: block j / 64 of the blocked layout -- consecutive slots are 512 bytes apart
    const float* col = partials + (static_cast<int64_t>(j >> 6) * total_slots * 2) * 64 + (j & 63);
    for (int s = a; s < b; s += 8) {
      float nb = 0.f, sx = 0.f, sxx = 0.f;
      float vx[8], vxx[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const bool on = s + t < b;
        vx[t] = on ? col[static_cast<int64_t>(s + t) * 128] : 0.f;
        vxx[t] = on ? col[static_cast<int64_t>(s + t) * 128 + 64] : 0.f;
        nb += on ? static_cast<float>(slot_n[s + t]) : 0.f;
      }
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        sx += vx[t];
        sxx += vxx[t];
      }
      const float mb = sx / nb;
      const float m2b = fmaxf(sxx - sx * mb, 0.f);
      const float tot = n + nb;
      const float delta = mb - mean;
      mean += delta * (nb / tot);
      m2 += m2b + delta * delta * (n * nb / tot);
      n = tot;
    }
  }
  if constexpr (kLanes > 1) {
    sh[0][part][cl] = n;
    sh[1][part][cl] = mean;
    sh[2][part][cl] = m2;
    __syncthreads();
    if (part != 0) return;
    n = 0.f, mean = 0.f, m2 = 0.f;
#pragma unroll
    for (int p = 0; p < kLanes; ++p) {
      const float nb = sh[0][p][cl];
      if (nb > 0.f) {
        const float tot = n + nb;
        const float delta = sh[1][p][cl] - mean;
        mean += delta * (nb / tot);
        m2 += sh[2][p][cl] + delta * delta * (n * nb / tot);
        n = tot;
      }
    }
  }
  if (j < dim) {
    float* out = stats + static_cast<int64_t>(chunk) * ld_stats;
    out[j] = mean;
    if (output_stddevs) out[dim + j] = sqrtf(fmaxf(m2 / n, floor_));
  }
}

// One thread per 32-row group: binary search for the first chunk whose valid range reaches the
// group; that chunk's run and, if the next chunk starts inside the group too, a second run.  The
// host guarantees that no group holds three runs (engine.cu) and numbers a chunk's slots
// consecutively from slot_begin[c], one per group the chunk touches.
__global__ void __launch_bounds__(256) pool_meta_kernel(const int32_t* __restrict__ lo, const int32_t* __restrict__ hi,
                                                         const int32_t* __restrict__ slot_begin, int num_chunks,
                                                         int64_t num_groups, PoolGroupMeta* __restrict__ meta,
                                                         int32_t* __restrict__ slot_n) {
  const int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (g >= num_groups) return;
  const int64_t g0 = g * 32, g1 = g0 + 32;
  int a = 0, b = num_chunks;
  while (a < b) {
    const int mid = (a + b) >> 1;
    if (hi[mid] >= g0) b = mid;
    else a = mid + 1;
  }
  PoolGroupMeta m{0, 0, 0, 0, 0};
  if (a < num_chunks && lo[a] < g1) {
    const int32_t l = lo[a], h = hi[a];
    const int f = static_cast<int>(max(static_cast<int64_t>(l), g0) - g0);
    const int e = static_cast<int>(min(static_cast<int64_t>(h) + 1, g1) - g0);
    m.slot = slot_begin[a] + static_cast<int32_t>(g - l / 32);
    m.a0 = static_cast<uint8_t>(f);
    m.a1 = static_cast<uint8_t>(e);
    slot_n[m.slot] = e - f;
    if (a + 1 < num_chunks && lo[a + 1] < g1) {  // head of the next chunk: its first slot = m.slot + 1
      const int f2 = static_cast<int>(lo[a + 1] - g0);
      const int e2 = static_cast<int>(min(static_cast<int64_t>(hi[a + 1]) + 1, g1) - g0);
      m.b0 = static_cast<uint8_t>(f2);
      m.b1 = static_cast<uint8_t>(e2);
      slot_n[m.slot + 1] = e2 - f2;
    }
  }
  meta[g] = m;
}

}  // namespace

cudaError_t LaunchPoolMeta(const int32_t* lo, const int32_t* hi, const int32_t* slot_begin, int32_t num_chunks,
                           int64_t num_groups, PoolGroupMeta* meta, int32_t* slot_n, cudaStream_t stream) {
  if (num_groups <= 0) return cudaSuccess;
  pool_meta_kernel<<<static_cast<unsigned>((num_groups + 255) / 256), 256, 0, stream>>>(lo, hi, slot_begin, num_chunks,
                                                                                       num_groups, meta, slot_n);
  return cudaGetLastError();
}

cudaError_t LaunchStatsPoolFinalizeFused(const float* partials, int32_t ld, const int32_t* slot_begin,
                                         const int32_t* slot_n, int32_t num_chunks, int32_t total_slots, int32_t dim,
                                         int32_t output_stddevs, float variance_floor, float* stats,
                                         int32_t ld_stats, cudaStream_t stream) {
  if (num_chunks <= 0) return cudaSuccess;
  if (total_slots / num_chunks >= 32) {  // long chunks: split each chunk's slots over 4 lanes
    dim3 grid(num_chunks, (dim + 63) / 64);
    stats_pool_finalize_fused_kernel<4><<<grid, 256, 0, stream>>>(partials, total_slots, slot_begin, slot_n, dim, output_stddevs,
                                                                  variance_floor, stats, ld_stats);
  } else {
    dim3 grid(num_chunks, (dim + 255) / 256);
    stats_pool_finalize_fused_kernel<1><<<grid, 256, 0, stream>>>(partials, total_slots, slot_begin, slot_n, dim, output_stddevs,
                                                                  variance_floor, stats, ld_stats);
  }
  return cudaGetLastError();
}

cudaError_t LaunchStatsPoolPartial(const void* act, int32_t elem, int32_t ld, const PoolSlab* slabs,
                                   int32_t num_slabs, float* partials, cudaStream_t stream) {
  if (num_slabs <= 0) return cudaSuccess;
  const int cols = 32 * (16 / static_cast<int>(ElemSize(elem)));
  dim3 grid(num_slabs, (ld + cols - 1) / cols);
  if (elem == kF16)
    stats_pool_partial_kernel<__half><<<grid, kPoolWarps * 32, 0, stream>>>(static_cast<const __half*>(act), ld, slabs, partials);
  else if (elem == kBF16)
    stats_pool_partial_kernel<__nv_bfloat16><<<grid, kPoolWarps * 32, 0, stream>>>(static_cast<const __nv_bfloat16*>(act), ld, slabs, partials);
  else
    stats_pool_partial_kernel<float><<<grid, kPoolWarps * 32, 0, stream>>>(static_cast<const float*>(act), ld, slabs, partials);
  return cudaGetLastError();
}

cudaError_t LaunchStatsPoolFinalize(const float* partials, int32_t ld, const PoolSlab* slabs,
                                    const int32_t* chunk_slab_off, int32_t num_chunks, int32_t dim,
                                    int32_t output_stddevs, float variance_floor, float* stats,
                                    int32_t ld_stats, cudaStream_t stream) {
  if (num_chunks <= 0) return cudaSuccess;
  dim3 grid(num_chunks, (dim + 255) / 256);
  stats_pool_finalize_kernel<<<grid, 256, 0, stream>>>(partials, ld, slabs, chunk_slab_off, dim, output_stddevs,
                                                       variance_floor, stats, ld_stats);
  return cudaGetLastError();
}

}  // namespace cvx
