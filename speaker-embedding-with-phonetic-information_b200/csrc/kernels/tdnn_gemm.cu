// K1 -- TDNN frame layer as an implicit GEMM on tcgen05 tensor cores (sm_100a).
//
// Replaces, per layer, upstream Kaldi's descriptor splice copy (CopyRows/AddRows)
// + cublasSgemm (AffineComponent::Propagate) + ApplyFloor (ReLU) + BatchNorm
// scale/offset passes.  A spliced layer
// Append(Offset(x,o_1), ..., Offset(x,o_n)) is  sum_taps X[r+o_tap, :] . W_tap^T:
// the K loop walks (tap, 64-column block) pairs and the producer warp fetches the
// activation tile of each pair with TMA at row coordinate m0 + o_tap -- the splice
// is never materialised, and a two-branch Append just walks two source buffers.
//
//   warp 0      TMA producer (one lane): A tile 128x64 + W tile 256x64 per stage
//   warp 1      TMEM owner; one lane issues tcgen05.mma (128 x N x 16, fp32 accum in TMEM)
//   warps 2-9   epilogue (two warps per TMEM lane quarter, alternating 64-column
//               chunks): tcgen05.ld -> +bias, ReLU, +BN offset -> fp16/bf16 ->
//               swizzled smem -> TMA store, overlapped with the next tile's MMAs
//               through two 256-column accumulator stages.
// Padding is not computed: a tap whose source has d columns issues only
// ceil(d/16) K steps (KBlockDesc::k16), and the last N tile uses UMMA_N =
// roundup(out_dim,16) - n0; pad columns beyond that are never read by a consumer.
// Persistent CTAs (one per SM), static round-robin over (m_tile, n_tile).
#include "epilogue.cuh"
#include "kernels.h"
#include "ptx.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace cvx {
namespace {

constexpr int kStages = 4;
constexpr int kABytes = kGemmBlockM * kGemmBlockK * 2;  // 16 KB
constexpr int kBBytes = kGemmBlockN * kGemmBlockK * 2;  // 32 KB
constexpr int kCWarpBytes = 32 * 128;                   // 32 rows x 64 cols x 2 B per epilogue warp
constexpr int kSmemA = 0;
constexpr int kSmemB = kSmemA + kStages * kABytes;           // 65536
constexpr int kSmemC = kSmemB + kStages * kBBytes;           // 196608
constexpr int kSmemPost = kSmemC + 8 * kCWarpBytes;          // 229376 (one staging buffer per epilogue warp)
constexpr int kSmemBars = kSmemPost + kGemmBlockN * 8;       // 231424
constexpr int kNumBars = 2 * kStages + 4;
constexpr int kSmemTmemPtr = kSmemBars + kNumBars * 8;       // 231520
constexpr int kSmemTotal = kSmemTmemPtr + 32;                // 231552 <= 232448
constexpr int kTmemCols = 512;
constexpr int kThreads = 320;
constexpr int kEpilogueThreads = 256;

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
__global__ void __launch_bounds__(kThreads, 1) tdnn_gemm_kernel(const __grid_constant__ TdnnGemmArgs args) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smem_a = smem + kSmemA;
  uint8_t* smem_b = smem + kSmemB;
  uint8_t* smem_c = smem + kSmemC;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kSmemBars);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + kSmemTmemPtr);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x) >> 5, 0);
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    if (ptx::smem_u32(smem) & 1023u) __trap();  // SWIZZLE_128B tiles need 1024-byte alignment
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tmem_full_bar[i], 1);
      ptx::mbar_init(&tmem_empty_bar[i], kEpilogueThreads);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&args.b_map);
    ptx::prefetch_tensormap(&args.c_map);
    ptx::prefetch_tensormap(&args.a_maps[0]);
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_ptr, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int num_n_tiles = args.num_n_tiles;
  const int num_tiles = args.num_m_tiles * num_n_tiles;
  const int num_kb = args.num_kb;
  const int nstages = min(max(args.num_stages, 1), kStages);

  if (warp == 0) {
    // ------------------------------------------------------ TMA producer --
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / num_n_tiles) * kGemmBlockM;
        const int n0 = (tile % num_n_tiles) * args.tile_n;
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          if (ptx::elect_one_sync()) {
            ptx::mbar_arrive_expect_tx(&full_bar[stage], kABytes + args.b_box_rows * kGemmBlockK * 2);
            const KBlockDesc d = args.kb[kb];
            ptx::tma_load_2d(smem_a + stage * kABytes, &args.a_maps[d.src], &full_bar[stage], d.col, m0 + d.rowoff);
            ptx::tma_load_2d(smem_b + stage * kBBytes, &args.b_map, &full_bar[stage], kb * kGemmBlockK, n0);
          }
          __syncwarp();
          if (++stage == nstages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------- MMA issuer --
    {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int n0 = (tile % num_n_tiles) * args.tile_n;
        const int n_cols = min(args.tile_n, args.n_mma - n0);  // multiple of 16
        const uint32_t idesc = ptx::make_idesc_f16(kBF16 ? 1u : 0u, kGemmBlockM, n_cols);
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        ptx::mbar_wait(&tmem_empty_bar[as], aphase ^ 1);  // epilogue drained this accumulator
        ptx::tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + as * kGemmBlockN;
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tcgen05_fence_after();
          const uint64_t a_desc = ptx::make_umma_desc_sw128(ptx::smem_u32(smem_a + stage * kABytes));
          const uint64_t b_desc = ptx::make_umma_desc_sw128(ptx::smem_u32(smem_b + stage * kBBytes));
          const int k16 = args.kb[kb].k16;  // K steps that hold real columns (1..4)
          if (ptx::elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < kGemmBlockK / 16; ++k) {
              // +32 bytes per 16-element K step inside the 128-byte swizzle row (>>4 encoded)
              if (k < k16) ptx::umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
            }
            ptx::umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
            if (kb == num_kb - 1) ptx::umma_commit(&tmem_full_bar[as]);  // accumulator complete -> epilogue
          }
          __syncwarp();
          if (++stage == nstages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else {
    // ---------------------------------------------------------- epilogue --
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;  // which of the quarter's two warps: takes chunks c % 2 == half
    const int et = threadIdx.x - 64;
    uint8_t* buf = smem_c + (warp - 2) * kCWarpBytes;
    int it = 0;
    static_assert(kEpilogueThreads == kGemmBlockN, "one parameter column per epilogue thread");
    int staged_n0 = -1, pre_n0 = -1;
    float2 pre = make_float2(0.f, 0.f);
    const int tile_step = gridDim.x;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m0 = (tile / num_n_tiles) * kGemmBlockM;
      const int n0 = (tile % num_n_tiles) * args.tile_n;
      // whole 32-column units; columns beyond n_mma (a 16-column tail no MMA wrote) are clipped by the store maps
      const int n_cols = (min(args.tile_n, args.n_mma - n0) + 31) & ~31;
      const int n_chunks = (n_cols + 63) >> 6;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;

      // Per-column parameters of this N tile: restaged only when the tile column changes, and then
      // from registers loaded during the previous tile (the L2 latency is off the critical path).
      if (n0 != staged_n0) {
        asm volatile("bar.sync 1, %0;" ::"n"(kEpilogueThreads) : "memory");  // previous tile's params no longer read
        const float2 p = (pre_n0 == n0) ? pre : epi::load_post(args.post, n0 + et, args.n_pad);
        epi::store_post<kBF16, false>(smem + kSmemPost, et, p);
        asm volatile("bar.sync 1, %0;" ::"n"(kEpilogueThreads) : "memory");
        staged_n0 = n0;
      }
      if (tile + tile_step < num_tiles) {
        const int nn0 = ((tile + tile_step) % num_n_tiles) * args.tile_n;
        if (nn0 != n0) {
          pre = epi::load_post(args.post, nn0 + et, args.n_pad);
          pre_n0 = nn0;
        }
      }

      ptx::mbar_wait(&tmem_full_bar[as], aphase);
      ptx::tcgen05_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * kGemmBlockN;

      for (int c = half; c < n_chunks; c += 2) {
        if (ptx::elect_one_sync()) ptx::tma_store_wait_read<0>();  // this warp's previous store is done with `buf`
        __syncwarp();
        const bool half_chunk = (n_cols - c * 64) < 64;  // see tdnn_gemm_pair.cu
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (h == 1 && half_chunk) break;
          uint32_t r[32];
          ptx::tmem_ld_32x32b_x32(t_row + c * 64 + h * 32, r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint32_t pk[4];
            epi::pack8<kBF16>(&r[j], smem + kSmemPost, c * 64 + h * 32 + j, args.relu != 0, pk);
            const int unit = h * 4 + (j >> 3);  // 16-byte unit inside the 128-byte row
            uint8_t* dst = half_chunk ? buf + lane * 64 + (unit << 4) : buf + lane * 128 + ((unit ^ (lane & 7)) << 4);
            *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (ptx::elect_one_sync()) {
          ptx::tma_store_2d(half_chunk ? &args.c_map_half : &args.c_map, buf, n0 + c * 64, m0 + q * 32);
          ptx::tma_store_commit();
        }
      }
      ptx::tcgen05_fence_before();
      ptx::mbar_arrive(&tmem_empty_bar[as]);
    }
    if (ptx::elect_one_sync()) ptx::tma_store_wait<0>();
  }

  ptx::tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------- SIMT ----
constexpr int kSimtTile = 64;

__global__ void __launch_bounds__(256) tdnn_gemm_simt_kernel(const __grid_constant__ TdnnGemmSimtArgs args) {
  __shared__ float As[kSimtTile][kGemmBlockK + 1];
  __shared__ float Ws[kSimtTile][kGemmBlockK + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t m0 = static_cast<int64_t>(blockIdx.x) * kSimtTile;
  const int n0 = blockIdx.y * kSimtTile;
  float acc[4][4] = {};
  for (int kb = 0; kb < args.num_kb; ++kb) {
    const KBlockDesc d = args.kb[kb];
    const float* src = args.a_src[d.src];
    const int ld = args.a_ld[d.src];
    for (int i = threadIdx.x; i < kSimtTile * kGemmBlockK; i += 256) {
      const int r = i / kGemmBlockK, c = i % kGemmBlockK;
      const int64_t row = m0 + r + d.rowoff;
      As[r][c] = (row >= 0 && row < args.rows_cap) ? src[row * ld + d.col + c] : 0.f;
      const int n = n0 + r;
      Ws[r][c] = (n < args.n_pad) ? args.w[static_cast<int64_t>(n) * args.k_pad + kb * kGemmBlockK + c] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < kGemmBlockK; ++k) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a[i] = As[ty * 4 + i][k];
        w[i] = Ws[tx * 4 + i][k];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t row = m0 + ty * 4 + i;
    if (row >= args.rows) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= args.n_pad) continue;
      const float2 p = args.post[n];
      float v = acc[i][j] + p.x;
      if (args.relu) v = fmaxf(v, 0.f);
      args.out[row * args.ld_out + n] = v + p.y;
    }
  }
}

}  // namespace

size_t TdnnGemmSmemBytes() { return kSmemTotal; }

cudaError_t LaunchTdnnGemm(const TdnnGemmArgs& args, int num_sms, cudaStream_t stream) {
  const bool bf16 = args.ab_format == 1;
  auto kern = bf16 ? tdnn_gemm_kernel<true> : tdnn_gemm_kernel<false>;
  // per-device attribute; cheap enough to set on every launch
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal);
  if (e != cudaSuccess) return e;
  const int tiles = args.num_m_tiles * args.num_n_tiles;
  if (tiles <= 0) return cudaSuccess;
  const int grid = tiles < num_sms ? tiles : num_sms;
  kern<<<grid, kThreads, kSmemTotal, stream>>>(args);
  return cudaGetLastError();
}

cudaError_t LaunchTdnnGemmSimt(const TdnnGemmSimtArgs& args, cudaStream_t stream) {
  if (args.rows <= 0) return cudaSuccess;
  dim3 grid(static_cast<unsigned>((args.rows + kSimtTile - 1) / kSimtTile), (args.n_pad + kSimtTile - 1) / kSimtTile);
  tdnn_gemm_simt_kernel<<<grid, 256, 0, stream>>>(args);
  return cudaGetLastError();
}

}  // namespace cvx
