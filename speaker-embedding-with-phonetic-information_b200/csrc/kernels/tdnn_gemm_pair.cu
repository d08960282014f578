// K1 (CTA-pair variant) -- TDNN frame layer as an implicit GEMM with cta_group::2.
//
// Same contraction and epilogue as tdnn_gemm.cu, but two CTAs on the two SMs of a TPC
// form one 256 x N tile: each CTA stages its own 128 activation rows and only HALF of the
// weight tile (N/2 rows) per K block; one thread of the leader CTA issues
// tcgen05.mma.cta_group::2 (M = 256), which reads both CTAs' shared memory and writes a
// 128 x N accumulator into each CTA's TMEM.  Per K block a CTA therefore moves 32 KB
// (16 A + 16 W) instead of 48 KB for the same 512 MMA cycles, which is what the 1-CTA
// kernel was bound by (ncu: ~590 cycles per K block at 48 KB regardless of MMA work),
// and the freed shared memory deepens the ring from 4 to 6 stages.
//
//   both CTAs   warp 0: TMA producer -> own smem, completion counted on the LEADER's
//                       full barrier (cp.async.bulk.tensor ... .cta_group::2)
//   leader      warp 1: waits full, issues the MMAs, tcgen05.commit multicast frees the
//                       stage in both CTAs / publishes the accumulator to both epilogues
//   both CTAs   warps 2-9: epilogue of their own 128 rows; lane 0 of each warp arrives on
//                       the leader's tmem_empty barrier.
#include "epilogue.cuh"
#include "kernels.h"
#include "ptx.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace cvx {
namespace {

constexpr int kABytes = kGemmBlockM * kGemmBlockK * 2;        // 16 KB: 128 rows of this CTA
constexpr int kBBytes = (kGemmBlockN / 2) * kGemmBlockK * 2;  // 16 KB: this CTA's half of W
constexpr int kCWarpBytes = 32 * 128;
constexpr int kTmemCols = 512;
// Epilogue warps: 8 (two per TMEM lane quarter, 64-column chunks) for the ordinary storing variant.
// kPool: the fused-pooling epilogue is latency-bound (tcgen05.ld -> convert -> smem transpose ->
// dependent FHADD/FHFMA chains; ncu r2: ~5 600 cycles per 128 x 256 tile with two warps per
// scheduler, issue slots 40 % used), so it runs 16 warps (four per quarter) on 32-column units with
// half the per-warp staging buffer each.
// kWide: the first TDNN layers (K <= 128: two K blocks, weights resident) are all epilogue -- 1 024
// MMA cycles per tile against ~2 900 cycles of TMEM -> convert -> staging -> TMA store, again
// latency-bound with two warps per scheduler (issue slots 32 % used, and slower in the sustained
// bench than under ncu exactly by the SM clock ratio).  They run 16 epilogue warps, one 64-column chunk
// each, with 4 KB staging buffers paid for by a 4-stage ring (they only ever have 2 K blocks in flight).
template <bool kPool, bool kWide>
struct PairCfg {
  static_assert(!(kPool && kWide), "one or the other");
  static constexpr int kStages = kWide ? 4 : 6;
  static constexpr int kEpiWarps = (kPool || kWide) ? 16 : 8;
  static constexpr int kEpiThreads = 32 * kEpiWarps;
  static constexpr int kThreads = 64 + kEpiThreads;
  static constexpr int kParts = kEpiWarps / 4;                  // warps sharing one TMEM lane quarter
  static constexpr int kWarpBuf = kPool ? kCWarpBytes / 2 : kCWarpBytes;  // bytes of smem_c per warp
  static constexpr int kSmemA = 0;
  static constexpr int kSmemB = kSmemA + kStages * kABytes;
  static constexpr int kSmemC = kSmemB + kStages * kBBytes;
  static constexpr int kSmemPost = kSmemC + kEpiWarps * kWarpBuf;
  static constexpr int kSmemBars = kSmemPost + kGemmBlockN * 8;
  static constexpr int kNumBars = 2 * kStages + 5;
  static constexpr int kSmemTmemPtr = kSmemBars + kNumBars * 8;
  static constexpr int kSmemTotal = kSmemTmemPtr + 32;
  static_assert(kSmemB % 1024 == 0 && kSmemC % 1024 == 0, "swizzle alignment");
  static_assert(kSmemTotal <= 232448, "shared memory budget");
};

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

// (sx, sxx) += (x, x * x) for a 16-bit x: sm_100a's mixed-precision FHADD / FHFMA take the 16-bit
// operand directly and accumulate in fp32.
template <bool kBF16>
__device__ __forceinline__ void accum16(uint16_t x, float& sx, float& sxx) {
  if constexpr (kBF16)
    asm("add.rn.f32.bf16 %0, %2, %0;\n\tfma.rn.f32.bf16 %1, %2, %2, %1;" : "+f"(sx), "+f"(sxx) : "h"(x));
  else
    asm("add.rn.f32.f16 %0, %2, %0;\n\tfma.rn.f32.f16 %1, %2, %2, %1;" : "+f"(sx), "+f"(sxx) : "h"(x));
}

// kOutF32: the epilogue writes fp32 (32-column chunks) instead of the operand type -- used for
// the segment-level layers, whose fp32 inputs/weights are fed as fp16 hi/lo splits.
// kPool:   fused statistics pooling -- the epilogue of the pooled layer rounds its outputs to the
// operand type (as if stored) and reduces sum x / sum x^2 over the valid rows of each 32-row
// group (PoolGroupMeta) into per-slot partials through a swizzled smem transpose; the
// [rows, 1500] activation is never written.
template <bool kBF16, bool kOutF32, bool kPool, bool kWide = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PairCfg<kPool, kWide>::kThreads, 1)
    tdnn_gemm_pair_kernel(const __grid_constant__ TdnnGemmArgs args) {
  using Cfg = PairCfg<kPool, kWide>;
  constexpr int kEpilogueThreads = Cfg::kEpiThreads;
  constexpr int kEpilogueWarps = Cfg::kEpiWarps;
  constexpr int kStages = Cfg::kStages, kSmemA = Cfg::kSmemA, kSmemB = Cfg::kSmemB, kSmemC = Cfg::kSmemC,
                kSmemPost = Cfg::kSmemPost, kSmemBars = Cfg::kSmemBars, kSmemTmemPtr = Cfg::kSmemTmemPtr;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smem_a = smem + kSmemA;
  uint8_t* smem_b = smem + kSmemB;
  uint8_t* smem_c = smem + kSmemC;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kSmemBars);  // used in the leader
  uint64_t* empty_bar = full_bar + kStages;                            // per CTA
  uint64_t* tmem_full_bar = empty_bar + kStages;                       // per CTA
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;                        // used in the leader
  uint64_t* b_res_bar = tmem_empty_bar + 2;                            // used in the leader (resident weights)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + kSmemTmemPtr);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x) >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();

  if (threadIdx.x == 0) {
    if (ptx::smem_u32(smem) & 1023u) __trap();
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(&full_bar[i], 2);   // one arrive.expect_tx per CTA of the pair
      ptx::mbar_init(&empty_bar[i], 1);  // the leader's multicast commit
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tmem_full_bar[i], 1);
      ptx::mbar_init(&tmem_empty_bar[i], 2 * kEpilogueWarps);
    }
    ptx::mbar_init(b_res_bar, 2);
    ptx::fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&args.b_map_pair);
    ptx::prefetch_tensormap(&args.c_map);
    ptx::prefetch_tensormap(&args.a_maps[0]);
  }
  if (warp == 1) {
    ptx::tmem_alloc_pair(tmem_ptr, kTmemCols);
    ptx::tmem_relinquish_pair();
  }
  ptx::tcgen05_fence_before();
  ptx::cluster_sync_all();  // barriers of both CTAs are initialised before any remote arrive
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int num_n_tiles = args.num_n_tiles;
  const int num_tiles = ((args.num_m_tiles + 1) >> 1) * num_n_tiles;  // 256-row tiles
  const int num_kb = args.num_kb;
  const int nstages = min(max(args.num_stages, 1), kStages);
  const bool resident = args.resident_b != 0;
  const int first_tile = blockIdx.x >> 1;
  const int tile_step = gridDim.x >> 1;

  if (warp == 0) {
    // ------------------------------------------------------ TMA producer --
    {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t b_bytes = args.b_box_rows_pair * kGemmBlockK * 2;
      const uint32_t stage_bytes = kABytes + (resident ? 0u : b_bytes);
      if (resident && first_tile < num_tiles) {
        // this pair's tile column never changes (tile_step is a multiple of num_n_tiles): W slot kb holds
        // K block kb for the whole launch
        const int n0 = (first_tile % num_n_tiles) * args.tile_n;
        const int n_cols = min(args.tile_n, args.n_mma - n0);
        const int nb0 = n0 + static_cast<int>(rank) * (n_cols >> 1);
        if (ptx::elect_one_sync()) {
          ptx::mbar_arrive_expect_tx_cluster(b_res_bar, b_bytes * num_kb, 0);
          for (int kb = 0; kb < num_kb; ++kb)
            ptx::tma_load_2d_pair(smem_b + kb * kBBytes, &args.b_map_pair, b_res_bar, 0, kb * kGemmBlockK, nb0);
        }
        __syncwarp();
      }
      for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
        const int m0 = (tile / num_n_tiles) * (2 * kGemmBlockM) + static_cast<int>(rank) * kGemmBlockM;
        const int n0 = (tile % num_n_tiles) * args.tile_n;
        const int n_cols = min(args.tile_n, args.n_mma - n0);
        const int nb0 = n0 + static_cast<int>(rank) * (n_cols >> 1);  // this CTA's half of the W tile
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          if (ptx::elect_one_sync()) {
            ptx::mbar_arrive_expect_tx_cluster(&full_bar[stage], stage_bytes, 0);
            const KBlockDesc d = args.kb[kb];
            ptx::tma_load_2d_pair(smem_a + stage * kABytes, &args.a_maps[d.src], &full_bar[stage], 0, d.col,
                                  m0 + d.rowoff);
            if (!resident)
              ptx::tma_load_2d_pair(smem_b + stage * kBBytes, &args.b_map_pair, &full_bar[stage], 0,
                                    kb * kGemmBlockK, nb0);
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
    // ------------------------------------------- MMA issuer (leader CTA) --
    if (rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      if (resident && first_tile < num_tiles) ptx::mbar_wait(b_res_bar, 0);  // both CTAs' weight half-tiles have landed
      for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++it) {
        const int n0 = (tile % num_n_tiles) * args.tile_n;
        const int n_cols = min(args.tile_n, args.n_mma - n0);  // multiple of 16
        const uint32_t idesc = ptx::make_idesc_f16(kBF16 ? 1u : 0u, 2 * kGemmBlockM, n_cols);
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        ptx::mbar_wait(&tmem_empty_bar[as], aphase ^ 1);  // both epilogues drained this accumulator
        ptx::tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + as * kGemmBlockN;
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);  // both CTAs' A and W halves have landed
          ptx::tcgen05_fence_after();
          const uint64_t a_desc = ptx::make_umma_desc_sw128(ptx::smem_u32(smem_a + stage * kABytes));
          const uint64_t b_desc = ptx::make_umma_desc_sw128(ptx::smem_u32(smem_b + (resident ? kb : stage) * kBBytes));
          const int k16 = args.kb[kb].k16;
          if (ptx::elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < kGemmBlockK / 16; ++k) {
              if (k < k16) ptx::umma_f16_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
            }
            ptx::umma_commit_pair_mc(&empty_bar[stage], 0x3);  // frees the stage in both CTAs
            if (kb == num_kb - 1) ptx::umma_commit_pair_mc(&tmem_full_bar[as], 0x3);  // accumulators complete
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
    const int half = (warp - 2) >> 2;  // which of the quarter's warps: takes chunks / units c % kParts == half
    const int et = threadIdx.x - 64;
    uint8_t* buf = smem_c + (warp - 2) * Cfg::kWarpBuf;
    int it = 0;
    static_assert(kEpilogueThreads >= kGemmBlockN, "one parameter column per epilogue thread");
    int staged_n0 = -1, pre_n0 = -1;
    float2 pre = make_float2(0.f, 0.f);
    for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++it) {
      const int m0 = (tile / num_n_tiles) * (2 * kGemmBlockM) + static_cast<int>(rank) * kGemmBlockM;
      const int n0 = (tile % num_n_tiles) * args.tile_n;
      constexpr int kChunkCols = kOutF32 ? 32 : 64;  // 128 bytes of output per row and chunk
      // whole 32-column units; columns beyond n_mma (a 16-column tail no MMA wrote) are clipped by the store maps
      const int n_cols = (min(args.tile_n, args.n_mma - n0) + 31) & ~31;
      const int n_chunks = (n_cols + kChunkCols - 1) / kChunkCols;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;

      // Per-column parameters of this N tile: restaged only when the tile column changes, and then
      // from registers loaded during the previous tile (the L2 latency is off the critical path).
      if (n0 != staged_n0) {
        asm volatile("bar.sync 1, %0;" ::"n"(kEpilogueThreads) : "memory");  // previous tile's params no longer read
        if (et < kGemmBlockN) {
          const float2 p = (pre_n0 == n0) ? pre : epi::load_post(args.post, n0 + et, args.n_pad);
          epi::store_post<kBF16, kOutF32>(smem + kSmemPost, et, p);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kEpilogueThreads) : "memory");
        staged_n0 = n0;
      }
      if (tile + tile_step < num_tiles) {
        const int nn0 = ((tile + tile_step) % num_n_tiles) * args.tile_n;
        if (nn0 != n0) {
          if (et < kGemmBlockN) pre = epi::load_post(args.post, nn0 + et, args.n_pad);
          pre_n0 = nn0;
        }
      }

      ptx::mbar_wait(&tmem_full_bar[as], aphase);
      ptx::tcgen05_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * kGemmBlockN;

      if constexpr (kPool) {
        // 32-column units: this warp's 32 rows x 32 columns are loaded, rounded to the operand type (the
        // values a consumer would read), transposed through the warp's 2 KB buffer as 16-bit values
        // (lane = row writes its 16 words at word index ^ (row >> 1); lane = column then reads one
        // 16-bit value per row: both phases bank-conflict free) and summed with the mixed-precision
        // FHADD / FHFMA (fp32 += fp16, no unpacking), even and odd rows in separate chains.
        const PoolGroupMeta meta = args.pool_meta[(m0 + q * 32) >> 5];
        const bool has_a = meta.a1 > meta.a0, has_b = meta.b1 > meta.b0;
        const int n_units = n_cols >> 5;
        if (has_a || has_b) {  // else: no valid row in this 32-row group (edge / gap rows)
          uint32_t* tile = reinterpret_cast<uint32_t*>(buf);
          const uint32_t lp = static_cast<uint32_t>(lane) << 1;  // byte offset of column `lane` in an un-swizzled row
          auto row_value = [&](int r) -> uint16_t {
            return *reinterpret_cast<const uint16_t*>(buf + r * 64 + (lp ^ (static_cast<uint32_t>(r >> 1) << 2)));
          };
          for (int u = half; u < n_units; u += Cfg::kParts) {
            uint32_t r[32];
            ptx::tmem_ld_32x32b_x32(t_row + u * 32, r);
            ptx::tmem_ld_wait();
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint32_t p4[4];
              epi::pack8<kBF16>(&r[j], smem + kSmemPost, u * 32 + j, args.relu != 0, p4);
#pragma unroll
              for (int e = 0; e < 4; ++e) pk[(j >> 1) + e] = p4[e];
            }
            __syncwarp();  // the previous unit's reads of `tile` are done
#pragma unroll
            for (int j = 0; j < 16; ++j) tile[lane * 16 + (j ^ (lane >> 1))] = pk[j];
            __syncwarp();
#pragma unroll 1
            for (int seg = 0; seg < 2; ++seg) {
              const int lo = seg ? meta.b0 : meta.a0, hi = seg ? meta.b1 : meta.a1;
              if (hi <= lo) continue;
              float sx0 = 0.f, sxx0 = 0.f, sx1 = 0.f, sxx1 = 0.f;  // even / odd rows of column `lane`
              if (lo == 0 && hi == 32) {  // interior group: every row is valid
#pragma unroll
                for (int rr = 0; rr < 32; rr += 2) {
                  accum16<kBF16>(row_value(rr), sx0, sxx0);
                  accum16<kBF16>(row_value(rr + 1), sx1, sxx1);
                }
              } else {
                int rr = lo;
                for (; rr + 1 < hi; rr += 2) {
                  accum16<kBF16>(row_value(rr), sx0, sxx0);
                  accum16<kBF16>(row_value(rr + 1), sx1, sxx1);
                }
                if (rr < hi) accum16<kBF16>(row_value(rr), sx0, sxx0);
              }
              const int slot = meta.slot + ((seg && has_a) ? 1 : 0);
              // blocked layout [column block of 64][slot][sum | sum of squares][64] (the tile column base is a
              // multiple of 32, so a unit's 32 columns stay inside one block)
              const int col = n0 + u * 32;
              float* out = args.pool_partials +
                           ((static_cast<int64_t>(col >> 6) * args.pool_slots + slot) * 2) * 64 + (col & 63) + lane;
              out[0] = sx0 + sx1;
              out[64] = sxx0 + sxx1;
            }
          }
        }
      }
      for (int c = half; !kPool && c < n_chunks; c += Cfg::kParts) {
        epi::store_chunk<kBF16, kOutF32>(t_row, c, n_cols, n0, args.c_row0 + m0 + q * 32, lane, buf, smem + kSmemPost, args.relu != 0,
                                         &args.c_map, &args.c_map_half);
      }
      ptx::tcgen05_fence_before();
      __syncwarp();
      if (ptx::elect_one_sync()) ptx::mbar_arrive_cluster(&tmem_empty_bar[as], 0);
    }
    if (ptx::elect_one_sync()) ptx::tma_store_wait<0>();
  }

  ptx::tcgen05_fence_before();
  ptx::cluster_sync_all();  // the peer may still be reading this CTA's smem / arriving on its barriers
  if (warp == 1) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc_pair(tmem_base, kTmemCols);
  }
}

}  // namespace

namespace {
template <typename Kern>
cudaError_t LaunchPair(Kern kern, int threads, int smem, int stages, const TdnnGemmArgs& args, int num_sms, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  const int tiles = ((args.num_m_tiles + 1) / 2) * args.num_n_tiles;
  if (tiles <= 0) return cudaSuccess;
  int clusters = tiles < num_sms / 2 ? tiles : num_sms / 2;
  if (args.resident_b) {
    if (args.num_kb > stages) return cudaErrorInvalidValue;
    // a pair must keep its tile column: the stride over tiles is a multiple of the column count
    // (tiles itself is one, so at least num_n_tiles pairs remain)
    clusters = clusters / args.num_n_tiles * args.num_n_tiles;
    if (clusters == 0) clusters = args.num_n_tiles;
  }
  kern<<<2 * clusters, threads, smem, stream>>>(args);
  return cudaGetLastError();
}
}  // namespace

cudaError_t LaunchTdnnGemmPair(const TdnnGemmArgs& args, int num_sms, cudaStream_t stream) {
  const bool bf16 = args.ab_format == 1;
  using Pool = PairCfg<true, false>;
  using Wide = PairCfg<false, true>;
  using Plain = PairCfg<false, false>;
  if (args.pool_meta)
    return bf16 ? LaunchPair(tdnn_gemm_pair_kernel<true, false, true>, Pool::kThreads, Pool::kSmemTotal, Pool::kStages, args, num_sms, stream)
                : LaunchPair(tdnn_gemm_pair_kernel<false, false, true>, Pool::kThreads, Pool::kSmemTotal, Pool::kStages, args, num_sms, stream);
  if (args.out_f32)
    return bf16 ? LaunchPair(tdnn_gemm_pair_kernel<true, true, false>, Plain::kThreads, Plain::kSmemTotal, Plain::kStages, args, num_sms, stream)
                : LaunchPair(tdnn_gemm_pair_kernel<false, true, false>, Plain::kThreads, Plain::kSmemTotal, Plain::kStages, args, num_sms, stream);
  // first layers (resident weights, at most two K blocks): the 16-warp epilogue variant
  if (args.resident_b && args.wide_epilogue && args.num_kb <= Wide::kStages)
    return bf16 ? LaunchPair(tdnn_gemm_pair_kernel<true, false, false, true>, Wide::kThreads, Wide::kSmemTotal, Wide::kStages, args, num_sms, stream)
                : LaunchPair(tdnn_gemm_pair_kernel<false, false, false, true>, Wide::kThreads, Wide::kSmemTotal, Wide::kStages, args, num_sms, stream);
  return bf16 ? LaunchPair(tdnn_gemm_pair_kernel<true, false, false>, Plain::kThreads, Plain::kSmemTotal, Plain::kStages, args, num_sms, stream)
              : LaunchPair(tdnn_gemm_pair_kernel<false, false, false>, Plain::kThreads, Plain::kSmemTotal, Plain::kStages, args, num_sms, stream);
}

}  // namespace cvx
