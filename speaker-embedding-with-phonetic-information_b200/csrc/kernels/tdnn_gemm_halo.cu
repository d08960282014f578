// K1 (production variant) -- CTA-pair TDNN implicit GEMM with tap-halo reuse.
//
// A spliced layer Append(Offset(x,o_1..o_n)) reads, for one 64-column block of x, the
// row windows [m0+o_t, m0+o_t+128) -- n nearly identical activation tiles.  This kernel
// loads ONE tile of 128+span rows (span = max o - min o <= 16) per column block and
// issues the MMAs of every tap from it: tap t reads rows [o_t - o_min, +128) of the tile
// through a shared-memory descriptor whose start address is advanced by whole 128-byte
// rows (the swizzle phase follows the absolute address, as it does for the 32-byte K steps).  Activation-tile traffic (L2 -> SM and smem fill) drops by
// the tap count; weight half-tiles stream through their own ring.
//
//   A ring   4 slots x (144 rows x 128 B)   one activation tile per (source, column block)
//   B ring   7 slots x (128 rows x 128 B)   this CTA's half of one W tile per (group, tap)
//   (narrow layers, N <= 128: 7 activation slots + 7 weight slots of 64 rows, see Ring<>)
//   both CTAs of the pair: warp 0 = TMA producer (A tile, then its taps' W tiles, in
//   consumption order); leader CTA warp 1 = MMA issuer (tcgen05.mma.cta_group::2, M=256);
//   warps 2-9 = epilogue (see tdnn_gemm_pair.cu, unchanged).
#include "epilogue.cuh"
#include "kernels.h"
#include "ptx.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace cvx {
namespace {

constexpr int kASlotBytes = (kGemmBlockM + kMaxHaloSpan) * 128;  // 18432
constexpr int kCWarpBytes = 32 * 128;
constexpr int kTmemCols = 512;
constexpr int kThreads = 320;
constexpr int kEpilogueThreads = 256;
constexpr int kEpilogueWarps = 8;

// Ring geometry.  Wide tiles (N > 128): 4 activation slots + 7 weight slots of 16 KB.  Narrow
// layers (N <= 128, e.g. the 1950 -> 128 bottleneck) need only 8 KB per weight half-tile and their
// tap groups last half as many MMA cycles, so the same lookahead in TIME needs more activation
// slots: 7 + 7 x 8 KB (with 4 the MMA warp waited on HBM latency: 63 % tensor-pipe active).
template <bool kNarrow>
struct Ring {
  static constexpr int kAStages = kNarrow ? 7 : 4;
  static constexpr int kBStages = 7;
  static constexpr int kBSlotBytes = kNarrow ? (kGemmBlockN / 4) * 128 : (kGemmBlockN / 2) * 128;  // 8192 : 16384
  static constexpr int kSmemA = 0;
  static constexpr int kSmemB = kSmemA + kAStages * kASlotBytes;
  static constexpr int kSmemC = kSmemB + kBStages * kBSlotBytes;
  static constexpr int kSmemPost = kSmemC + 8 * kCWarpBytes;
  static constexpr int kSmemBars = kSmemPost + kGemmBlockN * 8;
  static constexpr int kNumBars = 2 * kAStages + 2 * kBStages + 4;
  static constexpr int kSmemTmemPtr = kSmemBars + kNumBars * 8;
  static constexpr int kSmemTotal = kSmemTmemPtr + 32;
  static_assert(kASlotBytes % 1024 == 0 && kSmemB % 1024 == 0 && kSmemC % 1024 == 0, "swizzle alignment");
  static_assert(kSmemTotal <= 232448, "shared memory budget");
};

template <bool kBF16, bool kOutF32, bool kNarrow>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
    tdnn_gemm_halo_kernel(const __grid_constant__ TdnnGemmHaloArgs args) {
  using R = Ring<kNarrow>;
  constexpr int kAStages = R::kAStages, kBStages = R::kBStages, kBSlotBytes = R::kBSlotBytes;
  constexpr int kSmemA = R::kSmemA, kSmemB = R::kSmemB, kSmemC = R::kSmemC, kSmemPost = R::kSmemPost,
                kSmemBars = R::kSmemBars, kSmemTmemPtr = R::kSmemTmemPtr;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smem_a = smem + kSmemA;
  uint8_t* smem_b = smem + kSmemB;
  uint8_t* smem_c = smem + kSmemC;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + kSmemBars);  // leader's are used
  uint64_t* a_empty = a_full + kAStages;                             // per CTA
  uint64_t* b_full = a_empty + kAStages;                             // leader's are used
  uint64_t* b_empty = b_full + kBStages;                             // per CTA
  uint64_t* tmem_full_bar = b_empty + kBStages;                      // per CTA
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;                      // leader's are used
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + kSmemTmemPtr);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x) >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();

  if (threadIdx.x == 0) {
    if (ptx::smem_u32(smem) & 1023u) __trap();
    for (int i = 0; i < kAStages; ++i) {
      ptx::mbar_init(&a_full[i], 2);
      ptx::mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < kBStages; ++i) {
      ptx::mbar_init(&b_full[i], 2);
      ptx::mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tmem_full_bar[i], 1);
      ptx::mbar_init(&tmem_empty_bar[i], 2 * kEpilogueWarps);
    }
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
  ptx::cluster_sync_all();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int num_n_tiles = args.num_n_tiles;
  const int num_tiles = ((args.num_m_tiles + 1) >> 1) * num_n_tiles;  // 256-row tiles
  const int num_groups = args.num_groups;
  const int first_tile = blockIdx.x >> 1;
  const int tile_step = gridDim.x >> 1;

  if (warp == 0) {
    // ------------------------------------------------------ TMA producer --
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    const uint32_t a_bytes = static_cast<uint32_t>(args.a_box_rows) * 128u;
    const uint32_t b_bytes = static_cast<uint32_t>(args.b_box_rows_pair) * 128u;
    for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
      const int m0 = (tile / num_n_tiles) * (2 * kGemmBlockM) + static_cast<int>(rank) * kGemmBlockM;
      const int n0 = (tile % num_n_tiles) * args.tile_n;
      const int n_cols = min(args.tile_n, args.n_mma - n0);
      const int nb0 = n0 + static_cast<int>(rank) * (n_cols >> 1);  // this CTA's half of the W tile
      for (int g = 0; g < num_groups; ++g) {
        const HaloGroup& grp = args.groups[g];
        ptx::mbar_wait(&a_empty[sa], pa ^ 1);
        if (ptx::elect_one_sync()) {
          ptx::mbar_arrive_expect_tx_cluster(&a_full[sa], a_bytes, 0);
          ptx::tma_load_2d_pair(smem_a + sa * kASlotBytes, &args.a_maps[grp.src], &a_full[sa], 0, grp.col,
                                m0 + grp.rowoff);
        }
        __syncwarp();
        if (++sa == kAStages) {
          sa = 0;
          pa ^= 1;
        }
        for (int t = 0; t < grp.ntaps; ++t) {
          ptx::mbar_wait(&b_empty[sb], pb ^ 1);
          if (ptx::elect_one_sync()) {
            ptx::mbar_arrive_expect_tx_cluster(&b_full[sb], b_bytes, 0);
            ptx::tma_load_2d_pair(smem_b + sb * kBSlotBytes, &args.b_map_pair, &b_full[sb], 0,
                                  grp.wkb[t] * kGemmBlockK, nb0);
          }
          __syncwarp();
          if (++sb == kBStages) {
            sb = 0;
            pb ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------- MMA issuer (leader CTA) --
    if (rank == 0) {
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      int it = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++it) {
        const int n0 = (tile % num_n_tiles) * args.tile_n;
        const int n_cols = min(args.tile_n, args.n_mma - n0);  // multiple of 16
        const uint32_t idesc = ptx::make_idesc_f16(kBF16 ? 1u : 0u, 2 * kGemmBlockM, n_cols);
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        ptx::mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
        ptx::tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + as * kGemmBlockN;
        uint32_t first = 0;  // becomes non-zero after the first MMA of the tile
        for (int g = 0; g < num_groups; ++g) {
          const HaloGroup& grp = args.groups[g];
          ptx::mbar_wait(&a_full[sa], pa);
          const uint32_t a_addr = ptx::smem_u32(smem_a + sa * kASlotBytes);
          for (int t = 0; t < grp.ntaps; ++t) {
            ptx::mbar_wait(&b_full[sb], pb);
            ptx::tcgen05_fence_after();
            // tap t: rows [shift, shift+128) of the halo tile; base_offset = swizzle phase of the start row
            const uint32_t shift = static_cast<uint32_t>(grp.shift[t]);
            // (the 128-byte swizzle is a function of the absolute smem address bits, which TMA and the
            // MMA unit both see, so a start address advanced by whole rows needs no base_offset;
            // args.base_offset_mode = 1 sets the descriptor's base_offset field instead -- test aid)
            const uint64_t a_desc = ptx::make_umma_desc_sw128(a_addr + shift * 128u) |
                                    (args.base_offset_mode ? (static_cast<uint64_t>(shift & 7u) << 49) : 0ull);
            const uint64_t b_desc = ptx::make_umma_desc_sw128(ptx::smem_u32(smem_b + sb * kBSlotBytes));
            if (ptx::elect_one_sync()) {
#pragma unroll
              for (int k = 0; k < kGemmBlockK / 16; ++k) {
                if (k < grp.k16) {
                  ptx::umma_f16_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, first | static_cast<uint32_t>(k));
                }
              }
              ptx::umma_commit_pair_mc(&b_empty[sb], 0x3);                          // W slot free in both CTAs
              if (t == grp.ntaps - 1) ptx::umma_commit_pair_mc(&a_empty[sa], 0x3);  // activation slot free
              if (t == grp.ntaps - 1 && g == num_groups - 1) ptx::umma_commit_pair_mc(&tmem_full_bar[as], 0x3);
            }
            __syncwarp();
            first = 1;
            if (++sb == kBStages) {
              sb = 0;
              pb ^= 1;
            }
          }
          if (++sa == kAStages) {
            sa = 0;
            pa ^= 1;
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
        const float2 p = (pre_n0 == n0) ? pre : epi::load_post(args.post, n0 + et, args.n_pad);
        epi::store_post<kBF16, kOutF32>(smem + kSmemPost, et, p);
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
        epi::store_chunk<kBF16, kOutF32>(t_row, c, n_cols, n0, m0 + q * 32, lane, buf, smem + kSmemPost, args.relu != 0,
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

template <bool kNarrow>
cudaError_t LaunchHalo(const TdnnGemmHaloArgs& args, int num_sms, cudaStream_t stream) {
  const bool bf16 = args.ab_format == 1;
  auto kern = args.out_f32 ? (bf16 ? tdnn_gemm_halo_kernel<true, true, kNarrow> : tdnn_gemm_halo_kernel<false, true, kNarrow>)
                           : (bf16 ? tdnn_gemm_halo_kernel<true, false, kNarrow> : tdnn_gemm_halo_kernel<false, false, kNarrow>);
  constexpr int kSmemTotal = Ring<kNarrow>::kSmemTotal;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal);
  if (e != cudaSuccess) return e;
  const int tiles = ((args.num_m_tiles + 1) / 2) * args.num_n_tiles;
  if (tiles <= 0) return cudaSuccess;
  const int clusters = tiles < num_sms / 2 ? tiles : num_sms / 2;
  kern<<<2 * clusters, kThreads, kSmemTotal, stream>>>(args);
  return cudaGetLastError();
}

cudaError_t LaunchTdnnGemmHalo(const TdnnGemmHaloArgs& args, int num_sms, cudaStream_t stream) {
  // a W half-tile of <= 64 rows fits the narrow ring's 8 KB weight slots
  if (args.b_box_rows_pair <= kGemmBlockN / 4) return LaunchHalo<true>(args, num_sms, stream);
  return LaunchHalo<false>(args, num_sms, stream);
}

}  // namespace cvx
