// Host-callable launchers of the CUDA kernels (one .cu per kernel family).
#ifndef CVX_KERNELS_KERNELS_H_
#define CVX_KERNELS_KERNELS_H_

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cvx {

enum ElemType : int32_t { kF16 = 0, kBF16 = 1, kF32 = 2 };
inline size_t ElemSize(int32_t e) { return e == kF32 ? 4 : 2; }

// ---- K1: TDNN layer as implicit GEMM -------------------------------------
// out[r, n] = post( sum_kb A_kb[r + rowoff_kb, col_kb : col_kb+64] . Wp[n, 64*kb : 64*kb+64] )
// post(v)  = (relu ? max(v + b'[n], 0) : v + b'[n]) + o[n]
// (test-mode BatchNorm scale is folded into Wp/b' at load; o is its offset.)
constexpr int kGemmBlockM = 128;
constexpr int kGemmBlockN = 256;
constexpr int kGemmBlockK = 64;
constexpr int kMaxKBlocks = 160;
constexpr int kMaxASources = 4;

struct KBlockDesc {
  int32_t src;     // index into the layer's A sources
  int32_t col;     // first column in that source buffer (multiple of 64)
  int32_t rowoff;  // time offset of the tap
  int32_t k16;     // 16-column K steps of this block that hold real columns (1..4)
};

// Fused pooling (pair kernel, pooled layer only): the epilogue reduces each 32-row group of the
// output tile over its valid rows instead of storing the activations.  A group holds at most two
// runs of valid rows (tail of one chunk, head of the next): rows [a0, a1) -> slot, rows [b0, b1)
// -> the following slot.  Empty runs have a1 <= a0 / b1 <= b0.
struct PoolGroupMeta {
  int32_t slot;
  uint8_t a0, a1, b0, b1;
};

struct TdnnGemmArgs {
  CUtensorMap a_maps[kMaxASources];  // [rows_cap, ld] activations, box 64 x 128, SWIZZLE_128B
  CUtensorMap b_map;                 // Wp [n_pad, k_pad], box 64 x 256, SWIZZLE_128B
  CUtensorMap b_map_pair;            // same matrix, box 64 x 128 (one CTA's half of a W tile, cta_group::2)
  CUtensorMap c_map;                 // out [rows_cap, ld_out], box 64 x 32, SWIZZLE_128B
  CUtensorMap c_map_half;            // same matrix, box 32 x 32, no swizzle (last half chunk of a tile)
  KBlockDesc kb[kMaxKBlocks];
  int32_t num_kb;
  int32_t n_pad;          // output row pitch, multiple of 64
  int32_t n_mma;          // roundup(out_dim, 16 or 32): columns the MMAs produce; the store maps end here
  int32_t tile_n;         // N tile width: n_mma split into equal tiles, multiple of 32, <= 256
  int32_t num_m_tiles;    // ceil(rows / 128)
  int32_t num_n_tiles;    // ceil(n_pad / 256)
  int32_t b_box_rows;     // rows of the W box = min(256, n_pad)
  int32_t b_box_rows_pair;  // rows of the half-tile W box = min(128, n_pad)
  const float2* post;     // [n_pad] {b', o}
  int32_t relu;
  int32_t ab_format;      // 0 = fp16, 1 = bf16 (UMMA F16F32Format)
  int32_t out_f32;        // pair kernel only: write fp32 output (c_map is an fp32 map with 32-column boxes)
  int32_t c_row0;         // row offset added to every output store (segment layers write into a slice)
  const PoolGroupMeta* pool_meta;  // non-NULL: fused pooling (see PoolGroupMeta), no activation store
  // partials, blocked by 64 columns so that the finalize kernel streams them:
  // [column block][pool_slots][2 (sum x, sum x^2)][64]
  float* pool_partials;
  int32_t pool_ld, pool_slots;
  int32_t num_stages;     // smem ring depth actually used (<= the kernel's maximum; tuning aid)
  // pair kernel, layers with at most 2 K blocks (K <= 128, the first TDNN layers): every CTA pair
  // keeps ONE tile column for the whole launch and loads its weight half-tiles once instead of with
  // every tile -- a quarter of the bytes such a tile moves through the SM
  int32_t resident_b;
  int32_t wide_epilogue;  // with resident_b: run the 16-epilogue-warp variant (tuning aid: CVX_WIDE_EPILOGUE=0 turns it off)
};

size_t TdnnGemmSmemBytes();
cudaError_t LaunchTdnnGemm(const TdnnGemmArgs& args, int num_sms, cudaStream_t stream);
// CTA-pair (cta_group::2, 256-row tiles) variant of the same layer; the production path.
cudaError_t LaunchTdnnGemmPair(const TdnnGemmArgs& args, int num_sms, cudaStream_t stream);

// Tap-halo variant (tdnn_gemm_halo.cu): one activation tile of 128 + span rows per (source,
// 64-column block) serves every tap of that source; the production kernel.
constexpr int kMaxHaloSpan = 16;    // max (largest - smallest) tap offset sharing one tile
constexpr int kMaxHaloTaps = 4;
constexpr int kMaxHaloGroups = 160;
struct HaloGroup {
  int32_t src;                  // index into a_maps
  int32_t col;                  // first column of the 64-column block in the source
  int32_t rowoff;               // smallest tap offset of the group (row coordinate of the tile = m0 + rowoff)
  int32_t ntaps;
  int32_t k16;                  // 16-column K steps with real columns
  int32_t shift[kMaxHaloTaps];  // tap offset - rowoff (rows into the tile)
  int32_t wkb[kMaxHaloTaps];    // 64-column block index of the tap's weights in Wp
};
struct TdnnGemmHaloArgs {
  CUtensorMap a_maps[kMaxASources];  // box 64 x a_box_rows
  CUtensorMap b_map_pair;            // box 64 x b_box_rows_pair
  CUtensorMap c_map;
  CUtensorMap c_map_half;
  HaloGroup groups[kMaxHaloGroups];
  int32_t num_groups;
  int32_t a_box_rows;       // 128 + span of this layer (<= 144)
  int32_t b_box_rows_pair;
  int32_t n_pad, n_mma, tile_n, num_m_tiles, num_n_tiles;
  const float2* post;
  int32_t relu, ab_format, out_f32;
  int32_t base_offset_mode;  // test aid: also set the descriptor's base_offset to (shift & 7)
};
cudaError_t LaunchTdnnGemmHalo(const TdnnGemmHaloArgs& args, int num_sms, cudaStream_t stream);

// fp32 CUDA-core version of the same contraction (CVX_PRECISION_FP32; validation aid).
struct TdnnGemmSimtArgs {
  const float* a_src[kMaxASources];
  int32_t a_ld[kMaxASources];
  KBlockDesc kb[kMaxKBlocks];
  int32_t num_kb;
  const float* w;   // Wp [n_pad, k_pad] fp32
  int32_t k_pad;
  int32_t n_pad;
  const float2* post;
  int32_t relu;
  float* out;
  int32_t ld_out;
  int64_t rows;      // rows to compute
  int64_t rows_cap;  // rows that exist in the source buffers
};
cudaError_t LaunchTdnnGemmSimt(const TdnnGemmSimtArgs& args, cudaStream_t stream);

// ---- input staging: fp32 features -> spliced operand rows ----------------
constexpr int kMaxStageTaps = 16;
constexpr int kStageMaxOffset = 2047;  // |time offset| representable in the column map
struct StageTap {
  int32_t stream;
  int32_t offset;
  int32_t dim;
  int32_t col0;  // first output column of this tap
};
struct StageArgs {
  const float* feats[kMaxASources];  // per stream, row 0 = first row of the batch
  int32_t feat_dim[kMaxASources];
  const uint32_t* colmap;            // [ld_out] device: output column -> (stream, offset, source column)
  void* out;        // [rows, ld_out] of `elem`
  int32_t ld_out;   // multiple of 64; columns beyond the taps are zero-filled
  int32_t elem;
  int64_t rows;
  int64_t src_row_min, src_row_max;  // source rows that exist, relative to feats[] row 0
  // tiled fast path: every mapped column reads stream `tile_stream` at offsets in [off_min, off_max]
  int32_t tile_stream;               // -1: generic gather
  int32_t off_min, off_max;
};
cudaError_t LaunchStageInput(const StageArgs& args, cudaStream_t stream);

// counts the fp16 values of buf [rows, dim] (row pitch ld) that sit at the range limit (saturation.cu)
cudaError_t LaunchCountSaturated(const void* buf, int64_t rows, int32_t ld, int32_t dim, unsigned long long* counter,
                                 cudaStream_t stream);

// ---- K2: ragged mean + stddev pooling -------------------------------------
struct PoolSlab {
  int64_t row_begin;  // first row (in the activation buffer) of this slab
  int32_t nrows;
  int32_t chunk;
};
constexpr int kPoolSlabRows = 256;
// partials: [num_slabs][2][ld] fp32 = per-column (mean, M2) of each slab
cudaError_t LaunchStatsPoolPartial(const void* act, int32_t elem, int32_t ld, const PoolSlab* slabs,
                                   int32_t num_slabs, float* partials, cudaStream_t stream);
// stats[c] = [mean(dim) | sqrt(max(var, floor))(dim)] (or just mean), merging the chunk's slabs
cudaError_t LaunchStatsPoolFinalize(const float* partials, int32_t ld, const PoolSlab* slabs,
                                    const int32_t* chunk_slab_off, int32_t num_chunks, int32_t dim,
                                    int32_t output_stddevs, float variance_floor, float* stats,
                                    int32_t ld_stats, cudaStream_t stream);

// fp32 [m, k] -> fp16 hi / lo split (x = hi + lo to ~2^-22), pad columns [k, ld_out) zeroed
cudaError_t LaunchSplitF32(const float* in, int32_t ld_in, int64_t m, int32_t k, void* hi, void* lo,
                           int32_t ld_out, int32_t elem, cudaStream_t stream);

// fused-pooling tables of one batch, built on the device from the chunks' valid row ranges
// [lo[c], hi[c]] (relative to the batch, hi inclusive, increasing): per 32-row group the (at most
// two) runs of valid rows and their slots, and the number of rows of every slot
cudaError_t LaunchPoolMeta(const int32_t* lo, const int32_t* hi, const int32_t* slot_begin, int32_t num_chunks,
                           int64_t num_groups, PoolGroupMeta* meta, int32_t* slot_n, cudaStream_t stream);

// fused-pooling finalize: chunk c owns slots [slot_begin[c], slot_begin[c+1]) with slot_n[] valid rows each;
// partials in the blocked layout of TdnnGemmArgs::pool_partials with pool_slots = total_slots
cudaError_t LaunchStatsPoolFinalizeFused(const float* partials, int32_t ld, const int32_t* slot_begin,
                                         const int32_t* slot_n, int32_t num_chunks, int32_t total_slots,
                                         int32_t dim, int32_t output_stddevs, float variance_floor, float* stats,
                                         int32_t ld_stats, cudaStream_t stream);

// ---- K3: segment-level affine and per-utterance weighted average ----------
cudaError_t LaunchSegmentAffine(const float* in, int32_t ld_in, int64_t m, int32_t k, const float* w,
                                const float* bias, int32_t relu, const float* scale, const float* offset,
                                int32_t n, float* out, int32_t ld_out, cudaStream_t stream);
cudaError_t LaunchUttAverage(const float* chunk_emb, int32_t dim, const int64_t* utt_chunk_off,
                             const float* chunk_weight, int32_t num_utts, float* out, cudaStream_t stream);

// ---- feature pre-step: sliding-window CMN + voiced-frame selection (prestep.cu) ----
// Both kernels take the call's whole index arrays and a sub-range (utterances [utt0, utt0 + num_utts) =
// runs [run0, run0 + num_runs), kept rows [kept0, kept0 + num_kept)), so that a call can be
// pre-processed piece by piece as its raw features arrive.
// Sliding CMN + selection: utt_run_start[u] = number of CmnRunFrames()-frame runs before utterance u.
cudaError_t LaunchCmnSliding(const float* x, int32_t dim, const int64_t* utt_start, const int64_t* utt_run_start,
                             const int64_t* utt_kept_start, const int32_t* keep, int32_t utt0, int32_t num_utts, int64_t run0,
                             int64_t num_runs, int32_t cmn_window, int32_t center, int32_t min_window, float* out,
                             cudaStream_t stream);
int32_t CmnRunFrames();
// plain selection: out[j] = x[frame keep[j] of utterance kept_utt[j]]
cudaError_t LaunchCmnSelect(const float* x, int32_t dim, const int64_t* utt_start, const int32_t* kept_utt,
                            const int32_t* keep, int64_t kept0, int64_t num_kept, float* out, cudaStream_t stream);

// ---- frame-level outputs (frame_output.cu): edge replication in, compaction (+log-softmax) out ----
// utt_start [num_utts + 1] in compact rows; utterance u owns padded rows from utt_start[u] + u * (left + right)
cudaError_t LaunchPadEdges(const float* src, int32_t dim, const int64_t* utt_start, int32_t num_utts, int32_t left,
                           int32_t right, int64_t padded_rows, float* dst, cudaStream_t stream);
cudaError_t LaunchFrameOutput(const float* act, int32_t ld, int32_t dim, const int64_t* utt_start, int32_t num_utts,
                              int32_t left, int32_t right, int64_t rows, int32_t log_softmax, int32_t apply_exp,
                              float* out, cudaStream_t stream);

// ---- embedding post-step (vector_chain.cu): subtract mean -> transform -> length normalisation ----
// out [num_vectors, transform ? t_rows : dim]; ratios [num_vectors] or NULL
cudaError_t LaunchVectorChain(const float* in, int64_t num_vectors, int32_t dim, const float* mean, const float* transform,
                              int32_t t_rows, int32_t t_cols, int32_t normalize, int32_t scaleup, int32_t keep_zero,
                              float* out, float* ratios, cudaStream_t stream);

// ---- PLDA scoring (plda.cu): transform + length normalisation of vectors, then one warp per trial ----
cudaError_t LaunchPldaTransform(const float* in, int64_t num_vectors, int32_t dim_in, const double* mean, const double* transform,
                                int32_t dim, const double* psi, const int32_t* num_examples, int32_t normalize_length,
                                int32_t simple_norm, double* out, cudaStream_t stream);
cudaError_t LaunchPldaScore(const double* train, const int32_t* train_n, const double* test, const double* psi, int32_t dim,
                            const int32_t* trial_train, const int32_t* trial_test, int64_t num_trials, float* scores,
                            cudaStream_t stream);

}  // namespace cvx
#endif  // CVX_KERNELS_KERNELS_H_
