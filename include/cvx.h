/*
 * cvx.h -- C ABI of the B200-native x-vector / c-vector extraction path.
 *
 * The reference (mycrazycracy/speaker-embedding-with-phonetic-information) has no
 * FFI for this path: its boundary is the command line of upstream Kaldi's
 * `nnet3-xvector-compute` as driven by
 *   egs/sre/v5/sid/nnet3/xvector/extract_xvectors_new.sh:57-94
 * (and `nnet3-cvector-compute-multiple-input`,
 *   egs/sre/v5/sid/nnet3_cvector/cvector/extract_cvectors_with_am.sh:99-110).
 * The drop-in programs of this repo (csrc/bin/) keep that command line and call
 * the CUDA side only through the functions below; INTEGRATION.md shows the stub a
 * Kaldi maintainer would add to call them from the original binary instead.
 *
 * Conventions: plain pointers and sizes, no C++ / torch types; every function
 * returns CVX_OK (0) or a negative error code and never throws; the message for
 * the last error on the calling thread is cvx_last_error().  A cvx_engine is
 * bound to one CUDA device and is NOT thread-safe (one engine per worker/GPU);
 * a cvx_model is immutable and may be shared.  The caller owns all buffers it
 * passes in.  There is no CPU fallback: without a usable CUDA device
 * cvx_engine_create fails with CVX_ERR_CUDA.
 */
#ifndef CVX_H_
#define CVX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CVX_ABI_VERSION 1

enum {
  CVX_OK = 0,
  CVX_ERR_INVALID = -1,     /* bad argument */
  CVX_ERR_IO = -2,          /* cannot open / read */
  CVX_ERR_FORMAT = -3,      /* malformed model file */
  CVX_ERR_UNSUPPORTED = -4, /* graph uses something outside the extraction path */
  CVX_ERR_CUDA = -5,        /* CUDA runtime / driver failure, or no device */
  CVX_ERR_NOMEM = -6
};

/* Operand type of the frame-level (TDNN) contractions; accumulation is fp32. */
enum {
  CVX_PRECISION_FP16 = 0, /* tcgen05 kind::f16, fp16 operands (default: meets 1e-3) */
  CVX_PRECISION_BF16 = 1, /* tcgen05 kind::f16, bf16 operands */
  CVX_PRECISION_FP32 = 2  /* fp32 CUDA-core kernels; validation aid, slow */
};

const char* cvx_last_error(void);
const char* cvx_version(void);

/* ------------------------------------------------------------------ plan --
 * What a host that already holds the network in memory hands over.  This is
 * the information Kaldi's nnet3 compiler derives for the sub-graph reachable
 * from output-node `output` (replaces Nnet / CachingOptimizingCompiler /
 * NnetComputer in nnet3-xvector-compute.cc).
 */
typedef struct cvx_tap {
  int32_t src;    /* >= 0: index of an earlier frame layer; < 0: input stream (-1 - src) */
  int32_t offset; /* Offset(src, offset): value at t is the source at t+offset */
} cvx_tap;

typedef struct cvx_layer {
  const char* name;     /* for logs; may be NULL */
  int32_t num_taps;     /* Append(...) arity; segment layers: 0 (= previous layer) */
  const cvx_tap* taps;
  int32_t in_dim;       /* sum of the tapped sources' dims */
  int32_t out_dim;
  const float* weight;  /* [out_dim, in_dim] row-major (Kaldi <LinearParams>), y = x W^T + b */
  const float* bias;    /* [out_dim] or NULL */
  int32_t relu;         /* apply max(.,0) after the affine */
  const float* scale;   /* [out_dim] or NULL: test-mode BatchNorm, y = r*scale + offset */
  const float* offset;  /* [out_dim] or NULL */
} cvx_layer;

typedef struct cvx_plan {
  int32_t num_inputs;
  const char* const* input_names;
  const int32_t* input_dims;
  int32_t num_frame_layers;
  const cvx_layer* frame_layers; /* topological order */
  int32_t pool_src;              /* frame layer feeding StatisticsExtraction */
  int32_t pool_left_context;     /* StatisticsPooling window [t-left, t+right] at t=0 */
  int32_t pool_right_context;
  int32_t pool_output_stddevs;   /* 1: [mean | stddev], 0: [mean] */
  float pool_variance_floor;     /* 1e-10 upstream */
  int32_t num_segment_layers;
  const cvx_layer* segment_layers; /* chain applied to the pooled vector; last one = embedding */
} cvx_plan;

/* ----------------------------------------------------------------- model -- */
typedef struct cvx_model cvx_model;

/* Reads an nnet3 raw model (text or binary; `rxfilename` may be a path, "-",
 * or a command ending in '|' as in extract_xvectors_new.sh:59), optionally
 * redirects output-node `output` to `output_node` (what
 * `nnet3-copy --nnet-config=extract.config` does, :58-59), prunes to the
 * reachable sub-graph, precomputes test-mode BatchNorm scale/offset and
 * lowers it to a plan.  Replaces ReadKaldiObject + SetBatchnormTestMode +
 * SetDropoutTestMode + CollapseModel in nnet3-xvector-compute.cc. */
int cvx_model_load(const char* rxfilename, const char* output_node, cvx_model** out);

/* The same for the frame-level program nnet3-compute: the output is a layer of the
 * network evaluated at every frame, e.g. the phonetic bottleneck `tdnn5.batchnorm` or the
 * log-posteriors `output.log-softmax`
 * (egs/sre/v5/sid/nnet3_cvector/cvector/extract_am_embedding.sh:56-85,
 *  egs/sre/v5/sid/nnet3_cvector/cvector/extract_log_post.sh:68-85).  Engines of such a model
 * serve cvx_compute_frames only; cvx_model_info.pooled_dim is 0 and embedding_dim is the
 * output dimension. */
int cvx_model_load_frames(const char* rxfilename, const char* output_node, cvx_model** out);

/* Deep-copies a caller-built plan. */
int cvx_model_from_plan(const cvx_plan* plan, cvx_model** out);

void cvx_model_free(cvx_model* model);

typedef struct cvx_model_info {
  int32_t num_inputs;
  int32_t embedding_dim;
  int32_t left_context;   /* of the pooled frame layer: frame t valid iff L <= t <= T-1-R */
  int32_t right_context;
  int32_t pooled_dim;     /* dim of the pooled vector (e.g. 3000) */
  int32_t num_frame_layers;
  int32_t num_segment_layers;
  int64_t mac_per_frame;  /* algorithmic multiply-accumulates per input frame, frame layers */
  int64_t mac_per_chunk;  /* segment layers, per chunk */
} cvx_model_info;

int cvx_model_get_info(const cvx_model* model, cvx_model_info* info);
const char* cvx_model_input_name(const cvx_model* model, int32_t i);
int32_t cvx_model_input_dim(const cvx_model* model, int32_t i);
/* One line per fused layer, for --verbose logs.  Pointer valid until model is freed. */
const char* cvx_model_describe(const cvx_model* model);

/* ---------------------------------------------------------------- engine -- */
typedef struct cvx_engine cvx_engine;

typedef struct cvx_engine_options {
  int32_t device;          /* CUDA ordinal */
  int32_t precision;       /* CVX_PRECISION_* */
  int64_t max_batch_rows;  /* rows (frames) per packed ragged batch; 0 = default */
  int32_t profile;         /* 1: record per-kernel CUDA-event times (cvx_engine_get_profile) */
  int32_t reserved[7];     /* zero.  reserved[0]: debug kernel selection (tests); reserved[1] != 0: range check mode
                            * for fp16 operands -- every layer's output is stored and inspected for values clipped
                            * at +-65504 (cvx_engine_saturation_counts); slower, for validating a new model */
} cvx_engine_options;

void cvx_engine_options_default(cvx_engine_options* opts);

/* Uploads the re-packed weights, allocates activation buffers lazily, builds
 * the TMA descriptors.  Replaces CuDevice::SelectGpuId + compiler setup. */
int cvx_engine_create(const cvx_model* model, const cvx_engine_options* opts, cvx_engine** out);
void cvx_engine_destroy(cvx_engine* engine);

#define CVX_FLAG_DEVICE_INPUT 1u  /* feats[] are device pointers on the engine's device */
#define CVX_FLAG_DEVICE_OUTPUT 2u /* out is a device pointer */

/* THE HOT PATH.  Replaces the per-utterance / per-chunk loop of
 * nnet3-xvector-compute.cc main() + RunNnetComputation():
 * forward of every chunk and the frame-weighted average per utterance.
 *
 *   feats[s]      [total_rows, input_dim(s)] fp32 row-major, all utterances packed
 *                 row-wise; the same row ranges apply to every stream
 *   chunk_start[] first row of chunk c (sorted ascending, chunks do not overlap)
 *   chunk_len[]   rows of chunk c (after any --pad-input replication done by the host)
 *   chunk_utt[]   utterance index of chunk c (non-decreasing)
 *   chunk_weight[] averaging weight of chunk c (upstream: its un-padded frame
 *                 count); NULL = chunk_len
 *   out           [num_utts, embedding_dim] fp32; utterances without chunks get zeros
 *
 * A chunk shorter than left_context+right_context+1 rows has no computable
 * frame: CVX_ERR_INVALID (the host filters these with min-chunk-size). */
int cvx_extract(cvx_engine* engine, const float* const* feats, int32_t num_streams,
                int64_t total_rows, const int64_t* chunk_start, const int32_t* chunk_len,
                const int32_t* chunk_utt, const float* chunk_weight, int64_t num_chunks,
                int32_t num_utts, float* out, uint32_t flags);

/* ----------------------------------------------------- feature pre-step --
 * Optional: the step immediately before the hot path in the reference's feature pipe
 * (extract_xvectors_new.sh:79), done on the device instead of by two extra processes:
 *   apply-cmvn-sliding --norm-vars=false --center=true --cmn-window=300   (per stream)
 *   select-voiced-frames ... vad.scp                                      (same frames for every stream)
 */
typedef struct cvx_prestep {
  int32_t cmn_window;     /* 0: no sliding CMN for this stream */
  int32_t cmn_center;     /* --center */
  int32_t min_cmn_window; /* --min-cmn-window (only used when center = 0) */
  int32_t norm_vars;      /* must be 0 */
} cvx_prestep;

/* cvx_extract on RAW features: raw_feats[s] [total_raw_rows, dim(s)] fp32 host rows of all
 * utterances; utterance u owns raw rows [utt_raw_start[u], utt_raw_start[u+1]); keep[] lists, per
 * utterance and in increasing order, the frame indices (inside the utterance) that survive VAD
 * selection -- kept rows of utterance u are [utt_kept_start[u], utt_kept_start[u+1]) of the
 * selected matrix (keep == NULL: keep every frame).  prestep[s] configures stream s.  Chunk arrays
 * are in selected-row coordinates, exactly as for cvx_extract. */
int cvx_extract_raw(cvx_engine* engine, const float* const* raw_feats, const cvx_prestep* prestep,
                    int32_t num_streams, int64_t total_raw_rows, const int64_t* utt_raw_start,
                    const int32_t* keep, const int64_t* utt_kept_start, int32_t num_utts,
                    const int64_t* chunk_start, const int32_t* chunk_len, const int32_t* chunk_utt,
                    const float* chunk_weight, int64_t num_chunks, float* out, uint32_t flags);

/* ------------------------------------------------- embedding post-step --
 * Optional: the step immediately after the hot path in the reference's back-end
 * (egs/sre/v5/run_sre10.sh:118-129), done on the device instead of by three more processes:
 *   ivector-subtract-global-mean mean.vec ... | transform-vec transform.mat ... |
 *   ivector-normalize-length ...
 * y = T (v - mean) (or T [v - mean; 1] when transform_cols == dim + 1), then
 * y *= sqrt(rows) / ||y||  (scaleup = 1, upstream's default) or 1 / ||y|| (scaleup = 0). */
typedef struct cvx_vector_chain {
  const float* mean;        /* [dim] host, or NULL: nothing subtracted */
  const float* transform;   /* [transform_rows, transform_cols] row-major host, or NULL */
  int32_t transform_rows, transform_cols;
  int32_t normalize_length; /* ivector-normalize-length --normalize */
  int32_t scaleup;          /* ivector-normalize-length --scaleup */
} cvx_vector_chain;

/* Installs (chain != NULL) or removes the post-step of an extraction engine: cvx_extract /
 * cvx_extract_raw then return [num_utts, cvx_engine_output_dim()] vectors that went through
 * the chain on the device (failed utterances stay all-zero).  mean / transform are copied. */
int cvx_engine_set_vector_chain(cvx_engine* engine, const cvx_vector_chain* chain);
int32_t cvx_engine_output_dim(const cvx_engine* engine);

/* The chain alone on host vectors in [num_vectors, dim] -> out [num_vectors, transform ?
 * transform_rows : dim]; ratios (may be NULL) receives per vector the factor
 * ivector-normalize-length averages in its log (||y|| / sqrt(dim) with scaleup). */
int cvx_transform_vectors(int32_t device, const float* in, int64_t num_vectors, int32_t dim,
                          const cvx_vector_chain* chain, float* out, float* ratios);

/* PLDA scoring, the last step of the back-end (`ivector-plda-scoring`, run_sre10.sh:122-129):
 * enrolment vectors train [num_train, dim_in] (speaker means over train_num_utts[i] utterances;
 * NULL = 1 each) and test vectors test [num_test, dim_in] are centred and projected by the model
 * (y = transform (x - mean)), length-normalised the PLDA way when normalize_length is set
 * (simple_length_norm: to sqrt(dim) instead), and every trial (trial_train[t], trial_test[t])
 * gets its log-likelihood ratio.  Double precision on the device, like upstream. */
typedef struct cvx_plda {
  const double* mean;      /* [dim_in] */
  const double* transform; /* [dim, dim_in] row-major: diagonalises the within- and between-class covariances */
  const double* psi;       /* [dim] between-class variances in the transformed space */
  int32_t dim_in, dim;
  int32_t normalize_length, simple_length_norm; /* --normalize-length, --simple-length-normalization */
} cvx_plda;
int cvx_plda_score(int32_t device, const cvx_plda* plda, const float* train, const int32_t* train_num_utts,
                   int32_t num_train, const float* test, int32_t num_test, const int32_t* trial_train,
                   const int32_t* trial_test, int64_t num_trials, float* scores);

/* ------------------------------------------------- frame-level outputs --
 * Replaces DecodableNnetSimple::GetOutput over a whole utterance as nnet3-compute uses it
 * [UPSTREAM-RECALL nnet3/nnet-am-decodable-simple.cc]: utterance u owns rows
 * [utt_start[u], utt_start[u+1]) of feats[s]; out [total_rows, embedding_dim] receives one output
 * row per input frame.  Frames the network needs beyond either end of an utterance are copies
 * of its first / last frame (upstream's edge rule).  CVX_FLAG_APPLY_EXP exponentiates the result
 * (nnet3-compute --apply-exp).  Only for engines of a cvx_model_load_frames model. */
#define CVX_FLAG_APPLY_EXP 4u
int cvx_compute_frames(cvx_engine* engine, const float* const* feats, int32_t num_streams,
                       int64_t total_rows, const int64_t* utt_start, int32_t num_utts, float* out,
                       uint32_t flags);

/* Test entry: the pre-step alone on one stream; out = [total_kept, dim] fp32 host. */
int cvx_test_prestep(cvx_engine* engine, const float* raw, int32_t dim, const cvx_prestep* prestep,
                     int64_t total_raw_rows, const int64_t* utt_raw_start, const int32_t* keep,
                     const int64_t* utt_kept_start, int32_t num_utts, float* out);

/* Per-chunk embeddings of the last cvx_extract call (before averaging),
 * [num_chunks, embedding_dim] copied to host memory. */
int cvx_engine_last_chunk_embeddings(cvx_engine* engine, float* out, int64_t num_chunks);

/* Pinned host memory for cvx_extract inputs (async H2D needs it). */
int cvx_host_alloc(void** ptr, size_t bytes);
void cvx_host_free(void* ptr);
int cvx_device_alloc(int32_t device, void** ptr, size_t bytes);
void cvx_device_free(int32_t device, void* ptr);
int cvx_memcpy_h2d(int32_t device, void* dst, const void* src, size_t bytes);
int cvx_device_count(int32_t* count);

/* ------------------------------------------------------- job fan-out --
 * The recipe starts `nj` processes per stage with --use-gpu=yes and no device hint
 * (extract_xvectors_new.sh:83-89); upstream picks a device per process in CuDevice::SelectGpuId.
 * cvx_acquire_device gives every process of a box its own GPU: it takes an advisory lock file
 * (flock, $CVX_LOCK_DIR or /tmp, named by PCI bus id, released when the process exits) on the
 * first device nobody holds; when all devices are held the jobs spread evenly and *sharers
 * (may be NULL) is the number of jobs already on the chosen device.  wait != 0 (--use-gpu=wait):
 * block until a device is free instead of sharing. */
int cvx_acquire_device(int32_t wait, int32_t* device, int32_t* sharers);
/* The same protocol for a host that wants its device BEFORE it initialises CUDA (so that it can restrict
 * CUDA_VISIBLE_DEVICES to it and not pay context set-up on every GPU of the box): ids[] are stable
 * device names, e.g. the PCI bus ids under /proc/driver/nvidia/gpus; *index = the one acquired.  Names
 * are matched case-insensitively with cvx_acquire_device's, so both kinds of job share one set of locks. */
int cvx_acquire_device_by_id(const char* const* ids, int32_t num_ids, int32_t wait, int32_t* index, int32_t* sharers);
/* Test entry: the same lock protocol over `num_devices` made-up devices in directory `dir`. */
int cvx_test_acquire_lock(const char* dir, int32_t num_devices, int32_t wait, int32_t* device, int32_t* sharers);

/* Length-balanced utterance shards (replaces utils/split_data.sh for this path,
 * extract_xvectors_new.sh:72-74): greedy longest-processing-time over frame counts -- utterances
 * sorted by length (descending, ties in input order) go to the currently lightest shard (ties: the
 * lowest shard index).  shard_of[u] receives the 0-based shard of utterance u.  Deterministic:
 * every job computes the same partition from utt2num_frames without communicating. */
int cvx_lpt_shards(const int64_t* lengths, int64_t num_utts, int32_t num_shards, int32_t* shard_of);

/* Profile of the calls since the last reset (opts.profile = 1): one record per
 * kernel class. */
typedef struct cvx_kernel_profile {
  char name[48];
  int64_t launches;
  double total_ms;     /* CUDA-event time on the engine's compute stream */
  double flops;        /* algorithmic FLOPs of those launches (0 if n/a) */
  double bytes;        /* algorithmic HBM bytes of those launches (0 if n/a) */
} cvx_kernel_profile;

int cvx_engine_get_profile(cvx_engine* engine, cvx_kernel_profile* records, int32_t max_records,
                           int32_t* num_records);
/* Range check mode (options.reserved[1]): counts[0] = staged input values, counts[1 + l] = outputs of frame
 * layer l that were clipped at the fp16 range limit since the engine was created.  The conversions
 * saturate (cvt.rn.satfinite), so without this check a model whose activations exceed 65504 -- un-normalised
 * features, a missing BatchNorm -- would be clipped silently.  *num_counts = 0 when the mode is off. */
int cvx_engine_saturation_counts(cvx_engine* engine, int64_t* counts, int32_t max_counts, int32_t* num_counts);
void cvx_engine_reset_profile(cvx_engine* engine);
/* Kernels launched by this engine since creation (all classes). */
int64_t cvx_engine_launch_count(const cvx_engine* engine);
/* Device time of the last cvx_extract call: CUDA events on the engine's compute stream
 * bracketing all of its batches (H2D copies of a host-input call are waited on by that
 * stream, so they are inside the bracket). */
double cvx_engine_last_extract_ms(const cvx_engine* engine);

/* ---------------------------------------------------- kernel unit entries --
 * Single-kernel entry points used by tests/ to check each kernel in isolation
 * against the oracle (device pointers, default stream, synchronous). */

/* out[r, n] = post(sum_taps A[r+off_tap, :] W_tap^T): one TDNN layer on fp32 host data;
 * converts to `precision`, runs the production kernel, returns fp32 [rows, out_dim].
 * The input is treated like a previous layer's output (one 64-column K block walk per
 * tap, TMA row offsets); `precision | 0x100` selects the compact first-layer splice. */
int cvx_test_tdnn_layer(int32_t device, int32_t precision, const float* in, int64_t rows,
                        int32_t in_dim, const int32_t* offsets, int32_t num_offsets,
                        const float* weight, const float* bias, int32_t relu, const float* scale,
                        const float* offset, int32_t out_dim, float* out);

/* Average device time (ms, CUDA events) of the production kernel for one such layer on
 * `rows` rows of pseudo-random data, `iters` timed launches after warm-up. */
int cvx_test_tdnn_layer_bench(int32_t device, int32_t precision, int64_t rows, int32_t in_dim,
                              const int32_t* offsets, int32_t num_offsets, int32_t out_dim,
                              int32_t iters, double* avg_ms);

/* One segment-level layer (tdnn6: [chunks, 3000] x [3000, 512]) on fp32 host data through the production path:
 * out[m, n] = post(in[m, k] w[n, k]^T), post = (+bias, optional ReLU, optional scale/offset).  With 16-bit
 * `precision` that is the hi/lo-split tensor-core GEMM (fp32-accurate, ~2^-22); with CVX_PRECISION_FP32 the
 * CUDA-core kernel. */
int cvx_test_segment_affine(int32_t device, int32_t precision, const float* in, int64_t m, int32_t k,
                            const float* w, const float* bias, int32_t relu, const float* scale,
                            const float* offset, int32_t n, float* out);

/* mean|stddev over row ranges [begin[c], end[c]) of an fp32 host matrix [rows, dim]. */
int cvx_test_stats_pool(int32_t device, int32_t precision, const float* in, int64_t rows,
                        int32_t dim, const int64_t* begin, const int64_t* end, int64_t num_chunks,
                        float variance_floor, float* out /* [num_chunks, 2*dim] */);

#ifdef __cplusplus
}
#endif
#endif /* CVX_H_ */
