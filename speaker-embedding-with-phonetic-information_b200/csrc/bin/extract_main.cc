// Drop-in command-line driver: same argv / rspecifier / wspecifier / exit-code
// contract as upstream Kaldi's nnet3-xvector-compute, as invoked by
//   egs/sre/v5/sid/nnet3/xvector/extract_xvectors_new.sh:84-93
//     nnet3-xvector-compute --use-gpu=no --min-chunk-size=$min --chunk-size=$chunk "$nnet" "$feat" ark,scp:A,B
// and the two-stream variant invoked by
//   egs/sre/v5/sid/nnet3_cvector/cvector/extract_cvectors_with_am.sh:99-110
//     nnet3-cvector-compute-multiple-input ... --name="$input_name" 2 "$nnet" "$feat" "$feat_am" ark,scp:A,B
// (that binary is referenced but not shipped by the reference; semantics = the
// single-input program with N named inputs read in lock-step).
//
// Chunking / averaging follows nnet3-xvector-compute.cc main();
// the forward pass itself goes through the C ABI (include/cvx.h) to the CUDA side.
// Utterances are packed into row batches; a worker thread per GPU runs cvx_extract on full batches
// while the reader keeps filling a free one.  One process can drive several GPUs (--num-gpus=G):
// with scp inputs every GPU gets its own reader thread over a length-balanced share of the table
// ("lanes": nothing is shared between them, the main thread merges their results back into table
// order); with archive / pipe inputs (one sequential stream) the single reader hands batches to
// the GPUs in turn.  --shard=J/N restricts the job to the J-th of N length-balanced shards of the
// table (replaces utils/split_data.sh, extract_xvectors_new.sh:72-74).
#include "extract_main.h"

#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <memory>
#include <mutex>
#include <thread>
#include <unordered_map>

#include <unistd.h>

#include "../../../include/cvx.h"
#include "../kaldi_io.h"
#include "job_util.h"

namespace cvx {
namespace {

struct HostBatch {
  std::vector<float*> feats;        // pinned, per stream
  std::vector<size_t> cap_rows;     // per stream capacity in rows
  std::vector<int32_t> dims;
  int64_t rows = 0;
  std::vector<std::string> keys;
  std::vector<int64_t> chunk_start;
  std::vector<int32_t> chunk_len, chunk_utt;
  std::vector<float> chunk_weight;
  float* out = nullptr;             // pinned: the embeddings come back with an asynchronous copy
  size_t out_cap = 0;
  int64_t frames = 0;  // un-padded input frames of the successful utterances
  // raw-feature mode (device pre-step): rows are raw frames, chunks index the selected frames
  int64_t kept_rows = 0;
  std::vector<int64_t> utt_raw_start{0}, utt_kept_start{0};
  std::vector<int32_t> keep;

  void Reset() {
    rows = 0;
    frames = 0;
    keys.clear();
    chunk_start.clear();
    chunk_len.clear();
    chunk_utt.clear();
    chunk_weight.clear();
    kept_rows = 0;
    utt_raw_start.assign(1, 0);
    utt_kept_start.assign(1, 0);
    keep.clear();
  }
  void Reserve(int64_t need_rows) {
    for (size_t s = 0; s < feats.size(); ++s) {
      if (static_cast<size_t>(need_rows) <= cap_rows[s]) continue;
      size_t ncap = std::max<size_t>(static_cast<size_t>(need_rows) * 5 / 4, 1 << 16);
      void* p = nullptr;
      if (cvx_host_alloc(&p, ncap * dims[s] * sizeof(float)) != CVX_OK) throw KaldiError(cvx_last_error());
      if (rows) std::memcpy(p, feats[s], static_cast<size_t>(rows) * dims[s] * sizeof(float));
      cvx_host_free(feats[s]);
      feats[s] = static_cast<float*>(p);
      cap_rows[s] = ncap;
    }
  }
  void ReserveOut(size_t n) {
    if (n <= out_cap) return;
    cvx_host_free(out);
    out = nullptr;
    out_cap = 0;
    void* p = nullptr;
    const size_t ncap = std::max<size_t>(n * 5 / 4, 1 << 16);
    if (cvx_host_alloc(&p, ncap * sizeof(float)) != CVX_OK) throw KaldiError(cvx_last_error());
    out = static_cast<float*>(p);
    out_cap = ncap;
  }
  void AppendRow(const std::vector<const float*>& src_rows) {
    for (size_t s = 0; s < feats.size(); ++s)
      std::memcpy(feats[s] + static_cast<size_t>(rows) * dims[s], src_rows[s], dims[s] * sizeof(float));
    ++rows;
  }
  ~HostBatch() {
    for (float* p : feats) cvx_host_free(p);
    cvx_host_free(out);
  }
};

// Pins host batches on a helper thread, first batch first: pinning ~100 MB takes longer than
// building an engine, and a job of less than one batch never waits for the second.
struct BatchPool {
  std::vector<HostBatch>& hb;
  int64_t rows;
  std::mutex mu;
  std::condition_variable cv;
  int ready = 0;
  std::string error;
  std::thread th;
  BatchPool(std::vector<HostBatch>& h, int64_t r) : hb(h), rows(r) {
    th = std::thread([this] {
      for (size_t i = 0; i < hb.size(); ++i) {
        try {
          hb[i].Reserve(rows);
        } catch (const std::exception& e) {
          std::lock_guard<std::mutex> lk(mu);
          error = e.what();
          ready = static_cast<int>(hb.size());
          cv.notify_all();
          return;
        }
        std::lock_guard<std::mutex> lk(mu);
        ready = static_cast<int>(i) + 1;
        cv.notify_all();
      }
    });
  }
  void WaitFor(int i) {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] { return ready > i; });
    if (!error.empty()) throw KaldiError("allocating pinned host memory: " + error);
  }
  ~BatchPool() {
    if (th.joinable()) th.join();
  }
};

// Results of one lane, consumed in order by the merging writer.
struct ResultQueue {
  std::mutex mu;
  std::condition_variable cv;
  std::deque<std::pair<std::string, std::vector<float>>> q;
  bool done = false;
  void Push(const std::string& key, const float* v, int32_t n) {
    std::lock_guard<std::mutex> lk(mu);
    q.emplace_back(key, std::vector<float>(v, v + n));
    cv.notify_all();
  }
  void Finish() {
    std::lock_guard<std::mutex> lk(mu);
    done = true;
    cv.notify_all();
  }
};

struct LaneResult {
  int64_t num_success = 0, num_fail = 0, frame_count = 0;
  std::string error;
  // where the lane's wall time went (logged with --verbose=1)
  double setup_s = 0, gpu_wait_s = 0, extract_s = 0, device_s = 0, total_s = 0, first_call_s = 0, max_call_s = 0;
  int64_t calls = 0;
  std::chrono::steady_clock::time_point end{};  // last result handed over (before the lane frees its pinned buffers)
};

uint32_t KeyHash(const std::string& s) {  // FNV-1a: the same in every job of a stage
  uint32_t h = 2166136261u;
  for (unsigned char c : s) h = (h ^ c) * 16777619u;
  return h;
}

}  // namespace

std::string ReadOutputNodeFromConfig(const std::string& path) {
  // accepts the one-line extract.config written by run_xvector_new.sh:123 /
  // extract_xvectors_new.sh:58: "output-node name=output input=<descriptor>"
  std::ifstream f(path);
  if (!f) throw KaldiError("cannot open --nnet-config file " + path);
  std::string line;
  while (std::getline(f, line)) {
    size_t h = line.find('#');
    if (h != std::string::npos) line = line.substr(0, h);
    if (line.find("output-node") == std::string::npos) continue;
    if (line.find("name=output ") == std::string::npos && line.find("name=output\t") == std::string::npos) continue;
    size_t p = line.find("input=");
    if (p == std::string::npos) continue;
    std::string v = line.substr(p + 6);
    size_t q = v.find(" objective=");
    if (q != std::string::npos) v = v.substr(0, q);
    while (!v.empty() && std::isspace(static_cast<unsigned char>(v.back()))) v.pop_back();
    return v;
  }
  throw KaldiError("no 'output-node name=output input=...' line in " + path);
}

int ExtractMain(int argc, const char* const* argv, bool multiple_input) {
  try {
    const char* usage_single =
        "Computes one embedding (x-vector / c-vector) per utterance: runs the frame-level\n"
        "TDNN layers, the statistics pooling and the segment-level layers of an nnet3 raw\n"
        "model up to output-node 'output' on the GPU.  With --chunk-size, long utterances\n"
        "are cut into chunks, one embedding is computed per chunk and the chunk embeddings\n"
        "are averaged weighted by their frame counts.\n"
        "\n"
        "Usage: nnet3-xvector-compute [options] <raw-nnet-in> "
        "<features-rspecifier> <vector-wspecifier>\n"
        "e.g.: nnet3-xvector-compute final.raw scp:feats.scp "
        "ark:nnet_prediction.ark\n"
        "See also: nnet3-compute\n";
    const char* usage_multi =
        "Like nnet3-xvector-compute, for networks with several input nodes\n"
        "(e.g. speaker features + phonetic/AM features).  The i-th name in --name\n"
        "is the input-node fed by the i-th feature rspecifier; all streams are read\n"
        "in lock-step and must have the same keys and row counts.\n"
        "\n"
        "Usage: nnet3-cvector-compute-multiple-input [options] --name=n1,n2 <num-inputs> "
        "<raw-nnet-in> <features-rspecifier-1> ... <features-rspecifier-n> <vector-wspecifier>\n";
    ParseOptions po(multiple_input ? usage_multi : usage_single);

    std::string use_gpu = "yes", output_node, nnet_config, precision = "fp16", names, vad_rspecifier;
    std::string shard_spec, utt2num_frames;
    std::string mean_vec, transform_mat;
    bool normalize_length = false, scaleup = true;
    int32_t chunk_size = -1, min_chunk_size = 100, gpu_id = -1, cmn_window = 0, min_cmn_window = 100, num_gpus = 1;

    bool cmn_center = true, cmn_norm_vars = false;
    int64_t batch_rows = 1 << 20, max_batch_rows = 0;
    bool pad_input = false, profile = false, check_saturation = false;
    // accepted for command-line compatibility with nnet3 programs; no effect here
    int32_t ign_i[4] = {0, 0, 0, 0};
    float ign_f = 1.0f;
    bool ign_b[2] = {false, false};
    std::string ign_s;

    po.Register("use-gpu", &use_gpu, "yes|no|optional|wait, accepted for compatibility; this program always runs on a GPU");
    po.Register("chunk-size", &chunk_size,
                "Frames per chunk; embeddings of the chunks are averaged. -1 = one chunk per utterance.");
    po.Register("min-chunk-size", &min_chunk_size, "Shortest chunk (in frames) an embedding is computed from.");
    po.Register("pad-input", &pad_input,
                "If true, duplicate the first and last frames of the input features as required to equal min-chunk-size "
                "(newer Kaldi default); if false, chunks shorter than min-chunk-size are skipped (older Kaldi).");
    po.Register("output-node", &output_node,
                "Descriptor to use as output-node 'output' (same effect as piping the model through "
                "nnet3-copy --nnet-config=extract.config)");
    po.Register("nnet-config", &nnet_config, "File with an 'output-node name=output input=...' line (extract.config)");
    po.Register("precision", &precision, "TDNN operand type: fp16 | bf16 | fp32 (fp32 = slow validation kernels)");
    po.Register("gpu-id", &gpu_id,
                "CUDA device ordinal (-1: $CVX_DEVICE, else a free GPU is picked and held through a lock file, so that "
                "the nj jobs of a recipe stage spread over the GPUs of the box)");
    po.Register("num-gpus", &num_gpus,
                "Use this many GPUs (devices gpu-id .. gpu-id + num-gpus - 1) from ONE process, one engine per GPU. "
                "scp inputs: one reader thread per GPU over a length-balanced share of the table; archive / pipe "
                "inputs: row batches go to the GPUs in turn.  Output order is the table's either way");
    po.Register("shard", &shard_spec,
                "J/N: process only the J-th (1-based, like JOB) of N length-balanced shards of the feature table "
                "(greedy LPT over frame counts; replaces utils/split_data.sh)");
    po.Register("utt2num-frames", &utt2num_frames,
                "utt2num_frames file of the data dir: frame counts for --shard / --num-gpus balancing (scp inputs can do "
                "without: the counts are then read from the matrix headers)");
    po.Register("batch-rows", &batch_rows, "Frames read from the feature pipe(s) per cvx_extract call");
    po.Register("max-batch-rows", &max_batch_rows, "Frames per packed device batch (0 = engine default)");
    po.Register("profile", &profile, "Print per-kernel CUDA-event times at the end");
    po.Register("check-saturation", &check_saturation,
                "Validation mode for a new model (fp16 operands): store every layer and report activations that were clipped "
                "at the fp16 range limit; slower (no fused pooling)");
    po.Register("vad-rspecifier", &vad_rspecifier,
                "If set, the features are RAW: keep only frames whose VAD value is non-zero (what piping through "
                "select-voiced-frames does), on the GPU");
    po.Register("cmn-window", &cmn_window,
                "If > 0, the features are RAW: apply sliding-window CMN with this window to the first stream before "
                "frame selection (what piping through apply-cmvn-sliding does), on the GPU");
    po.Register("center", &cmn_center, "apply-cmvn-sliding --center");
    po.Register("min-cmn-window", &min_cmn_window, "apply-cmvn-sliding --min-cmn-window (only if --center=false)");
    po.Register("norm-vars", &cmn_norm_vars, "apply-cmvn-sliding --norm-vars (only false is supported)");
    po.Register("mean-vec", &mean_vec,
                "If set, subtract this vector from every embedding on the GPU (what piping the output through "
                "ivector-subtract-global-mean <mean> does)");
    po.Register("transform-mat", &transform_mat,
                "If set, multiply every embedding by this matrix on the GPU, after --mean-vec (transform-vec, e.g. the LDA)");
    po.Register("normalize-length", &normalize_length,
                "If true, scale every embedding to length sqrt(dim) on the GPU, last (ivector-normalize-length)");
    po.Register("scaleup", &scaleup, "ivector-normalize-length --scaleup");
    if (multiple_input) po.Register("name", &names, "Comma-separated input-node names, one per feature stream");
    po.Register("frame-subsampling-factor", &ign_i[0], "ignored");
    po.Register("extra-left-context", &ign_i[1], "ignored");
    po.Register("extra-right-context", &ign_i[2], "ignored");
    po.Register("frames-per-chunk", &ign_i[3], "ignored");
    po.Register("acoustic-scale", &ign_f, "ignored");
    po.Register("debug-computation", &ign_b[0], "ignored");
    po.Register("optimization.optimize", &ign_b[1], "ignored");
    po.Register("cache-capacity", &ign_i[0], "ignored");

    po.Read(argc, argv);
    int num_inputs = 1;
    int arg = 1;
    if (multiple_input) {
      if (po.NumArgs() < 4) {
        po.PrintUsage();
        return 1;
      }
      num_inputs = std::atoi(po.GetArg(arg++).c_str());
      if (num_inputs < 1 || po.NumArgs() != num_inputs + 3) {
        po.PrintUsage();
        return 1;
      }
    } else if (po.NumArgs() != 3) {
      po.PrintUsage();
      return 1;
    }
    const std::string nnet_rxfilename = po.GetArg(arg++);
    std::vector<std::string> feat_rspecifiers;
    for (int i = 0; i < num_inputs; ++i) feat_rspecifiers.push_back(po.GetArg(arg++));
    const std::string vector_wspecifier = po.GetArg(arg++);

    if (use_gpu != "yes" && use_gpu != "no" && use_gpu != "optional" && use_gpu != "wait")
      throw KaldiError("invalid value for --use-gpu: " + use_gpu);
    if (use_gpu == "no") CVX_LOG("--use-gpu=no requested; this implementation has no CPU path and runs on the GPU");
    if (!nnet_config.empty()) output_node = ReadOutputNodeFromConfig(nnet_config);
    const bool raw_mode = !vad_rspecifier.empty() || cmn_window > 0;
    if (raw_mode && pad_input) throw KaldiError("--pad-input=true is not supported together with --vad-rspecifier / --cmn-window");
    if (cmn_norm_vars) throw KaldiError("--norm-vars=true is not supported (the recipes use false)");
    int32_t prec;
    if (precision == "fp16") prec = CVX_PRECISION_FP16;
    else if (precision == "bf16") prec = CVX_PRECISION_BF16;
    else if (precision == "fp32") prec = CVX_PRECISION_FP32;
    else throw KaldiError("invalid --precision " + precision);
    if (pad_input && chunk_size != -1 && chunk_size < min_chunk_size)
      CVX_WARN("--chunk-size=" << chunk_size << " is below --min-chunk-size=" << min_chunk_size
                               << ": with --pad-input=true every chunk is padded to the minimum size");
    if (num_gpus < 1) throw KaldiError("--num-gpus must be >= 1");
    const std::pair<int32_t, int32_t> shard = ParseShard(shard_spec);
    const auto t_proc = std::chrono::steady_clock::now();
    auto since_start = [&]() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t_proc).count(); };
    gpu_id = ChooseDevice(gpu_id, use_gpu, num_gpus);
    const double t_device = since_start();

    cvx_model* model = nullptr;
    if (cvx_model_load(nnet_rxfilename.c_str(), output_node.empty() ? nullptr : output_node.c_str(), &model) != CVX_OK)
      throw KaldiError(std::string("reading ") + nnet_rxfilename + ": " + cvx_last_error());
    const double t_model = since_start();
    cvx_model_info info;
    cvx_model_get_info(model, &info);
    if (po.verbose >= 1) CVX_LOG("extraction plan:\n" << cvx_model_describe(model));
    if (info.num_inputs != num_inputs)
      throw KaldiError("the network reads " + std::to_string(info.num_inputs) + " input node(s) but " +
                       std::to_string(num_inputs) + " feature stream(s) were given");
    // stream order: --name maps rspecifier i to input node name[i]
    std::vector<int> stream_of_arg(num_inputs);
    for (int i = 0; i < num_inputs; ++i) stream_of_arg[i] = i;
    if (multiple_input && !names.empty()) {
      std::vector<std::string> nm;
      size_t p = 0;
      while (p <= names.size()) {
        size_t q = names.find(',', p);
        if (q == std::string::npos) q = names.size();
        nm.push_back(names.substr(p, q - p));
        p = q + 1;
      }
      if (static_cast<int>(nm.size()) != num_inputs) throw KaldiError("--name must list " + std::to_string(num_inputs) + " names");
      for (int i = 0; i < num_inputs; ++i) {
        int found = -1;
        for (int s = 0; s < info.num_inputs; ++s)
          if (nm[i] == cvx_model_input_name(model, s)) found = s;
        if (found < 0) throw KaldiError("the network has no input-node named '" + nm[i] + "'");
        stream_of_arg[i] = found;
      }
    }

    cvx_engine_options eo;
    cvx_engine_options_default(&eo);
    eo.device = gpu_id;
    eo.precision = prec;
    eo.max_batch_rows = max_batch_rows;
    eo.profile = profile ? 1 : 0;
    eo.reserved[1] = check_saturation ? 1 : 0;
    int32_t ndev = 0;
    if (cvx_device_count(&ndev) != CVX_OK) throw KaldiError(std::string("no usable CUDA device: ") + cvx_last_error());
    // test aid: CVX_WRAP_DEVICES=1 lets --num-gpus exceed the visible devices (engine g runs on device
    // (gpu-id + g) mod ndev), so the multi-engine / multi-lane paths can be exercised on a one-GPU box
    const bool wrap_devices = std::getenv("CVX_WRAP_DEVICES") != nullptr && ndev > 0;
    if (gpu_id + num_gpus > ndev && !wrap_devices)
      throw KaldiError("--gpu-id=" + std::to_string(gpu_id) + " --num-gpus=" + std::to_string(num_gpus) + " but only " +
                       std::to_string(ndev) + " CUDA device(s) are visible");

    // ---- work partition: this job's shard, and inside it one lane per GPU when the tables are scp ----
    std::vector<TableSpec> feat_specs;
    bool scp_input = true;
    for (const std::string& r : feat_rspecifiers) {
      feat_specs.push_back(ParseRspecifier(r));
      scp_input = scp_input && feat_specs.back().is_scp;
    }
    if (!vad_rspecifier.empty()) scp_input = scp_input && ParseRspecifier(vad_rspecifier).is_scp;
    // (More than one lane per GPU was tried and dropped: a second engine's pinned-buffer allocations and
    // device-wide frees stall the first one's CUDA calls for hundreds of ms.)
    const int num_lanes = (scp_input && num_gpus > 1) ? num_gpus : 1;
    const int num_engines = num_gpus;
    const int32_t num_bins = shard.second * num_lanes;
    std::unordered_map<std::string, int32_t> bin_of;
    if (num_bins > 1) {
      bin_of = AssignShards(utt2num_frames, feat_specs[0].is_scp ? feat_specs[0].scp_name : std::string(), num_bins, nullptr);
      CVX_LOG("shard " << (shard.first + 1) << "/" << shard.second << ": " << num_lanes << " reader lane(s), "
                       << bin_of.size() << " utterances balanced by frame count over " << num_bins << " bins");
    }
    // bin of a key; keys without a frame count (not in utt2num_frames) are spread by hash
    auto key_bin = [&](const std::string& key) -> int32_t {
      auto it = bin_of.find(key);
      return it != bin_of.end() ? it->second : static_cast<int32_t>(KeyHash(key) % static_cast<uint32_t>(num_bins));
    };

    // one engine per GPU (SURVEY 8e mode ii); engines share nothing, there is no collective.  Built
    // concurrently: each is ~0.3 s of weight re-packing and uploads.
    std::vector<cvx_engine*> engines(num_engines, nullptr);
    {
      std::vector<std::string> errs(num_engines);
      std::vector<std::thread> th;
      for (int g = 0; g < num_engines; ++g)
        th.emplace_back([&, g] {
          cvx_engine_options o = eo;
          o.device = wrap_devices ? (gpu_id + g) % ndev : gpu_id + g;
          if (cvx_engine_create(model, &o, &engines[g]) != CVX_OK) errs[g] = cvx_last_error();
        });
      for (std::thread& t : th) t.join();
      for (const std::string& e : errs)
        if (!e.empty()) throw KaldiError("creating the GPU engine: " + e);
    }
    cvx_engine* engine = engines[0];
    if (po.verbose >= 1)
      CVX_LOG("start-up: device selection " << t_device << " s, model load " << (t_model - t_device) << " s, CUDA context + "
                                             << num_engines << " engine(s) " << (since_start() - t_model) << " s");

    if (!mean_vec.empty() || !transform_mat.empty() || normalize_length) {
      // the back-end's per-vector chain, applied before the embeddings leave the device
      std::vector<float> mean;
      Matrix tf;
      cvx_vector_chain chain;
      std::memset(&chain, 0, sizeof(chain));
      if (!mean_vec.empty()) {
        mean = ReadVectorFile(mean_vec);
        if (static_cast<int32_t>(mean.size()) != info.embedding_dim)
          throw KaldiError("--mean-vec has dim " + std::to_string(mean.size()) + ", the embeddings have dim " +
                           std::to_string(info.embedding_dim));
        chain.mean = mean.data();
      }
      if (!transform_mat.empty()) {
        tf = ReadMatrixFile(transform_mat);
        chain.transform = tf.Data();
        chain.transform_rows = tf.rows;
        chain.transform_cols = tf.cols;
      }
      chain.normalize_length = normalize_length ? 1 : 0;
      chain.scaleup = scaleup ? 1 : 0;
      for (cvx_engine* e : engines)
        if (cvx_engine_set_vector_chain(e, &chain) != CVX_OK)
          throw KaldiError(std::string("installing the vector post-step: ") + cvx_last_error());
    }

    std::vector<cvx_prestep> prestep(num_inputs);
    for (int s = 0; s < num_inputs; ++s) prestep[s] = cvx_prestep{s == 0 ? cmn_window : 0, cmn_center ? 1 : 0, min_cmn_window, 0};
    const int32_t E = cvx_engine_output_dim(engine);
    const int32_t ctx = info.left_context + info.right_context + 1;
    auto t_start = std::chrono::steady_clock::now();

    // ---- one lane: reader(s) -> pinned row batches -> cvx_extract on `engs` (in turn) -> emit, in table order ----
    auto run_lane = [&](const std::vector<cvx_engine*>& engs, const std::function<bool(const std::string&)>& keep,
                        const std::function<void(const std::string&, const float*)>& emit, LaneResult* res) {
      const int ngpu = static_cast<int>(engs.size());
      const auto lane_t0 = std::chrono::steady_clock::now();
      auto since = [](std::chrono::steady_clock::time_point t) {
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - t).count();
      };
      // pinned host batches: one being filled + one in flight per GPU of this lane
      std::vector<HostBatch> hb(ngpu + 1);
      for (HostBatch& b : hb) {
        b.feats.assign(num_inputs, nullptr);
        b.cap_rows.assign(num_inputs, 0);
        b.dims.resize(num_inputs);
        for (int s = 0; s < num_inputs; ++s) b.dims[s] = cvx_model_input_dim(model, s);
      }
      BatchPool batch_pool(hb, batch_rows + batch_rows / 8 + 65536);
      std::vector<std::unique_ptr<SequentialMatrixReader>> readers;
      for (int i = 0; i < num_inputs; ++i) readers.emplace_back(new SequentialMatrixReader(feat_rspecifiers[i], false, keep));
      std::unique_ptr<SequentialMatrixReader> vad_reader;
      if (!vad_rspecifier.empty()) vad_reader.reset(new SequentialMatrixReader(vad_rspecifier, /*allow_vectors=*/true, keep));
      batch_pool.WaitFor(0);
      res->setup_s = since(lane_t0);
      int64_t& num_success = res->num_success;
      int64_t& num_fail = res->num_fail;
      int64_t& frame_count = res->frame_count;

      // Jobs finish in launch order (outputs keep the input order); job k runs on GPU k mod ngpu.
      struct Job {
        int batch = -1;
        std::thread th;
        int status = CVX_OK;
        std::string error;
        double seconds = 0;  // inside cvx_extract
        double device_ms = 0;
      };
      std::deque<std::unique_ptr<Job>> inflight;
      struct Joiner {
        std::deque<std::unique_ptr<Job>>& q;
        ~Joiner() {
          for (auto& j : q)
            if (j->th.joinable()) j->th.join();
        }
      } joiner{inflight};
      std::vector<int> free_batches;
      for (int i = 1; i <= ngpu; ++i) free_batches.push_back(i);
      int64_t launch_seq = 0;
      auto finish_oldest = [&]() {
        if (inflight.empty()) return;
        std::unique_ptr<Job> job = std::move(inflight.front());
        inflight.pop_front();
        const auto tw = std::chrono::steady_clock::now();
        job->th.join();
        res->gpu_wait_s += since(tw);
        res->extract_s += job->seconds;
        res->device_s += job->device_ms * 1e-3;
        if (res->calls == 0) res->first_call_s = job->seconds;
        res->max_call_s = std::max(res->max_call_s, job->seconds);
        res->calls += 1;
        if (job->status != CVX_OK) throw KaldiError("cvx_extract failed: " + job->error);
        HostBatch& b = hb[job->batch];
        for (size_t u = 0; u < b.keys.size(); ++u) emit(b.keys[u], b.out + u * E);
        num_success += static_cast<int64_t>(b.keys.size());
        frame_count += b.frames;
        b.Reset();
        free_batches.push_back(job->batch);
      };
      // Hands batch `which` to the next GPU and returns the index of a free batch to fill next.
      auto launch = [&](int which) -> int {
        HostBatch& b = hb[which];
        if (b.keys.empty()) return which;
        // the GPU this job goes to is the one the oldest job in flight runs on
        if (static_cast<int>(inflight.size()) == ngpu) finish_oldest();
        b.ReserveOut(b.keys.size() * static_cast<size_t>(E));
        cvx_engine* eng = engs[launch_seq++ % ngpu];
        std::unique_ptr<Job> job(new Job);
        job->batch = which;
        Job* jp = job.get();
        job->th = std::thread([&, which, eng, jp]() {
          const auto tj = std::chrono::steady_clock::now();
          HostBatch& w = hb[which];
          std::vector<const float*> f(w.feats.begin(), w.feats.end());
          if (raw_mode)
            jp->status = cvx_extract_raw(eng, f.data(), prestep.data(), num_inputs, w.rows, w.utt_raw_start.data(),
                                         vad_reader ? w.keep.data() : nullptr, w.utt_kept_start.data(),
                                         static_cast<int32_t>(w.keys.size()), w.chunk_start.data(), w.chunk_len.data(),
                                         w.chunk_utt.data(), w.chunk_weight.data(),
                                         static_cast<int64_t>(w.chunk_start.size()), w.out, 0);
          else
            jp->status = cvx_extract(eng, f.data(), num_inputs, w.rows, w.chunk_start.data(), w.chunk_len.data(),
                                     w.chunk_utt.data(), w.chunk_weight.data(), static_cast<int64_t>(w.chunk_start.size()),
                                     static_cast<int32_t>(w.keys.size()), w.out, 0);
          if (jp->status != CVX_OK) jp->error = cvx_last_error();
          jp->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - tj).count();
          jp->device_ms = cvx_engine_last_extract_ms(eng);
        });
        inflight.push_back(std::move(job));
        const int next = free_batches.back();
        free_batches.pop_back();
        batch_pool.WaitFor(next);
        return next;
      };

      int cur = 0;
      if (num_inputs == 1 && !pad_input) {
        // binary float matrices are read straight into the pinned batch at its tail
        readers[0]->Value().placer = [&](int32_t rows, int32_t cols) -> float* {
          HostBatch& b = hb[cur];
          if (cols != b.dims[0]) return nullptr;
          b.Reserve(b.rows + rows + std::max(min_chunk_size, 0));
          return b.feats[0] + static_cast<size_t>(b.rows) * b.dims[0];
        };
      }
      std::vector<const Matrix*> mats(num_inputs);
      std::vector<const float*> row_ptrs(num_inputs);
      for (; !readers[0]->Done(); ) {
        const std::string utt = readers[0]->Key();
        bool ok = true;
        for (int i = 0; i < num_inputs; ++i) {
          if (readers[i]->Done() || readers[i]->Key() != utt)
            throw KaldiError("feature streams are not in lock-step at utterance " + utt);
          mats[stream_of_arg[i]] = &readers[i]->Value();
        }
        const int32_t Traw = mats[0]->rows;
        // raw mode: select-voiced-frames' rule (frames with VAD != 0, in order).  The VAD table is looked
        // up by key like the random-access table of the pipe it replaces (scp,s,cs: sorted keys): entries
        // for other utterances are passed over, a missing one fails this utterance only.
        std::vector<int32_t> utt_keep;
        if (vad_reader) {
          while (!vad_reader->Done() && vad_reader->Key() < utt) vad_reader->Next();
          if (vad_reader->Done() || vad_reader->Key() != utt) {
            CVX_WARN("No VAD entry for utterance " << utt << " in " << vad_rspecifier);
            ok = false;
          } else {
            const Matrix& v = vad_reader->Value();
            const int64_t n = static_cast<int64_t>(v.rows) * v.cols;
            if (n != Traw) {
              CVX_WARN("Mismatch in number for frames " << Traw << " for features and VAD " << n << ", for utterance " << utt);
              ok = false;
            } else {
              for (int32_t i = 0; i < Traw; ++i)
                if (v.Data()[i] != 0.0f) utt_keep.push_back(i);
              if (utt_keep.empty()) {
                CVX_WARN("No features were judged as voiced for utterance " << utt);
                ok = false;
              }
            }
          }
        }
        const int32_t T = vad_reader ? static_cast<int32_t>(utt_keep.size()) : Traw;  // frames the chunks are cut from
        for (int s = 0; s < num_inputs; ++s) {
          if (mats[s]->rows != Traw) {
            CVX_WARN("Feature streams of " << utt << " have different numbers of rows (" << Traw << " vs " << mats[s]->rows << ")");
            ok = false;
          }
          if (Traw > 0 && mats[s]->cols != hb[cur].dims[s])
            throw KaldiError("feature dimension " + std::to_string(mats[s]->cols) + " of " + utt + " does not match the network's input dim " +
                             std::to_string(hb[cur].dims[s]));
        }
        if (ok && Traw == 0) {
          CVX_WARN("Zero-length utterance: " << utt);
          ok = false;
        }
        if (ok && !pad_input && T < min_chunk_size) {
          CVX_WARN("Minimum chunk size of " << min_chunk_size << " is greater than the number of rows in utterance: " << utt);
          ok = false;
        }
        if (ok) {
          // nnet3-xvector-compute.cc main(): this_chunk_size, num_chunks, per-chunk offset
          int32_t this_chunk = chunk_size;
          if (chunk_size == -1 || T < chunk_size) this_chunk = T;
          const int32_t num_chunks = (T + this_chunk - 1) / this_chunk;
          HostBatch& b = hb[cur];
          // rows this utterance adds to the batch: its own, or with --pad-input every chunk brought
          // up to min-chunk-size (with --chunk-size < --min-chunk-size that is EVERY chunk)
          int64_t need_rows = Traw;
          if (pad_input) {
            need_rows = 0;
            for (int32_t ci = 0; ci < num_chunks; ++ci)
              need_rows += std::max(std::min(this_chunk, T - ci * this_chunk), min_chunk_size);
          }
          b.Reserve(b.rows + std::max<int64_t>(need_rows, Traw) + std::max(min_chunk_size, 0));
          const int64_t chunk_base = raw_mode ? b.kept_rows : b.rows;
          const int32_t utt_index = static_cast<int32_t>(b.keys.size());
          const size_t chunks_before = b.chunk_start.size();
          const int64_t rows_before = b.rows;
          // Fast path (no replicated rows needed): the utterance's rows go into the batch as one
          // block -- or are already there when the reader placed them (single stream).
          const bool tail_needs_pad = pad_input && (T - (num_chunks - 1) * this_chunk) < min_chunk_size;
          bool block_copied = false;
          if (!tail_needs_pad) {
            for (int s = 0; s < num_inputs; ++s) {
              float* dst = b.feats[s] + static_cast<size_t>(b.rows) * b.dims[s];
              if (mats[s]->Data() != dst) std::memcpy(dst, mats[s]->Data(), static_cast<size_t>(Traw) * b.dims[s] * sizeof(float));
            }
            block_copied = true;
          }
          for (int32_t ci = 0; ci < num_chunks; ++ci) {
            const int32_t begin = ci * this_chunk;
            const int32_t offset = std::min(this_chunk, T - begin);
            int32_t left_pad = 0, right_pad = 0;
            bool use = true;
            if (offset < min_chunk_size) {
              if (!pad_input) use = false;
              else {
                left_pad = (min_chunk_size - offset) / 2;
                right_pad = min_chunk_size - offset - left_pad;
              }
            }
            const int32_t rows_here = left_pad + offset + right_pad;
            if (use && rows_here < ctx) {
              CVX_WARN("Chunk of " << rows_here << " frames of utterance " << utt << " is shorter than the network's context ("
                                   << ctx << " frames); skipping it");
              use = false;
            }
            const int64_t start = raw_mode ? chunk_base + begin : b.rows;
            if (block_copied) {
              if (!raw_mode) b.rows += offset;
            } else {
              auto put = [&](int32_t r) {
                for (int s = 0; s < num_inputs; ++s) row_ptrs[s] = mats[s]->Row(r);
                b.AppendRow(row_ptrs);
              };
              for (int32_t k = 0; k < left_pad; ++k) put(begin);
              for (int32_t r = begin; r < begin + offset; ++r) put(r);
              for (int32_t k = 0; k < right_pad; ++k) put(begin + offset - 1);
            }
            if (use) {
              b.chunk_start.push_back(start);
              b.chunk_len.push_back(rows_here);
              b.chunk_utt.push_back(utt_index);
              b.chunk_weight.push_back(static_cast<float>(offset));
            }
          }
          if (b.chunk_start.size() == chunks_before) {
            CVX_WARN("No usable chunk in utterance " << utt);
            b.rows = rows_before;
            ok = false;
          } else {
            b.keys.push_back(utt);
            b.frames += T;
            if (raw_mode) {
              b.rows = rows_before + Traw;
              b.kept_rows += T;
              b.utt_raw_start.push_back(b.rows);
              b.utt_kept_start.push_back(b.kept_rows);
              b.keep.insert(b.keep.end(), utt_keep.begin(), utt_keep.end());
            }
          }
        }
        if (!ok) ++num_fail;
        for (int i = 0; i < num_inputs; ++i) readers[i]->Next();
        if (hb[cur].rows >= batch_rows) cur = launch(cur);
      }
      for (int i = 1; i < num_inputs; ++i)
        if (!readers[i]->Done()) CVX_WARN("feature stream " << (i + 1) << " has extra utterances after the first stream ended");
      launch(cur);
      while (!inflight.empty()) finish_oldest();
      for (auto& r : readers) {
        int st = r->Close();
        if (st != 0) CVX_WARN("a feature input pipe exited with status " << st);
      }
      res->total_s = since(lane_t0);
      res->end = std::chrono::steady_clock::now();
      if (po.verbose >= 1)
        CVX_LOG("lane timing: " << res->total_s << " s = " << res->setup_s << " s waiting for pinned buffers + readers, "
                                << res->gpu_wait_s << " s blocked on the GPU, the rest reading / packing; " << res->calls
                                << " cvx_extract calls took " << res->extract_s << " s (first " << res->first_call_s << " s, longest "
                                << res->max_call_s << " s), of which " << res->device_s << " s between the engine's CUDA events");
    };

    VectorWriter writer(vector_wspecifier);
    int64_t num_success = 0, num_fail = 0, frame_count = 0;
    // the processing loop ends with the last vector handed to the writer; un-pinning the lanes' host buffers
    // (tens of ms per 100 MB) is tear-down, like the destructors that run after upstream's timer
    std::chrono::steady_clock::time_point t_loop_end{};
    if (num_lanes == 1) {
      LaneResult res;
      std::function<bool(const std::string&)> keep;
      if (num_bins > 1) keep = [&](const std::string& k) { return key_bin(k) == shard.first; };
      run_lane(engines, keep, [&](const std::string& key, const float* v) { writer.Write(key, v, E); }, &res);
      num_success = res.num_success;
      num_fail = res.num_fail;
      frame_count = res.frame_count;
      t_loop_end = res.end;
    } else {
      // lanes run side by side, each on its own GPU and its own share of the scp; this thread merges
      // their results back into table order (a lane emits in table order; an utterance missing at the
      // head of its lane's queue failed there)
      std::vector<LaneResult> res(num_lanes);
      std::vector<std::unique_ptr<ResultQueue>> queues;
      for (int g = 0; g < num_lanes; ++g) queues.emplace_back(new ResultQueue);
      std::vector<std::thread> threads;
      for (int g = 0; g < num_lanes; ++g)
        threads.emplace_back([&, g] {
          try {
            const int32_t my_bin = shard.first * num_lanes + g;
            run_lane({engines[g]}, [&, my_bin](const std::string& k) { return key_bin(k) == my_bin; },
                     [&, g](const std::string& key, const float* v) { queues[g]->Push(key, v, E); }, &res[g]);
          } catch (const std::exception& e) {
            res[g].error = e.what();
          }
          queues[g]->Finish();
        });
      struct LaneJoiner {
        std::vector<std::thread>& t;
        ~LaneJoiner() {
          for (std::thread& th : t)
            if (th.joinable()) th.join();
        }
      } lane_joiner{threads};
      for (const auto& kv : SequentialMatrixReader::ReadScpList(feat_specs[0].scp_name)) {
        const int32_t bin = key_bin(kv.first);
        if (bin / num_lanes != shard.first) continue;
        ResultQueue& q = *queues[bin % num_lanes];
        std::unique_lock<std::mutex> lk(q.mu);
        q.cv.wait(lk, [&] { return !q.q.empty() || q.done; });
        if (q.q.empty() || q.q.front().first != kv.first) continue;  // failed in its lane
        std::pair<std::string, std::vector<float>> item = std::move(q.q.front());
        q.q.pop_front();
        lk.unlock();
        writer.Write(item.first, item.second.data(), E);
      }
      t_loop_end = std::chrono::steady_clock::now();
      for (std::thread& th : threads) th.join();
      for (const LaneResult& r : res) {
        if (!r.error.empty()) throw KaldiError(r.error);
        num_success += r.num_success;
        num_fail += r.num_fail;
        frame_count += r.frame_count;
      }
    }
    const auto t_close = std::chrono::steady_clock::now();
    writer.Close();
    double elapsed = std::chrono::duration<double>(t_loop_end - t_start).count() +
                     std::chrono::duration<double>(std::chrono::steady_clock::now() - t_close).count();
    CVX_LOG("Time taken " << elapsed << "s: real-time factor assuming 100 frames/sec is "
                          << (frame_count > 0 ? elapsed * 100.0 / static_cast<double>(frame_count) : 0.0));
    CVX_LOG("Done " << num_success << " utterances, failed for " << num_fail);
    for (int g = 0; profile && g < num_engines; ++g) {
      cvx_kernel_profile recs[16];
      int32_t n = 0;
      cvx_engine_get_profile(engines[g], recs, 16, &n);
      for (int32_t i = 0; i < n; ++i)
        CVX_LOG("profile gpu" << (gpu_id + g) << " " << recs[i].name << ": launches=" << recs[i].launches
                              << " ms=" << recs[i].total_ms
                              << (recs[i].flops > 0 ? " TFLOP/s=" + std::to_string(recs[i].flops / recs[i].total_ms * 1e-9) : "")
                              << (recs[i].bytes > 0 ? " GB/s=" + std::to_string(recs[i].bytes / recs[i].total_ms * 1e-6) : ""));
    }
    if (check_saturation) {
      // names of the inspected buffers: the staged inputs, then the frame layers in plan order
      std::vector<std::string> names{"the staged input features"};
      std::istringstream desc(cvx_model_describe(model));
      for (std::string ln; std::getline(desc, ln);)
        if (ln.rfind("frame[", 0) == 0) {
          std::istringstream w(ln);
          std::string idx, name;
          w >> idx >> name;
          names.push_back("layer " + name);
        }
      int64_t total = 0;
      for (int g = 0; g < num_engines; ++g) {
        int64_t counts[80];
        int32_t n = 0;
        if (cvx_engine_saturation_counts(engines[g], counts, 80, &n) != CVX_OK) continue;
        for (int32_t i = 0; i < n; ++i) {
          if (counts[i] == 0) continue;
          total += counts[i];
          CVX_WARN(counts[i] << " values of " << (static_cast<size_t>(i) < names.size() ? names[i] : std::string("buffer ") + std::to_string(i))
                             << " were clipped at the fp16 range limit (+-65504); the embeddings are not reliable -- use --precision=bf16 "
                                "(fp32 exponent range) or --precision=fp32 for this model");
        }
      }
      if (total == 0) CVX_LOG("range check: no activation reached the fp16 range limit");
    }
    const int status = num_success != 0 ? 0 : 1;
    if (std::getenv("CVX_FULL_TEARDOWN") == nullptr) {
      // Everything this job produced is written and closed.  Freeing GBs of device buffers one by one and
      // destroying the CUDA context in order costs 0.2-0.5 s per job; the driver reclaims all of it (and the
      // device lock file) when the process ends, so end it here.  CVX_FULL_TEARDOWN=1 keeps the orderly path.
      if (po.verbose >= 1) CVX_LOG("done " << since_start() << " s since main(); leaving device tear-down to process exit");
      std::fflush(nullptr);
      _exit(status);
    }
    const double t_done = since_start();
    for (cvx_engine* e : engines) cvx_engine_destroy(e);
    cvx_model_free(model);
    if (po.verbose >= 1) CVX_LOG("tear-down (engines, device memory): " << (since_start() - t_done) << " s; " << since_start() << " s since main()");
    return status;
  } catch (const std::exception& e) {
    LogMessage("ERROR", "main", __FILE__, __LINE__, e.what());
    return -1 & 0xff;
  }
}

}  // namespace cvx
