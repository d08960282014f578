// Drop-in command-line driver for the frame-level program of the same family: upstream Kaldi's
// nnet3-compute as the reference's recipes call it
//   egs/sre/v5/sid/nnet3_cvector/cvector/extract_am_embedding.sh:79-85
//     nnet3-compute $compute_gpu_opt "$nnet" "$feats" ark:- | copy-feats --compress=... ark:- ark,scp:A,B
//   egs/sre/v5/sid/nnet3_cvector/am/extract_bn.sh:57-69          (same form)
//   egs/sre/v5/sid/nnet3_cvector/cvector/extract_log_post.sh:80-85
//     nnet3-compute --use-gpu=no "$nnet" "$feats" ark,scp:A,B
// One output row per input frame; frames the network needs beyond the ends of an utterance are
// copies of its first / last frame (upstream's DecodableNnetSimple rule).  The forward pass goes
// through the C ABI (cvx_model_load_frames + cvx_compute_frames); the main thread keeps reading
// the feature pipe into one pinned batch while a worker thread computes and writes the other.
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>

#include "../../../include/cvx.h"
#include "../kaldi_io.h"
#include "extract_main.h"
#include "job_util.h"

namespace cvx {
namespace {

struct FrameBatch {
  float* feats = nullptr;  // pinned [cap_rows, dim]
  size_t cap_rows = 0;
  int32_t dim = 0;
  std::vector<std::string> keys;
  std::vector<int64_t> utt_start{0};
  float* out = nullptr;  // pinned [out_cap]
  size_t out_cap = 0;

  int64_t rows() const { return utt_start.back(); }
  void Reset() {
    keys.clear();
    utt_start.assign(1, 0);
  }
  void Reserve(int64_t need_rows) {
    if (static_cast<size_t>(need_rows) <= cap_rows) return;
    size_t ncap = std::max<size_t>(static_cast<size_t>(need_rows) * 5 / 4, 1 << 16);
    void* p = nullptr;
    if (cvx_host_alloc(&p, ncap * dim * sizeof(float)) != CVX_OK) throw KaldiError(cvx_last_error());
    if (rows()) std::memcpy(p, feats, static_cast<size_t>(rows()) * dim * sizeof(float));
    cvx_host_free(feats);
    feats = static_cast<float*>(p);
    cap_rows = ncap;
  }
  void ReserveOut(size_t n) {
    if (n <= out_cap) return;
    cvx_host_free(out);
    out = nullptr;
    void* p = nullptr;
    if (cvx_host_alloc(&p, n * sizeof(float)) != CVX_OK) throw KaldiError(cvx_last_error());
    out = static_cast<float*>(p);
    out_cap = n;
  }
  ~FrameBatch() {
    cvx_host_free(feats);
    cvx_host_free(out);
  }
};

}  // namespace

int ComputeMain(int argc, const char* const* argv) {
  try {
    ParseOptions po(
        "Runs the frame-level layers of an nnet3 raw model on the GPU and writes, per utterance,\n"
        "the matrix of output-node 'output' with one row per input frame (e.g. bottleneck\n"
        "features or log-posteriors).  Frames needed beyond the ends of an utterance repeat its\n"
        "first / last frame.\n"
        "\n"
        "Usage: nnet3-compute [options] <raw-nnet-in> <features-rspecifier> <matrix-wspecifier>\n"
        "e.g.: nnet3-compute final.raw scp:feats.scp ark:nnet_prediction.ark\n"
        "See also: nnet3-xvector-compute\n");
    std::string use_gpu = "yes", output_node, nnet_config, precision = "fp16", ivectors, online_ivectors;
    std::string shard_spec, utt2num_frames;
    int32_t gpu_id = -1, subsampling = 1;
    int64_t batch_rows = 1 << 19, max_batch_rows = 0;
    bool apply_exp = false, use_priors = false, profile = false;
    int32_t ign_i[8] = {};
    float ign_f = 1.0f;
    bool ign_b[2] = {false, false};

    po.Register("use-gpu", &use_gpu, "yes|no|optional|wait, accepted for compatibility; this program always runs on a GPU");
    po.Register("apply-exp", &apply_exp, "If true, apply exp to the output (log-posteriors -> posteriors)");
    po.Register("use-priors", &use_priors, "Only false is supported (raw models carry no priors)");
    po.Register("ivectors", &ivectors, "not supported (the recipes' networks have no ivector input)");
    po.Register("online-ivectors", &online_ivectors, "not supported");
    po.Register("frame-subsampling-factor", &subsampling, "Only 1 is supported");
    po.Register("output-node", &output_node,
                "Descriptor to use as output-node 'output' (same effect as piping the model through "
                "nnet3-copy --nnet-config=extract.config)");
    po.Register("nnet-config", &nnet_config, "File with an 'output-node name=output input=...' line (extract.config)");
    po.Register("precision", &precision, "TDNN operand type: fp16 | bf16 | fp32 (fp32 = slow validation kernels)");
    po.Register("gpu-id", &gpu_id, "CUDA device ordinal (-1: $CVX_DEVICE, else a free GPU is picked and held through a lock file)");
    po.Register("shard", &shard_spec, "J/N: process only the J-th (1-based) of N length-balanced shards of the feature table");
    po.Register("utt2num-frames", &utt2num_frames, "frame counts for --shard (scp inputs can do without)");
    po.Register("batch-rows", &batch_rows, "Frames read from the feature pipe per cvx_compute_frames call");
    po.Register("max-batch-rows", &max_batch_rows, "Frames per packed device batch (0 = engine default)");
    po.Register("profile", &profile, "Print per-kernel CUDA-event times at the end");
    // accepted for command-line compatibility with upstream; no effect on the result here
    po.Register("online-ivector-period", &ign_i[0], "ignored");
    po.Register("extra-left-context", &ign_i[1], "ignored");
    po.Register("extra-right-context", &ign_i[2], "ignored");
    po.Register("extra-left-context-initial", &ign_i[3], "ignored");
    po.Register("extra-right-context-final", &ign_i[4], "ignored");
    po.Register("frames-per-chunk", &ign_i[5], "ignored (whole utterances are evaluated at once)");
    po.Register("acoustic-scale", &ign_f, "ignored");
    po.Register("debug-computation", &ign_b[0], "ignored");
    po.Register("optimization.optimize", &ign_b[1], "ignored");

    po.Read(argc, argv);
    if (po.NumArgs() != 3) {
      po.PrintUsage();
      return 1;
    }
    const std::string nnet_rxfilename = po.GetArg(1), feat_rspecifier = po.GetArg(2), matrix_wspecifier = po.GetArg(3);
    if (use_gpu != "yes" && use_gpu != "no" && use_gpu != "optional" && use_gpu != "wait")
      throw KaldiError("invalid value for --use-gpu: " + use_gpu);
    if (use_gpu == "no") CVX_LOG("--use-gpu=no requested; this implementation has no CPU path and runs on the GPU");
    if (!ivectors.empty() || !online_ivectors.empty()) throw KaldiError("ivector inputs are not supported on this path");
    if (use_priors) throw KaldiError("--use-priors=true is not supported (raw models carry no priors)");
    if (subsampling != 1) throw KaldiError("--frame-subsampling-factor != 1 is not supported");
    if (!nnet_config.empty()) output_node = ReadOutputNodeFromConfig(nnet_config);
    int32_t prec;
    if (precision == "fp16") prec = CVX_PRECISION_FP16;
    else if (precision == "bf16") prec = CVX_PRECISION_BF16;
    else if (precision == "fp32") prec = CVX_PRECISION_FP32;
    else throw KaldiError("invalid --precision " + precision);
    const std::pair<int32_t, int32_t> shard = ParseShard(shard_spec);
    gpu_id = ChooseDevice(gpu_id, use_gpu, 1);

    cvx_model* model = nullptr;
    if (cvx_model_load_frames(nnet_rxfilename.c_str(), output_node.empty() ? nullptr : output_node.c_str(), &model) != CVX_OK)
      throw KaldiError(std::string("reading ") + nnet_rxfilename + ": " + cvx_last_error());
    cvx_model_info info;
    cvx_model_get_info(model, &info);
    if (po.verbose >= 1) CVX_LOG("frame plan:\n" << cvx_model_describe(model));
    if (info.num_inputs != 1)
      throw KaldiError("the network reads " + std::to_string(info.num_inputs) + " input nodes; nnet3-compute feeds one");
    cvx_engine_options eo;
    cvx_engine_options_default(&eo);
    eo.device = gpu_id;
    eo.precision = prec;
    eo.max_batch_rows = max_batch_rows;
    eo.profile = profile ? 1 : 0;
    cvx_engine* engine = nullptr;
    if (cvx_engine_create(model, &eo, &engine) != CVX_OK) throw KaldiError(std::string("creating the GPU engine: ") + cvx_last_error());

    std::unordered_map<std::string, int32_t> bin_of;
    std::function<bool(const std::string&)> keep;
    if (shard.second > 1) {
      const TableSpec fs = ParseRspecifier(feat_rspecifier);
      bin_of = AssignShards(utt2num_frames, fs.is_scp ? fs.scp_name : std::string(), shard.second, nullptr);
      keep = [&](const std::string& k) {
        auto it = bin_of.find(k);
        return it != bin_of.end() && it->second == shard.first;
      };
    }
    SequentialMatrixReader reader(feat_rspecifier, false, keep);
    MatrixWriter writer(matrix_wspecifier);
    const int32_t dim = cvx_model_input_dim(model, 0), E = info.embedding_dim;
    // wide outputs (log-posteriors over thousands of senones): keep a batch's result under 1 GB of pinned memory
    batch_rows = std::min<int64_t>(batch_rows, std::max<int64_t>(4096, (int64_t{1} << 28) / std::max(E, 1)));
    FrameBatch fb[2];
    for (FrameBatch& b : fb) {
      b.dim = dim;
      b.Reserve(batch_rows + batch_rows / 8 + 65536);
    }
    int64_t num_success = 0, num_fail = 0, frame_count = 0;
    auto t_start = std::chrono::steady_clock::now();

    // worker: compute one batch and write its matrices while the main thread reads the next
    std::thread worker;
    struct Joiner {
      std::thread& t;
      ~Joiner() {
        if (t.joinable()) t.join();
      }
    } joiner{worker};
    std::string worker_error;
    int inflight = -1;
    auto finish_inflight = [&]() {
      if (inflight < 0) return;
      worker.join();
      if (!worker_error.empty()) throw KaldiError(worker_error);
      FrameBatch& b = fb[inflight];
      num_success += static_cast<int64_t>(b.keys.size());
      frame_count += b.rows();
      b.Reset();
      inflight = -1;
    };
    auto launch = [&](int which) {
      if (fb[which].keys.empty()) return;
      finish_inflight();
      inflight = which;
      worker = std::thread([&, which]() {
        try {
          FrameBatch& w = fb[which];
          w.ReserveOut(static_cast<size_t>(w.rows()) * E);
          const float* f[1] = {w.feats};
          if (cvx_compute_frames(engine, f, 1, w.rows(), w.utt_start.data(), static_cast<int32_t>(w.keys.size()),
                                 w.out, apply_exp ? CVX_FLAG_APPLY_EXP : 0) != CVX_OK)
            throw KaldiError(std::string("cvx_compute_frames failed: ") + cvx_last_error());
          for (size_t u = 0; u < w.keys.size(); ++u)
            writer.Write(w.keys[u], w.out + static_cast<size_t>(w.utt_start[u]) * E,
                         static_cast<int32_t>(w.utt_start[u + 1] - w.utt_start[u]), E);
        } catch (const std::exception& e) {
          worker_error = e.what();
        }
      });
    };

    int cur = 0;
    // binary float matrices are read straight into the pinned batch at its tail
    reader.Value().placer = [&](int32_t rows, int32_t cols) -> float* {
      FrameBatch& b = fb[cur];
      if (cols != b.dim) return nullptr;
      b.Reserve(b.rows() + rows);
      return b.feats + static_cast<size_t>(b.rows()) * b.dim;
    };
    for (; !reader.Done(); reader.Next()) {
      const std::string& key = reader.Key();
      const Matrix& m = reader.Value();
      if (m.rows == 0) {
        CVX_WARN("Zero-length utterance: " << key);
        ++num_fail;
        continue;
      }
      if (m.cols != dim)
        throw KaldiError("feature dimension " + std::to_string(m.cols) + " of " + key + " does not match the network's input (" +
                         std::to_string(dim) + ")");
      FrameBatch& b = fb[cur];
      b.Reserve(b.rows() + m.rows);
      float* dst = b.feats + static_cast<size_t>(b.rows()) * dim;
      if (m.Data() != dst) std::memcpy(dst, m.Data(), static_cast<size_t>(m.rows) * dim * sizeof(float));
      b.keys.push_back(key);
      b.utt_start.push_back(b.rows() + m.rows);
      if (b.rows() >= batch_rows) {
        launch(cur);
        cur ^= 1;
        if (inflight == cur) finish_inflight();
      }
    }
    launch(cur);
    finish_inflight();
    int st = reader.Close();
    if (st != 0) CVX_WARN("the feature input pipe exited with status " << st);
    writer.Close();

    double elapsed = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
    CVX_LOG("Time taken " << elapsed << "s: real-time factor assuming 100 frames/sec is "
                          << (frame_count > 0 ? elapsed * 100.0 / static_cast<double>(frame_count) : 0.0));
    CVX_LOG("Done " << num_success << " utterances, failed for " << num_fail);
    if (profile) {
      cvx_kernel_profile recs[16];
      int32_t n = 0;
      cvx_engine_get_profile(engine, recs, 16, &n);
      for (int32_t i = 0; i < n; ++i)
        CVX_LOG("profile " << recs[i].name << ": launches=" << recs[i].launches << " ms=" << recs[i].total_ms);
    }
    cvx_engine_destroy(engine);
    cvx_model_free(model);
    return num_success != 0 ? 0 : 1;
  } catch (const std::exception& e) {
    LogMessage("ERROR", "main", __FILE__, __LINE__, e.what());
    return -1 & 0xff;
  }
}

}  // namespace cvx
