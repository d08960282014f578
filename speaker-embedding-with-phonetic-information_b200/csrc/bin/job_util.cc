#include "job_util.h"

#include <cstdlib>
#include <sstream>

#include "../../../include/cvx.h"
#include "../kaldi_io.h"

namespace cvx {

int32_t ChooseDevice(int32_t gpu_id, const std::string& use_gpu, int32_t num_gpus) {
  if (gpu_id >= 0) return gpu_id;
  if (const char* e = std::getenv("CVX_DEVICE")) return std::atoi(e);
  if (num_gpus > 1) return 0;  // --num-gpus=G takes devices 0 .. G-1
  int32_t dev = 0, sharers = 0;
  if (cvx_acquire_device(use_gpu == "wait" ? 1 : 0, &dev, &sharers) != CVX_OK)
    throw KaldiError(std::string("selecting a GPU: ") + cvx_last_error());
  if (sharers == 0) CVX_LOG("Selected GPU " << dev << " (free; held for this job)");
  else CVX_LOG("All GPUs are in use by other jobs; sharing GPU " << dev << " with " << sharers << " other job(s)");
  return dev;
}

std::pair<int32_t, int32_t> ParseShard(const std::string& spec) {
  if (spec.empty()) return {0, 1};
  const size_t slash = spec.find('/');
  char* e1 = nullptr;
  char* e2 = nullptr;
  long j = std::strtol(spec.c_str(), &e1, 10);
  long n = slash == std::string::npos ? 0 : std::strtol(spec.c_str() + slash + 1, &e2, 10);
  if (slash == std::string::npos || e1 != spec.c_str() + slash || !e2 || *e2 || n < 1 || j < 1 || j > n)
    throw KaldiError("--shard expects J/N with 1 <= J <= N, got '" + spec + "'");
  return {static_cast<int32_t>(j - 1), static_cast<int32_t>(n)};
}

int64_t PeekEntryRows(const std::string& rxfilename) {
  Input in(rxfilename);
  if (!ReadBinaryHeader(in)) return -1;
  const std::string tag = in.ReadToken();
  if (tag == "FV" || tag == "DV") return 1;
  if (tag == "FM" || tag == "DM") return ReadInt32(in, true);
  if (tag == "CM" || tag == "CM2" || tag == "CM3") {
    float hdr[2];
    int32_t rows = 0;
    in.Read(hdr, 8);
    in.Read(&rows, 4);
    return rows;
  }
  throw KaldiError("unknown matrix tag '" + tag + "' in " + rxfilename);
}

std::vector<std::pair<std::string, int64_t>> ReadUtt2NumFrames(const std::string& path) {
  std::vector<std::pair<std::string, int64_t>> out;
  Input in(path);
  while (!in.Eof()) {
    std::istringstream line(in.ReadLine());
    std::string key;
    long long n = 0;
    if (!(line >> key)) continue;
    if (!(line >> n) || n < 0) throw KaldiError("bad line for " + key + " in " + path + " (expected '<utt> <frames>')");
    out.emplace_back(key, static_cast<int64_t>(n));
  }
  in.Close();
  return out;
}

std::unordered_map<std::string, int32_t> AssignShards(const std::string& utt2num_frames, const std::string& scp_rxfilename,
                                                      int32_t num_shards, std::vector<std::string>* order) {
  std::vector<std::pair<std::string, int64_t>> table;
  if (!utt2num_frames.empty()) {
    table = ReadUtt2NumFrames(utt2num_frames);
  } else {
    if (scp_rxfilename.empty())
      throw KaldiError("length-balanced sharding of an archive (ark:) input needs --utt2num-frames=<file>");
    for (const auto& kv : SequentialMatrixReader::ReadScpList(scp_rxfilename)) {
      int64_t rows = PeekEntryRows(kv.second);
      if (rows < 0) throw KaldiError("cannot read the frame count of text entry " + kv.first + "; give --utt2num-frames=<file>");
      table.emplace_back(kv.first, rows);
    }
  }
  std::vector<int64_t> lengths(table.size());
  for (size_t i = 0; i < table.size(); ++i) lengths[i] = table[i].second;
  std::vector<int32_t> shard_of(table.size(), 0);
  if (cvx_lpt_shards(lengths.data(), static_cast<int64_t>(lengths.size()), num_shards, shard_of.data()) != CVX_OK)
    throw KaldiError(std::string("cvx_lpt_shards: ") + cvx_last_error());
  std::unordered_map<std::string, int32_t> out;
  out.reserve(table.size() * 2);
  for (size_t i = 0; i < table.size(); ++i) {
    out[table[i].first] = shard_of[i];
    if (order) order->push_back(table[i].first);
  }
  return out;
}

}  // namespace cvx
