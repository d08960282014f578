#include "job_util.h"

#include <dirent.h>
#include <unistd.h>

#include <algorithm>
#include <cstdlib>
#include <fstream>
#include <sstream>

#include "../../../include/cvx.h"
#include "../kaldi_io.h"

namespace cvx {

namespace {
// PCI bus ids of the GPUs this process can open, in PCI order, WITHOUT touching CUDA:
// /proc/driver/nvidia/gpus/<bus id>/information holds "Device Minor: N"; kept if /dev/nvidiaN is accessible.
std::vector<std::string> ListGpusFromProc() {
  std::vector<std::string> ids;
  const std::string root = "/proc/driver/nvidia/gpus";
  DIR* d = opendir(root.c_str());
  if (!d) return ids;
  while (dirent* e = readdir(d)) {
    const std::string name = e->d_name;
    if (name == "." || name == "..") continue;
    std::ifstream info(root + "/" + name + "/information");
    std::string line;
    int minor = -1;
    while (std::getline(info, line)) {
      const size_t p = line.find("Device Minor:");
      if (p != std::string::npos) minor = std::atoi(line.c_str() + p + 13);
    }
    if (minor < 0) continue;
    if (access(("/dev/nvidia" + std::to_string(minor)).c_str(), R_OK | W_OK) != 0) continue;
    ids.push_back(name);
  }
  closedir(d);
  std::sort(ids.begin(), ids.end());
  return ids;
}
}  // namespace

int32_t ChooseDevice(int32_t gpu_id, const std::string& use_gpu, int32_t num_gpus) {
  if (gpu_id >= 0) return gpu_id;
  if (const char* e = std::getenv("CVX_DEVICE")) return std::atoi(e);
  if (num_gpus > 1) return 0;  // --num-gpus=G takes devices 0 .. G-1
  int32_t dev = 0, sharers = 0;
  // Preferred: pick the GPU before CUDA is initialised and make it the only visible one, so that this job
  // does not pay (and hold) driver state on every GPU of the box -- with nj jobs starting together
  // that is most of their start-up time.
  if (!std::getenv("CUDA_VISIBLE_DEVICES") && !std::getenv("CVX_SELECT_WITH_CUDA")) {
    const std::vector<std::string> ids = ListGpusFromProc();
    if (!ids.empty()) {
      std::vector<const char*> c;
      for (const std::string& s : ids) c.push_back(s.c_str());
      if (cvx_acquire_device_by_id(c.data(), static_cast<int32_t>(c.size()), use_gpu == "wait" ? 1 : 0, &dev, &sharers) == CVX_OK) {
        setenv("CUDA_DEVICE_ORDER", "PCI_BUS_ID", 1);
        setenv("CUDA_VISIBLE_DEVICES", std::to_string(dev).c_str(), 1);
        if (sharers == 0) CVX_LOG("Selected GPU " << dev << " (" << ids[dev] << ": free; held for this job)");
        else CVX_LOG("All " << ids.size() << " GPUs are in use by other jobs; sharing GPU " << dev << " (" << ids[dev] << ") with "
                            << sharers << " other job(s)");
        return 0;  // the only device CUDA will see
      }
    }
  }
  // CUDA_VISIBLE_DEVICES already set (a container runtime, a scheduler): the choice is among its entries, and
  // it can still be made before CUDA starts -- the lock is named after the entry, the winner becomes the only
  // visible device.  (All jobs of a stage inherit the same list; a job that names the same GPU differently, or
  // selects through CUDA, does not see these locks.)
  if (const char* cvd_env = std::getenv("CUDA_VISIBLE_DEVICES")) {
    const std::string cvd = cvd_env;  // setenv below may free the old value
    std::vector<std::string> entries;
    bool usable = !std::getenv("CVX_SELECT_WITH_CUDA");
    std::istringstream list(cvd);
    for (std::string e; usable && std::getline(list, e, ',');) {
      const size_t a = e.find_first_not_of(" \t"), b = e.find_last_not_of(" \t");
      if (a == std::string::npos || e[a] == '-') usable = false;  // CUDA stops at an empty or negative entry
      else entries.push_back(e.substr(a, b - a + 1));
    }
    if (usable && !entries.empty()) {
      std::vector<std::string> names;
      for (const std::string& e : entries) names.push_back("cvd-" + e);
      std::vector<const char*> c;
      for (const std::string& n : names) c.push_back(n.c_str());
      if (cvx_acquire_device_by_id(c.data(), static_cast<int32_t>(c.size()), use_gpu == "wait" ? 1 : 0, &dev, &sharers) == CVX_OK) {
        setenv("CUDA_VISIBLE_DEVICES", entries[dev].c_str(), 1);
        if (sharers == 0) CVX_LOG("Selected GPU " << entries[dev] << " (entry " << dev << " of CUDA_VISIBLE_DEVICES=" << cvd << ": free; held for this job)");
        else CVX_LOG("All " << entries.size() << " GPUs of CUDA_VISIBLE_DEVICES=" << cvd << " are in use by other jobs; sharing GPU "
                            << entries[dev] << " with " << sharers << " other job(s)");
        return 0;  // the only device CUDA will see
      }
    }
  }
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
