// Job fan-out helpers shared by the drop-in programs: which GPU this process uses and which
// utterances of the table are its share (see csrc/device_select.cc for the C-ABI side).
#ifndef CVX_BIN_JOB_UTIL_H_
#define CVX_BIN_JOB_UTIL_H_

#include <cstdint>
#include <memory>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

namespace cvx {

// --gpu-id=N, else $CVX_DEVICE, else (one GPU per process) an automatically acquired device:
// the recipe's GPU branch starts nj jobs with --use-gpu=yes and no device hint
// (egs/sre/v5/sid/nnet3/xvector/extract_xvectors_new.sh:83-89).  Logs the choice.
int32_t ChooseDevice(int32_t gpu_id, const std::string& use_gpu, int32_t num_gpus);

// "--shard=J/N": J is 1-based like run.pl's JOB; returns (J-1, N).  "" -> (0, 1).
std::pair<int32_t, int32_t> ParseShard(const std::string& spec);

// Rows of one table entry "path[:offset]" read from its header only (binary FM / DM / CM / CM2 /
// CM3 matrices, FV / DV vectors); -1 for text entries.
int64_t PeekEntryRows(const std::string& rxfilename);

// (key, frames) per line of an utt2num_frames file (steps/make_mfcc.sh --write-utt2num-frames true,
// egs/sre/v2/run_sre16.sh:48).
std::vector<std::pair<std::string, int64_t>> ReadUtt2NumFrames(const std::string& path);

// key -> shard (0-based) for `num_shards` length-balanced shards (cvx_lpt_shards).  Frame counts come
// from `utt2num_frames` if non-empty, else from the matrix headers of `scp_rxfilename`'s entries.
// *order (optional) receives the keys in table order.
std::unordered_map<std::string, int32_t> AssignShards(const std::string& utt2num_frames, const std::string& scp_rxfilename,
                                                      int32_t num_shards, std::vector<std::string>* order);

}  // namespace cvx
#endif  // CVX_BIN_JOB_UTIL_H_
