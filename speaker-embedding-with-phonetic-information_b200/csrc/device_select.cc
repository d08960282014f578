// Device selection and utterance sharding for the recipe's job fan-out.
//
// The reference launches `nj` independent processes per stage, each with `--use-gpu=yes` and NO
// device hint (egs/sre/v5/sid/nnet3/xvector/extract_xvectors_new.sh:83-89), and shards the data
// with utils/split_data.sh (per speaker, not length balanced; :72-74).  Upstream picks a GPU per
// process in CuDevice::SelectGpuId (exclusive-mode devices, or the one with most free memory).
// Here: cvx_acquire_device hands every process of a box its own GPU through advisory lock files
// (first free device; when all are taken the jobs spread evenly), and cvx_lpt_shards is the
// length-balanced partition (greedy LPT over frame counts) that replaces split_data.sh.
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/file.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cctype>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <queue>
#include <string>
#include <thread>
#include <vector>

#include "../../include/cvx.h"
#include "api_error.h"

namespace cvx {
std::string NormalizeDeviceId(const std::string& id);
namespace {

constexpr int kMaxSharers = 64;
int g_lock_fd = -1;  // held until the process exits (the kernel drops the lock with the descriptor)

}  // namespace

// "0000:1B:00.0" (CUDA) and "0000:1b:00.0" (/proc) name the same lock file
std::string NormalizeDeviceId(const std::string& id_in) {
  std::string id = id_in;
  for (char& c : id) {
    if (c == ':' || c == '/') c = '_';
    else c = static_cast<char>(std::tolower(static_cast<unsigned char>(c)));
  }
  return id;
}

namespace {

std::string LockDir() {
  const char* e = std::getenv("CVX_LOCK_DIR");
  return (e && *e) ? e : "/tmp";
}

// Tries, round by round, to take lock `<dir>/cvx-gpu-<uid>-<id>.<round>` of every device: round 0
// grants exclusive use, round k means k other jobs already run on that device.
int AcquireLock(const std::string& dir, const std::vector<std::string>& ids, bool wait, int32_t* device,
                int32_t* sharers) {
  if (ids.empty()) return CVX_ERR_INVALID;
  if (g_lock_fd >= 0) {
    close(g_lock_fd);
    g_lock_fd = -1;
  }
  auto try_round = [&](int round) -> bool {
    for (size_t d = 0; d < ids.size(); ++d) {
      const std::string path =
          dir + "/cvx-gpu-" + std::to_string(static_cast<long>(getuid())) + "-" + ids[d] + "." + std::to_string(round);
      int fd = open(path.c_str(), O_CREAT | O_RDWR | O_CLOEXEC, 0666);
      if (fd < 0) continue;
      if (flock(fd, LOCK_EX | LOCK_NB) == 0) {
        g_lock_fd = fd;
        *device = static_cast<int32_t>(d);
        *sharers = round;
        return true;
      }
      close(fd);
    }
    return false;
  };
  if (wait) {
    // --use-gpu=wait: block until some device is free
    while (!try_round(0)) std::this_thread::sleep_for(std::chrono::milliseconds(200));
    return CVX_OK;
  }
  for (int round = 0; round < kMaxSharers; ++round)
    if (try_round(round)) return CVX_OK;
  // lock directory unusable or absurdly many jobs: spread by pid
  *device = static_cast<int32_t>(static_cast<unsigned long>(getpid()) % ids.size());
  *sharers = kMaxSharers;
  return CVX_OK;
}

}  // namespace
}  // namespace cvx

extern "C" {

int cvx_acquire_device(int32_t wait, int32_t* device, int32_t* sharers) {
  if (!device) return cvx::SetApiError(CVX_ERR_INVALID, "cvx_acquire_device: device is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return cvx::SetApiError(CVX_ERR_CUDA, std::string("no usable CUDA device (") + cudaGetErrorString(e) +
                                              "); this path has no CPU fallback");
  // lock names carry the PCI bus id, so that processes with different CUDA_VISIBLE_DEVICES agree
  std::vector<std::string> ids;
  for (int d = 0; d < n; ++d) {
    char buf[32] = {0};
    if (cudaDeviceGetPCIBusId(buf, sizeof(buf), d) != cudaSuccess) std::snprintf(buf, sizeof(buf), "ordinal%d", d);
    ids.push_back(cvx::NormalizeDeviceId(buf));
  }
  int32_t dev = 0, sh = 0;
  int rc = cvx::AcquireLock(cvx::LockDir(), ids, wait != 0, &dev, &sh);
  if (rc != CVX_OK) return cvx::SetApiError(rc, "cvx_acquire_device: could not take a device lock");
  *device = dev;
  if (sharers) *sharers = sh;
  return CVX_OK;
}

int cvx_acquire_device_by_id(const char* const* ids, int32_t num_ids, int32_t wait, int32_t* index, int32_t* sharers) {
  if (!ids || num_ids <= 0 || !index) return cvx::SetApiError(CVX_ERR_INVALID, "cvx_acquire_device_by_id: bad argument");
  std::vector<std::string> names;
  for (int32_t i = 0; i < num_ids; ++i) {
    if (!ids[i]) return cvx::SetApiError(CVX_ERR_INVALID, "cvx_acquire_device_by_id: NULL id");
    names.push_back(cvx::NormalizeDeviceId(ids[i]));
  }
  int32_t dev = 0, sh = 0;
  int rc = cvx::AcquireLock(cvx::LockDir(), names, wait != 0, &dev, &sh);
  if (rc != CVX_OK) return cvx::SetApiError(rc, "cvx_acquire_device_by_id: could not take a device lock");
  *index = dev;
  if (sharers) *sharers = sh;
  return CVX_OK;
}

int cvx_test_acquire_lock(const char* dir, int32_t num_devices, int32_t wait, int32_t* device, int32_t* sharers) {
  if (!dir || num_devices <= 0 || !device || !sharers)
    return cvx::SetApiError(CVX_ERR_INVALID, "cvx_test_acquire_lock: bad argument");
  std::vector<std::string> ids;
  for (int d = 0; d < num_devices; ++d) ids.push_back("test" + std::to_string(d));
  return cvx::AcquireLock(dir, ids, wait != 0, device, sharers);
}

int cvx_lpt_shards(const int64_t* lengths, int64_t num_utts, int32_t num_shards, int32_t* shard_of) {
  if (num_utts < 0 || num_shards < 1 || (num_utts > 0 && (!lengths || !shard_of)))
    return cvx::SetApiError(CVX_ERR_INVALID, "cvx_lpt_shards: bad argument");
  // longest first (ties: input order) onto the currently lightest shard (ties: lowest shard)
  std::vector<int64_t> order(static_cast<size_t>(num_utts));
  std::iota(order.begin(), order.end(), int64_t{0});
  std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return lengths[a] > lengths[b]; });
  using Load = std::pair<int64_t, int32_t>;
  std::priority_queue<Load, std::vector<Load>, std::greater<Load>> heap;
  for (int32_t s = 0; s < num_shards; ++s) heap.push({0, s});
  for (int64_t i : order) {
    Load l = heap.top();
    heap.pop();
    shard_of[i] = l.second;
    heap.push({l.first + lengths[i], l.second});
  }
  return CVX_OK;
}

}  // extern "C"
