// Range guard for fp16 operands: the epilogues and the input staging convert
// with cvt.rn.satfinite, so a value beyond the fp16 range is clipped to +-65504 silently.  In check
// mode (cvx_engine_options.reserved[1] / --check-saturation) the engine stores every layer and this
// kernel counts the 16-bit values sitting at the limit, per buffer.
#include "kernels.h"

namespace cvx {
namespace {

__global__ void __launch_bounds__(256) count_saturated_kernel(const uint16_t* __restrict__ buf, int64_t rows, int32_t ld,
                                                              int32_t dim, unsigned long long* __restrict__ counter) {
  unsigned long long n = 0;
  const int64_t total = rows * dim;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / dim;
    const int32_t c = static_cast<int32_t>(i - r * dim);
    n += (buf[r * ld + c] & 0x7FFFu) >= 0x7BFFu;  // max finite fp16, or (never produced by satfinite) inf / nan
  }
  for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
  if ((threadIdx.x & 31) == 0 && n) atomicAdd(counter, n);
}

}  // namespace

cudaError_t LaunchCountSaturated(const void* buf, int64_t rows, int32_t ld, int32_t dim, unsigned long long* counter,
                                 cudaStream_t stream) {
  if (rows <= 0 || dim <= 0) return cudaSuccess;
  const int64_t total = rows * dim;
  const unsigned grid = static_cast<unsigned>(total / 256 / 8 + 1 < 148 * 8 ? total / 256 / 8 + 1 : 148 * 8);
  count_saturated_kernel<<<grid, 256, 0, stream>>>(static_cast<const uint16_t*>(buf), rows, ld, dim, counter);
  return cudaGetLastError();
}

}  // namespace cvx
