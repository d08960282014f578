// Segment-level affine layers (tdnn6 and later): parameters, the hi/lo-split tensor-core path and
// the stand-alone test entry.
#include <algorithm>
#include <cstring>
#include <vector>

#include "engine.h"
#include "engine_detail.h"

namespace cvx {

using detail::ConvertToElem;
using detail::ElemToFloat;
using detail::Encode2D;
using detail::RoundUp;

namespace {
template <typename T>
T* Upload(const T* src, size_t n) {
  void* p = nullptr;
  CVX_CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
  if (n) CVX_CUDA_CHECK(cudaMemcpy(p, src, n * sizeof(T), cudaMemcpyHostToDevice));
  return static_cast<T*>(p);
}
}  // namespace

void BuildSegmentOp(const Layer& L, int32_t elem, SegmentOp* op) {
  SegmentOp& s = *op;
  s.in_dim = L.in_dim;
  s.out_dim = L.out_dim;
  s.relu = L.relu;
  s.d_w = Upload(L.weight.data(), L.weight.size());
  if (!L.bias.empty()) s.d_b = Upload(L.bias.data(), L.bias.size());
  if (!L.scale.empty()) {
    s.d_scale = Upload(L.scale.data(), L.scale.size());
    s.d_offset = Upload(L.offset.data(), L.offset.size());
  }
  bool scale_ok = true;
  for (float v : L.scale) scale_ok = scale_ok && (!L.relu || v > 0.f);
  s.kp = RoundUp(L.in_dim, kGemmBlockK);
  if (elem == kF32 || L.out_dim % 4 != 0 || !scale_ok || 3 * (s.kp / kGemmBlockK) > kMaxKBlocks) return;
  // tensor path: BatchNorm scale folded into W and b; Wh = round16(W), Wl = round16(W - Wh)
  s.n_pad = RoundUp(L.out_dim, 64);
  const size_t kc = static_cast<size_t>(3) * s.kp;
  std::vector<float> wf(static_cast<size_t>(s.n_pad) * s.kp, 0.f);
  std::vector<float2> post(s.n_pad, make_float2(0.f, 0.f));
  for (int32_t n = 0; n < L.out_dim; ++n) {
    const float sc = L.scale.empty() ? 1.f : L.scale[n];
    for (int32_t k = 0; k < L.in_dim; ++k)
      wf[static_cast<size_t>(n) * s.kp + k] = L.weight[static_cast<size_t>(n) * L.in_dim + k] * sc;
    post[n] = make_float2((L.bias.empty() ? 0.f : L.bias[n]) * sc, L.offset.empty() ? 0.f : L.offset[n]);
  }
  std::vector<uint8_t> h16(wf.size() * 2), l16(wf.size() * 2);
  ConvertToElem(wf.data(), wf.size(), elem, h16.data());
  std::vector<float> resid(wf.size());
  for (size_t i = 0; i < resid.size(); ++i) resid[i] = wf[i] - ElemToFloat(h16.data(), i, elem);
  ConvertToElem(resid.data(), resid.size(), elem, l16.data());
  std::vector<uint8_t> wcat(static_cast<size_t>(s.n_pad) * kc * 2);
  const size_t row_bytes = static_cast<size_t>(s.kp) * 2;
  for (int32_t n = 0; n < s.n_pad; ++n) {
    uint8_t* row = wcat.data() + static_cast<size_t>(n) * kc * 2;
    std::memcpy(row, h16.data() + n * row_bytes, row_bytes);
    std::memcpy(row + row_bytes, h16.data() + n * row_bytes, row_bytes);
    std::memcpy(row + 2 * row_bytes, l16.data() + n * row_bytes, row_bytes);
  }
  const int32_t k_real = RoundUp(L.in_dim, 16);
  for (int pass = 0; pass < 3; ++pass)
    for (int32_t j = 0; j * kGemmBlockK < k_real; ++j)
      s.kb.push_back({pass == 1 ? 1 : 0, j * kGemmBlockK, 0, std::min(kGemmBlockK, k_real - j * kGemmBlockK) / 16});
  // every pass must cover the same kp/64 blocks so that W's column offset is kb*64
  if (static_cast<int32_t>(s.kb.size()) != 3 * (s.kp / kGemmBlockK)) {
    s.kb.clear();
    return;
  }
  s.d_wcat = Upload(wcat.data(), wcat.size());
  s.d_post = Upload(post.data(), post.size());
  s.tensor = true;
}

void FreeSegmentOp(SegmentOp* op) noexcept {
  cudaFree(op->d_wcat);
  cudaFree(op->d_post);
  cudaFree(op->d_w);
  cudaFree(op->d_b);
  cudaFree(op->d_scale);
  cudaFree(op->d_offset);
  *op = SegmentOp();
}

void RunSegmentOp(const SegmentOp& s, int32_t elem, int num_sms, const float* in, int32_t ld_in, int64_t m, float* out,
                  int32_t ld_out, void* hi, void* lo, int64_t split_rows, cudaStream_t stream) {
  if (m <= 0) return;
  if (!s.tensor) {
    CVX_CUDA_CHECK(LaunchSegmentAffine(in, ld_in, m, s.in_dim, s.d_w, s.d_b, s.relu ? 1 : 0, s.d_scale, s.d_offset, s.out_dim,
                                       out, ld_out, stream));
    return;
  }
  CVX_CUDA_CHECK(LaunchSplitF32(in, ld_in, m, s.in_dim, hi, lo, s.kp, elem, stream));
  TdnnGemmArgs a;
  std::memset(&a, 0, sizeof(a));
  Encode2D(&a.a_maps[0], elem, hi, split_rows, s.kp, kGemmBlockK, kGemmBlockM);
  Encode2D(&a.a_maps[1], elem, lo, split_rows, s.kp, kGemmBlockK, kGemmBlockM);
  a.a_maps[2] = a.a_maps[3] = a.a_maps[0];
  a.b_box_rows = std::min(kGemmBlockN, s.n_pad);
  a.b_box_rows_pair = std::min(kGemmBlockN / 2, s.n_pad);
  Encode2D(&a.b_map, elem, s.d_wcat, s.n_pad, 3 * s.kp, kGemmBlockK, a.b_box_rows);
  Encode2D(&a.b_map_pair, elem, s.d_wcat, s.n_pad, 3 * s.kp, kGemmBlockK, a.b_box_rows_pair);
  // fp32 output; rows beyond m and columns beyond out_dim are clipped by the map
  Encode2D(&a.c_map, kF32, out, m, ld_out, 32, 32, s.out_dim);
  a.num_kb = static_cast<int32_t>(s.kb.size());
  for (size_t i = 0; i < s.kb.size(); ++i) a.kb[i] = s.kb[i];
  a.n_pad = s.n_pad;
  a.n_mma = RoundUp(s.out_dim, 32);
  a.tile_n = kGemmBlockN;
  a.c_map_half = a.c_map;
  a.num_m_tiles = static_cast<int32_t>((m + kGemmBlockM - 1) / kGemmBlockM);
  a.num_n_tiles = (a.n_mma + kGemmBlockN - 1) / kGemmBlockN;
  a.post = s.d_post;
  a.relu = s.relu ? 1 : 0;
  a.ab_format = elem == kBF16 ? 1 : 0;
  a.out_f32 = 1;
  a.num_stages = 64;
  CVX_CUDA_CHECK(LaunchTdnnGemmPair(a, num_sms, stream));
}

void TestSegmentAffine(int32_t device, int32_t precision, const float* in, int64_t m, int32_t k, const float* w,
                       const float* bias, int32_t relu, const float* scale, const float* offset, int32_t n, float* out) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) throw CudaError("no usable CUDA device; this path has no CPU fallback");
  if (device < 0 || device >= ndev) throw InvalidArgument("device ordinal out of range");
  if (m <= 0 || k <= 0 || n <= 0 || !in || !w || !out) throw InvalidArgument("bad segment-layer test arguments");
  if ((scale == nullptr) != (offset == nullptr)) throw InvalidArgument("scale and offset must come together");
  CVX_CUDA_CHECK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CVX_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  const int32_t elem = precision == CVX_PRECISION_FP16 ? kF16 : precision == CVX_PRECISION_BF16 ? kBF16 : kF32;
  Layer L;
  L.in_dim = k;
  L.out_dim = n;
  L.weight.assign(w, w + static_cast<size_t>(n) * k);
  if (bias) L.bias.assign(bias, bias + n);
  L.relu = relu != 0;
  if (scale) {
    L.scale.assign(scale, scale + n);
    L.offset.assign(offset, offset + n);
  }
  SegmentOp op;
  float *d_in = nullptr, *d_out = nullptr;
  void *hi = nullptr, *lo = nullptr;
  auto cleanup = [&]() {
    FreeSegmentOp(&op);
    cudaFree(d_in);
    cudaFree(d_out);
    cudaFree(hi);
    cudaFree(lo);
  };
  try {
    BuildSegmentOp(L, elem, &op);
    d_in = Upload(in, static_cast<size_t>(m) * k);
    CVX_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&d_out), static_cast<size_t>(m) * n * 4));
    const int64_t split_rows = (m + 2 * kGemmBlockM - 1) / (2 * kGemmBlockM) * (2 * kGemmBlockM);
    if (op.tensor) {
      CVX_CUDA_CHECK(cudaMalloc(&hi, static_cast<size_t>(split_rows) * op.kp * 2));
      CVX_CUDA_CHECK(cudaMalloc(&lo, static_cast<size_t>(split_rows) * op.kp * 2));
      CVX_CUDA_CHECK(cudaMemset(hi, 0, static_cast<size_t>(split_rows) * op.kp * 2));
      CVX_CUDA_CHECK(cudaMemset(lo, 0, static_cast<size_t>(split_rows) * op.kp * 2));
    }
    RunSegmentOp(op, elem, prop.multiProcessorCount, d_in, k, m, d_out, n, hi, lo, split_rows, nullptr);
    CVX_CUDA_CHECK(cudaMemcpy(out, d_out, static_cast<size_t>(m) * n * 4, cudaMemcpyDeviceToHost));
  } catch (...) {
    cleanup();
    throw;
  }
  cleanup();
}

}  // namespace cvx
