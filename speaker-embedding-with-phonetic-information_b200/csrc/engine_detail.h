// Internal helpers shared by the engine's translation units (engine*.cu).
#ifndef CVX_ENGINE_DETAIL_H_
#define CVX_ENGINE_DETAIL_H_

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstring>
#include <string>

#include "engine.h"

namespace cvx {
namespace detail {

// 2-D map over a row-major [rows, ld] matrix of `elem`, box = box_cols x box_rows, 128-byte swizzle
// (box_cols * elem size == 128) unless swizzle = false; out-of-bounds reads give zeros.
void Encode2D(CUtensorMap* map, int32_t elem, void* base, int64_t rows, int64_t ld, uint32_t box_cols, uint32_t box_rows,
              int64_t cols = -1, bool swizzle = true);
void ConvertToElem(const float* src, size_t n, int32_t elem, void* dst);
float ElemToFloat(const void* p, size_t i, int32_t elem);
inline int32_t RoundUp(int32_t v, int32_t m) { return (v + m - 1) / m * m; }

}  // namespace detail
}  // namespace cvx
#endif  // CVX_ENGINE_DETAIL_H_
