// datagen.cu — counter-based synthetic columns: x_i = splitmix64(seed + first + i).
// Used by bench.py and the GPU tests; the numpy twin lives in oracle/datagen.py.
#include "common.cuh"
#include "device_utils.cuh"

#include <algorithm>

namespace b2 {
namespace {
__global__ void fill_kernel(void* __restrict__ dst, int64_t n, uint64_t seed, int64_t first, int kind, uint64_t modulus)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if (kind == 4) {
      // validity words: word w holds bits for rows 32w .. 32w+31, bit b = top bit of the row's draw
      // (n = number of 32-bit words here)
      uint32_t w = 0;
      for (int b = 0; b < 32; ++b) w |= (uint32_t)(splitmix64(seed + (uint64_t)(first + i * 32 + b)) >> 63) << b;
      static_cast<uint32_t*>(dst)[i] = w;
      continue;
    }
    const uint64_t x = splitmix64(seed + (uint64_t)(first + i));
    switch (kind) {
      case 0: static_cast<uint64_t*>(dst)[i] = x; break;
      case 1: static_cast<double*>(dst)[i] = (double)(x >> 11) * (1.0 / 9007199254740992.0); break;
      case 2: static_cast<int64_t*>(dst)[i] = (int64_t)(x % modulus); break;
      case 3: static_cast<uint32_t*>(dst)[i] = (uint32_t)x; break;
      default: break;
    }
  }
}
}  // namespace
}  // namespace b2

extern "C" b2_status b2_fill_splitmix64(void* dst, int64_t n, uint64_t seed, int64_t first, int32_t kind, uint64_t modulus,
                                        b2_stream stream)
{
  try {
    B2_EXPECTS(kind >= 0 && kind <= 4, B2_ERR_INVALID_ARGUMENT, "unknown generator kind");
    B2_EXPECTS(kind != 2 || modulus != 0, B2_ERR_INVALID_ARGUMENT, "modulus must be non-zero");
    if (n <= 0) return B2_OK;
    int64_t items = kind == 4 ? (n + 31) / 32 : n;
    int grid = (int)std::max<int64_t>(1, std::min<int64_t>((items + 255) / 256, b2::NUM_SMS_B200 * 16));
    B2_LAUNCH(b2::fill_kernel, grid, 256, 0, static_cast<cudaStream_t>(stream), dst, items, seed, first, kind, modulus);
  } catch (const b2::error& e) {
    b2::set_last_error(e.what());
    return e.code;
  }
  return B2_OK;
}
