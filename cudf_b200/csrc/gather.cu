// gather.cu — fixed-width gather with fused validity gather.
// Replaces thrust::gather + valid_if_n_kernel of cpp/include/cudf/detail/gather.cuh:108-133,506-527,
// 627-675 (public API cpp/include/cudf/copying.hpp:81-126).
//   out[i] = src[map[i]] ; negative map values wrap once (i + n), as the public API documents;
//   NULLIFY: out-of-range rows become null; the output carries a mask iff the source has nulls or
//   NULLIFY is requested (gather.cuh:650-672).
// Each thread owns 4 consecutive output rows: one 128-bit load of the map, 4 independent random
// reads in flight, vector stores; the 4 validity bits are merged across 8-lane groups with shuffles
// so that one lane writes a full 32-bit mask word (no atomics, no second pass over the map).
#include "common.cuh"
#include "device_utils.cuh"

#include <algorithm>

namespace b2 {
namespace {

template <typename T, bool MASK>
__global__ void __launch_bounds__(256) gather_kernel(const T* __restrict__ src, const uint32_t* __restrict__ src_mask,
                                                     int64_t src_bit_offset, int32_t src_n, const int32_t* __restrict__ map,
                                                     int64_t n, bool check_bounds, T* __restrict__ out,
                                                     uint32_t* __restrict__ out_mask, unsigned long long* __restrict__ null_count)
{
  const int64_t ngroups = (n + 3) / 4;
  const int64_t ngroups_round = (ngroups + 31) / 32 * 32;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const bool map_aligned = (reinterpret_cast<uintptr_t>(map) & 15) == 0;
  unsigned long long nulls = 0;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups_round; g += stride) {
    const int64_t i0 = g * 4;
    int32_t m[4] = {0, 0, 0, 0};
    const int cnt = (int)max((int64_t)0, min((int64_t)4, n - i0));
    if (cnt == 4 && map_aligned) {
      int4 q = ld_nc_v4(map + i0);
      m[0] = q.x; m[1] = q.y; m[2] = q.z; m[3] = q.w;
    } else {
      for (int j = 0; j < cnt; ++j) m[j] = map[i0 + j];
    }
    T v[4];
    uint32_t vbits = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int32_t r = m[j];
      if (r < 0) r += src_n;
      bool ok = j < cnt;
      if (check_bounds) ok = ok && (r >= 0 && r < src_n);
      v[j] = ok ? src[r] : T(0);
      if (MASK) {
        bool valid = ok && (src_mask == nullptr || bit_is_set(src_mask, src_bit_offset + r));
        vbits |= (valid ? 1u : 0u) << j;
        if (j < cnt && !valid) ++nulls;
      }
    }
    if (cnt == 4 && ((reinterpret_cast<uintptr_t>(out + i0) & (sizeof(T) * 4 - 1)) == 0)) {
      if constexpr (sizeof(T) == 8) {
        int4 a, b;
        memcpy(&a, &v[0], 16);
        memcpy(&b, &v[2], 16);
        st_na_v4(out + i0, a);
        st_na_v4(out + i0 + 2, b);
      } else if constexpr (sizeof(T) == 4) {
        int4 a;
        memcpy(&a, &v[0], 16);
        st_na_v4(out + i0, a);
      } else if constexpr (sizeof(T) == 2) {
        uint2 a;
        memcpy(&a, &v[0], 8);
        *reinterpret_cast<uint2*>(out + i0) = a;
      } else {
        uint32_t a;
        memcpy(&a, &v[0], 4);
        *reinterpret_cast<uint32_t*>(out + i0) = a;
      }
    } else {
      for (int j = 0; j < cnt; ++j) out[i0 + j] = v[j];
    }
    if (MASK) {
      // merge 8 lanes x 4 bits into one word
      uint32_t w = vbits << (4 * (lane_id() & 7));
      w |= __shfl_xor_sync(0xffffffffu, w, 1);
      w |= __shfl_xor_sync(0xffffffffu, w, 2);
      w |= __shfl_xor_sync(0xffffffffu, w, 4);
      if ((lane_id() & 7) == 0 && i0 < n) out_mask[i0 >> 5] = w;
    }
  }
  if (MASK) {
    nulls = warp_sum(nulls);
    if (lane_id() == 0 && nulls) atomicAdd(null_count, nulls);
  }
}

template <typename T>
void launch_gather(const b2_column_view& src, const int32_t* map, int64_t n, bool nullify, b2_column& out, bool with_mask,
                   cudaStream_t stream)
{
  const int64_t groups = (n + 3) / 4;
  int grid = (int)std::max<int64_t>(1, std::min<int64_t>((groups + 255) / 256, NUM_SMS_B200 * 16));
  const T* data = static_cast<const T*>(src.data) + src.offset;
  prof_scope ps("gather", stream);
  if (with_mask) {
    B2_LAUNCH((gather_kernel<T, true>), grid, 256, 0, stream, data, has_nulls(src) ? src.null_mask : nullptr,
              (int64_t)src.offset, src.size, map, n, nullify, out.data.as<T>(), out.mask.as<uint32_t>(),
              out.pending.as<unsigned long long>());
  } else {
    B2_LAUNCH((gather_kernel<T, false>), grid, 256, 0, stream, data, (const uint32_t*)nullptr, (int64_t)0, src.size, map, n,
              nullify, out.data.as<T>(), (uint32_t*)nullptr, (unsigned long long*)nullptr);
  }
}

}  // namespace

column_ptr gather_column(const b2_column_view& src, const int32_t* map, int32_t n, bool nullify_oob, cudaStream_t stream)
{
  const bool with_mask = has_nulls(src) || nullify_oob;
  auto out = make_column(src.type_id, n, with_mask, stream);
  if (n == 0) return out;
  if (with_mask) {
    out->pending = dbuf(sizeof(unsigned long long), stream);
    out->pending_stream = stream;
    out->null_count = -1;
    B2_CUDA_TRY(cudaMemsetAsync(out->pending.ptr, 0, sizeof(unsigned long long), stream));
  }
  switch (type_width(src.type_id)) {
    case 1: launch_gather<uint8_t>(src, map, n, nullify_oob, *out, with_mask, stream); break;
    case 2: launch_gather<uint16_t>(src, map, n, nullify_oob, *out, with_mask, stream); break;
    case 4: launch_gather<uint32_t>(src, map, n, nullify_oob, *out, with_mask, stream); break;
    case 8: launch_gather<uint64_t>(src, map, n, nullify_oob, *out, with_mask, stream); break;
    default: B2_FAIL(B2_ERR_DATA_TYPE, "gather: unsupported (non fixed-width) column type");
  }
  return out;
}

table_ptr gather_table(const std::vector<b2_column_view>& cols, const int32_t* map, int32_t n, bool nullify_oob,
                       cudaStream_t stream)
{
  auto t = std::make_unique<b2_table>();
  for (const auto& c : cols) t->cols.push_back(gather_column(c, map, n, nullify_oob, stream));
  return t;
}

}  // namespace b2
