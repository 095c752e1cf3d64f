// bitmask.cu — null-mask utilities (Arrow validity bitmaps, LSB first, 32-bit words).
// Restates the behaviour of cpp/src/bitmask/null_mask.cu:48-86 (create_null_mask), :339-404
// (set_null_mask), :409-560 (copy_bitmask with bit offsets), count_set_bits, :608-735 (bitmask_and).
#include "common.cuh"
#include "device_utils.cuh"

#include <algorithm>

namespace b2 {
namespace {

__global__ void set_bits_kernel(uint32_t* __restrict__ mask, int64_t begin, int64_t end, uint32_t fill)
{
  // words [begin>>5, (end-1)>>5]; first and last word are partial
  const int64_t w0 = begin >> 5, w1 = (end - 1) >> 5;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t w = w0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w <= w1; w += stride) {
    uint32_t m = 0xffffffffu;
    if (w == w0) m &= 0xffffffffu << (begin & 31);
    if (w == w1) {
      int hi = (int)(end & 31);
      if (hi) m &= (1u << hi) - 1u;
    }
    if (m == 0xffffffffu) mask[w] = fill;
    else mask[w] = (mask[w] & ~m) | (fill & m);  // boundary words are owned by a single thread
  }
}

__global__ void count_bits_kernel(const uint32_t* __restrict__ mask, int64_t start, int64_t stop,
                                  unsigned long long* __restrict__ out)
{
  const int64_t nbits = stop - start;
  const int64_t nwords = (nbits + 31) / 32;
  const int64_t last_word = (stop - 1) >> 5;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  unsigned long long c = 0;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += stride) {
    uint32_t bits = load_mask_word_unaligned(mask, start + w * 32, last_word);
    int64_t rem = nbits - w * 32;
    if (rem < 32) bits &= (1u << rem) - 1u;
    c += __popc(bits);
  }
  c = warp_sum(c);
  if (lane_id() == 0 && c) atomicAdd(out, c);
}

__global__ void copy_bits_kernel(const uint32_t* __restrict__ src, int64_t begin, int64_t end, uint32_t* __restrict__ dst)
{
  const int64_t nbits = end - begin;
  const int64_t nwords = (nbits + 31) / 32;
  const int64_t last_word = (end - 1) >> 5;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += stride) {
    uint32_t bits = load_mask_word_unaligned(src, begin + w * 32, last_word);
    int64_t rem = nbits - w * 32;
    if (rem < 32) bits &= (1u << rem) - 1u;
    dst[w] = bits;
  }
}

struct and_args {
  const uint32_t* masks[16];
  int64_t offsets[16];
  int n;
};

__global__ void and_bits_kernel(and_args a, int64_t nbits, uint32_t* __restrict__ dst, unsigned long long* __restrict__ valid_count)
{
  const int64_t nwords = (nbits + 31) / 32;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  unsigned long long c = 0;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += stride) {
    uint32_t bits = 0xffffffffu;
    for (int i = 0; i < a.n; ++i) {
      const int64_t last_word = (a.offsets[i] + nbits - 1) >> 5;
      bits &= load_mask_word_unaligned(a.masks[i], a.offsets[i] + w * 32, last_word);
    }
    int64_t rem = nbits - w * 32;
    if (rem < 32) bits &= (1u << rem) - 1u;
    dst[w] = bits;
    c += __popc(bits);
  }
  c = warp_sum(c);
  if (lane_id() == 0 && c) atomicAdd(valid_count, c);
}

int grid_for(int64_t items, int block = 256)
{
  int64_t g = (items + block - 1) / block;
  return (int)std::max<int64_t>(1, std::min<int64_t>(g, NUM_SMS_B200 * 8));
}

}  // namespace

void set_null_mask(uint32_t* mask, int64_t begin, int64_t end, bool valid, cudaStream_t stream)
{
  B2_EXPECTS(begin >= 0 && begin <= end, B2_ERR_LOGIC, "Invalid bit range.");
  if (mask == nullptr || begin == end) return;
  const int64_t words = ((end - 1) >> 5) - (begin >> 5) + 1;
  B2_LAUNCH(set_bits_kernel, grid_for(words), 256, 0, stream, mask, begin, end, valid ? 0xffffffffu : 0u);
}

int32_t count_set_bits(const uint32_t* mask, int64_t start, int64_t stop, cudaStream_t stream)
{
  B2_EXPECTS(start >= 0 && start <= stop, B2_ERR_LOGIC, "Invalid bit range.");
  if (mask == nullptr) return 0;  // reference: a null bitmask pointer counts as zero set bits
  if (start == stop) return 0;
  dbuf cnt(sizeof(unsigned long long), stream);
  B2_CUDA_TRY(cudaMemsetAsync(cnt.ptr, 0, sizeof(unsigned long long), stream));
  B2_LAUNCH(count_bits_kernel, grid_for((stop - start + 31) / 32), 256, 0, stream, mask, start, stop,
            cnt.as<unsigned long long>());
  unsigned long long h = 0;
  B2_CUDA_TRY(cudaMemcpyAsync(&h, cnt.ptr, sizeof(h), cudaMemcpyDeviceToHost, stream));
  B2_CUDA_TRY(cudaStreamSynchronize(stream));
  return (int32_t)h;
}

dbuf copy_bitmask(const uint32_t* mask, int64_t begin, int64_t end, cudaStream_t stream)
{
  B2_EXPECTS(begin >= 0, B2_ERR_LOGIC, "Invalid range.");
  B2_EXPECTS(begin <= end, B2_ERR_LOGIC, "Invalid bit range.");
  if (mask == nullptr || begin == end) return dbuf{};
  dbuf out(bitmask_bytes(end - begin), stream);
  // zero the padding so that whole-word consumers never see garbage
  B2_CUDA_TRY(cudaMemsetAsync(out.ptr, 0, out.bytes, stream));
  B2_LAUNCH(copy_bits_kernel, grid_for((end - begin + 31) / 32), 256, 0, stream, mask, begin, end, out.as<uint32_t>());
  return out;
}

dbuf bitmask_and(const std::vector<b2_column_view>& cols, int32_t rows, int32_t* null_count, cudaStream_t stream)
{
  *null_count = 0;
  and_args a{};
  for (const auto& c : cols) {
    if (c.null_mask != nullptr && c.null_count > 0) {
      B2_EXPECTS(a.n < 16, B2_ERR_INVALID_ARGUMENT, "bitmask_and: at most 16 nullable columns supported");
      a.masks[a.n]   = c.null_mask;
      a.offsets[a.n] = c.offset;
      ++a.n;
    }
  }
  if (a.n == 0 || rows == 0) return dbuf{};
  dbuf out(bitmask_bytes(rows), stream);
  B2_CUDA_TRY(cudaMemsetAsync(out.ptr, 0, out.bytes, stream));
  dbuf cnt(sizeof(unsigned long long), stream);
  B2_CUDA_TRY(cudaMemsetAsync(cnt.ptr, 0, sizeof(unsigned long long), stream));
  B2_LAUNCH(and_bits_kernel, grid_for(num_words(rows)), 256, 0, stream, a, (int64_t)rows, out.as<uint32_t>(),
            cnt.as<unsigned long long>());
  unsigned long long h = 0;
  B2_CUDA_TRY(cudaMemcpyAsync(&h, cnt.ptr, sizeof(h), cudaMemcpyDeviceToHost, stream));
  B2_CUDA_TRY(cudaStreamSynchronize(stream));
  *null_count = rows - (int32_t)h;
  return out;
}

}  // namespace b2
