// partition.cu — stable partition of table rows into P buckets (the bucket step either side of the
// NCCL all-to-all of the sharded sort / join; role of cudf::hash_partition,
// cpp/include/cudf/partitioning.hpp:103-145 / cpp/src/partitioning/partitioning.cu, and of the range
// partition by sampled splitters in python/cudf_polars/cudf_polars/streaming/sort.py:169-235).
//   mode 0 (range): bucket(row) = #splitters <= key  (P-1 ascending splitters of the key's type)
//   mode 1 (hash):  bucket(row) = mix64(key bits) % P
// Implementation: bucket ids (uint8) + per-bucket counts in one pass over the key column, a stable
// one-pass radix sorted_order of the ids (radix_sort.cu) as gather map, then the fused gather.
#include "common.cuh"
#include "device_utils.cuh"
#include "key_pack.cuh"

#include <algorithm>

namespace b2 {
namespace {

template <typename UK>
__device__ __forceinline__ uint64_t order_bits(UK raw, int kind)
{
  if (kind == (int)key_kind::SIGNED) return (uint64_t)twiddle_in<UK, key_kind::SIGNED>(raw);
  if (kind == (int)key_kind::FLOAT) {
    if constexpr (sizeof(UK) >= 4) return (uint64_t)twiddle_in<UK, key_kind::FLOAT>(raw);
  }
  return (uint64_t)raw;
}

constexpr int PT_TILE = 4096;  // rows per CTA tile (256 threads x 16 steps)

// pass 1: bucket id per row (uint8) + per-tile bucket counts
template <typename UK>
__global__ void __launch_bounds__(256) bucket_kernel(const UK* __restrict__ keys, int64_t n, int mode, int kind,
                                                     const UK* __restrict__ splitters, int P, uint8_t* __restrict__ ids,
                                                     uint32_t* __restrict__ tile_counts)
{
  __shared__ uint64_t s_split[256];
  __shared__ unsigned int s_cnt[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    s_cnt[i] = 0;
    s_split[i] = (mode == 0 && i < P - 1) ? order_bits<UK>(splitters[i], kind) : ~0ull;
  }
  __syncthreads();
  const int64_t tile = blockIdx.x;
  const int64_t end = min(n, (tile + 1) * (int64_t)PT_TILE);
  for (int64_t i = tile * PT_TILE + threadIdx.x; i < end; i += blockDim.x) {
    const UK raw = keys[i];
    int b;
    if (mode == 0) {
      const uint64_t k = order_bits<UK>(raw, kind);
      int lo = 0, hi = P - 1;  // number of splitters <= k
      while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (s_split[mid] <= k) lo = mid + 1; else hi = mid;
      }
      b = lo;
    } else {
      b = (int)(mix64((uint64_t)raw) % (uint64_t)P);
    }
    ids[i] = (uint8_t)b;
    atomicAdd(&s_cnt[b], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < P; i += blockDim.x) tile_counts[tile * P + i] = s_cnt[i];
}

// ---- cudf::hash_partition's row hash (cpp/src/partitioning/partitioning.cu:875-945, row_operator/hashing.cuh:40-140) ----
// MurmurHash3_x86_32 of a fixed-width value's little-endian bytes (cuco::murmurhash3_32 = the public algorithm; restated
// from its published description, not from cuCollections, which is not in the reference tree).
__device__ __forceinline__ uint32_t murmur3_32_fixed(uint64_t bits, int len, uint32_t seed)
{
  constexpr uint32_t c1 = 0xcc9e2d51u, c2 = 0x1b873593u;
  uint32_t h = seed;
  auto mixk = [&](uint32_t k) {
    k *= c1;
    k = (k << 15) | (k >> 17);
    return k * c2;
  };
  if (len >= 4) {
    h ^= mixk((uint32_t)bits);
    h = ((h << 13) | (h >> 19)) * 5u + 0xe6546b64u;
    if (len == 8) {
      h ^= mixk((uint32_t)(bits >> 32));
      h = ((h << 13) | (h >> 19)) * 5u + 0xe6546b64u;
    }
  } else {
    h ^= mixk((uint32_t)bits & (len == 1 ? 0xffu : 0xffffu));  // tail bytes only
  }
  h ^= (uint32_t)len;
  h ^= h >> 16; h *= 0x85ebca6bu;
  h ^= h >> 13; h *= 0xc2b2ae35u;
  return h ^ (h >> 16);
}

// hash_fn: 0 = IdentityHash (integral keys only here), 1 = MurmurHash3_x86_32.  Floats hash their normalised value (-0 -> +0,
// NaN -> canonical quiet NaN: key_col_bits does exactly that), BOOL8 hashes one byte 0 / 1, a null hashes to UINT32_MAX,
// the columns are folded left to right with hash_combine (hashing.hpp:83-86), the first column's hash being the start.
__device__ __forceinline__ uint32_t row_hash32(const key_cols& kc, int64_t r, int hash_fn, uint32_t seed, uint32_t bool_cols)
{
  uint32_t h = 0;
#pragma unroll 1
  for (int c = 0; c < kc.n; ++c) {
    const int64_t e = r + kc.offset[c];
    uint32_t hc = 0xffffffffu;
    if (kc.mask[c] == nullptr || bit_is_set(kc.mask[c], e)) {
      uint64_t bits = key_col_bits(kc, c, e);
      if ((bool_cols >> c) & 1u) bits = bits != 0;
      hc = hash_fn == 0 ? (uint32_t)bits : murmur3_32_fixed(bits, kc.width[c], seed);
    }
    h = c == 0 ? hc : (h ^ (hc + 0x9e3779b9u + (h << 6) + (h >> 2)));
  }
  return h;
}

__global__ void __launch_bounds__(256) bucket_rowhash_kernel(key_cols kc, int64_t n, int P, int hash_fn, uint32_t seed, uint32_t bool_cols,
                                                             uint8_t* __restrict__ ids, uint32_t* __restrict__ tile_counts)
{
  __shared__ unsigned int s_cnt[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_cnt[i] = 0;
  __syncthreads();
  const int64_t tile = blockIdx.x;
  const int64_t end = min(n, (tile + 1) * (int64_t)PT_TILE);
  for (int64_t i = tile * PT_TILE + threadIdx.x; i < end; i += blockDim.x) {
    const int b = (int)(row_hash32(kc, i, hash_fn, seed, bool_cols) % (uint32_t)P);
    ids[i] = (uint8_t)b;
    atomicAdd(&s_cnt[b], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < P; i += blockDim.x) tile_counts[tile * P + i] = s_cnt[i];
}

// pass 2 (one CTA): tile_counts[t][b] -> global start of bucket b's rows of tile t; totals[b]
__global__ void __launch_bounds__(1024) tile_scan_kernel(uint32_t* __restrict__ tile_counts, int64_t ntiles, int P,
                                                         unsigned long long* __restrict__ totals)
{
  __shared__ uint32_t wsum[32];
  __shared__ uint32_t carry;
  __shared__ uint32_t bucket_start;
  if (threadIdx.x == 0) bucket_start = 0;
  __syncthreads();
  for (int b = 0; b < P; ++b) {
    if (threadIdx.x == 0) carry = bucket_start;
    __syncthreads();
    for (int64_t t0 = 0; t0 < ntiles; t0 += 1024) {
      const int64_t t = t0 + threadIdx.x;
      const uint32_t v = t < ntiles ? tile_counts[t * P + b] : 0u;
      uint32_t inc = warp_inclusive_sum(v);
      if (lane_id() == 31) wsum[threadIdx.x >> 5] = inc;
      __syncthreads();
      uint32_t woff = 0;
      for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) woff += wsum[w];
      const uint32_t c = carry;
      if (t < ntiles) tile_counts[t * P + b] = c + woff + inc - v;
      __syncthreads();
      if (threadIdx.x == 1023) carry = c + woff + inc;
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      totals[b] = (unsigned long long)(carry - bucket_start);
      bucket_start = carry;
    }
    __syncthreads();
  }
}

// pass 3: stable destination of every row: dest[row] = start(tile, bucket) + rank of the row among the
// tile's rows of the same bucket (warp counts by smem atomics, MATCH.ANY ranking: few distinct ids)
__global__ void __launch_bounds__(256) dest_kernel(const uint8_t* __restrict__ ids, int64_t n, int P,
                                                   const uint32_t* __restrict__ tile_starts, int32_t* __restrict__ dest)
{
  __shared__ uint32_t woff[8][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = lane; i < P; i += 32) woff[warp][i] = 0;
  __syncwarp();
  const int64_t tile = blockIdx.x;
  const int64_t wbase = tile * PT_TILE + (int64_t)warp * (PT_TILE / 8);
  uint8_t id[PT_TILE / 8 / 32];
#pragma unroll
  for (int s = 0; s < PT_TILE / 8 / 32; ++s) {
    const int64_t r = wbase + s * 32 + lane;
    id[s] = r < n ? ids[r] : (uint8_t)255;
    if (r < n) atomicAdd(&woff[warp][id[s]], 1u);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < P; b += blockDim.x) {
    uint32_t run = tile_starts[tile * P + b];
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const uint32_t c = woff[w][b];
      woff[w][b] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int s = 0; s < PT_TILE / 8 / 32; ++s) {
    const int64_t r = wbase + s * 32 + lane;
    const bool in = r < n;
    const unsigned peers = __match_any_sync(0xffffffffu, in ? (unsigned)id[s] : 0x100u);
    const unsigned lt = __popc(peers & lanemask_lt());
    uint32_t prev = 0;
    if (in && lt == 0) {
      prev = woff[warp][id[s]];
      woff[warp][id[s]] = prev + __popc(peers);
    }
    __syncwarp();
    prev = __shfl_sync(0xffffffffu, prev, __ffs(peers) - 1);
    if (in) dest[r] = (int32_t)(prev + lt);
  }
}

struct dest_table {
  void* ptr[128];            // destination base address per bucket
  uint32_t bucket_start[128];  // global start of the bucket in the plan's dest numbering
};
template <typename T>
__global__ void __launch_bounds__(256) scatter_to_kernel(const T* __restrict__ in, const uint8_t* __restrict__ ids,
                                                         const int32_t* __restrict__ dest, int64_t n, dest_table dt)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int b = ids[i];
    static_cast<T*>(dt.ptr[b])[(uint32_t)dest[i] - dt.bucket_start[b]] = ld_stream(in + i);
  }
}

// Staged variant for peer destinations (EXPERIMENTAL in round 1: compiled, not yet run on hardware, not the
// default): one CTA per 4096-row tile stages the tile's rows grouped by bucket in shared memory and then writes
// each (tile, bucket) run with consecutive threads, so that a peer sees runs of ~4096/P rows (4 KB at P = 8)
// instead of the ~4-row segments of scatter_to_kernel (measured 60 GB/s over NVLink at 8 ranks).
struct dest_table_staged {
  void* ptr[128];
  uint32_t bucket_start[129];  // [P] = total rows
};
template <typename T>
__global__ void __launch_bounds__(256) scatter_to_staged_kernel(const T* __restrict__ in, const uint8_t* __restrict__ ids,
                                                                const int32_t* __restrict__ dest,
                                                                const uint32_t* __restrict__ tile_starts, int64_t n, int P,
                                                                int64_t ntiles, dest_table_staged dt)
{
  __shared__ T stage[PT_TILE];
  __shared__ uint8_t sb[PT_TILE];
  __shared__ uint32_t gbase[128];  // global start of this tile's rows of bucket b
  __shared__ uint32_t soff[129];   // staged offset of bucket b inside the tile
  const int64_t tile = blockIdx.x;
  const int64_t row0 = tile * PT_TILE;
  const int rows = (int)min((int64_t)PT_TILE, n - row0);
  for (int b = threadIdx.x; b < P; b += blockDim.x) {
    const uint32_t g0 = tile_starts[tile * P + b];
    const uint32_t g1 = (tile + 1 < ntiles) ? tile_starts[(tile + 1) * P + b] : dt.bucket_start[b + 1];
    gbase[b] = g0;
    soff[b + 1] = g1 - g0;  // count, turned into offsets below
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    soff[0] = 0;
    for (int b = 0; b < P; ++b) {
      const uint32_t c = soff[b + 1];
      soff[b + 1] = run + c;
      if (b == 0) soff[0] = 0;
      run += c;
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < rows; j += blockDim.x) {
    const int64_t r = row0 + j;
    const int b = ids[r];
    const uint32_t p = soff[b] + ((uint32_t)dest[r] - gbase[b]);
    stage[p] = ld_stream(in + r);
    sb[p] = (uint8_t)b;
  }
  __syncthreads();
  for (int q = threadIdx.x; q < rows; q += blockDim.x) {
    const int b = sb[q];
    static_cast<T*>(dt.ptr[b])[(gbase[b] - dt.bucket_start[b]) + ((uint32_t)q - soff[b])] = stage[q];
  }
}

template <typename T>
__global__ void __launch_bounds__(256) scatter_kernel(const T* __restrict__ in, const int32_t* __restrict__ dest, int64_t n,
                                                      T* __restrict__ out)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[dest[i]] = ld_stream(in + i);
}

}  // namespace

// ---- more than 256 partitions: ids as a 32-bit column, stable radix order of the ids (radix_sort.cu), fused gather ----
__global__ void __launch_bounds__(256) rowhash_ids32_kernel(key_cols kc, int64_t n, uint32_t P, int hash_fn, uint32_t seed, uint32_t bool_cols,
                                                            uint32_t* __restrict__ ids)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) ids[i] = row_hash32(kc, i, hash_fn, seed, bool_cols) % P;
}
template <typename M>
__global__ void __launch_bounds__(256) map_ids32_kernel(const M* __restrict__ map, int64_t n, uint32_t* __restrict__ ids)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) ids[i] = (uint32_t)map[i];
}
// offsets[p] = number of rows whose id is < p (ids read through the sorted order), p in [0, P]
__global__ void __launch_bounds__(256) id_offsets_kernel(const uint32_t* __restrict__ ids, const int32_t* __restrict__ order, int64_t n, uint32_t P,
                                                         int32_t* __restrict__ offsets)
{
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p > (int64_t)P) return;
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = lo + (hi - lo) / 2;
    if (ids[order[mid]] < (uint32_t)p) lo = mid + 1;
    else hi = mid;
  }
  offsets[p] = (int32_t)lo;
}

static table_ptr partition_by_ids32(const std::vector<b2_column_view>& input, const dbuf& ids, int64_t n, int P, int32_t* out_offsets,
                                    cudaStream_t stream)
{
  b2_column_view idv{B2_UINT32, (int32_t)n, ids.ptr, nullptr, 0, 0};
  auto order = sorted_order({idv}, {}, {}, true, stream);
  auto out = gather_table(input, order->data.as<int32_t>(), (int32_t)n, false, stream);
  dbuf offs(sizeof(int32_t) * ((size_t)P + 1), stream);
  B2_LAUNCH(id_offsets_kernel, (unsigned)((P + 1 + 255) / 256), 256, 0, stream, ids.as<uint32_t>(), order->data.as<int32_t>(), n, (uint32_t)P,
            offs.as<int32_t>());
  B2_CUDA_TRY(cudaMemcpyAsync(out_offsets, offs.ptr, sizeof(int32_t) * ((size_t)P + 1), cudaMemcpyDeviceToHost, stream));
  B2_CUDA_TRY(cudaStreamSynchronize(stream));
  return out;
}

// common tail of every partition flavour: `buckets(ids, tile_counts)` launches the kernel that writes one bucket id per row
// and the per-tile bucket counts
template <typename BucketFn>
static table_ptr partition_rows(const std::vector<b2_column_view>& input, int64_t n, int P, int32_t* out_offsets, cudaStream_t stream,
                                BucketFn&& buckets)
{
  B2_EXPECTS(P >= 1 && P <= 256, B2_ERR_INVALID_ARGUMENT, "num_partitions must be in [1, 256]");
  for (auto& c : input) B2_EXPECTS(c.size == n, B2_ERR_LOGIC, "Column size mismatch.");
  const int64_t ntiles = (n + PT_TILE - 1) / PT_TILE;
  dbuf ids(std::max<int64_t>(n, 1), stream), totals(sizeof(unsigned long long) * 256, stream);
  dbuf tile_counts(sizeof(uint32_t) * std::max<int64_t>(ntiles, 1) * P, stream);
  B2_CUDA_TRY(cudaMemsetAsync(totals.ptr, 0, totals.bytes, stream));
  auto out = std::make_unique<b2_table>();
  if (n > 0) {
    {
      prof_scope ps("partition_bucket", stream);
      buckets(ids.as<uint8_t>(), tile_counts.as<uint32_t>(), (unsigned)ntiles);
    }
    B2_LAUNCH(tile_scan_kernel, 1, 1024, 0, stream, tile_counts.as<uint32_t>(), ntiles, P, totals.as<unsigned long long>());
    bool any_nullable = false;
    for (auto& c : input) any_nullable |= has_nulls(c);
    dbuf dest(sizeof(int32_t) * n, stream);
    {
      prof_scope ps("partition_dest", stream);
      B2_LAUNCH(dest_kernel, (unsigned)ntiles, 256, 0, stream, ids.as<uint8_t>(), n, P, tile_counts.as<uint32_t>(), dest.as<int32_t>());
    }
    if (!any_nullable) {
      prof_scope ps("partition_scatter", stream);
      const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, NUM_SMS_B200 * 16));
      for (auto& c : input) {
        auto oc = make_column(c.type_id, (int32_t)n, false, stream);
        dispatch_width(type_width(c.type_id), [&](auto tag) {
          using T = decltype(tag);
          B2_LAUNCH((scatter_kernel<T>), grid, 256, 0, stream, static_cast<const T*>(c.data) + c.offset, dest.as<int32_t>(), n, oc->data.as<T>());
        });
        out->cols.push_back(std::move(oc));
      }
    } else {
      // nullable payloads: invert through a stable one-pass radix order of the ids and the fused gather
      b2_column_view idv{B2_UINT8, (int32_t)n, ids.ptr, nullptr, 0, 0};
      auto order = sorted_order({idv}, {}, {}, true, stream);
      out = gather_table(input, order->data.as<int32_t>(), (int32_t)n, false, stream);
    }
  } else {
    for (auto& c : input) out->cols.push_back(make_column(c.type_id, 0, false, stream));
  }
  unsigned long long h[256];
  B2_CUDA_TRY(cudaMemcpyAsync(h, totals.ptr, sizeof(unsigned long long) * 256, cudaMemcpyDeviceToHost, stream));
  B2_CUDA_TRY(cudaStreamSynchronize(stream));
  out_offsets[0] = 0;
  for (int b = 0; b < P; ++b) out_offsets[b + 1] = out_offsets[b] + (int32_t)h[b];
  return out;
}

table_ptr partition_table(const std::vector<b2_column_view>& input, const b2_column_view& keys, int mode, const void* splitters,
                          int P, int32_t* out_offsets, cudaStream_t stream)
{
  B2_EXPECTS(!has_nulls(keys), B2_ERR_INVALID_ARGUMENT, "partition key column must not contain nulls");
  B2_EXPECTS(mode == 1 || P == 1 || splitters != nullptr, B2_ERR_INVALID_ARGUMENT, "range partition needs splitters");
  const int64_t n = keys.size;
  const int sid  = storage_type(keys.type_id);
  const int kind = is_float_id(sid) ? (int)key_kind::FLOAT : (is_signed_id(sid) ? (int)key_kind::SIGNED : (int)key_kind::UNSIGNED);
  return partition_rows(input, n, P, out_offsets, stream, [&](uint8_t* ids, uint32_t* tile_counts, unsigned grid) {
    dispatch_width(type_width(keys.type_id), [&](auto tag) {
      using T = decltype(tag);
      B2_LAUNCH((bucket_kernel<T>), grid, 256, 0, stream, static_cast<const T*>(keys.data) + keys.offset, n, mode, kind, static_cast<const T*>(splitters), P,
                ids, tile_counts);
    });
  });
}

// cudf::partition(t, partition_map, num_partitions) — cpp/src/partitioning/partitioning.cu:898-915: the map names each row's partition
table_ptr partition_by_map(const std::vector<b2_column_view>& input, const b2_column_view& map, int P, int32_t* out_offsets, cudaStream_t stream)
{
  B2_EXPECTS(is_integral_id(map.type_id) && map.type_id != B2_BOOL8, B2_ERR_LOGIC, "Unexpected, non-integral partition map.");
  B2_EXPECTS(!has_nulls(map), B2_ERR_LOGIC, "Unexpected null values in partition_map.");
  const int64_t n = map.size;
  for (auto& c : input) B2_EXPECTS(c.size == n, B2_ERR_LOGIC, "Size mismatch between table and partition map.");
  if (P <= 0 || n == 0) {
    auto out = std::make_unique<b2_table>();
    for (auto& c : input) out->cols.push_back(make_column(c.type_id, 0, false, stream));
    for (int b = 0; b <= std::max(P, 0); ++b) out_offsets[b] = 0;
    return out;
  }
  dbuf ids(sizeof(uint32_t) * n, stream);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, NUM_SMS_B200 * 16));
  dispatch_width(type_width(map.type_id), [&](auto tag) {
    using M = decltype(tag);
    B2_LAUNCH((map_ids32_kernel<M>), grid, 256, 0, stream, static_cast<const M*>(map.data) + map.offset, n, ids.as<uint32_t>());
  });
  return partition_by_ids32(input, ids, n, P, out_offsets, stream);
}

// cudf::hash_partition(input, keys, num_partitions, hash_function, seed) — cpp/src/partitioning/partitioning.cu:875-945
table_ptr hash_partition_table(const std::vector<b2_column_view>& input, const std::vector<b2_column_view>& keys, int P, int hash_fn,
                               uint32_t seed, int32_t* out_offsets, cudaStream_t stream)
{
  const int64_t n = input.empty() ? 0 : input[0].size;
  B2_EXPECTS(hash_fn == 0 || hash_fn == 1, B2_ERR_LOGIC, "Unsupported hash function in hash_partition");
  B2_EXPECTS(keys.empty() || keys[0].size == n, B2_ERR_INVALID_ARGUMENT,
             "Input table and key table must have same number of rows, or key table should have no columns.");
  if (P <= 0 || n == 0 || keys.empty()) {  // empty result with num_partitions + 1 zero offsets
    auto out = std::make_unique<b2_table>();
    for (auto& c : input) out->cols.push_back(make_column(c.type_id, 0, false, stream));
    for (int b = 0; b <= std::max(P, 0); ++b) out_offsets[b] = 0;
    return out;
  }
  uint32_t bool_cols = 0;
  for (size_t c = 0; c < keys.size(); ++c) {
    if (hash_fn == 0) B2_EXPECTS(is_integral_id(storage_type(keys[c].type_id)), B2_ERR_LOGIC, "IdentityHash does not support this data type");
    if (keys[c].type_id == B2_BOOL8) bool_cols |= 1u << c;
  }
  const key_cols kc = make_key_cols(keys, true);
  if (P > 256) {
    dbuf ids(sizeof(uint32_t) * n, stream);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, NUM_SMS_B200 * 16));
    B2_LAUNCH(rowhash_ids32_kernel, grid, 256, 0, stream, kc, n, (uint32_t)P, hash_fn, seed, bool_cols, ids.as<uint32_t>());
    return partition_by_ids32(input, ids, n, P, out_offsets, stream);
  }
  return partition_rows(input, n, P, out_offsets, stream, [&](uint8_t* ids, uint32_t* tile_counts, unsigned grid) {
    B2_LAUNCH(bucket_rowhash_kernel, grid, 256, 0, stream, kc, n, P, hash_fn, seed, bool_cols, ids, tile_counts);
  });
}

}  // namespace b2

extern "C" b2_status b2_partition_by_map(const b2_table_view* input, const b2_column_view* partition_map, int32_t num_partitions, b2_stream stream,
                                          b2_table** out, int32_t* out_offsets)
{
  B2_TRY_BEGIN
    B2_EXPECTS(input && partition_map && out && out_offsets, B2_ERR_INVALID_ARGUMENT, "null argument");
    std::vector<b2_column_view> cols;
    b2::validate_table(input, cols);
    b2::validate_column(*partition_map);
    *out = b2::partition_by_map(cols, *partition_map, num_partitions, out_offsets, static_cast<cudaStream_t>(stream)).release();
  B2_TRY_END
}

extern "C" b2_status b2_hash_partition(const b2_table_view* input, const b2_table_view* keys, int32_t num_partitions, int32_t hash_function,
                                       uint32_t seed, b2_stream stream, b2_table** out, int32_t* out_offsets)
{
  B2_TRY_BEGIN
    B2_EXPECTS(input && keys && out && out_offsets, B2_ERR_INVALID_ARGUMENT, "null argument");
    std::vector<b2_column_view> cols, kcols;
    b2::validate_table(input, cols);
    B2_EXPECTS(keys->num_columns >= 0 && (keys->num_columns == 0 || keys->columns != nullptr), B2_ERR_INVALID_ARGUMENT, "invalid table_view");
    kcols.assign(keys->columns, keys->columns + keys->num_columns);
    for (auto& c : kcols) b2::validate_column(c);
    *out = b2::hash_partition_table(cols, kcols, num_partitions, hash_function, seed, out_offsets, static_cast<cudaStream_t>(stream)).release();
  B2_TRY_END
}

extern "C" b2_status b2_partition(const b2_table_view* input, const b2_column_view* keys, int32_t mode, const void* splitters,
                                  int32_t num_partitions, b2_stream stream, b2_table** out, int32_t* out_offsets)
{
  B2_TRY_BEGIN
    B2_EXPECTS(input && keys && out && out_offsets, B2_ERR_INVALID_ARGUMENT, "null argument");
    std::vector<b2_column_view> cols;
    b2::validate_table(input, cols);
    b2::validate_column(*keys);
    *out = b2::partition_table(cols, *keys, mode, splitters, num_partitions, out_offsets, static_cast<cudaStream_t>(stream)).release();
  B2_TRY_END
}

// ---- two-phase partition (plan + scatter to arbitrary destinations) and CUDA-IPC buffers --------------------
struct b2_partition_plan {
  int64_t n = 0;
  int32_t P = 0;
  b2::dbuf ids, dest;
  b2::dbuf tile_starts;  // [ntiles][P] global start of the rows of (tile, bucket) in the plan's dest numbering
  uint32_t bucket_start[257] = {0};
};

namespace b2 {
static std::unique_ptr<b2_partition_plan> make_plan(const b2_column_view& keys, int mode, const void* splitters, int P,
                                                    int64_t* out_counts, cudaStream_t stream)
{
  B2_EXPECTS(P >= 1 && P <= 128, B2_ERR_INVALID_ARGUMENT, "num_partitions must be in [1, 128]");
  B2_EXPECTS(!has_nulls(keys), B2_ERR_INVALID_ARGUMENT, "partition key column must not contain nulls");
  B2_EXPECTS(mode == 1 || P == 1 || splitters != nullptr, B2_ERR_INVALID_ARGUMENT, "range partition needs splitters");
  auto plan = std::make_unique<b2_partition_plan>();
  const int64_t n = keys.size;
  plan->n = n;
  plan->P = P;
  for (int b = 0; b <= P; ++b) plan->bucket_start[b] = 0;
  if (n == 0) {
    for (int b = 0; b < P; ++b) out_counts[b] = 0;
    return plan;
  }
  const int64_t ntiles = (n + PT_TILE - 1) / PT_TILE;
  plan->ids  = dbuf(n, stream);
  plan->dest = dbuf(sizeof(int32_t) * n, stream);
  dbuf totals(sizeof(unsigned long long) * 256, stream);
  plan->tile_starts = dbuf(sizeof(uint32_t) * ntiles * P, stream);
  dbuf& tile_counts = plan->tile_starts;
  B2_CUDA_TRY(cudaMemsetAsync(totals.ptr, 0, totals.bytes, stream));
  const int sid  = storage_type(keys.type_id);
  const int kind = is_float_id(sid) ? (int)key_kind::FLOAT : (is_signed_id(sid) ? (int)key_kind::SIGNED : (int)key_kind::UNSIGNED);
  {
    prof_scope ps("partition_bucket", stream);
    const unsigned grid = (unsigned)ntiles;
    dispatch_width(type_width(keys.type_id), [&](auto tag) {
      using T = decltype(tag);
      B2_LAUNCH((bucket_kernel<T>), grid, 256, 0, stream, static_cast<const T*>(keys.data) + keys.offset, n, mode, kind, static_cast<const T*>(splitters), P,
                plan->ids.as<uint8_t>(), tile_counts.as<uint32_t>());
    });
  }
  B2_LAUNCH(tile_scan_kernel, 1, 1024, 0, stream, tile_counts.as<uint32_t>(), ntiles, P, totals.as<unsigned long long>());
  {
    prof_scope ps("partition_dest", stream);
    B2_LAUNCH(dest_kernel, (unsigned)ntiles, 256, 0, stream, plan->ids.as<uint8_t>(), n, P, tile_counts.as<uint32_t>(), plan->dest.as<int32_t>());
  }
  unsigned long long h[256];
  B2_CUDA_TRY(cudaMemcpyAsync(h, totals.ptr, sizeof(unsigned long long) * 256, cudaMemcpyDeviceToHost, stream));
  B2_CUDA_TRY(cudaStreamSynchronize(stream));
  for (int b = 0; b < P; ++b) {
    out_counts[b] = (int64_t)h[b];
    plan->bucket_start[b + 1] = plan->bucket_start[b] + (uint32_t)h[b];
  }
  return plan;
}
}  // namespace b2

extern "C" {

b2_status b2_partition_plan_create(const b2_column_view* keys, int32_t mode, const void* splitters, int32_t num_partitions,
                                   b2_stream stream, b2_partition_plan** out, int64_t* out_counts)
{
  B2_TRY_BEGIN
    B2_EXPECTS(keys && out && out_counts, B2_ERR_INVALID_ARGUMENT, "null argument");
    b2::validate_column(*keys);
    *out = b2::make_plan(*keys, mode, splitters, num_partitions, out_counts, static_cast<cudaStream_t>(stream)).release();
  B2_TRY_END
}

b2_status b2_partition_scatter(const b2_partition_plan* plan, const b2_column_view* column, void* const* dest_ptrs, b2_stream stream)
{
  B2_TRY_BEGIN
    B2_EXPECTS(plan && column && dest_ptrs, B2_ERR_INVALID_ARGUMENT, "null argument");
    b2::validate_column(*column);
    B2_EXPECTS(column->size == plan->n, B2_ERR_LOGIC, "Column size mismatch.");
    B2_EXPECTS(!b2::has_nulls(*column), B2_ERR_INVALID_ARGUMENT, "b2_partition_scatter: nullable columns are not supported");
    if (plan->n == 0) return B2_OK;
    b2::dest_table dt{};
    for (int b = 0; b < plan->P; ++b) { dt.ptr[b] = dest_ptrs[b]; dt.bucket_start[b] = plan->bucket_start[b]; }
    const int64_t n = plan->n;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, b2::NUM_SMS_B200 * 16));
    auto s = static_cast<cudaStream_t>(stream);
    b2::prof_scope ps("partition_scatter_p2p", s);
    b2::dispatch_width(b2::type_width(column->type_id), [&](auto tag) {
      using T = decltype(tag);
      B2_LAUNCH((b2::scatter_to_kernel<T>), grid, 256, 0, s, static_cast<const T*>(column->data) + column->offset, plan->ids.as<uint8_t>(),
                plan->dest.as<int32_t>(), n, dt);
    });
  B2_TRY_END
}

void b2_partition_plan_free(b2_partition_plan* plan) { delete plan; }

b2_status b2_ipc_alloc(size_t bytes, void** out_ptr, uint8_t* out_handle64)
{
  B2_TRY_BEGIN
    B2_EXPECTS(out_ptr && out_handle64, B2_ERR_INVALID_ARGUMENT, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* p = nullptr;
    B2_CUDA_TRY(cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); B2_CUDA_TRY(e); }
    memcpy(out_handle64, &h, 64);
    *out_ptr = p;
  B2_TRY_END
}
b2_status b2_ipc_open(const uint8_t* handle64, void** out_ptr)
{
  B2_TRY_BEGIN
    B2_EXPECTS(out_ptr && handle64, B2_ERR_INVALID_ARGUMENT, "null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    B2_CUDA_TRY(cudaIpcOpenMemHandle(out_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  B2_TRY_END
}
b2_status b2_ipc_close(void* ptr)
{
  B2_TRY_BEGIN
  if (ptr) B2_CUDA_TRY(cudaIpcCloseMemHandle(ptr));
  B2_TRY_END
}
b2_status b2_ipc_free(void* ptr)
{
  B2_TRY_BEGIN
  if (ptr) B2_CUDA_TRY(cudaFree(ptr));
  B2_TRY_END
}

}  // extern "C"

namespace b2 {
namespace {
// grid-stride copy with the widest vector both ends allow: src and dst are aligned to G bytes RELATIVE to each other
// (bucket offsets are multiples of the element size, so 8-byte columns usually give G = 8 or 16); the few bytes before the
// first G-aligned address and after the last whole vector go one by one
template <typename V>
__global__ void __launch_bounds__(256) peer_copy_kernel(char* __restrict__ dst, const char* __restrict__ src, size_t bytes)
{
  constexpr size_t G = sizeof(V);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t head = (G - (reinterpret_cast<uintptr_t>(src) & (G - 1))) & (G - 1);
  if (head > bytes) head = bytes;
  const size_t nvec = (bytes - head) / G;
  for (size_t i = tid; i < head; i += stride) dst[i] = src[i];
  const V* s = reinterpret_cast<const V*>(src + head);
  V* d = reinterpret_cast<V*>(dst + head);
  for (size_t i = tid; i < nvec; i += stride) d[i] = s[i];
  for (size_t i = head + nvec * G + tid; i < bytes; i += stride) dst[i] = src[i];
}
}  // namespace
}  // namespace b2

// Range partition fused with the first radix pass machinery (radix_sort.cu::range_partition_*): counts, then ONE stable pass that
// writes every bucket's rows to its destination (peer) buffer.
extern "C" b2_status b2_range_partition_counts(const b2_column_view* keys, const void* splitters, int32_t num_partitions, b2_stream stream,
                                               int64_t* out_counts)
{
  B2_TRY_BEGIN
    B2_EXPECTS(keys && out_counts, B2_ERR_INVALID_ARGUMENT, "null argument");
    b2::validate_column(*keys);
    B2_EXPECTS(num_partitions >= 1 && num_partitions <= 256, B2_ERR_INVALID_ARGUMENT, "num_partitions must be in [1, 256]");
    B2_EXPECTS(b2::type_width(keys->type_id) == 8 && !b2::is_float_id(b2::storage_type(keys->type_id)) && !b2::has_nulls(*keys), B2_ERR_DATA_TYPE,
               "b2_range_partition_*: one null-free 8-byte integer-like key column");
    b2::range_partition_counts(*keys, splitters, num_partitions, out_counts, static_cast<cudaStream_t>(stream));
  B2_TRY_END
}
extern "C" b2_status b2_range_partition_scatter(const b2_column_view* keys, const b2_column_view* values, const void* splitters,
                                                int32_t num_partitions, void* const* key_dst, void* const* val_dst, b2_stream stream)
{
  B2_TRY_BEGIN
    B2_EXPECTS(keys && key_dst, B2_ERR_INVALID_ARGUMENT, "null argument");
    b2::validate_column(*keys);
    B2_EXPECTS(num_partitions >= 1 && num_partitions <= 256, B2_ERR_INVALID_ARGUMENT, "num_partitions must be in [1, 256]");
    B2_EXPECTS(b2::type_width(keys->type_id) == 8 && !b2::is_float_id(b2::storage_type(keys->type_id)) && !b2::has_nulls(*keys), B2_ERR_DATA_TYPE,
               "b2_range_partition_*: one null-free 8-byte integer-like key column");
    if (values) {
      b2::validate_column(*values);
      B2_EXPECTS(val_dst != nullptr, B2_ERR_INVALID_ARGUMENT, "null argument");
      B2_EXPECTS(values->size == keys->size, B2_ERR_LOGIC, "Column size mismatch.");
      const int vw = b2::type_width(values->type_id);
      B2_EXPECTS((vw == 4 || vw == 8) && !b2::has_nulls(*values), B2_ERR_DATA_TYPE, "b2_range_partition_scatter: one null-free 4- or 8-byte payload column");
    }
    b2::range_partition_scatter(*keys, values, splitters, num_partitions, key_dst, val_dst, static_cast<cudaStream_t>(stream));
  B2_TRY_END
}

extern "C" b2_status b2_peer_copy(void* dst, const void* src, size_t bytes, b2_stream stream)
{
  B2_TRY_BEGIN
    B2_EXPECTS(bytes == 0 || (dst && src), B2_ERR_INVALID_ARGUMENT, "null argument");
    if (bytes == 0) return B2_OK;
    const int grid = (int)std::max<size_t>(1, std::min<size_t>((bytes / 16 + 255) / 256, (size_t)b2::NUM_SMS_B200 * 8));
    b2::prof_scope ps("peer_copy", static_cast<cudaStream_t>(stream));
    const uintptr_t rel = reinterpret_cast<uintptr_t>(dst) ^ reinterpret_cast<uintptr_t>(src);  // low bits equal <=> same relative alignment
    char* d = static_cast<char*>(dst);
    const char* s = static_cast<const char*>(src);
    auto st = static_cast<cudaStream_t>(stream);
    if ((rel & 15) == 0) B2_LAUNCH((b2::peer_copy_kernel<uint4>), grid, 256, 0, st, d, s, bytes);
    else if ((rel & 7) == 0) B2_LAUNCH((b2::peer_copy_kernel<uint2>), grid, 256, 0, st, d, s, bytes);
    else if ((rel & 3) == 0) B2_LAUNCH((b2::peer_copy_kernel<uint32_t>), grid, 256, 0, st, d, s, bytes);
    else B2_LAUNCH((b2::peer_copy_kernel<uint8_t>), grid, 256, 0, st, d, s, bytes);
  B2_TRY_END
}

// EXPERIMENTAL (see scatter_to_staged_kernel): same contract as b2_partition_scatter.
extern "C" b2_status b2_partition_scatter_staged(const b2_partition_plan* plan, const b2_column_view* column, void* const* dest_ptrs,
                                                 b2_stream stream)
{
  B2_TRY_BEGIN
    B2_EXPECTS(plan && column && dest_ptrs, B2_ERR_INVALID_ARGUMENT, "null argument");
    b2::validate_column(*column);
    B2_EXPECTS(column->size == plan->n, B2_ERR_LOGIC, "Column size mismatch.");
    B2_EXPECTS(!b2::has_nulls(*column), B2_ERR_INVALID_ARGUMENT, "b2_partition_scatter_staged: nullable columns are not supported");
    if (plan->n == 0) return B2_OK;
    b2::dest_table_staged dt{};
    for (int b = 0; b < plan->P; ++b) { dt.ptr[b] = dest_ptrs[b]; dt.bucket_start[b] = plan->bucket_start[b]; }
    dt.bucket_start[plan->P] = plan->bucket_start[plan->P];
    const int64_t n = plan->n;
    const int64_t ntiles = (n + b2::PT_TILE - 1) / b2::PT_TILE;
    auto s = static_cast<cudaStream_t>(stream);
    b2::prof_scope ps("partition_scatter_p2p_staged", s);
    const uint32_t* ts = plan->tile_starts.as<uint32_t>();
    b2::dispatch_width(b2::type_width(column->type_id), [&](auto tag) {
      using T = decltype(tag);
      B2_LAUNCH((b2::scatter_to_staged_kernel<T>), (unsigned)ntiles, 256, 0, s, static_cast<const T*>(column->data) + column->offset,
                plan->ids.as<uint8_t>(), plan->dest.as<int32_t>(), ts, n, plan->P, ntiles, dt);
    });
  B2_TRY_END
}
