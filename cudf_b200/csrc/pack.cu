// pack.cu — cudf::pack / packed_size / pack_metadata / unpack for tables of fixed-width columns: the wire format of
// libcudf's contiguous_split (what dask_cudf / rapidsmpf style shufflers ship between workers).
//
// Replaces cpp/src/copying/pack.cpp:36-85,246-330 (serialized_table_header, serialized_column, unpack, metadata builder)
// and the single-partition case of cpp/src/copying/contiguous_split.cu (buffer order :505-560, split_align = 64 :50).
//   metadata (host bytes) = serialized_table_header {int32 version = 2, int32 num_columns, int32 num_rows, int32 pad}
//                           followed by one serialized_column per column, depth first:
//                           {int32 type_id, int32 scale, int32 size, int32 null_count, int64 data_offset,
//                            int64 null_mask_offset, int32 num_children, int32 pad}; offsets are into gpu_data, -1 = absent
//   gpu_data (device)     = for each column: [validity words, only if the column is nullable][data], every buffer padded
//                           to 64 bytes; a sliced input (offset != 0) is copied out with its mask re-based to bit 0.
// unpack allocates nothing: the views point into the caller's gpu_data.
#include "common.cuh"
#include "device_utils.cuh"

#include <cstring>

namespace b2 {
namespace {

constexpr int32_t PACKED_METADATA_VERSION = 2;  // cpp/include/cudf/detail/contiguous_split.hpp:127
constexpr size_t SPLIT_ALIGN = 64;              // contiguous_split.cu:50

struct table_header {
  int32_t version, num_columns, num_rows, pad;
};
struct column_entry {
  int32_t type_id, scale, size, null_count;
  int64_t data_offset, null_mask_offset;
  int32_t num_children, pad;
};
static_assert(sizeof(table_header) == 16 && sizeof(column_entry) == 40, "wire format of pack.cpp:36-85");

inline size_t round_up(size_t v) { return (v + SPLIT_ALIGN - 1) / SPLIT_ALIGN * SPLIT_ALIGN; }
// contiguous_split copies a nullable column's validity even when its null count is 0 (column_view::nullable())
inline bool nullable(const b2_column_view& c) { return c.null_mask != nullptr; }
inline size_t mask_bytes(int64_t rows) { return (size_t)((rows + 31) / 32) * 4; }

}  // namespace

size_t packed_size(const std::vector<b2_column_view>& cols)
{
  size_t total = 0;
  for (auto& c : cols) {
    if (c.size == 0) continue;
    if (nullable(c)) total += round_up(mask_bytes(c.size));
    total += round_up((size_t)c.size * type_width(c.type_id));
  }
  return total;
}

void pack_table(const std::vector<b2_column_view>& cols, int32_t num_rows, cudaStream_t stream, std::vector<uint8_t>& metadata, dbuf& gpu_data)
{
  const size_t bytes = packed_size(cols);
  gpu_data = dbuf(bytes, stream);
  metadata.assign(sizeof(table_header) + cols.size() * sizeof(column_entry), 0);
  table_header h{PACKED_METADATA_VERSION, (int32_t)cols.size(), cols.empty() ? num_rows : cols[0].size, 0};
  memcpy(metadata.data(), &h, sizeof(h));
  size_t off = 0;
  char* base = static_cast<char*>(gpu_data.ptr);
  for (size_t i = 0; i < cols.size(); ++i) {
    const auto& c = cols[i];
    column_entry e{c.type_id, 0, c.size, nullable(c) ? std::max(c.null_count, 0) : 0, -1, -1, 0, 0};
    if (c.size > 0) {
      if (nullable(c)) {
        const size_t mb = mask_bytes(c.size);
        dbuf m = copy_bitmask(c.null_mask, c.offset, (int64_t)c.offset + c.size, stream);  // re-based to bit 0
        B2_CUDA_TRY(cudaMemcpyAsync(base + off, m.ptr, mb, cudaMemcpyDeviceToDevice, stream));
        if (round_up(mb) > mb) B2_CUDA_TRY(cudaMemsetAsync(base + off + mb, 0, round_up(mb) - mb, stream));
        e.null_mask_offset = (int64_t)off;
        off += round_up(mb);
      }
      const size_t w = type_width(c.type_id), db = (size_t)c.size * w;
      B2_CUDA_TRY(cudaMemcpyAsync(base + off, static_cast<const char*>(c.data) + (size_t)c.offset * w, db, cudaMemcpyDeviceToDevice, stream));
      if (round_up(db) > db) B2_CUDA_TRY(cudaMemsetAsync(base + off + db, 0, round_up(db) - db, stream));
      e.data_offset = (int64_t)off;
      off += round_up(db);
    }
    memcpy(metadata.data() + sizeof(table_header) + i * sizeof(column_entry), &e, sizeof(e));
  }
}

// metadata describing columns that already live in [buffer, buffer + buffer_size) — cudf::pack_metadata (pack.cpp:262-272)
void pack_metadata(const std::vector<b2_column_view>& cols, int32_t num_rows, const uint8_t* buffer, size_t buffer_size, std::vector<uint8_t>& metadata)
{
  metadata.assign(sizeof(table_header) + cols.size() * sizeof(column_entry), 0);
  table_header h{PACKED_METADATA_VERSION, (int32_t)cols.size(), cols.empty() ? num_rows : cols[0].size, 0};
  memcpy(metadata.data(), &h, sizeof(h));
  for (size_t i = 0; i < cols.size(); ++i) {
    const auto& c = cols[i];
    B2_EXPECTS(c.offset == 0, B2_ERR_LOGIC, "pack_metadata: sliced columns cannot be described in place");
    column_entry e{c.type_id, 0, c.size, nullable(c) ? std::max(c.null_count, 0) : 0, -1, -1, 0, 0};
    if (c.size > 0 && c.data) {
      const uint8_t* p = static_cast<const uint8_t*>(c.data);
      B2_EXPECTS(p >= buffer && p < buffer + buffer_size, B2_ERR_LOGIC, "Encountered column data outside the range of input buffer");
      e.data_offset = p - buffer;
    }
    if (c.size > 0 && nullable(c)) {
      const uint8_t* p = reinterpret_cast<const uint8_t*>(c.null_mask);
      B2_EXPECTS(p >= buffer && p < buffer + buffer_size, B2_ERR_LOGIC, "Encountered column null mask outside the range of input buffer");
      e.null_mask_offset = p - buffer;
    }
    memcpy(metadata.data() + sizeof(table_header) + i * sizeof(column_entry), &e, sizeof(e));
  }
}

// cudf::unpack(metadata, gpu_data) — pack.cpp:246-296, with the bounds checks of packed_metadata_view (:100-128)
void unpack_table(const uint8_t* metadata, size_t metadata_size, const uint8_t* gpu_data, std::vector<b2_column_view>& out, int32_t& num_rows)
{
  B2_EXPECTS(metadata != nullptr, B2_ERR_LOGIC, "Encountered invalid packed column input");
  B2_EXPECTS(metadata_size >= sizeof(table_header), B2_ERR_LOGIC, "packed metadata access is out of bounds");
  table_header h;
  memcpy(&h, metadata, sizeof(h));
  B2_EXPECTS(h.version == PACKED_METADATA_VERSION, B2_ERR_LOGIC, "packed metadata has an unsupported format version");
  B2_EXPECTS(h.num_columns >= 0, B2_ERR_LOGIC, "packed metadata header has negative column count");
  B2_EXPECTS(h.num_rows >= 0, B2_ERR_LOGIC, "packed metadata header has negative row count");
  B2_EXPECTS(metadata_size == sizeof(table_header) + (size_t)h.num_columns * sizeof(column_entry), B2_ERR_LOGIC,
             "packed metadata size does not match its column count (nested columns are not supported on this path)");
  out.clear();
  for (int32_t i = 0; i < h.num_columns; ++i) {
    column_entry e;
    memcpy(&e, metadata + sizeof(table_header) + (size_t)i * sizeof(column_entry), sizeof(e));
    B2_EXPECTS(e.num_children == 0, B2_ERR_DATA_TYPE, "unpack: nested columns are not supported on this path");
    B2_EXPECTS(is_fixed_width(e.type_id), B2_ERR_DATA_TYPE, "unpack: only fixed-width columns are supported on this path");
    b2_column_view v{};
    v.type_id    = e.type_id;
    v.size       = e.size;
    v.data       = e.data_offset != -1 ? gpu_data + e.data_offset : nullptr;
    v.null_mask  = e.null_mask_offset != -1 ? reinterpret_cast<const uint32_t*>(gpu_data + e.null_mask_offset) : nullptr;
    v.null_count = e.null_count;
    v.offset     = 0;
    out.push_back(v);
  }
  num_rows = h.num_rows;
  if (!out.empty()) B2_EXPECTS(h.num_rows == out[0].size, B2_ERR_LOGIC, "packed metadata row count does not match the columns");
}

}  // namespace b2

using namespace b2;

extern "C" {

b2_status b2_packed_size(const b2_table_view* input, size_t* out_bytes)
{
  B2_TRY_BEGIN
    B2_EXPECTS(input && out_bytes, B2_ERR_INVALID_ARGUMENT, "null argument");
    std::vector<b2_column_view> cols;
    validate_table(input, cols);
    *out_bytes = packed_size(cols);
  B2_TRY_END
}

// metadata: caller-provided host buffer of metadata_capacity bytes (16 + 40 per column); *metadata_size receives the size
b2_status b2_pack(const b2_table_view* input, b2_stream stream, uint8_t* metadata, size_t metadata_capacity, size_t* metadata_size,
                  b2_buffer** gpu_data)
{
  B2_TRY_BEGIN
    B2_EXPECTS(input && metadata && metadata_size && gpu_data, B2_ERR_INVALID_ARGUMENT, "null argument");
    std::vector<b2_column_view> cols;
    validate_table(input, cols);
    std::vector<uint8_t> md;
    auto buf = std::make_unique<b2_buffer>();
    pack_table(cols, 0, static_cast<cudaStream_t>(stream), md, buf->buf);
    B2_EXPECTS(md.size() <= metadata_capacity, B2_ERR_INVALID_ARGUMENT, "metadata buffer too small");
    memcpy(metadata, md.data(), md.size());
    *metadata_size = md.size();
    *gpu_data = buf.release();
  B2_TRY_END
}

b2_status b2_pack_metadata(const b2_table_view* input, const uint8_t* contiguous_buffer, size_t buffer_size, uint8_t* metadata,
                           size_t metadata_capacity, size_t* metadata_size)
{
  B2_TRY_BEGIN
    B2_EXPECTS(input && metadata && metadata_size, B2_ERR_INVALID_ARGUMENT, "null argument");
    std::vector<b2_column_view> cols;
    validate_table(input, cols);
    std::vector<uint8_t> md;
    pack_metadata(cols, 0, contiguous_buffer, buffer_size, md);
    B2_EXPECTS(md.size() <= metadata_capacity, B2_ERR_INVALID_ARGUMENT, "metadata buffer too small");
    memcpy(metadata, md.data(), md.size());
    *metadata_size = md.size();
  B2_TRY_END
}

// out_columns: caller array of `capacity` views pointing into gpu_data; *num_columns / *num_rows from the header
b2_status b2_unpack(const uint8_t* metadata, size_t metadata_size, const void* gpu_data, b2_column_view* out_columns, int32_t capacity,
                    int32_t* num_columns, int32_t* num_rows)
{
  B2_TRY_BEGIN
    B2_EXPECTS(num_columns && num_rows, B2_ERR_INVALID_ARGUMENT, "null argument");
    std::vector<b2_column_view> cols;
    int32_t rows = 0;
    unpack_table(metadata, metadata_size, static_cast<const uint8_t*>(gpu_data), cols, rows);
    B2_EXPECTS((int32_t)cols.size() <= capacity && (cols.empty() || out_columns), B2_ERR_INVALID_ARGUMENT, "column array too small");
    for (size_t i = 0; i < cols.size(); ++i) out_columns[i] = cols[i];
    *num_columns = (int32_t)cols.size();
    *num_rows = rows;
  B2_TRY_END
}

}  // extern "C"
