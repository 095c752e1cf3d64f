// arrow_interop.cu — Arrow C Data / C Device Data interface for fixed-width columns.
//
// Replaces, for the types on this path, cpp/src/interop/to_arrow_schema.cpp:25-70 + arrow_utilities.cpp:65-101 (type mapping),
// to_arrow_device.cu:489-520 (ArrowDeviceArray over device buffers, CUDA event as sync_event), to_arrow_host.cu (deep copy to
// host) and from_arrow_device.cu / from_arrow_host.cu (the inverse).  The struct layouts are the published Arrow ABI
// (https://arrow.apache.org/docs/format/CDataInterface.html, CDeviceDataInterface.html); nanoarrow, which the reference uses
// to fill them, is not needed for flat fixed-width arrays.
//   buffers[0] = validity bitmap (LSB first, same as libcudf) or NULL, buffers[1] = values; `offset` applies to both.
//   BOOL8 is one byte per value here and one BIT per value in Arrow: converted in both directions (bools_to_mask in the
//   reference, to_arrow_device.cu:142).
#include "common.cuh"
#include "device_utils.cuh"

#include <cstdlib>
#include <cstring>

using namespace b2;

namespace {

const char* format_of(int32_t id)
{
  switch (id) {
    case B2_INT8: return "c";
    case B2_UINT8: return "C";
    case B2_INT16: return "s";
    case B2_UINT16: return "S";
    case B2_INT32: return "i";
    case B2_UINT32: return "I";
    case B2_INT64: return "l";
    case B2_UINT64: return "L";
    case B2_FLOAT32: return "f";
    case B2_FLOAT64: return "g";
    case B2_BOOL8: return "b";
    case B2_TIMESTAMP_DAYS: return "tdD";
    case B2_TIMESTAMP_SECONDS: return "tss:";
    case B2_TIMESTAMP_MILLISECONDS: return "tsm:";
    case B2_TIMESTAMP_MICROSECONDS: return "tsu:";
    case B2_TIMESTAMP_NANOSECONDS: return "tsn:";
    case B2_DURATION_SECONDS: return "tDs";
    case B2_DURATION_MILLISECONDS: return "tDm";
    case B2_DURATION_MICROSECONDS: return "tDu";
    case B2_DURATION_NANOSECONDS: return "tDn";
    default: return nullptr;  // DURATION_DAYS has no Arrow type (to_arrow_schema.cpp: data_type_error)
  }
}

int32_t type_of_format(const char* f)
{
  if (!f) return B2_EMPTY;
  for (int32_t id = B2_INT8; id <= B2_DURATION_NANOSECONDS; ++id) {
    const char* g = format_of(id);
    if (!g) continue;
    if (g[0] == 't' && g[1] == 's') {  // timestamps: any timezone suffix
      if (strncmp(f, g, 4) == 0) return id;
    } else if (strcmp(f, g) == 0) {
      return id;
    }
  }
  return B2_EMPTY;
}

__global__ void bytes_to_bits_kernel(const uint8_t* __restrict__ bytes, int64_t n, uint32_t* __restrict__ bits)
{
  // one warp per 32 values: ballot packs them into a word
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwords = (n + 31) / 32;
  if (w >= nwords) return;
  const int64_t i = w * 32 + lane_id();
  const unsigned b = __ballot_sync(0xffffffffu, i < n && bytes[i] != 0);
  if (lane_id() == 0) bits[w] = b;
}
__global__ void bits_to_bytes_kernel(const uint32_t* __restrict__ bits, int64_t bit_offset, int64_t n, uint8_t* __restrict__ bytes)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) bytes[i] = bit_is_set(bits, bit_offset + i) ? 1 : 0;
}

struct schema_private {
  std::string format, name;
};
void release_schema(b2_arrow_schema* s)
{
  if (!s || !s->release) return;
  delete static_cast<schema_private*>(s->private_data);
  s->release = nullptr;
}

struct array_private {
  const void* buffers[2] = {nullptr, nullptr};
  void* host_blocks[2]   = {nullptr, nullptr};  // malloc'ed (host export)
  dbuf device_block;                             // owned device buffer (bool conversion)
  cudaEvent_t event = nullptr;
};
void release_array(b2_arrow_array* a)
{
  if (!a || !a->release) return;
  auto* p = static_cast<array_private*>(a->private_data);
  if (p) {
    std::free(p->host_blocks[0]);
    std::free(p->host_blocks[1]);
    if (p->event) cudaEventDestroy(p->event);
    delete p;
  }
  a->release = nullptr;
}

void fill_array(b2_arrow_array* a, array_private* p, int64_t length, int64_t null_count, int64_t offset)
{
  memset(a, 0, sizeof(*a));
  a->length       = length;
  a->null_count   = null_count;
  a->offset       = offset;
  a->n_buffers    = 2;
  a->n_children   = 0;
  a->buffers      = p->buffers;
  a->release      = release_array;
  a->private_data = p;
}

}  // namespace

extern "C" {

// cudf::to_arrow_schema for one column (to_arrow_schema.cpp:25-70): flags = ARROW_FLAG_NULLABLE
b2_status b2_to_arrow_schema(const b2_column_view* col, const char* name, b2_arrow_schema* out)
{
  B2_TRY_BEGIN
    B2_EXPECTS(col && out, B2_ERR_INVALID_ARGUMENT, "null argument");
    const char* f = format_of(col->type_id);
    B2_EXPECTS(f != nullptr, B2_ERR_DATA_TYPE, "Unsupported type for to_arrow_schema");
    auto* p = new schema_private{f, name ? name : ""};
    memset(out, 0, sizeof(*out));
    out->format       = p->format.c_str();
    out->name         = p->name.c_str();
    out->flags        = 2;  // ARROW_FLAG_NULLABLE
    out->release      = release_schema;
    out->private_data = p;
  B2_TRY_END
}

// cudf::to_arrow_device(column_view) (to_arrow_device.cu:489-520): zero copy (BOOL8: bit-packed copy), sync_event recorded on
// `stream`; the caller keeps the column's memory alive until the array is released
b2_status b2_to_arrow_device(const b2_column_view* col, b2_stream stream, b2_arrow_device_array* out)
{
  B2_TRY_BEGIN
    B2_EXPECTS(col && out, B2_ERR_INVALID_ARGUMENT, "null argument");
    validate_column(*col);
    B2_EXPECTS(format_of(col->type_id) != nullptr, B2_ERR_DATA_TYPE, "Unsupported type for to_arrow_device");
    auto s = static_cast<cudaStream_t>(stream);
    auto p = std::make_unique<array_private>();
    int64_t offset = col->offset;
    p->buffers[0] = has_nulls(*col) ? col->null_mask : nullptr;
    p->buffers[1] = col->data;
    if (col->type_id == B2_BOOL8 && col->size > 0) {
      p->device_block = dbuf(bitmask_bytes(col->size), s);
      const int64_t nthreads = ((int64_t)col->size + 31) / 32 * 32;
      B2_LAUNCH(bytes_to_bits_kernel, (unsigned)((nthreads + 255) / 256), 256, 0, s, static_cast<const uint8_t*>(col->data) + col->offset, (int64_t)col->size,
                p->device_block.as<uint32_t>());
      p->buffers[1] = p->device_block.ptr;
      if (p->buffers[0] && offset) {  // the bit-packed values start at bit 0: re-base the validity too
        // (kept simple: a sliced nullable BOOL8 column is exported through its own mask copy)
        dbuf m = copy_bitmask(col->null_mask, col->offset, (int64_t)col->offset + col->size, s);
        // both buffers live in one private block list: append the mask behind the values
        dbuf both(p->device_block.bytes + m.bytes, s);
        B2_CUDA_TRY(cudaMemcpyAsync(both.ptr, p->device_block.ptr, p->device_block.bytes, cudaMemcpyDeviceToDevice, s));
        B2_CUDA_TRY(cudaMemcpyAsync(static_cast<char*>(both.ptr) + p->device_block.bytes, m.ptr, m.bytes, cudaMemcpyDeviceToDevice, s));
        p->buffers[1] = both.ptr;
        p->buffers[0] = static_cast<char*>(both.ptr) + p->device_block.bytes;
        p->device_block = std::move(both);
      }
      offset = 0;
    }
    B2_CUDA_TRY(cudaEventCreateWithFlags(&p->event, cudaEventDisableTiming));
    B2_CUDA_TRY(cudaEventRecord(p->event, s));
    memset(out, 0, sizeof(*out));
    fill_array(&out->array, p.get(), col->size, has_nulls(*col) ? col->null_count : 0, offset);
    int dev = 0;
    B2_CUDA_TRY(cudaGetDevice(&dev));
    out->device_id   = dev;
    out->device_type = 2;  // ARROW_DEVICE_CUDA
    out->sync_event  = &p->event;
    p.release();
  B2_TRY_END
}

// cudf::to_arrow_host(column_view): deep copy into host memory owned by the ArrowArray (offset re-based to 0)
b2_status b2_to_arrow_host(const b2_column_view* col, b2_stream stream, b2_arrow_array* out)
{
  B2_TRY_BEGIN
    B2_EXPECTS(col && out, B2_ERR_INVALID_ARGUMENT, "null argument");
    validate_column(*col);
    B2_EXPECTS(format_of(col->type_id) != nullptr, B2_ERR_DATA_TYPE, "Unsupported type for to_arrow_host");
    auto s = static_cast<cudaStream_t>(stream);
    auto p = std::make_unique<array_private>();
    const int64_t n = col->size;
    const size_t w = type_width(col->type_id);
    const bool is_bool = col->type_id == B2_BOOL8;
    if (n > 0) {
      const size_t vbytes = is_bool ? bitmask_bytes(n) : (size_t)n * w;
      p->host_blocks[1] = std::calloc(vbytes + 64, 1);
      if (is_bool) {
        dbuf bits(bitmask_bytes(n), s);
        const int64_t nthreads = (n + 31) / 32 * 32;
        B2_LAUNCH(bytes_to_bits_kernel, (unsigned)((nthreads + 255) / 256), 256, 0, s, static_cast<const uint8_t*>(col->data) + col->offset, n, bits.as<uint32_t>());
        B2_CUDA_TRY(cudaMemcpyAsync(p->host_blocks[1], bits.ptr, (size_t)((n + 31) / 32) * 4, cudaMemcpyDeviceToHost, s));
        B2_CUDA_TRY(cudaStreamSynchronize(s));
      } else {
        B2_CUDA_TRY(cudaMemcpyAsync(p->host_blocks[1], static_cast<const char*>(col->data) + (size_t)col->offset * w, (size_t)n * w, cudaMemcpyDeviceToHost, s));
      }
      if (has_nulls(*col)) {
        dbuf m = copy_bitmask(col->null_mask, col->offset, (int64_t)col->offset + n, s);
        p->host_blocks[0] = std::calloc(bitmask_bytes(n) + 64, 1);
        B2_CUDA_TRY(cudaMemcpyAsync(p->host_blocks[0], m.ptr, (size_t)((n + 31) / 32) * 4, cudaMemcpyDeviceToHost, s));
        B2_CUDA_TRY(cudaStreamSynchronize(s));
      }
      B2_CUDA_TRY(cudaStreamSynchronize(s));
    }
    p->buffers[0] = p->host_blocks[0];
    p->buffers[1] = p->host_blocks[1];
    fill_array(out, p.get(), n, has_nulls(*col) ? col->null_count : 0, 0);
    p.release();
  B2_TRY_END
}

// cudf::from_arrow_device_column: a view over the producer's device buffers (zero copy). BOOL8 needs a byte-per-value copy:
// then *out_owner receives an owning column and *out_view views it. The caller waits on array->sync_event if it is set.
b2_status b2_from_arrow_device(const b2_arrow_schema* schema, const b2_arrow_device_array* in, b2_stream stream, b2_column_view* out_view,
                               b2_column** out_owner)
{
  B2_TRY_BEGIN
    B2_EXPECTS(schema && in && out_view && out_owner, B2_ERR_INVALID_ARGUMENT, "null argument");
    B2_EXPECTS(in->device_type == 2 || in->device_type == 3 /* CUDA_HOST */ || in->device_type == 13 /* CUDA_MANAGED */, B2_ERR_INVALID_ARGUMENT,
               "ArrowDeviceArray memory not accessible by the current device");
    const int32_t id = type_of_format(schema->format);
    B2_EXPECTS(id != B2_EMPTY, B2_ERR_DATA_TYPE, "Unsupported Arrow format for this path (fixed-width types only)");
    const b2_arrow_array& a = in->array;
    B2_EXPECTS(a.n_buffers == 2 && a.n_children == 0 && a.dictionary == nullptr, B2_ERR_DATA_TYPE, "only flat fixed-width arrays are supported");
    B2_EXPECTS(a.length >= 0 && a.length <= INT32_MAX && a.offset >= 0 && a.offset <= INT32_MAX, B2_ERR_INVALID_ARGUMENT, "array too long for size_type");
    auto s = static_cast<cudaStream_t>(stream);
    if (in->sync_event) B2_CUDA_TRY(cudaStreamWaitEvent(s, *static_cast<cudaEvent_t*>(in->sync_event), 0));
    const uint32_t* mask = static_cast<const uint32_t*>(a.buffers[0]);
    int32_t nulls = 0;
    if (mask && a.length > 0) nulls = a.null_count >= 0 ? (int32_t)a.null_count : (int32_t)a.length - count_set_bits(mask, a.offset, a.offset + a.length, s);
    if (nulls == 0) mask = nullptr;
    *out_owner = nullptr;
    if (id == B2_BOOL8 && a.length > 0) {
      auto c = make_column(B2_BOOL8, (int32_t)a.length, false, s);
      const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((a.length + 255) / 256, NUM_SMS_B200 * 8));
      B2_LAUNCH(bits_to_bytes_kernel, grid, 256, 0, s, static_cast<const uint32_t*>(a.buffers[1]), a.offset, a.length, c->data.as<uint8_t>());
      if (mask) {
        c->mask = copy_bitmask(mask, a.offset, a.offset + a.length, s);
        c->null_count = nulls;
      }
      *out_view = b2_column_view{B2_BOOL8, (int32_t)a.length, c->data.ptr, static_cast<const uint32_t*>(c->mask.ptr), nulls, 0};
      *out_owner = c.release();
    } else {
      *out_view = b2_column_view{id, (int32_t)a.length, a.buffers[1], mask, nulls, (int32_t)a.offset};
    }
  B2_TRY_END
}

// cudf::from_arrow(schema, array) for host memory: copies to the device, returns an owning column
b2_status b2_from_arrow_host(const b2_arrow_schema* schema, const b2_arrow_array* a, b2_stream stream, b2_column** out)
{
  B2_TRY_BEGIN
    B2_EXPECTS(schema && a && out, B2_ERR_INVALID_ARGUMENT, "null argument");
    const int32_t id = type_of_format(schema->format);
    B2_EXPECTS(id != B2_EMPTY, B2_ERR_DATA_TYPE, "Unsupported Arrow format for this path (fixed-width types only)");
    B2_EXPECTS(a->n_buffers == 2 && a->n_children == 0 && a->dictionary == nullptr, B2_ERR_DATA_TYPE, "only flat fixed-width arrays are supported");
    B2_EXPECTS(a->length >= 0 && a->length <= INT32_MAX && a->offset >= 0, B2_ERR_INVALID_ARGUMENT, "array too long for size_type");
    auto s = static_cast<cudaStream_t>(stream);
    const int64_t n = a->length, off = a->offset;
    const size_t w = type_width(id);
    const uint8_t* hmask = static_cast<const uint8_t*>(a->buffers[0]);
    auto c = make_column(id, (int32_t)n, false, s);
    if (n > 0) {
      // stage the host bitmaps on the device with their bit offset, then re-base with the library's kernels
      auto upload_bits = [&](const uint8_t* bits, int64_t first_bit, int64_t nbits, dbuf& staged) {
        const int64_t w0 = first_bit / 32, w1 = (first_bit + nbits + 31) / 32;
        staged = dbuf((size_t)(w1 - w0) * 4 + 64, s);
        B2_CUDA_TRY(cudaMemsetAsync(staged.ptr, 0, staged.bytes, s));
        const int64_t byte0 = w0 * 4, byte1 = std::min<int64_t>((first_bit + nbits + 7) / 8, w1 * 4);
        B2_CUDA_TRY(cudaMemcpyAsync(staged.ptr, bits + byte0, (size_t)(byte1 - byte0), cudaMemcpyHostToDevice, s));
        return first_bit - w0 * 32;  // bit offset inside the staged buffer
      };
      if (id == B2_BOOL8) {
        dbuf staged;
        const int64_t bo = upload_bits(static_cast<const uint8_t*>(a->buffers[1]), off, n, staged);
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, NUM_SMS_B200 * 8));
        B2_LAUNCH(bits_to_bytes_kernel, grid, 256, 0, s, staged.as<uint32_t>(), bo, n, c->data.as<uint8_t>());
        B2_CUDA_TRY(cudaStreamSynchronize(s));
      } else {
        B2_CUDA_TRY(cudaMemcpyAsync(c->data.ptr, static_cast<const char*>(a->buffers[1]) + (size_t)off * w, (size_t)n * w, cudaMemcpyHostToDevice, s));
      }
      if (hmask && a->null_count != 0) {
        dbuf staged;
        const int64_t bo = upload_bits(hmask, off, n, staged);
        c->mask = copy_bitmask(staged.as<uint32_t>(), bo, bo + n, s);
        const int32_t valid = count_set_bits(c->mask.as<uint32_t>(), 0, n, s);
        c->null_count = (int32_t)n - valid;
        if (c->null_count == 0) c->mask.reset();
      }
      B2_CUDA_TRY(cudaStreamSynchronize(s));  // the host buffers may go away once we return
    }
    *out = c.release();
  B2_TRY_END
}

void b2_arrow_schema_release(b2_arrow_schema* s) { if (s && s->release) s->release(s); }
void b2_arrow_array_release(b2_arrow_array* a) { if (a && a->release) a->release(a); }

}  // extern "C"
