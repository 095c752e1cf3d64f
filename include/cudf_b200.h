/*
 * cudf_b200.h — flat C ABI of the B200-native hot path (sort / hash join / hash groupby /
 * scan / reduce / segmented reduce / gather / null-mask utilities).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ types, no torch types.
 * The reference (rapidsai/cudf) has no C ABI of its own; its boundary is the libcudf C++ API that
 * pylibcudf's .pxd files bind.  Every entry point below names the reference interface it replaces
 * (paths relative to the reference tree).  `include/cudf/ *.hpp` re-creates that C++ surface as
 * header-only wrappers over these functions; `cudf_b200/pylibcudf` is the Python twin.
 *
 * Conventions
 *  - All pointers inside views are DEVICE pointers (Arrow layout, fixed-width types only).
 *  - `b2_column_view` mirrors cudf::column_view (cpp/include/cudf/column/column_view.hpp:237-244):
 *    element i lives at data[(offset+i)], validity bit at bit (offset+i) of null_mask (LSB first,
 *    32-bit words, 1 = valid); null_mask may be NULL (all valid); null_count must be exact.
 *  - Every call is ordered on `stream` (a cudaStream_t passed as void*) and may return before the
 *    device work finishes, except where a size has to be read back (join size, groupby growth).
 *  - Outputs are library-owned handles released with the matching *_free function.
 *  - Return value: b2_status; on failure b2_last_error() holds a thread-local message.  The status
 *    maps 1:1 on the reference exception taxonomy (cpp/include/cudf/utilities/error.hpp:35-118):
 *      LOGIC -> cudf::logic_error, INVALID_ARGUMENT -> std::invalid_argument,
 *      DATA_TYPE -> cudf::data_type_error, OUT_OF_RANGE -> std::out_of_range,
 *      BAD_ALLOC -> std::bad_alloc, CUDA -> cudf::cuda_error.
 */
#ifndef CUDF_B200_H
#define CUDF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2_API __attribute__((visibility("default")))

typedef void* b2_stream; /* cudaStream_t */

typedef enum b2_status {
  B2_OK                   = 0,
  B2_ERR_LOGIC            = 1,
  B2_ERR_INVALID_ARGUMENT = 2,
  B2_ERR_DATA_TYPE        = 3,
  B2_ERR_OUT_OF_RANGE     = 4,
  B2_ERR_BAD_ALLOC        = 5,
  B2_ERR_CUDA             = 6
} b2_status;

/* cudf::type_id values (cpp/include/cudf/types.hpp:183-216). Only fixed-width ids are accepted. */
enum {
  B2_EMPTY = 0, B2_INT8 = 1, B2_INT16 = 2, B2_INT32 = 3, B2_INT64 = 4,
  B2_UINT8 = 5, B2_UINT16 = 6, B2_UINT32 = 7, B2_UINT64 = 8,
  B2_FLOAT32 = 9, B2_FLOAT64 = 10, B2_BOOL8 = 11,
  B2_TIMESTAMP_DAYS = 12, B2_TIMESTAMP_SECONDS = 13, B2_TIMESTAMP_MILLISECONDS = 14,
  B2_TIMESTAMP_MICROSECONDS = 15, B2_TIMESTAMP_NANOSECONDS = 16,
  B2_DURATION_DAYS = 17, B2_DURATION_SECONDS = 18, B2_DURATION_MILLISECONDS = 19,
  B2_DURATION_MICROSECONDS = 20, B2_DURATION_NANOSECONDS = 21
};

/* cudf::order / null_order / null_policy / null_equality (types.hpp:99-150): bool enums. */
enum { B2_ASCENDING = 0, B2_DESCENDING = 1 };
enum { B2_NULL_AFTER = 0, B2_NULL_BEFORE = 1 };
enum { B2_NULL_EXCLUDE = 0, B2_NULL_INCLUDE = 1 };
enum { B2_NULLS_EQUAL = 0, B2_NULLS_UNEQUAL = 1 };
/* cudf::out_of_bounds_policy (cpp/include/cudf/copying.hpp:37-40) */
enum { B2_OOB_NULLIFY = 0, B2_OOB_DONT_CHECK = 1 };
/* cudf::mask_state (types.hpp:172-177) */
enum { B2_MASK_UNALLOCATED = 0, B2_MASK_UNINITIALIZED = 1, B2_MASK_ALL_VALID = 2, B2_MASK_ALL_NULL = 3 };
/* cudf::aggregation::Kind subset (cpp/include/cudf/aggregation.hpp:78-121), same numeric values. */
enum {
  B2_AGG_SUM = 0, B2_AGG_PRODUCT = 2, B2_AGG_MIN = 3, B2_AGG_MAX = 4,
  B2_AGG_COUNT_VALID = 5, B2_AGG_COUNT_ALL = 6, B2_AGG_SUM_OF_SQUARES = 9, B2_AGG_MEAN = 10, B2_AGG_M2 = 11,
  B2_AGG_VARIANCE = 12, B2_AGG_STD = 13, /* groupby only; ddof = 1 unless given with B2_AGG_WITH_DDOF */
  B2_AGG_ARGMAX = 16, B2_AGG_ARGMIN = 17 /* groupby only: INT32 row index of the extreme value (first row among ties) */
};
/* make_variance_aggregation(ddof) / make_std_aggregation(ddof) (aggregation.hpp:231-260): the kind word carries ddof */
#define B2_AGG_WITH_DDOF(kind, ddof) ((int32_t)(kind) | (1 << 30) | (((int32_t)(ddof) & 0xFFFF) << 8))
/* cudf::scan_type (cpp/include/cudf/reduction.hpp) */
enum { B2_SCAN_INCLUSIVE = 0, B2_SCAN_EXCLUSIVE = 1 };
/* JoinNoMatch sentinel (cpp/include/cudf/join/join.hpp:72) */
#define B2_JOIN_NO_MATCH INT32_MIN

typedef struct b2_column_view {
  int32_t         type_id;
  int32_t         size;
  const void*     data;
  const uint32_t* null_mask;
  int32_t         null_count;
  int32_t         offset;
} b2_column_view;

typedef struct b2_table_view {
  const b2_column_view* columns;
  int32_t               num_columns;
} b2_table_view;

typedef struct b2_column    b2_column;    /* owning cudf::column     (column.hpp:36-334)  */
typedef struct b2_table     b2_table;     /* owning cudf::table      (table.hpp:31-215)   */
typedef struct b2_scalar    b2_scalar;    /* owning numeric_scalar<T> (scalar/scalar.hpp) */
typedef struct b2_buffer    b2_buffer;    /* owning rmm::device_buffer                    */
typedef struct b2_hash_join b2_hash_join; /* cudf::hash_join         (join/hash_join.hpp) */
typedef struct b2_groupby   b2_groupby;   /* cudf::groupby::groupby  (groupby.hpp)        */

/* cudf::groupby::aggregation_request (cpp/include/cudf/groupby.hpp:60-64) */
typedef struct b2_agg_request {
  b2_column_view values;
  const int32_t* kinds;     /* B2_AGG_* */
  int32_t        num_kinds;
} b2_agg_request;

/* ---- errors / runtime ------------------------------------------------------------------- */
B2_API const char* b2_last_error(void);
B2_API const char* b2_version(void);
/* Number of kernels this library has launched in this process (bench.py gpu_launches). */
B2_API uint64_t b2_kernel_launch_count(void);
/* Optional per-kernel-family timing with CUDA events on the caller's stream (off by default; the
 * NVTX-range analogue of CUDF_FUNC_RANGE, used by bench.py for the live roofline figure).
 * b2_profile_get synchronises the device and returns accumulated milliseconds and launch count. */
B2_API void      b2_profile_enable(int32_t on);
B2_API void      b2_profile_reset(void);
B2_API b2_status b2_profile_get(const char* name, double* total_ms, int64_t* launches);
/* same, restricted to the scopes that took at least `min_ms` (separates executed radix passes from skipped ones) */
B2_API b2_status b2_profile_get_over(const char* name, double min_ms, double* total_ms, int64_t* launches);
/* Trim the stream-ordered pool back to the driver (rmm pool release analogue). */
B2_API b2_status b2_trim_pool(void);

/* ---- owning handles --------------------------------------------------------------------- */
B2_API b2_status b2_column_view_of(const b2_column* col, b2_column_view* out);
B2_API void      b2_column_free(b2_column* col);
B2_API int32_t   b2_table_num_columns(const b2_table* tbl);
B2_API int32_t   b2_table_num_rows(const b2_table* tbl);
/* borrowed pointer, valid while the table lives */
B2_API const b2_column* b2_table_column(const b2_table* tbl, int32_t i);
/* cudf::table::release(): moves the columns out (caller frees each), table becomes empty */
B2_API b2_status b2_table_release(b2_table* tbl, b2_column** out_cols, int32_t capacity);
B2_API void      b2_table_free(b2_table* tbl);
B2_API void*     b2_buffer_data(const b2_buffer* buf);
B2_API size_t    b2_buffer_size(const b2_buffer* buf);
B2_API void      b2_buffer_free(b2_buffer* buf);
/* numeric_scalar<T>: value bytes are the native representation of type_id (<= 8 bytes). */
B2_API b2_status b2_scalar_create(int32_t type_id, const void* host_value, int32_t is_valid,
                                  b2_stream stream, b2_scalar** out);
B2_API int32_t     b2_scalar_type(const b2_scalar* s);
B2_API const void* b2_scalar_device_data(const b2_scalar* s);
/* synchronises `stream`; copies the value (<= 8 bytes) and validity to the host */
B2_API b2_status b2_scalar_get(const b2_scalar* s, b2_stream stream, void* host_value, int32_t* is_valid);
B2_API void      b2_scalar_free(b2_scalar* s);

/* ---- null masks: cpp/include/cudf/null_mask.hpp, cpp/src/bitmask/null_mask.cu ------------ */
B2_API size_t    b2_bitmask_allocation_size_bytes(int32_t number_of_bits);             /* null_mask.hpp:55 */
B2_API b2_status b2_create_null_mask(int32_t size, int32_t mask_state, b2_stream stream,
                                     b2_buffer** out);                                  /* null_mask.cu:48-86 */
B2_API b2_status b2_set_null_mask(uint32_t* bitmask, int32_t begin_bit, int32_t end_bit, int32_t valid,
                                  b2_stream stream);                                    /* null_mask.cu:339-404 */
B2_API b2_status b2_copy_bitmask(const uint32_t* mask, int32_t begin_bit, int32_t end_bit, b2_stream stream,
                                 b2_buffer** out);                                      /* null_mask.cu:409-560 */
B2_API b2_status b2_count_set_bits(const uint32_t* bitmask, int32_t start, int32_t stop, b2_stream stream,
                                   int32_t* out);                                       /* cudf::detail::count_set_bits */
B2_API b2_status b2_null_count(const uint32_t* bitmask, int32_t start, int32_t stop, b2_stream stream,
                                   int32_t* out);                                       /* cudf::null_count */
/* AND of the masks of all columns -> (mask, null_count); mask NULL when no column is nullable */
B2_API b2_status b2_bitmask_and(const b2_table_view* view, b2_stream stream, b2_buffer** out_mask,
                                int32_t* out_null_count);                               /* null_mask.cu:608-735 */

/* ---- gather: cpp/include/cudf/copying.hpp:81-86, cpp/include/cudf/detail/gather.cuh:627-675 */
B2_API b2_status b2_gather(const b2_table_view* source, const b2_column_view* gather_map, int32_t oob_policy,
                           b2_stream stream, b2_table** out);

/* ---- sort: cpp/include/cudf/sorting.hpp:44-163, cpp/src/sort/{sort,stable_sort}.cu -------- */
/* column_order / null_precedence: arrays of B2_ASCENDING.. / B2_NULL_AFTER.. of length n_order /
 * n_null_prec; 0 length = defaults (ASCENDING, BEFORE: sort_impl.cuh:56-57). */
B2_API b2_status b2_sorted_order(const b2_table_view* keys, const uint8_t* column_order, int32_t n_order,
                                 const uint8_t* null_precedence, int32_t n_null_prec, int32_t stable,
                                 b2_stream stream, b2_column** out);
B2_API b2_status b2_sort(const b2_table_view* input, const uint8_t* column_order, int32_t n_order,
                         const uint8_t* null_precedence, int32_t n_null_prec, int32_t stable,
                         b2_stream stream, b2_table** out);
B2_API b2_status b2_sort_by_key(const b2_table_view* values, const b2_table_view* keys,
                                const uint8_t* column_order, int32_t n_order,
                                const uint8_t* null_precedence, int32_t n_null_prec, int32_t stable,
                                b2_stream stream, b2_table** out);

/* cudf::{stable_,}segmented_sorted_order / segmented_sort_by_key (sorting.hpp:232-366): segment_offsets is an INT32
 * column of start offsets (the last entry ends the last segment); rows outside [offsets[0], offsets[last]) keep their
 * place; fewer than two offsets sort nothing; a non-INT32 offsets column -> LOGIC error. */
B2_API b2_status b2_segmented_sorted_order(const b2_table_view* keys, const b2_column_view* segment_offsets,
                                           const uint8_t* column_order, int32_t n_order,
                                           const uint8_t* null_precedence, int32_t n_null_prec, int32_t stable,
                                           b2_stream stream, b2_column** out);
B2_API b2_status b2_segmented_sort_by_key(const b2_table_view* values, const b2_table_view* keys,
                                          const b2_column_view* segment_offsets, const uint8_t* column_order,
                                          int32_t n_order, const uint8_t* null_precedence, int32_t n_null_prec,
                                          int32_t stable, b2_stream stream, b2_table** out);
/* cudf::top_k / top_k_order (sorting.hpp:370-416, top_k.cu:100-170): the k first rows of the stable sorted order
 * (nulls last for ASCENDING, first for DESCENDING); k >= size returns all rows; k < 0 -> INVALID_ARGUMENT. */
B2_API b2_status b2_top_k(const b2_column_view* col, int32_t k, int32_t topk_order, b2_stream stream, b2_column** out);
B2_API b2_status b2_top_k_order(const b2_column_view* col, int32_t k, int32_t topk_order, b2_stream stream,
                                b2_column** out);
/* cudf::rank (sorting.hpp:165-230, cpp/src/sort/rank.cu:236-356). method = cudf::rank_method (0 FIRST, 1 AVERAGE,
 * 2 MIN, 3 MAX, 4 DENSE); the result is FLOAT64 for AVERAGE or percentage, INT32 otherwise; under
 * null_handling = EXCLUDE the result carries the input's validity; percentage divides by the row count (DENSE: by
 * the number of distinct values). */
B2_API b2_status b2_rank(const b2_column_view* input, int32_t method, int32_t column_order, int32_t null_handling,
                         int32_t null_precedence, int32_t percentage, b2_stream stream, b2_column** out);

/* ---- hash join: cpp/include/cudf/join/join.hpp:127-249, join/hash_join.hpp ---------------- */
/* Results are two INT32 columns of equal length (cudf returns device_uvector<size_type>), row
 * order unspecified (join.hpp:130-136). */
B2_API b2_status b2_inner_join(const b2_table_view* left_keys, const b2_table_view* right_keys,
                               int32_t compare_nulls, b2_stream stream, b2_column** out_left,
                               b2_column** out_right);
B2_API b2_status b2_left_join(const b2_table_view* left_keys, const b2_table_view* right_keys,
                              int32_t compare_nulls, b2_stream stream, b2_column** out_left,
                              b2_column** out_right);
B2_API b2_status b2_full_join(const b2_table_view* left_keys, const b2_table_view* right_keys,
                              int32_t compare_nulls, b2_stream stream, b2_column** out_left,
                              b2_column** out_right);
/* hash_join(build, has_nulls, compare_nulls, load_factor, stream): hash_join.hpp ctor #2.
 * has_nulls < 0 = derive from the build table (ctor #1). load_factor outside (0,1] ->
 * INVALID_ARGUMENT (cpp/tests/join/join_tests.cpp:346-366). */
B2_API b2_status b2_hash_join_create(const b2_table_view* build, int32_t has_nulls, int32_t compare_nulls,
                                     double load_factor, b2_stream stream, b2_hash_join** out);
B2_API void      b2_hash_join_destroy(b2_hash_join* hj);
/* has_output_size = 0 -> the size is computed (std::optional<size_t> output_size = nullopt) */
B2_API b2_status b2_hash_join_inner_join(const b2_hash_join* hj, const b2_table_view* probe,
                                         int32_t has_output_size, size_t output_size, b2_stream stream,
                                         b2_column** out_left, b2_column** out_right);
B2_API b2_status b2_hash_join_left_join(const b2_hash_join* hj, const b2_table_view* probe,
                                        int32_t has_output_size, size_t output_size, b2_stream stream,
                                        b2_column** out_left, b2_column** out_right);
B2_API b2_status b2_hash_join_full_join(const b2_hash_join* hj, const b2_table_view* probe,
                                        int32_t has_output_size, size_t output_size, b2_stream stream,
                                        b2_column** out_left, b2_column** out_right);
B2_API b2_status b2_hash_join_inner_join_size(const b2_hash_join* hj, const b2_table_view* probe,
                                              b2_stream stream, size_t* out);
B2_API b2_status b2_hash_join_left_join_size(const b2_hash_join* hj, const b2_table_view* probe,
                                             b2_stream stream, size_t* out);
B2_API b2_status b2_hash_join_full_join_size(const b2_hash_join* hj, const b2_table_view* probe,
                                             b2_stream stream, size_t* out);

/* hash_join::{inner,left,full}_join_match_context (hash_join.hpp:254-330, join.hpp:81-125): per probe row, the
 * number of matching build rows as an INT32 column of probe.num_rows (join_kind: 0 inner, 1 left, 2 full; for
 * left / full a row without a match counts 1 — its null-placeholder output row). Golden: join_tests.cpp:2339-2527. */
B2_API b2_status b2_hash_join_match_counts(const b2_hash_join* hj, const b2_table_view* probe, int32_t join_kind,
                                           b2_stream stream, b2_column** out_counts);
/* hash_join::partitioned_{inner,left,full}_join (hash_join.hpp:331-411): join rows [left_start, left_end) of the
 * probe table given the match counts of the WHOLE probe table (from b2_hash_join_match_counts with the same
 * join_kind). Left indices are relative to the whole probe table. join_kind 2 does not append the unmatched build
 * rows: b2_hash_join_finalize_full_join does. Bounds outside [0, num_rows] -> INVALID_ARGUMENT. */
B2_API b2_status b2_hash_join_partitioned_join(const b2_hash_join* hj, const b2_table_view* probe,
                                               const b2_column_view* match_counts, int32_t left_start,
                                               int32_t left_end, int32_t join_kind, b2_stream stream,
                                               b2_column** out_left, b2_column** out_right);
/* hash_join::finalize_partitioned_full_join (hash_join.hpp:413-440): concatenates the per-partition index pairs and
 * appends (JoinNoMatch, r) for every build row r that no partition matched. */
B2_API b2_status b2_hash_join_finalize_full_join(const b2_column_view* left_partials,
                                                 const b2_column_view* right_partials, int32_t num_partials,
                                                 int32_t left_table_num_rows, int32_t right_table_num_rows,
                                                 b2_stream stream, b2_column** out_left, b2_column** out_right);

/* ---- groupby: cpp/include/cudf/groupby.hpp:121-125,181-184, cpp/src/groupby/groupby.cu ----- */
B2_API b2_status b2_groupby_create(const b2_table_view* keys, int32_t null_handling, int32_t keys_are_sorted,
                                   const uint8_t* column_order, int32_t n_order,
                                   const uint8_t* null_precedence, int32_t n_null_prec, b2_groupby** out);
B2_API void      b2_groupby_destroy(b2_groupby* gb);
/* aggregate(): out_keys = distinct key rows; out_results = one column per (request, kind) in
 * request-major order (aggregation_result::results flattened). Row order arbitrary (groupby.hpp:148). */
B2_API b2_status b2_groupby_aggregate(b2_groupby* gb, const b2_agg_request* requests, int32_t num_requests,
                                      b2_stream stream, b2_table** out_keys, b2_table** out_results);
/* scan(): rows in sorted key order (cpp/src/groupby/sort/scan.cpp) */
B2_API b2_status b2_groupby_scan(b2_groupby* gb, const b2_agg_request* requests, int32_t num_requests,
                                 b2_stream stream, b2_table** out_keys, b2_table** out_results);

/* ---- reduce / scan / segmented reduce: cpp/include/cudf/reduction.hpp ---------------------- */
/* init may be NULL (std::nullopt). */
B2_API b2_status b2_reduce(const b2_column_view* col, int32_t agg_kind, int32_t output_type_id,
                           const b2_scalar* init, b2_stream stream, b2_scalar** out);
B2_API b2_status b2_segmented_reduce(const b2_column_view* values, const int32_t* offsets, int32_t num_offsets,
                                     int32_t agg_kind, int32_t output_type_id, int32_t null_handling,
                                     const b2_scalar* init, b2_stream stream, b2_column** out);
B2_API b2_status b2_scan(const b2_column_view* col, int32_t agg_kind, int32_t scan_type, int32_t null_handling,
                         b2_stream stream, b2_column** out);

/* ---- sharded path helpers (no libcudf equivalent on one GPU; role of cudf::hash_partition,
 *      cpp/include/cudf/partitioning.hpp:103-145, and of cudf_polars' sort splitters) ----------- */
/* Stable partition of `input` rows into num_partitions buckets. out_offsets (host int32[P+1]).
 * mode 0: bucket = number of splitters <= key (range partition on the single key column `keys`,
 *         splitters = device array of P-1 ascending keys of the same type)
 * mode 1: bucket = mix64(key bits) % P (hash partition)                                          */
B2_API b2_status b2_partition(const b2_table_view* input, const b2_column_view* keys, int32_t mode,
                              const void* splitters, int32_t num_partitions, b2_stream stream,
                              b2_table** out, int32_t* out_offsets);

/* cudf::hash_partition(input, keys, num_partitions, hash_function, seed) — cpp/include/cudf/partitioning.hpp:103-145,
 * cpp/src/partitioning/partitioning.cu:875-945.  Row hash as in libcudf (hash_function 1 = HASH_MURMUR3: MurmurHash3_x86_32
 * per column with `seed`, floats normalised, null = UINT32_MAX, columns folded with hash_combine; 0 = HASH_IDENTITY on
 * integral keys), partition = hash % num_partitions: a row lands in the same partition as under libcudf, so the output
 * interoperates with dask_cudf / rapidsmpf style shuffles.  out_offsets: host int32[num_partitions + 1].  Rows keep their
 * input order inside a partition.  Up to 256 partitions take the tile-based partition kernels, more go through a stable
 * radix order of the 32-bit partition ids; zero key columns / rows give an empty result. */
B2_API b2_status b2_hash_partition(const b2_table_view* input, const b2_table_view* keys, int32_t num_partitions,
                                   int32_t hash_function, uint32_t seed, b2_stream stream, b2_table** out, int32_t* out_offsets);

/* cudf::partition(t, partition_map, num_partitions) for any partition count (cpp/include/cudf/partitioning.hpp:58-101): the
 * integer map names each row's partition; stable order of the map values + fused gather. out_offsets: int32[P + 1]. */
B2_API b2_status b2_partition_by_map(const b2_table_view* input, const b2_column_view* partition_map, int32_t num_partitions,
                                     b2_stream stream, b2_table** out, int32_t* out_offsets);

/* Two-phase form of b2_partition for the fused partition + exchange: the plan holds the bucket id and the
 * stable in-bucket rank of every row; out_counts[b] = rows of bucket b.  b2_partition_scatter then writes one
 * fixed-width column straight to P destination base addresses — local buffers or PEER device memory mapped with
 * b2_ipc_open — so that the all-to-all exchange happens inside the scatter kernel over NVLink
 * (dest_ptrs[b] = address of bucket b's first row; host array of P device pointers). */
typedef struct b2_partition_plan b2_partition_plan;
B2_API b2_status b2_partition_plan_create(const b2_column_view* keys, int32_t mode, const void* splitters,
                                          int32_t num_partitions, b2_stream stream, b2_partition_plan** out,
                                          int64_t* out_counts);
B2_API b2_status b2_partition_scatter(const b2_partition_plan* plan, const b2_column_view* column,
                                      void* const* dest_ptrs, b2_stream stream);
/* EXPERIMENTAL variant of b2_partition_scatter: stages per-destination runs of a 4096-row tile in shared memory
 * before the (remote) stores; compiled in round 1 but not yet validated on hardware and not used by default. */
B2_API b2_status b2_partition_scatter_staged(const b2_partition_plan* plan, const b2_column_view* column,
                                             void* const* dest_ptrs, b2_stream stream);
B2_API void      b2_partition_plan_free(b2_partition_plan* plan);
/* Range partition fused into ONE one-sweep pass whose digit is the bucket (number of splitters <= key): first the bucket counts of
 * this rank's rows, then — after the ranks exchanged their counts — a stable pass that writes bucket b's keys (and optionally one
 * 4- / 8-byte payload column) to key_dst[b] / val_dst[b], local or peer memory (host arrays of num_partitions device pointers: the
 * address of THIS rank's first row of bucket b). One null-free 8-byte integer-like key column; splitters = device array of P - 1
 * ascending keys of the column's type. The sharded sort's partition + exchange ("fuse with the first radix pass").
 * splitters == NULL selects a HASH partition instead (bucket = high-multiply of a 64-bit mix of the key by num_partitions; the
 * same function in both calls and on every rank): the sharded join's shuffle. */
B2_API b2_status b2_range_partition_counts(const b2_column_view* keys, const void* splitters, int32_t num_partitions, b2_stream stream,
                                           int64_t* out_counts);
B2_API b2_status b2_range_partition_scatter(const b2_column_view* keys, const b2_column_view* values, const void* splitters,
                                            int32_t num_partitions, void* const* key_dst, void* const* val_dst, b2_stream stream);
/* CUDA-IPC exchange buffers (cudaMalloc + cudaIpcGetMemHandle / cudaIpcOpenMemHandle); handle = 64 bytes */
B2_API b2_status b2_ipc_alloc(size_t bytes, void** out_ptr, uint8_t* out_handle64);
B2_API b2_status b2_ipc_open(const uint8_t* handle64, void** out_ptr);
B2_API b2_status b2_ipc_close(void* ptr);
B2_API b2_status b2_ipc_free(void* ptr);
/* Copy `bytes` (any alignment) from local device memory to `dst`, which may be PEER memory mapped with b2_ipc_open: a
 * plain copy kernel whose stores travel over NVLink (the bucket exchange after b2_partition: one contiguous run per
 * destination rank). Stream-ordered. */
B2_API b2_status b2_peer_copy(void* dst, const void* src, size_t bytes, b2_stream stream);

/* ---- cudf::pack / packed_size / pack_metadata / unpack (cpp/include/cudf/contiguous_split.hpp:233-317, cpp/src/copying/pack.cpp):
 *      libcudf's contiguous wire format for tables of fixed-width columns. metadata = host bytes (16-byte table header +
 *      40 bytes per column), gpu_data = one device buffer (validity then data per column, 64-byte padded). -------------- */
B2_API b2_status b2_packed_size(const b2_table_view* input, size_t* out_bytes);
B2_API b2_status b2_pack(const b2_table_view* input, b2_stream stream, uint8_t* metadata, size_t metadata_capacity,
                         size_t* metadata_size, b2_buffer** gpu_data);
B2_API b2_status b2_pack_metadata(const b2_table_view* input, const uint8_t* contiguous_buffer, size_t buffer_size,
                                  uint8_t* metadata, size_t metadata_capacity, size_t* metadata_size);
/* No allocation: out_columns[i] point into gpu_data (the caller keeps it alive). */
B2_API b2_status b2_unpack(const uint8_t* metadata, size_t metadata_size, const void* gpu_data, b2_column_view* out_columns,
                           int32_t capacity, int32_t* num_columns, int32_t* num_rows);

/* ---- Arrow C Data / C Device Data interface (cpp/include/cudf/interop.hpp to_arrow_schema / to_arrow_device / to_arrow_host /
 *      from_arrow_device_column / from_arrow; cpp/src/interop/*).  The structs are the published Arrow ABI: a pointer to
 *      b2_arrow_schema / b2_arrow_array / b2_arrow_device_array IS a pointer to ArrowSchema / ArrowArray / ArrowDeviceArray.
 *      Fixed-width types only (BOOL8 is bit-packed on the Arrow side and converted). --------------------------------------- */
typedef struct b2_arrow_schema {
  const char* format; const char* name; const char* metadata; int64_t flags; int64_t n_children;
  struct b2_arrow_schema** children; struct b2_arrow_schema* dictionary;
  void (*release)(struct b2_arrow_schema*); void* private_data;
} b2_arrow_schema;
typedef struct b2_arrow_array {
  int64_t length; int64_t null_count; int64_t offset; int64_t n_buffers; int64_t n_children;
  const void** buffers; struct b2_arrow_array** children; struct b2_arrow_array* dictionary;
  void (*release)(struct b2_arrow_array*); void* private_data;
} b2_arrow_array;
typedef struct b2_arrow_device_array {
  b2_arrow_array array; int64_t device_id; int32_t device_type; void* sync_event; int64_t reserved[3];
} b2_arrow_device_array;
B2_API b2_status b2_to_arrow_schema(const b2_column_view* col, const char* name, b2_arrow_schema* out);
/* zero copy: the caller keeps the column alive until out->array.release is called; sync_event is recorded on `stream` */
B2_API b2_status b2_to_arrow_device(const b2_column_view* col, b2_stream stream, b2_arrow_device_array* out);
B2_API b2_status b2_to_arrow_host(const b2_column_view* col, b2_stream stream, b2_arrow_array* out);
/* out_view points into the producer's buffers (which must outlive it); BOOL8 input is converted: *out_owner owns the copy */
B2_API b2_status b2_from_arrow_device(const b2_arrow_schema* schema, const b2_arrow_device_array* in, b2_stream stream,
                                      b2_column_view* out_view, b2_column** out_owner);
B2_API b2_status b2_from_arrow_host(const b2_arrow_schema* schema, const b2_arrow_array* in, b2_stream stream, b2_column** out);
B2_API void      b2_arrow_schema_release(b2_arrow_schema* schema);
B2_API void      b2_arrow_array_release(b2_arrow_array* array);

/* ---- synthetic data (SURVEY §8d generator): x_i = splitmix64(seed + first + i) -------------- */
/* kind 0: raw uint64 -> int64 ; 1: float64 uniform [0,1) ; 2: x mod modulus as int64 ;
 * 3: int32 low bits ; 4: validity bitmask words with P(valid)=0.5 (n = number of bits) */
B2_API b2_status b2_fill_splitmix64(void* dst, int64_t n, uint64_t seed, int64_t first, int32_t kind,
                                    uint64_t modulus, b2_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* CUDF_B200_H */
