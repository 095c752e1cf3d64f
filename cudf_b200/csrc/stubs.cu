// stubs.cu — entry points declared in include/cudf_b200.h whose kernels are not written yet.
// They fail loudly (B2_ERR_LOGIC) instead of silently falling back to anything.
#include "common.cuh"
#define NOT_YET(name) do { b2::set_last_error(name ": not implemented yet"); return B2_ERR_LOGIC; } while (0)
extern "C" {
b2_status b2_inner_join(const b2_table_view*, const b2_table_view*, int32_t, b2_stream, b2_column**, b2_column**) { NOT_YET("b2_inner_join"); }
b2_status b2_left_join(const b2_table_view*, const b2_table_view*, int32_t, b2_stream, b2_column**, b2_column**) { NOT_YET("b2_left_join"); }
b2_status b2_full_join(const b2_table_view*, const b2_table_view*, int32_t, b2_stream, b2_column**, b2_column**) { NOT_YET("b2_full_join"); }
b2_status b2_hash_join_create(const b2_table_view*, int32_t, int32_t, double, b2_stream, b2_hash_join**) { NOT_YET("b2_hash_join_create"); }
void b2_hash_join_destroy(b2_hash_join*) {}
b2_status b2_hash_join_inner_join(const b2_hash_join*, const b2_table_view*, int32_t, size_t, b2_stream, b2_column**, b2_column**) { NOT_YET("b2_hash_join_inner_join"); }
b2_status b2_hash_join_left_join(const b2_hash_join*, const b2_table_view*, int32_t, size_t, b2_stream, b2_column**, b2_column**) { NOT_YET("b2_hash_join_left_join"); }
b2_status b2_hash_join_full_join(const b2_hash_join*, const b2_table_view*, int32_t, size_t, b2_stream, b2_column**, b2_column**) { NOT_YET("b2_hash_join_full_join"); }
b2_status b2_hash_join_inner_join_size(const b2_hash_join*, const b2_table_view*, b2_stream, size_t*) { NOT_YET("b2_hash_join_inner_join_size"); }
b2_status b2_hash_join_left_join_size(const b2_hash_join*, const b2_table_view*, b2_stream, size_t*) { NOT_YET("b2_hash_join_left_join_size"); }
b2_status b2_hash_join_full_join_size(const b2_hash_join*, const b2_table_view*, b2_stream, size_t*) { NOT_YET("b2_hash_join_full_join_size"); }
b2_status b2_groupby_create(const b2_table_view*, int32_t, int32_t, const uint8_t*, int32_t, const uint8_t*, int32_t, b2_groupby**) { NOT_YET("b2_groupby_create"); }
void b2_groupby_destroy(b2_groupby*) {}
b2_status b2_groupby_aggregate(b2_groupby*, const b2_agg_request*, int32_t, b2_stream, b2_table**, b2_table**) { NOT_YET("b2_groupby_aggregate"); }
b2_status b2_groupby_scan(b2_groupby*, const b2_agg_request*, int32_t, b2_stream, b2_table**, b2_table**) { NOT_YET("b2_groupby_scan"); }
b2_status b2_partition(const b2_table_view*, const b2_column_view*, int32_t, const void*, int32_t, b2_stream, b2_table**, int32_t*) { NOT_YET("b2_partition"); }
}
