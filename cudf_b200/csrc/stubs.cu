// stubs.cu — entry points declared in include/cudf_b200.h whose kernels are not written yet.
// They fail loudly (B2_ERR_LOGIC) instead of silently falling back to anything.
#include "common.cuh"
#define NOT_YET(name) do { b2::set_last_error(name ": not implemented yet"); return B2_ERR_LOGIC; } while (0)
extern "C" {
b2_status b2_partition(const b2_table_view*, const b2_column_view*, int32_t, const void*, int32_t, b2_stream, b2_table**, int32_t*) { NOT_YET("b2_partition"); }
}
