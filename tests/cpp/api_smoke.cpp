// Reads like a libcudf gtest (cpp/tests/sort/sort_test.cpp:50-85, cpp/tests/join/join_tests.cpp:2316-2337,
// cpp/tests/groupby/sum_tests.cpp:68-80): exercises the cudf:: C++ surface in include/cudf over the C ABI.
#include <cudf/column/column_view.hpp>
#include <cudf/contiguous_split.hpp>
#include <cudf/partitioning.hpp>
#include <cudf/groupby.hpp>
#include <cudf/join/hash_join.hpp>
#include <cudf/join/join.hpp>
#include <cudf/reduction.hpp>
#include <cudf/sorting.hpp>
#include <cudf/table/table_view.hpp>

#include <cuda_runtime_api.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <vector>

template <typename T>
struct dev_vec {
  T* p = nullptr;
  size_t n;
  explicit dev_vec(std::vector<T> const& h) : n(h.size()) { cudaMalloc(reinterpret_cast<void**>(&p), n * sizeof(T) + 16); cudaMemcpy(p, h.data(), n * sizeof(T), cudaMemcpyHostToDevice); }
  ~dev_vec() { cudaFree(p); }
};
template <typename T>
std::vector<T> to_host(T const* d, size_t n) { std::vector<T> h(n); cudaDeviceSynchronize(); cudaMemcpy(h.data(), d, n * sizeof(T), cudaMemcpyDeviceToHost); return h; }
#define EXPECT(c) do { if (!(c)) { std::printf("FAILED: %s (line %d)\n", #c, __LINE__); return 1; } } while (0)

int main()
{
  using namespace cudf;
  // sorted_order / sort_by_key
  dev_vec<int64_t> k({5, 4, 3, 5, 8, 5});
  column_view kc{data_type{type_id::INT64}, 6, k.p};
  auto order = sorted_order(table_view{{kc}}, {order::ASCENDING});
  auto ho = to_host(order->view().data<int32_t>(), 6);
  EXPECT((ho == std::vector<int32_t>{2, 1, 0, 3, 5, 4}));
  auto sorted = sort_by_key(table_view{{kc}}, table_view{{kc}}, {order::DESCENDING});
  EXPECT((to_host(sorted->get_column(0).view().data<int64_t>(), 6) == std::vector<int64_t>{8, 5, 5, 5, 4, 3}));
  try { sorted_order(table_view{{kc, kc}}, {order::ASCENDING}); EXPECT(false); } catch (cudf::logic_error const&) {}
  {  // segmented sort (sorting.hpp:238-244) and top-k
    dev_vec<int32_t> sk({9, 8, 7, 6, 5, 4, 3, 2, 1, 0}), so({0, 3, 7, 10});
    column_view skc{data_type{type_id::INT32}, 10, sk.p}, soc{data_type{type_id::INT32}, 4, so.p};
    auto seg = segmented_sorted_order(table_view{{skc}}, soc);
    EXPECT((to_host(seg->view().data<int32_t>(), 10) == std::vector<int32_t>{2, 1, 0, 6, 5, 4, 3, 9, 8, 7}));
    dev_vec<int32_t> rk({3, 4, 5, 4, 1, 2});  // sorting.hpp:172-179: AVERAGE = {3, 4.5, 6, 4.5, 1, 2}
    column_view rkc{data_type{type_id::INT32}, 6, rk.p};
    auto avg = cudf::rank(rkc, rank_method::AVERAGE, order::ASCENDING, null_policy::INCLUDE, null_order::AFTER, false);
    EXPECT((to_host(avg->view().data<double>(), 6) == std::vector<double>{3, 4.5, 6, 4.5, 1, 2}));
    auto top = top_k(skc, 3);
    auto ht  = to_host(top->view().data<int32_t>(), 3);
    std::sort(ht.begin(), ht.end());
    EXPECT((ht == std::vector<int32_t>{7, 8, 9}));
  }
  // hash join gold maps (join_tests.cpp:2316-2337)
  dev_vec<int32_t> l({3, 1, 2, 0, 2}), r({2, 2, 0, 4, 3});
  column_view lc{data_type{type_id::INT32}, 5, l.p}, rc{data_type{type_id::INT32}, 5, r.p};
  cudf::hash_join hj(table_view{{rc}}, null_equality::EQUAL);
  auto [li, ri] = hj.inner_join(table_view{{lc}});
  EXPECT(li->size() == 6 && hj.inner_join_size(table_view{{lc}}) == 6);
  auto hl = to_host(li->data(), 6), hr = to_host(ri->data(), 6);
  std::vector<std::pair<int, int>> pairs;
  for (int i = 0; i < 6; ++i) pairs.emplace_back(hl[i], hr[i]);
  std::sort(pairs.begin(), pairs.end());
  EXPECT((pairs == std::vector<std::pair<int, int>>{{0, 4}, {2, 0}, {2, 1}, {3, 2}, {4, 0}, {4, 1}}));
  try { cudf::hash_join bad(table_view{{rc}}, nullable_join::NO, null_equality::EQUAL, 1.5); EXPECT(false); } catch (std::invalid_argument const&) {}
  // match context + partitioned probes (join_tests.cpp:2363-2377: counts {1, 0, 2, 1, 2})
  {
    auto ctx = hj.inner_join_match_context(table_view{{lc}});
    EXPECT((to_host(ctx._match_counts->data(), 5) == std::vector<int32_t>{1, 0, 2, 1, 2}));
    cudf::join_partition_context part{std::make_unique<cudf::join_match_context>(std::move(ctx)), 0, 3};
    auto [pl, pr] = hj.partitioned_inner_join(part);
    EXPECT(pl->size() == 3);  // rows 0..2 of the probe: (0,4), (2,0), (2,1)
    part.left_start_idx = 3; part.left_end_idx = 5;
    auto [ql, qr] = hj.partitioned_inner_join(part);
    EXPECT(ql->size() == 3);  // (3,2), (4,0), (4,1)
    part.left_start_idx = 4; part.left_end_idx = 9;
    try { (void)hj.partitioned_inner_join(part); EXPECT(false); } catch (std::invalid_argument const&) {}
    auto fctx = hj.full_join_match_context(table_view{{lc}});
    cudf::join_partition_context fpart{std::make_unique<cudf::join_match_context>(std::move(fctx)), 0, 5};
    auto [fl, fr] = hj.partitioned_full_join(fpart);
    auto [gl, gr] = cudf::hash_join::finalize_partitioned_full_join({{fl->data(), fl->size()}}, {{fr->data(), fr->size()}}, 5, 5);
    EXPECT(gl->size() == hj.full_join_size(table_view{{lc}}));
  }
  // groupby sum (sum_tests.cpp:68-80)
  dev_vec<int32_t> gk({1, 2, 3, 1, 2, 2, 1, 3, 3, 2});
  dev_vec<double> gv({0, 1, 2, 3, 4, 5, 6, 7, 8, 9});
  column_view gkc{data_type{type_id::INT32}, 10, gk.p}, gvc{data_type{type_id::FLOAT64}, 10, gv.p};
  groupby::groupby gb(table_view{{gkc}});
  std::vector<groupby::aggregation_request> reqs(1);
  reqs[0].values = gvc;
  reqs[0].aggregations.push_back(make_sum_aggregation<groupby_aggregation>());
  reqs[0].aggregations.push_back(make_count_aggregation<groupby_aggregation>());
  reqs[0].aggregations.push_back(make_variance_aggregation<groupby_aggregation>());      // var_tests.cpp:30-41: {9, 131/12, 31/3}
  reqs[0].aggregations.push_back(make_std_aggregation<groupby_aggregation>(0));
  auto [gkeys, gres] = gb.aggregate(reqs);
  EXPECT(gkeys->num_rows() == 3);
  auto hk = to_host(gkeys->get_column(0).view().data<int32_t>(), 3);
  auto hs = to_host(gres[0].results[0]->view().data<double>(), 3);
  auto hc = to_host(gres[0].results[1]->view().data<int32_t>(), 3);
  for (int i = 0; i < 3; ++i) {
    double es = hk[i] == 1 ? 9 : (hk[i] == 2 ? 19 : 17);
    int ec = hk[i] == 2 ? 4 : 3;
    EXPECT(hs[i] == es && hc[i] == ec);
  }
  {
    auto hv = to_host(gres[0].results[2]->view().data<double>(), 3);
    auto hd = to_host(gres[0].results[3]->view().data<double>(), 3);
    for (int i = 0; i < 3; ++i) {
      double ev = hk[i] == 1 ? 9.0 : (hk[i] == 2 ? 131.0 / 12 : 31.0 / 3);
      int n = hk[i] == 2 ? 4 : 3;
      EXPECT(std::abs(hv[i] - ev) < 1e-9 && std::abs(hd[i] * hd[i] - ev * (n - 1) / n) < 1e-9);
    }
  }
  // reduce + scan
  auto s = reduce(gvc, *make_sum_aggregation<reduce_aggregation>(), data_type{type_id::FLOAT64});
  EXPECT(s->is_valid());
  auto sc = scan(gkc, *make_sum_aggregation<scan_aggregation>(), scan_type::INCLUSIVE);
  EXPECT((to_host(sc->view().data<int32_t>(), 10).back() == 20));
  {  // partition (partition_test.cpp:166-181 Reverse), hash_partition contract, pack / unpack round trip (pack_tests.cpp:69-75)
    dev_vec<int32_t> pv({0, 1, 3, 7, 5, 13}), pm({5, 4, 3, 2, 1, 0});
    column_view pvc{data_type{type_id::INT32}, 6, pv.p}, pmc{data_type{type_id::INT32}, 6, pm.p};
    auto [pt, poff] = cudf::partition(table_view{{pvc}}, pmc, 6);
    EXPECT((poff == std::vector<size_type>{0, 1, 2, 3, 4, 5, 6}));
    EXPECT((to_host(pt->get_column(0).view().data<int32_t>(), 6) == std::vector<int32_t>{13, 5, 7, 3, 1, 0}));
    auto [ht, hoff] = cudf::hash_partition(table_view{{pvc, pmc}}, std::vector<size_type>{0}, 3);
    EXPECT(hoff.size() == 4 && hoff.front() == 0 && hoff.back() == 6 && ht->num_rows() == 6);
    auto packed = cudf::pack(table_view{{pvc, pmc}});
    EXPECT(packed.gpu_data->size() == cudf::packed_size(table_view{{pvc, pmc}}) && packed.metadata->size() == 16 + 2 * 40);
    auto un = cudf::unpack(packed);
    EXPECT(un.num_columns() == 2 && un.num_rows() == 6);
    EXPECT((to_host(un.column(1).data<int32_t>(), 6) == std::vector<int32_t>{5, 4, 3, 2, 1, 0}));
  }
  std::printf("CPP_API_OK\n");
  return 0;
}
