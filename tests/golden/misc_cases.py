"""Join / groupby / reduce / scan / segmented-reduce known-answer vectors transcribed from the
reference gtests (file:line cited per case; N = null)."""
N = None
NO_MATCH = -(2 ** 31)  # cudf::JoinNoMatch (join.hpp:72)

# ---- join (cpp/tests/join/join_tests.cpp) -------------------------------------------------------
JOIN_GOLD_MAPS = dict(cite="join_tests.cpp:2316-2337 (HashJoinMemoryResource), col0 of InnerJoinNoNulls :1163-1237",
                      probe=[3, 1, 2, 0, 2], build=[2, 2, 0, 4, 3], left=[0, 2, 2, 3, 4, 4], right=[4, 0, 1, 2, 0, 1])
# HashJoinWithNullsOneSide :2191-2298 — two int32 key columns; results compared after sorting each
# index column independently
JOIN_NULLS_ONE_SIDE = dict(
    cite="join_tests.cpp:2191-2298",
    build=[[2, 2, 0, 4, 3], [1, 10, 1, 2, 1]],
    probe=[[1, 2, 3, 4, 5, 2, 2, 0, 4, 3, 1, 2, 3, 4, 5], [1, 2, 3, 4, 5, 1, N, 1, 2, 1, 1, 2, 3, 4, 5]],
    left_join=(list(range(15)), [NO_MATCH] * 11 + [0, 2, 3, 4]),
    inner_join=([5, 7, 8, 9], [0, 2, 3, 4]),
    full_join=([NO_MATCH] + list(range(15)), [NO_MATCH] * 11 + [0, 1, 2, 3, 4]),
)
JOIN_EQUAL_VALUES = dict(cite="join_tests.cpp:1906-1940 (numeric column only)", left=[0, 0], right=[0, 0], pairs=4)
JOIN_LARGE = dict(cite="join_tests.cpp:2299-2314", n=65567)

# ---- groupby (cpp/tests/groupby/*.cpp), keys int32 -------------------------------------------------
GB_KEYS = [1, 2, 3, 1, 2, 2, 1, 3, 3, 2]
GB_VALS = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9]
GROUPBY_CASES = [
    dict(name="sum_basic", cite="sum_tests.cpp:68-80", keys=GB_KEYS, vals=GB_VALS, kind="sum", ekeys=[1, 2, 3], evals=[9, 19, 17]),
    dict(name="sum_zero_valid_keys", cite="sum_tests.cpp:95-107", keys=[N, N, N], vals=[3, 4, 5], kind="sum", ekeys=[], evals=[]),
    dict(name="sum_zero_valid_values", cite="sum_tests.cpp:109-121", keys=[1, 1, 1], vals=[N, N, N], kind="sum", ekeys=[1], evals=[N]),
    dict(name="sum_null_keys_and_values", cite="sum_tests.cpp:123-143",
         keys=[1, 2, 3, 1, 2, 2, 1, N, 3, 2, 4], vals=[N, 1, 2, 3, 4, N, 6, 7, 8, 9, N], kind="sum",
         ekeys=[1, 2, 3, 4], evals=[9, 14, 10, N]),
    dict(name="count_basic", cite="count_tests.cpp:21-40", keys=GB_KEYS, vals=GB_VALS, kind="count", ekeys=[1, 2, 3], evals=[3, 4, 3]),
    dict(name="count_all_basic", cite="count_tests.cpp:37-39", keys=GB_KEYS, vals=GB_VALS, kind="count_all", ekeys=[1, 2, 3], evals=[3, 4, 3]),
    dict(name="mean_basic", cite="mean_tests.cpp:37-55", keys=GB_KEYS, vals=GB_VALS, kind="mean", ekeys=[1, 2, 3],
         evals=[3.0, 19.0 / 4, 17.0 / 3]),
    dict(name="empty", cite="sum_tests.cpp:82-93", keys=[], vals=[], kind="sum", ekeys=[], evals=[]),
    # min_tests.cpp / max_tests.cpp (every case there runs the hash and the sort implementation)
    dict(name="min_basic", cite="min_tests.cpp:25-41", keys=GB_KEYS, vals=GB_VALS, kind="min", ekeys=[1, 2, 3], evals=[0, 1, 2]),
    dict(name="min_zero_valid_keys", cite="min_tests.cpp:61-77", keys=[N, N, N], vals=[3, 4, 5], kind="min", ekeys=[], evals=[]),
    dict(name="min_zero_valid_values", cite="min_tests.cpp:79-95", keys=[1, 1, 1], vals=[N, N, N], kind="min", ekeys=[1], evals=[N]),
    dict(name="min_null_keys_and_values", cite="min_tests.cpp:97-119",
         keys=[1, 2, 3, 1, 2, 2, 1, N, 3, 2, 4], vals=[N, 1, 2, 3, 4, N, 6, 7, 8, 9, N], kind="min",
         ekeys=[1, 2, 3, 4], evals=[3, 1, 2, N]),
    dict(name="max_basic", cite="max_tests.cpp:27-46", keys=GB_KEYS, vals=GB_VALS, kind="max", ekeys=[1, 2, 3], evals=[6, 9, 8]),
    dict(name="max_null_keys_and_values", cite="max_tests.cpp:102-124",
         keys=[1, 2, 3, 1, 2, 2, 1, N, 3, 2, 4], vals=[0, 1, 2, 3, 4, 5, N, 7, 8, N, N], kind="max",
         ekeys=[1, 2, 3, 4], evals=[3, 5, 8, N]),
]
GROUPBY_SCAN_CASES = [
    dict(name="sum_scan_basic", cite="sum_scan_tests.cpp:33-49", keys=GB_KEYS, vals=GB_VALS, kind="sum",
         ekeys=[1, 1, 1, 2, 2, 2, 2, 3, 3, 3], evals=[0, 3, 9, 1, 5, 10, 19, 2, 9, 17]),
    dict(name="sum_scan_zero_valid_values", cite="sum_scan_tests.cpp:99-111", keys=[1, 1, 1], vals=[N, N, N], kind="sum",
         ekeys=[1, 1, 1], evals=[N, N, N]),
    dict(name="sum_scan_null_keys_and_values", cite="sum_scan_tests.cpp:113-130",
         keys=[1, 2, 3, 1, 2, 2, 1, N, 3, 2, 4], vals=[N, 1, 2, 3, 4, N, 6, 7, 8, 9, N], kind="sum",
         ekeys=[1, 1, 1, 2, 2, 2, 2, 3, 3, 4], evals=[N, 3, 9, 1, 5, N, 14, 2, 10, N]),
    dict(name="count_scan_basic", cite="count_scan_tests.cpp:27-46", keys=GB_KEYS, vals=GB_VALS, kind="count",
         ekeys=[1, 1, 1, 2, 2, 2, 2, 3, 3, 3], evals=[1, 2, 3, 1, 2, 3, 4, 1, 2, 3]),
    dict(name="min_scan_basic", cite="min_scan_tests.cpp:27-42", keys=GB_KEYS, vals=[5, 6, 7, 8, 9, 0, 1, 2, 3, 4], kind="min",
         ekeys=[1, 1, 1, 2, 2, 2, 2, 3, 3, 3], evals=[5, 5, 1, 6, 6, 0, 0, 7, 2, 2]),
    dict(name="min_scan_null_keys_and_values", cite="min_scan_tests.cpp:109-127",
         keys=[1, 2, 3, 1, 2, 2, 1, N, 3, 2, 4], vals=[N, 6, 7, 8, 9, N, 1, 2, 3, 4, N], kind="min",
         ekeys=[1, 1, 1, 2, 2, 2, 2, 3, 3, 4], evals=[N, 8, 1, 6, 6, N, 4, 7, 3, N]),
    dict(name="max_scan_basic", cite="max_scan_tests.cpp:29-45", keys=GB_KEYS, vals=[5, 6, 7, 8, 9, 0, 1, 2, 3, 4], kind="max",
         ekeys=[1, 1, 1, 2, 2, 2, 2, 3, 3, 3], evals=[5, 8, 8, 6, 9, 9, 9, 7, 7, 7]),
    dict(name="max_scan_null_keys_and_values", cite="max_scan_tests.cpp:112-130",
         keys=[1, 2, 3, 1, 2, 2, 1, N, 3, 2, 4], vals=[N, 6, 7, 8, 9, N, 1, 2, 3, 4, N], kind="max",
         ekeys=[1, 1, 1, 2, 2, 2, 2, 3, 3, 4], evals=[N, 8, 8, 6, 9, N, 9, 7, 7, N]),
]

# ---- scan (cpp/tests/reductions/scan_tests.cpp:160-216) ----------------------------------------------
SCAN_COL = [5, 4, 6, 0, 1, 6, 5, 3]
SCAN_CASES = [
    dict(name="InclusiveNoNulls", vals=SCAN_COL, inclusive=True, policy="EXCLUDE", expected=[5, 9, 15, 15, 16, 22, 27, 30]),
    dict(name="ExclusiveNoNulls", vals=SCAN_COL, inclusive=False, policy="EXCLUDE", expected=[0, 5, 9, 15, 15, 16, 22, 27]),
    dict(name="InclusiveWithNullsExclude", vals=[5, 4, 6, N, 1, 6, 5, 3], inclusive=True, policy="EXCLUDE",
         expected=[5, 9, 15, N, 16, 22, 27, 30]),
    dict(name="InclusiveWithNullsInclude", vals=[5, 4, 6, N, 1, 6, 5, 3], inclusive=True, policy="INCLUDE",
         expected=[5, 9, 15, N, N, N, N, N]),
]

# ---- segmented reduce (cpp/tests/reductions/segmented_reduction_tests.cpp:30-76) -----------------------
SEGMENTED_SUM = dict(vals=[1, 2, 3, 1, N, 3, 1, N, N, N], offsets=[0, 3, 6, 7, 8, 10, 10],
                     expected=[6, 4, 1, N, N, N], init=3, expected_init=[9, 7, 4, 3, 3, 3])
