"""Sort golden vectors — transcribed from /root/reference/cpp/tests/sort/{sort_test.cpp,stable_sort_tests.cpp}.

Where the reference test mixes in a strings column, the strings are replaced by integers with the
same relative order ("a"<"d"<"e"<"k" -> 1<4<5<11), which leaves the expected permutation unchanged.
T = the typed-test element type; cases marked `skip_bool` are the ones the reference itself does not
pin for bool.  N = None (null)."""
import math

NaN, Inf = math.nan, math.inf
ASC, DESC = 0, 1
AFTER, BEFORE = 0, 1
N = None

# name, columns [(values, kind)], column_order, null_precedence, expected sorted_order, citation
# kind "T": cast to the typed-test type; "i": int32 stand-in for the strings column
SORTED_ORDER_CASES = [
    dict(name="WithNullMax", cite="sort_test.cpp:50-85",
         cols=[([5, 4, N, 5, 8, 5], "T"), ([4, 5, N, 4, 11, 4], "i"), ([10, 40, N, 5, 2, 10], "T")],
         order=[ASC, ASC, DESC], nulls=[AFTER, AFTER, AFTER], expected=[1, 0, 5, 3, 4, 2], skip_bool=True),
    dict(name="WithNullMin", cite="sort_test.cpp:87-117",
         cols=[([5, 4, N, 5, 8], "T"), ([4, 5, N, 4, 11], "i"), ([10, 40, N, 5, 2], "T")],
         order=[ASC, ASC, DESC], nulls=[], expected=[2, 1, 0, 3, 4], skip_bool=True),
    dict(name="WithMixedNullOrder", cite="sort_test.cpp:119-148",
         cols=[([N, N, 3, 5, N], "T"), ([N, 5, N, N, 11], "i"), ([10, N, 70, N, 2], "T")],
         order=[ASC, ASC, ASC], nulls=[AFTER, BEFORE, AFTER], expected=[2, 3, 0, 1, 4], skip_bool=True),
    dict(name="WithAllValid", cite="sort_test.cpp:150-175",
         cols=[([5, 4, 3, 5, 8], "T"), ([4, 5, 1, 4, 11], "i"), ([10, 40, 70, 5, 2], "T")],
         order=[ASC, ASC, DESC], nulls=[], expected=[2, 1, 0, 3, 4], skip_bool=True),
    dict(name="StableMixedNullOrder", cite="stable_sort_tests.cpp:41-58",
         # strings {"2","a","b","x",null,"a","x","a"} -> ranks 2<a<b<x : 0,1,2,3,N,1,3,1
         cols=[([N, 1, 1, 0, 0, 1, 0, 1], "T"), ([0, 1, 2, 3, N, 1, 3, 1], "i")],
         order=[ASC, ASC], nulls=[AFTER, BEFORE], expected=[4, 3, 6, 1, 5, 7, 2, 0], skip_bool=False),
    dict(name="StableWithNullMax", cite="stable_sort_tests.cpp:60-86",
         cols=[([5, 4, N, 5, 8, 5], "T"), ([4, 5, N, 4, 11, 4], "i"), ([10, 40, N, 10, 2, 10], "T")],
         order=[ASC, ASC, DESC], nulls=[AFTER, AFTER, AFTER], expected=[1, 0, 3, 5, 4, 2], skip_bool=True),
]

# sort_test.cpp:475-508 SlicedColumns: strings -> ranks  ""(null) , aaa=1 < ab=2 < abc=3 < b=4 < bbe=5 < za=6
SLICED = dict(cite="sort_test.cpp:475-508",
              col1=[5, 5, 1, 3, 2, 6, 4, N], col2=[7, 8, 1, 1, 9, 5, 7, 3],
              expected=[7, 2, 4, 3, 6, 0, 1, 5], split=3, expected_sliced=[4, 1, 0, 3, 2])

# stable_sort_tests.cpp:88-123: results are compared after gather (values, not indices)
SINGLE_NO_NULL = dict(cite="stable_sort_tests.cpp:88-103", values=[7, 1, -2, 5, 1, 0, 1, -2, 0, 5],
                      expected_signed=[2, 7, 5, 8, 1, 4, 6, 3, 9, 0], expected_unsigned=[5, 8, 1, 4, 6, 3, 9, 0, 2, 7])
SINGLE_WITH_NULL = dict(cite="stable_sort_tests.cpp:105-123", values=[7, 1, -2, 5, 1, 0, 1, -2, 0, 5],
                        valid=[1, 1, 0, 0, 1, 0, 1, 0, 1, 0],
                        expected_signed=[2, 7, 5, 3, 9, 8, 1, 4, 6, 0], expected_unsigned=[5, 3, 9, 2, 7, 8, 1, 4, 6, 0])

# sort_test.cpp:1070-1083 and stable_sort_tests.cpp:277-289
INF_NAN = dict(cite="sort_test.cpp:1070-1083", values=[-0.0, -NaN, -NaN, NaN, Inf, -Inf, 7.0, 5.0, 6.0, NaN, Inf, -Inf, -NaN, -NaN, -0.0],
               expected=[5, 11, 0, 14, 7, 8, 6, 4, 10, 1, 2, 3, 9, 12, 13])
