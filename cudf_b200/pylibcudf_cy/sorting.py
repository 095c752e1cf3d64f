"""pylibcudf.sorting (python/pylibcudf/pylibcudf/sorting.pyx:37-79,333-520): compiled in _core.pyx."""
from ._core import (rank, segmented_sort_by_key, segmented_sorted_order, sort, sort_by_key, sorted_order, stable_segmented_sort_by_key,
                    stable_segmented_sorted_order, stable_sort, stable_sort_by_key, stable_sorted_order, top_k, top_k_order)

__all__ = ["sorted_order", "stable_sorted_order", "sort", "stable_sort", "sort_by_key", "stable_sort_by_key", "segmented_sorted_order",
           "stable_segmented_sorted_order", "segmented_sort_by_key", "stable_segmented_sort_by_key", "top_k", "top_k_order", "rank"]
