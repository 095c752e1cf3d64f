"""Sort semantics of cpp/src/sort/{sort.cu,sort_impl.cuh,sort_column.cu,sort_column_impl.cuh,
sorted_order_radix.cu,sort_radix.cu,stable_sort.cu} restated with numpy (dense ranks + stable lexsort;
deliberately NOT the radix/twiddle formulation the CUDA path uses)."""
from __future__ import annotations

import numpy as np

ASCENDING, DESCENDING = 0, 1
AFTER, BEFORE = 0, 1


def _dense_rank(values: np.ndarray) -> tuple[np.ndarray, int]:
    """Rank under relational_compare (cpp/include/cudf/detail/row_operator/common_utils.cuh:157-169):
    -0 == +0, NaN greater than everything, NaN == NaN."""
    v = np.asarray(values)
    if v.dtype == np.bool_:
        v = v.astype(np.uint8)
    if v.dtype.kind == "f":
        v = np.where(v == 0, 0.0, v).astype(v.dtype)  # -0.0 -> +0.0
        nan = np.isnan(v)
        uniq, inv = np.unique(v[~nan], return_inverse=True)
        rank = np.empty(len(v), dtype=np.int64)
        rank[~nan] = inv
        rank[nan] = len(uniq)
        return rank, len(uniq) + 1
    uniq, inv = np.unique(v, return_inverse=True)
    return inv.astype(np.int64), len(uniq)


def _column_rank(values, valid, order, null_order) -> np.ndarray:
    rank, k = _dense_rank(values)
    if order == DESCENDING:
        rank = (k - 1) - rank
    if valid is not None and not valid.all():
        # sort_column_impl.cuh:35-57: the null flags are swapped for DESCENDING
        nulls_first = (null_order == BEFORE) != (order == DESCENDING)
        rank = np.where(valid, rank, -1 if nulls_first else k)
    return rank


def sorted_order(columns, column_order=None, null_precedence=None) -> np.ndarray:
    """cudf::sorted_order / stable_sorted_order (sort_impl.cuh:31-96). columns: list of (values, valid)."""
    if not columns or len(columns[0][0]) == 0:
        return np.empty(0, dtype=np.int32)
    ncol = len(columns)
    if column_order and len(column_order) != ncol:
        raise RuntimeError("Mismatch between number of columns and column order.")  # cudf::logic_error
    if null_precedence and len(null_precedence) != ncol:
        raise RuntimeError("Mismatch between number of columns and null_precedence size.")
    order = list(column_order) if column_order else [ASCENDING] * ncol
    nprec = list(null_precedence) if null_precedence else [BEFORE] * ncol  # sort_impl.cuh:56-57
    ranks = [_column_rank(v, m, o, p) for (v, m), o, p in zip(columns, order, nprec)]
    n = len(ranks[0])
    out = np.lexsort(tuple(reversed(ranks))).astype(np.int32) if ncol > 1 else np.argsort(ranks[0], kind="stable").astype(np.int32)
    # single non-null float column, DESCENDING: the radix path sorts the tuple (isnan*(idx+1), f)
    # descending (sorted_order_radix.cu:41-50,121-131) => NaNs come first in DESCENDING row order.
    v0, m0 = columns[0]
    if ncol == 1 and order[0] == DESCENDING and np.asarray(v0).dtype.kind == "f" and (m0 is None or m0.all()):
        k = int(np.isnan(v0).sum())
        out[:k] = out[:k][::-1]
    assert len(out) == n
    return out


def gather(columns, gather_map, nullify_oob=False):
    """cudf::gather (cpp/include/cudf/detail/gather.cuh:627-675): negative indices wrap once."""
    out = []
    gm = np.asarray(gather_map, dtype=np.int64)
    for values, valid in columns:
        n = len(values)
        idx = np.where(gm < 0, gm + n, gm)
        inb = (idx >= 0) & (idx < n)
        safe = np.where(inb, idx, 0)
        vals = np.asarray(values)[safe] if n else np.zeros(len(gm), dtype=np.asarray(values).dtype)
        if valid is not None and not np.asarray(valid).all() or nullify_oob:
            v = (np.asarray(valid)[safe] if valid is not None else np.ones(len(gm), dtype=bool))
            if nullify_oob:
                v = v & inb
            out.append((vals, v))
        else:
            out.append((vals, None))
    return out


def sort_by_key(values, keys, column_order=None, null_precedence=None):
    """cudf::sort_by_key (sort.cu:31-50)."""
    nv = len(values[0][0]) if values else 0
    nk = len(keys[0][0]) if keys else 0
    if nv != nk:
        raise RuntimeError("Mismatch in number of rows for values and keys")
    return gather(values, sorted_order(keys, column_order, null_precedence))


def sort(columns, column_order=None, null_precedence=None):
    """cudf::sort (sort.cu:52-67)."""
    return sort_by_key(columns, columns, column_order, null_precedence)


def segmented_sorted_order(columns, segment_offsets, column_order=None, null_precedence=None) -> np.ndarray:
    """cudf::{stable_,}segmented_sorted_order (cpp/include/cudf/sorting.hpp:232-296): every segment
    [offsets[i], offsets[i+1]) is sorted on its own; rows outside the segments keep their place; fewer than two offsets
    sort nothing. Ties are broken stably (the unstable variant leaves them unspecified)."""
    if not columns:
        return np.empty(0, dtype=np.int32)
    offs = np.asarray(segment_offsets)
    if offs.dtype != np.int32:
        raise RuntimeError("segment offsets should be size_type")  # cudf::logic_error
    n = len(columns[0][0])
    out = np.arange(n, dtype=np.int32)
    for b, e in zip(offs[:-1], offs[1:]):
        b, e = int(b), int(e)
        if e - b > 1:
            seg = [(np.asarray(v)[b:e], None if m is None else np.asarray(m)[b:e]) for v, m in columns]
            out[b:e] = b + sorted_order(seg, column_order, null_precedence)
    return out


def segmented_sort_by_key(values, keys, segment_offsets, column_order=None, null_precedence=None):
    nv = len(values[0][0]) if values else 0
    nk = len(keys[0][0]) if keys else 0
    if nv != nk:
        raise RuntimeError("Mismatch in number of rows for values and keys")
    return gather(values, segmented_sorted_order(keys, segment_offsets, column_order, null_precedence))


def top_k_order(column, k, order=DESCENDING) -> np.ndarray:
    """cudf::top_k_order (cpp/src/sort/top_k.cu:143-170): the first k rows of the stable sorted order, nulls last for
    ASCENDING and first for DESCENDING; the reference may return them in any order."""
    if k < 0:
        raise ValueError("k must be non-negative")
    if k == 0 or len(column[0]) == 0:
        return np.empty(0, dtype=np.int32)
    o = sorted_order([column], [order], [AFTER if order == ASCENDING else BEFORE])
    return o[: min(k, len(o))]


def top_k(column, k, order=DESCENDING):
    return gather([column], top_k_order(column, k, order))[0]


def rank(column, method, column_order=ASCENDING, null_handling=1, null_precedence=AFTER, percentage=False):
    """cudf::rank (cpp/src/sort/rank.cu:236-356, sorting.hpp:165-230). method: 0 FIRST, 1 AVERAGE, 2 MIN, 3 MAX, 4 DENSE;
    null_handling: 0 EXCLUDE (result carries the input validity), 1 INCLUDE. -> (values, valid | None)."""
    values, valid = column
    values = np.asarray(values)
    n = len(values)
    as_double = percentage or method == 1
    if n == 0:
        return np.empty(0, np.float64 if as_double else np.int32), None
    m = np.ones(n, bool) if valid is None else np.asarray(valid, bool)
    o = sorted_order([(values, None if valid is None else m)], [column_order], [null_precedence]).astype(np.int64)
    sv, sm = values[o], m[o]
    same = np.zeros(n, bool)
    if n > 1:
        a, b = sv[1:], sv[:-1]
        eqv = (a == b) | ((a != a) & (b != b)) if values.dtype.kind == "f" else (a == b)
        same[1:] = (sm[1:] & sm[:-1] & eqv) | (~sm[1:] & ~sm[:-1])
    dense = np.cumsum(~same)
    starts = np.nonzero(~same)[0]
    ends = np.concatenate([starts[1:], [n]]) - 1
    gfirst, glast = starts[dense - 1], ends[dense - 1]
    pos = np.arange(n)
    r = {0: pos + 1.0, 1: gfirst + 1 + (glast - gfirst) / 2.0, 2: gfirst + 1.0, 3: glast + 1.0, 4: dense.astype(np.float64)}[method]
    if percentage:
        count = int(m.sum()) if null_handling == 0 else n
        count = count if count > 0 else n
        r = r / (dense[count - 1] if method == 4 else count)
    out = np.empty(n, np.float64 if as_double else np.int32)
    out[o] = r if as_double else r.astype(np.int32)
    ov = None
    if null_handling == 0 and valid is not None and not m.all():
        ov = m.copy()
    return out, ov
