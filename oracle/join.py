"""Hash-join semantics of cpp/src/join/join.cu:27-110, hash_join/{hash_join.cu,retrieve_impl.cuh,
size_impl.cuh,dispatch.cuh} and primitive_row_operators.cuh:121-143 restated as a sort-merge join.
Outputs are canonical: pairs sorted by (left, right). JoinNoMatch = INT32_MIN (join.hpp:72)."""
from __future__ import annotations

import numpy as np

EQUAL, UNEQUAL = 0, 1
JOIN_NO_MATCH = np.iinfo(np.int32).min


def _norm(values: np.ndarray) -> np.ndarray:
    v = np.asarray(values)
    if v.dtype == np.bool_:
        return v.astype(np.uint8)
    if v.dtype.kind == "f":
        v = np.where(v == 0, 0.0, v).astype(v.dtype)  # -0 == +0 under operator==
    return v


def _row_ids(left_cols, right_cols, nulls_equal):
    """Map every row of both tables to an integer id such that equal rows (row_equality with NaN==NaN,
    common_utils.cuh:214-220; nulls equal iff EQUAL) share the id; -1 = row can never match."""
    if len(left_cols) != len(right_cols):
        raise ValueError("Mismatch in number of columns to be joined on")  # std::invalid_argument
    nl = len(left_cols[0][0]) if left_cols else 0
    nr = len(right_cols[0][0]) if right_cols else 0
    if len(left_cols) == 1 and left_cols[0][1] is None and right_cols[0][1] is None and np.asarray(left_cols[0][0]).dtype.kind in "iu" \
            and np.asarray(left_cols[0][0]).dtype == np.asarray(right_cols[0][0]).dtype and nl + nr > 1_000_000:
        # large single integer key column without nulls: rank the values directly (same ids, minutes faster than the row-wise unique)
        v = np.concatenate([np.asarray(left_cols[0][0]), np.asarray(right_cols[0][0])])
        _, ids = np.unique(v, return_inverse=True)
        return ids[:nl].astype(np.int64), ids[nl:].astype(np.int64)
    ids = np.zeros(nl + nr, dtype=np.int64)
    dead = np.zeros(nl + nr, dtype=bool)
    for (lv, lm), (rv, rm) in zip(left_cols, right_cols):
        if np.asarray(lv).dtype != np.asarray(rv).dtype:
            raise TypeError("Mismatch in joining column data types")  # cudf::data_type_error
        v = np.concatenate([_norm(lv), _norm(rv)])
        valid = np.concatenate([lm if lm is not None else np.ones(nl, bool), rm if rm is not None else np.ones(nr, bool)])
        if v.dtype.kind == "f":
            nan = np.isnan(v)
            _, inv = np.unique(np.where(nan, 0, v), return_inverse=True)
            col_id = np.where(nan, inv.max(initial=0) + 1, inv)
        else:
            _, col_id = np.unique(v, return_inverse=True)
        col_id = np.where(valid, col_id + 1, 0)  # 0 = null
        if nulls_equal == UNEQUAL:
            dead |= ~valid
        _, ids = np.unique(np.stack([ids, col_id], axis=1), axis=0, return_inverse=True)
        ids = ids.reshape(-1)
    ids = np.where(dead, -1, ids)
    return ids[:nl], ids[nl:]


def _matches(lid, rid):
    order = np.argsort(rid, kind="stable")
    rs = rid[order]
    lo = np.searchsorted(rs, lid, side="left")
    hi = np.searchsorted(rs, lid, side="right")
    cnt = np.where(lid >= 0, hi - lo, 0)
    left = np.repeat(np.arange(len(lid), dtype=np.int64), cnt)
    starts = np.repeat(lo, cnt)
    within = np.arange(cnt.sum(), dtype=np.int64) - np.repeat(np.cumsum(cnt) - cnt, cnt)
    right = order[starts + within] if len(left) else np.empty(0, dtype=np.int64)
    return left, right, cnt


def _canon(left, right):
    o = np.lexsort((right, left))
    return left[o].astype(np.int32), right[o].astype(np.int32)


def inner_join(left_cols, right_cols, nulls_equal=EQUAL):
    lid, rid = _row_ids(left_cols, right_cols, nulls_equal)
    l, r, _ = _matches(lid, rid)
    return _canon(l, r)


def left_join(left_cols, right_cols, nulls_equal=EQUAL):
    lid, rid = _row_ids(left_cols, right_cols, nulls_equal)
    l, r, cnt = _matches(lid, rid)
    un = np.nonzero(cnt == 0)[0]
    l = np.concatenate([l, un])
    r = np.concatenate([r, np.full(len(un), JOIN_NO_MATCH, dtype=np.int64)])
    return _canon(l, r)


def full_join(left_cols, right_cols, nulls_equal=EQUAL):
    lid, rid = _row_ids(left_cols, right_cols, nulls_equal)
    l, r, cnt = _matches(lid, rid)
    un = np.nonzero(cnt == 0)[0]
    matched_r = np.zeros(len(rid), dtype=bool)
    matched_r[r] = True
    ur = np.nonzero(~matched_r)[0]
    l = np.concatenate([l, un, np.full(len(ur), JOIN_NO_MATCH, dtype=np.int64)])
    r = np.concatenate([r, np.full(len(un), JOIN_NO_MATCH, dtype=np.int64), ur])
    return _canon(l, r)


def inner_join_size(left_cols, right_cols, nulls_equal=EQUAL) -> int:
    lid, rid = _row_ids(left_cols, right_cols, nulls_equal)
    return int(_matches_count(lid, rid))


def match_counts(left_cols, right_cols, nulls_equal=EQUAL, kind="inner") -> np.ndarray:
    """hash_join::{inner,left,full}_join_match_context (cpp/include/cudf/join/hash_join.hpp:254-330): matching build
    rows per left row as int32; for left / full a row without a match counts 1 (join_tests.cpp:2403-2492)."""
    lid, rid = _row_ids(left_cols, right_cols, nulls_equal)
    rs = np.sort(rid)
    c = np.where(lid >= 0, np.searchsorted(rs, lid, side="right") - np.searchsorted(rs, lid, side="left"), 0)
    if kind != "inner":
        c = np.maximum(c, 1)
    return c.astype(np.int32)


def _matches_count(lid, rid):
    rs = np.sort(rid)
    lo = np.searchsorted(rs, lid, side="left")
    hi = np.searchsorted(rs, lid, side="right")
    return np.where(lid >= 0, hi - lo, 0).sum()


def canonical(left_idx, right_idx):
    """Canonical form of a gather-map pair produced by any implementation."""
    return _canon(np.asarray(left_idx, dtype=np.int64), np.asarray(right_idx, dtype=np.int64))
