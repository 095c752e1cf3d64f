"""Aggregation factories (python/pylibcudf/pylibcudf/aggregation.pyx:234-369; kinds from
cpp/include/cudf/aggregation.hpp:78-121). Only the kinds on the hot path exist."""
from __future__ import annotations

import enum

from .types import NullPolicy


class Kind(enum.IntEnum):
    SUM = 0
    PRODUCT = 2
    MIN = 3
    MAX = 4
    COUNT_VALID = 5
    COUNT_ALL = 6
    SUM_OF_SQUARES = 9
    MEAN = 10
    M2 = 11
    VARIANCE = 12
    STD = 13
    MEDIAN = 14
    ARGMAX = 16
    ARGMIN = 17
    NUNIQUE = 18
    NTH_ELEMENT = 19


class Aggregation:
    __slots__ = ("_kind", "_ddof")

    def __init__(self, kind: Kind, ddof: int | None = None):
        self._kind = Kind(kind)
        self._ddof = ddof

    def kind(self) -> Kind:
        return self._kind

    def abi_kind(self) -> int:
        """The kind word of the C ABI: B2_AGG_WITH_DDOF(kind, parameter) — ddof for VARIANCE / STD, the (signed) n for
        NTH_ELEMENT (include/cudf_b200.h)."""
        if self._ddof is None:
            return int(self._kind)
        return int(self._kind) | (1 << 30) | ((int(self._ddof) & 0xFFFF) << 8)

    def __repr__(self):
        return f"Aggregation({self._kind.name})"


def sum() -> Aggregation:  # noqa: A001
    return Aggregation(Kind.SUM)


def product() -> Aggregation:
    return Aggregation(Kind.PRODUCT)


def min() -> Aggregation:  # noqa: A001
    return Aggregation(Kind.MIN)


def max() -> Aggregation:  # noqa: A001
    return Aggregation(Kind.MAX)


def mean() -> Aggregation:
    return Aggregation(Kind.MEAN)


def count(null_handling: NullPolicy = NullPolicy.EXCLUDE) -> Aggregation:
    return Aggregation(Kind.COUNT_VALID if null_handling == NullPolicy.EXCLUDE else Kind.COUNT_ALL)


def sum_of_squares() -> Aggregation:
    return Aggregation(Kind.SUM_OF_SQUARES)


def m2() -> Aggregation:
    return Aggregation(Kind.M2)


def variance(ddof: int = 1) -> Aggregation:
    return Aggregation(Kind.VARIANCE, ddof)


def std(ddof: int = 1) -> Aggregation:
    return Aggregation(Kind.STD, ddof)


def argmax() -> Aggregation:
    return Aggregation(Kind.ARGMAX)


def argmin() -> Aggregation:
    return Aggregation(Kind.ARGMIN)


def median() -> Aggregation:
    """make_median_aggregation (sort-based groupby path)."""
    return Aggregation(Kind.MEDIAN)


def nunique(null_handling: NullPolicy = NullPolicy.EXCLUDE) -> Aggregation:
    """make_nunique_aggregation: distinct valid values per group (sort-based groupby path; nulls are not counted)."""
    if null_handling != NullPolicy.EXCLUDE:
        raise ValueError("nunique: only null_policy::EXCLUDE is supported on this path")
    return Aggregation(Kind.NUNIQUE)


def nth_element(n: int, null_handling: NullPolicy = NullPolicy.INCLUDE) -> Aggregation:
    """make_nth_element_aggregation(n): the n-th row of each group in input order (negative n counts from the end); a group
    that is too short gives null (sort-based groupby path; null_policy::INCLUDE)."""
    if null_handling != NullPolicy.INCLUDE:
        raise ValueError("nth_element: only null_policy::INCLUDE is supported on this path")
    if not -32768 <= int(n) <= 32767:
        raise ValueError("nth_element: n must fit 16 bits on this path")
    return Aggregation(Kind.NTH_ELEMENT, int(n))
