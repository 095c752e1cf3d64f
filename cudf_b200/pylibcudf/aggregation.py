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
    MEAN = 10


class Aggregation:
    __slots__ = ("_kind",)

    def __init__(self, kind: Kind):
        self._kind = Kind(kind)

    def kind(self) -> Kind:
        return self._kind

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
