"""pylibcudf-named API of the B200 hot path (sorting, join, groupby, reduce, copying, aggregation, null_mask)."""
from . import aggregation, contiguous_split, copying, groupby, interop, join, null_mask, partitioning, reduce, sorting, types
from .column import Column, DeviceSpan, Scalar, Table
from .types import DataType, MaskState, NullEquality, NullOrder, NullPolicy, Order, OutOfBoundsPolicy, RankMethod, Sorted, TypeId

__all__ = [
    "aggregation", "contiguous_split", "copying", "groupby", "interop", "join", "null_mask", "partitioning", "reduce", "sorting", "types",
    "Column", "DeviceSpan", "Scalar", "Table", "DataType", "MaskState", "NullEquality", "NullOrder", "NullPolicy",
    "Order", "OutOfBoundsPolicy", "RankMethod", "Sorted", "TypeId",
]
