"""pylibcudf.reduce (python/pylibcudf/pylibcudf/reduce.pyx:48-157) + segmented_reduce: compiled in _core.pyx."""
from ..pylibcudf.reduce import ScanType
from ._core import reduce, scan, segmented_reduce

__all__ = ["reduce", "scan", "segmented_reduce", "ScanType"]
