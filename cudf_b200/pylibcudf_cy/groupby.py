"""pylibcudf.groupby (python/pylibcudf/pylibcudf/groupby.pyx:36-243): compiled in _core.pyx."""
from ._core import GroupBy, GroupByRequest

__all__ = ["GroupBy", "GroupByRequest"]
