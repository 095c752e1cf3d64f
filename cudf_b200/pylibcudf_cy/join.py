"""pylibcudf.join (python/pylibcudf/pylibcudf/join.pyx:63-205) + cudf::hash_join with its match contexts: compiled in _core.pyx."""
from ._core import HashJoin, JoinMatchContext, JoinPartitionContext, full_join, inner_join, left_join

__all__ = ["inner_join", "left_join", "full_join", "HashJoin", "JoinMatchContext", "JoinPartitionContext"]
