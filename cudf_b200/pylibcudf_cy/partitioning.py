"""pylibcudf.partitioning (python/pylibcudf/pylibcudf/partitioning.pyx; cpp/include/cudf/partitioning.hpp:58-175): compiled in _core.pyx."""
from ..pylibcudf.partitioning import DEFAULT_HASH_SEED, HashId
from ._core import hash_partition, partition

__all__ = ["hash_partition", "partition", "HashId", "DEFAULT_HASH_SEED"]
