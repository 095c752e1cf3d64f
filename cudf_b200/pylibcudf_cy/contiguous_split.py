"""pylibcudf.contiguous_split (python/pylibcudf/pylibcudf/contiguous_split.pyx): pack / unpack compiled in _core.pyx."""
from ..pylibcudf.contiguous_split import PackedColumns
from ._core import pack, pack_metadata, packed_size, unpack, unpack_from_memoryviews

__all__ = ["pack", "unpack", "packed_size", "pack_metadata", "unpack_from_memoryviews", "PackedColumns"]
