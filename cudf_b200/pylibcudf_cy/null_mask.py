"""pylibcudf.null_mask (python/pylibcudf/pylibcudf/null_mask.pyx): compiled in _core.pyx; DeviceBuffer is the ctypes twin's class."""
from ..pylibcudf.null_mask import DeviceBuffer
from ..pylibcudf.types import MaskState
from ._core import bitmask_allocation_size_bytes, bitmask_and, copy_bitmask, count_set_bits, create_null_mask, null_count, set_null_mask

__all__ = ["bitmask_allocation_size_bytes", "create_null_mask", "copy_bitmask", "bitmask_and", "null_count", "count_set_bits",
           "set_null_mask", "DeviceBuffer", "MaskState"]
