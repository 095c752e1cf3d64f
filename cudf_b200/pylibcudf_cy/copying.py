"""pylibcudf.copying.gather (python/pylibcudf/pylibcudf/copying.pyx:64-113): compiled in _core.pyx."""
from ._core import gather

__all__ = ["gather"]
