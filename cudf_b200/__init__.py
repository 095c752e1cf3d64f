"""cudf_b200 — B200-native (sm_100a) sort / hash join / hash groupby / scan / reduce hot path.

The product is the CUDA library ``libcudf_b200.so`` behind the C ABI in ``include/cudf_b200.h``.
``cudf_b200.pylibcudf`` mirrors the pylibcudf API names of the reference for this path
(python/pylibcudf/pylibcudf/{sorting,join,groupby,reduce,copying,aggregation}.pyx).
There is no CPU fallback: importing the bindings without the built library raises.
"""
from . import _lib  # noqa: F401  (fails loudly when the extension is missing)
from . import pylibcudf  # noqa: F401

__all__ = ["pylibcudf"]
