"""CPU: the C-ABI library loads and exports every symbol include/cudf_b200.h declares (no compute calls)."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _declared():
    text = (ROOT / "include" / "cudf_b200.h").read_text()
    return sorted(set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    import __graft_entry__ as g

    g.build()
    lib = ctypes.CDLL(str(ROOT / "cudf_b200" / "libcudf_b200.so"))
    names = _declared()
    assert len(names) >= 50
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_binding_list_matches_header():
    from cudf_b200 import _lib

    assert sorted(_lib.DECLARED_SYMBOLS) == _declared()
    assert _lib.MISSING == []


def test_error_codes_without_gpu():
    """Argument validation that fails before any CUDA call maps to the reference's exception classes."""
    import ctypes as C

    import pytest

    from cudf_b200 import _lib

    out = C.c_void_p()
    with pytest.raises(ValueError):  # std::invalid_argument
        _lib.check(_lib.lib.b2_sorted_order(None, None, 0, None, 0, 0, None, C.byref(out)))
    assert _lib.lib.b2_bitmask_allocation_size_bytes(1) == 64 and _lib.lib.b2_bitmask_allocation_size_bytes(513) == 128
    assert b"cudf_b200" in _lib.lib.b2_version()


def test_no_oracle_import_in_product():
    """The product path must not route through the oracle (or any CPU fallback)."""
    for p in (ROOT / "cudf_b200").rglob("*.py"):
        txt = p.read_text()
        assert "import oracle" not in txt and "from oracle" not in txt, p
