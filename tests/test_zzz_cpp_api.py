"""The cudf:: C++ header surface (include/cudf/*.hpp) compiles against the C ABI (CPU) and runs (GPU)."""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
EXE = ROOT / "tests" / "cpp" / "api_smoke"


def _build():
    import __graft_entry__ as g

    g.build()
    cmd = ["g++", "-std=c++17", f"-I{ROOT / 'include'}", "-I/usr/local/cuda/include", str(ROOT / "tests/cpp/api_smoke.cpp"), "-o", str(EXE),
           f"-L{ROOT / 'cudf_b200'}", "-lcudf_b200", "-L/usr/local/cuda/lib64", "-lcudart", f"-Wl,-rpath,{ROOT / 'cudf_b200'}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_cpp_headers_compile():
    _build()
    assert EXE.exists()


@pytest.mark.gpu
def test_cpp_api_runs():
    _build()
    r = subprocess.run([str(EXE)], capture_output=True, text=True, timeout=120)
    assert "CPP_API_OK" in r.stdout, r.stdout + r.stderr
