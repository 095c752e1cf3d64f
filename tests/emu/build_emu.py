"""Build the CPU emulation of the kernel library: cudf_b200/csrc/*.cu compiled as C++ against tests/emu/include.

TEST INFRASTRUCTURE ONLY (see include/cuda_runtime.h). Output: tests/emu/_build/libcudf_b200_emu.so (git-ignored).
"""
from __future__ import annotations

import hashlib
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
CSRC = ROOT / "cudf_b200" / "csrc"
OUT = HERE / "_build"
LIB = OUT / "libcudf_b200_emu.so"

FLAGS = ["-std=c++17", "-O1", "-g", "-fPIC", "-DB2_EMU", "-fno-strict-aliasing", "-Wno-unknown-pragmas", "-Wno-attributes",
         f"-I{HERE / 'include'}", f"-I{CSRC}"]


def _digest(paths) -> str:
    h = hashlib.sha256(" ".join(FLAGS).encode())
    for p in paths:
        h.update(p.read_bytes())
    return h.hexdigest()


def build(verbose: bool = False) -> Path:
    OUT.mkdir(exist_ok=True)
    sources = sorted(CSRC.glob("*.cu")) + [HERE / "emu_runtime.cpp"]
    deps = sources + sorted(CSRC.glob("*.cuh")) + [ROOT / "include" / "cudf_b200.h", HERE / "include" / "cuda_runtime.h"]
    stamp = OUT / "stamp"
    want = _digest(deps)
    if LIB.exists() and stamp.exists() and stamp.read_text() == want:
        return LIB

    def compile_one(src: Path) -> Path:
        obj = OUT / (src.stem + ".o")
        cmd = ["g++", "-x", "c++", "-c", str(src), "-o", str(obj)] + FLAGS
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"emu build failed for {src.name}:\n{r.stderr[-6000:]}")
        if verbose and r.stderr:
            print(r.stderr[-2000:], file=sys.stderr)
        return obj

    with ThreadPoolExecutor(8) as ex:
        objs = list(ex.map(compile_one, sources))
    # -Bsymbolic: the real library may be loaded RTLD_GLOBAL in the same process; keep our internal calls ours
    r = subprocess.run(["g++", "-shared", "-o", str(LIB)] + [str(o) for o in objs] + ["-Wl,-Bsymbolic", "-lpthread"], capture_output=True,
                       text=True)
    if r.returncode != 0:
        raise RuntimeError("emu link failed:\n" + r.stderr[-4000:])
    stamp.write_text(want)
    return LIB


if __name__ == "__main__":
    print(build(verbose=True))
