"""Builds the Cython extension `_core` in-tree: cython -> C -> gcc, linked against libcudf_b200.so next to the package
(rpath $ORIGIN/..). `lib` / `out_dir` let tests/emu build the same sources against the kernel emulator's library."""
from __future__ import annotations

import hashlib
import subprocess
import sys
import sysconfig
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
EXT = sysconfig.get_config_var("EXT_SUFFIX")


def build(lib: Path | None = None, out_dir: Path | None = None, force: bool = False) -> Path:
    lib = Path(lib) if lib else HERE.parent / "libcudf_b200.so"
    out_dir = Path(out_dir) if out_dir else HERE
    out_dir.mkdir(parents=True, exist_ok=True)
    target = out_dir / f"_core{EXT}"
    srcs = [HERE / "_core.pyx", HERE / "libcudf_b200.pxd", ROOT / "include" / "cudf_b200.h"]
    h = hashlib.sha256()
    for p in srcs:
        h.update(p.read_bytes())
    default = lib.parent == HERE.parent and out_dir == HERE
    h.update(b"in-tree" if default else str(lib.resolve()).encode())  # the in-tree build is relocatable (rpath $ORIGIN/..)
    stamp = out_dir / "_core.stamp"
    if not force and target.exists() and stamp.exists() and stamp.read_text() == h.hexdigest():
        return target
    if not lib.exists():
        raise FileNotFoundError(f"{lib} is missing: build the CUDA library first")
    work = out_dir / "_build"
    work.mkdir(exist_ok=True)
    c_file = work / "_core.c"
    subprocess.run([sys.executable, "-m", "cython", "-3", "--module-name", "cudf_b200.pylibcudf_cy._core", "-I", str(ROOT), str(HERE / "_core.pyx"),
                    "-o", str(c_file)], check=True, cwd=ROOT)
    inc = sysconfig.get_paths()["include"]
    rpath = "$ORIGIN/.." if default else str(lib.parent.resolve())
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-fvisibility=hidden", "-Wno-unused-function", "-Wno-unreachable-code", f"-I{inc}", f"-I{ROOT / 'include'}", str(c_file), "-o", str(target),
           f"-L{lib.parent}", f"-l:{lib.name}", f"-Wl,-rpath,{rpath}"]
    subprocess.run(cmd, check=True, cwd=ROOT)
    stamp.write_text(h.hexdigest())
    return target


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
