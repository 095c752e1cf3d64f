#!/usr/bin/env python
"""Which kernels of an object file changed?   python scripts/sass_identity.py old.o new.o

Compares the SASS of every kernel in `old.o` with `new.o` (cuobjdump -sass, encodings stripped). Kernels are matched by
BODY, so a kernel whose mangled name changed (e.g. a new defaulted template parameter) still counts as identical.
Used to show that an edit did not touch the hardware-validated kernels."""
import re
import subprocess
import sys


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    d, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            d[cur] = []
        elif cur and line.strip():
            d[cur].append(re.sub(r"/\* 0x[0-9a-f]+ \*/", "", line).rstrip())
    return {k: "\n".join(v) for k, v in d.items()}


def main():
    old, new = kernels(sys.argv[1]), kernels(sys.argv[2])
    new_bodies = set(new.values())
    changed = [k for k, body in old.items() if body not in new_bodies]
    added = [k for k, body in new.items() if body not in set(old.values())]
    print(f"{len(old)} kernels before, {len(new)} after; {len(old) - len(changed)} unchanged")
    for k in changed:
        print("  changed or removed:", subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()[:160])
    for k in added:
        print("  new or changed:    ", subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()[:160])
    return 1 if changed else 0


if __name__ == "__main__":
    sys.exit(main())
