"""CPU model of the radix-sort ping-pong plan (plan_kernel in cudf_b200/csrc/radix_sort.cu): exhaustive check of the
buffer-role invariants for every combination of skipped passes.  The device code is a line-by-line twin of `plan`."""
import itertools

import pytest


def plan(npass, trivial_mask, raw, pre_idx_buf=1):
    triv = [(trivial_mask >> p) & 1 for p in range(npass)]
    nexec = sum(1 for t in triv if not t)
    k = 0
    key_cur = 0 if raw else 1
    idx_cur = -1 if raw else pre_idx_buf
    out = []
    for p in range(npass):
        if triv[p]:
            out.append(None)
            continue
        remaining_after = nexec - 1 - k
        key_src, idx_src = key_cur, idx_cur
        key_dst = ((1 if remaining_after % 2 == 0 else 2) if raw else (2 if key_cur == 1 else 1))
        idx_dst = 0 if remaining_after % 2 == 0 else (2 if idx_cur == 1 else 1)
        last = remaining_after == 0
        key_cur, idx_cur = key_dst, idx_dst
        k += 1
        out.append(dict(key_src=key_src, key_dst=key_dst, idx_src=idx_src, idx_dst=idx_dst, last=last))
    return out, nexec


@pytest.mark.parametrize("npass", [1, 2, 4, 8])
@pytest.mark.parametrize("raw", [True, False])
def test_plan_invariants(npass, raw):
    for mask in range(1 << npass):
        steps, nexec = plan(npass, mask, raw)
        ex = [s for s in steps if s]
        assert len(ex) == nexec
        if not ex:
            continue
        for i, s in enumerate(ex):
            assert s["key_src"] != s["key_dst"], (npass, mask, raw, s)       # never scatter in place
            assert s["idx_src"] != s["idx_dst"], (npass, mask, raw, s)
            assert s["key_dst"] in (1, 2) and s["idx_dst"] in (0, 1, 2)
            if i > 0:
                assert s["key_src"] == ex[i - 1]["key_dst"] and s["idx_src"] == ex[i - 1]["idx_dst"]
            assert s["last"] == (i == len(ex) - 1)
        assert ex[-1]["idx_dst"] == 0                                          # row ids end in the output buffer
        if raw:
            assert ex[0]["key_src"] == 0 and ex[0]["idx_src"] == -1
            assert ex[-1]["key_dst"] == 1                                      # keys-only output / kept keys end in buffer 1
            assert all(s["idx_dst"] in (0, 1) for s in ex)                     # the third row-id buffer is never needed
        else:
            assert ex[0]["key_src"] == 1 and ex[0]["idx_src"] == 1
            # only the first executed pass may need the third row-id buffer
            assert all(s["idx_dst"] != 2 for s in ex[1:])
