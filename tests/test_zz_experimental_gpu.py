"""Opt-in experimental paths that were written after the round-1 GPU budget was spent. They are OFF by default; these
checks run them in a subprocess with their environment switch and are marked xfail(strict=False) so that the first
hardware run reports their state (XPASS / XFAIL) without gating the validated suite."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.xfail(strict=False, reason="payload-carrying sort_by_key (B2_SORT_CARRY=1) not yet validated on hardware")
def test_sort_by_key_carry_payload():
    code = r"""
import numpy as np, sys
sys.path.insert(0, '.')
import cudf_b200.pylibcudf as plc
from oracle import sort as osort
rng = np.random.default_rng(5)
for n in (1, 33, 6144, 6145, 200_003):
    for kdt in (np.int64, np.int32, np.uint16, np.float64):
        for vdt in (np.int64, np.float64, np.int32, np.float32):
            keys = (rng.standard_normal(n) * 50).astype(kdt)
            vals = rng.integers(0, 1 << 30, n).astype(vdt)
            for order in ((0, 1) if np.dtype(kdt).kind != 'f' else (0,)):
                got = plc.sorting.sort_by_key(plc.Table([plc.Column.from_numpy(vals)]), plc.Table([plc.Column.from_numpy(keys)]), [order], [])
                exp = osort.sort_by_key([(vals, None)], [(keys, None)], [order])[0][0]
                assert np.array_equal(got.columns()[0].to_numpy()[0], exp), (n, kdt, vdt, order)
print('CARRY_OK')
"""
    env = dict(os.environ, B2_SORT_CARRY="1")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)
    assert "CARRY_OK" in r.stdout, r.stdout[-1500:] + r.stderr[-2500:]


@pytest.mark.xfail(strict=False, reason="aliased sort_by_key(T, T) -> keys-only radix (B2_SORT_ALIAS=1) not yet run on hardware")
def test_sort_by_key_alias_shortcut():
    code = r"""
import numpy as np, sys
sys.path.insert(0, '.')
import cudf_b200.pylibcudf as plc
rng = np.random.default_rng(11)
for n in (1, 33, 6145, 200_003, 3_000_001):
    for dt in (np.int64, np.int32, np.uint16, np.int8, np.uint64):
        keys = rng.integers(np.iinfo(dt).min, np.iinfo(dt).max, n, dtype=dt, endpoint=True)
        for order in (0, 1):
            c = plc.Column.from_numpy(keys)
            got = plc.sorting.sort_by_key(plc.Table([c]), plc.Table([c]), [order], []).columns()[0].to_numpy()[0]
            exp = np.sort(keys, kind='stable')
            if order == 1:
                exp = exp[::-1]
            assert np.array_equal(got, exp), (n, dt, order)
print('ALIAS_OK')
"""
    env = dict(os.environ, B2_SORT_ALIAS="1")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)
    assert "ALIAS_OK" in r.stdout, r.stdout[-1500:] + r.stderr[-2500:]
