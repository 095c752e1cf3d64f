"""Hybrid sort (partial LSD + segment fix-up) on the GPU at sizes where the default size limit would not select it,
plus one run at the default settings large enough to take it (2^22 rows)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

PRELUDE = r"""
import sys
sys.path.insert(0, '.')
import numpy as np
import cudf_b200.pylibcudf as plc
from cudf_b200 import _lib as L
from oracle import sort as osort
"""


@pytest.mark.parametrize("carry,fix_fast", [("1", "0"), ("1", "1"), ("0", "0"), ("0", "1")])
def test_hybrid_small_inputs(carry, fix_fast):
    """fix_fast: both flavours of the segment fix-up (the plan picks one from its expected segment length; B2_SORT_FIX_FAST forces it)."""
    from tests.snippets.hybrid_sort import CODE

    env = dict(os.environ, B2_SORT_HYBRID_MIN="0", B2_SORT_CARRY=carry, B2_SORT_FIX_FAST=fix_fast)
    r = subprocess.run([sys.executable, "-c", PRELUDE + "SIZES = (3, 100, 2047, 2049, 6145, 20011, 300_007)\n" + CODE], capture_output=True,
                       text=True, env=env, cwd=ROOT, timeout=600)
    assert "HYBRID_OK" in r.stdout, r.stdout[-1500:] + r.stderr[-2500:]


def test_hybrid_default_settings(plc):
    from oracle import sort as osort

    rng = np.random.default_rng(21)
    n = (1 << 22) + 13
    keys = rng.integers(-2**63, 2**63 - 1, n, dtype=np.int64)
    vals = rng.integers(0, 1 << 50, n).astype(np.int64)
    kc = plc.Column.from_numpy(keys)
    assert np.array_equal(plc.sorting.sorted_order(plc.Table([kc]), [], []).to_numpy()[0], osort.sorted_order([(keys, None)]))
    got = plc.sorting.sort_by_key(plc.Table([plc.Column.from_numpy(vals)]), plc.Table([kc]), [1], []).columns()[0].to_numpy()[0]
    assert np.array_equal(got, osort.sort_by_key([(vals, None)], [(keys, None)], [1])[0][0])
    assert np.array_equal(plc.sorting.sort(plc.Table([kc]), [0], []).columns()[0].to_numpy()[0], np.sort(keys))
