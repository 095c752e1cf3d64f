"""Partitioned hash groupby (groupby.cu::pgb_agg_kernel) on the GPU: small inputs forced onto the path, a tiny shared table
that spills, and one input large enough (2^24 + rows) to take the path at the default settings with 1e6 groups."""
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
from tests.helpers import assert_columns_equal
from tests.impls import OracleImpl, PlcImpl, sort_groups
cu, o = PlcImpl(plc), OracleImpl()
"""


@pytest.mark.parametrize("smem_slots,est", [("0", "0"), ("64", "0"), ("0", "1"), ("0", "cap")])
def test_partitioned_groupby_forced(smem_slots, est):
    """est: the histogram-free partition pass (estimated bases, B2_GROUPBY_EST=1); "cap" forces a tiny capacity = its overflow fallback."""
    from tests.snippets.partitioned_groupby import CODE

    env = dict(os.environ, B2_GROUPBY_PARTITION_ROWS="1", B2_GROUPBY_SMEM_SLOTS=smem_slots, B2_GROUPBY_EST="0" if est == "0" else "1",
               B2_GROUPBY_EST_MIN="1")
    if est == "cap":
        env["B2_GROUPBY_EST_CAP"] = "40"
    cases = "CASES = [(1, 1), (100, 7), (5000, 300), (40_000, 20_000), (60_000, 3), (3_000_000, 1_500_000)]\n"
    r = subprocess.run([sys.executable, "-c", PRELUDE + cases + CODE], capture_output=True, text=True, env=env, cwd=ROOT, timeout=900)
    assert "PGB_OK" in r.stdout, r.stdout[-1500:] + r.stderr[-2500:]


def test_partitioned_groupby_default_settings(plc):
    """BASELINE configs[3] shape at 2^24 + 5 rows: int64 key with 1e6 groups, sum(float64) + count(int32); per-group equality
    against np.bincount (counts and integer keys bit-exact, float sums within 1e-9 relative)."""
    rng = np.random.default_rng(9)
    n, G = (1 << 24) + 5, 1_000_000
    k = rng.integers(0, G, n).astype(np.int64)
    v = rng.random(n)
    c = rng.integers(0, 100, n).astype(np.int32)
    gb = plc.groupby.GroupBy(plc.Table([plc.Column.from_numpy(k)]))
    agg = plc.aggregation
    keys_out, res = gb.aggregate([plc.groupby.GroupByRequest(plc.Column.from_numpy(v), [agg.sum()]),
                                  plc.groupby.GroupByRequest(plc.Column.from_numpy(c), [agg.count()])])
    gk = keys_out.columns()[0].to_numpy()[0]
    gs = res[0].columns()[0].to_numpy()[0]
    gc = res[1].columns()[0].to_numpy()[0]
    order = np.argsort(gk, kind="stable")
    exp_c = np.bincount(k, minlength=G)
    present = np.nonzero(exp_c)[0]
    assert np.array_equal(gk[order], present)
    assert gc.dtype == np.int32 and np.array_equal(gc[order], exp_c[present])
    np.testing.assert_allclose(gs[order], np.bincount(k, weights=v, minlength=G)[present], rtol=1e-9)
