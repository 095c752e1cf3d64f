"""The compiled (Cython) binding `cudf_b200.pylibcudf_cy` on CPU: it builds, links against libcudf_b200.so by a relocatable
rpath, its .pxd declares only functions the header declares, and the same _core.pyx linked against the kernel emulator's library
reproduces the oracle for every operation it binds (the binding is host code: what the emulator run cannot show is the GPU)."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_extension_builds_and_links():
    import __graft_entry__ as g

    g.build()
    import cudf_b200.pylibcudf_cy as cy

    assert cy._core.version().startswith("cudf_b200")
    so = cy._core.__file__
    dyn = subprocess.run(["readelf", "-d", so], capture_output=True, text=True).stdout
    assert "libcudf_b200.so" in dyn and "$ORIGIN/.." in dyn, dyn
    for mod, names in ((cy.sorting, ["sorted_order", "stable_sorted_order", "sort", "stable_sort", "sort_by_key", "stable_sort_by_key"]),
                       (cy.join, ["inner_join", "left_join", "full_join", "HashJoin"]), (cy.groupby, ["GroupBy", "GroupByRequest"]),
                       (cy.reduce, ["reduce", "scan", "segmented_reduce", "ScanType"]), (cy.copying, ["gather"]),
                       (cy.partitioning, ["hash_partition", "partition", "HashId"]),
                       (cy.contiguous_split, ["pack", "unpack", "packed_size", "pack_metadata", "PackedColumns"]),
                       (cy.null_mask, ["create_null_mask", "copy_bitmask", "bitmask_and", "null_count", "count_set_bits", "set_null_mask"])):
        for n in names:
            assert hasattr(mod, n), (mod.__name__, n)
    assert type(cy.Column).__module__ != "ctypes" and cy.Column.__module__.endswith("_core")


def test_pxd_matches_header():
    """Every function the .pxd declares is declared by include/cudf_b200.h (the C compiler checks the prototypes themselves when
    the extension is built: the generated C includes the header)."""
    pxd = open(os.path.join(ROOT, "cudf_b200", "pylibcudf_cy", "libcudf_b200.pxd")).read()
    hdr = open(os.path.join(ROOT, "include", "cudf_b200.h")).read()
    declared = set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", pxd))
    in_header = set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", hdr))
    assert declared and declared <= in_header, sorted(declared - in_header)


CODE = r"""
import sys
sys.path.insert(0, '.')
from tests.emu import harness
cy = harness.install_cy()
import numpy as np
from oracle import sort as osort
from tests.impls import OracleImpl, PlcImpl, sort_groups
from tests.helpers import assert_columns_equal
before = cy._core.kernel_launch_count()
cu, o = PlcImpl(cy), OracleImpl()
rng = np.random.default_rng(3)
for n in (0, 1, 5000):
    keys = rng.integers(-50, 50, n).astype(np.int64); vals = rng.random(n)
    kv = rng.random(n) < 0.9
    for order in (0, 1):
        got = cy.sorting.sort_by_key(cy.Table([cy.Column.from_numpy(vals)]), cy.Table([cy.Column.from_numpy(keys, kv)]), [order], [0])
        exp = osort.sort_by_key([(vals, None)], [(keys, kv)], [order], [0])[0][0]
        assert np.array_equal(got.columns()[0].to_numpy()[0], exp), (n, order)
        for fn in ("sorted_order", "stable_sorted_order"):
            so = getattr(cy.sorting, fn)(cy.Table([cy.Column.from_numpy(keys, kv)]), [order], [1])
            assert np.array_equal(so.to_numpy()[0], osort.sorted_order([(keys, kv)], [order], [1])), (n, order, fn)
keys = rng.integers(-50, 50, 5000).astype(np.int64); vals = rng.random(5000)
st = cy.sorting.stable_sort(cy.Table([cy.Column.from_numpy(keys), cy.Column.from_numpy(vals)]), [0, 1], [])
ek = osort.sorted_order([(keys, None), (vals, None)], [0, 1])
assert np.array_equal(st.columns()[0].to_numpy()[0], keys[ek]) and np.array_equal(st.columns()[1].to_numpy()[0], vals[ek])
l = [(rng.integers(0, 300, 4000), rng.random(4000) < 0.95)]; r = [(rng.integers(0, 300, 1000), rng.random(1000) < 0.95)]
for ne in (0, 1):
    for kind in ("inner_join", "left_join", "full_join"):
        g, e = getattr(cu, kind)(l, r, ne), getattr(o, kind)(l, r, ne)
        assert np.array_equal(g[0], e[0]) and np.array_equal(g[1], e[1]), (kind, ne)
assert cu.inner_join_size(l, r) == o.inner_join_size(l, r)
hj = cy.join.HashJoin(cu._tbl(r), 0)
a = hj.left_join(cu._tbl(l)); b = cy.join.left_join(cu._tbl(l), cu._tbl(r), 0)
assert a[0].size() == b[0].size() == hj.left_join_size(cu._tbl(l))
k = [(rng.integers(0, 40, 3000).astype(np.int32), rng.random(3000) < 0.9)]
v = (rng.integers(-100, 100, 3000).astype(np.int64), rng.random(3000) < 0.8)
kinds = ["sum", "min", "max", "count", "count_all", "mean"]
for inc in (False, True):
    gk, gr = sort_groups(*cu.groupby(k, [(v, kinds)], inc)); ek, er = sort_groups(*o.groupby(k, [(v, kinds)], inc))
    assert_columns_equal(gk[0], ek[0], what="keys")
    for j, kind in enumerate(kinds):
        assert_columns_equal(gr[0][j], er[0][j], what=kind)
gk, gr = cu.groupby_scan(k, [(v, ["sum", "count"])]); ek, er = o.groupby_scan(k, [(v, ["sum", "count"])])
assert_columns_equal(gk[0], ek[0], what="scan keys")
for j in range(2):
    assert_columns_equal(gr[0][j], er[0][j], what="groupby scan")
x = (rng.integers(-100, 100, 7000).astype(np.int64), rng.random(7000) < 0.9)
assert_columns_equal(cu.scan(x, "sum"), o.scan(x, "sum"), what="scan")
assert_columns_equal(cu.scan(x, "max", inclusive=False), o.scan(x, "max", inclusive=False), what="exclusive scan")
assert cu.reduce(x, "sum", np.int64) == o.reduce(x, "sum", np.int64)
assert cu.reduce(x, "min", np.int64, init=(-1000, True)) == o.reduce(x, "min", np.int64, init=(-1000, True))
offs = np.sort(rng.integers(0, 7000, 30)).astype(np.int32); offs[0] = 0
assert_columns_equal(cu.segmented_reduce(x, offs, "sum", np.int64), o.segmented_reduce(x, offs, "sum", np.int64), what="segmented")
src = cy.Table([cy.Column.from_numpy(vals)])
gm = cy.Column.from_numpy(np.array([4, 0, 4999, 7], dtype=np.int32))
assert np.array_equal(cy.copying.gather(src, gm, cy.OutOfBoundsPolicy.DONT_CHECK).columns()[0].to_numpy()[0], vals[[4, 0, 4999, 7]])
sl = cy.Column.from_numpy(keys).slice(10, 60)
assert np.array_equal(cy.sorting.sort(cy.Table([sl]), [0], []).columns()[0].to_numpy()[0], np.sort(keys[10:60]))
# error classes of python/pylibcudf/pylibcudf/exception_handler.pxd:29-66
for fn, exc in ((lambda: cy.sorting.sort_by_key(cy.Table([cy.Column.from_numpy(vals[:5])]), cy.Table([cy.Column.from_numpy(keys)]), [0], []), RuntimeError),
                (lambda: cy.join.HashJoin(cu._tbl(r), 0, None, 1.5), ValueError),
                (lambda: cy.Table([cy.Column.from_numpy(vals[:5]), cy.Column.from_numpy(keys)]), ValueError)):
    try:
        fn()
        raise SystemExit("no error")
    except exc:
        pass
# ---- the other modules of the path: segmented sort / top-k / rank, match contexts, partitioning, null masks ----
from oracle import partition as opart
kcol = (rng.integers(0, 50, 3000).astype(np.int32), rng.random(3000) < 0.9)
soffs = np.array([0, 10, 10, 700, 2999], dtype=np.int32)
for order in (0, 1):
    got = cy.sorting.stable_segmented_sorted_order(cy.Table([cy.Column.from_numpy(*kcol)]), cy.Column.from_numpy(soffs), [order], [order]).to_numpy()[0]
    assert np.array_equal(got, osort.segmented_sorted_order([kcol], soffs, [order], [order])), order
    tk = cy.sorting.top_k_order(cy.Column.from_numpy(*kcol), 17, order).to_numpy()[0]
    assert np.array_equal(tk, osort.top_k_order(kcol, 17, order)), order
sv = cy.sorting.segmented_sort_by_key(cy.Table([cy.Column.from_numpy(np.arange(3000, dtype=np.int64))]), cy.Table([cy.Column.from_numpy(*kcol)]),
                                      cy.Column.from_numpy(soffs), [0], [1]).columns()[0].to_numpy()[0]
assert sorted(sv.tolist()) == list(range(3000))
tv = cy.sorting.top_k(cy.Column.from_numpy(kcol[0]), 5, 1).to_numpy()[0]
assert sorted(tv.tolist()) == sorted(np.sort(kcol[0])[-5:].tolist())
for method in range(5):
    g = cy.sorting.rank(cy.Column.from_numpy(*kcol), method, 0, 0, 1, False).to_numpy()
    assert_columns_equal(g, osort.rank(kcol, method, 0, 0, 1, False), what=f"rank {method}")
for kind in ("inner", "left", "full"):
    assert np.array_equal(cu.match_counts(l, r, 0, kind), o.match_counts(l, r, 0, kind)), kind
    g, e = cu.partitioned_join(l, r, 0, kind, (0, 1000, 2500)), o.partitioned_join(l, r, 0, kind)
    assert np.array_equal(g[0], e[0]) and np.array_equal(g[1], e[1]), kind
tcols = [(rng.integers(-5, 5, 2000).astype(np.int64), rng.random(2000) < 0.9), (rng.random(2000).astype(np.float32), None)]
for P in (1, 7, 300):
    out, offs = cy.partitioning.hash_partition(cy.Table([cy.Column.from_numpy(*c) for c in tcols]), [0, 1], P)
    ecols, eoffs = opart.hash_partition(tcols, tcols, P, 0, False)
    assert offs == list(eoffs), P
    for gc, ec in zip(out.columns(), ecols):
        assert_columns_equal(gc.to_numpy(), ec, what=f"hash_partition {P}")
pmap = rng.integers(0, 9, 2000).astype(np.int32)
pout, poffs = cy.partitioning.partition(cy.Table([cy.Column.from_numpy(np.arange(2000, dtype=np.int64))]), cy.Column.from_numpy(pmap), 9)
assert np.array_equal(pout.columns()[0].to_numpy()[0], np.argsort(pmap, kind="stable")) and poffs == np.concatenate([[0], np.cumsum(np.bincount(pmap, minlength=9))]).tolist()
try:
    cy.partitioning.hash_partition(cy.Table([cy.Column.from_numpy(pmap)]), [3], 4)
    raise SystemExit("no error")
except IndexError:
    pass
nm = cy.null_mask
assert nm.bitmask_allocation_size_bytes(100) == 64
mcol = cy.Column.from_numpy(*kcol)
mptr = mcol.null_mask().ptr
assert nm.null_count(mptr, 0, 3000) == int((~kcol[1]).sum()) and nm.count_set_bits(mptr, 5, 2000) == int(kcol[1][5:2000].sum())
buf, nc = nm.bitmask_and([mcol, cy.Column.from_numpy(kcol[0], ~kcol[1] | (rng.random(3000) < 0.5))])
assert nc >= int((~kcol[1]).sum()) and buf.size >= 3000 // 8
cp = nm.copy_bitmask(mcol.slice(3, 2000))
assert nm.count_set_bits(cp.ptr, 0, 1997) == int(kcol[1][3:2000].sum())
all_valid = nm.create_null_mask(200, nm.MaskState.ALL_VALID)
nm.set_null_mask(all_valid.ptr, 10, 20, False)
assert nm.null_count(all_valid.ptr, 0, 200) == 10
# pack / unpack round trip and metadata bytes against the oracle's wire format
from oracle import pack as opack
pcols = [(rng.integers(-9, 9, 777).astype(np.int64), rng.random(777) < 0.8), (rng.random(777).astype(np.float32), None)]
pt = cy.Table([cy.Column.from_numpy(*c) for c in pcols])
packed = cy.contiguous_split.pack(pt)
assert cy.contiguous_split.packed_size(pt) == packed.gpu_data_size == opack.packed_size(pcols)
md, _data = opack.pack(pcols, [int(c.type().id()) for c in pt.columns()])
assert packed.metadata == md
up = cy.contiguous_split.unpack(packed)
for gc, ec in zip(up.columns(), pcols):
    assert_columns_equal(gc.to_numpy(), ec, what="unpack")
assert cy.contiguous_split.pack_metadata(up, packed.gpu_data_ptr, packed.gpu_data_size) == md
# ctypes twin <-> compiled twin share memory
import cudf_b200.pylibcudf as plc
pc = plc.Column.from_numpy(keys)
assert np.array_equal(cy.Column.from_plc(pc).to_plc().to_numpy()[0], keys)
assert cy._core.kernel_launch_count() > before
print('CY_EMU_OK')
"""


def test_cython_binding_on_the_emulator():
    r = subprocess.run([sys.executable, "-c", CODE], capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert "CY_EMU_OK" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]
