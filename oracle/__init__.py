"""CPU oracle — TEST INFRASTRUCTURE ONLY.

A numpy restatement of the reference's *semantics* for the hot path (the arithmetic itself lives in
CCCL / cuCollections, which are not in the reference tree and cannot be built here — see DESIGN.md).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this
package; the product (cudf_b200/) never does.

Pinned against the reference's own known-answer tests, transcribed under tests/golden/ with
file:line citations (tests/test_oracle_golden.py).  Hash-table layout / hash values are NOT pinned
(unobservable: join and groupby outputs are order-free and compared after canonical sorting).

A column is the pair (values: np.ndarray, valid: np.ndarray[bool] | None).
"""
from . import bitmask, datagen, groupby, join, reduce, sort  # noqa: F401
