"""CPU restatement of cudf::pack's wire format for tables of fixed-width columns — TEST INFRASTRUCTURE ONLY.

Follows cpp/src/copying/pack.cpp:36-85 (serialized_column: {data_type{int32 id, int32 scale}, int32 size, int32 null_count,
int64 data_offset, int64 null_mask_offset, int32 num_children, int32 pad}; serialized_table_header: {int32 version = 2
(cpp/include/cudf/detail/contiguous_split.hpp:127), int32 num_columns, int32 num_rows, int32 pad}), pack.cpp:297-330 (header
then the column entries, depth first) and cpp/src/copying/contiguous_split.cu:50,505-560,1004 (per column: the validity
buffer first when the column is nullable, then the data; every buffer padded to split_align = 64 bytes; a sliced
column's mask is re-based to bit 0). The reference's tests (cpp/tests/copying/pack_tests.cpp:20-66) pin the round trip,
packed_size == gpu_data size and metadata sizes, not the bytes: the byte layout here is pinned on the struct definitions."""
from __future__ import annotations

import struct

import numpy as np

VERSION = 2
ALIGN = 64


def _round_up(n: int) -> int:
    return (n + ALIGN - 1) // ALIGN * ALIGN


def pack(cols, type_ids):
    """cols: list of (values, valid-or-None); type_ids: libcudf type ids. -> (metadata bytes, gpu_data bytes)."""
    n = len(cols[0][0]) if cols else 0
    md = struct.pack("<iiii", VERSION, len(cols), n, 0)
    data = bytearray()
    for (values, valid), tid in zip(cols, type_ids):
        v = np.ascontiguousarray(values)
        size = len(v)
        doff = moff = -1
        nulls = 0
        if size:
            if valid is not None:
                valid = np.asarray(valid, bool)
                nulls = int((~valid).sum())
                bits = np.packbits(valid, bitorder="little").tobytes()
                words = bits + b"\0" * ((size + 31) // 32 * 4 - len(bits))
                moff = len(data)
                data += words + b"\0" * (_round_up(len(words)) - len(words))
            raw = (v.astype(np.uint8) if v.dtype == np.bool_ else v).tobytes()
            doff = len(data)
            data += raw + b"\0" * (_round_up(len(raw)) - len(raw))
        md += struct.pack("<iiiiqqii", int(tid), 0, size, nulls, doff, moff, 0, 0)
    return md, bytes(data)


def packed_size(cols) -> int:
    total = 0
    for values, valid in cols:
        size = len(values)
        if not size:
            continue
        if valid is not None:
            total += _round_up((size + 31) // 32 * 4)
        total += _round_up(size * np.asarray(values).dtype.itemsize)
    return total
