"""Property tests of the PLY codec (dimo_b200/ply.py): any table of scalar columns survives write -> read, in binary
(both byte orders, hand-assembled) and ascii."""
import io

import numpy as np
from hypothesis import given, settings, strategies as st

from dimo_b200 import ply

CODES = ["i1", "u1", "i2", "u2", "i4", "u4", "f4", "f8"]
NAMES = {"i1": "char", "u1": "uchar", "i2": "short", "u2": "ushort", "i4": "int", "u4": "uint", "f4": "float", "f8": "double"}


@st.composite
def tables(draw):
    n_cols = draw(st.integers(1, 6))
    n_rows = draw(st.integers(0, 17))
    codes = [draw(st.sampled_from(CODES)) for _ in range(n_cols)]
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    arr = np.empty(n_rows, dtype=[(f"p{i}", c) for i, c in enumerate(codes)])
    for i, c in enumerate(codes):
        if c[0] == "f":
            arr[f"p{i}"] = rng.normal(size=n_rows).astype(c)
        else:
            info = np.iinfo(c)
            arr[f"p{i}"] = rng.integers(info.min, info.max, size=n_rows, endpoint=True).astype(c)
    return arr


@settings(max_examples=60, deadline=None, derandomize=True, database=None)
@given(tables())
def test_binary_little_endian_round_trip(arr):
    buf = io.BytesIO()
    ply.write_structured_ply(buf, arr)
    back = ply.read_ply(io.BytesIO(buf.getvalue())).first
    assert back.dtype.names == arr.dtype.names and back.shape == arr.shape
    for n in arr.dtype.names:
        assert np.array_equal(back[n], arr[n]), n


@settings(max_examples=40, deadline=None, derandomize=True, database=None)
@given(tables())
def test_big_endian_and_ascii_files_read_back(arr):
    header = ["ply", "format {} 1.0", "comment hypothesis", f"element vertex {arr.shape[0]}"]
    header += [f"property {NAMES[arr.dtype[n].str.lstrip('<>|=')]} {n}" for n in arr.dtype.names]
    header.append("end_header")
    be = np.empty(arr.shape[0], dtype=[(n, ">" + arr.dtype[n].str.lstrip("<>|=")) for n in arr.dtype.names])
    for n in arr.dtype.names:
        be[n] = arr[n]
    blob = ("\n".join(header).format("binary_big_endian") + "\n").encode() + be.tobytes()
    back = ply.read_ply(io.BytesIO(blob)).first
    for n in arr.dtype.names:
        assert np.array_equal(back[n], arr[n]), n
    rows = "\n".join(" ".join(repr(v.item()) for v in row) for row in arr) + ("\n" if arr.shape[0] else "")
    txt = ("\n".join(header).format("ascii") + "\n" + rows).encode()
    back = ply.read_ply(io.BytesIO(txt)).first
    for n in arr.dtype.names:
        assert np.array_equal(back[n], arr[n]), n
