"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports
every symbol include/mzgpu.h declares, the ctypes table covers them all, and
without a CUDA device the product path fails loudly (no CPU fallback)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "mzgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mzgpu_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    from materialize_b200 import _ffi

    names = declared_functions()
    assert len(names) >= 55
    for name in names:
        assert hasattr(_ffi.lib, name), f"libmzgpu.so does not export {name}"
    assert set(names) == set(_ffi.SIGNATURES), set(names) ^ set(_ffi.SIGNATURES)


def test_rust_shim_declares_every_entry_point():
    """rust-shim/src/sys.rs (uncompiled: no Rust toolchain here) stays one to one with the header:
    the same functions, each with the same number of arguments."""
    src = open(os.path.join(ROOT, "include", "mzgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    c_args = {}
    for name, args in re.findall(r"\b(mzgpu_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src):
        args = " ".join(args.split())
        c_args[name] = 0 if args in ("", "void") else len(args.split(","))
    rs = open(os.path.join(ROOT, "rust-shim", "src", "sys.rs")).read()
    rs = re.sub(r"//[^\n]*", "", rs)
    r_args = {}
    for name, args in re.findall(r"pub fn (mzgpu_[a-z0-9_]+)\s*\(([^)]*)\)", rs):
        args = " ".join(args.split())
        r_args[name] = 0 if args == "" else len([a for a in args.split(",") if a.strip()])
    assert set(c_args) == set(declared_functions())
    assert set(r_args) == set(c_args), set(r_args) ^ set(c_args)
    assert {n: r_args[n] for n in c_args} == c_args


def test_row_layouts_match_header():
    from materialize_b200 import _ffi

    assert _ffi.R16.itemsize == 16 and _ffi.R32.itemsize == 32 and _ffi.R40.itemsize == 40
    assert _ffi.RACC.itemsize == 80 and _ffi.ROUT.itemsize == 64
    import ctypes as C

    assert C.sizeof(_ffi.Closure) == 144
    assert C.sizeof(_ffi.Desc) == 24 and C.sizeof(_ffi.Stats) == 64


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import materialize_b200 as mz

    with pytest.raises(mz.MzGpuError) as e:
        mz.Context(0)
    assert e.value.status == -2  # MZGPU_E_CUDA


def test_route_matches_oracle(oracle):
    import materialize_b200 as mz

    for peers in (1, 2, 3, 8):
        for key in (0, 1, 2, 12345678901234567, 2**64 - 1):
            assert mz.route(key, peers) == oracle.lib().mzo_route(key, peers)


def test_product_does_not_touch_oracle():
    """Nothing under materialize_b200/ may import, link or load oracle/."""
    pkg = os.path.join(ROOT, "materialize_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "libmzoracle" not in text, f
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f
                assert not re.search(r'#include\s+"[^"]*oracle/', text), f


def test_row_keys_pack_order_preserving_and_invertible():
    """f1, first step: Rows of at most 7 bytes as one u64 that orders exactly like RowRef::cmp
    (length first, then bytes: src/repr/src/row.rs:704-722) and maps back.  Host-only functions of
    the C ABI: no GPU needed."""
    import ctypes as C
    import random

    from materialize_b200 import _ffi as F

    rng = random.Random(5)
    rows = [bytes(rng.randrange(256) for _ in range(rng.randrange(0, 8))) for _ in range(4000)]
    rows += [b"", b"\x00", b"\x00\x00", b"\xff", b"\xff" * 7, b"\x01\x00", b"\x00\x01"]
    data = b"".join(rows)
    offs = [0]
    for r in rows:
        offs.append(offs[-1] + len(r))
    buf = (C.c_uint8 * max(1, len(data))).from_buffer_copy(data or b"\x00")
    offsets = (C.c_uint64 * len(offs))(*offs)
    keys = (C.c_uint64 * len(rows))()
    done = C.c_uint64(0)
    assert F.lib.mzgpu_rowkeys_pack(buf, offsets, len(rows), keys, C.byref(done)) == F.OK and done.value == len(rows)
    keys = list(keys)
    # same order as the reference's Row comparison
    by_row = sorted(range(len(rows)), key=lambda i: (len(rows[i]), rows[i]))
    by_key = sorted(range(len(rows)), key=lambda i: keys[i])
    assert [rows[i] for i in by_row] == [rows[i] for i in by_key]
    assert len(set(keys)) == len(set(rows))  # injective
    out, n = (C.c_uint8 * 7)(), C.c_uint64(0)
    for r, k in zip(rows[:500] + rows[-7:], keys[:500] + keys[-7:]):
        assert F.lib.mzgpu_rowkey_unpack(k, out, C.byref(n)) == F.OK and bytes(out[: n.value]) == r
    # longer rows are outside the subset, reported at the first one
    two = b"\x01\x02" + b"\x09" * 8
    buf2 = (C.c_uint8 * len(two)).from_buffer_copy(two)
    offs2 = (C.c_uint64 * 3)(0, 2, 10)
    keys2 = (C.c_uint64 * 2)()
    assert F.lib.mzgpu_rowkeys_pack(buf2, offs2, 2, keys2, C.byref(done)) == F.E_UNSUPPORTED and done.value == 1
    assert F.lib.mzgpu_rowkey_unpack((8 << 56), out, C.byref(n)) == F.E_INVALID


def test_column_size_arithmetic_matches_oracle(oracle):
    """f4: the pure host functions of the columnar wire format (indexed::length_in_words, the ship signal of
    columnar.rs:164-175 / builder.rs:48-52, rows per minted container) agree with the oracle restatement."""
    import numpy as np

    import materialize_b200._ffi as F

    rng = np.random.default_rng(1)
    for words in [0, 1, 235930, 235931, 262143, 262144, 262145, 471860, 471861, 524288] + rng.integers(0, 1 << 24, size=200).tolist():
        assert bool(F.lib.mzgpu_column_at_capacity(int(words))) == oracle.col_at_capacity(int(words))
    for n, kb, vb in [(0, 0, 0), (1, 0, 7), (1000, 3001, 12), (58982, 1, 1)]:
        assert F.lib.mzgpu_column_length_in_words(F.COLUMN_U64X4, n, kb, vb) == oracle.col_length_in_words([8 * n] * 4)
        assert F.lib.mzgpu_column_length_in_words(F.COLUMN_U64X2, n, kb, vb) == oracle.col_length_in_words([8 * n] * 2)
        assert F.lib.mzgpu_column_length_in_words(F.COLUMN_ROWROW, n, kb, vb) == oracle.col_length_in_words([8 * n, kb, 8 * n, vb, 8 * n, 8 * n])
    a = np.zeros(120_000, dtype=oracle.R32)
    for layout in (F.COLUMN_U64X4, F.COLUMN_U64X2):
        first = oracle.column_builder(layout, a)[0]
        assert F.lib.mzgpu_column_ship_rows(layout) == len(oracle.column_rows(layout, first))
    assert F.lib.mzgpu_column_ship_rows(F.COLUMN_ROWROW) == 0
