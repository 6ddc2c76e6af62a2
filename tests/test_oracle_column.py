"""Oracle restatement of the columnar wire format: `columnar::bytes::indexed` as used by
`Column<C>` (src/timely-util/src/columnar.rs:54-222) and `ColumnBuilder`
(src/timely-util/src/columnar/builder.rs:28-111).  The `columnar` crate is not vendored: the restatement is
pinned to the bytes the reference's own tests hold (`raw_columnar_bytes`, columnar.rs:247-258, and the three
tests around it, :260-339), then checked for round trips and for the builder's ship rule."""
import struct

import numpy as np
import pytest

# columnar.rs:247-258: offsets (16), length (28), 1i32, 2i32, 3i32, four bytes of padding
RAW_COLUMNAR_BYTES = struct.pack("<QQiii4x", 16, 28, 1, 2, 3)


def test_column_known_bytes(oracle):
    """test_column_known_bytes (columnar.rs:289-298): Column<i32> of [1, 2, 3] serializes to exactly these bytes."""
    words = oracle.col_encode_slices([struct.pack("<iii", 1, 2, 3)])
    assert words.tobytes() == RAW_COLUMNAR_BYTES
    assert oracle.col_length_in_words([12]) == len(RAW_COLUMNAR_BYTES) // 8


def test_column_from_bytes_and_clone(oracle):
    """test_column_from_bytes / test_column_clone (columnar.rs:260-287,300-339): the bytes borrow back as
    [1, 2, 3], whether they arrive aligned (Column::Bytes) or are relocated first (Column::Align)."""
    words = np.frombuffer(RAW_COLUMNAR_BYTES, dtype="<u8")
    (s,) = oracle.col_decode_slices(words)
    assert struct.unpack("<iii", s) == (1, 2, 3)
    # misaligned arrival: from_bytes relocates into a Vec<u64> (pod_collect_to_vec), then decodes the same
    shifted = bytearray(len(RAW_COLUMNAR_BYTES) + 9)
    shifted[1 : 1 + len(RAW_COLUMNAR_BYTES)] = RAW_COLUMNAR_BYTES
    relocated = np.frombuffer(bytes(shifted[1 : 1 + len(RAW_COLUMNAR_BYTES)]), dtype="<u8")
    assert struct.unpack("<iii", oracle.col_decode_slices(relocated)[0]) == (1, 2, 3)


def test_indexed_layout_of_several_slices(oracle):
    """Every slice starts word aligned; the index holds the unpadded ends; padding is zero."""
    slices = [b"\x01\x02\x03", b"", b"ABCDEFGHIJ", struct.pack("<Q", 7)]
    words = oracle.col_encode_slices(slices)
    assert oracle.col_length_in_words([len(s) for s in slices]) == len(words)
    idx = words[:5].tolist()
    assert idx == [40, 43, 48, 58, 72]
    assert oracle.col_decode_slices(words) == slices
    raw = words.tobytes()
    assert raw[43:48] == b"\0" * 5 and raw[58:64] == b"\0" * 6
    # malformed indexes are rejected, not read
    bad = words.copy()
    bad[0] = 41
    assert oracle.col_decode_slices(bad) is None
    bad = words.copy()
    bad[4] = 8 * len(words) + 1
    assert oracle.col_decode_slices(bad) is None


def _rows(oracle, rng, n, row_keys=False):
    a = np.zeros(n, dtype=oracle.R32)
    if row_keys:
        # Rows of 0..7 bytes packed as mzgpu_rowkey_pack does: len << 56 | bytes big-endian, zero padded
        for f in ("key", "val"):
            lens = rng.integers(0, 8, size=n, dtype=np.uint64)
            body = rng.integers(0, 1 << 56, size=n, dtype=np.uint64)
            mask = np.where(lens == 0, np.uint64(0), ~((np.uint64(1) << (np.uint64(56) - np.uint64(8) * lens)) - np.uint64(1)) & np.uint64((1 << 56) - 1))
            a[f] = (lens << np.uint64(56)) | (body & mask)
    else:
        a["key"] = rng.integers(0, 1 << 63, size=n, dtype=np.uint64) * 2 + 1
        a["val"] = rng.integers(0, 1 << 40, size=n, dtype=np.uint64)
    a["time"] = rng.integers(0, 50, size=n, dtype=np.uint64)
    a["diff"] = rng.integers(-3, 4, size=n)
    return a


@pytest.mark.parametrize("layout,n", [(0, 0), (0, 1), (0, 1000), (1, 777), (2, 0), (2, 1), (2, 1500)])
def test_typed_containers_round_trip(oracle, layout, n):
    rng = np.random.default_rng(layout * 100 + n)
    a = _rows(oracle, rng, n, row_keys=layout == 2)
    words = oracle.column_encode(layout, a)
    back = oracle.column_rows(layout, words)
    if layout == 1:  # (u64, i64): key and diff only
        assert back["key"].tobytes() == a["key"].tobytes() and back["diff"].tobytes() == a["diff"].tobytes()
        assert len(words) == 3 + 2 * n
    else:
        assert back.tobytes() == a.tobytes()
    slices = oracle.col_decode_slices(words)
    assert len(slices) == {0: 4, 1: 2, 2: 6}[layout]
    if layout == 0:
        assert len(words) == 5 + 4 * n
        assert [np.frombuffer(s, dtype="<u8").tolist() for s in slices[:3]] == [a[f].tolist() for f in ("key", "val", "time")]
    if layout == 2 and n:
        # `Rows`: bounds are END offsets (row.rs:606-611), bytes back to back
        kb = np.frombuffer(slices[0], dtype="<u8")
        lens = (a["key"] >> np.uint64(56)).astype(np.uint64)
        assert kb.tolist() == np.cumsum(lens).tolist() and len(slices[1]) == int(lens.sum())
        first = int(lens[0])
        want = [(int(a["key"][0]) >> (8 * (6 - i))) & 0xFF for i in range(first)]
        assert list(slices[1][:first]) == want


def test_row_longer_than_seven_bytes_is_reported(oracle):
    bounds = np.array([8], dtype="<u8").tobytes()
    one = np.array([5], dtype="<u8").tobytes()
    words = oracle.col_encode_slices([bounds, b"12345678", np.array([0], dtype="<u8").tobytes(), b"", one, one])
    with pytest.raises(NotImplementedError):
        oracle.column_rows(2, words)
    with pytest.raises(ValueError):
        oracle.column_rows(0, words)  # six slices are not a ((u64, u64), u64, i64) container


def test_ship_rule_and_builder(oracle):
    """at_serialized_capacity (columnar.rs:164-175) / ColumnBuilder::push_into (builder.rs:44-52): a container is
    minted at the first push that brings the serialized size within 10 % of the next multiple of 2^18 words."""
    assert not oracle.col_at_capacity(0) and not oracle.col_at_capacity(235930)
    assert oracle.col_at_capacity(235931) and oracle.col_at_capacity(262144)
    assert not oracle.col_at_capacity(262145) and oracle.col_at_capacity(2 * 262144 - 52428 + 1)
    rng = np.random.default_rng(5)
    a = _rows(oracle, rng, 130_000)
    chunks = oracle.column_builder(0, a)
    # 5 + 4 n >= 235931 first at n = 58982
    assert [len(c) for c in chunks] == [5 + 4 * 58982, 5 + 4 * 58982, 5 + 4 * (130_000 - 2 * 58982)]
    assert all(oracle.col_at_capacity(len(c)) for c in chunks[:-1])
    back = np.concatenate([oracle.column_rows(0, c) for c in chunks])
    assert back.tobytes() == a.tobytes()
    # Row containers: the cut depends on the bytes; every minted container sits in the ship window and
    # would not have been at capacity one row earlier
    b = _rows(oracle, rng, 150_000, row_keys=True)
    chunks = oracle.column_builder(2, b)
    assert len(chunks) >= 3
    at = 0
    for c in chunks[:-1]:
        assert oracle.col_at_capacity(len(c))
        n = len(oracle.column_rows(2, c))
        assert not oracle.col_at_capacity(len(oracle.column_encode(2, b[at : at + n - 1])))
        at += n
    assert np.concatenate([oracle.column_rows(2, c) for c in chunks]).tobytes() == b.tobytes()
