"""Pin the CPU oracle to the reference's own golden vectors.

Every case here is transcribed from a unit test, property test or datadriven
trace under /root/reference (citations inside tests/golden/*.json).
"""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def load(name):
    with open(os.path.join(HERE, "golden", name)) as f:
        return json.load(f)


VEC = load("consolidate_vectors.json")
TRACES = load("spine_traces.json")


def gen_rows(spec):
    if spec["kind"] == "cancel_pairs":
        rows = []
        for _ in range(spec["pairs"]):
            rows.append([spec["key"], 0, 1])
            rows.append([spec["key"], 0, -1])
        return rows
    if spec["kind"] == "repeat":
        return [spec["row"]] * spec["n"]
    if spec["kind"] == "distinct":
        return [[d, 0, 1] for d in range(spec["n"])]
    raise ValueError(spec)


def as_r32_from_dtr(rows, B):
    # (data, time, diff) -> R32 with val = 0
    return B.rows(B.R32, [(d, 0, t, r) for d, t, r in rows])


def as_r32_from_kvtr(rows, B):
    return B.rows(B.R32, [tuple(r) for r in rows])


@pytest.mark.parametrize("case", VEC["chunker_u64"]["cases"] + VEC["builder_u64"]["cases"], ids=lambda c: c["name"])
def test_consolidate_u64_vectors(oracle, case):
    B = oracle
    rows = case["input"] if "input" in case else gen_rows(case["gen"])
    expected = case["expected"] if "expected" in case else gen_rows(case["expected_gen"])
    got = B.consolidate(as_r32_from_dtr(rows, B))
    want = as_r32_from_dtr(expected, B)
    assert got.tolist() == want.tolist()
    # the same vectors through the merge batcher (Chunker::push_into + seal)
    batcher = B.Batcher(32)
    batcher.push(as_r32_from_dtr(rows, B))
    batch = batcher.seal(B.FRONTIER_EMPTY)
    assert batch.rows().tolist() == want.tolist()


@pytest.mark.parametrize("case", VEC["chunker_keyval"]["cases"], ids=lambda c: c["name"])
def test_consolidate_keyval_vectors(oracle, case):
    B = oracle
    got = B.consolidate(as_r32_from_kvtr(case["input"], B))
    assert got.tolist() == as_r32_from_kvtr(case["expected"], B).tolist()


@pytest.mark.parametrize("case", VEC["merger_keyval"]["cases"], ids=lambda c: c["name"])
def test_merger_vectors(oracle, case):
    B = oracle
    c1 = [r for chunk in case["chain1"] for r in chunk]
    c2 = [r for chunk in case["chain2"] for r in chunk]
    a, b = as_r32_from_kvtr(c1, B), as_r32_from_kvtr(c2, B)
    out = np.zeros(len(a) + len(b), dtype=B.R32)
    chunk_rows = len(case["chain1"][0])
    n = B.lib().mzo_merge_chains_r32(a.ctypes.data, len(a), b.ctypes.data, len(b), chunk_rows, out.ctypes.data)
    assert out[:n].tolist() == as_r32_from_kvtr(case["expected"], B).tolist()
    # and as a batch merge (Batch::Merger) with since = 0
    b1 = B.Batch.build(a, 0, 1)
    b2 = B.Batch.build(b, 1, 2)
    assert b1.merge(b2, 0).rows().tolist() == as_r32_from_kvtr(case["expected"], B).tolist()


def model_consolidate(rows):
    """The reference model, src/timely-util/src/columnar/batcher.rs:1116-1130."""
    out = []
    for r in sorted(rows):
        if out and out[-1][:-1] == r[:-1]:
            out[-1] = out[-1][:-1] + (out[-1][-1] + r[-1],)
        else:
            out.append(r)
    return [r for r in out if r[-1] != 0]


def arb_consolidated(rng):
    """arb_consolidated, batcher.rs:1134-1137: small ranges so collisions are common."""
    n = int(rng.integers(0, 30))
    rows = [
        (int(rng.integers(0, 5)), int(rng.integers(0, 5)), int(rng.integers(0, 3)), int(rng.integers(-3, 4)))
        for _ in range(n)
    ]
    return model_consolidate(rows)


def test_property_merge_equals_consolidated_union(oracle):
    """merge_from_equals_consolidated_union, batcher.rs:1185-1199."""
    B = oracle
    rng = np.random.default_rng(1)
    for _ in range(300):
        a, b = arb_consolidated(rng), arb_consolidated(rng)
        ra, rb = B.rows(B.R32, a), B.rows(B.R32, b)
        out = np.zeros(len(a) + len(b) + 1, dtype=B.R32)
        for chunk in (1, 3, 64):
            n = B.lib().mzo_merge_chains_r32(ra.ctypes.data, len(ra), rb.ctypes.data, len(rb), chunk, out.ctypes.data)
            assert [tuple(x) for x in out[:n].tolist()] == model_consolidate(a + b)


def test_property_extract_partitions_by_frontier(oracle):
    """extract_partitions_by_frontier, batcher.rs:1257-1310."""
    B = oracle
    rng = np.random.default_rng(2)
    for _ in range(300):
        data = arb_consolidated(rng)
        upper = int(rng.integers(0, 5))
        rows = B.rows(B.R32, data)
        ship = np.zeros(len(rows) + 1, dtype=B.R32)
        keep = np.zeros(len(rows) + 1, dtype=B.R32)
        import ctypes as C

        ns, nk = C.c_uint64(0), C.c_uint64(0)
        frontier = B.lib().mzo_extract_r32(
            rows.ctypes.data, len(rows), upper, ship.ctypes.data, C.addressof(ns), keep.ctypes.data, C.addressof(nk)
        )
        kept = [tuple(x) for x in keep[: nk.value].tolist()]
        shipped = [tuple(x) for x in ship[: ns.value].tolist()]
        assert all(t >= upper for (_, _, t, _) in kept)
        assert all(t < upper for (_, _, t, _) in shipped)
        assert sorted(kept + shipped) == sorted(data)
        if kept:
            assert frontier == min(t for (_, _, t, _) in kept)
        else:
            assert frontier == B.FRONTIER_EMPTY


def test_property_consolidate_u64_i64(oracle):
    """test_consolidate_sorted, src/ore/src/iter.rs:260-268: consolidate on Vec<(u64,i64)>
    equals streaming consolidation of the sorted input (BASELINE config 1's exact type)."""
    B = oracle
    rng = np.random.default_rng(3)
    for n in (0, 1, 2, 17, 1000):
        keys = rng.integers(0, max(1, n // 3 + 1), size=n, dtype=np.uint64)
        diffs = rng.integers(-3, 4, size=n, dtype=np.int64)
        a = np.zeros(n, dtype=B.R16)
        a["key"], a["diff"] = keys, diffs
        got = B.consolidate(a)
        want = model_consolidate([(int(k), int(d)) for k, d in zip(keys, diffs)])
        assert [tuple(x) for x in got.tolist()] == want
    # wrapping i64 addition (Overflowing<i64> in release mode)
    a = B.rows(B.R16, [(1, 2**63 - 1), (1, 1)])
    assert B.consolidate(a).tolist() == [(1, -(2**63))]


@pytest.mark.parametrize("trace", TRACES["traces"], ids=lambda t: t["name"])
def test_spine_golden_traces(oracle, trace):
    B = oracle
    s = B.HollowSpine()
    for cmd, arg in trace["script"]:
        if cmd == "push":
            lower, upper, since, length, name = arg
            s.push(lower, upper, since, length, name)
        elif cmd == "since":
            s.downgrade_since(arg)
        elif cmd == "expect-batches":
            got = [ln.rstrip() for ln in s.describe().splitlines()]
            assert got == arg, (trace["name"], got)
        elif cmd == "expect-since-upper":
            assert [s.since(), s.upper()] == arg
        else:
            raise ValueError(cmd)


def test_batch_csr_layout(oracle):
    """OrdValBatch CSR arrays (keys / vals.offs / upds.offs), src/compute/src/extensions/arrange.rs:325-330."""
    B = oracle
    rows = B.rows(B.R32, [(1, 10, 0, 1), (1, 10, 1, 1), (1, 11, 0, 1), (2, 10, 0, 1), (5, 1, 3, -1)])
    b = B.Batch.build(rows, 0, 4)
    assert len(b) == 5 and b.keys() == 3
    assert b.csr_sizes() == (3, 4, 5)  # 3 keys, 3+1 key offsets, 4 (key,val) runs + 1
    assert b.desc() == (0, 4, 0)
