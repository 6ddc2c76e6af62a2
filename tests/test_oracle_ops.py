"""The oracle's join / half_join / reduce restatements have no unit-level golden
vectors in the reference ("parity unpinned"); they are validated
by algebraic identities and by agreement between independent formulations."""
import numpy as np
import pytest


def rand_r32(B, rng, n, key_hi, val_hi, time_lo, time_hi):
    a = np.zeros(n, dtype=B.R32)
    a["key"] = rng.integers(0, key_hi, size=n, dtype=np.uint64)
    a["val"] = rng.integers(0, val_hi, size=n, dtype=np.uint64)
    a["time"] = rng.integers(time_lo, time_hi, size=n, dtype=np.uint64)
    a["diff"] = rng.integers(-2, 3, size=n, dtype=np.int64)
    return a


def brute_join(a, b):
    by_key = {}
    for r in b.tolist():
        by_key.setdefault(r[0], []).append(r)
    out = []
    for k, v1, t1, d1 in a.tolist():
        for _, v2, t2, d2 in by_key.get(k, ()):
            out.append((k, v1, v2, max(t1, t2), d1 * d2))
    return out


@pytest.mark.parametrize("strategy", [0, 1, 2])
def test_join_core_equals_bruteforce(oracle, strategy):
    """A7: every pair of updates is joined exactly once, whatever the batch arrival
    order, and both join_key strategies give the same consolidated output."""
    B = oracle
    rng = np.random.default_rng(strategy)
    s1, s2 = B.Spine(32, 1, True), B.Spine(32, 1, True)
    j = B.Join(s1, s2, None, strategy)
    A, Bb = [], []
    for t in range(0, 12, 2):
        order = [0, 1] if (t // 2) % 2 == 0 else [1, 0]
        for side in order:
            # few keys, many times per key so the linear time scan path (>= 10 edits) runs
            a = rand_r32(B, rng, int(rng.integers(0, 400)), 6, 3, t, t + 2)
            (A if side == 0 else Bb).append(a)
            batch = B.Batch.build(a, t, t + 2)
            (s1 if side == 0 else s2).insert(batch)
            j.push(side, batch, t)
        j.work()
    got = B.consolidate(j.results())
    want = B.consolidate(B.rows(B.R40, brute_join(np.concatenate(A), np.concatenate(Bb))))
    assert got.tobytes() == want.tobytes()


def test_half_join_tiebreak_counts_each_pair_once(oracle):
    """A8: with `le` on one path and `lt` on the other, the two delta paths together
    produce exactly the change of the full join (no pair counted twice or missed)."""
    B = oracle
    rng = np.random.default_rng(7)
    sa, sb = B.Spine(32, 1, True), B.Spine(32, 1, True)
    total = []
    accA, accB = [], []
    for t in range(5):
        a = rand_r32(B, rng, 300, 40, 5, t, t + 1)
        b = rand_r32(B, rng, 300, 40, 5, t, t + 1)
        ba, bb = B.Batch.build(a, t, t + 1), B.Batch.build(b, t, t + 1)
        sa.insert(ba)
        sb.insert(bb)
        # path A (relation 0): lookups into B use `le`; path B (relation 1): lookups into A use `lt`
        cl_a = B.make_closure(key_fields=[(0, 0, 64, 0)], val_fields=[(1, 0, 8, 0), (2, 0, 8, 8)])
        cl_b = B.make_closure(key_fields=[(0, 0, 64, 0)], val_fields=[(2, 0, 8, 0), (1, 0, 8, 8)])
        total.append(B.half_join(B.update_stream(ba), sb, 0, cl_a))
        total.append(B.half_join(B.update_stream(bb), sa, 1, cl_b))
        accA.append(a)
        accB.append(b)
    got = B.consolidate(np.concatenate(total))
    full = brute_join(B.consolidate(np.concatenate(accA)), B.consolidate(np.concatenate(accB)))
    want = B.consolidate(B.rows(B.R32, [(k, v1 | (v2 << 8), t, d) for k, v1, v2, t, d in full]))
    assert got.tobytes() == want.tobytes()


def test_reduce_accumulates_to_group_by(oracle):
    """A9: the accumulated output collection equals GROUP BY over the accumulated input, at every time."""
    B = oracle
    rng = np.random.default_rng(9)
    r = B.Reduce(0)
    outputs, inputs = [], []
    for t in range(6):
        a = np.zeros(500, dtype=B.R32)
        a["key"] = rng.integers(0, 30, size=500, dtype=np.uint64)
        a["val"] = rng.integers(-50, 50, size=500, dtype=np.int64).astype(np.uint64)
        a["time"] = t
        a["diff"] = rng.integers(-1, 3, size=500, dtype=np.int64)
        inputs.append(a)
        outputs.append(r.step(a, t + 1))
        acc_in = np.concatenate(inputs)
        want = {}
        for k, v, _, d in acc_in.tolist():
            c, s = want.get(k, (0, 0))
            v = v - 2**64 if v >= 2**63 else v
            want[k] = (c + d, s + v * d)
        want = {k: cs for k, cs in want.items() if cs != (0, 0)}
        acc_out = np.concatenate(outputs).copy()
        acc_out["time"] = 0
        acc_out = B.consolidate(acc_out)
        got = {}
        for k, c, lo, hi, flags, _, d, _ in acc_out.tolist():
            assert d == 1
            got[k] = (c, (hi << 64) + lo if hi >= 0 else (hi << 64) + lo)
        assert got == {k: (c, s) for k, (c, s) in want.items()}


def q3_bruteforce(tables):
    """Direct evaluation of Q3 on the accumulated base tables (multisets)."""
    cust = {}
    for k, v, _, d in tables[0]:
        cust[(k, v)] = cust.get((k, v), 0) + d
    orders = {}
    for k, v, _, d in tables[1]:
        orders[(k, v)] = orders.get((k, v), 0) + d
    li = {}
    for k, v, _, d in tables[3]:
        li.setdefault(k, {})
        li[k][v] = li[k].get(v, 0) + d
    building = {}
    for (ck, seg), d in cust.items():
        if seg == 1 and d:
            building[ck] = building.get(ck, 0) + d
    out = {}
    for (ok, v), d in orders.items():
        if d == 0:
            continue
        custkey, odate, prio = v & 0xFFFFFF, (v >> 24) & 0xFFF, (v >> 36) & 1
        if odate >= 1169 or custkey not in building:
            continue
        for lv, ld in li.get(ok, {}).items():
            ext, disc, ship = (lv >> 3) & 0x1FFFF, (lv >> 20) & 0xF, (lv >> 24) & 0xFFF
            if ship <= 1169 or ld == 0:
                continue
            g = ok | (odate << 32) | (prio << 44)
            mult = d * ld * building[custkey]
            c, s = out.get(g, (0, 0))
            out[g] = (c + mult, s + mult * ext * (100 - disc))
    return {g: cs for g, cs in out.items() if cs != (0, 0)}


@pytest.mark.parametrize("workers", [1, 3])
def test_q3_dataflow_equals_direct_query(oracle, workers):
    """Sum over the three delta paths == change of the full join; reduce on top of it
    accumulates to the GROUP BY of the direct query — checked after hydration and
    after every update batch, for 1 and for several key-sharded workers."""
    B = oracle
    q = B.Q3(seed=7, n_customer=300, n_orders=3000, n_part=400, workers=workers, per_batch=50)
    tables = {0: [], 1: [], 3: []}
    outs = []

    def check():
        acc = np.concatenate(outs).copy()
        acc["time"] = 0
        acc = B.consolidate(acc)
        got = {}
        for k, c, lo, hi, flags, _, d, _ in acc.tolist():
            assert d == 1 and flags == 0
            got[k] = (c, (hi << 64) + lo)
        assert got == q3_bruteforce(tables)

    q.hydrate()
    for a in (0, 1, 3):
        tables[a] += q.inputs(a).tolist()
    outs.append(q.drain())
    assert len(outs[0]) > 0
    check()
    for b in range(4):
        q.step(b)
        for a in (1, 3):
            tables[a] += q.inputs(a).tolist()
        outs.append(q.drain())
        check()


def test_q3_workers_agree(oracle):
    B = oracle
    res = []
    for w in (1, 2, 4):
        q = B.Q3(seed=3, n_customer=200, n_orders=2000, n_part=300, workers=w, per_batch=40)
        q.hydrate()
        outs = [q.drain()]
        for b in range(3):
            q.step(b)
            outs.append(q.drain())
        res.append(np.concatenate(outs).tobytes())
    assert res[0] == res[1] == res[2]


def test_distinct_and_threshold_accumulate_to_their_definitions():
    """ReducePlan::Distinct keeps (key, ()) once while the multiplicity is non-zero (error flag when it is
    negative, reduce.rs:264-334); ThresholdPlan::Basic keeps a row with its multiplicity while that is
    positive (threshold.rs:33-77).  The oracle's output corrections must accumulate to exactly that at
    every time."""
    import numpy as np

    from oracle import binding as oracle

    rng = np.random.default_rng(3)
    for kind in (2, 3):
        r = oracle.Reduce(kind)
        acc, out_acc = {}, {}
        for step in range(6):
            n = 2000
            a = np.zeros(n, dtype=oracle.R32)
            a["key"] = rng.integers(0, 150, size=n, dtype=np.uint64)
            a["time"] = step
            a["diff"] = rng.integers(-2, 3, size=n)
            out = r.step(a, step + 1)
            for k, d in zip(a["key"].tolist(), a["diff"].tolist()):
                acc[k] = acc.get(k, 0) + d
            for row in out.tolist():
                key = (row[0], row[1], row[2], row[3], row[4])
                out_acc[key] = out_acc.get(key, 0) + row[6]
            out_acc = {k: v for k, v in out_acc.items() if v}
            if kind == 2:
                want = {(k, 1, 0, 0, 2 if c < 0 else 0): 1 for k, c in acc.items() if c != 0}
            else:
                want = {(k, 0, 0, 0, 0): c for k, c in acc.items() if c > 0}
            assert out_acc == want, (kind, step)


@pytest.mark.parametrize("kind", [4, 5])
def test_min_max_accumulate_to_their_definitions(oracle, kind):
    """MIN / MAX (reduce.rs:1050-1135): the accumulated output equals func(values with positive
    count) per key, or the error row when any count is negative."""
    rng = np.random.default_rng(5 + kind)
    r = oracle.Reduce(kind)
    acc, out_acc = {}, {}
    for step in range(8):
        n = 1500
        a = np.zeros(n, dtype=oracle.R32)
        a["key"] = rng.integers(0, 120, size=n, dtype=np.uint64)
        a["val"] = rng.integers(0, 12, size=n, dtype=np.uint64)
        a["time"] = rng.integers(step * 2, step * 2 + 2, size=n, dtype=np.uint64)
        a["diff"] = rng.integers(-1, 3, size=n)
        out = r.step(a, step * 2 + 2)
        for k, v, d in zip(a["key"].tolist(), a["val"].tolist(), a["diff"].tolist()):
            acc.setdefault(k, {})
            acc[k][v] = acc[k].get(v, 0) + d
        for row in out.tolist():
            key = (row[0], row[1], row[2], row[3], row[4])
            out_acc[key] = out_acc.get(key, 0) + row[6]
        out_acc = {k: v for k, v in out_acc.items() if v}
        want = {}
        for k, m in acc.items():
            nz = {v: c for v, c in m.items() if c != 0}
            if not nz:
                continue
            if any(c < 0 for c in nz.values()):
                want[(k, 0, 0, 0, 2)] = 1
            else:
                want[(k, 0, (min if kind == 4 else max)(nz), 0, 0)] = 1
        assert out_acc == want


@pytest.mark.parametrize("limit,offset,desc", [(3, 0, False), (2, 1, True), (-1, 2, False), (0, 0, False), (5, 0, True)])
def test_topk_accumulates_to_its_definition(oracle, limit, offset, desc):
    """TopK (top_k.rs:521-673): the accumulated output is, per key, the rows [offset, offset+limit) of
    the live values in order with multiplicities, or the error row when a count is negative."""
    rng = np.random.default_rng(9)
    r = oracle.TopK(limit, offset, desc)
    acc, out_acc = {}, {}
    for step in range(8):
        n = 1200
        a = np.zeros(n, dtype=oracle.R32)
        a["key"] = rng.integers(0, 100, size=n, dtype=np.uint64)
        a["val"] = rng.integers(0, 10, size=n, dtype=np.uint64)
        a["time"] = rng.integers(step * 2, step * 2 + 2, size=n, dtype=np.uint64)
        a["diff"] = rng.integers(-1, 3, size=n)
        out = r.step(a, step * 2 + 2)
        for k, v, d in zip(a["key"].tolist(), a["val"].tolist(), a["diff"].tolist()):
            acc.setdefault(k, {})
            acc[k][v] = acc[k].get(v, 0) + d
        for row in out.tolist():
            key = (row[0], row[2], row[4])
            out_acc[key] = out_acc.get(key, 0) + row[6]
        out_acc = {k: v for k, v in out_acc.items() if v}
        want = {}
        for k, m in acc.items():
            nz = {v: c for v, c in m.items() if c != 0}
            if not nz:
                continue
            if any(c < 0 for c in nz.values()):
                want[(k, 0, 2)] = 1
                continue
            rows = []
            for v in sorted(nz, reverse=desc):
                rows += [v] * nz[v]
            rows = rows[offset:]
            if limit >= 0:
                rows = rows[:limit]
            for v in rows:
                want[(k, v, 0)] = want.get((k, v, 0), 0) + 1
        assert out_acc == want


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_spine_preserves_contents_under_random_maintenance(oracle, seed):
    """Whatever the fueled spine does (merges, roll-ups, idle effort, physical/logical compaction —
    trace.rs:1698-2047), its accumulated contents are the consolidation of everything inserted with
    times advanced to `since`, and every batch's description tiles [0, upper)."""
    rng = np.random.default_rng(seed)
    sp = oracle.Spine(32, 1, True)
    inserted = []
    t = 0
    since = 0
    for step in range(40):
        n = int(rng.integers(0, 3000))
        a = np.zeros(n, dtype=oracle.R32)
        a["key"] = rng.integers(0, 300, size=n, dtype=np.uint64)
        a["val"] = rng.integers(0, 5, size=n, dtype=np.uint64)
        a["time"] = t
        a["diff"] = rng.integers(-2, 3, size=n)
        sp.insert(oracle.Batch.build(a, t, t + 1))
        inserted.append(a)
        t += 1
        sp.set_physical_compaction(t)
        if rng.integers(0, 3) == 0:
            since = max(since, int(rng.integers(0, t)))
            sp.set_logical_compaction(since)
        if rng.integers(0, 4) == 0:
            e = sp.exert_logic(16)
            if e:
                sp.exert(e)
        assert sp.read_upper() == t
        got = sp.export()
        allr = np.concatenate(inserted)
        allr["time"] = np.maximum(allr["time"], np.uint64(since))
        # export folds with advance_by(since) applied: compare as consolidated collections
        want = oracle.consolidate(allr)
        assert oracle.consolidate(got).tobytes() == want.tobytes()


# ------------------------------------------------ reference-held SQL known answers
def test_oracle_operators_reproduce_sqllogictest_answers(oracle):
    """Pins the join / distinct / reduce restatements to answers the reference itself holds: the
    integer-only cases of test/sqllogictest/joins.slt and aggregates.slt (fixture cites the lines)."""
    import sql_golden as sg

    fx = sg.load()
    ops = sg.OracleOps(oracle)
    ran = 0
    for case in fx["cases"]:
        if case["shape"] == "sum_of_nulls":
            continue
        got = sg.norm(sg.evaluate(ops, case, fx["tables"]))
        want = sg.norm([tuple(r) for r in case["expect"]])
        assert got == want, (case["name"], case["cite"], got, want)
        ran += 1
    assert ran >= 12


def test_oracle_sum_of_nulls_is_null(oracle):
    """aggregates.slt:156-176: SUM over only-NULL inputs is NULL.  NULL inputs never reach explode (the
    fixed-width subset has no NULL datum), so this pins finalize_accum alone (reduce.rs:1905-1937): an
    accumulator that counted two rows and no non-NULL value finalizes to the NULL-sum flag."""
    acc = np.zeros(1, dtype=oracle.RACC)
    acc["key"], acc["total"], acc["non_nulls"] = 0, 2, 0
    out = oracle.finalize(acc, 0)
    assert int(out["flags"][0]) & 1
    acc["non_nulls"] = 2
    assert int(oracle.finalize(acc, 0)["flags"][0]) & 1 == 0


def test_oracle_topk_reproduces_sqllogictest_answers(oracle):
    """Pins the oracle's TopK operator to reference-held known answers: the integer-encodable cases of
    test/sqllogictest/topk.slt (tests/golden/sqllogictest_topk.json cites the lines).  A city is one
    (key = state, val = pop << 8 | city) row: the operator orders a key's values, so the population leads the
    word and the city rides in its low bits; NULL sorts above every population (DESC NULLS FIRST) or below
    (DESC NULLS LAST), as `order_by=[#2{pop} desc nulls_first]` / `nulls_last` ask."""
    import json
    import os

    fx = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sqllogictest_topk.json")))
    rows = fx["cities"]["rows"]
    states = sorted({r[1] for r in rows})
    names = [r[0] for r in rows]
    for case in fx["per_group"]:
        a = np.zeros(len(rows), dtype=oracle.R32)
        for i, (name, state, pop) in enumerate(rows):
            if pop is None:
                pop = (1 << 40) - 1 if case["nulls_first"] else 0
            a[i] = (states.index(state), (pop << 8) | names.index(name), 0, 1)
        out = oracle.TopK(case["limit"], 0, case["descending"]).step(a, 1)
        got = set()
        for row in out.tolist():
            assert row[4] == 0 and row[6] == 1, row  # no error row, every city once
            got.add((states[row[0]], names[row[2] & 0xFF]))
        assert got == {tuple(x) for x in case["answer"]}, case["name"]
        assert len(out) == len(case["answer"])
    for case in fx["global"]:
        cur = np.zeros(len(case["t"]), dtype=oracle.R32)
        cur["val"] = np.array(case["t"], dtype=np.uint64)
        cur["diff"] = 1
        for st in case["stages"]:
            out = oracle.TopK(st["limit"], st["offset"], False).step(cur, 1)
            cur = np.zeros(len(out), dtype=oracle.R32)
            cur["val"] = np.array([row[2] for row in out.tolist()], dtype=np.uint64)
            cur["diff"] = np.array([row[6] for row in out.tolist()], dtype=np.int64)
        assert sorted(cur["val"].tolist()) == case["answer"], case["name"]
        assert (cur["diff"] == 1).all()

