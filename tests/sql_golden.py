"""Evaluates the reference-held SQL known answers of tests/golden/sqllogictest_join_reduce.json with
the operator core, the way the reference's planner renders each query shape:

  inner equi-join           ArrangeBy both sides -> mz_join_core (src/compute/src/render/join/mz_join_core.rs)
  LEFT / RIGHT JOIN         inner join  UNION  (outer side  MINUS  outer side semijoin Distinct(keys of the
                            other side)) padded with NULLs  (src/sql/src/plan/lowering.rs: outer joins lower to
                            Join + Distinct + Negate + Union)
  x IN (subquery)           semijoin against Distinct(subquery)  (the plan printed at joins.slt:147-196)
  multi-way join            linear join: re-arrange the running result by the next key, join_core again
                            (src/compute/src/render/join/linear_join.rs:327-527)
  GROUP BY count/sum        accumulable reduce (src/compute/src/render/reduce.rs:1261-1471)
  GROUP BY min/max          hierarchical reduce's result (reduce.rs:796-1135)
  GROUP BY (no aggregates)  Distinct (reduce.rs:264-334)

Scalar expressions in keys (la + 1, 2 * a) are the MapFilterProject in front of ArrangeBy and are applied
on the host here -- they are not part of the operator core.  `ops` is an adapter over either the CPU oracle
(oracle/binding.py) or the GPU library (materialize_b200), so the same evaluation pins both.
"""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURE = os.path.join(HERE, "golden", "sqllogictest_join_reduce.json")

AGG_COUNT_SUM_I64, AGG_DISTINCT, AGG_MIN, AGG_MAX = 0, 2, 4, 5


def load():
    with open(FIXTURE) as f:
        return json.load(f)


class OracleOps:
    name = "oracle"

    def __init__(self, oracle):
        self.o = oracle
        self.R32, self.R40 = oracle.R32, oracle.R40

    def arrange(self, rows):
        s = self.o.Spine(32, 1, True)
        s.insert(self.o.Batch.build(rows, 0, 1))
        return s

    def join(self, a, b):
        j = self.o.Join(a, b)
        j.work()
        return self.o.consolidate(j.results())

    def reduce(self, kind, rows):
        return self.o.Reduce(kind).step(rows, 1)

    def consolidate(self, rows):
        return self.o.consolidate(rows)


class GpuOps:
    name = "gpu"

    def __init__(self, mz, ctx, oracle_consolidate=None):
        self.mz, self.ctx = mz, ctx
        self.R32, self.R40 = mz.R32, mz.R40

    def arrange(self, rows):
        s = self.mz.Spine(self.ctx, 32)
        s.insert(self.mz.Batch.build(self.ctx, rows, 0, 1))
        return s

    def join(self, a, b):
        j = self.mz.JoinCore(self.ctx, a, b)
        j.work()
        return self.ctx.consolidate(j.results())

    def reduce(self, kind, rows):
        return self.mz.ReduceAccumulable(self.ctx, kind).step(rows, 1)

    def consolidate(self, rows):
        return self.ctx.consolidate(rows)


def kv(ops, pairs):
    """(key, val) pairs -> R32 updates at time 0 with diff +1 (an INSERTed table)."""
    a = np.zeros(len(pairs), dtype=ops.R32)
    a["key"] = np.array([p[0] for p in pairs], dtype=np.int64).view(np.uint64)
    a["val"] = np.array([p[1] for p in pairs], dtype=np.int64).view(np.uint64)
    a["diff"] = 1
    return a


def collection(rows, cols):
    """Consolidated update rows -> list of tuples, one per unit of (positive) multiplicity."""
    out = []
    for r in rows:
        assert int(r["diff"]) > 0, "negative multiplicity in a query result"
        out.extend([tuple(int(np.int64(r[c])) if c is not None else None for c in cols)] * int(r["diff"]))
    return out


def inner_join(ops, left, right):
    """[(key, lval, rval)] of left JOIN right ON key."""
    return collection(ops.join(ops.arrange(kv(ops, left)), ops.arrange(kv(ops, right))), ["key", "val1", "val2"])


def distinct_keys(ops, keys):
    out = ops.reduce(AGG_DISTINCT, kv(ops, [(k, 0) for k in keys]))
    assert (out["flags"] == 0).all()
    return [k for (k,) in collection(out, ["key"])]


def semijoin(ops, rows, keys):
    """rows whose key is in `keys` (a Distinct collection): join_core against (key, ()) rows."""
    return [(k, v) for (k, v, _) in inner_join(ops, rows, [(k, 0) for k in keys])]


def minus(ops, rows, sub):
    """rows EXCEPT ALL sub through Negate + Union + consolidate."""
    a, b = kv(ops, rows), kv(ops, sub)
    b["diff"] = -1
    return collection(ops.consolidate(np.concatenate([a, b])), ["key", "val"])


def outer_side(ops, outer, other):
    """Rows of `outer` with no partner in `other`."""
    matched = semijoin(ops, outer, distinct_keys(ops, [k for k, _ in other]))
    return minus(ops, outer, matched)


def group(ops, kind, pairs):
    out = ops.reduce(kind, kv(ops, pairs))
    res = {}
    for r in out:
        assert int(r["diff"]) == 1 and int(r["flags"]) == 0
        res[int(np.int64(r["key"]))] = r
    return res


def sums(ops, pairs):
    """{key: (count, i128 sum)}"""
    return {
        k: (int(np.int64(r["count"])), (int(np.int64(r["sum_hi"])) << 64) + int(r["sum_lo"]))
        for k, r in group(ops, AGG_COUNT_SUM_I64, pairs).items()
    }


def evaluate(ops, case, tables):
    """Result rows of one fixture case, as a sorted list of tuples (None = NULL)."""
    shape = case["shape"]
    L = [tuple(r) for r in tables["l"]["rows"]]
    R = [tuple(r) for r in tables["r"]["rows"]]
    T = [tuple(r) for r in tables["t"]["rows"]]
    TB = [tuple(r) for r in tables["t_bigint"]["rows"]]
    if shape in ("left_join", "right_join"):
        both = [(k, lv, k, rv) for (k, lv, rv) in inner_join(ops, L, R)]
        if shape == "left_join":
            both += [(k, v, None, None) for (k, v) in outer_side(ops, L, R)]
        else:
            both += [(None, None, k, v) for (k, v) in outer_side(ops, R, L)]
        return both
    if shape == "three_way_scalar_keys":
        # stage 1: l1 arranged by (la + 1), l2 by la
        s1 = inner_join(ops, [(la + 1, la) for la, _ in L], L)  # (l2.la, l1.la, l2.lb)
        # stage 2 (linear join): the running result re-arranged by l1.la + l2.la, joined with l3 by la
        packed = [(l1a + l2a, l1a | (l2b << 16)) for (l2a, l1a, l2b) in s1]
        s2 = inner_join(ops, packed, L)  # (l3.la, packed, l3.lb)
        return [(p & 0xFFFF, p >> 16, l3b) for (_, p, l3b) in s2]
    if shape == "nested_in_subqueries":
        s3 = distinct_keys(ops, [la + 1 for la, _ in L])
        l2 = semijoin(ops, L, s3)
        s2 = distinct_keys(ops, [la + 1 for la, _ in l2])
        return semijoin(ops, L, s2)
    if shape == "global_sums":
        (ca, sa), (cb, sb) = sums(ops, [(0, a) for a, _ in T])[0], sums(ops, [(0, b) for _, b in T])[0]
        assert ca == cb == len(T)
        return [(1, sa, sb, sa / ca)]
    if shape == "having_sum":
        return [(k,) for k, (_, s) in sums(ops, T).items() if s == 3]
    if shape == "having_sum_expr_key":
        return [(k,) for k, (_, s) in sums(ops, [(a + 1, b) for a, b in T]).items() if s == 3]
    if shape == "group_sum":
        return [(k, s) for k, (_, s) in sums(ops, T).items()]
    if shape == "group_distinct_keys":
        return [(k,) for k in distinct_keys(ops, [a for a, _ in T])]
    if shape == "count_min_sum_max":
        cs = sums(ops, T)
        mn, mx = group(ops, AGG_MIN, T), group(ops, AGG_MAX, T)
        return [(k, cs[k][0], int(np.int64(mn[k]["sum_lo"])), cs[k][1], int(np.int64(mx[k]["sum_lo"]))) for k in cs]
    if shape == "expr_key_sum":
        return [(k, s) for k, (_, s) in sums(ops, [(2 * a, b) for a, b in T]).items()]
    if shape == "bigint_sums":
        return [(k, s) for k, (_, s) in sums(ops, TB).items()]
    raise KeyError(shape)


def norm(rows):
    return sorted((tuple(r) for r in rows), key=lambda r: tuple((x is None, x if x is not None else 0) for x in r))
