"""Oracle restatement of the MV sink correction buffer (src/compute/src/sink/correction_v2.rs; SURVEY.md
§8(f)-3, groundwork for a device version).  The reference file has no unit tests, so the restatement is
pinned to the documented contract: `updates_before(upper)` = the consolidated updates whose time, advanced
by `since`, lies before `upper`, in (time, data) order — against a brute-force model — and to the chain
invariant its docs state (every chain at least `chain_proportionality` times as long, in chunks, as the next)."""
import math

import numpy as np
import pytest

EMPTY = (1 << 64) - 1


class Model:
    def __init__(self):
        self.rows, self.since = [], 0

    def insert(self, a, negate=False):
        if self.since == EMPTY:
            return
        for k, v, t, d in a.tolist():
            self.rows.append((k, v, max(t, self.since), -d if negate else d))

    def updates_before(self, upper):
        if self.since == EMPTY or not (upper == EMPTY or self.since < upper):
            return []
        acc = {}
        for k, v, t, d in self.rows:
            t = max(t, self.since)
            if upper == EMPTY or t < upper:
                acc[(t, k, v)] = acc.get((t, k, v), 0) + d
        return [(k, v, t, d) for (t, k, v), d in sorted(acc.items()) if d != 0]


def rand(oracle, rng, n, t_lo, t_hi):
    a = np.zeros(n, dtype=oracle.R32)
    a["key"] = rng.integers(0, 40, size=n, dtype=np.uint64)
    a["val"] = rng.integers(0, 3, size=n, dtype=np.uint64)
    a["time"] = rng.integers(t_lo, t_hi, size=n, dtype=np.uint64)
    a["diff"] = rng.integers(-2, 3, size=n)
    return a


@pytest.mark.parametrize("prop,cap", [(3.0, 16), (2.0, 1), (1.5, 64), (4.0, 7)])
def test_correction_contract_and_chain_invariant(oracle, prop, cap):
    rng = np.random.default_rng(int(prop * 10) + cap)
    c, m = oracle.Correction(prop, cap), Model()
    since = 0
    for step in range(60):
        op = rng.integers(0, 10)
        if op < 5:
            a = rand(oracle, rng, int(rng.integers(0, 90)), max(0, since - 3), since + 12)
            neg = bool(rng.integers(0, 2))
            c.insert(a, neg)
            m.insert(a, neg)
            # insert restores the chain invariant
            lens, _ = c.chains()
            chunks = [math.ceil(n / cap) for n in lens]
            for x, y in zip(chunks, chunks[1:]):
                assert y * prop <= x
        elif op < 7:
            since += int(rng.integers(0, 4))
            c.advance_since(since)
            m.since = since
        elif op < 8:
            c.consolidate_at_since()
        else:
            upper = since + int(rng.integers(-1, 8))
            upper = max(upper, 0)
            got = c.updates_before(upper).tolist()
            assert got == m.updates_before(upper)
            lens, staged = c.chains()
            if m.since < upper:
                assert staged == 0  # the stage was flushed into the chains
                chunks = [math.ceil(n / cap) for n in lens]
                for x, y in zip(chunks, chunks[1:]):
                    assert y * prop <= x
    assert c.updates_before(EMPTY).tolist() == m.updates_before(EMPTY)


def test_correction_retractions_cancel_and_empty_since_discards(oracle):
    c = oracle.Correction(3.0, 4)
    a = np.zeros(10, dtype=oracle.R32)
    a["key"] = np.arange(10)
    a["time"] = 5
    a["diff"] = 1
    c.insert(a)
    assert len(c.updates_before(6)) == 10
    assert len(c.updates_before(5)) == 0  # nothing before time 5
    c.insert(a, negate=True)  # what was written comes back negated
    assert len(c.updates_before(EMPTY)) == 0
    c.insert(a)
    c.advance_since(9)
    got = c.updates_before(10)
    assert len(got) == 10 and set(got["time"].tolist()) == {9}  # times advanced by since
    assert len(c.updates_before(9)) == 0  # since is not before upper
    c.advance_since(EMPTY)
    c.insert(a)
    assert len(c.updates_before(EMPTY)) == 0
