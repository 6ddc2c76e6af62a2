#!/usr/bin/env python
"""A batch merge whose advance_by(since) collapses thousands of times of one (key, val) onto one time: the run of
equal (key, val, time') rows exceeds a merge tile's slack, so mergepath.cu's k_mrg_tiles walks it with one
thread (its slow path), and the partition moves tile boundaries across whole tiles.  GPU result vs the oracle's
Batch::Merger restatement, row for row.  Run on a GPU box:  python tools/merge_long_runs_check.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(mz, ctx, B, log=print):
    """GPU merges (mz, ctx) against the oracle's (B); returns True when every case matches row for row."""
    rng = np.random.default_rng(11)
    ok = True
    for case, (n_long, cancel) in enumerate([(700, False), (700, True), (3000, False), (40, False)]):
        halves = []
        for h in range(2):
            n_bg = 6000
            a = np.zeros(n_bg + 2 * n_long, dtype=B.R32)
            a["key"][:n_bg] = rng.integers(0, 300, size=n_bg, dtype=np.uint64)
            a["val"][:n_bg] = rng.integers(0, 4, size=n_bg, dtype=np.uint64)
            a["time"][:n_bg] = rng.integers(h * 5000, (h + 1) * 5000, size=n_bg, dtype=np.uint64)
            a["diff"][:n_bg] = rng.integers(-2, 3, size=n_bg)
            # two long runs: (key 150, val 2) in the middle of the key range, (key 299, val 3) at its end
            for r, (k, v) in enumerate(((150, 2), (299, 3))):
                s = slice(n_bg + r * n_long, n_bg + (r + 1) * n_long)
                a["key"][s], a["val"][s] = k, v
                a["time"][s] = h * 5000 + np.arange(n_long, dtype=np.uint64)
                a["diff"][s] = (-1 if (cancel and h == 1) else 1)
            halves.append(a)
        since = 20000  # every time collapses onto `since`
        g1, g2 = (mz.Batch.build(ctx, halves[0], 0, 5000), mz.Batch.build(ctx, halves[1], 5000, 10000))
        o1, o2 = (B.Batch.build(halves[0], 0, 5000), B.Batch.build(halves[1], 5000, 10000))
        for s in (0, 5000, since):
            gm, om = g1.merge(g2, s), o1.merge(o2, s)
            got, want = gm.rows(), om.rows()
            same = got.tobytes() == want.tobytes() and gm.keys() == om.keys()
            ok = ok and same
            log(f"case {case} (run {n_long} per batch, cancel={cancel}) since={s}: {len(got)} rows vs {len(want)}  {'OK' if same else 'MISMATCH'}")
    return ok


if __name__ == "__main__":
    import materialize_b200 as mz
    from oracle import binding as B  # (the checker)

    print("LONG RUNS OK" if run(mz, mz.Context(0), B) else "LONG RUNS FAILED")
