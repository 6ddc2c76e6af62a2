#!/usr/bin/env python
"""Phase breakdown of the fused consolidate kernel (globaltimer stamps taken by CTA 0),
(a) on synthetic R32 inputs of update-batch size, (b) over real Q3 steps.
Writes one JSON document; used to decide where the kernel's time goes."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import materialize_b200 as mz  # noqa: E402
from materialize_b200 import _ffi as F  # noqa: E402
from materialize_b200 import harness  # noqa: E402

NAMES = ["zero+minmax", "plan", "radix", "gather", "heads", "segsum", "count", "emit", "barrier", "index"]


def phases(ctx):
    buf = (C.c_uint64 * (32 * 4096))()
    n = C.c_uint32(0)
    ctx.check(F.lib.mzgpu_profile_fused_phases(ctx.h, buf, 4096, C.byref(n)))
    a = np.frombuffer(buf, dtype=np.uint64).reshape(-1, 32)[: n.value]
    out = []
    for r in a:
        st = sorted((int(x), i) for i, x in enumerate(r[:16]) if int(x))
        d = {"rows": int(r[16]), "rounds": int(r[17]), "bits": [int((int(r[18]) >> (8 * i)) & 255) for i in range(4)],
             "ctas": int(r[19]), "total_us": (st[-1][0] - st[0][0]) / 1e3,
             "path": {0: "lsd/merge", 1: "msd-warp", 2: "msd-cta", 3: "lsd(overflow)", 4: "fast-msd(64-bit)", 5: "fast-msd(128-bit)"}.get(int(r[21]), "?"),
             "max_bucket": int(r[22]), "buckets": int(r[23]), "row_bytes": int(r[20]),
             # stamp ids: 0 start, 1 after minmax+plan, 10 after MSD pack, 11 after MSD scatter, 7 after emit,
             # 8 after the barrier before the index, 9 end; LSD path: 2 radix, 3 gather, 4 heads, 5 segsum, 6 count
             "stamps": [(i, round((t - st[0][0]) / 1e3, 2)) for t, i in st]}
        out.append(d)
    return out


def main():
    ctx = mz.Context(0)
    res = {"synthetic": [], "q3": []}
    rng = np.random.default_rng(1)
    for n, kb, vb in ((20000, 26, 36), (80000, 28, 36), (80000, 28, 8), (300000, 28, 36), (1000000, 28, 36), (2000, 20, 30)):
        rows = np.zeros(n, dtype=mz.R32)
        rows["key"] = rng.integers(0, 1 << kb, size=n, dtype=np.uint64)
        rows["val"] = rng.integers(0, 1 << vb, size=n, dtype=np.uint64)
        rows["time"] = 5
        rows["diff"] = 1
        d = mz.DeviceRows(ctx, 32)
        for rep in range(3):
            d.upload(rows)
            d.consolidate()
            len(d)
        ctx.profile(True)
        d.upload(rows)
        d.consolidate()
        len(d)
        ph = phases(ctx)
        ctx.profile_report()
        ctx.profile(False)
        res["synthetic"].append({"n": n, "key_bits": kb, "val_bits": vb, "phases": ph})
    # real Q3 steps
    q = harness.Q3Dataflow(ctx, 7, per_batch=10000, n_customer=1500000, n_orders=15000000, n_part=2000000)
    q.hydrate()
    q.clear_out()
    for b in range(6):
        q.stage_batch(b, q.time())
        q.step()
        q.out_rows()
        q.clear_out()
    ctx.profile(True)
    for b in range(6, 9):
        q.stage_batch(b, q.time())
        q.step()
        q.out_rows()
        q.clear_out()
    res["q3"] = phases(ctx)
    rep = ctx.profile_report()
    ctx.profile(False)
    res["q3_kernels"] = rep
    res["stats"] = ctx.stats()
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
