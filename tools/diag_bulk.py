#!/usr/bin/env python
"""Where the wall time of the bulk regimes (BASELINE configs 2 and 4) goes: wall clock per call, host time in
the allocator and in waits (MZGPU_DEBUG counters), per-kernel device time.  Run on a GPU box."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["MZGPU_DEBUG"] = "1"
import materialize_b200 as mz  # noqa: E402
from materialize_b200 import _ffi as F  # noqa: E402
from materialize_b200 import harness  # noqa: E402

ctx = mz.Context(0)
which = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000_000


def stamp(label, t0):
    ctx.sync()
    print(f"{label}: {1e3 * (time.perf_counter() - t0):.2f} ms", flush=True)
    ctx.stats()


if which == "cfg4":
    nk = 1_000_000
    w = 1.0 / np.power(np.arange(1, nk + 1, dtype=np.float64), 0.9)
    cdf = np.cumsum(w / w.sum())
    cdf[-1] = 1.0
    for rep in range(3):
        d = harness.gen_cfg4(ctx, 3, n, cdf)
        ctx.sync()
        ctx.stats()
        if rep == 2:
            ctx.profile(True)
            ctx.profile_report()
        t0 = time.perf_counter()
        r = mz.ReduceAccumulable(ctx, mz.AGG_COUNT_SUM_I64)
        out = mz.DeviceRows(ctx, 64)
        ctx.check(F.lib.mzgpu_reduce_accumulable(r.h, d.device_ptr(), len(d), F.MEM_DEVICE, 1, out.h))
        print(f"rep {rep} call returned: {1e3 * (time.perf_counter() - t0):.2f} ms", flush=True)
        stamp(f"rep {rep} synced", t0)
        print("groups", len(out))
        if rep == 2:
            for k, v in sorted(ctx.profile_report().items(), key=lambda kv: -kv[1]["ms"])[:14]:
                print(f"   {k:40s} {v['launches']:4d} {v['ms']:9.3f} ms")
        del r, out, d
else:
    for rep in range(3):
        a, b = harness.gen_cfg2(ctx, 1, n, n), harness.gen_cfg2(ctx, 2, n, n)
        ctx.sync()
        ctx.stats()
        if rep == 2:
            ctx.profile(True)
            ctx.profile_report()
        t0 = time.perf_counter()
        ba, bb = mz.Batcher(ctx, 32), mz.Batcher(ctx, 32)
        ba.push_device(a)
        bb.push_device(b)
        xa, xb = ba.seal(1), bb.seal(1)
        stamp(f"rep {rep} seals", t0)
        sa, sb = mz.Spine(ctx, 32), mz.Spine(ctx, 32)
        j = mz.JoinCore(ctx, sa, sb)
        sa.insert(xa)
        j.push(0, xa, 0)
        sb.insert(xb)
        j.push(1, xb, 0)
        stamp(f"rep {rep} pushes", t0)
        j.work()
        stamp(f"rep {rep} work", t0)
        print("out rows", len(j.out))
        if rep == 2:
            for k, v in sorted(ctx.profile_report().items(), key=lambda kv: -kv[1]["ms"])[:14]:
                print(f"   {k:40s} {v['launches']:4d} {v['ms']:9.3f} ms")
        del j, sa, sb, xa, xb, ba, bb, a, b
