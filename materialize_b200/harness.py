"""Binding of libmzgpu_harness.so: the C++ mini-timely worker that drives the
TPC-H-Q3-shaped delta join + reduce dataflow over the C ABI (csrc/harness.cu),
plus device-side seeded generators for the other BASELINE configs."""
import ctypes as C
import os

import numpy as np

from . import _ffi as F
from .api import DeviceRows

_lib = C.CDLL(os.path.join(F.HERE, "libmzgpu_harness.so"))
vp, u64, u32, i32, i64 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int32, C.c_int64
_SIG = {
    "mzh_q3_new": (i32, [vp, u64, u64, u64, u64, u64, u32, u32, C.POINTER(vp)]),
    "mzh_q3_free": (None, [vp]),
    "mzh_q3_hydrate": (i32, [vp, C.POINTER(u64)]),
    "mzh_q3_stage_batch": (i32, [vp, u64, u64, C.POINTER(u64)]),
    "mzh_q3_stage_host": (i32, [vp, i32, vp, u64]),
    "mzh_q3_stage_device": (i32, [vp, i32, vp, u64]),
    "mzh_q3_stage_commit": (i32, [vp]),
    "mzh_q3_h2d_bytes": (u64, [vp]),
    "mzh_q3_d2h_bytes": (u64, [vp]),
    "mzh_q3_pipeline_out": (i32, [vp]),
    "mzh_q3_fetch_out": (i32, [vp, i32, vp, u64, C.POINTER(u64)]),
    "mzh_q3_maintain": (i32, [vp]),
    "mzh_q3_use_p2p": (i32, [vp, i32]),
    "mzh_q3_host_ns": (i32, [vp, C.POINTER(u64)]),
    "mzh_q3_input": (vp, [vp, i32]),
    "mzh_q3_staged": (i32, [vp, i32, vp, u64, C.POINTER(u64)]),
    "mzh_q3_step": (i32, [vp]),
    "mzh_q3_out": (vp, [vp]),
    "mzh_q3_clear_out": (i32, [vp]),
    "mzh_q3_time": (u64, [vp]),
    "mzh_q3_spine": (vp, [vp, i32]),
    "mzh_gen_cfg1": (i32, [vp, u64, u64, u64, u32, vp]),
    "mzh_gen_cfg2": (i32, [vp, u64, u64, u64, u64, vp]),
    "mzh_gen_cfg4": (i32, [vp, u64, u64, u64, vp, u64, i32, u64, i64, vp]),
}
for _n, (_r, _a) in _SIG.items():
    _f = getattr(_lib, _n)
    _f.restype, _f.argtypes = _r, _a


class Q3Dataflow:
    """One worker's share of the Q3 delta-join + reduce dataflow."""

    def __init__(self, ctx, seed, n_customer, n_orders, n_part, per_batch, worker=0, peers=1):
        self.ctx = ctx
        h = vp()
        ctx.check(_lib.mzh_q3_new(ctx.h, seed, n_customer, n_orders, n_part, per_batch, worker, peers, C.byref(h)))
        self.h = h

    def hydrate(self):
        rows = u64(0)
        self.ctx.check(_lib.mzh_q3_hydrate(self.h, C.byref(rows)))
        return rows.value

    def stage_batch(self, b, t):
        rows = u64(0)
        self.ctx.check(_lib.mzh_q3_stage_batch(self.h, b, t, C.byref(rows)))
        return rows.value

    def stage_host(self, a, rows):
        self.ctx.check(_lib.mzh_q3_stage_host(self.h, a, rows.ctypes.data_as(vp), len(rows)))

    def stage_commit(self):
        """The host batch staged with stage_host() is complete (the next step() consumes it)."""
        self.ctx.check(_lib.mzh_q3_stage_commit(self.h))

    def pipeline_out(self):
        """Alternate output buffers per timestamp so that fetch_out(0) can read the previous
        timestamp's corrections while the current one runs."""
        self.ctx.check(_lib.mzh_q3_pipeline_out(self.h))

    def fetch_out(self, which, into):
        """which=0: corrections of the timestamp before the one enqueued last; 1: the latest."""
        n = u64(0)
        self.ctx.check(_lib.mzh_q3_fetch_out(self.h, which, into.ctypes.data_as(vp), len(into), C.byref(n)))
        return into[: n.value]

    def d2h_bytes(self):
        return _lib.mzh_q3_d2h_bytes(self.h)

    def h2d_bytes(self):
        return _lib.mzh_q3_h2d_bytes(self.h)

    def host_ns(self):
        """Host nanoseconds spent so far in the phases of step(): inputs, seals, maintenance, the delta
        paths' two stages, result exchange, reduce, the whole step."""
        out = (C.c_uint64 * 8)()
        self.ctx.check(_lib.mzh_q3_host_ns(self.h, out))
        return dict(zip(("inputs", "seals", "maintenance", "stage1", "stage2", "result_exchange", "reduce", "step"), [int(x) for x in out]))

    def use_p2p(self, on=True):
        """Update-batch exchange rounds over peer memory (landing zones connected by the caller)."""
        self.ctx.check(_lib.mzh_q3_use_p2p(self.h, 1 if on else 0))

    def maintain(self):
        self.ctx.check(_lib.mzh_q3_maintain(self.h))

    def stage_device(self, a, dev_rows):
        self.ctx.check(_lib.mzh_q3_stage_device(self.h, a, dev_rows.device_ptr(), len(dev_rows)))

    def staged(self, a):
        buf = _lib.mzh_q3_input(self.h, a)
        n = F.lib.mzgpu_buf_len(buf)
        out = np.zeros(n, dtype=F.R32)
        got = u64(0)
        self.ctx.check(_lib.mzh_q3_staged(self.h, a, out.ctypes.data_as(vp), n, C.byref(got)))
        return out

    def staged_copy(self, a):
        """Device-resident copy of the staged rows of arrangement `a`."""
        buf = _lib.mzh_q3_input(self.h, a)
        d = DeviceRows(self.ctx, 32)
        self.ctx.check(F.lib.mzgpu_buf_upload(d.h, F.lib.mzgpu_buf_device_ptr(buf), F.lib.mzgpu_buf_len(buf), F.MEM_DEVICE))
        return d

    def step(self):
        self.ctx.check(_lib.mzh_q3_step(self.h))

    def out_len(self):
        return F.lib.mzgpu_buf_len(_lib.mzh_q3_out(self.h))

    def out_rows(self, into=None):
        """Download the output corrections accumulated since the last clear_out()."""
        buf = _lib.mzh_q3_out(self.h)
        n = F.lib.mzgpu_buf_len(buf)
        out = into[:n] if into is not None else np.zeros(n, dtype=F.ROUT)
        got = u64(0)
        self.ctx.check(F.lib.mzgpu_buf_download(buf, out.ctypes.data_as(vp), n, F.MEM_HOST, C.byref(got)))
        return out

    def keep_out(self, acc, max_rows):
        """Append this timestamp's output corrections to the device buffer `acc` (at most `max_rows`
        of them, checked on the device) without reading anything back."""
        self.ctx.check(F.lib.mzgpu_buf_append_buf_at_most(acc.h, _lib.mzh_q3_out(self.h), max_rows))

    def clear_out(self):
        self.ctx.check(_lib.mzh_q3_clear_out(self.h))

    def time(self):
        return _lib.mzh_q3_time(self.h)

    def __del__(self):
        if getattr(self, "h", None) and self.ctx.h:
            _lib.mzh_q3_free(self.h)
            self.h = None


def gen_cfg1(ctx, seed, n, key_bits, first=0):
    d = DeviceRows(ctx, 16)
    ctx.check(_lib.mzh_gen_cfg1(ctx.h, seed, first, n, key_bits, d.h))
    return d


def gen_cfg2(ctx, seed, n, n_keys, first=0):
    d = DeviceRows(ctx, 32)
    ctx.check(_lib.mzh_gen_cfg2(ctx.h, seed, first, n, n_keys, d.h))
    return d


def gen_cfg4(ctx, seed, n, cdf, as_f64=False, first=0, t=0, diff=1):
    d = DeviceRows(ctx, 32)
    cdf = np.ascontiguousarray(cdf, dtype=np.float64)
    ctx.check(_lib.mzh_gen_cfg4(ctx.h, seed, first, n, cdf.ctypes.data_as(vp), len(cdf), 1 if as_f64 else 0, t, diff, d.h))
    return d
