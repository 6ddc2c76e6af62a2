"""ctypes binding of the CPU oracle (oracle/libmzoracle.so).

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / ``--impl reference`` legs may import this module;
nothing under materialize_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmzoracle.so")

R16 = np.dtype([("key", "<u8"), ("diff", "<i8")])
R32 = np.dtype([("key", "<u8"), ("val", "<u8"), ("time", "<u8"), ("diff", "<i8")])
R40 = np.dtype([("key", "<u8"), ("val1", "<u8"), ("val2", "<u8"), ("time", "<u8"), ("diff", "<i8")])
RACC = np.dtype(
    [
        ("key", "<u8"),
        ("time", "<u8"),
        ("total", "<i8"),
        ("non_nulls", "<i8"),
        ("acc_lo", "<u8"),
        ("acc_hi", "<i8"),
        ("pos_infs", "<i8"),
        ("neg_infs", "<i8"),
        ("nans", "<i8"),
        ("_pad", "<i8"),
    ]
)
ROUT = np.dtype(
    [
        ("key", "<u8"),
        ("count", "<i8"),
        ("sum_lo", "<u8"),
        ("sum_hi", "<i8"),
        ("flags", "<u8"),
        ("time", "<u8"),
        ("diff", "<i8"),
        ("_pad", "<i8"),
    ]
)
DTYPES = {16: R16, 32: R32, 40: R40, 80: RACC, 64: ROUT}
FRONTIER_EMPTY = 2**64 - 1


class Field(C.Structure):
    _fields_ = [("src", C.c_uint8), ("shift", C.c_uint8), ("bits", C.c_uint8), ("dst_shift", C.c_uint8)]


class Filter(C.Structure):
    _fields_ = [("field", Field), ("op", C.c_uint32), ("rhs", C.c_uint64)]


class Closure(C.Structure):
    _fields_ = [
        ("n_key_fields", C.c_uint32),
        ("n_val_fields", C.c_uint32),
        ("n_filters", C.c_uint32),
        ("expr_kind", C.c_uint32),
        ("key_fields", Field * 6),
        ("val_fields", Field * 6),
        ("filters", Filter * 4),
        ("expr_a", Field),
        ("expr_b", Field),
        ("expr_c", C.c_uint64),
    ]


SRC_KEY, SRC_VAL1, SRC_VAL2 = 0, 1, 2
CMP = {"eq": 0, "ne": 1, "lt": 2, "le": 3, "gt": 4, "ge": 5}


def make_closure(key_fields=(), val_fields=(), filters=(), expr=None):
    """key_fields/val_fields: (src, shift, bits, dst_shift); filters: (src, shift, bits, op, rhs);
    expr: ((src,shift,bits), (src,shift,bits), c) meaning a * (c - b)."""
    c = Closure()
    c.n_key_fields = len(key_fields)
    c.n_val_fields = len(val_fields)
    c.n_filters = len(filters)
    for i, f in enumerate(key_fields):
        c.key_fields[i] = Field(*f)
    for i, f in enumerate(val_fields):
        c.val_fields[i] = Field(*f)
    for i, (src, shift, bits, op, rhs) in enumerate(filters):
        c.filters[i] = Filter(Field(src, shift, bits, 0), CMP[op], rhs)
    if expr is not None:
        a, b, k = expr
        c.expr_kind = 1
        c.expr_a = Field(a[0], a[1], a[2], 0)
        c.expr_b = Field(b[0], b[1], b[2], 0)
        c.expr_c = k
    return c


def build():
    """Compile the oracle with the committed Makefile (gcc only)."""
    subprocess.check_call(["make", "-s", "-C", HERE])


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int32
        sig = {
            "mzo_vec_new": (vp, [u32]),
            "mzo_vec_free": (None, [vp]),
            "mzo_vec_len": (u64, [vp]),
            "mzo_vec_data": (vp, [vp]),
            "mzo_vec_clear": (None, [vp]),
            "mzo_consolidate_r16": (u64, [vp, u64]),
            "mzo_consolidate_r32": (u64, [vp, u64]),
            "mzo_consolidate_r40": (u64, [vp, u64]),
            "mzo_consolidate_racc": (u64, [vp, u64]),
            "mzo_consolidate_rout": (u64, [vp, u64]),
            "mzo_consolidate_r16_inplace": (u64, [vp, u64]),
            "mzo_merge_chains_r32": (u64, [vp, u64, vp, u64, u64, vp]),
            "mzo_extract_r32": (u64, [vp, u64, u64, vp, vp, vp, vp]),
            "mzo_batcher_new": (vp, [u32]),
            "mzo_batcher_free": (None, [vp]),
            "mzo_batcher_push": (None, [vp, vp, u64]),
            "mzo_batcher_seal": (vp, [vp, u64]),
            "mzo_batcher_frontier": (u64, [vp]),
            "mzo_batcher_len": (u64, [vp]),
            "mzo_batch_build": (vp, [u32, vp, u64, u64, u64, u64]),
            "mzo_batch_free": (None, [vp]),
            "mzo_batch_len": (u64, [vp]),
            "mzo_batch_keys": (u64, [vp]),
            "mzo_batch_desc": (None, [vp, vp]),
            "mzo_batch_export": (None, [vp, vp]),
            "mzo_batch_csr_sizes": (None, [vp, vp]),
            "mzo_batch_merge": (vp, [vp, vp, u64]),
            "mzo_spine_new": (vp, [u32, u32, i32]),
            "mzo_spine_free": (None, [vp]),
            "mzo_spine_insert": (None, [vp, vp]),
            "mzo_spine_exert": (i32, [vp, u64]),
            "mzo_spine_exert_logic": (u64, [vp, u32]),
            "mzo_spine_set_logical_compaction": (None, [vp, u64]),
            "mzo_spine_set_physical_compaction": (None, [vp, u64]),
            "mzo_spine_read_upper": (u64, [vp]),
            "mzo_spine_layers": (u32, [vp, vp, u32]),
            "mzo_spine_num_batches_through": (u32, [vp, u64]),
            "mzo_spine_export": (None, [vp, vp]),
            "mzo_hspine_new": (vp, []),
            "mzo_hspine_free": (None, [vp]),
            "mzo_hspine_push": (None, [vp, u64, u64, u64, u64, C.c_char_p]),
            "mzo_hspine_downgrade_since": (None, [vp, u64]),
            "mzo_hspine_describe": (C.c_char_p, [vp]),
            "mzo_hspine_merge_reqs": (C.c_char_p, [vp]),
            "mzo_hspine_since": (u64, [vp]),
            "mzo_hspine_upper": (u64, [vp]),
            "mzo_join_new": (vp, [vp, vp, vp, i32]),
            "mzo_join_free": (None, [vp]),
            "mzo_join_push": (None, [vp, i32, vp, u64]),
            "mzo_join_work": (i32, [vp, u64, vp]),
            "mzo_half_join": (None, [vp, u64, vp, i32, vp, i32, vp]),
            "mzo_update_stream": (None, [vp, vp, u64, vp]),
            "mzo_map_rows": (None, [vp, u64, vp, vp]),
            "mzo_correction_new": (vp, [C.c_double, u64]),
            "mzo_correction_free": (None, [vp]),
            "mzo_correction_insert": (None, [vp, vp, u64, i32]),
            "mzo_correction_updates_before": (None, [vp, u64, vp]),
            "mzo_correction_advance_since": (None, [vp, u64]),
            "mzo_correction_consolidate_at_since": (None, [vp]),
            "mzo_correction_chains": (u64, [vp, vp, u64, vp]),
            "mzo_reduce_new": (vp, [i32]),
            "mzo_topk_new": (vp, [C.c_int64, u64, i32]),
            "mzo_reduce_free": (None, [vp]),
            "mzo_reduce_step": (None, [vp, vp, u64, u64, vp]),
            "mzo_explode": (None, [vp, u64, i32, vp]),
            "mzo_finalize": (None, [vp, u64, i32, vp]),
            "mzo_route": (u32, [u64, u32]),
            "mzo_col_encode_slices": (None, [vp, vp, u32, vp]),
            "mzo_col_decode_slices": (i32, [vp, u64, vp, u32]),
            "mzo_col_length_in_words": (u64, [vp, u32]),
            "mzo_col_at_capacity": (i32, [u64]),
            "mzo_column_encode": (None, [i32, vp, u64, vp]),
            "mzo_column_rows": (i32, [i32, vp, u64, vp]),
            "mzo_column_builder": (u32, [i32, vp, u64, vp, vp, u32]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _cl(closure):
    return C.cast(C.pointer(closure), C.c_void_p) if closure is not None else None


def rows(dtype, data):
    """Build a structured row array from a list of tuples."""
    return np.array(list(data), dtype=dtype) if len(data) else np.zeros(0, dtype=dtype)


def consolidate(a):
    a = np.ascontiguousarray(a).copy()
    fn = {16: "r16", 32: "r32", 40: "r40", 80: "racc", 64: "rout"}[a.dtype.itemsize]
    n = getattr(lib(), "mzo_consolidate_" + fn)(_ptr(a), len(a))
    return a[:n].copy()


class Vec:
    def __init__(self, row_bytes):
        self.row_bytes = row_bytes
        self.h = lib().mzo_vec_new(row_bytes)

    def array(self):
        n = lib().mzo_vec_len(self.h)
        dt = DTYPES[self.row_bytes]
        if n == 0:
            return np.zeros(0, dtype=dt)
        buf = (C.c_char * (n * self.row_bytes)).from_address(lib().mzo_vec_data(self.h))
        return np.frombuffer(buf, dtype=dt).copy()

    def clear(self):
        lib().mzo_vec_clear(self.h)

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.mzo_vec_free(self.h)
            self.h = None


class Batch:
    def __init__(self, h, row_bytes):
        self.h, self.row_bytes = h, row_bytes

    @staticmethod
    def build(a, lower, upper, since=0):
        a = np.ascontiguousarray(a)
        return Batch(lib().mzo_batch_build(a.dtype.itemsize, _ptr(a), len(a), lower, upper, since), a.dtype.itemsize)

    def __len__(self):
        return lib().mzo_batch_len(self.h)

    def keys(self):
        return lib().mzo_batch_keys(self.h)

    def desc(self):
        d = np.zeros(3, dtype=np.uint64)
        lib().mzo_batch_desc(self.h, _ptr(d))
        return tuple(int(x) for x in d)

    def csr_sizes(self):
        d = np.zeros(3, dtype=np.uint64)
        lib().mzo_batch_csr_sizes(self.h, _ptr(d))
        return tuple(int(x) for x in d)

    def rows(self):
        a = np.zeros(len(self), dtype=DTYPES[self.row_bytes])
        if len(a):
            lib().mzo_batch_export(self.h, _ptr(a))
        return a

    def merge(self, other, since):
        return Batch(lib().mzo_batch_merge(self.h, other.h, since), self.row_bytes)


class Batcher:
    def __init__(self, row_bytes=32):
        self.row_bytes = row_bytes
        self.h = lib().mzo_batcher_new(row_bytes)

    def push(self, a):
        a = np.ascontiguousarray(a)
        assert a.dtype.itemsize == self.row_bytes
        lib().mzo_batcher_push(self.h, _ptr(a), len(a))

    def seal(self, upper):
        return Batch(lib().mzo_batcher_seal(self.h, upper), self.row_bytes)

    def frontier(self):
        return lib().mzo_batcher_frontier(self.h)

    def __len__(self):
        return lib().mzo_batcher_len(self.h)


class Spine:
    def __init__(self, row_bytes=32, effort=1, gate_physical=False):
        self.row_bytes = row_bytes
        self.h = lib().mzo_spine_new(row_bytes, effort, 1 if gate_physical else 0)
        self._keep = []

    def insert(self, batch):
        self._keep.append(batch)
        lib().mzo_spine_insert(self.h, batch.h)

    def exert(self, effort):
        return bool(lib().mzo_spine_exert(self.h, effort))

    def exert_logic(self, prop):
        return lib().mzo_spine_exert_logic(self.h, prop)

    def set_logical_compaction(self, f):
        lib().mzo_spine_set_logical_compaction(self.h, f)

    def set_physical_compaction(self, f):
        lib().mzo_spine_set_physical_compaction(self.h, f)

    def read_upper(self):
        return lib().mzo_spine_read_upper(self.h)

    def layers(self):
        out = np.zeros(4 * 64, dtype=np.uint64)
        n = lib().mzo_spine_layers(self.h, _ptr(out), 64)
        return [tuple(int(x) for x in out[4 * i : 4 * i + 4]) for i in range(n)]

    def num_batches_through(self, upper):
        return lib().mzo_spine_num_batches_through(self.h, upper)

    def export(self):
        v = Vec(self.row_bytes)
        lib().mzo_spine_export(self.h, v.h)
        return v.array()


class HollowSpine:
    def __init__(self):
        self.h = lib().mzo_hspine_new()

    def push(self, lower, upper, since, length, name=""):
        lib().mzo_hspine_push(self.h, lower, upper, since, length, name.encode())

    def downgrade_since(self, since):
        lib().mzo_hspine_downgrade_since(self.h, since)

    def describe(self):
        return lib().mzo_hspine_describe(self.h).decode()

    def merge_reqs(self):
        return lib().mzo_hspine_merge_reqs(self.h).decode()

    def since(self):
        return lib().mzo_hspine_since(self.h)

    def upper(self):
        return lib().mzo_hspine_upper(self.h)


class Join:
    def __init__(self, spine1, spine2, closure=None, strategy=0):
        self.closure = closure
        self.h = lib().mzo_join_new(spine1.h, spine2.h, _cl(closure), strategy)
        self.out = Vec(32 if closure is not None else 40)

    def push(self, side, batch, cap):
        lib().mzo_join_push(self.h, side, batch.h, cap)

    def work(self, fuel_rows=1 << 62):
        return bool(lib().mzo_join_work(self.h, fuel_rows, self.out.h))

    def results(self):
        return self.out.array()


def half_join(stream, spine, cmp_mode, closure=None, consolidate_output=True):
    stream = np.ascontiguousarray(stream)
    v = Vec(32)
    lib().mzo_half_join(_ptr(stream), len(stream), spine.h, cmp_mode, _cl(closure), 1 if consolidate_output else 0, v.h)
    return v.array()


def update_stream(batch, closure=None, skip_time=FRONTIER_EMPTY):
    v = Vec(32)
    lib().mzo_update_stream(batch.h, _cl(closure), skip_time, v.h)
    return v.array()


def map_rows(a, closure):
    a = np.ascontiguousarray(a)
    v = Vec(32)
    lib().mzo_map_rows(_ptr(a), len(a), _cl(closure), v.h)
    return v.array()


class Reduce:
    def __init__(self, agg_kind=0):
        self.h = lib().mzo_reduce_new(agg_kind)

    def step(self, a, upper):
        a = np.ascontiguousarray(a)
        v = Vec(64)
        lib().mzo_reduce_step(self.h, _ptr(a), len(a), upper, v.h)
        return v.array()


class Correction:
    """MV sink correction buffer (oracle CorrectionV2; src/compute/src/sink/correction_v2.rs)."""

    def __init__(self, chain_proportionality=3.0, chunk_capacity=64):
        self.h = lib().mzo_correction_new(chain_proportionality, chunk_capacity)
        self.chunk_capacity = chunk_capacity

    def insert(self, rows, negate=False):
        rows = np.ascontiguousarray(rows)
        lib().mzo_correction_insert(self.h, _ptr(rows), len(rows), 1 if negate else 0)

    def updates_before(self, upper):
        v = Vec(32)
        lib().mzo_correction_updates_before(self.h, upper, v.h)
        return v.array()

    def advance_since(self, since):
        lib().mzo_correction_advance_since(self.h, since)

    def consolidate_at_since(self):
        lib().mzo_correction_consolidate_at_since(self.h)

    def chains(self):
        lens = (C.c_uint64 * 256)()
        staged = C.c_uint64(0)
        n = lib().mzo_correction_chains(self.h, lens, 256, C.byref(staged))
        return [int(lens[i]) for i in range(min(n, 256))], int(staged.value)

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.mzo_correction_free(self.h)
            self.h = None


class TopK(Reduce):
    """TopK per key (oracle ReduceTopK): limit < 0 = none."""

    def __init__(self, limit, offset=0, descending=False):
        self.h = lib().mzo_topk_new(limit, offset, 1 if descending else 0)


def explode(a, agg_kind=0):
    a = np.ascontiguousarray(a)
    out = np.zeros(len(a), dtype=RACC)
    lib().mzo_explode(_ptr(a), len(a), agg_kind, _ptr(out))
    return out


def finalize(acc, agg_kind=0):
    acc = np.ascontiguousarray(acc)
    out = np.zeros(len(acc), dtype=ROUT)
    lib().mzo_finalize(_ptr(acc), len(acc), agg_kind, _ptr(out))
    return out


# ------------------------------------------------------------------ Q3 dataflow
def _q3_sigs():
    L = lib()
    vp, u64, u32, i32, dbl = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int32, C.c_double
    if getattr(L, "_q3_ready", False):
        return L
    L.mzo_q3_new.restype, L.mzo_q3_new.argtypes = vp, [u64, u64, u64, u64, u32, u64]
    L.mzo_q3_free.restype, L.mzo_q3_free.argtypes = None, [vp]
    L.mzo_q3_hydrate.restype, L.mzo_q3_hydrate.argtypes = dbl, [vp, C.POINTER(u64)]
    L.mzo_q3_step.restype, L.mzo_q3_step.argtypes = dbl, [vp, u64, C.POINTER(u64)]
    L.mzo_q3_drain.restype, L.mzo_q3_drain.argtypes = None, [vp, vp]
    L.mzo_q3_input.restype, L.mzo_q3_input.argtypes = u64, [vp, i32, vp, u64]
    L.mzo_gen_cfg1.restype, L.mzo_gen_cfg1.argtypes = None, [u64, u64, u32, vp]
    L.mzo_gen_cfg2.restype, L.mzo_gen_cfg2.argtypes = None, [u64, u64, u64, vp]
    L.mzo_zipf_cdf.restype, L.mzo_zipf_cdf.argtypes = None, [dbl, u64, vp]
    L.mzo_gen_cfg4.restype, L.mzo_gen_cfg4.argtypes = None, [u64, u64, u64, vp, u64, i32, vp]
    L._q3_ready = True
    return L


class Q3:
    """The CPU Q3 delta-join + reduce dataflow (W worker threads)."""

    def __init__(self, seed, n_customer, n_orders, n_part, workers=1, per_batch=100):
        self.L = _q3_sigs()
        self.h = self.L.mzo_q3_new(seed, n_customer, n_orders, n_part, workers, per_batch)

    def hydrate(self):
        rows = C.c_uint64(0)
        secs = self.L.mzo_q3_hydrate(self.h, C.byref(rows))
        return secs, rows.value

    def step(self, b):
        rows = C.c_uint64(0)
        secs = self.L.mzo_q3_step(self.h, b, C.byref(rows))
        return secs, rows.value

    def drain(self):
        v = Vec(64)
        self.L.mzo_q3_drain(self.h, v.h)
        return v.array()

    def inputs(self, a):
        n = self.L.mzo_q3_input(self.h, a, None, 0)
        out = np.zeros(n, dtype=R32)
        if n:
            self.L.mzo_q3_input(self.h, a, _ptr(out), n)
        return out

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            self.L.mzo_q3_free(self.h)
            self.h = None


def gen_cfg1(seed, n, key_bits):
    out = np.zeros(n, dtype=R16)
    _q3_sigs().mzo_gen_cfg1(seed, n, key_bits, _ptr(out))
    return out


def gen_cfg2(seed, n, n_keys):
    out = np.zeros(n, dtype=R32)
    _q3_sigs().mzo_gen_cfg2(seed, n, n_keys, _ptr(out))
    return out


def zipf_cdf(theta, n):
    cdf = np.zeros(n, dtype=np.float64)
    _q3_sigs().mzo_zipf_cdf(theta, n, _ptr(cdf))
    return cdf


def gen_cfg4(seed, first, n, cdf, as_f64=False):
    out = np.zeros(n, dtype=R32)
    _q3_sigs().mzo_gen_cfg4(seed, first, n, _ptr(cdf), len(cdf), 1 if as_f64 else 0, _ptr(out))
    return out


# ---- f4: columnar wire format (oracle/mzo_column.hpp)
COLUMN_U64X4, COLUMN_U64X2, COLUMN_ROWROW = 0, 1, 2


def _words(vec_h):
    n = lib().mzo_vec_len(vec_h)
    if n == 0:
        return np.zeros(0, dtype="<u8")
    buf = (C.c_char * (n * 8)).from_address(lib().mzo_vec_data(vec_h))
    return np.frombuffer(buf, dtype="<u8").copy()


def col_encode_slices(slices):
    """indexed::encode over raw byte slices -> words."""
    keep = [np.frombuffer(bytes(s), dtype=np.uint8) if len(s) else np.zeros(0, np.uint8) for s in slices]
    ptrs = (C.c_void_p * len(keep))(*[a.ctypes.data if len(a) else None for a in keep])
    lens = np.array([len(a) for a in keep], dtype="<u8")
    h = lib().mzo_vec_new(8)
    try:
        lib().mzo_col_encode_slices(ptrs, _ptr(lens), len(keep), h)
        return _words(h)
    finally:
        lib().mzo_vec_free(h)


def col_decode_slices(words):
    """indexed::decode -> list of bytes (None if malformed)."""
    words = np.ascontiguousarray(words, dtype="<u8")
    off_len = np.zeros(2 * 64, dtype="<u8")
    k = lib().mzo_col_decode_slices(_ptr(words), len(words), _ptr(off_len), 64)
    if k < 0:
        return None
    raw = words.tobytes()
    return [raw[int(off_len[2 * i]) : int(off_len[2 * i] + off_len[2 * i + 1])] for i in range(k)]


def col_length_in_words(lens):
    lens = np.array(list(lens), dtype="<u8")
    return lib().mzo_col_length_in_words(_ptr(lens), len(lens))


def col_at_capacity(words):
    return bool(lib().mzo_col_at_capacity(words))


def column_encode(layout, a):
    """One serialized container holding the R32 rows `a` (U64X2 uses key and diff only)."""
    a = np.ascontiguousarray(a, dtype=R32)
    h = lib().mzo_vec_new(8)
    try:
        lib().mzo_column_encode(layout, _ptr(a), len(a), h)
        return _words(h)
    finally:
        lib().mzo_vec_free(h)


def column_rows(layout, words):
    """The updates of a serialized container as R32 rows; raises ValueError (malformed) or
    NotImplementedError (a Row longer than 7 bytes)."""
    words = np.ascontiguousarray(words, dtype="<u8")
    v = Vec(32)
    rc = lib().mzo_column_rows(layout, _ptr(words), len(words), v.h)
    if rc == -1:
        raise ValueError("malformed container")
    if rc == -2:
        raise NotImplementedError("Row longer than 7 bytes")
    return v.array()


def column_builder(layout, a):
    """ColumnBuilder over the rows: list of serialized containers in order."""
    a = np.ascontiguousarray(a, dtype=R32)
    h = lib().mzo_vec_new(8)
    try:
        sizes = np.zeros(4096, dtype="<u8")
        k = lib().mzo_column_builder(layout, _ptr(a), len(a), h, _ptr(sizes), len(sizes))
        w = _words(h)
        out, at = [], 0
        for i in range(k):
            out.append(w[at : at + int(sizes[i])])
            at += int(sizes[i])
        return out
    finally:
        lib().mzo_vec_free(h)
