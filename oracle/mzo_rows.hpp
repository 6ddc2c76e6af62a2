// oracle/mzo_rows.hpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the reference's update-row arithmetic: ordering, diff
// addition (Semigroup::plus_equals), zero test, consolidation and closure
// evaluation over the fixed-width rows of include/mzgpu.h.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` legs may use anything under oracle/.
//
// Parity pin: the consolidation here is checked against the reference's own
// golden vectors (src/timely-util/src/columnar/batcher.rs:903-990,
// src/timely-util/src/columnar/consolidate.rs:296-385, src/ore/src/iter.rs:260-279)
// in tests/test_oracle_golden.py.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../include/mzgpu.h"

namespace mzo {

typedef uint64_t u64;
typedef int64_t i64;
typedef unsigned __int128 u128;
typedef __int128 i128;

// Overflowing<i64> in release mode wraps (src/ore/src/overflowing.rs:24-35).
static inline i64 wadd(i64 a, i64 b) { return (i64)((u64)a + (u64)b); }
static inline i64 wmul(i64 a, i64 b) { return (i64)((u64)a * (u64)b); }

// ---------------------------------------------------------------- traits
template <class R>
struct Tr;

template <>
struct Tr<mzgpu_r16> {
  typedef mzgpu_r16 R;
  static bool less(const R& a, const R& b) { return a.key < b.key; }
  static bool same(const R& a, const R& b) { return a.key == b.key; }
  static void add(R& a, const R& b) { a.diff = wadd(a.diff, b.diff); }
  static bool zero(const R& a) { return a.diff == 0; }
  static u64 time(const R&) { return 0; }
  static void set_time(R&, u64) {}
  static u64 key(const R& a) { return a.key; }
};

template <>
struct Tr<mzgpu_r32> {
  typedef mzgpu_r32 R;
  static bool less(const R& a, const R& b) {
    if (a.key != b.key) return a.key < b.key;
    if (a.val != b.val) return a.val < b.val;
    return a.time < b.time;
  }
  static bool same(const R& a, const R& b) {
    return a.key == b.key && a.val == b.val && a.time == b.time;
  }
  static void add(R& a, const R& b) { a.diff = wadd(a.diff, b.diff); }
  static bool zero(const R& a) { return a.diff == 0; }
  static u64 time(const R& a) { return a.time; }
  static void set_time(R& a, u64 t) { a.time = t; }
  static u64 key(const R& a) { return a.key; }
};

template <>
struct Tr<mzgpu_r40> {
  typedef mzgpu_r40 R;
  static bool less(const R& a, const R& b) {
    if (a.key != b.key) return a.key < b.key;
    if (a.val1 != b.val1) return a.val1 < b.val1;
    if (a.val2 != b.val2) return a.val2 < b.val2;
    return a.time < b.time;
  }
  static bool same(const R& a, const R& b) {
    return a.key == b.key && a.val1 == b.val1 && a.val2 == b.val2 && a.time == b.time;
  }
  static void add(R& a, const R& b) { a.diff = wadd(a.diff, b.diff); }
  static bool zero(const R& a) { return a.diff == 0; }
  static u64 time(const R& a) { return a.time; }
  static void set_time(R& a, u64 t) { a.time = t; }
  static u64 key(const R& a) { return a.key; }
};

// (Vec<Accum>, Diff) with one accumulated column; Semigroup::plus_equals of
// Accum::SimpleNumber / Accum::Float (src/compute/src/render/reduce.rs:1940-2041):
// component-wise, i128 accumulators wrap.
template <>
struct Tr<mzgpu_racc> {
  typedef mzgpu_racc R;
  static bool less(const R& a, const R& b) {
    if (a.key != b.key) return a.key < b.key;
    return a.time < b.time;
  }
  static bool same(const R& a, const R& b) { return a.key == b.key && a.time == b.time; }
  static void add(R& a, const R& b) {
    a.total = wadd(a.total, b.total);
    a.non_nulls = wadd(a.non_nulls, b.non_nulls);
    u128 x = ((u128)(u64)a.acc_hi << 64) | a.acc_lo;
    u128 y = ((u128)(u64)b.acc_hi << 64) | b.acc_lo;
    x += y;
    a.acc_lo = (u64)x;
    a.acc_hi = (i64)(u64)(x >> 64);
    a.pos_infs = wadd(a.pos_infs, b.pos_infs);
    a.neg_infs = wadd(a.neg_infs, b.neg_infs);
    a.nans = wadd(a.nans, b.nans);
  }
  // IsZero for (Vec<Accum>, Diff): every component zero (reduce.rs:1905-1938).
  static bool zero(const R& a) {
    return a.total == 0 && a.non_nulls == 0 && a.acc_lo == 0 && a.acc_hi == 0 &&
           a.pos_infs == 0 && a.neg_infs == 0 && a.nans == 0;
  }
  static u64 time(const R& a) { return a.time; }
  static void set_time(R& a, u64 t) { a.time = t; }
  static u64 key(const R& a) { return a.key; }
};

template <>
struct Tr<mzgpu_rout> {
  typedef mzgpu_rout R;
  static bool less(const R& a, const R& b) {
    if (a.key != b.key) return a.key < b.key;
    if (a.count != b.count) return (u64)a.count < (u64)b.count;
    if (a.sum_lo != b.sum_lo) return a.sum_lo < b.sum_lo;
    if (a.sum_hi != b.sum_hi) return (u64)a.sum_hi < (u64)b.sum_hi;
    if (a.flags != b.flags) return a.flags < b.flags;
    return a.time < b.time;
  }
  static bool same(const R& a, const R& b) {
    return a.key == b.key && a.count == b.count && a.sum_lo == b.sum_lo && a.sum_hi == b.sum_hi &&
           a.flags == b.flags && a.time == b.time;
  }
  static void add(R& a, const R& b) { a.diff = wadd(a.diff, b.diff); }
  static bool zero(const R& a) { return a.diff == 0; }
  static u64 time(const R& a) { return a.time; }
  static void set_time(R& a, u64 t) { a.time = t; }
  static u64 key(const R& a) { return a.key; }
};

// --------------------------------------------------------- consolidation
// differential_dataflow::consolidation::consolidate_updates (0.23.0, external
// crate): sort by (data, time), sum diffs of equal neighbours, drop zeros,
// truncate.  Restated from the in-tree equivalents
// src/timely-util/src/columnar/batcher.rs:74-121 (sort + fold + zero drop) and
// the reference model :1116-1130.  `consolidate_from(v, off)` only touches
// v[off..] (src/compute/src/render/join/mz_join_core.rs:828).
template <class R>
void consolidate_from(std::vector<R>& v, size_t off) {
  if (v.size() <= off) return;
  std::sort(v.begin() + off, v.end(), [](const R& a, const R& b) { return Tr<R>::less(a, b); });
  size_t w = off;
  size_t i = off;
  const size_t n = v.size();
  while (i < n) {
    R acc = v[i];
    size_t j = i + 1;
    while (j < n && Tr<R>::same(acc, v[j])) {
      Tr<R>::add(acc, v[j]);
      ++j;
    }
    if (!Tr<R>::zero(acc)) v[w++] = acc;
    i = j;
  }
  v.resize(w);
}

template <class R>
void consolidate(std::vector<R>& v) {
  consolidate_from(v, 0);
}

// ------------------------------------------------------ closure evaluation
// Fixed-width stand-in for JoinClosure::apply
// (src/compute-types/src/plan/join.rs:50-82): equality/range filters on
// columns + projection, over bit-field columns.  Same descriptor semantics as
// documented in include/mzgpu.h; the CUDA path evaluates the same struct.
static inline u64 field_get(const mzgpu_field& f, u64 key, u64 v1, u64 v2) {
  u64 w = f.src == MZGPU_SRC_KEY ? key : (f.src == MZGPU_SRC_VAL1 ? v1 : v2);
  w >>= f.shift;
  if (f.bits < 64) w &= ((u64)1 << f.bits) - 1;
  return w;
}

static inline bool closure_apply(const mzgpu_closure* c, u64 key, u64 v1, u64 v2, u64* okey,
                                 u64* oval) {
  for (uint32_t i = 0; i < c->n_filters; ++i) {
    const mzgpu_filter& f = c->filters[i];
    u64 x = field_get(f.field, key, v1, v2);
    bool ok;
    switch (f.op) {
      case MZGPU_CMP_EQ: ok = x == f.rhs; break;
      case MZGPU_CMP_NE: ok = x != f.rhs; break;
      case MZGPU_CMP_LT: ok = x < f.rhs; break;
      case MZGPU_CMP_LE: ok = x <= f.rhs; break;
      case MZGPU_CMP_GT: ok = x > f.rhs; break;
      default: ok = x >= f.rhs; break;
    }
    if (!ok) return false;
  }
  u64 k = 0, v = 0;
  for (uint32_t i = 0; i < c->n_key_fields; ++i)
    k |= field_get(c->key_fields[i], key, v1, v2) << c->key_fields[i].dst_shift;
  if (c->expr_kind == MZGPU_EXPR_MUL_CONST_MINUS) {
    u64 a = field_get(c->expr_a, key, v1, v2);
    u64 b = field_get(c->expr_b, key, v1, v2);
    v = (u64)wmul((i64)a, (i64)(c->expr_c - b));
  } else {
    for (uint32_t i = 0; i < c->n_val_fields; ++i)
      v |= field_get(c->val_fields[i], key, v1, v2) << c->val_fields[i].dst_shift;
  }
  *okey = k;
  *oval = v;
  return true;
}

// Hashable::hashed for u64 keys in DD 0.23 = FNV-1a 64 over the 8 LE bytes
// (SURVEY.md Appendix B); timely's Exchange routes by hash % peers
// (src/compute/src/extensions/arrange.rs:116).  The output collections do not
// depend on the routing function; the CUDA path uses the same one so that
// per-worker shards are comparable one to one.
static inline u64 fnv1a64(u64 key) {
  u64 h = 0xcbf29ce484222325ull;
  for (int i = 0; i < 8; ++i) {
    h ^= (key >> (8 * i)) & 0xff;
    h *= 0x100000001b3ull;
  }
  return h;
}
static inline uint32_t route(u64 key, uint32_t peers) { return (uint32_t)(fnv1a64(key) % peers); }

}  // namespace mzo
