// reduce.cu — accumulable reduce: COUNT / SUM moved into the diff.
//
// Reference (src/compute/src/render/reduce.rs):
//   explode_one + datum_to_accumulator   :1313-1334, 1530-1669
//   Multiply<Diff> for Accum             :2043-2104
//   Semigroup for Accum (i128 wrapping)  :1940-2041
//   reduce_abelian closure + finalize    :1388-1409, 1671-1835
//   AccumulableErrorCheck                :1410-1466
//   FLOAT_SCALE = 2^24                   :1528
// and the reduce operator contract src/compute/src/extensions/reduce.rs:52-107.
//
// GPU shape: values become 80-byte accumulator rows (k_explode), the generic
// sort + segmented sum arranges them by (key, time), and k_corrections walks
// each changed key once: it sums the key's history from the prior batches of
// the arrangement (one hash probe per batch), then replays the new batch's
// times in order, emitting (-old, +new) output rows whenever the finalized
// aggregate changes.  All arithmetic is integer (i64 / i128 with carries), so
// results do not depend on summation order and match the reference bit for bit.
#include "common.cuh"

namespace {

constexpr int RT = 256;

// (x * 2^24) as i128 with Rust's saturating float->int cast semantics.
__device__ __forceinline__ void f64_to_i128_sat(double x, u64* lo, u64* hi) {
  if (isnan(x)) {
    *lo = 0;
    *hi = 0;
    return;
  }
  const bool neg = x < 0.0;
  const double ax = fabs(x);
  if (ax < 9223372036854775808.0) {  // < 2^63: exact through i64 (truncation toward zero)
    long long v = (long long)x;
    *lo = (u64)v;
    *hi = v < 0 ? ~0ull : 0ull;
    return;
  }
  if (ax >= 170141183460469231731687303715884105728.0) {  // >= 2^127 (or inf): saturate
    if (neg) {
      *lo = 0;
      *hi = 0x8000000000000000ull;
    } else {
      *lo = ~0ull;
      *hi = 0x7fffffffffffffffull;
    }
    return;
  }
  const u64 bits = (u64)__double_as_longlong(ax);
  const int e = (int)((bits >> 52) & 0x7ff) - 1023;  // 63..126
  const u64 mant = (bits & 0xfffffffffffffull) | (1ull << 52);
  const int sh = e - 52;  // 11..74
  u64 l, h;
  if (sh >= 64) {
    l = 0;
    h = mant << (sh - 64);
  } else {
    l = mant << sh;
    h = mant >> (64 - sh);
  }
  if (neg) {  // two's complement negate
    l = ~l + 1;
    h = ~h + (l == 0 ? 1 : 0);
  }
  *lo = l;
  *hi = h;
}

// (i128 as f64): round to nearest, ties to even.
__device__ __forceinline__ double i128_to_f64(u64 lo, u64 hi) {
  const bool neg = (i64)hi < 0;
  if (neg) {
    lo = ~lo + 1;
    hi = ~hi + (lo == 0 ? 1 : 0);
  }
  double r;
  if (hi == 0) {
    r = __ull2double_rn(lo);
  } else {
    const int lz = __clzll((long long)hi);
    // top 64 bits of the 128-bit magnitude, sticky bit folded into bit 0
    const int shift = 64 - lz;  // bits shifted out of `lo`
    u64 top = lz == 0 ? hi : ((hi << lz) | (lo >> shift));
    u64 lost = lz == 0 ? lo : (lo << lz);
    if (lost) top |= 1;
    r = ldexp(__ull2double_rn(top), shift);
  }
  return neg ? -r : r;
}

// wrapping i128 * i64
__device__ __forceinline__ void mul_i128_i64(u64 lo, u64 hi, i64 d, u64* rlo, u64* rhi) {
  const u64 dl = (u64)d;
  const u64 dh = d < 0 ? ~0ull : 0ull;
  *rlo = lo * dl;
  *rhi = __umul64hi(lo, dl) + hi * dl + lo * dh;
}

__global__ void __launch_bounds__(RT) k_explode(const u64* __restrict__ rows, const DLen dn, int agg_kind,
                                                u64* __restrict__ out) {
  const u64 n = dlen_get(dn);
  for (u64 i = (u64)blockIdx.x * RT + threadIdx.x; i < n; i += (u64)gridDim.x * RT) {
  u64 r[4];
  load_row<4>(rows, i, r);
  const i64 diff = (i64)r[3];
  u64 o[10];
  o[0] = r[0];       // key
  o[1] = r[2];       // time
  o[2] = (u64)diff;  // total
  if (agg_kind == MZGPU_AGG_DISTINCT || agg_kind == MZGPU_AGG_THRESHOLD) {
    // only the multiplicity matters (build_distinct / threshold_arrangement)
#pragma unroll
    for (int w = 3; w < 10; ++w) o[w] = 0;
    store_row<10>(out, i, o);
    continue;
  }
  o[3] = (u64)diff;  // non_nulls
  u64 alo, ahi;
  u64 pinf = 0, ninf = 0, nan = 0;
  if (agg_kind == MZGPU_AGG_COUNT_SUM_F64) {
    const double v = __longlong_as_double((long long)r[1]);
    const bool is_nan = isnan(v);
    const bool is_pinf = isinf(v) && v > 0;
    const bool is_ninf = isinf(v) && v < 0;
    nan = is_nan ? (u64)diff : 0;
    pinf = is_pinf ? (u64)diff : 0;
    ninf = is_ninf ? (u64)diff : 0;
    if (is_nan || is_pinf || is_ninf) {
      alo = 0;
      ahi = 0;
    } else {
      f64_to_i128_sat(v * 16777216.0, &alo, &ahi);
    }
  } else {
    alo = r[1];
    ahi = (i64)r[1] < 0 ? ~0ull : 0ull;
  }
  mul_i128_i64(alo, ahi, diff, &o[4], &o[5]);
  o[6] = pinf;
  o[7] = ninf;
  o[8] = nan;
  o[9] = 0;
  store_row<10>(out, i, o);
  }
}

// finalize_accum + error-check flag for one accumulated diff S (words 2..8 of a RACC row)
__device__ __forceinline__ void finalize(const u64* S, int agg_kind, u64* o /* count, sum_lo, sum_hi, flags */) {
  const i64 total = (i64)S[0];
  if (agg_kind == MZGPU_AGG_DISTINCT) {  // (key, ()) once; error flag for a negative multiplicity
    o[0] = 1;
    o[1] = 0;
    o[2] = 0;
    o[3] = total < 0 ? 2 : 0;
    return;
  }
  const bool accum_zero = (S[1] | S[2] | S[3] | S[4] | S[5] | S[6]) == 0;
  u64 flags = 0;
  if (total > 0 && accum_zero) flags |= 1;
  if (total == 0 && !accum_zero) flags |= 2;
  o[0] = S[1];
  if (agg_kind == MZGPU_AGG_COUNT_SUM_F64) {
    const i64 pinf = (i64)S[4], ninf = (i64)S[5], nan = (i64)S[6];
    u64 bits;
    if (nan > 0 || (pinf > 0 && ninf > 0))
      bits = 0x7ff8000000000000ull;
    else if (pinf > 0)
      bits = 0x7ff0000000000000ull;
    else if (ninf > 0)
      bits = 0xfff0000000000000ull;
    else
      bits = (u64)__double_as_longlong(i128_to_f64(S[2], S[3]) / 16777216.0);
    o[1] = bits;
    o[2] = 0;
  } else {
    o[1] = S[2];
    o[2] = S[3];
  }
  if (flags & 1) {
    o[1] = 0;
    o[2] = 0;
  }
  o[3] = flags;
}

// sum of all prior updates of `key` (times before the new batch).  The first hash
// slot of GROUP batches is fetched before any is inspected (independent loads).
__device__ __forceinline__ void prior_sum(const TraceView& tv, u64 key, u64* S) {
  const u64 h0 = mix64(key);
  constexpr int GROUP = 8;
  for (u32 b0 = 0; b0 < tv.n_batches; b0 += GROUP) {
    ulonglong2 slot[GROUP];
    u64 hh[GROUP], mask[GROUP];
#pragma unroll
    for (int j = 0; j < GROUP; ++j) {
      if (b0 + j < tv.n_batches) {
        const BatchView& bv = tv.b[b0 + j];
        mask[j] = bv_mask(bv);
        hh[j] = h0 & mask[j];
        slot[j] = *reinterpret_cast<const ulonglong2*>(&bv.table[hh[j]]);
      }
    }
#pragma unroll
    for (int j = 0; j < GROUP; ++j) {
      if (b0 + j >= tv.n_batches) break;
      const BatchView& bv = tv.b[b0 + j];
      ulonglong2 sl = slot[j];
      u64 h = hh[j];
      while (true) {
        if (sl.y == 0) break;
        if (sl.x == key) {
          const u64 first = (sl.y & MZ_SLOT_ROW_MASK) - 1;
          const u32 len = (u32)(sl.y >> 44);
          const u64 bn = len != 0 ? first + len : bv_n(bv);
          for (u64 r = first; r < bn; ++r) {
            const u64* row = bv.rows + r * 10;
            if (len == 0 && row[0] != key) break;
            u64 d[8];
#pragma unroll
            for (int w = 0; w < 8; ++w) d[w] = row[2 + w];
            diff_add<8>(S, d);
          }
          break;
        }
        h = (h + 1) & mask[j];
        sl = *reinterpret_cast<const ulonglong2*>(&bv.table[h]);
      }
    }
  }
}

// The corrections of ONE key, out[pos .. pos + c), put into consolidated order (words 1..5:
// count, sum_lo, sum_hi, flags, time; the key is the same).  A key's corrections are distinct rows
// (-old and +new differ in their values, different times differ in the time word), and keys are
// written in ascending order by construction, so with this the kernel's whole output is already
// what consolidate() would return: the separate sort launch (40 us for a few hundred rows) is gone.
__device__ __forceinline__ void sort_key_corrections(u64* __restrict__ out, u64 pos, u32 c) {
  for (u32 x = 1; x < c; ++x) {
    u64 r[8];
    load_row<8>(out, pos + x, r);
    u32 y = x;
    while (y > 0) {
      u64 p[8];
      load_row<8>(out, pos + y - 1, p);
      bool less = false;
#pragma unroll
      for (int w = 1; w <= 5; ++w) {
        if (r[w] != p[w]) {
          less = r[w] < p[w];
          break;
        }
      }
      if (!less) break;
      store_row<8>(out, pos + y, p);
      --y;
    }
    if (y != x) store_row<8>(out, pos + y, r);
  }
}

// Corrections of one changed key: rows [i, ...) of the new batch with this key,
// given the key's prior accumulation S0.  Counts (and optionally writes at
// out[pos...]) the (-old, +new) output rows.
__device__ __forceinline__ u32 walk_key(const u64* __restrict__ rows, u64 n, u64 i, u64 key, const u64* S0,
                                        int agg_kind, bool do_write, u64* __restrict__ out, u64 pos) {
  u64 S[8];
#pragma unroll
  for (int w = 0; w < 8; ++w) S[w] = S0[w];
  if (agg_kind == MZGPU_AGG_THRESHOLD) {
    // output multiplicity = max(accumulated multiplicity, 0); one row per change, diff = the change
    i64 mult = (i64)S[0] > 0 ? (i64)S[0] : 0;
    u32 c = 0;
    for (u64 j = i; j < n; ++j) {
      const u64* row = rows + j * 10;
      if (row[0] != key) break;
      S[0] += row[2];
      const i64 m2 = (i64)S[0] > 0 ? (i64)S[0] : 0;
      if (m2 != mult) {
        if (do_write) {
          u64 r[8] = {key, 0, 0, 0, 0, row[1], (u64)(m2 - mult), 0};
          store_row<8>(out, pos + c, r);
        }
        ++c;
      }
      mult = m2;
    }
    if (do_write && c > 1) sort_key_corrections(out, pos, c);
    return c;
  }
  bool had = !diff_is_zero<8>(S);
  u64 oldv[4] = {0, 0, 0, 0};
  if (had) finalize(S, agg_kind, oldv);
  u32 c = 0;
  for (u64 j = i; j < n; ++j) {
    const u64* row = rows + j * 10;
    if (row[0] != key) break;
    u64 d[8];
#pragma unroll
    for (int w = 0; w < 8; ++w) d[w] = row[2 + w];
    diff_add<8>(S, d);
    const u64 t = row[1];
    const bool has = !diff_is_zero<8>(S);
    u64 newv[4] = {0, 0, 0, 0};
    if (has) finalize(S, agg_kind, newv);
    const bool same = had && has && oldv[0] == newv[0] && oldv[1] == newv[1] && oldv[2] == newv[2] &&
                      oldv[3] == newv[3];
    if (!same) {
      if (had) {
        if (do_write) {
          u64 r[8] = {key, oldv[0], oldv[1], oldv[2], oldv[3], t, ~0ull, 0};
          store_row<8>(out, pos + c, r);
        }
        ++c;
      }
      if (has) {
        if (do_write) {
          u64 r[8] = {key, newv[0], newv[1], newv[2], newv[3], t, 1, 0};
          store_row<8>(out, pos + c, r);
        }
        ++c;
      }
    }
    had = has;
#pragma unroll
    for (int w = 0; w < 4; ++w) oldv[w] = newv[w];
  }
  if (do_write && c > 1) sort_key_corrections(out, pos, c);
  return c;
}

template <bool WRITE>
__global__ void __launch_bounds__(RT) k_corrections(const u64* __restrict__ rows, u64 n,
                                                    const __grid_constant__ TraceView prior, int agg_kind,
                                                    u32* __restrict__ tile_counts,
                                                    const u32* __restrict__ tile_base,
                                                    u64* __restrict__ out) {
  __shared__ u32 sm[34];
  const u64 i = (u64)blockIdx.x * RT + threadIdx.x;
  u32 cnt = 0;
  const bool head = i < n && (i == 0 || rows[(i - 1) * 10] != rows[i * 10]);
  u64 key = 0;
  u64 S0[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (head) {
    key = rows[i * 10];
    prior_sum(prior, key, S0);
    cnt = walk_key(rows, n, i, key, S0, agg_kind, false, nullptr, 0);
  }
  u32 total;
  u32 ex = block_exclusive_scan(cnt, sm, &total);
  if (!WRITE) {
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = total;
  } else {
    if (head && cnt > 0) walk_key(rows, n, i, key, S0, agg_kind, true, out, (u64)tile_base[blockIdx.x] + ex);
  }
}

// single-pass form (sizes on the device, chained tiles): see probe.cu
__global__ void __launch_bounds__(RT) k_corrections_lb(const u64* __restrict__ rows, const DLen dn,
                                                       const __grid_constant__ TraceView prior, int agg_kind,
                                                       const LookBack lb, u64* __restrict__ out, u64 out_cap,
                                                       u64* __restrict__ out_len, u64* __restrict__ status) {
  __shared__ u32 sm[34];
  __shared__ u32 s_tile;
  __shared__ u64 s_b;
  const u64 n = dlen_get(dn);
  const u64 n_tiles = (n + RT - 1) / RT;
  while (true) {
    const u32 tile = lb_next_tile(lb, &s_tile);
    if ((u64)tile >= n_tiles) {
      if (n_tiles == 0 && tile == 0 && threadIdx.x == 0) *out_len = 0;
      break;
    }
    const u64 i = (u64)tile * RT + threadIdx.x;
    u32 cnt = 0;
    const bool head = i < n && (i == 0 || rows[(i - 1) * 10] != rows[i * 10]);
    u64 key = 0;
    u64 S0[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (head) {
      key = rows[i * 10];
      prior_sum(prior, key, S0);
      cnt = walk_key(rows, n, i, key, S0, agg_kind, false, nullptr, 0);
    }
    u32 total;
    const u32 ex = block_exclusive_scan(cnt, sm, &total);
    const u64 excl = lb_exclusive_prefix(lb, tile, (u64)total, &s_b);
    if (head && cnt > 0) {
      const u64 pos = excl + ex;
      if (pos + cnt > out_cap)
        atomicMax((unsigned long long*)status, (unsigned long long)(pos + cnt));
      else
        walk_key(rows, n, i, key, S0, agg_kind, true, out, pos);
    }
    if ((u64)tile == n_tiles - 1 && threadIdx.x == 0) *out_len = excl + total;
  }
}

// ------------------------------------------------------------------ MIN / MAX
// Result of the hierarchical reduce (reduce.rs:1050-1135): per key the live
// (value, count) pairs; a negative count -> the error row, else MIN / MAX of the
// values.  One thread per changed key keeps the key's pairs in a small local table.
constexpr int MM_CAP = 32;
struct MinMaxAcc {
  u64 vals[MM_CAP];
  i64 cnts[MM_CAP];
  int m;
  bool overflow;
};
__device__ __forceinline__ void mm_add(MinMaxAcc& a, u64 val, i64 d) {
  for (int j = 0; j < a.m; ++j)
    if (a.vals[j] == val) {
      a.cnts[j] += d;
      return;
    }
  // reuse a dead entry before growing
  for (int j = 0; j < a.m; ++j)
    if (a.cnts[j] == 0) {
      a.vals[j] = val;
      a.cnts[j] = d;
      return;
    }
  if (a.m < MM_CAP) {
    a.vals[a.m] = val;
    a.cnts[a.m] = d;
    ++a.m;
  } else {
    a.overflow = true;
  }
}
__device__ __forceinline__ bool mm_eval(const MinMaxAcc& a, int agg_kind, u64* o /* count, sum_lo, sum_hi, flags */) {
  bool any = false, bad = false, have = false;
  u64 best = 0;
  for (int j = 0; j < a.m; ++j) {
    const i64 c = a.cnts[j];
    if (c == 0) continue;
    any = true;
    if (c < 0) {
      bad = true;
      continue;
    }
    const u64 v = a.vals[j];
    if (!have || (agg_kind == MZGPU_AGG_MIN ? v < best : v > best)) best = v;
    have = true;
  }
  o[0] = 0;
  o[1] = bad ? 0 : best;
  o[2] = 0;
  o[3] = bad ? 2 : 0;
  return any;
}
__device__ __forceinline__ void mm_prior(const TraceView& tv, u64 key, MinMaxAcc& a) {
  const u64 h0 = mix64(key);
  for (u32 b = 0; b < tv.n_batches; ++b) {
    const BatchView& bv = tv.b[b];
    const u64 mask = bv_mask(bv);
    u64 h = h0 & mask;
    while (true) {
      const ulonglong2 sl = *reinterpret_cast<const ulonglong2*>(&bv.table[h]);
      if (sl.y == 0) break;
      if (sl.x == key) {
        const u64 first = (sl.y & MZ_SLOT_ROW_MASK) - 1;
        const u32 len = (u32)(sl.y >> 44);
        const u64 bn = len != 0 ? first + len : bv_n(bv);
        for (u64 r = first; r < bn; ++r) {
          const ulonglong2* row = reinterpret_cast<const ulonglong2*>(bv.rows + r * 4);
          const ulonglong2 kv = row[0];
          if (len == 0 && kv.x != key) break;
          mm_add(a, kv.y, (i64)row[1].y);
        }
        break;
      }
      h = (h + 1) & mask;
    }
  }
}
// ---- groups wider than the local table (more than MM_CAP distinct live values of one key).
// The reference bounds the work per update with a tree of hashed buckets (build_bucketed,
// reduce.rs:796-900; top_k.rs:251-380); what the tree COMPUTES is the same per-key function of
// the key's (value, count) pairs.  Every batch holds a key's updates as a run sorted by (value,
// time), so that function is a k-way merge of the runs in value order, one value at a time --
// any group width, no table.  Only a key that overflowed the table takes this path.
struct RunCur {
  const u64* rows;
  u64 lo, hi;  // rows [lo, hi) of the key, sorted by (val, time)
};
constexpr int MM_MAX_RUNS = MZ_MAX_TRACE_BATCHES + 1;
__device__ __noinline__ int mm_runs(const TraceView& tv, u64 key, RunCur* cur) {
  int nc = 0;
  const u64 h0 = mix64(key);
  for (u32 b = 0; b < tv.n_batches; ++b) {
    const BatchView& bv = tv.b[b];
    const u64 mask = bv_mask(bv);
    u64 h = h0 & mask;
    while (true) {
      const ulonglong2 sl = *reinterpret_cast<const ulonglong2*>(&bv.table[h]);
      if (sl.y == 0) break;
      if (sl.x == key) {
        const u64 first = (sl.y & MZ_SLOT_ROW_MASK) - 1;
        const u32 len = (u32)(sl.y >> 44);
        u64 end = first + len;
        if (len == 0) {  // run length not recorded: upper bound search (rows are sorted by key)
          u64 lo = first + 1, hi = bv_n(bv);
          while (lo < hi) {
            const u64 mid = (lo + hi) >> 1;
            if (bv.rows[mid * 4] == key)
              lo = mid + 1;
            else
              hi = mid;
          }
          end = lo;
        }
        cur[nc].rows = bv.rows;
        cur[nc].lo = first;
        cur[nc].hi = end;
        ++nc;
        break;
      }
      h = (h + 1) & mask;
    }
  }
  return nc;
}
// Visit the key's distinct values in ascending (or descending) order with their total count:
// all rows of the prior runs, and the rows of the new batch's run [nlo, nhi) whose time is
// <= t_limit (use_new).  f(value, count) returns false to stop early.  The cursors are copied:
// `cur` is left untouched.
template <class F>
__device__ __forceinline__ void mm_stream(const RunCur* cur, int nc, const u64* nrows, u64 nlo, u64 nhi, bool use_new,
                                          u64 t_limit, bool desc, F f) {
  u64 lo[MM_MAX_RUNS], hi[MM_MAX_RUNS];
  for (int c = 0; c < nc; ++c) {
    lo[c] = cur[c].lo;
    hi[c] = cur[c].hi;
  }
  const int nn = use_new ? nc + 1 : nc;
  if (use_new) {
    lo[nc] = nlo;
    hi[nc] = nhi;
  }
  while (true) {
    bool have = false;
    u64 v = 0;
    for (int c = 0; c < nn; ++c) {
      if (lo[c] >= hi[c]) continue;
      const u64* rows = c < nc ? cur[c].rows : nrows;
      const u64 hv = rows[(desc ? hi[c] - 1 : lo[c]) * 4 + 1];
      if (!have || (desc ? hv > v : hv < v)) {
        v = hv;
        have = true;
      }
    }
    if (!have) break;
    i64 cnt = 0;
    for (int c = 0; c < nn; ++c) {
      const u64* rows = c < nc ? cur[c].rows : nrows;
      while (lo[c] < hi[c]) {
        const u64 r = desc ? hi[c] - 1 : lo[c];
        if (rows[r * 4 + 1] != v) break;
        if (c < nc || rows[r * 4 + 2] <= t_limit) cnt += (i64)rows[r * 4 + 3];
        if (desc)
          --hi[c];
        else
          ++lo[c];
      }
    }
    if (!f(v, cnt)) break;
  }
}
// MIN / MAX of a wide group: one ascending pass (every value is looked at: the error row needs
// to know about any negative count)
__device__ __noinline__ bool mm_eval_stream(const RunCur* cur, int nc, const u64* nrows, u64 nlo, u64 nhi, bool use_new,
                                            u64 t_limit, int agg_kind, u64* o) {
  bool any = false, bad = false, have = false;
  u64 best = 0;
  mm_stream(cur, nc, nrows, nlo, nhi, use_new, t_limit, false, [&](u64 v, i64 c) -> bool {
    if (c == 0) return true;
    any = true;
    if (c < 0) {
      bad = true;
      return true;
    }
    if (!have || agg_kind == MZGPU_AGG_MAX) best = v;  // ascending: first positive = MIN, last = MAX
    have = true;
    return true;
  });
  o[0] = 0;
  o[1] = bad ? 0 : best;
  o[2] = 0;
  o[3] = bad ? 2 : 0;
  return any;
}

// rows [i, ...) of the new batch share `key` (sorted by (val, time)): replay them in
// time order on top of the prior pairs and emit (-old, +new) whenever the result changes.
__device__ __noinline__ u32 walk_minmax(const u64* __restrict__ rows, u64 n, u64 i, u64 key,
                                        const TraceView& prior, int agg_kind, bool do_write,
                                        u64* __restrict__ out, u64 pos, u64* __restrict__ status) {
  MinMaxAcc a;
  a.m = 0;
  a.overflow = false;
  mm_prior(prior, key, a);
  // the key's run in the new batch, and (only if the table overflows) its runs in the prior batches
  u64 i_end = i;
  while (i_end < n && rows[i_end * 4] == key) ++i_end;
  RunCur runs[MM_MAX_RUNS];
  int n_runs = -1;
  auto eval_at = [&](bool use_new, u64 t_limit, u64* o) -> bool {
    if (!a.overflow) return mm_eval(a, agg_kind, o);
    if (n_runs < 0) n_runs = mm_runs(prior, key, runs);
    return mm_eval_stream(runs, n_runs, rows, i, i_end, use_new, t_limit, agg_kind, o);
  };
  u64 oldv[4];
  bool had = eval_at(false, 0, oldv);
  u32 c = 0;
  bool first = true;
  u64 t_prev = 0;
  while (true) {
    // next distinct time of this key
    bool found = false;
    u64 t_cur = 0;
    for (u64 j = i; j < n; ++j) {
      const u64* row = rows + j * 4;
      if (row[0] != key) break;
      const u64 t = row[2];
      if ((first || t > t_prev) && (!found || t < t_cur)) {
        t_cur = t;
        found = true;
      }
    }
    if (!found) break;
    for (u64 j = i; j < n; ++j) {
      const u64* row = rows + j * 4;
      if (row[0] != key) break;
      if (row[2] == t_cur) mm_add(a, row[1], (i64)row[3]);
    }
    u64 newv[4];
    const bool has = eval_at(true, t_cur, newv);
    const bool same = (had == has) && (!has || (oldv[1] == newv[1] && oldv[3] == newv[3]));
    if (!same) {
      if (had) {
        if (do_write) {
          u64 r[8] = {key, oldv[0], oldv[1], oldv[2], oldv[3], t_cur, ~0ull, 0};
          store_row<8>(out, pos + c, r);
        }
        ++c;
      }
      if (has) {
        if (do_write) {
          u64 r[8] = {key, newv[0], newv[1], newv[2], newv[3], t_cur, 1, 0};
          store_row<8>(out, pos + c, r);
        }
        ++c;
      }
    }
    had = has;
#pragma unroll
    for (int w = 0; w < 4; ++w) oldv[w] = newv[w];
    first = false;
    t_prev = t_cur;
  }
  (void)status;
  return c;
}

// ---------------------------------------------------------------------- TopK
// build_topk_negated_stage (top_k.rs:521-673): order the key's live values, skip `offset`
// rows, keep at most `limit` (multiplicities counted); a negative count -> the error row.
struct TopKWin {
  u64 v[MM_CAP];
  i64 m[MM_CAP];
  int n;
  bool err;
};
__device__ __noinline__ void tk_eval(const MinMaxAcc& a, const TopKParams& tp, TopKWin& w) {
  w.n = 0;
  w.err = false;
  for (int j = 0; j < a.m; ++j)
    if (a.cnts[j] < 0) {
      w.err = true;
      return;
    }
  u64 skip = tp.offset;
  i64 left = tp.limit;
  bool has_last = false;
  u64 last = 0;
  while (!(tp.limit >= 0 && left == 0)) {
    // next value in order (selection over at most MM_CAP live entries)
    int best = -1;
    for (int j = 0; j < a.m; ++j) {
      if (a.cnts[j] <= 0) continue;
      const u64 v = a.vals[j];
      const bool after = !has_last || (tp.descending ? v < last : v > last);
      if (after && (best < 0 || (tp.descending ? v > a.vals[best] : v < a.vals[best]))) best = j;
    }
    if (best < 0) break;
    last = a.vals[best];
    has_last = true;
    i64 cnt = a.cnts[best];
    if (skip > 0) {
      const u64 s = skip < (u64)cnt ? skip : (u64)cnt;
      skip -= s;
      cnt -= (i64)s;
    }
    if (tp.limit >= 0) {
      cnt = cnt < left ? cnt : left;
      left -= cnt;
    }
    if (cnt > 0) {
      w.v[w.n] = last;
      w.m[w.n] = cnt;
      ++w.n;
    }
  }
}
// the same window for a group wider than the table: a first pass looks for a negative count (the
// error row), a second walks the values in the plan's order until the limit is spent.  Returns
// false if the WINDOW itself has more than MM_CAP distinct values (limit > MM_CAP or none).
__device__ __noinline__ bool tk_eval_stream(const RunCur* cur, int nc, const u64* nrows, u64 nlo, u64 nhi, bool use_new,
                                            u64 t_limit, const TopKParams& tp, TopKWin& w) {
  w.n = 0;
  w.err = false;
  mm_stream(cur, nc, nrows, nlo, nhi, use_new, t_limit, false, [&](u64, i64 c) -> bool {
    if (c < 0) w.err = true;
    return !w.err;
  });
  if (w.err) return true;
  u64 skip = tp.offset;
  i64 left = tp.limit;
  bool fits = true;
  if (tp.limit == 0) return true;
  mm_stream(cur, nc, nrows, nlo, nhi, use_new, t_limit, tp.descending != 0, [&](u64 v, i64 c) -> bool {
    if (c <= 0) return true;
    i64 cnt = c;
    if (skip > 0) {
      const u64 s_ = skip < (u64)cnt ? skip : (u64)cnt;
      skip -= s_;
      cnt -= (i64)s_;
    }
    if (tp.limit >= 0) {
      cnt = cnt < left ? cnt : left;
      left -= cnt;
    }
    if (cnt > 0) {
      if (w.n == MM_CAP) {
        fits = false;
        return false;
      }
      w.v[w.n] = v;
      w.m[w.n] = cnt;
      ++w.n;
    }
    return !(tp.limit >= 0 && left == 0);
  });
  return fits;
}
// changes old -> fresh at time t; returns the number of rows (written at out[pos...] if do_write)
__device__ __forceinline__ u32 tk_emit(u64 key, const TopKWin& old, const TopKWin& fresh, u64 t, bool do_write,
                                       u64* __restrict__ out, u64 pos) {
  u32 c = 0;
  auto put = [&](u64 val, u64 flags, i64 d) {
    if (do_write) {
      u64 r[8] = {key, 0, val, 0, flags, t, (u64)d, 0};
      store_row<8>(out, pos + c, r);
    }
    ++c;
  };
  if (old.err != fresh.err) put(0, 2, fresh.err ? 1 : -1);
  for (int i = 0; i < old.n; ++i) {
    i64 now = 0;
    for (int j = 0; j < fresh.n; ++j)
      if (fresh.v[j] == old.v[i]) now = fresh.m[j];
    if (now != old.m[i]) put(old.v[i], 0, now - old.m[i]);
  }
  for (int j = 0; j < fresh.n; ++j) {
    bool seen = false;
    for (int i = 0; i < old.n; ++i) seen = seen || old.v[i] == fresh.v[j];
    if (!seen) put(fresh.v[j], 0, fresh.m[j]);
  }
  return c;
}
__device__ __noinline__ u32 walk_topk(const u64* __restrict__ rows, u64 n, u64 i, u64 key,
                                      const TraceView& prior, const TopKParams& tp, bool do_write,
                                      u64* __restrict__ out, u64 pos, u64* __restrict__ status) {
  MinMaxAcc a;
  a.m = 0;
  a.overflow = false;
  mm_prior(prior, key, a);
  u64 i_end = i;
  while (i_end < n && rows[i_end * 4] == key) ++i_end;
  RunCur runs[MM_MAX_RUNS];
  int n_runs = -1;
  bool window_fits = true;
  auto eval_at = [&](bool use_new, u64 t_limit, TopKWin& w) {
    if (!a.overflow) {
      tk_eval(a, tp, w);
      return;
    }
    if (n_runs < 0) n_runs = mm_runs(prior, key, runs);
    if (!tk_eval_stream(runs, n_runs, rows, i, i_end, use_new, t_limit, tp, w)) window_fits = false;
  };
  TopKWin old, fresh;
  eval_at(false, 0, old);
  u32 c = 0;
  bool first = true;
  u64 t_prev = 0;
  while (true) {
    bool found = false;
    u64 t_cur = 0;
    for (u64 j = i; j < n; ++j) {
      const u64* row = rows + j * 4;
      if (row[0] != key) break;
      const u64 t = row[2];
      if ((first || t > t_prev) && (!found || t < t_cur)) {
        t_cur = t;
        found = true;
      }
    }
    if (!found) break;
    for (u64 j = i; j < n; ++j) {
      const u64* row = rows + j * 4;
      if (row[0] != key) break;
      if (row[2] == t_cur) mm_add(a, row[1], (i64)row[3]);
    }
    eval_at(true, t_cur, fresh);
    c += tk_emit(key, old, fresh, t_cur, do_write, out, pos + c);
    old = fresh;
    first = false;
    t_prev = t_cur;
  }
  // (a window of more than MM_CAP distinct values -- LIMIT beyond 32 on a wide group -- is the one
  // shape still outside the subset: reported, never wrong)
  if (!window_fits) status[1] = 1;
  return c;
}

__global__ void __launch_bounds__(RT) k_minmax_lb(const u64* __restrict__ rows, const DLen dn,
                                                  const __grid_constant__ TraceView prior, int agg_kind,
                                                  const TopKParams tp, const LookBack lb,
                                                  u64* __restrict__ out, u64 out_cap,
                                                  u64* __restrict__ out_len, u64* __restrict__ status) {
  __shared__ u32 sm[34];
  __shared__ u32 s_tile;
  __shared__ u64 s_b;
  const u64 n = dlen_get(dn);
  const u64 n_tiles = (n + RT - 1) / RT;
  const bool topk = agg_kind == MZGPU_AGG_TOPK;
  while (true) {
    const u32 tile = lb_next_tile(lb, &s_tile);
    if ((u64)tile >= n_tiles) {
      if (n_tiles == 0 && tile == 0 && threadIdx.x == 0) *out_len = 0;
      break;
    }
    const u64 i = (u64)tile * RT + threadIdx.x;
    u32 cnt = 0;
    const bool head = i < n && (i == 0 || rows[(i - 1) * 4] != rows[i * 4]);
    u64 key = 0;
    if (head) {
      key = rows[i * 4];
      cnt = topk ? walk_topk(rows, n, i, key, prior, tp, false, nullptr, 0, status)
                 : walk_minmax(rows, n, i, key, prior, agg_kind, false, nullptr, 0, status);
    }
    u32 total;
    const u32 ex = block_exclusive_scan(cnt, sm, &total);
    const u64 excl = lb_exclusive_prefix(lb, tile, (u64)total, &s_b);
    if (head && cnt > 0) {
      const u64 pos = excl + ex;
      if (pos + cnt > out_cap)
        atomicMax((unsigned long long*)status, (unsigned long long)(pos + cnt));
      else if (topk)
        walk_topk(rows, n, i, key, prior, tp, true, out, pos, status);
      else
        walk_minmax(rows, n, i, key, prior, agg_kind, true, out, pos, status);
    }
    if ((u64)tile == n_tiles - 1 && threadIdx.x == 0) *out_len = excl + total;
  }
}

}  // namespace

// MIN / MAX / TopK corrections of a sealed R32 batch against the prior R32 arrangement.
// MIN / MAX: at most two output rows per distinct (key, time), capacity 2 * n_ub suffices;
// TopK: at most 2 * min(limit, 32) + 2 (every window entry may change, plus the error row).
int32_t mz_reduce_minmax_async(mzgpu_ctx* ctx, const u64* d_batch_rows, DLen n, u64 n_ub,
                               const TraceView& prior, int agg_kind, const TopKParams& tp, u64* d_out,
                               u64 out_cap, u64* d_out_len) {
  LookBack lb;
  MZ_TRY(mz_lookback_begin(ctx, (n_ub + RT - 1) / RT, &lb));
  u64 grid = (n_ub + RT - 1) / RT;
  if (grid > (u64)ctx->num_sms * 8) grid = (u64)ctx->num_sms * 8;
  if (grid == 0) grid = 1;
  MZ_BYTES(ctx, n.p == nullptr ? n.imm * (32 + 16 + 32 + 128) : 0);
  MZ_LAUNCH(ctx, k_minmax_lb, (unsigned)grid, RT, 0, d_batch_rows, n, prior, agg_kind, tp, lb, d_out, out_cap,
            d_out_len, ctx->d_status);
  return MZGPU_OK;
}

int32_t mz_explode(mzgpu_ctx* ctx, const u64* d_r32, DLen n, u64 n_ub, int agg_kind, u64* d_racc) {
  if (n_ub == 0) return MZGPU_OK;
  u64 grid = (n_ub + RT - 1) / RT;
  if (grid > (u64)ctx->num_sms * 8) grid = (u64)ctx->num_sms * 8;
  MZ_BYTES(ctx, n.p == nullptr ? n.imm * 112 : 0);
  MZ_LAUNCH(ctx, k_explode, (unsigned)grid, RT, 0, d_r32, n, agg_kind, d_racc);
  return MZGPU_OK;
}

int32_t mz_reduce_corrections(mzgpu_ctx* ctx, const u64* d_batch_rows, u64 n, const TraceView& prior,
                              int agg_kind, DevMem* out, u64* n_out) {
  *n_out = 0;
  if (n == 0) return out->alloc(ctx, 16);
  const u64 n_tiles = (n + RT - 1) / RT;
  DevMem tiles;
  MZ_TRY(tiles.alloc(ctx, n_tiles * 4));
  u64* d_total = ctx->d_scratch + 30;
  MZ_LAUNCH(ctx, (k_corrections<false>), (unsigned)n_tiles, RT, 0, d_batch_rows, n, prior, agg_kind,
            tiles.as<u32>(), (const u32*)nullptr, (u64*)nullptr);
  MZ_LAUNCH(ctx, k_scan_tiles, 1, 1024, 0, tiles.as<u32>(), n_tiles, d_total);
  MZ_CUDA(ctx, cudaMemcpyAsync(ctx->h_scratch + 30, d_total, 8, cudaMemcpyDeviceToHost, ctx->stream));
  MZ_SYNC(ctx);
  ctx->stats.d2h_bytes += 8;
  const u64 total = ctx->h_scratch[30];
  MZ_TRY(out->alloc(ctx, total * 64));
  *n_out = total;
  if (total == 0) return MZGPU_OK;
  MZ_LAUNCH(ctx, (k_corrections<true>), (unsigned)n_tiles, RT, 0, d_batch_rows, n, prior, agg_kind,
            (u32*)nullptr, tiles.as<u32>(), out->as<u64>());
  return MZGPU_OK;
}

// Single-pass form: batch length read on the device; at most two output rows per
// new (key, time) row, so capacity 2 * n_ub always suffices.
int32_t mz_reduce_corrections_async(mzgpu_ctx* ctx, const u64* d_batch_rows, DLen n, u64 n_ub,
                                    const TraceView& prior, int agg_kind, u64* d_out, u64 out_cap,
                                    u64* d_out_len) {
  LookBack lb;
  MZ_TRY(mz_lookback_begin(ctx, (n_ub + RT - 1) / RT, &lb));
  u64 grid = (n_ub + RT - 1) / RT;
  if (grid > (u64)ctx->num_sms * 8) grid = (u64)ctx->num_sms * 8;
  if (grid == 0) grid = 1;
  MZ_BYTES(ctx, n.p == nullptr ? n.imm * (80 + 16 + 80 + 128) : 0);
  MZ_LAUNCH(ctx, k_corrections_lb, (unsigned)grid, RT, 0, d_batch_rows, n, prior, agg_kind, lb, d_out, out_cap,
            d_out_len, ctx->d_status);
  return MZGPU_OK;
}
