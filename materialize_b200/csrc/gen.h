// gen.h — seeded synthetic workload generators shared by the CUDA library's
// bench/test harness and by the CPU oracle, so that both sides see byte-identical
// inputs (SURVEY.md §8d "Distributions / seeds").  Header-only, host + device.
//
// All generators are counter-based (a pure function of (seed, index)), so rows
// can be produced in any order, on any device, in parallel.
//
// TPC-H-Q3-shaped tables are an integer-only restatement of Materialize's TPC-H
// load generator (src/storage/src/source/generator/tpch.rs:262-347 order_row,
// :351-354 partkey_retailprice, :385-397 order_key, :205-235 the tick pattern).
#pragma once
#include <stdint.h>

#include "../../include/mzgpu.h"

#if defined(__CUDACC__)
#define MZ_HD __host__ __device__ __forceinline__
#else
#define MZ_HD static inline
#endif

MZ_HD uint64_t mzg_splitmix64(uint64_t x) {
  x += 0x9e3779b97f4a7c15ull;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}
// independent stream `s` of the counter-based generator
MZ_HD uint64_t mzg_rand(uint64_t seed, uint64_t i, uint64_t s) {
  return mzg_splitmix64(mzg_splitmix64(seed * 0x2545f4914f6cdd1dull + s) ^ (i * 0x9e3779b97f4a7c15ull));
}
// uniform in [0, n) (multiply-shift; bias < 2^-32 for n < 2^32, irrelevant here)
MZ_HD uint64_t mzg_below(uint64_t r, uint64_t n) {
#if defined(__CUDA_ARCH__)
  return __umul64hi(r, n);
#else
  return (uint64_t)(((unsigned __int128)r * n) >> 64);
#endif
}

// ---------------------------------------------------------------- config 1
// consolidate(): (u64 key, i64 diff), key ~ U[0, 2^key_bits), diff ~ U{-3..3}
// (diff range as in src/timely-util/benches/columnar_merger.rs:72-77).
MZ_HD mzgpu_r16 mzg_cfg1_row(uint64_t seed, uint64_t i, uint32_t key_bits) {
  mzgpu_r16 r;
  uint64_t k = mzg_rand(seed, i, 0);
  r.key = key_bits >= 64 ? k : (k >> (64 - key_bits));
  r.diff = (int64_t)mzg_below(mzg_rand(seed, i, 1), 7) - 3;
  return r;
}

// ---------------------------------------------------------------- config 2
// arrange + join_core: key ~ U[0, n_keys), val = row index, time = 0, diff = +1.
MZ_HD mzgpu_r32 mzg_cfg2_row(uint64_t seed, uint64_t i, uint64_t n_keys) {
  mzgpu_r32 r;
  r.key = mzg_below(mzg_rand(seed, i, 0), n_keys);
  r.val = i;
  r.time = 0;
  r.diff = 1;
  return r;
}

// ---------------------------------------------------------------- config 4
// reduce: key = zipf rank (theta = 0.9 over n_keys) through a fixed bijection,
// val ~ U[-10^6, 10^6].  `cdf` is the host-built inverse-CDF table
// (cdf[k] = P(rank <= k), doubles, length n_keys) so host and device agree
// bit for bit.
MZ_HD uint64_t mzg_zipf_rank(const double* cdf, uint64_t n_keys, uint64_t r) {
  double u = (double)(r >> 11) * (1.0 / 9007199254740992.0);
  uint64_t lo = 0, hi = n_keys - 1;
  while (lo < hi) {
    uint64_t mid = (lo + hi) >> 1;
    if (cdf[mid] < u)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}
// bijection on [0, n): x -> (a*x + c) mod n with a coprime to n (a is prime and > n's factors)
MZ_HD uint64_t mzg_permute(uint64_t x, uint64_t n) {
  const uint64_t a = 2654435761ull;  // prime; coprime to any n not a multiple of it
#if defined(__CUDA_ARCH__)
  return (a * x + 40503ull) % n;     // x < n < 2^32 in all configs: no overflow
#else
  return (uint64_t)(((unsigned __int128)a * x + 40503ull) % n);
#endif
}
MZ_HD mzgpu_r32 mzg_cfg4_row(uint64_t seed, uint64_t i, const double* cdf, uint64_t n_keys,
                             int as_f64) {
  mzgpu_r32 r;
  uint64_t rank = mzg_zipf_rank(cdf, n_keys, mzg_rand(seed, i, 0));
  r.key = mzg_permute(rank, n_keys);
  int64_t v = (int64_t)mzg_below(mzg_rand(seed, i, 1), 2000001) - 1000000;
  if (as_f64) {
    double d = (double)v / 7.0;
    union { double d; uint64_t u; } cv;
    cv.d = d;
    r.val = cv.u;
  } else {
    r.val = (uint64_t)v;
  }
  r.time = 0;
  r.diff = 1;
  return r;
}

// ------------------------------------------------ TPC-H-Q3-shaped tables
// Scale: customer = 150_000 * SF, orders = 1_500_000 * SF, part = 200_000 * SF
// (src/sql/src/plan/statement/ddl.rs:2121-2124).
typedef struct mzg_q3_scale {
  uint64_t n_customer;
  uint64_t n_orders;
  uint64_t n_part;
} mzg_q3_scale;

MZ_HD mzg_q3_scale mzg_q3_scale_for(uint64_t sf) {
  mzg_q3_scale s;
  s.n_customer = 150000ull * sf;
  s.n_orders = 1500000ull * sf;
  s.n_part = 200000ull * sf;
  return s;
}

// order_key (tpch.rs:385-397): dbgen's sparse keys, 8 used of every 32.
MZ_HD uint64_t mzg_order_key(uint64_t i) {
  uint64_t low = i & 7;
  i >>= 3;
  i <<= 2;
  i <<= 3;
  return i + low;
}
// partkey_retailprice (tpch.rs:351-354), integer dollars.
MZ_HD uint64_t mzg_retailprice(uint64_t partkey) {
  return (90000 + ((partkey / 10) % 20001) + 100 * (partkey % 1000)) / 100;
}

// bit layouts of the packed value words (DESIGN.md "Q3 column packing")
#define MZG_Q3_DATE_CUTOFF 1169u /* 1995-03-15 as days since 1992-01-01 */
#define MZG_Q3_SEGMENT 1u        /* 'BUILDING' */
// customer val : mktsegment[0:3]
// orders-by-orderkey val : custkey[0:24] | orderdate[24:36] | shippriority[36:37]
// orders-by-custkey  val : orderkey[0:32] | orderdate[32:44] | shippriority[44:45]
// lineitem val : linenumber[0:3] | extendedprice[3:20] | discount[20:24] | shipdate[24:36]

MZ_HD mzgpu_r32 mzg_q3_customer(uint64_t seed, uint64_t i /* 0-based */) {
  mzgpu_r32 r;
  r.key = i + 1;
  r.val = mzg_below(mzg_rand(seed, i, 10), 5);
  r.time = 0;
  r.diff = 1;
  return r;
}

typedef struct mzg_q3_order {
  uint64_t orderkey;
  uint64_t custkey;
  uint32_t orderdate;
  uint32_t shippriority;
  uint32_t n_lineitems;
  uint64_t lineitem_val[7];
} mzg_q3_order;

// order_row (tpch.rs:262-347) for order index j (0-based) at `version` (how many
// times the order has been replaced by a tick).
MZ_HD void mzg_q3_order_row(uint64_t seed, mzg_q3_scale sc, uint64_t j, uint64_t version,
                            mzg_q3_order* o) {
  uint64_t oseed = mzg_rand(seed, j, 20 + version * 64);
  o->orderkey = mzg_order_key(j + 1);
  uint64_t ck = 0;
  for (uint64_t attempt = 0;; ++attempt) {  // custkey % 3 != 0 (tpch.rs:265-270)
    ck = 1 + mzg_below(mzg_rand(oseed, attempt, 1), sc.n_customer);
    if (ck % 3 != 0) break;
  }
  o->custkey = ck;
  o->orderdate = 1 + (uint32_t)mzg_below(mzg_rand(oseed, 0, 2), 2405);  // ORDER_END_DAYS
  o->shippriority = 0;
  o->n_lineitems = 1 + (uint32_t)mzg_below(mzg_rand(oseed, 0, 3), 7);
  for (uint32_t l = 0; l < o->n_lineitems; ++l) {
    uint64_t partkey = 1 + mzg_below(mzg_rand(oseed, l, 4), sc.n_part);
    uint64_t qty = 1 + mzg_below(mzg_rand(oseed, l, 5), 50);
    uint64_t ext = qty * mzg_retailprice(partkey);
    uint64_t disc = mzg_below(mzg_rand(oseed, l, 6), 9);
    uint64_t ship = o->orderdate + 1 + mzg_below(mzg_rand(oseed, l, 7), 121);
    o->lineitem_val[l] = (uint64_t)(l + 1) | (ext << 3) | (disc << 20) | (ship << 24);
  }
}
MZ_HD uint64_t mzg_q3_orders_by_orderkey_val(const mzg_q3_order* o) {
  return o->custkey | ((uint64_t)o->orderdate << 24) | ((uint64_t)o->shippriority << 36);
}
MZ_HD uint64_t mzg_q3_orders_by_custkey_val(const mzg_q3_order* o) {
  return o->orderkey | ((uint64_t)o->orderdate << 32) | ((uint64_t)o->shippriority << 44);
}

// The tick pattern (tpch.rs:205-235): replace one order (-old lineitems, +new
// lineitems, -old order, +new order).  Update batch `b` replaces the orders
// with sequence numbers [b*per_batch, (b+1)*per_batch); sequence number x maps
// to order index mzg_permute(x, n_orders), so each order is replaced at most
// once (version 0 -> 1) as long as batches * per_batch <= n_orders.
MZ_HD uint64_t mzg_q3_tick_order(uint64_t x, mzg_q3_scale sc) { return mzg_permute(x, sc.n_orders); }
