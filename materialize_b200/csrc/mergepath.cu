// mergepath.cu — Batch::Merger for two sorted, consolidated R32 batches as THREE ordinary kernels
// (no cooperative launch, no grid barrier), every size read on the device.
//
// Reference semantics: `Batch::Merger` with `advance_by(since)` (the spine's merges,
// src/persist-client/src/internal/trace.rs:1841-2047 schedule them; the chunk-level merge it
// replaces is `InternalMerge::merge_from`, src/timely-util/src/columnation.rs:579-634): the union
// of both inputs ordered by (key, val, max(time, since)), equal rows' diffs summed, zeros dropped;
// followed by the hash index of the result (OrdValBuilder's role, index.cu).
//
// Why not the fused cooperative kernel (fused.cu) for this: a merge of update-batch size there is
// 44-60 us whatever the size (two to five grid-wide barriers at ~5 us each, phases that re-read the
// rows four times), and cooperative launches cannot overlap with one another.  Both inputs are
// sorted, so the work is one pass:
//   k_mrg_partition  one warp per 1024-row output tile: merge-path split of (A, B) on the tile's
//                    diagonal, moved forward to the end of the run of equal (key, val, time') rows
//                    that straddles it -- a tile never shares a run with its neighbour, so the
//                    consolidation below needs no carry between tiles; also zeroes the hash table
//                    (size from the actual input count) and the result counters
//   k_mrg_tiles      CTA per tile (tickets): both pieces staged in shared memory with times
//                    advanced, merged by per-thread merge-path searches, runs summed, zeros dropped,
//                    the survivors written at the offset a decoupled look-back gives
//   k_mrg_index      hash index + distinct keys + longest key run of the result (as index.cu, with
//                    the row count and the mask read on the device)
// Algorithmic bytes: 2 * 32 B per input row for the merge + 8 B per row for the index pass.
#include "common.cuh"

namespace {

constexpr int MT = 256;                    // threads per CTA
constexpr int M_TILE = 1024;               // merged rows per tile before run alignment
constexpr int M_SLACK = 256;               // extra rows a tile may take to finish a run
constexpr int M_CAP = M_TILE + M_SLACK;    // rows staged in shared memory
constexpr int M_IT = (M_CAP + MT - 1) / MT;  // merged positions per thread

struct MergeArgs {
  const u64* a;
  const u64* b;
  DLen na, nb;
  u64 since;
  u64* out;
  u64 out_cap;
  HashSlot* table;
  u64 table_cap;  // slots allocated (a power of two >= 2 * cap)
  u64* st;        // [0] rows out, [1] table mask, [2] distinct keys, [3] longest key run
  u64* splits;    // [2 * (Tcap + 1)]: (ia, ib) at the start of every tile
  LookBack lb;
  u64* status;
};

__device__ __forceinline__ u64 adv(u64 t, u64 since) { return t < since ? since : t; }
// (key, val, time') of row i of a sorted input, time advanced
__device__ __forceinline__ void key_of(const u64* __restrict__ rows, u64 i, u64 since, u64* k) {
  const ulonglong2 kv = *reinterpret_cast<const ulonglong2*>(rows + i * 4);
  k[0] = kv.x;
  k[1] = kv.y;
  k[2] = adv(rows[i * 4 + 2], since);
}
__device__ __forceinline__ bool key_lt(const u64* x, const u64* y) {
  if (x[0] != y[0]) return x[0] < y[0];
  if (x[1] != y[1]) return x[1] < y[1];
  return x[2] < y[2];
}
__device__ __forceinline__ bool key_eq(const u64* x, const u64* y) { return x[0] == y[0] && x[1] == y[1] && x[2] == y[2]; }

__global__ void __launch_bounds__(MT) k_mrg_partition(const MergeArgs m) {
  const u64 na = dlen_get(m.na), nb = dlen_get(m.nb), n = na + nb;
  const u64 T = (n + M_TILE - 1) / M_TILE;
  const u64 gtid = (u64)blockIdx.x * MT + threadIdx.x, gstride = (u64)gridDim.x * MT;
  // hash table of the result: sized from the actual input count (as the fused kernel does)
  u64 slots = 2;
  while (slots < 2 * n) slots <<= 1;
  if (slots > m.table_cap) slots = m.table_cap;
  for (u64 i = gtid; i < slots * 2; i += gstride) ((u64*)m.table)[i] = 0;
  if (gtid == 0) {
    m.st[0] = 0;
    m.st[1] = slots - 1;
    m.st[2] = 0;
    m.st[3] = 0;
  }
  // one WARP per diagonal: a 32-ary search (every lane probes one candidate split per round, the ballot
  // narrows the range 32-fold) -- four or five dependent memory round trips instead of the twenty of a
  // binary search, which were the whole cost of this kernel for update-batch merges
  const u32 lane = threadIdx.x & 31;
  const u64 gwarp = gtid >> 5, nwarps = gstride >> 5;
  for (u64 t = gwarp; t <= T; t += nwarps) {
    const u64 d = t * M_TILE < n ? t * M_TILE : n;
    // merge path on diagonal d, ties taken from A first: ia = rows of A among the first d merged =
    // the first mid in [lo, hi] for which "A[mid] <= B[d - 1 - mid]" fails
    u64 lo = d > nb ? d - nb : 0, hi = d < na ? d : na;
    while (lo < hi) {
      const u64 width = hi - lo, step = (width + 31) / 32;
      const u64 mid = lo + (u64)lane * step;
      bool ok = false;  // predicate holds at mid (lanes past the range: treated as failing)
      if (mid < hi) {
        u64 ka[3], kb[3];
        key_of(m.a, mid, m.since, ka);
        key_of(m.b, d - 1 - mid, m.since, kb);
        ok = !key_lt(kb, ka);
      }
      const u32 okm = __ballot_sync(0xffffffffu, ok);
      const u32 f = okm == 0xffffffffu ? 32u : (u32)__ffs(~okm) - 1;  // first lane whose probe fails (monotone)
      // the answer lies after the last holding probe and at or before the first failing one
      const u64 new_hi = f < 32 && lo + (u64)f * step < hi ? lo + (u64)f * step : hi;
      const u64 new_lo = f == 0 ? lo : lo + (u64)(f - 1) * step + 1;
      lo = new_lo;
      hi = new_hi;
    }
    u64 ia = lo, ib = d - lo;
    if (lane == 0) {
      if (d > 0 && d < n) {
        // the last merged row before the split; rows equal to it on either side belong to its tile
        u64 prev[3], x[3];
        bool have = false;
        if (ia > 0) {
          key_of(m.a, ia - 1, m.since, prev);
          have = true;
        }
        if (ib > 0) {
          key_of(m.b, ib - 1, m.since, x);
          if (!have || key_lt(prev, x)) {
            prev[0] = x[0], prev[1] = x[1], prev[2] = x[2];
          }
        }
        while (ia < na) {
          key_of(m.a, ia, m.since, x);
          if (!key_eq(x, prev)) break;
          ++ia;
        }
        while (ib < nb) {
          key_of(m.b, ib, m.since, x);
          if (!key_eq(x, prev)) break;
          ++ib;
        }
      }
      m.splits[2 * t] = ia;
      m.splits[2 * t + 1] = ib;
    }
  }
}

struct MergeSmem {
  u64 rows[M_CAP][4];  // A piece then B piece, times advanced
  unsigned short ord[M_CAP];      // merged order: position -> staged row
  u32 scan[34];
  u32 tile;
  u64 bcast;
  u64 carry[4];  // slow path: running row
};

__global__ void __launch_bounds__(MT) k_mrg_tiles(const MergeArgs m) {
  __shared__ MergeSmem S;
  const u32 tid = threadIdx.x;
  const u64 na = dlen_get(m.na), nb = dlen_get(m.nb), n = na + nb;
  const u64 T = (n + M_TILE - 1) / M_TILE;
  while (true) {
    const u32 tile = lb_next_tile(m.lb, &S.tile);
    if ((u64)tile >= T) {
      if (T == 0 && tile == 0 && tid == 0) m.st[0] = 0;
      break;
    }
    // a run moved across a diagonal can swallow following tiles whole: their range is empty
    u64 a0 = m.splits[2 * (u64)tile], b0 = m.splits[2 * (u64)tile + 1];
    u64 a1 = m.splits[2 * (u64)tile + 2], b1 = m.splits[2 * (u64)tile + 3];
    if (a1 < a0) a1 = a0;  // (splits are monotone by construction; defensive)
    if (b1 < b0) b1 = b0;
    const u64 ma = a1 - a0, mb = b1 - b0, mm = ma + mb;
    u32 total = 0;
    u64 excl = 0;
    if (mm <= (u64)M_CAP) {
      // ---- stage both pieces, times advanced
      for (u64 i = tid; i < mm; i += MT) {
        const u64* src = i < ma ? m.a + (a0 + i) * 4 : m.b + (b0 + (i - ma)) * 4;
        const ulonglong2 kv = *reinterpret_cast<const ulonglong2*>(src);
        const ulonglong2 td = *reinterpret_cast<const ulonglong2*>(src + 2);
        S.rows[i][0] = kv.x;
        S.rows[i][1] = kv.y;
        S.rows[i][2] = adv(td.x, m.since);
        S.rows[i][3] = td.y;
      }
      __syncthreads();
      // ---- merged order: thread t places positions [t * M_IT, (t + 1) * M_IT)
      {
        const u32 d0 = tid * M_IT < (u32)mm ? tid * M_IT : (u32)mm;
        u32 lo = d0 > (u32)mb ? d0 - (u32)mb : 0, hi = d0 < (u32)ma ? d0 : (u32)ma;
        while (lo < hi) {
          const u32 mid = (lo + hi) >> 1;
          if (!key_lt(S.rows[ma + (d0 - 1 - mid)], S.rows[mid]))
            lo = mid + 1;
          else
            hi = mid;
        }
        u32 ia = lo, ib = d0 - lo;
#pragma unroll
        for (int k = 0; k < M_IT; ++k) {
          const u32 p = d0 + k;
          if (p < (u32)mm) {
            const bool take_a = ib >= (u32)mb || (ia < (u32)ma && !key_lt(S.rows[ma + ib], S.rows[ia]));
            S.ord[p] = (unsigned short)(take_a ? ia : ma + ib);
            if (take_a)
              ++ia;
            else
              ++ib;
          }
        }
      }
      __syncthreads();
      // ---- runs of equal (key, val, time'): the head sums its run; zeros drop out
      u64 sum[M_IT];
      bool keep[M_IT];
      u32 cnt = 0;
#pragma unroll
      for (int k = 0; k < M_IT; ++k) {
        const u32 p = tid * M_IT + k;
        keep[k] = false;
        sum[k] = 0;
        if (p < (u32)mm) {
          const u64* r = S.rows[S.ord[p]];
          const bool head = p == 0 || !key_eq(r, S.rows[S.ord[p - 1]]);
          if (head) {
            u64 s = r[3];
            for (u32 q = p + 1; q < (u32)mm && key_eq(S.rows[S.ord[q]], r); ++q) s += S.rows[S.ord[q]][3];
            sum[k] = s;
            keep[k] = s != 0;
            cnt += keep[k] ? 1u : 0u;
          }
        }
      }
      const u32 mine = block_exclusive_scan(cnt, S.scan, &total);
      excl = lb_exclusive_prefix(m.lb, tile, (u64)total, &S.bcast);
      u64 pos = excl + mine;
#pragma unroll
      for (int k = 0; k < M_IT; ++k) {
        if (keep[k]) {
          const u64* r = S.rows[S.ord[tid * M_IT + k]];
          if (pos >= m.out_cap) {
            atomicMax((unsigned long long*)m.status, (unsigned long long)(pos + 1));
          } else {
            u64* o = m.out + pos * 4;
            *reinterpret_cast<ulonglong2*>(o) = make_ulonglong2(r[0], r[1]);
            *reinterpret_cast<ulonglong2*>(o + 2) = make_ulonglong2(r[2], sum[k]);
          }
          ++pos;
        }
      }
    } else {
      // ---- a run longer than the slack (a (key, val) with hundreds of times collapsing onto
      // `since`): one thread walks the tile's pieces in global memory, twice (count, write)
      for (int pass = 0; pass < 2; ++pass) {
        if (tid == 0) {
          u64 ia = a0, ib = b0, cur[4] = {0, 0, 0, 0}, pos = excl;
          bool open = false;
          u32 c = 0;
          while (true) {
            const bool more = ia < a1 || ib < b1;
            u64 x[3] = {0, 0, 0}, d = 0;
            if (more) {
              u64 ka[3], kb[3];
              bool take_a = ib >= b1;
              if (!take_a && ia < a1) {
                key_of(m.a, ia, m.since, ka);
                key_of(m.b, ib, m.since, kb);
                take_a = !key_lt(kb, ka);
              }
              if (take_a) {
                key_of(m.a, ia, m.since, x);
                d = m.a[ia * 4 + 3];
                ++ia;
              } else {
                key_of(m.b, ib, m.since, x);
                d = m.b[ib * 4 + 3];
                ++ib;
              }
            }
            if (open && (!more || !key_eq(x, cur))) {
              if (cur[3] != 0) {
                if (pass == 1) {
                  if (pos >= m.out_cap) {
                    atomicMax((unsigned long long*)m.status, (unsigned long long)(pos + 1));
                  } else {
                    u64* o = m.out + pos * 4;
                    o[0] = cur[0], o[1] = cur[1], o[2] = cur[2], o[3] = cur[3];
                  }
                  ++pos;
                }
                ++c;
              }
              open = false;
            }
            if (!more) break;
            if (!open) {
              cur[0] = x[0], cur[1] = x[1], cur[2] = x[2], cur[3] = d;
              open = true;
            } else {
              cur[3] += d;
            }
          }
          if (pass == 0) S.scan[33] = c;
        }
        __syncthreads();
        if (pass == 0) {
          total = S.scan[33];
          excl = lb_exclusive_prefix(m.lb, tile, (u64)total, &S.bcast);
        }
      }
    }
    if ((u64)tile == T - 1 && tid == 0) m.st[0] = excl + total;
  }
}

// hash index, distinct keys, longest key run of the merged rows (index.cu's kernels with the row
// count and the mask taken from the result counters on the device)
__global__ void __launch_bounds__(MT) k_mrg_index(const u64* __restrict__ rows, u64* __restrict__ st,
                                                  HashSlot* __restrict__ table) {
  const u64 n = st[0], mask = st[1];
  __shared__ u32 s_heads[MT / 32], s_run[MT / 32];
  u32 heads = 0, longest = 0;
  for (u64 i = (u64)blockIdx.x * MT + threadIdx.x; i < n; i += (u64)gridDim.x * MT) {
    const u64 key = rows[i * 4];
    if (i != 0 && rows[(i - 1) * 4] == key) continue;
    u64 run = 1;
    while (run < 1024 && i + run < n && rows[(i + run) * 4] == key) ++run;
    ++heads;
    longest = (u32)run > longest ? (u32)run : longest;
    const u64 meta = (i + 1) | ((run <= 64 ? run : 0ull) << 44);
    u64 h = mix64(key) & mask;
    while (true) {
      const unsigned long long prev = atomicCAS((unsigned long long*)&table[h].meta, 0ull, (unsigned long long)meta);
      if (prev == 0ull) {
        table[h].key = key;
        break;
      }
      h = (h + 1) & mask;
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    heads += __shfl_xor_sync(0xffffffffu, heads, off);
    const u32 o = __shfl_xor_sync(0xffffffffu, longest, off);
    longest = o > longest ? o : longest;
  }
  if ((threadIdx.x & 31) == 0) {
    s_heads[threadIdx.x >> 5] = heads;
    s_run[threadIdx.x >> 5] = longest;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    u32 h = 0, r = 0;
    for (int w = 0; w < MT / 32; ++w) {
      h += s_heads[w];
      r = s_run[w] > r ? s_run[w] : r;
    }
    if (h) {
      atomicAdd((unsigned long long*)&st[2], (unsigned long long)h);
      atomicMax((unsigned long long*)&st[3], (unsigned long long)r);
    }
  }
}

}  // namespace

// Merge two sorted, consolidated R32 arrays (counts possibly device resident; cap = host bound on
// na + nb) with times advanced to `since`: rows, hash index and result counters as the fused
// kernel's merge leaves them (FusedOut), by three stream-ordered launches and no host wait.
int32_t mz_merge_r32_async(mzgpu_ctx* ctx, const void* d_a, DLen na, const void* d_b, DLen nb, u64 cap, u64 since,
                           FusedOut* res) {
  if (cap == 0) cap = 1;
  const u64 Tcap = (cap + M_TILE - 1) / M_TILE;
  if (Tcap > MZ_LB_TILES) {
    MZ_SET_ERR(ctx, "merge: %llu tiles exceed the look-back state", (unsigned long long)Tcap);
    return MZGPU_E_UNSUPPORTED;
  }
  u64 slots = 2;
  while (slots < 2 * cap) slots <<= 1;
  DevMem splits;
  MZ_TRY(splits.alloc(ctx, 16 * (Tcap + 2)));
  MZ_TRY(res->rows.alloc(ctx, cap * 32));
  res->rows_cap = cap;
  MZ_TRY(res->table.alloc(ctx, slots * sizeof(HashSlot)));
  MZ_TRY(res->st.make_pending(ctx));
  MergeArgs m;
  memset(&m, 0, sizeof(m));
  m.a = (const u64*)d_a;
  m.b = (const u64*)d_b;
  m.na = na;
  m.nb = nb;
  m.since = since;
  m.out = res->rows.as<u64>();
  m.out_cap = cap;
  m.table = res->table.as<HashSlot>();
  m.table_cap = slots;
  m.st = res->st.dptr();
  m.splits = splits.as<u64>();
  m.status = ctx->d_status;
  MZ_TRY(mz_lookback_begin(ctx, Tcap, &m.lb));
  // partition + table clear: enough CTAs to zero the table at bandwidth, at least one thread per tile
  u64 g1 = (slots * 2 + MT * 8 - 1) / (MT * 8);
  const u64 g1_min = (Tcap + 1 + MT / 32 - 1) / (MT / 32), g_max = (u64)ctx->num_sms * 8;  // a warp per diagonal
  if (g1 < g1_min) g1 = g1_min;
  if (g1 > g_max) g1 = g_max;
  const bool exact = na.p == nullptr && nb.p == nullptr;
  MZ_BYTES(ctx, 0);
  MZ_LAUNCH(ctx, k_mrg_partition, (unsigned)g1, MT, 0, m);
  u64 g2 = Tcap < (u64)ctx->num_sms * 4 ? Tcap : (u64)ctx->num_sms * 4;
  if (g2 == 0) g2 = 1;
  MZ_BYTES(ctx, exact ? (na.imm + nb.imm) * 64 : 0);
  MZ_LAUNCH(ctx, k_mrg_tiles, (unsigned)g2, MT, 0, m);
  // one row per thread where the machine can hold it: a head walks its key's run with dependent loads,
  // and four rows per thread made that chain four times as long for update-batch merges
  u64 g3 = (cap + MT - 1) / MT;
  if (g3 > g_max * 4) g3 = g_max * 4;
  if (g3 == 0) g3 = 1;
  MZ_BYTES(ctx, exact ? (na.imm + nb.imm) * 8 : 0);
  MZ_LAUNCH(ctx, k_mrg_index, (unsigned)g3, MT, 0, res->rows.as<u64>(), res->st.dptr(), res->table.as<HashSlot>());
  res->st.mark_written();
  return MZGPU_OK;
}
