// oracle/mzo_ops.hpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the join and reduce operators over arrangements:
//   mz_join_core       src/compute/src/render/join/mz_join_core.rs:56-934
//   half_join          differential-dogs3 0.23.0 half_join2 (external) as called
//                      from src/compute/src/render/join/delta_join.rs:401-431,
//                      semantics SURVEY.md A8 (delta_join.rs:173-263)
//   build_update_stream src/compute/src/render/join/delta_join.rs:600-707
//   accumulable reduce src/compute/src/render/reduce.rs:1261-1471 (build_accumulable),
//                      :1530-1669 (datum_to_accumulator), :1671-1835 (finalize_accum),
//                      :1940-2104 (Semigroup / Multiply) and the reduce_abelian
//                      contract src/compute/src/extensions/reduce.rs:52-107
//
// PARITY UNPINNED for these operators at the unit level: the reference holds no
// unit-level golden vectors for join / half_join / reduce (they
// are pinned only end-to-end by sqllogictest, which needs a running
// Materialize).  They are validated in tests/ by algebraic identities
// (join == consolidated brute-force cross product; sum of delta paths == change
// of the full join; reduce output accumulated == GROUP BY of accumulated
// input) and by the agreement of the two join strategies below.
#pragma once
#include <cmath>
#include <deque>
#include <map>

#include "mzo_spine.hpp"

namespace mzo {

typedef std::shared_ptr<Batch<mzgpu_r32>> ValBatch;
typedef std::shared_ptr<Batch<mzgpu_racc>> AccBatch;

template <class R>
SpineOps<std::shared_ptr<Batch<R>>> real_ops() {
  typedef std::shared_ptr<Batch<R>> B;
  SpineOps<B> o;
  o.len = [](const B& b) { return b->len(); };
  o.desc = [](const B& b) { return b->desc; };
  o.merge = [](const B& a, const B& b, u64 since) { return merge_batches(*a, *b, since); };
  o.empty = [](u64 lo, u64 up, u64 since) { return empty_batch<R>(lo, up, since); };
  return o;
}

typedef Spine<ValBatch> ValSpine;
typedef Spine<AccBatch> AccSpine;

// ------------------------------------------------------------ join_core
// Join result: closure == nullptr -> (key, val1, val2) rows (R40); otherwise
// the closure output as R32.  We always produce R40-shaped records internally
// and narrow at the end so both strategies share code.
struct JoinOut {
  std::vector<mzgpu_r40> r40;
  std::vector<mzgpu_r32> r32;
};

struct Edit {
  u64 val;
  u64 time;
  i64 diff;
};

// EditList::load (mz_join_core.rs:816-837): per value, (time.join(meet), diff)
// pairs consolidated with consolidate_from.
template <class C>
static void load_edits(C& cursor, u64 meet, std::vector<Edit>& edits) {
  edits.clear();
  std::vector<mzgpu_r16> td;  // (time as key, diff)
  while (cursor.val_valid()) {
    u64 val = cursor.val_row().val;
    td.clear();
    cursor.map_times([&](const mzgpu_r32& r) {
      mzgpu_r16 x;
      x.key = std::max(r.time, meet);
      x.diff = r.diff;
      td.push_back(x);
    });
    consolidate(td);
    for (auto& x : td) edits.push_back(Edit{val, x.key, x.diff});
    cursor.step_val();
  }
}

struct JoinEmitter {
  const mzgpu_closure* closure;
  JoinOut* out;
  size_t produced = 0;
  void emit(u64 key, u64 v1, u64 v2, u64 t, i64 d) {
    if (closure == nullptr) {
      mzgpu_r40 r{key, v1, v2, t, d};
      out->r40.push_back(r);
      ++produced;
    } else {
      u64 k, v;
      if (closure_apply(closure, key, v1, v2, &k, &v)) {
        mzgpu_r32 r{k, v, t, d};
        out->r32.push_back(r);
        ++produced;
      }
    }
  }
};

// Joiner::join_key_simple (mz_join_core.rs:714-726).
static inline void join_key_simple(u64 key, const std::vector<Edit>& h1, const std::vector<Edit>& h2,
                                   JoinEmitter& em) {
  for (auto& a : h1)
    for (auto& b : h2) em.emit(key, a.val, b.val, std::max(a.time, b.time), wmul(a.diff, b.diff));
}

// Joiner::join_key_linear_time_scan + ValueHistory (mz_join_core.rs:729-793,
// 853-934): replay both histories in time order; each edit is paired with the
// accumulated past of the other side.
struct ValueHistory {
  struct Fut {
    u64 time, meet;
    u64 val;
    i64 diff;
  };
  std::vector<Fut> future;
  std::vector<mzgpu_r32> past;  // (key unused, val, time, diff)
  void replay(const std::vector<Edit>& edits) {
    future.clear();
    past.clear();
    for (auto& e : edits) future.push_back(Fut{e.time, e.time, e.val, e.diff});
    // sort descending by (time, meet, value idx, diff); value order == idx order
    std::sort(future.begin(), future.end(), [](const Fut& x, const Fut& y) {
      if (x.time != y.time) return x.time > y.time;
      if (x.val != y.val) return x.val > y.val;
      return x.diff > y.diff;
    });
    for (size_t i = 1; i < future.size(); ++i)
      future[i].meet = std::min(future[i].meet, future[i - 1].meet);
  }
  bool is_empty() const { return future.empty(); }
  void step() {
    Fut f = future.back();
    future.pop_back();
    past.push_back(mzgpu_r32{0, f.val, f.time, f.diff});
  }
  void advance_past_by(u64 meet) {
    for (auto& p : past) p.time = std::max(p.time, meet);
    consolidate(past);
  }
};

static inline void join_key_linear(u64 key, const std::vector<Edit>& e1, const std::vector<Edit>& e2,
                                   JoinEmitter& em) {
  ValueHistory h1, h2;
  h1.replay(e1);
  h2.replay(e2);
  auto work1 = [&]() {
    auto f = h1.future.back();
    h2.advance_past_by(f.meet);
    for (auto& p : h2.past) em.emit(key, f.val, p.val, std::max(f.time, p.time), wmul(f.diff, p.diff));
    h1.step();
  };
  auto work2 = [&]() {
    auto f = h2.future.back();
    h1.advance_past_by(f.meet);
    for (auto& p : h1.past) em.emit(key, p.val, f.val, std::max(p.time, f.time), wmul(p.diff, f.diff));
    h2.step();
  };
  while (!h1.is_empty() && !h2.is_empty()) {
    if (h1.future.back().time < h2.future.back().time)
      work1();
    else
      work2();
  }
  while (!h1.is_empty()) work1();
  while (!h2.is_empty()) work2();
}

// strategy: 0 = as the reference chooses (<10 edits -> simple), 1 = always
// simple, 2 = always linear scan.
template <class C1, class C2>
static void join_cursors(C1& c1, C2& c2, u64 meet, int strategy, JoinEmitter& em) {
  std::vector<Edit> h1, h2;
  // Work::start_work key-merge loop (mz_join_core.rs:606-621).
  while (c1.key_valid() && c2.key_valid()) {
    u64 k1 = c1.key(), k2 = c2.key();
    if (k1 < k2) {
      c1.seek_key(k2);
    } else if (k2 < k1) {
      c2.seek_key(k1);
    } else {
      load_edits(c1, meet, h1);
      load_edits(c2, meet, h2);
      bool simple = strategy == 1 || (strategy == 0 && (h1.size() < 10 || h2.size() < 10));
      if (simple)
        join_key_simple(k1, h1, h2, em);
      else
        join_key_linear(k1, h1, h2, em);
      c1.step_key();
      c2.step_key();
    }
  }
}

// The operator: acknowledged frontiers + deferred work (mz_join_core.rs:109-327).
struct JoinCore {
  ValSpine* trace1;
  ValSpine* trace2;
  const mzgpu_closure* closure;
  int strategy = 0;
  u64 ack1 = 0, ack2 = 0;
  struct Work {
    int side;                        // which input's batch this is
    ValBatch batch;
    std::vector<ValBatch> others;    // cursor_through(ack_other)
    u64 cap;
  };
  std::deque<Work> todo;

  JoinCore(ValSpine* t1, ValSpine* t2, const mzgpu_closure* c) : trace1(t1), trace2(t2), closure(c) {
    // Pre-load: all existing trace1 batches are acknowledged, then each
    // existing trace2 batch is joined against trace1 through ack1
    // (mz_join_core.rs:109-190).
    for (auto& e : trace1->all_batches()) ack1 = e.batch->desc.upper;
    for (auto& e : trace2->all_batches()) {
      if (!e.batch->is_empty()) {
        Work w{1, e.batch, trace1->batches_through(ack1), 0};
        todo.push_back(w);
      }
      ack2 = e.batch->desc.upper;
    }
  }

  // A batch arrived on `side` (mz_join_core.rs:218-327).
  void push(int side, ValBatch batch, u64 cap) {
    u64& ack = side == 0 ? ack1 : ack2;
    u64 ack_other = side == 0 ? ack2 : ack1;
    ValSpine* other = side == 0 ? trace2 : trace1;
    // "ack <= batch.lower" guards against re-delivered batches.
    if (ack <= batch->desc.lower) {
      if (!batch->is_empty()) {
        Work w{side, batch, other->batches_through(ack_other), cap};
        todo.push_back(w);
      }
      ack = batch->desc.upper;
    }
    // physical compaction of each trace follows the acknowledged frontiers
    trace1->set_physical_compaction(ack1);
    trace2->set_physical_compaction(ack2);
  }

  // Work::process (mz_join_core.rs:534-582): each work item's output buffer is
  // consolidated before it is sent.  Returns true when the queue is empty.
  bool work(size_t fuel_rows, JoinOut& out) {
    size_t produced = 0;
    while (!todo.empty() && produced < fuel_rows) {
      Work w = todo.front();
      todo.pop_front();
      JoinOut local;
      JoinEmitter em{closure, &local};
      BatchCursor<mzgpu_r32> cb(w.batch.get());
      CursorList<mzgpu_r32> ct(w.others);
      if (w.side == 0) {
        join_cursors(cb, ct, w.cap, strategy, em);
      } else {
        join_cursors(ct, cb, w.cap, strategy, em);
      }
      consolidate(local.r40);
      consolidate(local.r32);
      produced += local.r40.size() + local.r32.size();
      out.r40.insert(out.r40.end(), local.r40.begin(), local.r40.end());
      out.r32.insert(out.r32.end(), local.r32.begin(), local.r32.end());
    }
    return todo.empty();
  }
};

// ------------------------------------------------------------ half_join
// dogs3 half_join2::half_join_internal_unsafe as configured at
// delta_join.rs:401-431.  `stream` rows are ((key, val1), time) with
// initial == time (total order: the data-time never moves).
static inline void half_join(const std::vector<mzgpu_r32>& stream_in, const std::vector<ValBatch>& trace,
                             int cmp_mode, const mzgpu_closure* closure, std::vector<mzgpu_r32>& out) {
  std::vector<mzgpu_r32> stream = stream_in;
  // the stash is sorted before probing so the cursor only moves forward
  std::sort(stream.begin(), stream.end(),
            [](const mzgpu_r32& a, const mzgpu_r32& b) { return Tr<mzgpu_r32>::less(a, b); });
  CursorList<mzgpu_r32> cursor(trace);
  std::vector<mzgpu_r16> buf;  // (time, diff) output buffer
  u64 cur_key = 0;
  bool have_key = false;
  for (const auto& s : stream) {
    if (!have_key || s.key != cur_key) {
      cursor.seek_key(s.key);
      have_key = true;
      cur_key = s.key;
    } else {
      // same key as the previous stream update: rewind values
      for (auto& c : cursor.cursors) c.rewind_vals();
      cursor.minimize_vals();
    }
    if (!cursor.key_valid() || cursor.key() != s.key) continue;
    while (cursor.val_valid()) {
      u64 v2 = cursor.val_row().val;
      buf.clear();
      cursor.map_times([&](const mzgpu_r32& r) {
        bool ok = cmp_mode == MZGPU_HALFJOIN_LE ? r.time <= s.time : r.time < s.time;
        if (ok) buf.push_back(mzgpu_r16{std::max(r.time, s.time), r.diff});
      });
      consolidate(buf);
      if (!buf.empty()) {
        u64 k = s.key, v = v2;
        bool keep = true;
        if (closure != nullptr) keep = closure_apply(closure, s.key, s.val, v2, &k, &v);
        if (keep)
          for (auto& td : buf) out.push_back(mzgpu_r32{k, v, td.key, wmul(s.diff, td.diff)});
      }
      cursor.step_val();
    }
  }
}

// build_update_stream (delta_join.rs:600-707).
static inline void update_stream(const Batch<mzgpu_r32>& batch, const mzgpu_closure* initial_closure,
                                 u64 skip_time, std::vector<mzgpu_r32>& out) {
  BatchCursor<mzgpu_r32> c(&batch);
  std::vector<mzgpu_r16> td;
  while (c.key_valid()) {
    while (c.val_valid()) {
      td.clear();
      c.map_times([&](const mzgpu_r32& r) {
        if (skip_time == FRONTIER_EMPTY || r.time != skip_time) td.push_back(mzgpu_r16{r.time, r.diff});
      });
      consolidate(td);
      if (!td.empty()) {
        u64 k = c.key(), v = c.val_row().val;
        bool keep = true;
        if (initial_closure != nullptr) keep = closure_apply(initial_closure, k, v, 0, &k, &v);
        if (keep)
          for (auto& x : td) out.push_back(mzgpu_r32{k, v, x.key, x.diff});
      }
      c.step_val();
    }
    c.step_key();
  }
}

// --------------------------------------------------- accumulable reduce
static const double FLOAT_SCALE = 16777216.0;  // 2^24, reduce.rs:1528

// `(n * FLOAT_SCALE) as i128`: Rust float->int casts saturate, NaN -> 0.
static inline i128 f64_to_i128_sat(double x) {
  if (std::isnan(x)) return 0;
  const double lim = 170141183460469231731687303715884105728.0;  // 2^127
  if (x >= lim) return (i128)(((u128)1 << 127) - 1);
  if (x <= -lim) return (i128)((u128)1 << 127);
  return (i128)x;
}

// explode_one + datum_to_accumulator + Multiply<Diff> (reduce.rs:1313-1334,
// 1530-1669, 2043-2104) for COUNT(val)+SUM(val).
static inline mzgpu_racc explode_row(const mzgpu_r32& r, int agg_kind) {
  mzgpu_racc a;
  std::memset(&a, 0, sizeof(a));
  a.key = r.key;
  a.time = r.time;
  a.total = r.diff;
  if (agg_kind == MZGPU_AGG_DISTINCT || agg_kind == MZGPU_AGG_THRESHOLD) return a;  // only the multiplicity matters
  a.non_nulls = r.diff;
  i128 acc;
  if (agg_kind == MZGPU_AGG_COUNT_SUM_F64) {
    double n;
    std::memcpy(&n, &r.val, 8);
    bool nan = std::isnan(n), pinf = n == INFINITY, ninf = n == -INFINITY;
    a.nans = nan ? r.diff : 0;
    a.pos_infs = pinf ? r.diff : 0;
    a.neg_infs = ninf ? r.diff : 0;
    acc = (nan || pinf || ninf) ? 0 : f64_to_i128_sat(n * FLOAT_SCALE);
  } else {
    acc = (i128)(i64)r.val;
  }
  u128 prod = (u128)acc * (u128)(i128)r.diff;  // wrapping_mul
  a.acc_lo = (u64)prod;
  a.acc_hi = (i64)(u64)(prod >> 64);
  return a;
}

// finalize_accum (reduce.rs:1671-1835) for COUNT and SUM over the accumulated
// diff, plus the AccumulableErrorCheck flag (reduce.rs:1418-1429).
static inline void finalize_row(const mzgpu_racc& s, int agg_kind, mzgpu_rout* o) {
  std::memset(o, 0, sizeof(*o));
  o->key = s.key;
  if (agg_kind == MZGPU_AGG_DISTINCT) {  // build_distinct: (key, ()) once; error row if the count is negative
    o->count = 1;
    if (s.total < 0) o->flags |= 2;
    return;
  }
  if (agg_kind == MZGPU_AGG_THRESHOLD) return;  // unit value; the multiplicity travels in the diff
  o->count = s.non_nulls;
  bool accum_zero = s.acc_lo == 0 && s.acc_hi == 0 && s.non_nulls == 0 && s.pos_infs == 0 &&
                    s.neg_infs == 0 && s.nans == 0;
  if (s.total > 0 && accum_zero) o->flags |= 1;  // SUM is NULL
  if (s.total == 0 && !accum_zero) o->flags |= 2; // error row
  if (agg_kind == MZGPU_AGG_COUNT_SUM_F64) {
    double v;
    if (s.nans > 0 || (s.pos_infs > 0 && s.neg_infs > 0))
      v = NAN;
    else if (s.pos_infs > 0)
      v = INFINITY;
    else if (s.neg_infs > 0)
      v = -INFINITY;
    else {
      i128 acc = (i128)(((u128)(u64)s.acc_hi << 64) | s.acc_lo);
      v = (double)acc / FLOAT_SCALE;
    }
    if (std::isnan(v)) v = NAN;  // canonical NaN bits
    std::memcpy(&o->sum_lo, &v, 8);
    if (std::isnan(v)) o->sum_lo = 0x7ff8000000000000ull;
    o->sum_hi = 0;
  } else {
    o->sum_lo = s.acc_lo;
    o->sum_hi = s.acc_hi;
  }
  if (o->flags & 1) {
    o->sum_lo = 0;
    o->sum_hi = 0;
  }
}

// build_accumulable as one operator: explode -> arrange (batcher + spine) ->
// reduce_abelian.  For every key of a new input batch and every time at which
// its accumulated input changes: out = logic(accumulated input) minus the
// accumulated previous output (extensions/reduce.rs:52-107).
struct ReduceAccumulable {
  int agg_kind;
  Batcher<mzgpu_racc> batcher;
  AccSpine input;
  // accumulated output per key (the output arrangement's current contents);
  // the output trace only ever holds at most one row per key at +1.
  std::map<u64, mzgpu_rout> output;

  explicit ReduceAccumulable(int kind) : agg_kind(kind), input(real_ops<mzgpu_racc>(), 1, false) {}

  void step(const mzgpu_r32* rows, size_t n, u64 upper, std::vector<mzgpu_rout>& out) {
    std::vector<mzgpu_racc> exploded;
    exploded.reserve(n);
    for (size_t i = 0; i < n; ++i) exploded.push_back(explode_row(rows[i], agg_kind));
    batcher.push_container(exploded.data(), exploded.size());
    AccBatch batch = batcher.seal(upper);
    // accumulated input *before* this batch
    std::vector<AccBatch> prior;
    for (auto& e : input.all_batches()) prior.push_back(e.batch);
    if (batch->desc.lower != batch->desc.upper) input.insert(batch);
    input.set_physical_compaction(input.upper);
    std::vector<mzgpu_rout> local;
    CursorList<mzgpu_racc> pc(prior);
    BatchCursor<mzgpu_racc> bc(batch.get());
    while (bc.key_valid()) {
      u64 key = bc.key();
      mzgpu_racc s;
      std::memset(&s, 0, sizeof(s));
      s.key = key;
      pc.seek_key(key);
      if (pc.key_valid() && pc.key() == key) {
        while (pc.val_valid()) {
          pc.map_times([&](const mzgpu_racc& r) { Tr<mzgpu_racc>::add(s, r); });
          pc.step_val();
        }
      }
      // walk the batch's times for this key in order
      while (bc.val_valid()) {
        bc.map_times([&](const mzgpu_racc& r) {
          Tr<mzgpu_racc>::add(s, r);
          u64 t = r.time;
          auto it = output.find(key);
          bool had = it != output.end();
          mzgpu_rout fresh;
          bool has = !Tr<mzgpu_racc>::zero(s);
          // threshold_arrangement keeps (record, count) only while count is positive
          if (agg_kind == MZGPU_AGG_THRESHOLD) has = s.total > 0;
          const int64_t mult = agg_kind == MZGPU_AGG_THRESHOLD ? s.total : 1;
          if (has) finalize_row(s, agg_kind, &fresh);
          if (had) {
            mzgpu_rout old = it->second;
            old.time = t;
            old.diff = -old.diff;  // the multiplicity the output held
            local.push_back(old);
          }
          if (has) {
            fresh.time = t;
            fresh.diff = mult;
            local.push_back(fresh);
            mzgpu_rout keep = fresh;
            keep.time = 0;
            keep.diff = mult;
            output[key] = keep;
          } else if (had) {
            output.erase(it);
          }
        });
        bc.step_val();
      }
      bc.step_key();
    }
    consolidate(local);
    out.insert(out.end(), local.begin(), local.end());
  }
};

// MIN / MAX per key: the final result of the hierarchical reduce
// (build_bucketed_negated_output, reduce.rs:1050-1135): source = the key's accumulated
// (value, count) pairs with non-zero count; a non-positive count -> the error row; otherwise
// func.eval(values).  Emission protocol as in ReduceAccumulable (reduce_abelian contract,
// extensions/reduce.rs:52-107).
struct ReduceMinMax {
  int agg_kind;
  Batcher<mzgpu_r32> batcher;
  ValSpine input;
  std::map<u64, mzgpu_rout> output;

  explicit ReduceMinMax(int kind) : agg_kind(kind), input(real_ops<mzgpu_r32>(), 1, false) {}

  bool evaluate(u64 key, const std::map<u64, i64>& acc, mzgpu_rout* o) const {
    bool any = false, bad = false, have = false;
    u64 best = 0;
    for (auto& kv : acc) {
      if (kv.second == 0) continue;
      any = true;
      if (kv.second < 0) {
        bad = true;
        continue;
      }
      if (!have || (agg_kind == MZGPU_AGG_MIN ? kv.first < best : kv.first > best)) best = kv.first;
      have = true;
    }
    if (!any) return false;
    std::memset(o, 0, sizeof(*o));
    o->key = key;
    if (bad)
      o->flags = 2;
    else
      o->sum_lo = best;
    return true;
  }

  void step(const mzgpu_r32* rows, size_t n, u64 upper, std::vector<mzgpu_rout>& out) {
    batcher.push_container(rows, n);
    ValBatch batch = batcher.seal(upper);
    std::vector<ValBatch> prior;
    for (auto& e : input.all_batches()) prior.push_back(e.batch);
    if (batch->desc.lower != batch->desc.upper) input.insert(batch);
    input.set_physical_compaction(input.upper);
    std::vector<mzgpu_rout> local;
    CursorList<mzgpu_r32> pc(prior);
    BatchCursor<mzgpu_r32> bc(batch.get());
    while (bc.key_valid()) {
      const u64 key = bc.key();
      std::map<u64, i64> acc;
      pc.seek_key(key);
      if (pc.key_valid() && pc.key() == key) {
        while (pc.val_valid()) {
          pc.map_times([&](const mzgpu_r32& r) { acc[r.val] += r.diff; });
          pc.step_val();
        }
      }
      // the batch's updates of this key, by time
      std::map<u64, std::vector<mzgpu_r32>> by_time;
      while (bc.val_valid()) {
        bc.map_times([&](const mzgpu_r32& r) { by_time[r.time].push_back(r); });
        bc.step_val();
      }
      for (auto& tv : by_time) {
        for (auto& r : tv.second) acc[r.val] += r.diff;
        const u64 t = tv.first;
        auto it = output.find(key);
        const bool had = it != output.end();
        mzgpu_rout fresh;
        const bool has = evaluate(key, acc, &fresh);
        if (had) {
          mzgpu_rout old = it->second;
          old.time = t;
          old.diff = -1;
          local.push_back(old);
        }
        if (has) {
          fresh.time = t;
          fresh.diff = 1;
          local.push_back(fresh);
          mzgpu_rout keep = fresh;
          keep.time = 0;
          output[key] = keep;
        } else if (had) {
          output.erase(it);
        }
      }
      bc.step_key();
    }
    consolidate(local);
    out.insert(out.end(), local.begin(), local.end());
  }
};

// TopK per key (build_topk_negated_stage, top_k.rs:521-673): the live (value, count) pairs of the
// key, a non-positive count -> the error row; otherwise order the values, skip `offset` rows and keep
// at most `limit` (multiplicities counted).  The reference's reduce emits the negated complement and
// the dataflow concatenates it with the input; the window itself is the resulting collection, and its
// changes are what this operator emits.
struct ReduceTopK {
  i64 limit;  // < 0: none
  u64 offset;
  bool descending;
  Batcher<mzgpu_r32> batcher;
  ValSpine input;
  struct Window {
    bool error = false;
    std::map<u64, i64> rows;  // value -> multiplicity inside the window
  };
  std::map<u64, Window> output;

  ReduceTopK(i64 lim, u64 off, bool desc)
      : limit(lim), offset(off), descending(desc), input(real_ops<mzgpu_r32>(), 1, false) {}

  Window evaluate(const std::map<u64, i64>& acc) const {
    Window w;
    for (auto& kv : acc)
      if (kv.second < 0) {
        w.error = true;
        return w;
      }
    u64 skip = offset;
    i64 left = limit;
    auto take = [&](u64 val, i64 cnt) {
      if (cnt <= 0) return;
      if (skip > 0) {
        const u64 s = std::min<u64>(skip, (u64)cnt);
        skip -= s;
        cnt -= (i64)s;
      }
      if (limit >= 0) {
        cnt = std::min(cnt, left);
        left -= cnt;
      }
      if (cnt > 0) w.rows[val] = cnt;
    };
    if (descending)
      for (auto it = acc.rbegin(); it != acc.rend(); ++it) take(it->first, it->second);
    else
      for (auto it = acc.begin(); it != acc.end(); ++it) take(it->first, it->second);
    return w;
  }

  void step(const mzgpu_r32* rows, size_t n, u64 upper, std::vector<mzgpu_rout>& out) {
    batcher.push_container(rows, n);
    ValBatch batch = batcher.seal(upper);
    std::vector<ValBatch> prior;
    for (auto& e : input.all_batches()) prior.push_back(e.batch);
    if (batch->desc.lower != batch->desc.upper) input.insert(batch);
    input.set_physical_compaction(input.upper);
    std::vector<mzgpu_rout> local;
    CursorList<mzgpu_r32> pc(prior);
    BatchCursor<mzgpu_r32> bc(batch.get());
    auto emit = [&](u64 key, u64 val, u64 flags, u64 t, i64 d) {
      mzgpu_rout o;
      std::memset(&o, 0, sizeof(o));
      o.key = key;
      o.sum_lo = val;
      o.flags = flags;
      o.time = t;
      o.diff = d;
      local.push_back(o);
    };
    while (bc.key_valid()) {
      const u64 key = bc.key();
      std::map<u64, i64> acc;
      pc.seek_key(key);
      if (pc.key_valid() && pc.key() == key) {
        while (pc.val_valid()) {
          pc.map_times([&](const mzgpu_r32& r) { acc[r.val] += r.diff; });
          pc.step_val();
        }
      }
      std::map<u64, std::vector<mzgpu_r32>> by_time;
      while (bc.val_valid()) {
        bc.map_times([&](const mzgpu_r32& r) { by_time[r.time].push_back(r); });
        bc.step_val();
      }
      for (auto& tv : by_time) {
        for (auto& r : tv.second) {
          acc[r.val] += r.diff;
          if (acc[r.val] == 0) acc.erase(r.val);
        }
        const u64 t = tv.first;
        Window fresh = evaluate(acc);
        Window& old = output[key];
        if (old.error != fresh.error) emit(key, 0, 2, t, fresh.error ? 1 : -1);
        for (auto& kv : old.rows) {
          auto it = fresh.rows.find(kv.first);
          const i64 now = it == fresh.rows.end() ? 0 : it->second;
          if (now != kv.second) emit(key, kv.first, 0, t, now - kv.second);
        }
        for (auto& kv : fresh.rows)
          if (!old.rows.count(kv.first)) emit(key, kv.first, 0, t, kv.second);
        if (!fresh.error && fresh.rows.empty())
          output.erase(key);
        else
          old = fresh;
      }
      bc.step_key();
    }
    consolidate(local);
    out.insert(out.end(), local.begin(), local.end());
  }
};

}  // namespace mzo
