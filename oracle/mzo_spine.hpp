// oracle/mzo_spine.hpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of spine_fueled::Spine (differential-dataflow 0.23.0,
// external) following its faithful in-tree fork
//   src/persist-client/src/internal/trace.rs:1565-2262
// (insert :1737-1770, introduce_batch :1805-1887, roll_up :1896-1928,
//  apply_fuel :1937-1965, insert_at :1972-1985, tidy_layers :1993-2047,
//  exert :1698-1727, MergeState :2139-2246, FuelingMerge :1483-1563,
//  begin_merge :893-916) and Materialize's ExertionLogic
//   src/cluster/src/client.rs:216-257.
// The compute layer's spine additionally gates merging on the physical
// compaction frontier (`pending` + `consider_merges`, DD spine_fueled) and
// exposes cursor_through(upper) — see SURVEY.md A5/A6.
//
// Parity pin: the layer structure after sequences of pushes is checked against
// the reference's datadriven golden traces
// (src/persist-client/tests/trace/{compaction,compaction_apply_res,
//  compaction_apply_res_since,compaction_regression_size_reduction,
//  empty_batch_optimization,since_upper}) in tests/test_oracle_golden.py,
// using the "hollow" batch policy below (len = sum, no consolidation), which
// is what those traces record.
#pragma once
#include <functional>
#include <string>

#include "mzo_batch.hpp"

namespace mzo {

// Batch operations the spine needs.  `B` is a cheap handle.
template <class B>
struct SpineOps {
  std::function<size_t(const B&)> len;
  std::function<Desc(const B&)> desc;
  std::function<B(const B&, const B&, u64 since)> merge;  // b1.upper == b2.lower
  std::function<B(u64 lower, u64 upper, u64 since)> empty;
};

template <class B>
struct Spine {
  struct Entry {
    B batch;
    size_t id0, id1;  // SpineId(lo, hi)
  };
  struct Fueling {
    u64 since;
    size_t remaining_work;
  };
  struct MergeState {
    std::vector<Entry> batches;  // at most BATCHES_PER_LEVEL = 2
    bool has_merge = false;
    Fueling merge{0, 0};
    bool is_vacant() const { return batches.empty(); }
    bool is_single() const { return batches.size() == 1; }
    bool is_full() const { return batches.size() == 2; }
    bool is_complete() const { return has_merge && merge.remaining_work == 0; }
  };
  // (id0, id1, lower, upper, since) of every completed merge, in order; the
  // analogue of the fork's FueledMergeReq log.
  struct MergeReq {
    size_t id0, id1;
    Desc desc;
  };

  SpineOps<B> ops;
  size_t effort = 1;
  size_t next_id_ = 0;
  u64 since = 0;           // logical compaction frontier
  u64 physical = 0;        // physical compaction frontier
  bool gate_physical;      // DD compute spine: true; persist fork: false
  u64 upper = 0;
  std::vector<MergeState> merging;
  std::vector<Entry> pending;  // DD: batches not yet allowed to merge
  std::vector<MergeReq> merge_log;

  Spine(SpineOps<B> o, size_t effort_, bool gate_physical_)
      : ops(std::move(o)), effort(effort_), gate_physical(gate_physical_) {}

  size_t layer_len(const MergeState& m) const {
    size_t n = 0;
    for (auto& e : m.batches) n += ops.len(e.batch);
    return n;
  }
  bool layer_is_empty(const MergeState& m) const { return layer_len(m) == 0; }

  void next_id(size_t* a, size_t* b) {
    *a = next_id_;
    next_id_ += 1;
    *b = next_id_;
  }

  // All batches, oldest to newest (spine_batches, trace.rs:1667-1669) followed
  // by pending ones.
  std::vector<Entry> all_batches() const {
    std::vector<Entry> out;
    for (size_t i = merging.size(); i-- > 0;)
      for (auto& e : merging[i].batches) out.push_back(e);
    for (auto& e : pending) out.push_back(e);
    return out;
  }

  // cursor_through(upper): batches whose upper <= `upper`
  // (src/compute/src/render/join/mz_join_core.rs:243-246).
  std::vector<B> batches_through(u64 through) const {
    std::vector<B> out;
    for (auto& e : all_batches()) {
      Desc d = ops.desc(e.batch);
      if (through == FRONTIER_EMPTY || (d.upper != FRONTIER_EMPTY && d.upper <= through))
        out.push_back(e.batch);
    }
    return out;
  }

  // Trace::insert.
  void insert(B batch) {
    Desc d = ops.desc(batch);
    assert(d.lower != d.upper);
    assert(d.lower == upper);
    size_t a, b;
    next_id(&a, &b);
    upper = d.upper;
    Entry e{batch, a, b};
    if (gate_physical) {
      pending.push_back(e);
      consider_merges();
    } else {
      insert_entry(e);
    }
  }

  // DD spine_fueled::consider_merges: a batch may enter the merge structure
  // once the physical compaction frontier has passed its upper.
  void consider_merges() {
    while (!pending.empty()) {
      Desc d = ops.desc(pending.front().batch);
      bool ok = physical == FRONTIER_EMPTY || (d.upper != FRONTIER_EMPTY && d.upper <= physical);
      if (!ok) break;
      Entry e = pending.front();
      pending.erase(pending.begin());
      insert_entry(e);
    }
  }

  void insert_entry(const Entry& e) {
    // If `batch` and the most recently inserted batch are both empty, fuse
    // them (trace.rs:1752-1764).
    if (ops.len(e.batch) == 0) {
      for (size_t pos = 0; pos < merging.size(); ++pos) {
        if (merging[pos].is_vacant()) continue;
        if (merging[pos].is_single() && layer_is_empty(merging[pos])) {
          insert_at(e, pos);
          Entry merged;
          if (complete_at(pos, &merged)) {
            merging[pos] = MergeState();
            merging[pos].batches.push_back(merged);
          }
          return;
        }
        break;
      }
    }
    size_t n = ops.len(e.batch);
    introduce_batch(e, level_of(n));
  }

  static size_t level_of(size_t n) {
    // usize::next_power_of_two().trailing_zeros(); next_power_of_two(0) == 1.
    // (next_power_of_two overflows beyond 2^63 in the reference; capped at level 63 here)
    if (n <= 1) return 0;
    const size_t l = 64 - (size_t)__builtin_clzll((unsigned long long)(n - 1));
    return l > 63 ? 63 : l;
  }

  void set_logical_compaction(u64 f) {
    if (f != FRONTIER_EMPTY && f > since) since = f;
    if (f == FRONTIER_EMPTY) since = f;
  }
  void set_physical_compaction(u64 f) {
    if (f == FRONTIER_EMPTY || (physical != FRONTIER_EMPTY && f > physical)) physical = f;
    if (gate_physical) consider_merges();
  }

  // DD spine_fueled::reduced: no merges in progress and fewer than two
  // non-empty layers.
  bool reduced() const {
    size_t non_empty = 0;
    for (auto& m : merging) {
      if (m.is_full()) return false;
      if (layer_len(m) > 0) ++non_empty;
      if (non_empty > 1) return false;
    }
    return true;
  }

  // Materialize's ExertionLogic (src/cluster/src/client.rs:227-254).
  size_t exert_logic(uint32_t proportionality) const {
    uint32_t prop = proportionality;
    if (prop == 0) return 0;
    bool skipping = true, first = true;
    for (size_t i = merging.size(); i-- > 0;) {
      size_t count = merging[i].batches.size();
      size_t len = layer_len(merging[i]);
      if (skipping && count == 0) continue;
      skipping = false;
      if (count > 1) return 1000;
      if (!first && prop > 0 && len > 0) return 1000;
      first = false;
      prop /= 2;
    }
    return 0;
  }

  // Trace::exert (trace.rs:1698-1727).
  bool exert(size_t eff) {
    tidy_layers();
    if (reduced()) return false;
    bool any = false;
    for (auto& m : merging) any = any || m.has_merge;
    if (any) {
      // isize::try_from(effort).unwrap_or(isize::MAX) (trace.rs:1706)
      apply_fuel(eff > (size_t)INT64_MAX ? (long long)INT64_MAX : (long long)eff);
    } else {
      size_t level = level_of(eff);
      size_t a, b;
      next_id(&a, &b);
      Entry e{ops.empty(upper, upper, since), a, b};
      introduce_batch(e, level);
    }
    return true;
  }

  void introduce_batch(const Entry& e, size_t batch_index) {
    long long fuel = (long long)(8ull << batch_index);
    fuel *= (long long)effort;
    apply_fuel(fuel);
    roll_up(batch_index);
    insert_at(e, batch_index);
    tidy_layers();
  }

  void roll_up(size_t index) {
    while (merging.size() <= index) merging.push_back(MergeState());
    bool any = false;
    for (size_t i = 0; i < index; ++i) any = any || !merging[i].is_vacant();
    if (any) {
      bool have = false;
      Entry merged;
      for (size_t i = 0; i < index; ++i) {
        if (have) {
          insert_at(merged, i);
          have = false;
        }
        have = complete_at(i, &merged);
      }
      if (have) insert_at(merged, index);
      if (merging[index].is_full()) {
        Entry m2;
        bool ok = complete_at(index, &m2);
        assert(ok);
        (void)ok;
        insert_at(m2, index + 1);
      }
    }
  }

  void apply_fuel(long long fuel_in) {
    for (size_t index = 0; index < merging.size(); ++index) {
      long long fuel = fuel_in;
      work(merging[index], &fuel);
      if (merging[index].is_complete()) {
        Entry complete;
        bool ok = complete_at(index, &complete);
        assert(ok);
        (void)ok;
        insert_at(complete, index + 1);
      }
    }
  }

  // FuelingMerge::work (trace.rs:1502-1506).
  void work(MergeState& m, long long* fuel) {
    if (!m.has_merge) return;
    size_t f = *fuel < 0 ? 0 : (size_t)*fuel;
    size_t used = std::min(f, m.merge.remaining_work);
    m.merge.remaining_work -= used;
    *fuel -= (long long)used;
  }

  // SpineBatch::begin_merge (trace.rs:893-916).
  bool begin_merge(const std::vector<Entry>& bs, bool with_frontier, Fueling* out) {
    if (bs.empty()) return false;
    u64 s = 0;
    size_t work = 0;
    for (auto& e : bs) {
      s = std::max(s, ops.desc(e.batch).since);
      work += ops.len(e.batch);
    }
    if (with_frontier) s = std::max(s, since);
    out->since = s;
    out->remaining_work = work;
    return true;
  }

  void insert_at(const Entry& e, size_t index) {
    while (merging.size() <= index) merging.push_back(MergeState());
    MergeState& m = merging[index];
    assert(!m.has_merge && "Attempted to insert batch into incomplete merge!");
    assert(m.batches.size() < 2 && "Attempted to insert batch into full layer!");
    if (!m.batches.empty()) {
      assert(m.batches.back().id1 == e.id0);
    }
    m.batches.push_back(e);
    if (m.is_full()) m.has_merge = begin_merge(m.batches, true, &m.merge);
  }

  // MergeState::complete + FuelingMerge::done (trace.rs:2216-2230,1512-1562).
  bool complete_at(size_t index, Entry* out) {
    MergeState m = std::move(merging[index]);
    merging[index] = MergeState();
    if (m.batches.size() <= 1) {
      if (m.batches.empty()) return false;
      *out = m.batches[0];
      return true;
    }
    Fueling f = m.merge;
    if (!m.has_merge) begin_merge(m.batches, false, &f);
    const Entry& first = m.batches.front();
    const Entry& last = m.batches.back();
    Desc d1 = ops.desc(first.batch), d2 = ops.desc(last.batch);
    Entry r;
    r.id0 = first.id0;
    r.id1 = last.id1;
    bool all_empty = true;
    for (auto& e : m.batches) all_empty = all_empty && ops.len(e.batch) == 0;
    if (all_empty) {
      r.batch = ops.empty(d1.lower, d2.upper, f.since);
    } else {
      r.batch = ops.merge(first.batch, last.batch, f.since);
      MergeReq req;
      req.id0 = r.id0;
      req.id1 = r.id1;
      req.desc.lower = d1.lower;
      req.desc.upper = d2.upper;
      req.desc.since = f.since;
      merge_log.push_back(req);
    }
    *out = r;
    return true;
  }

  void tidy_layers() {
    if (merging.empty()) return;
    size_t length = merging.size();
    if (!merging[length - 1].is_single()) return;
    size_t appropriate_level = level_of(layer_len(merging[length - 1]));
    while (appropriate_level < length - 1) {
      MergeState& current = merging[length - 2];
      if (current.is_vacant()) {
        merging.erase(merging.begin() + (length - 2));
        length = merging.size();
      } else {
        if (!current.is_full()) {
          size_t smaller = 0;
          for (size_t idx = 0; idx < length - 2; ++idx) smaller += merging[idx].batches.size() << idx;
          if (smaller <= ((size_t)1 << length) / 8) {
            MergeState state = std::move(merging[length - 2]);
            merging.erase(merging.begin() + (length - 2));
            assert(state.batches.size() == 1);
            for (auto& e : state.batches) insert_at(e, length - 2);
          }
        }
        break;
      }
    }
  }
};

}  // namespace mzo
