// mzo_correction.hpp — CPU restatement of the MV sink's correction buffer (SURVEY.md §8(f)-3:
// the step AFTER the hot path; groundwork for a device version).  TEST INFRASTRUCTURE ONLY.
//
// Follows src/compute/src/sink/correction_v2.rs:
//   CorrectionV2::{insert, insert_negated, insert_inner}   :213-277
//   updates_before / consolidate_before                    :280-374
//   advance_since / consolidate_at_since                   :377-390
//   merge_chains / merge_chains_up_to                      :439-498
//   Stage::{insert, flush, advance_times}                  :1170-1243
//   consolidate (sorts by (time, data))                    :1285-1321
// Updates are mzgpu_r32 rows: data = (key, val), time, diff; a frontier is one u64
// (MZGPU_FRONTIER_EMPTY = the empty antichain).  The reference has no unit tests for this file
// ("parity unpinned"): tests/test_oracle_correction.py pins the restatement to the documented
// contract (consolidated updates before `upper`, times advanced by `since`) and to the chain
// invariant.  A chain's length in chunks is taken as ceil(updates / chunk_capacity); the
// reference's remainder chains can carry one partially filled leading chunk more.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

#include "../include/mzgpu.h"

namespace mzo {

struct CorrectionV2 {
  typedef mzgpu_r32 U;
  std::vector<std::vector<U>> chains;  // each ordered by (time, data) and consolidated
  std::vector<U> stage;
  uint64_t since = 0;  // Timestamp::MIN
  double chain_proportionality;
  size_t chunk_capacity;

  CorrectionV2(double prop, size_t chunk_cap) : chain_proportionality(prop), chunk_capacity(chunk_cap ? chunk_cap : 1) {}

  static bool time_data_less(const U& a, const U& b) {
    if (a.time != b.time) return a.time < b.time;
    if (a.key != b.key) return a.key < b.key;
    return a.val < b.val;
  }
  // consolidate (:1285-1321): sort by (time, data), sum equal neighbours, drop zeros
  static void consolidate_td(std::vector<U>& v) {
    std::sort(v.begin(), v.end(), time_data_less);
    size_t o = 0;
    for (size_t i = 0; i < v.size();) {
      U acc = v[i];
      size_t j = i + 1;
      while (j < v.size() && v[j].time == acc.time && v[j].key == acc.key && v[j].val == acc.val) {
        acc.diff = (int64_t)((uint64_t)acc.diff + (uint64_t)v[j].diff);
        ++j;
      }
      if (acc.diff != 0) v[o++] = acc;
      i = j;
    }
    v.resize(o);
  }
  size_t chunks(const std::vector<U>& c) const { return (c.size() + chunk_capacity - 1) / chunk_capacity; }

  // merge_chains (:439-456): every update advanced by `since`, merged in (time, data) order
  std::vector<U> merge_chains(std::vector<std::vector<U>>&& cs) const {
    std::vector<U> all;
    if (since == MZGPU_FRONTIER_EMPTY) return all;
    for (auto& c : cs)
      for (auto& u : c) {
        U x = u;
        x.time = std::max(x.time, since);
        all.push_back(x);
      }
    consolidate_td(all);
    return all;
  }
  // merge_chains_up_to (:459-498): only the updates that lie before `upper` after the advance are
  // merged; what lies beyond stays in its chain, untouched
  void merge_chains_up_to(std::vector<std::vector<U>>&& cs, uint64_t upper, std::vector<U>* merged,
                          std::vector<std::vector<U>>* remains) const {
    merged->clear();
    remains->clear();
    if (since == MZGPU_FRONTIER_EMPTY) return;
    if (upper == MZGPU_FRONTIER_EMPTY) {
      *merged = merge_chains(std::move(cs));
      return;
    }
    if (since >= upper) {
      for (auto& c : cs)
        if (!c.empty()) remains->push_back(std::move(c));
      return;
    }
    for (auto& c : cs) {
      std::vector<U> keep;
      for (auto& u : c) {
        U x = u;
        x.time = std::max(x.time, since);
        if (x.time < upper)
          merged->push_back(x);
        else
          keep.push_back(u);  // times beyond `upper` are beyond `since`: unchanged
      }
      if (!keep.empty()) remains->push_back(std::move(keep));
    }
    consolidate_td(*merged);
  }

  // Stage::insert (:1170-1210): ships whole chunks' worth of updates as a consolidated chain
  bool stage_insert(const U* rows, size_t n, std::vector<U>* chain) {
    if (n == 0) return false;
    const size_t count = stage.size() + n;
    const size_t chunk_count = count / chunk_capacity;
    size_t taken = 0;
    bool shipped = false;
    if (chunk_count > 0) {
      const size_t ship = chunk_count * chunk_capacity;
      std::vector<U> buf;
      buf.swap(stage);
      while (buf.size() < ship) buf.push_back(rows[taken++]);
      consolidate_td(buf);
      *chain = std::move(buf);
      shipped = true;
    }
    stage.insert(stage.end(), rows + taken, rows + n);
    return shipped;
  }

  void insert_inner(const std::vector<U>& ups) {
    std::vector<U> chain;
    if (!stage_insert(ups.data(), ups.size(), &chain)) return;
    // (the reference pushes the new chain even when consolidation emptied it)
    chains.push_back(std::move(chain));
    auto merge_needed = [&]() {
      if (chains.size() < 2) return false;
      const double last = (double)chunks(chains[chains.size() - 1]);
      const double prev = (double)chunks(chains[chains.size() - 2]);
      return last * chain_proportionality > prev;
    };
    while (merge_needed()) {
      std::vector<std::vector<U>> two;
      two.push_back(std::move(chains.back()));
      chains.pop_back();
      two.push_back(std::move(chains.back()));
      chains.pop_back();
      chains.push_back(merge_chains(std::move(two)));
    }
  }
  void insert(const U* rows, size_t n, bool negate) {
    if (since == MZGPU_FRONTIER_EMPTY) return;  // the empty since discards everything
    std::vector<U> ups(rows, rows + n);
    for (auto& u : ups) {
      u.time = std::max(u.time, since);
      if (negate) u.diff = (int64_t)(0 - (uint64_t)u.diff);
    }
    insert_inner(ups);
  }

  void consolidate_before(uint64_t upper) {
    if (chains.empty() && stage.empty()) return;
    std::vector<std::vector<U>> cs;
    cs.swap(chains);
    {  // Stage::flush (:1213-1226)
      consolidate_td(stage);
      if (!stage.empty()) {
        cs.push_back(std::move(stage));
        stage.clear();
      }
    }
    // drop the empty chains a fully cancelled insert may have left (Chain::into_cursor is None)
    if (cs.empty()) return;
    std::vector<U> merged;
    std::vector<std::vector<U>> remains;
    merge_chains_up_to(std::move(cs), upper, &merged, &remains);
    chains = std::move(remains);
    if (!merged.empty()) chains.push_back(std::move(merged));
    // restore the chain invariant (:346-369)
    size_t i = chains.empty() ? 0 : chains.size() - 1;
    while (i > 0) {
      const bool needs = i < chains.size() &&
                         (double)chunks(chains[i]) * chain_proportionality > (double)chunks(chains[i - 1]);
      if (needs) {
        std::vector<std::vector<U>> two;
        two.push_back(std::move(chains[i]));
        chains.erase(chains.begin() + (long)i);
        two.push_back(std::move(chains[i - 1]));
        chains[i - 1] = merge_chains(std::move(two));
      } else {
        --i;
      }
    }
  }

  // updates_before (:280-304)
  void updates_before(uint64_t upper, std::vector<U>* out) {
    out->clear();
    // PartialOrder::less_than(since, upper) on one-element antichains
    const bool since_lt_upper = since != MZGPU_FRONTIER_EMPTY && (upper == MZGPU_FRONTIER_EMPTY || since < upper);
    if (!since_lt_upper) return;
    consolidate_before(upper);
    for (auto& c : chains) {
      if (c.empty()) continue;
      const bool first_before = upper == MZGPU_FRONTIER_EMPTY || c.front().time < upper;
      if (!first_before) continue;
      for (auto& u : c) {
        if (!(upper == MZGPU_FRONTIER_EMPTY || u.time < upper)) break;
        out->push_back(u);
      }
      return;  // at most one chain holds updates before `upper` now
    }
  }
  void advance_since(uint64_t s) {
    // (the reference asserts since <= s)
    if (s == MZGPU_FRONTIER_EMPTY) {
      stage.clear();
    } else {
      for (auto& u : stage) u.time = std::max(u.time, s);
    }
    since = s;
  }
  void consolidate_at_since() {
    if (since == MZGPU_FRONTIER_EMPTY || since == UINT64_MAX - 1) return;  // try_step_forward
    consolidate_before(since + 1);
  }
};

}  // namespace mzo
