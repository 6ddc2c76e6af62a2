// oracle/capi.cc — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// C entry points over the CPU restatement (mzo_*.hpp) so that tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
// legs can drive it through ctypes.  Nothing in materialize_b200/ links or
// loads this library.
#include <cstdio>
#include <sstream>
#include <string>

#include "mzo_column.hpp"
#include "mzo_correction.hpp"
#include "mzo_ops.hpp"
#include "mzo_vec.hpp"

using namespace mzo;

namespace {

// type-erased batch / batcher / spine handles keyed by row_bytes
struct BatchH {
  uint32_t row_bytes;
  ValBatch v;
  AccBatch a;
};
struct BatcherH {
  uint32_t row_bytes;
  Batcher<mzgpu_r32> v;
  Batcher<mzgpu_racc> a;
};
struct SpineH {
  uint32_t row_bytes;
  std::unique_ptr<ValSpine> v;
  std::unique_ptr<AccSpine> a;
};

// "hollow" batches for the datadriven golden traces: len + part names, merges
// concatenate parts and add lens (FuelingMerge::done, trace.rs:1512-1562).
struct Hollow {
  Desc desc;
  size_t len;
  std::vector<std::string> parts;
  size_t n_parts;
};
typedef std::shared_ptr<Hollow> HollowB;
struct HSpineH {
  std::unique_ptr<Spine<HollowB>> s;
  std::string text;
};

SpineOps<HollowB> hollow_ops() {
  SpineOps<HollowB> o;
  o.len = [](const HollowB& b) { return b->len; };
  o.desc = [](const HollowB& b) { return b->desc; };
  o.merge = [](const HollowB& a, const HollowB& b, u64 since) {
    auto r = std::make_shared<Hollow>();
    r->desc.lower = a->desc.lower;
    r->desc.upper = b->desc.upper;
    r->desc.since = since;
    r->len = a->len + b->len;
    r->parts = a->parts;
    r->parts.insert(r->parts.end(), b->parts.begin(), b->parts.end());
    r->n_parts = a->n_parts + b->n_parts;
    return r;
  };
  o.empty = [](u64 lo, u64 up, u64 since) {
    auto r = std::make_shared<Hollow>();
    r->desc.lower = lo;
    r->desc.upper = up;
    r->desc.since = since;
    r->len = 0;
    r->n_parts = 1;  // SpineBatch::empty holds one empty hollow part
    return r;
  };
  return o;
}

}  // namespace

extern "C" {

// ---------------------------------------------------------------- vectors
void* mzo_vec_new(uint32_t row_bytes) {
  Vec* v = new Vec();
  v->row_bytes = row_bytes;
  return v;
}
void mzo_vec_free(void* v) { delete (Vec*)v; }
uint64_t mzo_vec_len(void* v) { return ((Vec*)v)->bytes.size() / ((Vec*)v)->row_bytes; }
void* mzo_vec_data(void* v) { return ((Vec*)v)->bytes.data(); }
void mzo_vec_clear(void* v) { ((Vec*)v)->bytes.clear(); }

// ----------------------------------------------------------- consolidate
#define CONSOLIDATE(NAME, T)                         \
  uint64_t NAME(T* rows, uint64_t n) {               \
    std::vector<T> v(rows, rows + n);                \
    consolidate(v);                                  \
    if (!v.empty()) std::memcpy(rows, v.data(), v.size() * sizeof(T)); \
    return v.size();                                 \
  }
CONSOLIDATE(mzo_consolidate_r16, mzgpu_r16)
CONSOLIDATE(mzo_consolidate_r32, mzgpu_r32)
CONSOLIDATE(mzo_consolidate_r40, mzgpu_r40)
CONSOLIDATE(mzo_consolidate_racc, mzgpu_racc)
CONSOLIDATE(mzo_consolidate_rout, mzgpu_rout)

// in-place, no copy: what the CPU baseline times (config 1)
uint64_t mzo_consolidate_r16_inplace(mzgpu_r16* rows, uint64_t n) {
  std::sort(rows, rows + n, [](const mzgpu_r16& a, const mzgpu_r16& b) { return a.key < b.key; });
  uint64_t w = 0, i = 0;
  while (i < n) {
    mzgpu_r16 acc = rows[i];
    uint64_t j = i + 1;
    while (j < n && rows[j].key == acc.key) {
      acc.diff = wadd(acc.diff, rows[j].diff);
      ++j;
    }
    if (acc.diff != 0) rows[w++] = acc;
    i = j;
  }
  return w;
}

// ------------------------------------------------- chain merge / extract
// Merge two sorted consolidated lists split into chunks of `chunk_rows`
// (Merger::merge); returns the number of output rows.
uint64_t mzo_merge_chains_r32(const mzgpu_r32* a, uint64_t na, const mzgpu_r32* b, uint64_t nb,
                              uint64_t chunk_rows, mzgpu_r32* out) {
  auto mk = [&](const mzgpu_r32* p, uint64_t n) {
    Chain<mzgpu_r32> c;
    for (uint64_t i = 0; i < n; i += chunk_rows) {
      Chunk<mzgpu_r32> ch;
      ch.rows.assign(p + i, p + std::min(n, i + chunk_rows));
      c.push_back(std::move(ch));
    }
    return c;
  };
  Chain<mzgpu_r32> o;
  merge_chains(mk(a, na), mk(b, nb), o);
  uint64_t w = 0;
  for (auto& ch : o)
    for (auto& r : ch.rows) out[w++] = r;
  return w;
}

// Merger::extract; returns frontier of kept times.
uint64_t mzo_extract_r32(const mzgpu_r32* rows, uint64_t n, uint64_t upper, mzgpu_r32* ship,
                         uint64_t* n_ship, mzgpu_r32* keep, uint64_t* n_keep) {
  Chain<mzgpu_r32> c, s, k;
  Chunk<mzgpu_r32> ch;
  ch.rows.assign(rows, rows + n);
  c.push_back(std::move(ch));
  u64 frontier = FRONTIER_EMPTY;
  extract_chain(std::move(c), upper, &frontier, s, k);
  uint64_t w = 0;
  for (auto& x : s)
    for (auto& r : x.rows) ship[w++] = r;
  *n_ship = w;
  w = 0;
  for (auto& x : k)
    for (auto& r : x.rows) keep[w++] = r;
  *n_keep = w;
  return frontier;
}

// ---------------------------------------------------------------- batcher
void* mzo_batcher_new(uint32_t row_bytes) {
  BatcherH* b = new BatcherH();
  b->row_bytes = row_bytes;
  return b;
}
void mzo_batcher_free(void* b) { delete (BatcherH*)b; }
void mzo_batcher_push(void* h, const void* rows, uint64_t n) {
  BatcherH* b = (BatcherH*)h;
  if (b->row_bytes == 32)
    b->v.push_container((const mzgpu_r32*)rows, n);
  else
    b->a.push_container((const mzgpu_racc*)rows, n);
}
void* mzo_batcher_seal(void* h, uint64_t upper) {
  BatcherH* b = (BatcherH*)h;
  BatchH* out = new BatchH();
  out->row_bytes = b->row_bytes;
  if (b->row_bytes == 32)
    out->v = b->v.seal(upper);
  else
    out->a = b->a.seal(upper);
  return out;
}
uint64_t mzo_batcher_frontier(void* h) {
  BatcherH* b = (BatcherH*)h;
  return b->row_bytes == 32 ? b->v.frontier : b->a.frontier;
}
uint64_t mzo_batcher_len(void* h) {
  BatcherH* b = (BatcherH*)h;
  return b->row_bytes == 32 ? b->v.len() : b->a.len();
}

// ---------------------------------------------------------------- batches
void* mzo_batch_build(uint32_t row_bytes, const void* rows, uint64_t n, uint64_t lower,
                      uint64_t upper, uint64_t since) {
  BatchH* out = new BatchH();
  out->row_bytes = row_bytes;
  Desc d;
  d.lower = lower;
  d.upper = upper;
  d.since = since;
  if (row_bytes == 32) {
    const mzgpu_r32* p = (const mzgpu_r32*)rows;
    out->v = build_batch_from_rows(std::vector<mzgpu_r32>(p, p + n), d);
  } else {
    const mzgpu_racc* p = (const mzgpu_racc*)rows;
    out->a = build_batch_from_rows(std::vector<mzgpu_racc>(p, p + n), d);
  }
  return out;
}
void mzo_batch_free(void* b) { delete (BatchH*)b; }
uint64_t mzo_batch_len(void* h) {
  BatchH* b = (BatchH*)h;
  return b->row_bytes == 32 ? b->v->len() : b->a->len();
}
uint64_t mzo_batch_keys(void* h) {
  BatchH* b = (BatchH*)h;
  return b->row_bytes == 32 ? b->v->keys.size() : b->a->keys.size();
}
void mzo_batch_desc(void* h, uint64_t* out3) {
  BatchH* b = (BatchH*)h;
  Desc d = b->row_bytes == 32 ? b->v->desc : b->a->desc;
  out3[0] = d.lower;
  out3[1] = d.upper;
  out3[2] = d.since;
}
void mzo_batch_export(void* h, void* rows) {
  BatchH* b = (BatchH*)h;
  if (b->row_bytes == 32) {
    if (!b->v->rows.empty()) std::memcpy(rows, b->v->rows.data(), b->v->rows.size() * 32);
  } else {
    if (!b->a->rows.empty())
      std::memcpy(rows, b->a->rows.data(), b->a->rows.size() * sizeof(mzgpu_racc));
  }
}
// CSR arrays of the OrdValBatch layout (keys / key_offs / val_offs sizes)
void mzo_batch_csr_sizes(void* h, uint64_t* out3) {
  BatchH* b = (BatchH*)h;
  if (b->row_bytes == 32) {
    out3[0] = b->v->keys.size();
    out3[1] = b->v->key_offs.size();
    out3[2] = b->v->val_offs.size();
  } else {
    out3[0] = b->a->keys.size();
    out3[1] = b->a->key_offs.size();
    out3[2] = b->a->val_offs.size();
  }
}
void* mzo_batch_merge(void* h1, void* h2, uint64_t since) {
  BatchH* a = (BatchH*)h1;
  BatchH* b = (BatchH*)h2;
  BatchH* out = new BatchH();
  out->row_bytes = a->row_bytes;
  if (a->row_bytes == 32)
    out->v = merge_batches(*a->v, *b->v, since);
  else
    out->a = merge_batches(*a->a, *b->a, since);
  return out;
}

// ------------------------------------------------------------------ spine
void* mzo_spine_new(uint32_t row_bytes, uint32_t effort, int32_t gate_physical) {
  SpineH* s = new SpineH();
  s->row_bytes = row_bytes;
  if (row_bytes == 32)
    s->v.reset(new ValSpine(real_ops<mzgpu_r32>(), effort, gate_physical != 0));
  else
    s->a.reset(new AccSpine(real_ops<mzgpu_racc>(), effort, gate_physical != 0));
  return s;
}
void mzo_spine_free(void* s) { delete (SpineH*)s; }
void mzo_spine_insert(void* h, void* batch) {
  SpineH* s = (SpineH*)h;
  BatchH* b = (BatchH*)batch;
  if (s->row_bytes == 32)
    s->v->insert(b->v);
  else
    s->a->insert(b->a);
}
int32_t mzo_spine_exert(void* h, uint64_t effort) {
  SpineH* s = (SpineH*)h;
  return s->row_bytes == 32 ? s->v->exert(effort) : s->a->exert(effort);
}
uint64_t mzo_spine_exert_logic(void* h, uint32_t prop) {
  SpineH* s = (SpineH*)h;
  return s->row_bytes == 32 ? s->v->exert_logic(prop) : s->a->exert_logic(prop);
}
void mzo_spine_set_logical_compaction(void* h, uint64_t f) {
  SpineH* s = (SpineH*)h;
  if (s->row_bytes == 32)
    s->v->set_logical_compaction(f);
  else
    s->a->set_logical_compaction(f);
}
void mzo_spine_set_physical_compaction(void* h, uint64_t f) {
  SpineH* s = (SpineH*)h;
  if (s->row_bytes == 32)
    s->v->set_physical_compaction(f);
  else
    s->a->set_physical_compaction(f);
}
uint64_t mzo_spine_read_upper(void* h) {
  SpineH* s = (SpineH*)h;
  return s->row_bytes == 32 ? s->v->upper : s->a->upper;
}
// layers largest first: {n_batches, len0, len1, remaining_work}
uint32_t mzo_spine_layers(void* h, uint64_t* out4, uint32_t cap) {
  SpineH* s = (SpineH*)h;
  uint32_t n = 0;
  auto emit = [&](auto& sp) {
    for (size_t i = sp.merging.size(); i-- > 0;) {
      if (n >= cap) break;
      auto& m = sp.merging[i];
      out4[4 * n + 0] = m.batches.size();
      out4[4 * n + 1] = m.batches.size() > 0 ? sp.ops.len(m.batches[0].batch) : 0;
      out4[4 * n + 2] = m.batches.size() > 1 ? sp.ops.len(m.batches[1].batch) : 0;
      out4[4 * n + 3] = m.has_merge ? m.merge.remaining_work : 0;
      ++n;
    }
  };
  if (s->row_bytes == 32)
    emit(*s->v);
  else
    emit(*s->a);
  return n;
}
uint32_t mzo_spine_num_batches_through(void* h, uint64_t upper) {
  SpineH* s = (SpineH*)h;
  return s->row_bytes == 32 ? (uint32_t)s->v->batches_through(upper).size()
                            : (uint32_t)s->a->batches_through(upper).size();
}
// consolidated contents of the whole trace, times advanced to since
void mzo_spine_export(void* h, void* vec) {
  SpineH* s = (SpineH*)h;
  Vec* out = (Vec*)vec;
  if (s->row_bytes == 32) {
    std::vector<mzgpu_r32> all;
    for (auto& e : s->v->all_batches())
      for (auto r : e.batch->rows) {
        if (r.time < s->v->since) r.time = s->v->since;
        all.push_back(r);
      }
    consolidate(all);
    vec_append(out, all);
  } else {
    std::vector<mzgpu_racc> all;
    for (auto& e : s->a->all_batches())
      for (auto r : e.batch->rows) {
        if (r.time < s->a->since) r.time = s->a->since;
        all.push_back(r);
      }
    consolidate(all);
    vec_append(out, all);
  }
}

// --------------------------------------------- hollow spine (golden traces)
void* mzo_hspine_new() {
  HSpineH* h = new HSpineH();
  h->s.reset(new Spine<HollowB>(hollow_ops(), 1, false));
  return h;
}
void mzo_hspine_free(void* h) { delete (HSpineH*)h; }
void mzo_hspine_push(void* h, uint64_t lower, uint64_t upper, uint64_t since, uint64_t len,
                     const char* name) {
  auto b = std::make_shared<Hollow>();
  b->desc.lower = lower;
  b->desc.upper = upper;
  b->desc.since = since;
  b->len = len;
  b->n_parts = 1;
  if (name != nullptr && name[0] != 0) b->parts.push_back(name);
  ((HSpineH*)h)->s->insert(b);
}
void mzo_hspine_downgrade_since(void* h, uint64_t since) {
  ((HSpineH*)h)->s->set_logical_compaction(since);
}
// `spine-batches` rendering of SpineBatch::describe(extended = true)
// (trace.rs:919-965): "[id0-id1][lower][upper][since] parts/len names..."
const char* mzo_hspine_describe(void* hh) {
  HSpineH* h = (HSpineH*)hh;
  std::ostringstream os;
  for (auto& e : h->s->all_batches()) {
    os << "[" << e.id0 << "-" << e.id1 << "][" << e.batch->desc.lower << "][" << e.batch->desc.upper
       << "][" << e.batch->desc.since << "] " << e.batch->n_parts << "/" << e.batch->len;
    for (auto& p : e.batch->parts) os << " " << p;
    os << "\n";
  }
  h->text = os.str();
  return h->text.c_str();
}
// merge requests logged since creation: "[lower][upper][since]" per line
const char* mzo_hspine_merge_reqs(void* hh) {
  HSpineH* h = (HSpineH*)hh;
  std::ostringstream os;
  for (auto& r : h->s->merge_log)
    os << "[" << r.desc.lower << "][" << r.desc.upper << "][" << r.desc.since << "] " << r.id0
       << "-" << r.id1 << "\n";
  h->text = os.str();
  return h->text.c_str();
}
uint64_t mzo_hspine_since(void* h) { return ((HSpineH*)h)->s->since; }
uint64_t mzo_hspine_upper(void* h) { return ((HSpineH*)h)->s->upper; }

// ------------------------------------------------------------------- join
struct JoinH {
  mzgpu_closure closure;
  bool has_closure;
  std::unique_ptr<JoinCore> j;
};
void* mzo_join_new(void* spine1, void* spine2, const mzgpu_closure* closure, int32_t strategy) {
  JoinH* h = new JoinH();
  h->has_closure = closure != nullptr;
  if (closure) h->closure = *closure;
  h->j.reset(new JoinCore(((SpineH*)spine1)->v.get(), ((SpineH*)spine2)->v.get(),
                          h->has_closure ? &h->closure : nullptr));
  h->j->strategy = strategy;
  return h;
}
void mzo_join_free(void* h) { delete (JoinH*)h; }
void mzo_join_push(void* h, int32_t side, void* batch, uint64_t cap) {
  ((JoinH*)h)->j->push(side, ((BatchH*)batch)->v, cap);
}
// appends R40 (no closure) or R32 rows to `vec`; returns 1 when done
int32_t mzo_join_work(void* h, uint64_t fuel_rows, void* vec) {
  JoinOut out;
  bool done = ((JoinH*)h)->j->work(fuel_rows, out);
  if (((JoinH*)h)->has_closure)
    vec_append((Vec*)vec, out.r32);
  else
    vec_append((Vec*)vec, out.r40);
  return done ? 1 : 0;
}

// -------------------------------------------------------------- half_join
void mzo_half_join(const mzgpu_r32* stream, uint64_t n, void* spine, int32_t cmp_mode,
                   const mzgpu_closure* closure, int32_t consolidate_output, void* vec) {
  std::vector<mzgpu_r32> s(stream, stream + n), out;
  SpineH* sp = (SpineH*)spine;
  std::vector<ValBatch> batches;
  for (auto& e : sp->v->all_batches()) batches.push_back(e.batch);
  half_join(s, batches, cmp_mode, closure, out);
  if (consolidate_output) consolidate(out);
  vec_append((Vec*)vec, out);
}
void mzo_update_stream(void* batch, const mzgpu_closure* closure, uint64_t skip_time, void* vec) {
  std::vector<mzgpu_r32> out;
  update_stream(*((BatchH*)batch)->v, closure, skip_time, out);
  vec_append((Vec*)vec, out);
}
void mzo_map_rows(const mzgpu_r32* rows, uint64_t n, const mzgpu_closure* closure, void* vec) {
  std::vector<mzgpu_r32> out;
  for (uint64_t i = 0; i < n; ++i) {
    u64 k, v;
    if (closure_apply(closure, rows[i].key, rows[i].val, 0, &k, &v))
      out.push_back(mzgpu_r32{k, v, rows[i].time, rows[i].diff});
  }
  vec_append((Vec*)vec, out);
}

// ----------------------------------------------------------------- reduce
struct AnyReduce {
  ReduceAccumulable* acc = nullptr;
  ReduceMinMax* mm = nullptr;
  ReduceTopK* tk = nullptr;
  ~AnyReduce() {
    delete acc;
    delete mm;
    delete tk;
  }
};
void* mzo_topk_new(int64_t limit, uint64_t offset, int32_t descending) {
  AnyReduce* r = new AnyReduce();
  r->tk = new ReduceTopK(limit, offset, descending != 0);
  return r;
}
void* mzo_reduce_new(int32_t agg_kind) {
  AnyReduce* r = new AnyReduce();
  if (agg_kind == MZGPU_AGG_MIN || agg_kind == MZGPU_AGG_MAX)
    r->mm = new ReduceMinMax(agg_kind);
  else
    r->acc = new ReduceAccumulable(agg_kind);
  return r;
}
void mzo_reduce_free(void* r) { delete (AnyReduce*)r; }
void mzo_reduce_step(void* rv, const mzgpu_r32* rows, uint64_t n, uint64_t upper, void* vec) {
  std::vector<mzgpu_rout> out;
  AnyReduce* r = (AnyReduce*)rv;
  if (r->mm)
    r->mm->step(rows, n, upper, out);
  else if (r->tk)
    r->tk->step(rows, n, upper, out);
  else
    r->acc->step(rows, n, upper, out);
  vec_append((Vec*)vec, out);
}
void mzo_explode(const mzgpu_r32* rows, uint64_t n, int32_t agg_kind, mzgpu_racc* out) {
  for (uint64_t i = 0; i < n; ++i) out[i] = explode_row(rows[i], agg_kind);
}
void mzo_finalize(const mzgpu_racc* acc, uint64_t n, int32_t agg_kind, mzgpu_rout* out) {
  for (uint64_t i = 0; i < n; ++i) finalize_row(acc[i], agg_kind, &out[i]);
}

uint32_t mzo_route(uint64_t key, uint32_t peers) { return route(key, peers); }

// ------------------------------------------------- MV sink correction buffer
void* mzo_correction_new(double chain_proportionality, uint64_t chunk_capacity) {
  return new CorrectionV2(chain_proportionality, (size_t)chunk_capacity);
}
void mzo_correction_free(void* c) { delete (CorrectionV2*)c; }
void mzo_correction_insert(void* c, const mzgpu_r32* rows, uint64_t n, int32_t negate) {
  ((CorrectionV2*)c)->insert(rows, (size_t)n, negate != 0);
}
void mzo_correction_updates_before(void* c, uint64_t upper, void* vec) {
  std::vector<mzgpu_r32> out;
  ((CorrectionV2*)c)->updates_before(upper, &out);
  vec_append((Vec*)vec, out);
}
void mzo_correction_advance_since(void* c, uint64_t since) { ((CorrectionV2*)c)->advance_since(since); }
void mzo_correction_consolidate_at_since(void* c) { ((CorrectionV2*)c)->consolidate_at_since(); }
// chain lengths in updates (n_out entries written, at most cap); returns the number of chains;
// *staged = updates in the stage
uint64_t mzo_correction_chains(void* cv, uint64_t* lens, uint64_t cap, uint64_t* staged) {
  CorrectionV2* c = (CorrectionV2*)cv;
  for (size_t i = 0; i < c->chains.size() && i < cap; ++i) lens[i] = c->chains[i].size();
  if (staged) *staged = c->stage.size();
  return c->chains.size();
}

// ---- f4: columnar wire format (mzo_column.hpp)
// generic indexed::encode over caller slices (the reference's Column<i32> golden bytes)
void mzo_col_encode_slices(const void* const* ptrs, const uint64_t* lens, uint32_t k, void* vec) {
  std::vector<ColSlice> s;
  for (uint32_t i = 0; i < k; i++) s.push_back(ColSlice{(const uint8_t*)ptrs[i], (size_t)lens[i]});
  std::vector<uint64_t> store;
  col_encode(store, s);
  vec_append((Vec*)vec, store);
}
// indexed::decode: (byte offset, byte length) of every slice; -1 if malformed
int32_t mzo_col_decode_slices(const uint64_t* words, uint64_t n_words, uint64_t* off_len, uint32_t cap) {
  std::vector<ColSlice> s;
  if (!col_decode(words, n_words, &s)) return -1;
  for (size_t i = 0; i < s.size() && i < cap; i++) {
    off_len[2 * i] = (uint64_t)(s[i].p - (const uint8_t*)words);
    off_len[2 * i + 1] = s[i].len;
  }
  return (int32_t)s.size();
}
uint64_t mzo_col_length_in_words(const uint64_t* lens, uint32_t k) {
  std::vector<ColSlice> s;
  for (uint32_t i = 0; i < k; i++) s.push_back(ColSlice{nullptr, (size_t)lens[i]});
  return col_length_in_words(s);
}
int32_t mzo_col_at_capacity(uint64_t words) { return col_at_capacity(words) ? 1 : 0; }
// one container holding all the rows
void mzo_column_encode(int32_t layout, const mzgpu_r32* rows, uint64_t n, void* vec) {
  ColumnTyped c;
  c.layout = layout;
  for (uint64_t i = 0; i < n; i++) column_push(c, rows[i]);
  std::vector<uint64_t> store;
  col_encode(store, c.as_bytes());
  vec_append((Vec*)vec, store);
}
int32_t mzo_column_rows(int32_t layout, const uint64_t* words, uint64_t n_words, void* vec) {
  std::vector<mzgpu_r32> out;
  int rc = column_rows(layout, words, n_words, &out);
  if (rc == 0) vec_append((Vec*)vec, out);
  return rc;
}
// ColumnBuilder: push every row, finish; containers back to back in `vec`, sizes in chunk_words
uint32_t mzo_column_builder(int32_t layout, const mzgpu_r32* rows, uint64_t n, void* vec, uint64_t* chunk_words,
                            uint32_t cap) {
  ColumnBuilder b(layout);
  for (uint64_t i = 0; i < n; i++) b.push(rows[i]);
  b.finish();
  for (size_t i = 0; i < b.pending.size(); i++) {
    if (i < cap) chunk_words[i] = b.pending[i].size();
    vec_append((Vec*)vec, b.pending[i]);
  }
  return (uint32_t)b.pending.size();
}

}  // extern "C"
