// oracle/dataflow.cc — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// The CPU side of the benchmark dataflows: a "mini-timely" with W worker
// threads, key-hash Exchange between them, and the rendered shape of the
// TPC-H-Q3 delta join + accumulable reduce, assembled from the restated
// operators in mzo_ops.hpp exactly as mz_compute::render assembles them:
//
//   render_delta_join        src/compute/src/render/join/delta_join.rs:50-311
//     per path: build_update_stream (:600-707) -> [half_join per stage
//     (:324-376)] -> concatenate (:302-308)
//   plan                     test/sqllogictest/tpch_create_materialized_view.slt:320-369
//     %0:customer » %1:orders[#1]KAif » %2:lineitem[#0]KAif
//     %1:orders   » %0:customer[#0]KAef » %2:lineitem[#0]KAif
//     %2:lineitem » %1:orders[#0]KAif » %0:customer[#0]KAef
//   arrange                  src/compute/src/extensions/arrange.rs:86-119 (Exchange by key hash)
//   reduce                   src/compute/src/render/reduce.rs:1261-1471
//
// It is both the parity oracle for the GPU harness (same seeded inputs from
// materialize_b200/csrc/gen.h) and the timed CPU baseline ("C++ restatement of
// the reference CPU algorithms"; the Rust reference cannot be built here).
#include <pthread.h>

#include <chrono>
#include <thread>

#include "../materialize_b200/csrc/gen.h"
#include "../materialize_b200/csrc/q3_plan.h"
#include "mzo_ops.hpp"
#include "mzo_vec.hpp"

using namespace mzo;

namespace {

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

typedef mzg_q3_plan Q3Plan;
extern "C" void mzo_q3_plan(Q3Plan* p) { mzg_q3_plan_init(p); }

namespace {

struct Q3;

struct Q3Worker {
  Q3* df;
  uint32_t me;
  Batcher<mzgpu_r32> batcher[4];
  std::unique_ptr<ValSpine> spine[4];
  std::unique_ptr<ReduceAccumulable> reduce;
  std::vector<mzgpu_rout> out;  // reduce output of the current run
  uint64_t rows_in = 0;
};

struct Q3 {
  uint64_t seed;
  mzg_q3_scale sc;
  uint32_t W;
  uint64_t per_batch;
  Q3Plan plan;
  std::vector<std::unique_ptr<Q3Worker>> workers;
  pthread_barrier_t barrier;
  // mailboxes[src][dst]
  std::vector<std::vector<std::vector<mzgpu_r32>>> mail;
  // shared input for the current step: rows for the 4 arrangements
  std::vector<mzgpu_r32> input[4];
  uint64_t next_time = 0;  // time of the next step (0 = hydration)
};

// timely Exchange pact: route by key hash, all-to-all between worker threads.
void exchange(Q3* df, uint32_t me, std::vector<mzgpu_r32>& rows) {
  if (df->W == 1) return;
  for (uint32_t d = 0; d < df->W; ++d) df->mail[me][d].clear();
  for (auto& r : rows) df->mail[me][route(r.key, df->W)].push_back(r);
  pthread_barrier_wait(&df->barrier);
  rows.clear();
  for (uint32_t s = 0; s < df->W; ++s)
    rows.insert(rows.end(), df->mail[s][me].begin(), df->mail[s][me].end());
  pthread_barrier_wait(&df->barrier);
}

// One timestamp `t` on one worker.  `df->input[a]` holds the whole step's
// updates for arrangement a; this worker takes the slice [me*n/W, (me+1)*n/W)
// as "its" source output and exchanges it to the owning workers.
void worker_step(Q3Worker* w, uint64_t t) {
  Q3* df = w->df;
  const uint64_t upper = t + 1;
  ValBatch batch[4];
  for (int a = 0; a < 4; ++a) {
    const auto& in = df->input[a];
    size_t lo = in.size() * w->me / df->W, hi = in.size() * (w->me + 1) / df->W;
    std::vector<mzgpu_r32> mine(in.begin() + lo, in.begin() + hi);
    w->rows_in += mine.size();
    exchange(df, w->me, mine);
    // arrange: Batcher::push_container, seal at the new frontier, Trace::insert
    w->batcher[a].push_container(mine.data(), mine.size());
    batch[a] = w->batcher[a].seal(upper);
    w->spine[a]->insert(batch[a]);
    w->spine[a]->set_physical_compaction(upper);
  }
  // delta paths
  std::vector<mzgpu_r32> results;
  for (int path = 0; path < 3; ++path) {
    std::vector<mzgpu_r32> stream;
    // as_of rule: paths other than the first skip updates at as_of (= 0)
    update_stream(*batch[df->plan.source[path]], &df->plan.initial[path],
                  path == 0 ? FRONTIER_EMPTY : (uint64_t)0, stream);
    for (int st = 0; st < 2; ++st) {
      exchange(df, w->me, stream);  // half_join exchanges its stream by key
      std::vector<ValBatch> trace;
      for (auto& e : w->spine[df->plan.lookup[path][st]]->all_batches()) trace.push_back(e.batch);
      std::vector<mzgpu_r32> next;
      half_join(stream, trace, df->plan.cmp[path][st], &df->plan.stage[path][st], next);
      stream.swap(next);
    }
    results.insert(results.end(), stream.begin(), stream.end());
  }
  // reduce: Exchange by group key, then explode/arrange/reduce
  exchange(df, w->me, results);
  w->reduce->step(results.data(), results.size(), upper, w->out);
  // half_join's frontier_func holds logical compaction at input_frontier.step_back()
  for (int a = 0; a < 4; ++a) {
    w->spine[a]->set_logical_compaction(t);
    size_t e = w->spine[a]->exert_logic(16);
    if (e) w->spine[a]->exert(e);
  }
  w->reduce->input.set_logical_compaction(t);
}

void gen_hydration(Q3* df) {
  for (int a = 0; a < 4; ++a) df->input[a].clear();
  for (uint64_t i = 0; i < df->sc.n_customer; ++i) df->input[0].push_back(mzg_q3_customer(df->seed, i));
  mzg_q3_order o;
  for (uint64_t j = 0; j < df->sc.n_orders; ++j) {
    mzg_q3_order_row(df->seed, df->sc, j, 0, &o);
    df->input[1].push_back(mzgpu_r32{o.orderkey, mzg_q3_orders_by_orderkey_val(&o), 0, 1});
    df->input[2].push_back(mzgpu_r32{o.custkey, mzg_q3_orders_by_custkey_val(&o), 0, 1});
    for (uint32_t l = 0; l < o.n_lineitems; ++l)
      df->input[3].push_back(mzgpu_r32{o.orderkey, o.lineitem_val[l], 0, 1});
  }
}

void gen_batch(Q3* df, uint64_t b, uint64_t t) {
  for (int a = 0; a < 4; ++a) df->input[a].clear();
  mzg_q3_order o;
  for (uint64_t x = b * df->per_batch; x < (b + 1) * df->per_batch; ++x) {
    uint64_t j = mzg_q3_tick_order(x, df->sc);
    for (int ver = 0; ver < 2; ++ver) {
      int64_t d = ver == 0 ? -1 : 1;
      mzg_q3_order_row(df->seed, df->sc, j, ver, &o);
      df->input[1].push_back(mzgpu_r32{o.orderkey, mzg_q3_orders_by_orderkey_val(&o), t, d});
      df->input[2].push_back(mzgpu_r32{o.custkey, mzg_q3_orders_by_custkey_val(&o), t, d});
      for (uint32_t l = 0; l < o.n_lineitems; ++l)
        df->input[3].push_back(mzgpu_r32{o.orderkey, o.lineitem_val[l], t, d});
    }
  }
}

// run one timestamp on all workers; returns wall seconds of the parallel region
double run_step(Q3* df, uint64_t t) {
  double t0 = now_s();
  if (df->W == 1) {
    worker_step(df->workers[0].get(), t);
  } else {
    std::vector<std::thread> th;
    for (uint32_t i = 0; i < df->W; ++i) th.emplace_back(worker_step, df->workers[i].get(), t);
    for (auto& x : th) x.join();
  }
  return now_s() - t0;
}

}  // namespace

extern "C" {

void* mzo_q3_new(uint64_t seed, uint64_t n_customer, uint64_t n_orders, uint64_t n_part,
                 uint32_t workers, uint64_t per_batch) {
  Q3* df = new Q3();
  df->seed = seed;
  df->sc.n_customer = n_customer;
  df->sc.n_orders = n_orders;
  df->sc.n_part = n_part;
  df->W = workers;
  df->per_batch = per_batch;
  mzo_q3_plan(&df->plan);
  pthread_barrier_init(&df->barrier, nullptr, workers);
  df->mail.resize(workers);
  for (auto& m : df->mail) m.resize(workers);
  for (uint32_t i = 0; i < workers; ++i) {
    auto w = std::make_unique<Q3Worker>();
    w->df = df;
    w->me = i;
    for (int a = 0; a < 4; ++a) w->spine[a].reset(new ValSpine(real_ops<mzgpu_r32>(), 1, true));
    w->reduce.reset(new ReduceAccumulable(MZGPU_AGG_COUNT_SUM_I64));
    df->workers.push_back(std::move(w));
  }
  return df;
}
void mzo_q3_free(void* h) {
  Q3* df = (Q3*)h;
  pthread_barrier_destroy(&df->barrier);
  delete df;
}
// Hydration: all three inputs arrive at time 0.  Returns seconds; *rows = input rows.
double mzo_q3_hydrate(void* h, uint64_t* rows) {
  Q3* df = (Q3*)h;
  gen_hydration(df);
  *rows = df->input[0].size() + df->input[1].size() + df->input[3].size();
  double s = run_step(df, 0);
  df->next_time = 1;
  return s;
}
// Update batch `b` at the next timestamp.  Returns seconds for the step (input
// generation excluded); *rows = input update rows (orders counted once).
double mzo_q3_step(void* h, uint64_t b, uint64_t* rows) {
  Q3* df = (Q3*)h;
  uint64_t t = df->next_time;
  gen_batch(df, b, t);
  *rows = df->input[1].size() + df->input[3].size();
  double s = run_step(df, t);
  df->next_time = t + 1;
  return s;
}
// Drain the reduce output produced since the last drain (all workers,
// concatenated then consolidated) into `vec` (ROUT rows).
void mzo_q3_drain(void* h, void* vec_handle) {
  Q3* df = (Q3*)h;
  std::vector<mzgpu_rout> all;
  for (auto& w : df->workers) {
    all.insert(all.end(), w->out.begin(), w->out.end());
    w->out.clear();
  }
  consolidate(all);
  vec_append((Vec*)vec_handle, all);
}
// The current step's generated input for arrangement `a` (for feeding the GPU
// harness the same bytes in tests): copies rows, returns count.
uint64_t mzo_q3_input(void* h, int32_t a, mzgpu_r32* out, uint64_t cap) {
  Q3* df = (Q3*)h;
  uint64_t n = df->input[a].size();
  if (out != nullptr && cap >= n && n) std::memcpy(out, df->input[a].data(), n * sizeof(mzgpu_r32));
  return n;
}

// ---- stand-alone generators (tests / bench inputs on the host)
void mzo_gen_cfg1(uint64_t seed, uint64_t n, uint32_t key_bits, mzgpu_r16* out) {
  for (uint64_t i = 0; i < n; ++i) out[i] = mzg_cfg1_row(seed, i, key_bits);
}
void mzo_gen_cfg2(uint64_t seed, uint64_t n, uint64_t n_keys, mzgpu_r32* out) {
  for (uint64_t i = 0; i < n; ++i) out[i] = mzg_cfg2_row(seed, i, n_keys);
}
// zipf inverse-CDF table for theta over n keys
void mzo_zipf_cdf(double theta, uint64_t n, double* cdf) {
  double sum = 0;
  for (uint64_t k = 0; k < n; ++k) sum += 1.0 / std::pow((double)(k + 1), theta);
  double acc = 0;
  for (uint64_t k = 0; k < n; ++k) {
    acc += 1.0 / std::pow((double)(k + 1), theta) / sum;
    cdf[k] = acc;
  }
  cdf[n - 1] = 1.0;
}
void mzo_gen_cfg4(uint64_t seed, uint64_t first, uint64_t n, const double* cdf, uint64_t n_keys,
                  int32_t as_f64, mzgpu_r32* out) {
  for (uint64_t i = 0; i < n; ++i) out[i] = mzg_cfg4_row(seed, first + i, cdf, n_keys, as_f64);
}

}  // extern "C"
