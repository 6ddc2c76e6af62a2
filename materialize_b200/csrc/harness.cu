// harness.cu — a minimal "timely worker" over the C ABI (libmzgpu_harness.so).
//
// No Rust toolchain exists in this image, so this C++ harness stands in for the
// timely worker + mz_compute::render: it owns frontiers (plain u64 times), seals
// arrangements when the input frontier advances, and wires the rendered shape of
// the TPC-H-Q3 delta join + accumulable reduce together EXACTLY as
// render_delta_join (src/compute/src/render/join/delta_join.rs:50-311) and
// build_accumulable (src/compute/src/render/reduce.rs:1261-1471) do — using
// nothing but the public entry points of include/mzgpu.h with device pointers.
// The only CUDA code here is the seeded synthetic-input generation (gen.h), so
// that benchmark inputs are born in HBM.
//
// Per timestamp t (one update batch), on every worker/GPU:
//   inputs --Exchange(key)--> Batcher::push_container -> seal(t+1) -> Trace::insert   (x4 arrangements)
//   for each delta path: build_update_stream -> [Exchange -> half_join] x2 -> concat
//   --Exchange(group key)--> explode / arrange / reduce_abelian -> output corrections
//   logical compaction to t, physical compaction to t+1, idle merge effort
#include <cuda_runtime.h>
#include <stdio.h>

#include <chrono>
#include <memory>

#include <utility>
#include <vector>

#include "../../include/mzgpu.h"
#include "gen.h"
#include "q3_plan.h"

namespace {

struct DevArr {
  void* p = nullptr;
  ~DevArr() {
    if (p) cudaFree(p);
  }
  int alloc(size_t bytes) {
    if (p) cudaFree(p);
    p = nullptr;
    return cudaMalloc(&p, bytes ? bytes : 16) == cudaSuccess ? 0 : -1;
  }
};

__global__ void k_gen_customers(uint64_t seed, uint64_t first, uint64_t n, mzgpu_r32* out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = mzg_q3_customer(seed, first + i);
}

// orders [first, first + n) at `version` with multiplicity `diff` and time `t`;
// lineitems are appended through an atomic cursor (order is irrelevant: the
// batcher sorts).
__global__ void k_gen_orders(uint64_t seed, mzg_q3_scale sc, uint64_t first, uint64_t n, int tick,
                             uint64_t version, uint64_t t, int64_t diff, uint64_t out_base,
                             mzgpu_r32* o_ok, mzgpu_r32* o_ck, mzgpu_r32* li,
                             unsigned long long* li_cursor) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t j = tick ? mzg_q3_tick_order(first + i, sc) : first + i;
  mzg_q3_order o;
  mzg_q3_order_row(seed, sc, j, version, &o);
  mzgpu_r32 r;
  r.time = t;
  r.diff = diff;
  r.key = o.orderkey;
  r.val = mzg_q3_orders_by_orderkey_val(&o);
  o_ok[out_base + i] = r;
  r.key = o.custkey;
  r.val = mzg_q3_orders_by_custkey_val(&o);
  o_ck[out_base + i] = r;
  unsigned long long at = atomicAdd(li_cursor, (unsigned long long)o.n_lineitems);
  for (uint32_t l = 0; l < o.n_lineitems; ++l) {
    r.key = o.orderkey;
    r.val = o.lineitem_val[l];
    li[at + l] = r;
  }
}

__global__ void k_gen_cfg1(uint64_t seed, uint64_t first, uint64_t n, uint32_t key_bits, mzgpu_r16* out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = mzg_cfg1_row(seed, first + i, key_bits);
}
__global__ void k_gen_cfg2(uint64_t seed, uint64_t first, uint64_t n, uint64_t n_keys, mzgpu_r32* out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = mzg_cfg2_row(seed, first + i, n_keys);
}
__global__ void k_gen_cfg4(uint64_t seed, uint64_t first, uint64_t n, const double* cdf, uint64_t n_keys,
                           int as_f64, uint64_t t, int64_t diff, mzgpu_r32* out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    mzgpu_r32 r = mzg_cfg4_row(seed, first + i, cdf, n_keys, as_f64);
    r.time = t;
    r.diff = diff;
    out[i] = r;
  }
}

}  // namespace

#define H_TRY(expr)                 \
  do {                              \
    int32_t _s = (expr);            \
    if (_s != MZGPU_OK) return _s;  \
  } while (0)
#define H_CUDA(expr)                                   \
  do {                                                 \
    if ((expr) != cudaSuccess) return MZGPU_E_CUDA;    \
  } while (0)

struct mzh_q3 {
  mzgpu_ctx* ctx;
  cudaStream_t stream;
  uint64_t seed;
  mzg_q3_scale sc;
  uint64_t per_batch;
  uint32_t worker, peers;
  mzg_q3_plan plan;
  mzgpu_batcher* batcher[4];
  mzgpu_spine* spine[4];
  mzgpu_reduce* reduce;
  mzgpu_buf* input[4];  // staged inputs of the next step (device resident)
  mzgpu_buf *results, *xchg, *out;
  // pipelined result read-back (end-to-end path): timestamps alternate between two output
  // buffers; the previous timestamp's corrections are copied out on the copy stream while the
  // current one runs
  mzgpu_buf* out2 = nullptr;
  bool pipelined_out = false;
  cudaEvent_t ev_out[2] = {nullptr, nullptr};
  mzgpu_buf* out_of[2] = {nullptr, nullptr};
  bool out_pending[2] = {false, false};
  int out_parity = 0;
  mzgpu_buf *pstream[3], *pnext[3], *pxchg[3];  // per delta path: stream, next stage, exchange landing
  mzgpu_buf* axchg[4];                           // exchange landing per arrangement input
  DevArr gen_ok, gen_ck, gen_li, gen_cursor;
  uint64_t gen_cap_orders = 0;
  uint64_t next_time = 0;
  uint64_t last_rows_in = 0;
  uint64_t maintain_upper = 0;  // timestamp whose maintenance is still due
  uint64_t h2d_bytes = 0;       // host rows staged through mzh_q3_stage_host
  uint64_t d2h_bytes = 0;       // output rows copied out through mzh_q3_fetch_out
  // end-to-end path: host batches land in double-buffered device staging through a
  // copy stream, so the H2D copy of the next batch overlaps the current timestamp
  cudaStream_t copy_stream = nullptr;
  DevArr stage[2][4];
  uint64_t stage_cap[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
  uint64_t stage_n[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
  cudaEvent_t ev_up[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
  bool slot_full[2] = {false, false};
  bool slot_touched[2][4] = {{false, false, false, false}, {false, false, false, false}};
  int fill_slot = 0, run_slot = 0;
  bool static_rel[4] = {true, false, false, false};  // relations the generator never updates after hydration
  bool stepping = false;                             // false while hydrating
  bool use_p2p = false;                              // update-batch exchange rounds go over peer memory
  bool streams_prepared = false;                     // stage-0 streams already mapped + exchanged
  // host time spent inside the phases of a step (steady_clock; mzh_q3_host_ns): 0 inputs (staging hand-over,
  // arrangement-input exchange, pushes), 1 seals + spine inserts, 2 maintenance, 3 / 4 the delta paths' two
  // stages (exchange round, read-back, half joins), 5 result exchange, 6 reduce, 7 the whole step
  uint64_t host_ns[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};
namespace {
struct PhaseTimer {
  uint64_t* acc;
  std::chrono::steady_clock::time_point t0;
  explicit PhaseTimer(uint64_t* a) : acc(a), t0(std::chrono::steady_clock::now()) {}
  ~PhaseTimer() {
    *acc += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
  }
};
}  // namespace

static int32_t q3_gen_orders(mzh_q3* q, uint64_t first, uint64_t n, int tick, int n_versions, uint64_t t,
                             uint64_t* n_li) {
  // capacity for n orders x n_versions
  uint64_t need = n * n_versions;
  if (need > q->gen_cap_orders || q->gen_cursor.p == nullptr) {
    if (q->gen_ok.alloc(need * 32) || q->gen_ck.alloc(need * 32) || q->gen_li.alloc(need * 7 * 32) ||
        q->gen_cursor.alloc(8))
      return MZGPU_E_CUDA;
    q->gen_cap_orders = need;
  }
  H_CUDA(cudaMemsetAsync(q->gen_cursor.p, 0, 8, q->stream));
  if (n) {
    unsigned grid = (unsigned)((n + 255) / 256);
    for (int ver = 0; ver < n_versions; ++ver) {
      int64_t diff = (n_versions == 2 && ver == 0) ? -1 : 1;
      k_gen_orders<<<grid, 256, 0, q->stream>>>(q->seed, q->sc, first, n, tick, (uint64_t)ver, t, diff,
                                                (uint64_t)ver * n, (mzgpu_r32*)q->gen_ok.p,
                                                (mzgpu_r32*)q->gen_ck.p, (mzgpu_r32*)q->gen_li.p,
                                                (unsigned long long*)q->gen_cursor.p);
    }
    H_CUDA(cudaGetLastError());
  }
  unsigned long long h = 0;
  H_CUDA(cudaMemcpyAsync(&h, q->gen_cursor.p, 8, cudaMemcpyDeviceToHost, q->stream));
  H_CUDA(cudaStreamSynchronize(q->stream));
  *n_li = h;
  return MZGPU_OK;
}

// One exchange round for k buffers.  Update batches go over peer memory when the landing zones are
// connected (mzh_q3_use_p2p): no host wait, `recv_ub` = what this worker can receive at most, if
// the dataflow knows (the global batch size); bulk hydration chunks take the NCCL round.
static int32_t q3_exchange(mzh_q3* q, uint32_t k, mzgpu_buf** ins, mzgpu_buf** outs, const uint64_t* recv_ub) {
  if (q->use_p2p && q->stepping) return mzgpu_exchange_p2p(q->ctx, k, ins, outs, recv_ub);
  return mzgpu_exchange_many(q->ctx, k, ins, outs);
}

// Exchange(key) then Batcher::push_container for arrangement `a`
static int32_t q3_arrange_push(mzh_q3* q, int a, mzgpu_buf* rows) {
  mzgpu_buf* src = rows;
  // a relation that receives no updates on ANY worker (customer, in the tick
  // pattern) has nothing to exchange; every worker takes this branch together
  if (q->peers > 1 && !(q->static_rel[a] && q->stepping)) {
    H_TRY(mzgpu_exchange(q->ctx, rows, q->xchg));
    src = q->xchg;
  }
  return mzgpu_batcher_push_buf(q->batcher[a], src);
}

// everything after the inputs were pushed: seal, paths, reduce; then the
// maintenance the reference's worker loop does between operator activations
// (TraceManager::maintenance, src/compute/src/arrangement/manager.rs:55-73):
// physical compaction to the new upper, logical compaction, idle merge effort.
// Nothing in the operator part reads a row count back: counts flow between the
// operators in device memory (see include/mzgpu.h).
static int32_t q3_maintenance(mzh_q3* q);

static int32_t q3_run_timestamp(mzh_q3* q, uint64_t t) {
  const uint64_t upper = t + 1;
  mzgpu_batch* batch[4] = {nullptr, nullptr, nullptr, nullptr};
  int32_t st = MZGPU_OK;
  // the four arrange operators are activated by the same frontier advance: one batched seal
  {
    PhaseTimer pt(&q->host_ns[1]);
    st = mzgpu_batcher_seal_many(4, q->batcher, upper, batch);
    for (int a = 0; a < 4 && st == MZGPU_OK; ++a) st = mzgpu_spine_insert(q->spine[a], batch[a]);
  }
  // The previous timestamp's maintenance runs here: the seals above are already queued on
  // the device, so the merges it schedules (side stream) and the few lengths it has to read
  // back overlap with them instead of delaying them.
  if (st == MZGPU_OK) {
    PhaseTimer pt(&q->host_ns[2]);
    st = q3_maintenance(q);
  }
  if (st == MZGPU_OK) st = mzgpu_buf_clear(q->results);
  // the three delta paths run side by side, stage by stage, so that the exchange
  // points of one stage share a round (mzgpu_exchange_many).  A path whose source
  // relation is static (no updates on any worker) has nothing to do while stepping.
  bool active[3];
  for (int path = 0; path < 3 && st == MZGPU_OK; ++path) {
    active[path] = !(q->static_rel[q->plan.source[path]] && q->stepping);
    if (q->streams_prepared) continue;  // mapped and exchanged together with the inputs (mzh_q3_step)
    if (q->peers == 1) continue;        // one worker: the update stream is formed inside the first half join
    st = mzgpu_buf_clear(q->pstream[path]);
    // as_of rule: only the first relation's path sees the updates at as_of (= 0)
    if (st == MZGPU_OK && active[path])
      st = mzgpu_update_stream(q->ctx, batch[q->plan.source[path]], &q->plan.initial[path],
                               path == 0 ? MZGPU_FRONTIER_EMPTY : 0, q->pstream[path]);
  }
  for (int s = 0; s < 2 && st == MZGPU_OK; ++s) {
    PhaseTimer pt(&q->host_ns[3 + s]);
    if (q->peers > 1 && !(s == 0 && q->streams_prepared)) {  // half_join exchanges its stream by key
      mzgpu_buf *ins[3], *outs[3];
      uint32_t k = 0;
      for (int path = 0; path < 3; ++path)
        if (active[path]) {
          ins[k] = q->pstream[path];
          outs[k] = q->pxchg[path];
          ++k;
        }
      st = q3_exchange(q, k, ins, outs, nullptr);
      for (int path = 0; path < 3; ++path)
        if (active[path]) std::swap(q->pstream[path], q->pxchg[path]);
      // The one host wait of a timestamp over peer memory: the received stream lengths become
      // exact here (a worker cannot bound what the others send it), so the probes below run in
      // their bounded single-pass form; everything queued so far overlaps with this wait.
      if (st == MZGPU_OK && q->use_p2p && q->stepping)
        for (uint32_t i = 0; i < k; ++i) (void)mzgpu_buf_len(outs[i]);
    }
    // the active paths' stage-s half joins are independent operators: one launch.  Last stage:
    // the paths' outputs are concatenated (delta_join.rs:302-308), so every path appends
    // straight to the result collection (one chain, path order)
    mzgpu_buf *hs[3], *ho[3];
    mzgpu_spine* ht[3];
    int32_t hc[3];
    const mzgpu_closure* hcl[3];
    uint32_t hk = 0;
    for (int path = 0; path < 3 && st == MZGPU_OK; ++path) {
      if (!active[path]) continue;
      if (s == 0) st = mzgpu_buf_clear(q->pnext[path]);
      hs[hk] = q->pstream[path];
      ht[hk] = q->spine[q->plan.lookup[path][s]];
      hc[hk] = q->plan.cmp[path][s];
      hcl[hk] = &q->plan.stage[path][s];
      ho[hk] = s == 1 ? q->results : q->pnext[path];
      ++hk;
    }
    if (st == MZGPU_OK && s == 0 && q->peers == 1 && !q->streams_prepared) {
      // build_update_stream + the first half join of every active path in one launch
      mzgpu_batch* hb[3];
      const mzgpu_closure* hi[3];
      uint64_t hskip[3];
      uint32_t j = 0;
      for (int path = 0; path < 3; ++path) {
        if (!active[path]) continue;
        hb[j] = batch[q->plan.source[path]];
        hi[j] = &q->plan.initial[path];
        // as_of rule: only the first relation's path sees the updates at as_of (= 0)
        hskip[j] = path == 0 ? MZGPU_FRONTIER_EMPTY : 0;
        ++j;
      }
      st = mzgpu_delta_first_stage_many(q->ctx, hk, hb, hi, hskip, ht, hc, hcl, ho);
    } else if (st == MZGPU_OK) {
      st = mzgpu_half_join_many(q->ctx, hk, hs, ht, hc, hcl, ho);
    }
    if (s == 0)
      for (int path = 0; path < 3; ++path)
        if (active[path]) std::swap(q->pstream[path], q->pnext[path]);
  }
  if (st == MZGPU_OK && q->peers > 1) {
    PhaseTimer pt(&q->host_ns[5]);
    st = q3_exchange(q, 1, &q->results, &q->xchg, nullptr);
    std::swap(q->results, q->xchg);
  }
  mzgpu_buf* out_buf = q->out;
  if (q->pipelined_out) {
    q->out_parity ^= 1;
    out_buf = q->out_of[q->out_parity];
    if (q->out_pending[q->out_parity]) {  // the output two timestamps back was never fetched: dropped
      st = mzgpu_buf_clear(out_buf);
      q->out_pending[q->out_parity] = false;
    }
  }
  if (st == MZGPU_OK) {
    PhaseTimer pt(&q->host_ns[6]);
    st = mzgpu_reduce_accumulable_buf(q->reduce, q->results, upper, out_buf);
  }
  if (st == MZGPU_OK && q->pipelined_out) {
    H_CUDA(cudaEventRecord(q->ev_out[q->out_parity], q->stream));
    q->out_pending[q->out_parity] = true;
  }
  for (int a = 0; a < 4; ++a)
    if (batch[a]) mzgpu_batch_release(batch[a]);
  q->maintain_upper = upper;
  q->streams_prepared = false;
  return st;
}

// Between-activations maintenance for the timestamp that ran before.
static int32_t q3_maintenance(mzh_q3* q) {
  if (q->maintain_upper == 0) return MZGPU_OK;
  const uint64_t upper = q->maintain_upper, t = upper - 1;
  q->maintain_upper = 0;
  int32_t st = MZGPU_OK;
  for (int a = 0; a < 4 && st == MZGPU_OK; ++a) {
    // idle merge effort first: it looks at the layers as the previous tick left them (their
    // lengths have reached the host since), then the new batches are admitted
    uint64_t e = mzgpu_spine_exert_logic(q->spine[a], 16);
    if (e) st = mzgpu_spine_exert(q->spine[a], e, nullptr);
    if (st == MZGPU_OK) st = mzgpu_spine_set_physical_compaction(q->spine[a], upper);
    if (st == MZGPU_OK) st = mzgpu_spine_set_logical_compaction(q->spine[a], t);
  }
  if (st == MZGPU_OK) st = mzgpu_spine_set_logical_compaction(mzgpu_reduce_input_trace(q->reduce), t);
  return st;
}

extern "C" {

int32_t mzh_q3_new(mzgpu_ctx* ctx, uint64_t seed, uint64_t n_customer, uint64_t n_orders, uint64_t n_part,
                   uint64_t per_batch, uint32_t worker, uint32_t peers, mzh_q3** out) {
  if (ctx == nullptr || out == nullptr) return MZGPU_E_INVALID;
  mzh_q3* q = new mzh_q3();
  q->ctx = ctx;
  q->stream = (cudaStream_t)mzgpu_ctx_stream(ctx);
  q->seed = seed;
  q->sc.n_customer = n_customer;
  q->sc.n_orders = n_orders;
  q->sc.n_part = n_part;
  q->per_batch = per_batch;
  q->worker = worker;
  q->peers = peers;
  mzg_q3_plan_init(&q->plan);
  *out = q;
  for (int a = 0; a < 4; ++a) {
    H_TRY(mzgpu_batcher_new(ctx, MZGPU_ROW_R32, &q->batcher[a]));
    H_TRY(mzgpu_spine_new(ctx, MZGPU_ROW_R32, 1, &q->spine[a]));
    H_TRY(mzgpu_buf_new(ctx, MZGPU_ROW_R32, &q->input[a]));
  }
  H_TRY(mzgpu_reduce_new(ctx, MZGPU_AGG_COUNT_SUM_I64, &q->reduce));
  for (int p = 0; p < 3; ++p) {
    H_TRY(mzgpu_buf_new(ctx, MZGPU_ROW_R32, &q->pstream[p]));
    H_TRY(mzgpu_buf_new(ctx, MZGPU_ROW_R32, &q->pnext[p]));
    H_TRY(mzgpu_buf_new(ctx, MZGPU_ROW_R32, &q->pxchg[p]));
  }
  for (int a = 0; a < 4; ++a) H_TRY(mzgpu_buf_new(ctx, MZGPU_ROW_R32, &q->axchg[a]));
  H_TRY(mzgpu_buf_new(ctx, MZGPU_ROW_R32, &q->results));
  H_TRY(mzgpu_buf_new(ctx, MZGPU_ROW_R32, &q->xchg));
  H_TRY(mzgpu_buf_new(ctx, MZGPU_ROW_ROUT, &q->out));
  return MZGPU_OK;
}

void mzh_q3_free(mzh_q3* q) {
  if (q == nullptr) return;
  if (q->copy_stream) {
    cudaStreamSynchronize(q->copy_stream);
    cudaStreamDestroy(q->copy_stream);
    for (int i = 0; i < 2; ++i) {
      cudaEventDestroy(q->ev_up[i]);
      cudaEventDestroy(q->ev_free[i]);
    }
  }
  for (int a = 0; a < 4; ++a) {
    mzgpu_batcher_free(q->batcher[a]);
    mzgpu_spine_free(q->spine[a]);
    mzgpu_buf_free(q->input[a]);
  }
  mzgpu_reduce_free(q->reduce);
  for (int p = 0; p < 3; ++p) {
    mzgpu_buf_free(q->pstream[p]);
    mzgpu_buf_free(q->pnext[p]);
    mzgpu_buf_free(q->pxchg[p]);
  }
  for (int a = 0; a < 4; ++a) mzgpu_buf_free(q->axchg[a]);
  mzgpu_buf_free(q->results);
  mzgpu_buf_free(q->xchg);
  mzgpu_buf_free(q->out);
  if (q->out2) mzgpu_buf_free(q->out2);
  for (int i = 0; i < 2; ++i)
    if (q->ev_out[i]) cudaEventDestroy(q->ev_out[i]);
  delete q;
}

// Hydration: every base table arrives at time 0.  Each worker generates its
// slice of the source rows (chunked), exchanges them and pushes them into the
// batchers; then the timestamp runs.  *rows_in = rows this worker generated.
int32_t mzh_q3_hydrate(mzh_q3* q, uint64_t* rows_in) {
  if (q == nullptr) return MZGPU_E_INVALID;
  uint64_t total = 0;
  mzgpu_buf* tmp = q->input[0];
  {  // customers
    uint64_t lo = q->sc.n_customer * q->worker / q->peers, hi = q->sc.n_customer * (q->worker + 1) / q->peers;
    DevArr c;
    if (c.alloc((hi - lo) * 32)) return MZGPU_E_CUDA;
    if (hi > lo) {
      k_gen_customers<<<(unsigned)((hi - lo + 255) / 256), 256, 0, q->stream>>>(q->seed, lo, hi - lo,
                                                                              (mzgpu_r32*)c.p);
      H_CUDA(cudaGetLastError());
    }
    H_TRY(mzgpu_buf_upload(tmp, c.p, hi - lo, MZGPU_MEM_DEVICE));
    H_TRY(q3_arrange_push(q, 0, tmp));
    H_CUDA(cudaStreamSynchronize(q->stream));
    total += hi - lo;
  }
  const uint64_t lo = q->sc.n_orders * q->worker / q->peers, hi = q->sc.n_orders * (q->worker + 1) / q->peers;
  const uint64_t CHUNK = 4u << 20;
  // all peers must make the same number of exchange calls
  const uint64_t max_slice = (q->sc.n_orders + q->peers - 1) / q->peers + 1;
  const uint64_t n_chunks = (max_slice + CHUNK - 1) / CHUNK;
  for (uint64_t c = 0; c < n_chunks; ++c) {
    uint64_t first = lo + c * CHUNK;
    uint64_t n = first < hi ? (hi - first < CHUNK ? hi - first : CHUNK) : 0;
    uint64_t n_li = 0;
    H_TRY(q3_gen_orders(q, first, n, 0, 1, 0, &n_li));
    H_TRY(mzgpu_buf_upload(q->input[1], q->gen_ok.p, n, MZGPU_MEM_DEVICE));
    H_TRY(mzgpu_buf_upload(q->input[2], q->gen_ck.p, n, MZGPU_MEM_DEVICE));
    H_TRY(mzgpu_buf_upload(q->input[3], q->gen_li.p, n_li, MZGPU_MEM_DEVICE));
    if (q->peers > 1) {
      // the three relations of a chunk share one exchange round (one host wait, not three)
      mzgpu_buf *ins[3] = {q->input[1], q->input[2], q->input[3]}, *outs[3] = {q->axchg[1], q->axchg[2], q->axchg[3]};
      H_TRY(mzgpu_exchange_many(q->ctx, 3, ins, outs));
      for (int a = 1; a < 4; ++a) H_TRY(mzgpu_batcher_push_buf(q->batcher[a], q->axchg[a]));
    } else {
      for (int a = 1; a < 4; ++a) H_TRY(mzgpu_batcher_push_buf(q->batcher[a], q->input[a]));
    }
    total += n + n_li;
  }
  for (int a = 0; a < 4; ++a) H_TRY(mzgpu_buf_clear(q->input[a]));
  if (rows_in) *rows_in = total;
  H_TRY(q3_run_timestamp(q, 0));
  H_TRY(q3_maintenance(q));
  q->next_time = 1;
  return MZGPU_OK;
}

// Stage update batch `b` (this worker's share of the tick's order replacements)
// in device memory: not part of the timed step.  *rows_in = staged update rows
// (orders counted once, as in the oracle).
int32_t mzh_q3_stage_batch(mzh_q3* q, uint64_t b, uint64_t t, uint64_t* rows_in) {
  if (q == nullptr) return MZGPU_E_INVALID;
  const uint64_t x0 = b * q->per_batch;
  const uint64_t lo = x0 + q->per_batch * q->worker / q->peers, hi = x0 + q->per_batch * (q->worker + 1) / q->peers;
  uint64_t n_li = 0;
  H_TRY(q3_gen_orders(q, lo, hi - lo, 1, 2, t, &n_li));
  H_TRY(mzgpu_buf_clear(q->input[0]));
  H_TRY(mzgpu_buf_upload(q->input[1], q->gen_ok.p, 2 * (hi - lo), MZGPU_MEM_DEVICE));
  H_TRY(mzgpu_buf_upload(q->input[2], q->gen_ck.p, 2 * (hi - lo), MZGPU_MEM_DEVICE));
  H_TRY(mzgpu_buf_upload(q->input[3], q->gen_li.p, n_li, MZGPU_MEM_DEVICE));
  q->last_rows_in = 2 * (hi - lo) + n_li;
  if (rows_in) *rows_in = q->last_rows_in;
  return mzgpu_ctx_sync(q->ctx);
}

// Stage host rows for arrangement `a` of the NEXT timestamp (the end-to-end path:
// the H2D copy is issued here, inside the caller's timed region, on a copy stream
// into the staging slot the next mzh_q3_step consumes; call mzh_q3_stage_commit
// after the last arrangement of the batch).
int32_t mzh_q3_stage_host(mzh_q3* q, int32_t a, const mzgpu_r32* rows, uint64_t n) {
  if (q == nullptr || a < 0 || a > 3) return MZGPU_E_INVALID;
  const int s = q->fill_slot;
  if (q->copy_stream == nullptr) {
    H_CUDA(cudaStreamCreateWithFlags(&q->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      H_CUDA(cudaEventCreateWithFlags(&q->ev_up[i], cudaEventDisableTiming));
      H_CUDA(cudaEventCreateWithFlags(&q->ev_free[i], cudaEventDisableTiming));
    }
  }
  if (q->slot_full[s]) return MZGPU_E_INVALID;  // both slots staged and not yet stepped
  bool any = false;
  for (int i = 0; i < 4; ++i) any = any || q->slot_touched[s][i];
  if (!any) {
    // first arrangement of this batch: the slot's previous contents must have been consumed
    H_CUDA(cudaStreamWaitEvent(q->copy_stream, q->ev_free[s], 0));
    for (int i = 0; i < 4; ++i) q->stage_n[s][i] = 0;
  }
  if (n > q->stage_cap[s][a]) {
    H_CUDA(cudaStreamSynchronize(q->stream));  // growing: rare, after the first batches never
    if (q->stage[s][a].alloc(n * 32 * 5 / 4)) return MZGPU_E_CUDA;
    q->stage_cap[s][a] = n * 5 / 4;
  }
  if (n) H_CUDA(cudaMemcpyAsync(q->stage[s][a].p, rows, n * 32, cudaMemcpyHostToDevice, q->copy_stream));
  q->stage_n[s][a] = n;
  q->slot_touched[s][a] = true;
  q->h2d_bytes += n * 32;
  return MZGPU_OK;
}
// The batch staged with mzh_q3_stage_host is complete.
int32_t mzh_q3_stage_commit(mzh_q3* q) {
  if (q == nullptr || q->copy_stream == nullptr) return MZGPU_E_INVALID;
  const int s = q->fill_slot;
  H_CUDA(cudaEventRecord(q->ev_up[s], q->copy_stream));
  q->slot_full[s] = true;
  for (int i = 0; i < 4; ++i) q->slot_touched[s][i] = false;
  q->fill_slot ^= 1;
  return MZGPU_OK;
}
uint64_t mzh_q3_h2d_bytes(mzh_q3* q) { return q ? q->h2d_bytes : 0; }

// Stage rows that already live in device memory (a D2D copy on the ctx stream).
int32_t mzh_q3_stage_device(mzh_q3* q, int32_t a, const mzgpu_r32* d_rows, uint64_t n) {
  if (q == nullptr || a < 0 || a > 3) return MZGPU_E_INVALID;
  return mzgpu_buf_upload(q->input[a], d_rows, n, MZGPU_MEM_DEVICE);
}
mzgpu_buf* mzh_q3_input(mzh_q3* q, int32_t a) { return (q && a >= 0 && a < 4) ? q->input[a] : nullptr; }

// Copy the staged device inputs of arrangement `a` to the host (to build the
// host-resident copies for the end-to-end measurement and for parity tests).
int32_t mzh_q3_staged(mzh_q3* q, int32_t a, mzgpu_r32* rows, uint64_t cap, uint64_t* n) {
  if (q == nullptr || a < 0 || a > 3) return MZGPU_E_INVALID;
  return mzgpu_buf_download(q->input[a], rows, cap, MZGPU_MEM_HOST, n);
}

// One timestamp over the staged inputs (inputs already resident in HBM).  The
// device work is enqueued and the call returns; the previous timestamp's
// maintenance runs first (its counts have reached the host by now).
int32_t mzh_q3_step(mzh_q3* q) {
  if (q == nullptr) return MZGPU_E_INVALID;
  PhaseTimer pt_all(&q->host_ns[7]);
  std::unique_ptr<PhaseTimer> pt_in(new PhaseTimer(&q->host_ns[0]));
  q->stepping = true;
  if (q->slot_full[q->run_slot]) {
    // a committed host batch: the ctx stream waits for its H2D copies, takes the
    // rows over (device copies), and frees the slot for the batch after next
    const int s = q->run_slot;
    H_CUDA(cudaStreamWaitEvent(q->stream, q->ev_up[s], 0));
    for (int a = 0; a < 4; ++a)
      H_TRY(mzgpu_buf_upload(q->input[a], q->stage[s][a].p, q->stage_n[s][a], MZGPU_MEM_DEVICE));
    H_CUDA(cudaEventRecord(q->ev_free[s], q->stream));
    q->slot_full[s] = false;
    q->run_slot ^= 1;
  }
  const uint64_t t = q->next_time;
  if (q->peers > 1) {
    // the arrangement inputs of one timestamp share one exchange round
    mzgpu_buf *ins[8], *outs[8];
    uint64_t recv_ub[8];
    // what one worker can receive of a relation's updates: all of the (global) batch -- two
    // versions of every replaced order, at most seven lineitems each
    const uint64_t ub_rel[4] = {0, 2 * q->per_batch, 2 * q->per_batch, 14 * q->per_batch};
    uint32_t k = 0;
    for (int a = 0; a < 4; ++a)
      if (!q->static_rel[a]) {
        ins[k] = q->input[a];
        outs[k] = q->axchg[a];
        recv_ub[k] = ub_rel[a];
        ++k;
      }
    // ... and so do the delta paths' update streams: build_update_stream is a per-row map, so
    // it can run on the raw input rows before they are exchanged (the multiset of stream rows
    // is the one the consolidated batch would give) — one exchange round fewer per timestamp
    uint32_t first_stream = k;
    for (int path = 0; path < 3; ++path) {
      const int src = q->plan.source[path];
      if (q->static_rel[src]) continue;
      H_TRY(mzgpu_buf_clear(q->pstream[path]));
      H_TRY(mzgpu_map_rows(q->ctx, (const mzgpu_r32*)mzgpu_buf_device_ptr(q->input[src]), mzgpu_buf_len(q->input[src]),
                           MZGPU_MEM_DEVICE, &q->plan.initial[path], q->pstream[path]));
      ins[k] = q->pstream[path];
      outs[k] = q->pxchg[path];
      recv_ub[k] = ub_rel[src];
      ++k;
    }
    H_TRY(q3_exchange(q, k, ins, outs, recv_ub));
    for (int a = 0; a < 4; ++a)
      if (!q->static_rel[a]) H_TRY(mzgpu_batcher_push_buf(q->batcher[a], q->axchg[a]));
    for (int path = 0; path < 3; ++path)
      if (!q->static_rel[q->plan.source[path]]) std::swap(q->pstream[path], q->pxchg[path]);
    q->streams_prepared = first_stream < k;
  } else {
    for (int a = 0; a < 4; ++a) H_TRY(q3_arrange_push(q, a, q->input[a]));
  }
  pt_in.reset();
  H_TRY(q3_run_timestamp(q, t));
  q->next_time = t + 1;
  return MZGPU_OK;
}
// Host nanoseconds spent in the phases of mzh_q3_step so far (see mzh_q3::host_ns).
int32_t mzh_q3_host_ns(mzh_q3* q, uint64_t out[8]) {
  if (q == nullptr || out == nullptr) return MZGPU_E_INVALID;
  for (int i = 0; i < 8; ++i) out[i] = q->host_ns[i];
  return MZGPU_OK;
}
// Update-batch exchange rounds over peer memory from now on (the caller has connected the landing
// zones: mzgpu_comm_p2p_export / _import on q's context).
int32_t mzh_q3_use_p2p(mzh_q3* q, int32_t on) {
  if (q == nullptr) return MZGPU_E_INVALID;
  q->use_p2p = on != 0;
  return MZGPU_OK;
}
// Run the maintenance that is due (tests call this before inspecting the spines).
int32_t mzh_q3_maintain(mzh_q3* q) { return q ? q3_maintenance(q) : MZGPU_E_INVALID; }

// Output corrections (ROUT rows) accumulated since the last clear.
mzgpu_buf* mzh_q3_out(mzh_q3* q) { return q ? q->out : nullptr; }

// Pipelined result read-back: from now on timestamps alternate between two output buffers.
int32_t mzh_q3_pipeline_out(mzh_q3* q) {
  if (q == nullptr) return MZGPU_E_INVALID;
  if (q->pipelined_out) return MZGPU_OK;
  if (q->copy_stream == nullptr) H_CUDA(cudaStreamCreateWithFlags(&q->copy_stream, cudaStreamNonBlocking));
  if (q->ev_up[0] == nullptr)
    for (int i = 0; i < 2; ++i) {
      H_CUDA(cudaEventCreateWithFlags(&q->ev_up[i], cudaEventDisableTiming));
      H_CUDA(cudaEventCreateWithFlags(&q->ev_free[i], cudaEventDisableTiming));
    }
  H_TRY(mzgpu_buf_new(q->ctx, MZGPU_ROW_ROUT, &q->out2));
  for (int i = 0; i < 2; ++i) H_CUDA(cudaEventCreateWithFlags(&q->ev_out[i], cudaEventDisableTiming));
  q->out_of[0] = q->out;
  q->out_of[1] = q->out2;
  H_TRY(mzgpu_buf_clear(q->out));
  q->pipelined_out = true;
  return MZGPU_OK;
}
// Copy out the corrections of a finished timestamp: `which` = 0 the one before the timestamp
// enqueued last (its kernels have long finished; nothing newer is waited for), 1 the latest
// (drains the stream's tail).  The copy runs on the copy stream behind the event recorded when
// that timestamp's reduce was enqueued.  *n = rows copied (0 if nothing is pending).
int32_t mzh_q3_fetch_out(mzh_q3* q, int32_t which, mzgpu_rout* rows, uint64_t cap, uint64_t* n) {
  if (q == nullptr || !q->pipelined_out || n == nullptr) return MZGPU_E_INVALID;
  const int p = which == 0 ? (q->out_parity ^ 1) : q->out_parity;
  *n = 0;
  if (!q->out_pending[p]) return MZGPU_OK;
  mzgpu_buf* b = q->out_of[p];
  // the length travelled to the host with the read-back of the timestamp that followed (or is
  // read back now, for the latest timestamp)
  const uint64_t len = mzgpu_buf_len(b);
  if (len > cap) return MZGPU_E_CAPACITY;
  H_CUDA(cudaStreamWaitEvent(q->copy_stream, q->ev_out[p], 0));
  if (len) H_CUDA(cudaMemcpyAsync(rows, mzgpu_buf_device_ptr(b), len * sizeof(mzgpu_rout), cudaMemcpyDeviceToHost, q->copy_stream));
  H_CUDA(cudaStreamSynchronize(q->copy_stream));
  q->d2h_bytes += len * sizeof(mzgpu_rout);
  *n = len;
  q->out_pending[p] = false;
  return mzgpu_buf_clear(b);
}
uint64_t mzh_q3_d2h_bytes(mzh_q3* q) { return q ? q->d2h_bytes : 0; }
int32_t mzh_q3_clear_out(mzh_q3* q) { return q ? mzgpu_buf_clear(q->out) : MZGPU_E_INVALID; }
uint64_t mzh_q3_time(mzh_q3* q) { return q ? q->next_time : 0; }
mzgpu_spine* mzh_q3_spine(mzh_q3* q, int32_t a) { return (q && a >= 0 && a < 4) ? q->spine[a] : nullptr; }

// ---- device-side generators for the other BASELINE configs
int32_t mzh_gen_cfg1(mzgpu_ctx* ctx, uint64_t seed, uint64_t first, uint64_t n, uint32_t key_bits,
                     mzgpu_buf* out) {
  DevArr d;
  if (d.alloc(n * 16)) return MZGPU_E_CUDA;
  cudaStream_t s = (cudaStream_t)mzgpu_ctx_stream(ctx);
  if (n) k_gen_cfg1<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(seed, first, n, key_bits, (mzgpu_r16*)d.p);
  H_CUDA(cudaGetLastError());
  H_TRY(mzgpu_buf_upload(out, d.p, n, MZGPU_MEM_DEVICE));
  return mzgpu_ctx_sync(ctx);
}
int32_t mzh_gen_cfg2(mzgpu_ctx* ctx, uint64_t seed, uint64_t first, uint64_t n, uint64_t n_keys,
                     mzgpu_buf* out) {
  DevArr d;
  if (d.alloc(n * 32)) return MZGPU_E_CUDA;
  cudaStream_t s = (cudaStream_t)mzgpu_ctx_stream(ctx);
  if (n) k_gen_cfg2<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(seed, first, n, n_keys, (mzgpu_r32*)d.p);
  H_CUDA(cudaGetLastError());
  H_TRY(mzgpu_buf_upload(out, d.p, n, MZGPU_MEM_DEVICE));
  return mzgpu_ctx_sync(ctx);
}
// `cdf` is the host-built zipf inverse-CDF table (n_keys doubles)
int32_t mzh_gen_cfg4(mzgpu_ctx* ctx, uint64_t seed, uint64_t first, uint64_t n, const double* cdf,
                     uint64_t n_keys, int32_t as_f64, uint64_t t, int64_t diff, mzgpu_buf* out) {
  DevArr d, dc;
  if (d.alloc(n * 32) || dc.alloc(n_keys * 8)) return MZGPU_E_CUDA;
  cudaStream_t s = (cudaStream_t)mzgpu_ctx_stream(ctx);
  H_CUDA(cudaMemcpyAsync(dc.p, cdf, n_keys * 8, cudaMemcpyHostToDevice, s));
  if (n)
    k_gen_cfg4<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(seed, first, n, (const double*)dc.p, n_keys,
                                                           as_f64, t, diff, (mzgpu_r32*)d.p);
  H_CUDA(cudaGetLastError());
  H_TRY(mzgpu_buf_upload(out, d.p, n, MZGPU_MEM_DEVICE));
  return mzgpu_ctx_sync(ctx);
}

}  // extern "C"
