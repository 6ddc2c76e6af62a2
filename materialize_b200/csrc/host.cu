// host.cu — the C ABI of libmzgpu (include/mzgpu.h): contexts, row buffers,
// batcher, batches, the fueled spine, join_core / half_join / reduce operator
// state, and the NCCL exchange.  Host-side bookkeeping only; every per-row
// operation is a CUDA kernel from the sibling .cu files.  There is no CPU
// fallback: without a CUDA device every entry point fails with MZGPU_E_CUDA.
#include <dlfcn.h>

#include <algorithm>
#include <deque>
#include <memory>
#include <new>

#include "common.cuh"

// ====================================================================== ctx
extern "C" int32_t mzgpu_ctx_create(int32_t device, int32_t worker_index, int32_t peers,
                                    mzgpu_ctx** out) {
  if (out == nullptr || peers < 1 || worker_index < 0 || worker_index >= peers) return MZGPU_E_INVALID;
  *out = nullptr;
  mzgpu_ctx* ctx = new (std::nothrow) mzgpu_ctx();
  if (ctx == nullptr) return MZGPU_E_INVALID;
  memset(&ctx->stats, 0, sizeof(ctx->stats));
  ctx->device = device;
  ctx->worker = worker_index;
  ctx->peers = peers;
  *out = ctx;  // returned even on failure so the caller can read mzgpu_last_error
  MZ_CUDA(ctx, cudaSetDevice(device));
  cudaDeviceProp prop;
  MZ_CUDA(ctx, cudaGetDeviceProperties(&prop, device));
  ctx->num_sms = prop.multiProcessorCount;
  MZ_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  MZ_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev, cudaEventDisableTiming));
  MZ_CUDA(ctx, cudaMallocHost((void**)&ctx->h_scratch, 128 * 8));
  MZ_CUDA(ctx, cudaMalloc((void**)&ctx->d_scratch, 128 * 8));
  MZ_CUDA(ctx, cudaMallocHost(&ctx->h_fused, 16384));
  // keep freed blocks cached in the stream-ordered pool
  cudaMemPool_t pool;
  MZ_CUDA(ctx, cudaDeviceGetDefaultMemPool(&pool, device));
  uint64_t thresh = UINT64_MAX;
  MZ_CUDA(ctx, cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
  return MZGPU_OK;
}

extern "C" void mzgpu_ctx_destroy(mzgpu_ctx* ctx) {
  if (ctx == nullptr) return;
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->nccl_comm && ctx->nccl_lib) {
    typedef int (*destroy_t)(void*);
    destroy_t f = (destroy_t)dlsym(ctx->nccl_lib, "ncclCommDestroy");
    if (f) f(ctx->nccl_comm);
  }
  if (ctx->h_scratch) cudaFreeHost(ctx->h_scratch);
  if (ctx->h_fused) cudaFreeHost(ctx->h_fused);
  if (ctx->d_scratch) cudaFree(ctx->d_scratch);
  if (ctx->ev) cudaEventDestroy(ctx->ev);
  for (auto& r : ctx->prof) {
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  for (auto e : ctx->ev_pool) cudaEventDestroy(e);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" const char* mzgpu_last_error(mzgpu_ctx* ctx) {
  return ctx ? ctx->last_error.c_str() : "null context";
}

extern "C" int32_t mzgpu_ctx_sync(mzgpu_ctx* ctx) {
  MZ_CHECK_CTX(ctx);
  MZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return MZGPU_OK;
}

extern "C" int32_t mzgpu_ctx_stats(mzgpu_ctx* ctx, mzgpu_stats* out) {
  if (ctx == nullptr || out == nullptr) return MZGPU_E_INVALID;
  *out = ctx->stats;
  return MZGPU_OK;
}

extern "C" int32_t mzgpu_profile_enable(mzgpu_ctx* ctx, int32_t on) {
  MZ_CHECK_CTX(ctx);
  ctx->profile = on != 0;
  return MZGPU_OK;
}

extern "C" int32_t mzgpu_profile_report(mzgpu_ctx* ctx, char* buf, uint64_t cap) {
  MZ_CHECK_CTX(ctx);
  if (buf == nullptr || cap == 0) return MZGPU_E_INVALID;
  MZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  struct Agg {
    std::string name;
    u64 launches = 0, bytes = 0;
    double ms = 0;
  };
  std::vector<Agg> aggs;
  for (auto& r : ctx->prof) {
    float ms = 0;
    cudaEventElapsedTime(&ms, r.e0, r.e1);
    Agg* a = nullptr;
    for (auto& x : aggs)
      if (x.name == r.name) a = &x;
    if (a == nullptr) {
      aggs.push_back(Agg());
      a = &aggs.back();
      a->name = r.name;
    }
    a->launches++;
    a->bytes += r.bytes;
    a->ms += ms;
    ctx->ev_pool.push_back(r.e0);
    ctx->ev_pool.push_back(r.e1);
  }
  ctx->prof.clear();
  std::string out;
  for (auto& a : aggs) {
    char line[384];
    std::string nm = a.name;
    for (auto& ch : nm)
      if (ch == ' ') ch = '_';
    snprintf(line, sizeof(line), "%s %llu %.6f %llu\n", nm.c_str(), (unsigned long long)a.launches, a.ms,
             (unsigned long long)a.bytes);
    out += line;
  }
  if (out.size() + 1 > cap) return MZGPU_E_CAPACITY;
  memcpy(buf, out.c_str(), out.size() + 1);
  return MZGPU_OK;
}

extern "C" void* mzgpu_ctx_stream(mzgpu_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

// ------------------------------------------------------------- transfers
static int32_t copy_in(mzgpu_ctx* ctx, void* d_dst, const void* src, size_t bytes, int32_t mem) {
  if (bytes == 0) return MZGPU_OK;
  if (mem == MZGPU_MEM_HOST) {
    MZ_CUDA(ctx, cudaMemcpyAsync(d_dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    ctx->stats.h2d_bytes += bytes;
  } else {
    MZ_CUDA(ctx, cudaMemcpyAsync(d_dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  }
  return MZGPU_OK;
}
static int32_t copy_out(mzgpu_ctx* ctx, void* dst, const void* d_src, size_t bytes, int32_t mem) {
  if (bytes == 0) return MZGPU_OK;
  if (mem == MZGPU_MEM_HOST) {
    MZ_CUDA(ctx, cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    MZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stats.d2h_bytes += bytes;
  } else {
    MZ_CUDA(ctx, cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  }
  return MZGPU_OK;
}

static bool valid_row_bytes(uint32_t rb) { return rb == 16 || rb == 32 || rb == 40 || rb == 80 || rb == 64; }

// ====================================================================== buf
struct mzgpu_buf {
  mzgpu_ctx* ctx;
  uint32_t rb;
  DevMem mem;
  u64 len = 0;
  u64 cap = 0;
};

static int32_t buf_reserve(mzgpu_buf* b, u64 n, bool keep) {
  if (n <= b->cap) return MZGPU_OK;
  u64 ncap = std::max<u64>(n, b->cap * 2);
  DevMem m;
  MZ_TRY(m.alloc(b->ctx, ncap * b->rb));
  if (keep && b->len) MZ_TRY(copy_in(b->ctx, m.p, b->mem.p, b->len * b->rb, MZGPU_MEM_DEVICE));
  b->mem = std::move(m);
  b->cap = ncap;
  return MZGPU_OK;
}
// take ownership of a device array as the buffer contents
static void buf_adopt(mzgpu_buf* b, DevMem&& m, u64 len) {
  b->cap = m.bytes / b->rb;
  b->mem = std::move(m);
  b->len = len;
}
static int32_t buf_append_dev(mzgpu_buf* b, const void* d_rows, u64 n) {
  if (n == 0) return MZGPU_OK;
  MZ_TRY(buf_reserve(b, b->len + n, true));
  MZ_TRY(copy_in(b->ctx, (char*)b->mem.p + b->len * b->rb, d_rows, n * b->rb, MZGPU_MEM_DEVICE));
  b->len += n;
  return MZGPU_OK;
}

extern "C" int32_t mzgpu_buf_new(mzgpu_ctx* ctx, uint32_t row_bytes, mzgpu_buf** out) {
  MZ_CHECK_CTX(ctx);
  if (out == nullptr || !valid_row_bytes(row_bytes)) return MZGPU_E_INVALID;
  mzgpu_buf* b = new mzgpu_buf();
  b->ctx = ctx;
  b->rb = row_bytes;
  *out = b;
  return MZGPU_OK;
}
extern "C" void mzgpu_buf_free(mzgpu_buf* b) { delete b; }
extern "C" uint64_t mzgpu_buf_len(const mzgpu_buf* b) { return b ? b->len : 0; }
extern "C" uint32_t mzgpu_buf_row_bytes(const mzgpu_buf* b) { return b ? b->rb : 0; }
extern "C" void* mzgpu_buf_device_ptr(mzgpu_buf* b) { return b ? b->mem.p : nullptr; }
extern "C" int32_t mzgpu_buf_clear(mzgpu_buf* b) {
  if (b == nullptr) return MZGPU_E_INVALID;
  b->len = 0;
  return MZGPU_OK;
}
extern "C" int32_t mzgpu_buf_upload(mzgpu_buf* b, const void* rows, uint64_t n, int32_t mem) {
  if (b == nullptr || (rows == nullptr && n)) return MZGPU_E_INVALID;
  MZ_CHECK_CTX(b->ctx);
  b->len = 0;
  MZ_TRY(buf_reserve(b, n, false));
  MZ_TRY(copy_in(b->ctx, b->mem.p, rows, n * b->rb, mem));
  b->len = n;
  b->ctx->stats.rows_in += n;
  return MZGPU_OK;
}
extern "C" int32_t mzgpu_buf_append(mzgpu_buf* b, const void* rows, uint64_t n, int32_t mem) {
  if (b == nullptr || (rows == nullptr && n)) return MZGPU_E_INVALID;
  MZ_CHECK_CTX(b->ctx);
  MZ_TRY(buf_reserve(b, b->len + n, true));
  MZ_TRY(copy_in(b->ctx, (char*)b->mem.p + b->len * b->rb, rows, n * b->rb, mem));
  b->len += n;
  b->ctx->stats.rows_in += n;
  return MZGPU_OK;
}
extern "C" int32_t mzgpu_buf_download(mzgpu_buf* b, void* rows, uint64_t cap, int32_t mem,
                                      uint64_t* n_out) {
  if (b == nullptr) return MZGPU_E_INVALID;
  MZ_CHECK_CTX(b->ctx);
  if (n_out) *n_out = b->len;
  if (cap < b->len) {
    MZ_SET_ERR(b->ctx, "buf_download: capacity %llu < %llu rows", (unsigned long long)cap,
               (unsigned long long)b->len);
    return MZGPU_E_CAPACITY;
  }
  MZ_TRY(copy_out(b->ctx, rows, b->mem.p, b->len * b->rb, mem));
  b->ctx->stats.rows_out += b->len;
  return MZGPU_OK;
}

// ============================================================ consolidation
static int32_t consolidate_ptr(mzgpu_ctx* ctx, int rb, void* rows, u64 n, int32_t mem, u64* n_out) {
  MZ_CHECK_CTX(ctx);
  if ((rows == nullptr && n) || n_out == nullptr) return MZGPU_E_INVALID;
  *n_out = 0;
  if (n == 0) return MZGPU_OK;
  ctx->stats.rows_in += n;
  DevMem in, out;
  const void* d_in = rows;
  if (mem == MZGPU_MEM_HOST) {
    MZ_TRY(in.alloc(ctx, n * rb));
    MZ_TRY(copy_in(ctx, in.p, rows, n * rb, mem));
    d_in = in.p;
  }
  MZ_TRY(mz_sort_consolidate(ctx, rb, d_in, n, &out, n_out));
  MZ_TRY(copy_out(ctx, rows, out.p, *n_out * rb, mem));
  ctx->stats.rows_out += *n_out;
  return MZGPU_OK;
}
extern "C" int32_t mzgpu_consolidate_r16(mzgpu_ctx* ctx, mzgpu_r16* rows, uint64_t n, int32_t mem,
                                         uint64_t* n_out) {
  return consolidate_ptr(ctx, 16, rows, n, mem, n_out);
}
extern "C" int32_t mzgpu_consolidate_r32(mzgpu_ctx* ctx, mzgpu_r32* rows, uint64_t n, int32_t mem,
                                         uint64_t* n_out) {
  return consolidate_ptr(ctx, 32, rows, n, mem, n_out);
}
extern "C" int32_t mzgpu_buf_consolidate(mzgpu_buf* b) {
  if (b == nullptr) return MZGPU_E_INVALID;
  MZ_CHECK_CTX(b->ctx);
  if (b->len == 0) return MZGPU_OK;
  DevMem out;
  u64 n_out = 0;
  MZ_TRY(mz_sort_consolidate(b->ctx, b->rb, b->mem.p, b->len, &out, &n_out));
  buf_adopt(b, std::move(out), n_out);
  return MZGPU_OK;
}

// ================================================================== batches
struct mzgpu_batch {
  mzgpu_ctx* ctx;
  uint32_t rb;
  DevMem rows;
  u64 len = 0;
  u64 n_keys = 0;
  DevMem table;
  u64 slots = 0;
  mzgpu_desc desc;
  int refs = 1;
};

// sorted + consolidated device rows -> indexed immutable batch
static int32_t make_batch(mzgpu_ctx* ctx, uint32_t rb, DevMem&& rows, u64 len, mzgpu_desc desc,
                          mzgpu_batch** out) {
  std::unique_ptr<mzgpu_batch> b(new mzgpu_batch());
  b->ctx = ctx;
  b->rb = rb;
  b->rows = std::move(rows);
  b->len = len;
  b->desc = desc;
  MZ_TRY(mz_count_keys(ctx, rb, b->rows.p, len, &b->n_keys));
  MZ_TRY(mz_build_index(ctx, rb, b->rows.p, len, b->n_keys, &b->table, &b->slots));
  *out = b.release();
  return MZGPU_OK;
}
// rows + index already built (fused path)
static int32_t make_batch_indexed(mzgpu_ctx* ctx, uint32_t rb, DevMem&& rows, u64 len, DevMem&& table,
                                  u64 slots, u64 n_keys, mzgpu_desc desc, mzgpu_batch** out) {
  mzgpu_batch* b = new mzgpu_batch();
  b->ctx = ctx;
  b->rb = rb;
  b->rows = std::move(rows);
  b->len = len;
  b->table = std::move(table);
  b->slots = slots;
  b->n_keys = n_keys;
  b->desc = desc;
  *out = b;
  return MZGPU_OK;
}
// unsorted device rows -> batch: the fused kernel for small inputs, else the
// multi-kernel path
static int32_t build_batch_from_unsorted(mzgpu_ctx* ctx, uint32_t rb, const void* d_in, u64 n,
                                         mzgpu_desc desc, mzgpu_batch** out) {
  if (n > 0 && n <= MZ_FUSED_MAX_ROWS) {
    FusedResult fr;
    MZ_TRY(mz_fused_sort_consolidate(ctx, rb, d_in, n, true, &fr));
    if (!fr.fallback)
      return make_batch_indexed(ctx, rb, std::move(fr.rows), fr.n_out, std::move(fr.table), fr.slots,
                                fr.n_keys, desc, out);
  }
  DevMem cons;
  u64 n_out = 0;
  MZ_TRY(mz_sort_consolidate(ctx, rb, d_in, n, &cons, &n_out));
  return make_batch(ctx, rb, std::move(cons), n_out, desc, out);
}
static int32_t make_empty_batch(mzgpu_ctx* ctx, uint32_t rb, mzgpu_desc desc, mzgpu_batch** out) {
  DevMem rows;
  MZ_TRY(rows.alloc(ctx, 16));
  return make_batch(ctx, rb, std::move(rows), 0, desc, out);
}

extern "C" int32_t mzgpu_batch_build(mzgpu_ctx* ctx, uint32_t row_bytes, const void* rows, uint64_t n,
                                     int32_t mem, mzgpu_desc desc, mzgpu_batch** out) {
  MZ_CHECK_CTX(ctx);
  if (out == nullptr || (rows == nullptr && n) || (row_bytes != 32 && row_bytes != 80)) return MZGPU_E_INVALID;
  DevMem in, cons;
  const void* d_in = rows;
  if (mem == MZGPU_MEM_HOST && n) {
    MZ_TRY(in.alloc(ctx, n * row_bytes));
    MZ_TRY(copy_in(ctx, in.p, rows, n * row_bytes, mem));
    d_in = in.p;
  }
  ctx->stats.rows_in += n;
  return build_batch_from_unsorted(ctx, row_bytes, d_in, n, desc, out);
}
extern "C" uint64_t mzgpu_batch_len(const mzgpu_batch* b) { return b ? b->len : 0; }
extern "C" uint64_t mzgpu_batch_keys(const mzgpu_batch* b) { return b ? b->n_keys : 0; }
extern "C" mzgpu_desc mzgpu_batch_desc(const mzgpu_batch* b) {
  mzgpu_desc d = {0, 0, 0};
  return b ? b->desc : d;
}
extern "C" void mzgpu_batch_retain(mzgpu_batch* b) {
  if (b) b->refs++;
}
extern "C" void mzgpu_batch_release(mzgpu_batch* b) {
  if (b && --b->refs == 0) delete b;
}
extern "C" int32_t mzgpu_batch_export(mzgpu_batch* b, void* rows, uint64_t cap, int32_t mem,
                                      uint64_t* n_out) {
  if (b == nullptr) return MZGPU_E_INVALID;
  MZ_CHECK_CTX(b->ctx);
  if (n_out) *n_out = b->len;
  if (cap < b->len) return MZGPU_E_CAPACITY;
  MZ_TRY(copy_out(b->ctx, rows, b->rows.p, b->len * b->rb, mem));
  return MZGPU_OK;
}
static int32_t merge_batches(mzgpu_batch* b1, mzgpu_batch* b2, u64 since, mzgpu_batch** out) {
  mzgpu_ctx* ctx = b1->ctx;
  DevMem merged;
  u64 n_out = 0;
  MZ_TRY(mz_merge_consolidate(ctx, b1->rb, b1->rows.p, b1->len, b2->rows.p, b2->len, since, &merged,
                              &n_out));
  mzgpu_desc d = {b1->desc.lower, b2->desc.upper, since};
  return make_batch(ctx, b1->rb, std::move(merged), n_out, d, out);
}
extern "C" int32_t mzgpu_batch_merge(mzgpu_batch* b1, mzgpu_batch* b2, uint64_t since,
                                     mzgpu_batch** out) {
  if (b1 == nullptr || b2 == nullptr || out == nullptr || b1->rb != b2->rb) return MZGPU_E_INVALID;
  MZ_CHECK_CTX(b1->ctx);
  if (b1->desc.upper != b2->desc.lower) {
    MZ_SET_ERR(b1->ctx, "batch_merge: b1.upper %llu != b2.lower %llu", (unsigned long long)b1->desc.upper,
               (unsigned long long)b2->desc.lower);
    return MZGPU_E_FRONTIER;
  }
  return merge_batches(b1, b2, since, out);
}

// ================================================================== batcher
struct Chain {
  DevMem rows;
  u64 len = 0;
  // known only when the chain came out of the fused kernel (or merges of such)
  bool time_known = false;
  u64 max_time = 0;
  // hash index built by the fused kernel (valid while the chain is unmerged)
  bool has_index = false;
  DevMem table;
  u64 slots = 0, n_keys = 0;
};
struct mzgpu_batcher {
  mzgpu_ctx* ctx;
  uint32_t rb;
  std::vector<Chain> chains;
  u64 lower = 0;
  u64 frontier = MZGPU_FRONTIER_EMPTY;
};

static int32_t batcher_merge_top(mzgpu_batcher* b) {
  Chain newer = std::move(b->chains.back());
  b->chains.pop_back();
  Chain older = std::move(b->chains.back());
  b->chains.pop_back();
  Chain m;
  MZ_TRY(mz_merge_consolidate(b->ctx, b->rb, older.rows.p, older.len, newer.rows.p, newer.len, 0,
                              &m.rows, &m.len));
  m.time_known = older.time_known && newer.time_known;
  m.max_time = std::max(older.max_time, newer.max_time);
  b->chains.push_back(std::move(m));
  return MZGPU_OK;
}
// MergeBatcher::insert_chain: keep chain lengths geometric
static int32_t batcher_insert_chain(mzgpu_batcher* b, Chain&& c) {
  if (c.len == 0) return MZGPU_OK;
  b->chains.push_back(std::move(c));
  while (b->chains.size() > 1 &&
         b->chains[b->chains.size() - 1].len >= b->chains[b->chains.size() - 2].len / 2)
    MZ_TRY(batcher_merge_top(b));
  return MZGPU_OK;
}
static int32_t batcher_push_dev(mzgpu_batcher* b, const void* d_rows, u64 n) {
  if (n == 0) return MZGPU_OK;
  Chain c;
  bool done = false;
  if (n <= MZ_FUSED_MAX_ROWS) {
    FusedResult fr;
    MZ_TRY(mz_fused_sort_consolidate(b->ctx, b->rb, d_rows, n, true, &fr));
    if (!fr.fallback) {
      c.rows = std::move(fr.rows);
      c.len = fr.n_out;
      c.time_known = true;
      c.max_time = fr.max_time;
      c.has_index = true;
      c.table = std::move(fr.table);
      c.slots = fr.slots;
      c.n_keys = fr.n_keys;
      done = true;
    }
  }
  if (!done) MZ_TRY(mz_sort_consolidate(b->ctx, b->rb, d_rows, n, &c.rows, &c.len));
  return batcher_insert_chain(b, std::move(c));
}
static int32_t batcher_seal(mzgpu_batcher* b, u64 upper, mzgpu_batch** batch_out, u64* new_lower) {
  mzgpu_ctx* ctx = b->ctx;
  if (upper != MZGPU_FRONTIER_EMPTY && upper < b->lower) {
    MZ_SET_ERR(ctx, "batcher_seal: upper %llu precedes lower %llu", (unsigned long long)upper,
               (unsigned long long)b->lower);
    return MZGPU_E_FRONTIER;
  }
  while (b->chains.size() > 1) MZ_TRY(batcher_merge_top(b));
  Chain merged;
  if (!b->chains.empty()) {
    merged = std::move(b->chains.back());
    b->chains.pop_back();
  }
  Chain ship, keep;
  b->frontier = MZGPU_FRONTIER_EMPTY;
  if (merged.len == 0) {
    MZ_TRY(ship.rows.alloc(ctx, 16));
  } else if (upper == MZGPU_FRONTIER_EMPTY || (merged.time_known && merged.max_time < upper)) {
    ship = std::move(merged);  // empty antichain, or every buffered time is < upper: everything ships
  } else {
    u64 min_keep = MZGPU_FRONTIER_EMPTY;
    MZ_TRY(mz_extract(ctx, b->rb, merged.rows.p, merged.len, upper, &ship.rows, &ship.len, &keep.rows,
                      &keep.len, &min_keep));
    b->frontier = keep.len ? min_keep : MZGPU_FRONTIER_EMPTY;
  }
  if (keep.len) b->chains.push_back(std::move(keep));
  mzgpu_desc d = {b->lower, upper, 0};
  if (ship.has_index)
    MZ_TRY(make_batch_indexed(ctx, b->rb, std::move(ship.rows), ship.len, std::move(ship.table), ship.slots,
                              ship.n_keys, d, batch_out));
  else
    MZ_TRY(make_batch(ctx, b->rb, std::move(ship.rows), ship.len, d, batch_out));
  b->lower = upper;
  if (new_lower) *new_lower = b->frontier;
  return MZGPU_OK;
}

extern "C" int32_t mzgpu_batcher_new(mzgpu_ctx* ctx, uint32_t row_bytes, mzgpu_batcher** out) {
  MZ_CHECK_CTX(ctx);
  if (out == nullptr || (row_bytes != 32 && row_bytes != 80)) return MZGPU_E_INVALID;
  mzgpu_batcher* b = new mzgpu_batcher();
  b->ctx = ctx;
  b->rb = row_bytes;
  *out = b;
  return MZGPU_OK;
}
extern "C" void mzgpu_batcher_free(mzgpu_batcher* b) { delete b; }
extern "C" int32_t mzgpu_batcher_push(mzgpu_batcher* b, const void* rows, uint64_t n, int32_t mem) {
  if (b == nullptr || (rows == nullptr && n)) return MZGPU_E_INVALID;
  MZ_CHECK_CTX(b->ctx);
  if (n == 0) return MZGPU_OK;
  b->ctx->stats.rows_in += n;
  DevMem in;
  const void* d_in = rows;
  if (mem == MZGPU_MEM_HOST) {
    MZ_TRY(in.alloc(b->ctx, n * b->rb));
    MZ_TRY(copy_in(b->ctx, in.p, rows, n * b->rb, mem));
    d_in = in.p;
  }
  return batcher_push_dev(b, d_in, n);
}
extern "C" int32_t mzgpu_batcher_seal(mzgpu_batcher* b, uint64_t upper, mzgpu_batch** batch_out,
                                      uint64_t* new_lower) {
  if (b == nullptr || batch_out == nullptr) return MZGPU_E_INVALID;
  MZ_CHECK_CTX(b->ctx);
  return batcher_seal(b, upper, batch_out, new_lower);
}
extern "C" uint64_t mzgpu_batcher_frontier(const mzgpu_batcher* b) {
  return b ? b->frontier : MZGPU_FRONTIER_EMPTY;
}
extern "C" uint64_t mzgpu_batcher_len(const mzgpu_batcher* b) {
  u64 n = 0;
  if (b)
    for (auto& c : b->chains) n += c.len;
  return n;
}

// ==================================================================== spine
// Host-side restatement of spine_fueled::Spine's scheduling (in-tree fork
// src/persist-client/src/internal/trace.rs:1565-2262) over
// device batches.  Fuel is bookkeeping on the host; a merge runs as one
// merge-path kernel sequence when the schedule completes it.
struct mzgpu_spine {
  struct Layer {
    std::vector<mzgpu_batch*> batches;  // at most 2 (owned references)
    bool has_merge = false;
    u64 merge_since = 0;
    u64 remaining = 0;
  };
  mzgpu_ctx* ctx;
  uint32_t rb;
  u64 effort = 1;
  u64 since = 0;
  u64 physical = 0;
  u64 upper = 0;
  std::vector<Layer> merging;
  std::vector<mzgpu_batch*> pending;  // not yet admitted (upper > physical compaction)
  std::vector<mzgpu_batch*> view;     // scratch for batches_through
  int32_t err = MZGPU_OK;             // first failure inside a scheduling step

  static u64 level_of(u64 n) {
    u64 p = 1, l = 0;
    while (p < n) {
      p <<= 1;
      ++l;
    }
    return l;
  }
  u64 layer_len(const Layer& m) const {
    u64 n = 0;
    for (auto* b : m.batches) n += b->len;
    return n;
  }
  bool reduced() const {
    int non_empty = 0;
    for (auto& m : merging) {
      if (m.batches.size() == 2) return false;
      if (layer_len(m) > 0) ++non_empty;
      if (non_empty > 1) return false;
    }
    return true;
  }
  void begin_merge(Layer& m, bool with_frontier) {
    u64 s = 0, work = 0;
    for (auto* b : m.batches) {
      s = std::max(s, b->desc.since);
      work += b->len;
    }
    if (with_frontier) s = std::max(s, since);
    m.has_merge = true;
    m.merge_since = s;
    m.remaining = work;
  }
  void insert_at(mzgpu_batch* b, size_t index) {
    while (merging.size() <= index) merging.push_back(Layer());
    Layer& m = merging[index];
    if (m.has_merge || m.batches.size() >= 2) {
      MZ_SET_ERR(ctx, "spine: attempted to insert batch into a full / merging layer %zu", index);
      err = MZGPU_E_INVALID;
      mzgpu_batch_release(b);
      return;
    }
    m.batches.push_back(b);
    if (m.batches.size() == 2) begin_merge(m, true);
  }
  // MergeState::complete: returns an owned batch or nullptr
  mzgpu_batch* complete_at(size_t index) {
    Layer m = std::move(merging[index]);
    merging[index] = Layer();
    if (m.batches.empty()) return nullptr;
    if (m.batches.size() == 1) return m.batches[0];
    if (!m.has_merge) begin_merge(m, false);
    mzgpu_batch *b1 = m.batches[0], *b2 = m.batches[1];
    mzgpu_batch* out = nullptr;
    int32_t st;
    if (b1->len == 0 && b2->len == 0) {
      mzgpu_desc d = {b1->desc.lower, b2->desc.upper, m.merge_since};
      st = make_empty_batch(ctx, rb, d, &out);
    } else {
      st = merge_batches(b1, b2, m.merge_since, &out);
    }
    if (st != MZGPU_OK && err == MZGPU_OK) err = st;
    mzgpu_batch_release(b1);
    mzgpu_batch_release(b2);
    return out;
  }
  void apply_fuel(long long fuel_in) {
    for (size_t index = 0; index < merging.size(); ++index) {
      Layer& m = merging[index];
      if (m.has_merge) {
        u64 f = fuel_in < 0 ? 0 : (u64)fuel_in;
        m.remaining -= std::min(f, m.remaining);
      }
      if (m.has_merge && m.remaining == 0) {
        mzgpu_batch* done = complete_at(index);
        if (done) insert_at(done, index + 1);
      }
    }
  }
  void roll_up(size_t index) {
    while (merging.size() <= index) merging.push_back(Layer());
    bool any = false;
    for (size_t i = 0; i < index; ++i) any = any || !merging[i].batches.empty();
    if (!any) return;
    mzgpu_batch* merged = nullptr;
    for (size_t i = 0; i < index; ++i) {
      if (merged) {
        insert_at(merged, i);
        merged = nullptr;
      }
      merged = complete_at(i);
    }
    if (merged) insert_at(merged, index);
    if (merging[index].batches.size() == 2) {
      mzgpu_batch* m2 = complete_at(index);
      if (m2) insert_at(m2, index + 1);
    }
  }
  void tidy_layers() {
    if (merging.empty()) return;
    size_t length = merging.size();
    if (merging[length - 1].batches.size() != 1) return;
    u64 appropriate = level_of(layer_len(merging[length - 1]));
    while (appropriate < length - 1) {
      Layer& cur = merging[length - 2];
      if (cur.batches.empty()) {
        merging.erase(merging.begin() + (length - 2));
        length = merging.size();
      } else {
        if (cur.batches.size() != 2) {
          u64 smaller = 0;
          for (size_t i = 0; i < length - 2; ++i) smaller += (u64)merging[i].batches.size() << i;
          if (smaller <= ((u64)1 << length) / 8) {
            Layer state = std::move(merging[length - 2]);
            merging.erase(merging.begin() + (length - 2));
            for (auto* b : state.batches) insert_at(b, length - 2);
          }
        }
        break;
      }
    }
  }
  void introduce_batch(mzgpu_batch* b, size_t index) {
    long long fuel = (long long)((8ull << index) * effort);
    apply_fuel(fuel);
    roll_up(index);
    insert_at(b, index);
    tidy_layers();
  }
  void insert_entry(mzgpu_batch* b) {
    if (b->len == 0) {
      for (size_t pos = 0; pos < merging.size(); ++pos) {
        if (merging[pos].batches.empty()) continue;
        if (merging[pos].batches.size() == 1 && layer_len(merging[pos]) == 0) {
          insert_at(b, pos);
          mzgpu_batch* merged = complete_at(pos);
          if (merged) {
            merging[pos] = Layer();
            merging[pos].batches.push_back(merged);
          }
          return;
        }
        break;
      }
    }
    introduce_batch(b, level_of(b->len));
  }
  void consider_merges() {
    while (!pending.empty()) {
      mzgpu_batch* b = pending.front();
      bool ok = physical == MZGPU_FRONTIER_EMPTY ||
                (b->desc.upper != MZGPU_FRONTIER_EMPTY && b->desc.upper <= physical);
      if (!ok) break;
      pending.erase(pending.begin());
      insert_entry(b);
    }
  }
  // oldest first
  void all_batches(std::vector<mzgpu_batch*>& out) const {
    for (size_t i = merging.size(); i-- > 0;)
      for (auto* b : merging[i].batches) out.push_back(b);
    for (auto* b : pending) out.push_back(b);
  }
  ~mzgpu_spine() {
    for (auto& m : merging)
      for (auto* b : m.batches) mzgpu_batch_release(b);
    for (auto* b : pending) mzgpu_batch_release(b);
  }
};

static int32_t spine_take_err(mzgpu_spine* s) {
  int32_t e = s->err;
  s->err = MZGPU_OK;
  return e;
}

extern "C" int32_t mzgpu_spine_new(mzgpu_ctx* ctx, uint32_t row_bytes, uint32_t effort,
                                   mzgpu_spine** out) {
  MZ_CHECK_CTX(ctx);
  if (out == nullptr || (row_bytes != 32 && row_bytes != 80)) return MZGPU_E_INVALID;
  mzgpu_spine* s = new mzgpu_spine();
  s->ctx = ctx;
  s->rb = row_bytes;
  s->effort = effort ? effort : 1;
  *out = s;
  return MZGPU_OK;
}
extern "C" void mzgpu_spine_free(mzgpu_spine* s) { delete s; }
extern "C" int32_t mzgpu_spine_insert(mzgpu_spine* s, mzgpu_batch* batch) {
  if (s == nullptr || batch == nullptr || batch->rb != s->rb) return MZGPU_E_INVALID;
  MZ_CHECK_CTX(s->ctx);
  if (batch->desc.lower == batch->desc.upper || batch->desc.lower != s->upper) {
    MZ_SET_ERR(s->ctx, "spine_insert: batch [%llu, %llu) does not extend trace upper %llu",
               (unsigned long long)batch->desc.lower, (unsigned long long)batch->desc.upper,
               (unsigned long long)s->upper);
    return MZGPU_E_FRONTIER;
  }
  mzgpu_batch_retain(batch);
  s->upper = batch->desc.upper;
  s->pending.push_back(batch);
  s->consider_merges();
  return spine_take_err(s);
}
extern "C" int32_t mzgpu_spine_exert(mzgpu_spine* s, uint64_t effort, int32_t* did_work) {
  if (s == nullptr) return MZGPU_E_INVALID;
  MZ_CHECK_CTX(s->ctx);
  if (did_work) *did_work = 0;
  s->tidy_layers();
  if (s->reduced()) return spine_take_err(s);
  bool any = false;
  for (auto& m : s->merging) any = any || m.has_merge;
  if (any) {
    s->apply_fuel((long long)effort);
  } else {
    mzgpu_batch* e = nullptr;
    mzgpu_desc d = {s->upper, s->upper, s->since};
    MZ_TRY(make_empty_batch(s->ctx, s->rb, d, &e));
    s->introduce_batch(e, mzgpu_spine::level_of(effort));
  }
  if (did_work) *did_work = 1;
  return spine_take_err(s);
}
extern "C" uint64_t mzgpu_spine_exert_logic(const mzgpu_spine* s, uint32_t proportionality) {
  if (s == nullptr || proportionality == 0) return 0;
  uint32_t prop = proportionality;
  bool skipping = true, first = true;
  for (size_t i = s->merging.size(); i-- > 0;) {
    size_t count = s->merging[i].batches.size();
    u64 len = s->layer_len(s->merging[i]);
    if (skipping && count == 0) continue;
    skipping = false;
    if (count > 1) return 1000;
    if (!first && prop > 0 && len > 0) return 1000;
    first = false;
    prop /= 2;
  }
  return 0;
}
extern "C" int32_t mzgpu_spine_set_logical_compaction(mzgpu_spine* s, uint64_t f) {
  if (s == nullptr) return MZGPU_E_INVALID;
  if (f == MZGPU_FRONTIER_EMPTY || f > s->since) s->since = f;
  return MZGPU_OK;
}
extern "C" int32_t mzgpu_spine_set_physical_compaction(mzgpu_spine* s, uint64_t f) {
  if (s == nullptr) return MZGPU_E_INVALID;
  MZ_CHECK_CTX(s->ctx);
  if (f == MZGPU_FRONTIER_EMPTY || (s->physical != MZGPU_FRONTIER_EMPTY && f > s->physical)) s->physical = f;
  s->consider_merges();
  return spine_take_err(s);
}
extern "C" uint64_t mzgpu_spine_get_logical_compaction(const mzgpu_spine* s) { return s ? s->since : 0; }
extern "C" uint64_t mzgpu_spine_get_physical_compaction(const mzgpu_spine* s) {
  return s ? s->physical : 0;
}
extern "C" uint64_t mzgpu_spine_read_upper(const mzgpu_spine* s) { return s ? s->upper : 0; }

static void spine_through(mzgpu_spine* s, u64 through, std::vector<mzgpu_batch*>& out) {
  std::vector<mzgpu_batch*> all;
  s->all_batches(all);
  for (auto* b : all) {
    if (through == MZGPU_FRONTIER_EMPTY || (b->desc.upper != MZGPU_FRONTIER_EMPTY && b->desc.upper <= through))
      out.push_back(b);
  }
}
extern "C" int32_t mzgpu_spine_batches_through(mzgpu_spine* s, uint64_t upper, mzgpu_batch** batches,
                                               uint32_t cap, uint32_t* n_out) {
  if (s == nullptr || n_out == nullptr) return MZGPU_E_INVALID;
  s->view.clear();
  spine_through(s, upper, s->view);
  *n_out = (uint32_t)s->view.size();
  if (s->view.size() > cap) return MZGPU_E_CAPACITY;
  for (size_t i = 0; i < s->view.size(); ++i) batches[i] = s->view[i];
  return MZGPU_OK;
}
extern "C" int32_t mzgpu_spine_layers(const mzgpu_spine* s, uint64_t* out4, uint32_t cap_layers,
                                      uint32_t* n_layers) {
  if (s == nullptr || n_layers == nullptr) return MZGPU_E_INVALID;
  uint32_t n = 0;
  for (size_t i = s->merging.size(); i-- > 0;) {
    if (n >= cap_layers) return MZGPU_E_CAPACITY;
    const auto& m = s->merging[i];
    out4[4 * n + 0] = m.batches.size();
    out4[4 * n + 1] = m.batches.size() > 0 ? m.batches[0]->len : 0;
    out4[4 * n + 2] = m.batches.size() > 1 ? m.batches[1]->len : 0;
    out4[4 * n + 3] = m.has_merge ? m.remaining : 0;
    ++n;
  }
  *n_layers = n;
  return MZGPU_OK;
}

static int32_t trace_view(mzgpu_ctx* ctx, const std::vector<mzgpu_batch*>& batches, TraceView* tv) {
  tv->n_batches = 0;
  for (auto* b : batches) {
    if (b->len == 0) continue;
    if (tv->n_batches >= MZ_MAX_TRACE_BATCHES) {
      MZ_SET_ERR(ctx, "trace has more than %d non-empty batches", MZ_MAX_TRACE_BATCHES);
      return MZGPU_E_UNSUPPORTED;
    }
    BatchView& v = tv->b[tv->n_batches++];
    v.rows = b->rows.as<u64>();
    v.table = b->table.as<HashSlot>();
    v.n = b->len;
    v.mask = b->slots - 1;
  }
  return MZGPU_OK;
}

extern "C" int32_t mzgpu_spine_export(mzgpu_spine* s, mzgpu_buf* out) {
  if (s == nullptr || out == nullptr || out->rb != s->rb) return MZGPU_E_INVALID;
  MZ_CHECK_CTX(s->ctx);
  std::vector<mzgpu_batch*> all;
  s->all_batches(all);
  // fold the batches oldest-first with the merge kernel (advancing to `since`)
  DevMem acc;
  u64 acc_len = 0;
  MZ_TRY(acc.alloc(s->ctx, 16));
  for (auto* b : all) {
    DevMem m;
    u64 n = 0;
    MZ_TRY(mz_merge_consolidate(s->ctx, s->rb, acc.p, acc_len, b->rows.p, b->len, s->since, &m, &n));
    acc = std::move(m);
    acc_len = n;
  }
  buf_adopt(out, std::move(acc), acc_len);
  return MZGPU_OK;
}

// ================================================================ join_core
struct mzgpu_join {
  mzgpu_ctx* ctx;
  mzgpu_spine *t1, *t2;
  bool has_closure;
  mzgpu_closure closure;
  u64 ack1 = 0, ack2 = 0;
  struct Work {
    int side;
    mzgpu_batch* batch;
    std::vector<mzgpu_batch*> others;
    u64 cap;
  };
  std::deque<Work> todo;
  void release_work(Work& w) {
    mzgpu_batch_release(w.batch);
    for (auto* b : w.others) mzgpu_batch_release(b);
  }
  ~mzgpu_join() {
    for (auto& w : todo) release_work(w);
  }
};

static void join_enqueue(mzgpu_join* j, int side, mzgpu_batch* batch, u64 cap) {
  mzgpu_join::Work w;
  w.side = side;
  w.batch = batch;
  mzgpu_batch_retain(batch);
  spine_through(side == 0 ? j->t2 : j->t1, side == 0 ? j->ack2 : j->ack1, w.others);
  for (auto* b : w.others) mzgpu_batch_retain(b);
  w.cap = cap;
  j->todo.push_back(std::move(w));
}

extern "C" int32_t mzgpu_join_new(mzgpu_ctx* ctx, mzgpu_spine* trace1, mzgpu_spine* trace2,
                                  const mzgpu_closure* closure, mzgpu_join** out) {
  MZ_CHECK_CTX(ctx);
  if (trace1 == nullptr || trace2 == nullptr || out == nullptr || trace1->rb != 32 || trace2->rb != 32)
    return MZGPU_E_INVALID;
  mzgpu_join* j = new mzgpu_join();
  j->ctx = ctx;
  j->t1 = trace1;
  j->t2 = trace2;
  j->has_closure = closure != nullptr;
  memset(&j->closure, 0, sizeof(j->closure));
  if (closure) j->closure = *closure;
  // pre-load (mz_join_core.rs:109-190): trace1's batches are acknowledged, then
  // each existing trace2 batch is joined against trace1 through ack1
  std::vector<mzgpu_batch*> all;
  trace1->all_batches(all);
  for (auto* b : all) j->ack1 = b->desc.upper;
  all.clear();
  trace2->all_batches(all);
  for (auto* b : all) {
    if (b->len) join_enqueue(j, 1, b, 0);
    j->ack2 = b->desc.upper;
  }
  *out = j;
  return MZGPU_OK;
}
extern "C" void mzgpu_join_free(mzgpu_join* j) { delete j; }

extern "C" int32_t mzgpu_join_core_push(mzgpu_join* j, int32_t side, mzgpu_batch* batch, uint64_t cap) {
  if (j == nullptr || batch == nullptr || (side != 0 && side != 1) || batch->rb != 32) return MZGPU_E_INVALID;
  MZ_CHECK_CTX(j->ctx);
  u64& ack = side == 0 ? j->ack1 : j->ack2;
  if (ack <= batch->desc.lower) {
    if (batch->len) join_enqueue(j, side, batch, cap);
    ack = batch->desc.upper;
  }
  // physical compaction of both traces follows the acknowledged frontiers
  MZ_TRY(mzgpu_spine_set_physical_compaction(j->t1, j->ack1));
  MZ_TRY(mzgpu_spine_set_physical_compaction(j->t2, j->ack2));
  return MZGPU_OK;
}

extern "C" int32_t mzgpu_join_core_work(mzgpu_join* j, uint64_t fuel_rows, mzgpu_buf* out,
                                        int32_t* done) {
  if (j == nullptr || out == nullptr) return MZGPU_E_INVALID;
  MZ_CHECK_CTX(j->ctx);
  const uint32_t out_rb = j->has_closure ? 32 : 40;
  if (out->rb != out_rb) {
    MZ_SET_ERR(j->ctx, "join_core_work: output buffer row width %u, expected %u", out->rb, out_rb);
    return MZGPU_E_INVALID;
  }
  u64 produced = 0;
  while (!j->todo.empty() && produced < fuel_rows) {
    mzgpu_join::Work w = std::move(j->todo.front());
    j->todo.pop_front();
    TraceView tv;
    int32_t st = trace_view(j->ctx, w.others, &tv);
    DevMem res, cons;
    u64 n_res = 0, n_cons = 0;
    if (st == MZGPU_OK) {
      ProbeParams pp;
      memset(&pp, 0, sizeof(pp));
      pp.mode = MZ_PROBE_JOIN;
      pp.meet = w.cap;
      pp.has_closure = j->has_closure ? 1 : 0;
      pp.swap_vals = w.side == 1 ? 1 : 0;
      pp.closure = j->closure;
      st = mz_probe(j->ctx, w.batch->rows.as<u64>(), w.batch->len, tv, pp, &res, &n_res);
    }
    // Work::process consolidates each work item's output buffer before sending
    if (st == MZGPU_OK) st = mz_sort_consolidate(j->ctx, out_rb, res.p, n_res, &cons, &n_cons);
    if (st == MZGPU_OK) st = buf_append_dev(out, cons.p, n_cons);
    j->release_work(w);
    if (st != MZGPU_OK) return st;
    produced += n_cons;
    j->ctx->stats.rows_out += n_cons;
  }
  if (done) *done = j->todo.empty() ? 1 : 0;
  return MZGPU_OK;
}

// ================================================================ half_join
extern "C" int32_t mzgpu_half_join(mzgpu_ctx* ctx, const mzgpu_r32* stream, uint64_t n, int32_t mem,
                                   mzgpu_spine* trace, int32_t cmp_mode, const mzgpu_closure* closure,
                                   int32_t consolidate_output, mzgpu_buf* out) {
  MZ_CHECK_CTX(ctx);
  if (trace == nullptr || out == nullptr || (stream == nullptr && n) || trace->rb != 32 || out->rb != 32 ||
      (cmp_mode != MZGPU_HALFJOIN_LE && cmp_mode != MZGPU_HALFJOIN_LT))
    return MZGPU_E_INVALID;
  if (n == 0) return MZGPU_OK;
  ctx->stats.rows_in += n;
  DevMem in;
  const u64* d_stream = (const u64*)stream;
  if (mem == MZGPU_MEM_HOST) {
    MZ_TRY(in.alloc(ctx, n * 32));
    MZ_TRY(copy_in(ctx, in.p, stream, n * 32, mem));
    d_stream = in.as<u64>();
  }
  std::vector<mzgpu_batch*> all;
  trace->all_batches(all);
  TraceView tv;
  MZ_TRY(trace_view(ctx, all, &tv));
  ProbeParams pp;
  memset(&pp, 0, sizeof(pp));
  pp.mode = cmp_mode == MZGPU_HALFJOIN_LE ? MZ_PROBE_HALF_LE : MZ_PROBE_HALF_LT;
  pp.has_closure = 1;
  if (closure) {
    pp.closure = *closure;
  } else {
    // identity on (key, val2): the lookup value replaces the stream value
    pp.closure.n_key_fields = 1;
    pp.closure.key_fields[0] = mzgpu_field{MZGPU_SRC_KEY, 0, 64, 0};
    pp.closure.n_val_fields = 1;
    pp.closure.val_fields[0] = mzgpu_field{MZGPU_SRC_VAL2, 0, 64, 0};
  }
  DevMem res;
  u64 n_res = 0;
  MZ_TRY(mz_probe(ctx, d_stream, n, tv, pp, &res, &n_res));
  if (consolidate_output && n_res) {
    DevMem cons;
    u64 n_cons = 0;
    MZ_TRY(mz_sort_consolidate(ctx, 32, res.p, n_res, &cons, &n_cons));
    MZ_TRY(buf_append_dev(out, cons.p, n_cons));
    ctx->stats.rows_out += n_cons;
  } else {
    MZ_TRY(buf_append_dev(out, res.p, n_res));
    ctx->stats.rows_out += n_res;
  }
  return MZGPU_OK;
}

extern "C" int32_t mzgpu_update_stream(mzgpu_ctx* ctx, mzgpu_batch* batch,
                                       const mzgpu_closure* initial_closure, uint64_t skip_time,
                                       mzgpu_buf* out) {
  MZ_CHECK_CTX(ctx);
  if (batch == nullptr || out == nullptr || batch->rb != 32 || out->rb != 32) return MZGPU_E_INVALID;
  if (batch->len == 0) return MZGPU_OK;
  DevMem res;
  u64 n_res = 0;
  MZ_TRY(mz_map_rows_dev(ctx, batch->rows.as<u64>(), batch->len, initial_closure, skip_time, &res, &n_res));
  return buf_append_dev(out, res.p, n_res);
}

extern "C" int32_t mzgpu_map_rows(mzgpu_ctx* ctx, const mzgpu_r32* rows, uint64_t n, int32_t mem,
                                  const mzgpu_closure* closure, mzgpu_buf* out) {
  MZ_CHECK_CTX(ctx);
  if (out == nullptr || (rows == nullptr && n) || out->rb != 32) return MZGPU_E_INVALID;
  if (n == 0) return MZGPU_OK;
  DevMem in;
  const u64* d_rows = (const u64*)rows;
  if (mem == MZGPU_MEM_HOST) {
    MZ_TRY(in.alloc(ctx, n * 32));
    MZ_TRY(copy_in(ctx, in.p, rows, n * 32, mem));
    d_rows = in.as<u64>();
  }
  DevMem res;
  u64 n_res = 0;
  MZ_TRY(mz_map_rows_dev(ctx, d_rows, n, closure, MZGPU_FRONTIER_EMPTY, &res, &n_res));
  return buf_append_dev(out, res.p, n_res);
}

// =================================================================== reduce
struct mzgpu_reduce {
  mzgpu_ctx* ctx;
  int agg_kind;
  mzgpu_batcher* batcher = nullptr;
  mzgpu_spine* input = nullptr;
  ~mzgpu_reduce() {
    delete batcher;
    delete input;
  }
};

extern "C" int32_t mzgpu_reduce_new(mzgpu_ctx* ctx, int32_t agg_kind, mzgpu_reduce** out) {
  MZ_CHECK_CTX(ctx);
  if (out == nullptr || (agg_kind != MZGPU_AGG_COUNT_SUM_I64 && agg_kind != MZGPU_AGG_COUNT_SUM_F64))
    return MZGPU_E_INVALID;
  std::unique_ptr<mzgpu_reduce> r(new mzgpu_reduce());
  r->ctx = ctx;
  r->agg_kind = agg_kind;
  MZ_TRY(mzgpu_batcher_new(ctx, 80, &r->batcher));
  MZ_TRY(mzgpu_spine_new(ctx, 80, 1, &r->input));
  *out = r.release();
  return MZGPU_OK;
}
extern "C" void mzgpu_reduce_free(mzgpu_reduce* r) { delete r; }
extern "C" mzgpu_spine* mzgpu_reduce_input_trace(mzgpu_reduce* r) { return r ? r->input : nullptr; }

extern "C" int32_t mzgpu_reduce_accumulable(mzgpu_reduce* r, const mzgpu_r32* rows, uint64_t n,
                                            int32_t mem, uint64_t upper, mzgpu_buf* out) {
  if (r == nullptr || out == nullptr || (rows == nullptr && n) || out->rb != 64) return MZGPU_E_INVALID;
  mzgpu_ctx* ctx = r->ctx;
  MZ_CHECK_CTX(ctx);
  ctx->stats.rows_in += n;
  // explode_one: values move into the diff
  if (n) {
    DevMem in, exploded;
    const u64* d_rows = (const u64*)rows;
    if (mem == MZGPU_MEM_HOST) {
      MZ_TRY(in.alloc(ctx, n * 32));
      MZ_TRY(copy_in(ctx, in.p, rows, n * 32, mem));
      d_rows = in.as<u64>();
    }
    MZ_TRY(exploded.alloc(ctx, n * 80));
    MZ_TRY(mz_explode(ctx, d_rows, n, r->agg_kind, exploded.as<u64>()));
    MZ_TRY(batcher_push_dev(r->batcher, exploded.p, n));
  }
  // arrange: seal the accumulable arrangement's batch at the new frontier
  mzgpu_batch* batch = nullptr;
  MZ_TRY(batcher_seal(r->batcher, upper, &batch, nullptr));
  // reduce_abelian over the keys of the new batch
  std::vector<mzgpu_batch*> prior;
  r->input->all_batches(prior);
  TraceView tv;
  int32_t st = trace_view(ctx, prior, &tv);
  DevMem corr, cons;
  u64 n_corr = 0, n_cons = 0;
  if (st == MZGPU_OK)
    st = mz_reduce_corrections(ctx, batch->rows.as<u64>(), batch->len, tv, r->agg_kind, &corr, &n_corr);
  if (st == MZGPU_OK && n_corr) st = mz_sort_consolidate(ctx, 64, corr.p, n_corr, &cons, &n_cons);
  if (st == MZGPU_OK && n_cons) st = buf_append_dev(out, cons.p, n_cons);
  if (st == MZGPU_OK && batch->desc.lower != batch->desc.upper) st = mzgpu_spine_insert(r->input, batch);
  if (st == MZGPU_OK) st = mzgpu_spine_set_physical_compaction(r->input, r->input->upper);
  mzgpu_batch_release(batch);
  ctx->stats.rows_out += n_cons;
  return st;
}

// ================================================================= exchange
struct NcclId {
  char internal[128];
};
static void* open_nccl() {
  const char* names[] = {getenv("MZGPU_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    if (nm == nullptr) continue;
    void* h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (h) return h;
  }
  return nullptr;
}
extern "C" int32_t mzgpu_comm_unique_id(uint8_t id[MZGPU_COMM_ID_BYTES]) {
  void* lib = open_nccl();
  if (lib == nullptr) return MZGPU_E_NCCL;
  typedef int (*fn_t)(NcclId*);
  fn_t f = (fn_t)dlsym(lib, "ncclGetUniqueId");
  if (f == nullptr) return MZGPU_E_NCCL;
  NcclId uid;
  if (f(&uid) != 0) return MZGPU_E_NCCL;
  memcpy(id, uid.internal, MZGPU_COMM_ID_BYTES);
  return MZGPU_OK;
}
extern "C" int32_t mzgpu_comm_init(mzgpu_ctx* ctx, const uint8_t id[MZGPU_COMM_ID_BYTES]) {
  MZ_CHECK_CTX(ctx);
  if (ctx->peers == 1) return MZGPU_OK;
  ctx->nccl_lib = open_nccl();
  if (ctx->nccl_lib == nullptr) {
    MZ_SET_ERR(ctx, "comm_init: cannot dlopen libnccl.so.2 (set MZGPU_NCCL_LIB)");
    return MZGPU_E_NCCL;
  }
  typedef int (*fn_t)(void**, int, NcclId, int);
  fn_t f = (fn_t)dlsym(ctx->nccl_lib, "ncclCommInitRank");
  if (f == nullptr) return MZGPU_E_NCCL;
  NcclId uid;
  memcpy(uid.internal, id, MZGPU_COMM_ID_BYTES);
  MZ_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = f(&ctx->nccl_comm, ctx->peers, uid, ctx->worker);
  if (rc != 0) {
    MZ_SET_ERR(ctx, "ncclCommInitRank failed with %d", rc);
    ctx->sticky = true;
    return MZGPU_E_NCCL;
  }
  return MZGPU_OK;
}

extern "C" int32_t mzgpu_exchange(mzgpu_ctx* ctx, mzgpu_buf* in, mzgpu_buf* out) {
  MZ_CHECK_CTX(ctx);
  if (in == nullptr || out == nullptr || in->rb != out->rb || in == out) return MZGPU_E_INVALID;
  const u32 P = (u32)ctx->peers;
  if (P == 1) {
    out->len = 0;
    return buf_append_dev(out, in->mem.p, in->len);
  }
  if (ctx->nccl_comm == nullptr) {
    MZ_SET_ERR(ctx, "exchange: mzgpu_comm_init has not been called");
    return MZGPU_E_NCCL;
  }
  typedef int (*grp_t)();
  typedef int (*sr_t)(const void*, size_t, int, int, void*, cudaStream_t);
  typedef int (*rv_t)(void*, size_t, int, int, void*, cudaStream_t);
  grp_t gstart = (grp_t)dlsym(ctx->nccl_lib, "ncclGroupStart");
  grp_t gend = (grp_t)dlsym(ctx->nccl_lib, "ncclGroupEnd");
  sr_t send = (sr_t)dlsym(ctx->nccl_lib, "ncclSend");
  rv_t recv = (rv_t)dlsym(ctx->nccl_lib, "ncclRecv");
  if (!gstart || !gend || !send || !recv) return MZGPU_E_NCCL;
  const int NCCL_INT8 = 0, NCCL_UINT64 = 5;
  // 1. bucket rows by destination
  DevMem parts;
  MZ_TRY(parts.alloc(ctx, in->len * in->rb));
  u64 send_counts[64], recv_counts[64];
  MZ_TRY(mz_partition(ctx, in->rb, in->mem.p, in->len, P, parts.p, send_counts));
  // 2. counts all-to-all (P x u64)
  DevMem d_send, d_recv;
  MZ_TRY(d_send.alloc(ctx, P * 8));
  MZ_TRY(d_recv.alloc(ctx, P * 8));
  MZ_CUDA(ctx, cudaMemcpyAsync(d_send.p, send_counts, P * 8, cudaMemcpyHostToDevice, ctx->stream));
#define NCCL_TRY(expr)                                        \
  do {                                                        \
    int _rc = (expr);                                         \
    if (_rc != 0) {                                           \
      MZ_SET_ERR(ctx, "NCCL call failed with %d at %s:%d", _rc, __FILE__, __LINE__); \
      ctx->sticky = true;                                     \
      return MZGPU_E_NCCL;                                    \
    }                                                         \
  } while (0)
  NCCL_TRY(gstart());
  for (u32 p = 0; p < P; ++p) {
    NCCL_TRY(send(d_send.as<u64>() + p, 1, NCCL_UINT64, (int)p, ctx->nccl_comm, ctx->stream));
    NCCL_TRY(recv(d_recv.as<u64>() + p, 1, NCCL_UINT64, (int)p, ctx->nccl_comm, ctx->stream));
  }
  NCCL_TRY(gend());
  MZ_CUDA(ctx, cudaMemcpyAsync(recv_counts, d_recv.p, P * 8, cudaMemcpyDeviceToHost, ctx->stream));
  MZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  // 3. payload all-to-all
  u64 total = 0;
  for (u32 p = 0; p < P; ++p) total += recv_counts[p];
  out->len = 0;
  MZ_TRY(buf_reserve(out, total, false));
  NCCL_TRY(gstart());
  u64 soff = 0, roff = 0;
  for (u32 p = 0; p < P; ++p) {
    if (send_counts[p])
      NCCL_TRY(send((const char*)parts.p + soff * in->rb, send_counts[p] * in->rb, NCCL_INT8, (int)p,
                    ctx->nccl_comm, ctx->stream));
    if (recv_counts[p])
      NCCL_TRY(recv((char*)out->mem.p + roff * in->rb, recv_counts[p] * in->rb, NCCL_INT8, (int)p,
                    ctx->nccl_comm, ctx->stream));
    soff += send_counts[p];
    roff += recv_counts[p];
  }
  NCCL_TRY(gend());
  out->len = total;
  // `parts` is freed stream-ordered after the sends
  return MZGPU_OK;
#undef NCCL_TRY
}
