// common.cuh — context, device memory, row traits and block-level primitives
// shared by every kernel file of libmzgpu (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/mzgpu.h"

typedef uint64_t u64;
typedef int64_t i64;
typedef uint32_t u32;

// ------------------------------------------------------------------ errors
struct mzgpu_ctx {
  int device = 0;
  int worker = 0;
  int peers = 1;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev = nullptr;
  int num_sms = 148;
  bool sticky = false;  // a CUDA/NCCL failure happened: every later call fails
  std::string last_error;
  mzgpu_stats stats;
  // pinned staging for small device->host reads (counts, min/max)
  u64* h_scratch = nullptr;  // 64 words
  u64* d_scratch = nullptr;  // 64 words
  // pinned bounce buffers for host<->device row copies
  void* h_bounce = nullptr;
  size_t h_bounce_bytes = 0;
  void* h_fused = nullptr;  // pinned image of the fused kernel's control block (16 KiB)
  // per-kernel profiling (mzgpu_profile_enable)
  struct ProfRec {
    const char* name;
    cudaEvent_t e0, e1;
    u64 bytes;
  };
  bool profile = false;
  u64 next_bytes = 0;  // algorithmic bytes of the next launch (MZ_BYTES)
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> ev_pool;
  // NCCL (resolved with dlopen at mzgpu_comm_init)
  void* nccl_lib = nullptr;
  void* nccl_comm = nullptr;
};

#define MZ_SET_ERR(ctx, ...)                              \
  do {                                                    \
    char _buf[512];                                       \
    snprintf(_buf, sizeof(_buf), __VA_ARGS__);            \
    (ctx)->last_error = _buf;                             \
  } while (0)

#define MZ_CUDA(ctx, expr)                                                               \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      MZ_SET_ERR(ctx, "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__,      \
                 __LINE__, cudaGetErrorString(_e));                                      \
      (ctx)->sticky = true;                                                              \
      return MZGPU_E_CUDA;                                                               \
    }                                                                                    \
  } while (0)

#define MZ_TRY(expr)                 \
  do {                               \
    int32_t _s = (expr);             \
    if (_s != MZGPU_OK) return _s;   \
  } while (0)

#define MZ_CHECK_CTX(ctx)                       \
  do {                                          \
    if ((ctx) == nullptr) return MZGPU_E_INVALID; \
    if ((ctx)->sticky) return MZGPU_E_CUDA;     \
  } while (0)

// Brackets one launch with CUDA events on the launching stream when profiling.
struct ProfScope {
  mzgpu_ctx* ctx;
  mzgpu_ctx::ProfRec rec;
  bool on;
  ProfScope(mzgpu_ctx* c, const char* name) : ctx(c), on(c->profile) {
    if (!on) return;
    rec.name = name;
    rec.bytes = c->next_bytes;
    for (cudaEvent_t* e : {&rec.e0, &rec.e1}) {
      if (!c->ev_pool.empty()) {
        *e = c->ev_pool.back();
        c->ev_pool.pop_back();
      } else {
        cudaEventCreate(e);
      }
    }
    cudaEventRecord(rec.e0, c->stream);
  }
  ~ProfScope() {
    ctx->next_bytes = 0;
    if (!on) return;
    cudaEventRecord(rec.e1, ctx->stream);
    ctx->prof.push_back(rec);
  }
};
// algorithmic bytes the next launch moves (documented per kernel in DESIGN.md)
#define MZ_BYTES(ctx, n) ((ctx)->next_bytes = (u64)(n))

// Kernel launch helper: counts launches (mzgpu_stats.kernel_launches) and
// surfaces launch-configuration errors immediately.
#define MZ_LAUNCH(ctx, kernel, grid, block, smem, ...)                              \
  do {                                                                              \
    {                                                                               \
      ProfScope _prof(ctx, #kernel);                                                \
      kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);              \
    }                                                                               \
    (ctx)->stats.kernel_launches++;                                                 \
    MZ_CUDA(ctx, cudaGetLastError());                                               \
  } while (0)

// ------------------------------------------------------------ device memory
// Stream-ordered allocations from the device's default mempool (kept warm by a
// release threshold of UINT64_MAX set at ctx creation), with byte accounting
// for the arrangement-size metrics.
struct DevMem {
  mzgpu_ctx* ctx = nullptr;
  void* p = nullptr;
  size_t bytes = 0;
  DevMem() {}
  DevMem(const DevMem&) = delete;
  DevMem& operator=(const DevMem&) = delete;
  DevMem(DevMem&& o) noexcept { *this = std::move(o); }
  DevMem& operator=(DevMem&& o) noexcept {
    if (this != &o) {
      release();
      ctx = o.ctx;
      p = o.p;
      bytes = o.bytes;
      o.p = nullptr;
      o.bytes = 0;
    }
    return *this;
  }
  ~DevMem() { release(); }
  int32_t alloc(mzgpu_ctx* c, size_t n) {
    release();
    ctx = c;
    if (n == 0) n = 16;
    cudaError_t e = cudaMallocAsync(&p, n, c->stream);
    if (e != cudaSuccess) {
      p = nullptr;
      MZ_SET_ERR(c, "cudaMallocAsync(%zu bytes) failed: %s", n, cudaGetErrorString(e));
      c->sticky = true;
      return MZGPU_E_CUDA;
    }
    bytes = n;
    c->stats.device_bytes_in_use += n;
    if (c->stats.device_bytes_in_use > c->stats.device_bytes_peak)
      c->stats.device_bytes_peak = c->stats.device_bytes_in_use;
    return MZGPU_OK;
  }
  void release() {
    if (p != nullptr) {
      cudaFreeAsync(p, ctx->stream);
      ctx->stats.device_bytes_in_use -= bytes;
      p = nullptr;
      bytes = 0;
    }
  }
  template <class T>
  T* as() const {
    return (T*)p;
  }
};

// ---------------------------------------------------------------- row traits
// A row is NW 64-bit words: NK sort-key words first (compared as unsigned, in
// order), then the diff words.  TW = index of the time word (or -1).
template <int RB>
struct RowT;
template <>
struct RowT<16> {  // mzgpu_r16 (key | diff)
  static constexpr int NW = 2, NK = 1, ND = 1, TW = -1, DK = 1;
};
template <>
struct RowT<32> {  // mzgpu_r32 (key, val, time | diff)
  static constexpr int NW = 4, NK = 3, ND = 1, TW = 2, DK = 2;
};
template <>
struct RowT<40> {  // mzgpu_r40 (key, val1, val2, time | diff)
  static constexpr int NW = 5, NK = 4, ND = 1, TW = 3, DK = 3;
};
template <>
struct RowT<80> {  // mzgpu_racc (key, time | total, non_nulls, acc_lo, acc_hi, pinf, ninf, nan, pad)
  static constexpr int NW = 10, NK = 2, ND = 8, TW = 1, DK = 1;
};
template <>
struct RowT<64> {  // mzgpu_rout (key, count, sum_lo, sum_hi, flags, time | diff, pad)
  static constexpr int NW = 8, NK = 6, ND = 2, TW = 5, DK = 5;
};
// DK = number of leading "data" words (key words before the time word): two
// rows with equal DK words are the same (key, val).

// Semigroup::plus_equals on the diff words.  ND == 8 is the accumulable diff:
// words 2,3 (acc_lo, acc_hi) form an i128 (src/compute/src/render/reduce.rs:1940-2041).
template <int ND>
__host__ __device__ __forceinline__ void diff_add(u64* a, const u64* b) {
  if (ND == 8) {
    a[0] += b[0];
    a[1] += b[1];
    u64 lo = a[2] + b[2];
    u64 carry = lo < a[2] ? 1 : 0;
    a[2] = lo;
    a[3] = a[3] + b[3] + carry;
    a[4] += b[4];
    a[5] += b[5];
    a[6] += b[6];
  } else {
    a[0] += b[0];  // ND == 2 is (diff, pad): pad stays 0
  }
}
template <int ND>
__host__ __device__ __forceinline__ bool diff_is_zero(const u64* a) {
  if (ND == 8) return (a[0] | a[1] | a[2] | a[3] | a[4] | a[5] | a[6]) == 0;
  return a[0] == 0;
}

// ----------------------------------------------------------- device helpers
#ifdef __CUDACC__
__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ u32 warp_id() { return threadIdx.x >> 5; }

// Block-wide exclusive scan of one u32 per thread (blockDim.x <= 1024, multiple
// of 32).  Returns the exclusive prefix; *total receives the block sum.
// `smem` must hold 33 u32.
__device__ __forceinline__ u32 block_exclusive_scan(u32 v, u32* smem, u32* total) {
  u32 lane = lane_id(), warp = warp_id();
  u32 incl = v;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    u32 o = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= (u32)off) incl += o;
  }
  if (lane == 31) smem[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    u32 nw = (blockDim.x + 31) >> 5;
    u32 w = lane < nw ? smem[lane] : 0;
    u32 wi = w;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      u32 o = __shfl_up_sync(0xffffffffu, wi, off);
      if (lane >= (u32)off) wi += o;
    }
    smem[lane] = wi - w;  // exclusive warp offsets
    if (lane == 31) smem[32] = wi;
  }
  __syncthreads();
  u32 res = smem[warp] + incl - v;
  *total = smem[32];
  __syncthreads();
  return res;
}

// Single-block exclusive scan of per-tile counts, in place; the grand total is
// written to *total_out.  n_tiles is small (n / 512), so one CTA suffices.
static __global__ void __launch_bounds__(1024) k_scan_tiles(u32* __restrict__ counts, u64 n_tiles,
                                                            u64* __restrict__ total_out) {
  __shared__ u32 sm[34];
  u32 carry = 0;
  for (u64 base = 0; base < n_tiles; base += 1024) {
    u64 i = base + threadIdx.x;
    u32 v = i < n_tiles ? counts[i] : 0;
    u32 total;
    u32 ex = block_exclusive_scan(v, sm, &total);
    if (i < n_tiles) counts[i] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) *total_out = carry;
}

// 128-bit vectorised global loads/stores of row words (rows are 16-byte aligned
// for every row width we use: 16, 32, 48, 64, 80).
template <int NW>
__device__ __forceinline__ void load_row(const u64* __restrict__ base, u64 idx, u64* r) {
  const u64* p = base + idx * NW;
  if (NW % 2 == 0) {
#pragma unroll
    for (int i = 0; i < NW; i += 2) {
      ulonglong2 v = *reinterpret_cast<const ulonglong2*>(p + i);
      r[i] = v.x;
      r[i + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < NW; ++i) r[i] = p[i];
  }
}
template <int NW>
__device__ __forceinline__ void store_row(u64* __restrict__ base, u64 idx, const u64* r) {
  u64* p = base + idx * NW;
  if (NW % 2 == 0) {
#pragma unroll
    for (int i = 0; i < NW; i += 2) {
      ulonglong2 v;
      v.x = r[i];
      v.y = r[i + 1];
      *reinterpret_cast<ulonglong2*>(p + i) = v;
    }
  } else {
#pragma unroll
    for (int i = 0; i < NW; ++i) p[i] = r[i];
  }
}

// 64-bit mixer for the open-addressing hash index (murmur3 finalizer).
__host__ __device__ __forceinline__ u64 mix64(u64 k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return k;
}

// Closure evaluation on device: same descriptor semantics as include/mzgpu.h.
__host__ __device__ __forceinline__ u64 field_get(const mzgpu_field& f, u64 key, u64 v1, u64 v2) {
  u64 w = f.src == MZGPU_SRC_KEY ? key : (f.src == MZGPU_SRC_VAL1 ? v1 : v2);
  w >>= f.shift;
  if (f.bits < 64) w &= ((u64)1 << f.bits) - 1;
  return w;
}
__host__ __device__ __forceinline__ bool closure_eval(const mzgpu_closure& c, u64 key, u64 v1, u64 v2,
                                                      u64* okey, u64* oval) {
  for (u32 i = 0; i < c.n_filters; ++i) {
    const mzgpu_filter& f = c.filters[i];
    u64 x = field_get(f.field, key, v1, v2);
    bool ok;
    switch (f.op) {
      case MZGPU_CMP_EQ: ok = x == f.rhs; break;
      case MZGPU_CMP_NE: ok = x != f.rhs; break;
      case MZGPU_CMP_LT: ok = x < f.rhs; break;
      case MZGPU_CMP_LE: ok = x <= f.rhs; break;
      case MZGPU_CMP_GT: ok = x > f.rhs; break;
      default: ok = x >= f.rhs; break;
    }
    if (!ok) return false;
  }
  u64 k = 0, v = 0;
  for (u32 i = 0; i < c.n_key_fields; ++i)
    k |= field_get(c.key_fields[i], key, v1, v2) << c.key_fields[i].dst_shift;
  if (c.expr_kind == MZGPU_EXPR_MUL_CONST_MINUS) {
    u64 a = field_get(c.expr_a, key, v1, v2);
    u64 b = field_get(c.expr_b, key, v1, v2);
    v = a * (c.expr_c - b);
  } else {
    for (u32 i = 0; i < c.n_val_fields; ++i)
      v |= field_get(c.val_fields[i], key, v1, v2) << c.val_fields[i].dst_shift;
  }
  *okey = k;
  *oval = v;
  return true;
}
#endif  // __CUDACC__

// ------------------------------------------------------- kernel entry points
// (implemented in the .cu files; all asynchronous on ctx->stream unless noted)

// sort.cu: stable LSD radix sort of a permutation by the key words of `rows`.
// On return d_perm (u32[n]) holds the sorted order.  Synchronises once to read
// the key ranges.
int32_t mz_sort_perm(mzgpu_ctx* ctx, int row_bytes, const void* d_rows, u64 n, DevMem* perm_out);

// consolidate.cu
// rows[perm] gathered into a dense sorted array
int32_t mz_gather_rows(mzgpu_ctx* ctx, int row_bytes, const void* d_rows, const u32* d_perm, u64 n,
                       void* d_out);
// Sum diffs of equal neighbours in a sorted array and drop zeros.  Output is
// written to d_out (capacity n rows); *n_out is read back (one sync).
int32_t mz_consolidate_sorted(mzgpu_ctx* ctx, int row_bytes, const void* d_sorted, u64 n,
                              void* d_out, u64* n_out);
// sort + gather + consolidate; output array is allocated (capacity n rows).
int32_t mz_sort_consolidate(mzgpu_ctx* ctx, int row_bytes, const void* d_rows, u64 n, DevMem* out,
                            u64* n_out);

// fused.cu: the same pipeline (sort + gather + consolidate, optionally + hash
// index) as ONE cooperative kernel, for small / medium inputs.  If the composite
// key needs more than 64 bits `fallback` is set and nothing else is valid.
struct FusedResult {
  DevMem rows;   // consolidated output (capacity n rows)
  DevMem table;  // hash index (if requested)
  u64 n_out = 0, n_keys = 0, slots = 0;
  u64 min_time = 0, max_time = 0;  // range of the time word over the INPUT rows
  bool fallback = false;
};
int32_t mz_fused_sort_consolidate(mzgpu_ctx* ctx, int row_bytes, const void* d_rows, u64 n, bool want_index,
                                  FusedResult* res);
// inputs up to this many rows take the fused path
#define MZ_FUSED_MAX_ROWS (2u << 20)

// merge.cu: merge two sorted consolidated arrays; times are advanced to
// max(time, since) on the way; result is consolidated.
int32_t mz_merge_consolidate(mzgpu_ctx* ctx, int row_bytes, const void* d_a, u64 na, const void* d_b,
                             u64 nb, u64 since, DevMem* out, u64* n_out);
// stable partition of sorted rows by time < upper: ship (t < upper) and keep.
int32_t mz_extract(mzgpu_ctx* ctx, int row_bytes, const void* d_rows, u64 n, u64 upper, DevMem* ship,
                   u64* n_ship, DevMem* keep, u64* n_keep, u64* min_keep_time);

// index.cu: open-addressing hash index over the distinct keys of a sorted array.
struct HashSlot {
  u64 key;
  u64 meta;  // 0 = empty, else (first row index + 1)
};
int32_t mz_count_keys(mzgpu_ctx* ctx, int row_bytes, const void* d_rows, u64 n, u64* n_keys);
int32_t mz_build_index(mzgpu_ctx* ctx, int row_bytes, const void* d_rows, u64 n, u64 n_keys,
                       DevMem* table, u64* table_slots);

// probe.cu
struct BatchView {  // device-visible description of one batch of a trace
  const u64* rows;
  const HashSlot* table;
  u64 n;
  u64 mask;  // table_slots - 1
};
#define MZ_MAX_TRACE_BATCHES 64
struct TraceView {
  BatchView b[MZ_MAX_TRACE_BATCHES];
  u32 n_batches;
};
#define MZ_PROBE_HALF_LE 0
#define MZ_PROBE_HALF_LT 1
#define MZ_PROBE_JOIN 2  // join_core: no time filter, time = max(t1, t2, meet)
struct ProbeParams {
  int mode;
  u64 meet;         // join_core capability time
  int has_closure;  // 0: identity -> R40 (key, v1, v2)
  int swap_vals;    // join_core side 1: probe rows are val2, lookup rows are val1
  mzgpu_closure closure;
};
// Probe `n` R32 stream rows against the trace; appends results to d_out
// (allocated here) and returns the count.  One sync (to size the output).
int32_t mz_probe(mzgpu_ctx* ctx, const u64* d_stream, u64 n, const TraceView& trace,
                 const ProbeParams& pp, DevMem* out, u64* n_out);
// Apply a closure to R32 rows (val2 = 0); optional skip of rows at `skip_time`.
int32_t mz_map_rows_dev(mzgpu_ctx* ctx, const u64* d_rows, u64 n, const mzgpu_closure* closure,
                        u64 skip_time, DevMem* out, u64* n_out);

// reduce.cu
int32_t mz_explode(mzgpu_ctx* ctx, const u64* d_r32, u64 n, int agg_kind, u64* d_racc);
int32_t mz_reduce_corrections(mzgpu_ctx* ctx, const u64* d_batch_rows, u64 n, const TraceView& prior,
                              int agg_kind, DevMem* out, u64* n_out);

// exchange.cu
int32_t mz_partition(mzgpu_ctx* ctx, int row_bytes, const void* d_rows, u64 n, u32 peers, void* d_out,
                     u64* h_counts /* peers */);
