// fused.cu — the whole "sort + consolidate (+ index)" pipeline as ONE cooperative
// kernel for small and medium inputs (update batches).
//
// Why: a 100K-row update batch (BASELINE config 3) moves ~3 MB; every kernel on
// it finishes in microseconds, so a chain of ~25 dependent launches plus two
// host round trips per sort is pure latency.  Here a persistent grid (<= the
// co-resident capacity, launched cooperatively) runs every phase back to back
// with a device-side grid barrier between phases:
//
//   analyze (min/max per key word)            -> plan computed ON DEVICE by every CTA
//   pack composite keys + all-digit histogram
//   exclusive scan of the histograms
//   P x 8-bit radix passes (rs_tile_pass, decoupled look-back, 1024-key tiles)
//   gather rows by the permutation + head flags (from the sorted composites)
//   warp-segmented diff sums (atomics per (warp, segment))
//   non-zero flags -> compaction -> output rows
//   [optional] hash index over the distinct keys of the output
//
// One launch and one host read-back (counts, time range, fallback flag) replace
// the ~25 launches + 2 syncs of the unfused path (sort.cu + consolidate.cu +
// index.cu), which remains the path for large inputs and for composites wider
// than 64 bits (`fallback`).  Same reference semantics as those files.
#include "common.cuh"
#include "radix.cuh"

namespace {

constexpr int FT = RS_THREADS;  // 256 threads per CTA
constexpr int FI = 4;           // radix items per thread: 1024-key tiles
constexpr int FTILE = FT * FI;

struct FusedCtl {
  u32 barrier;
  u32 fallback;
  u32 pad[2];
  u64 minmax[12];  // [2k] = min (init ~0), [2k+1] = max (init 0)
  u64 n_seg;
  u64 n_out;
  u64 n_keys;
  u32 hist[8 * 256];
};

struct FusedArgs {
  const u64* rows;
  u64 n;
  FusedCtl* ctl;
  u64* k0;
  u64* k1;
  u32* v0;
  u32* v1;
  u32* tile_state;  // [8][T][256], zeroed in phase 0
  u64 T;            // radix tiles
  u64* sorted;      // n rows
  u32* tile_cnt;    // [U] per-256-row-tile counts (heads, then non-zero segments)
  u64* seg_sums;    // [n][ND]
  u32* seg_first;   // [n]
  u64* out;         // n rows capacity
  HashSlot* table;  // optional
  u64 mask;
};

__device__ __forceinline__ void grid_barrier(u32* counter, u32 G, u32& epoch) {
  __syncthreads();
  epoch++;
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    const u32 target = epoch * G;
    while (*(volatile u32*)counter < target) {
    }
    __threadfence();
  }
  __syncthreads();
}

__device__ __forceinline__ int bit_width_dev(u64 x) { return x == 0 ? 0 : 64 - __clzll((long long)x); }

// sum of a u32 array prefix [0, m) by the whole CTA (m is small: n / 256 tiles)
__device__ __forceinline__ u32 block_sum_prefix(const u32* a, u64 m, u32* sm) {
  u32 v = 0;
  for (u64 i = threadIdx.x; i < m; i += FT) v += *(volatile const u32*)(a + i);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  if (lane_id() == 0) sm[warp_id()] = v;
  __syncthreads();
  u32 tot = 0;
  for (int w = 0; w < FT / 32; ++w) tot += sm[w];
  __syncthreads();
  return tot;
}

union FusedSmem {
  RsSmemT<FI> rs;
  u32 hist[8 * 256];
};

template <int RB>
__global__ void __launch_bounds__(FT) k_fused_sort_consolidate(const FusedArgs a) {
  constexpr int NW = RowT<RB>::NW, NK = RowT<RB>::NK, ND = RowT<RB>::ND;
  __shared__ FusedSmem sm;
  __shared__ u32 sm_scan[34];
  __shared__ int s_nwords, s_word[6], s_shift[6], s_npass;
  __shared__ u64 s_minv[6];
  const u32 G = gridDim.x, c = blockIdx.x, tid = threadIdx.x;
  const u64 gtid = (u64)c * FT + tid, gstride = (u64)G * FT;
  const u64 n = a.n;
  FusedCtl* ctl = a.ctl;
  u32 epoch = 0;

  // ---- phase 0: zero scratch; phase 1: min/max of every key word
  {
    const u64 n_state = 8ull * a.T * 256;
    for (u64 i = gtid; i < n_state; i += gstride) a.tile_state[i] = 0;
    for (u64 i = gtid; i < n * ND; i += gstride) a.seg_sums[i] = 0;
    if (a.table != nullptr)
      for (u64 i = gtid; i < (a.mask + 1) * 2; i += gstride) ((u64*)a.table)[i] = 0;
    u64 mn[NK], mx[NK];
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      mn[k] = ~0ull;
      mx[k] = 0;
    }
    for (u64 i = gtid; i < n; i += gstride) {
      const u64* p = a.rows + i * NW;
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        u64 v = p[k];
        mn[k] = v < mn[k] ? v : mn[k];
        mx[k] = v > mx[k] ? v : mx[k];
      }
    }
#pragma unroll
    for (int k = 0; k < NK; ++k) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        u64 x = __shfl_xor_sync(0xffffffffu, mn[k], off);
        u64 y = __shfl_xor_sync(0xffffffffu, mx[k], off);
        mn[k] = x < mn[k] ? x : mn[k];
        mx[k] = y > mx[k] ? y : mx[k];
      }
      if (lane_id() == 0) {
        atomicMin((unsigned long long*)&ctl->minmax[2 * k], (unsigned long long)mn[k]);
        atomicMax((unsigned long long*)&ctl->minmax[2 * k + 1], (unsigned long long)mx[k]);
      }
    }
  }
  grid_barrier(&ctl->barrier, G, epoch);

  // ---- plan (every CTA computes the same plan from the global min/max)
  if (tid == 0) {
    int used = 0, nwords = 0;
    bool fb = false;
    for (int k = NK - 1; k >= 0; --k) {
      u64 lo = *(volatile u64*)&ctl->minmax[2 * k], hi = *(volatile u64*)&ctl->minmax[2 * k + 1];
      int bits = bit_width_dev(hi - lo);
      if (bits == 0) continue;
      if (used + bits > 64) {
        fb = true;
        break;
      }
      s_word[nwords] = k;
      s_shift[nwords] = used;
      s_minv[nwords] = lo;
      nwords++;
      used += bits;
    }
    s_nwords = nwords;
    s_npass = fb ? -1 : (used + 7) / 8;
    if (fb && c == 0) ctl->fallback = 1;
  }
  __syncthreads();
  const int npass = s_npass;
  if (npass < 0) return;  // composite wider than 64 bits: the host takes the unfused path

  // ---- phase 2: pack composites + histogram of every digit place
  for (int i = tid; i < 8 * 256; i += FT) sm.hist[i] = 0;
  __syncthreads();
  for (u64 i = gtid; i < n; i += gstride) {
    const u64* p = a.rows + i * NW;
    u64 comp = 0;
    for (int j = 0; j < s_nwords; ++j) comp |= (p[s_word[j]] - s_minv[j]) << s_shift[j];
    a.k0[i] = comp;
    a.v0[i] = (u32)i;
    for (int ps = 0; ps < npass; ++ps) atomicAdd(&sm.hist[ps * 256 + (u32)((comp >> (8 * ps)) & 255)], 1u);
  }
  __syncthreads();
  for (int i = tid; i < npass * 256; i += FT) {
    u32 v = sm.hist[i];
    if (v) atomicAdd(&ctl->hist[i], v);
  }
  grid_barrier(&ctl->barrier, G, epoch);

  // ---- phase 3: exclusive scan of each pass's histogram (one CTA per pass)
  for (int ps = c; ps < npass; ps += G) {
    u32 v = *(volatile u32*)&ctl->hist[ps * 256 + tid];
    u32 total;
    u32 ex = block_exclusive_scan(v, sm_scan, &total);
    ctl->hist[ps * 256 + tid] = ex;
  }
  grid_barrier(&ctl->barrier, G, epoch);

  // ---- phase 4: radix passes.  CTA c takes tiles c, c+G, ...: every predecessor
  // of a tile is finished or in flight on a co-resident CTA (look-back progress).
  u64* kin = a.k0;
  u64* kout = a.k1;
  u32* vin = a.v0;
  u32* vout = a.v1;
  for (int ps = 0; ps < npass; ++ps) {
    for (u64 t = c; t < a.T; t += G)
      rs_tile_pass<FI>(sm.rs, (u32)t, kin, vin, kout, vout, n, 8 * ps, ctl->hist + ps * 256,
                       a.tile_state + (u64)ps * a.T * 256);
    grid_barrier(&ctl->barrier, G, epoch);
    u64* tk = kin;
    kin = kout;
    kout = tk;
    u32* tv = vin;
    vin = vout;
    vout = tv;
  }
  const u64* keys = kin;  // sorted composites
  const u32* perm = vin;  // sorted row indices (identity order if npass == 0)

  // ---- phase 5: gather rows, head flags, per-tile head counts
  const u64 U = (n + FT - 1) / FT;
  for (u64 u = c; u < U; u += G) {
    const u64 i = u * FT + tid;
    u32 flag = 0;
    if (i < n) {
      u64 r[NW];
      load_row<NW>(a.rows, perm[i], r);
      store_row<NW>(a.sorted, i, r);
      flag = (i == 0) ? 1u : (keys[i] != keys[i - 1] ? 1u : 0u);
    }
    u32 total;
    block_exclusive_scan(flag, sm_scan, &total);
    if (tid == 0) a.tile_cnt[u] = total;
  }
  grid_barrier(&ctl->barrier, G, epoch);

  // ---- phase 6: segmented sums
  for (u64 u = c; u < U; u += G) {
    const u32 base = block_sum_prefix(a.tile_cnt, u, sm_scan);
    const u64 i = u * FT + tid;
    const bool valid = i < n;
    u32 flag = 0;
    if (valid) flag = (i == 0) ? 1u : (keys[i] != keys[i - 1] ? 1u : 0u);
    u32 total;
    u32 ex = block_exclusive_scan(flag, sm_scan, &total);
    u32 seg = valid ? base + ex + flag - 1 : 0xffffffffu;
    u64 d[ND];
#pragma unroll
    for (int w = 0; w < ND; ++w) d[w] = valid ? a.sorted[i * NW + NK + w] : 0;
    if (valid && flag) a.seg_first[seg] = (u32)i;
    const u32 lane = lane_id();
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      u32 oseg = __shfl_up_sync(0xffffffffu, seg, off);
      u64 o[ND];
#pragma unroll
      for (int w = 0; w < ND; ++w) o[w] = __shfl_up_sync(0xffffffffu, d[w], off);
      if (lane >= (u32)off && oseg == seg) diff_add<ND>(d, o);
    }
    u32 nseg = __shfl_down_sync(0xffffffffu, seg, 1);
    if (valid && (lane == 31 || nseg != seg)) {
      u64* acc = a.seg_sums + (u64)seg * ND;
      if (ND == 8) {
        if (d[0]) atomicAdd((unsigned long long*)&acc[0], (unsigned long long)d[0]);
        if (d[1]) atomicAdd((unsigned long long*)&acc[1], (unsigned long long)d[1]);
        u64 old = atomicAdd((unsigned long long*)&acc[2], (unsigned long long)d[2]);
        u64 hi = d[3] + ((old + d[2]) < old ? 1 : 0);
        if (hi) atomicAdd((unsigned long long*)&acc[3], (unsigned long long)hi);
        if (d[4]) atomicAdd((unsigned long long*)&acc[4], (unsigned long long)d[4]);
        if (d[5]) atomicAdd((unsigned long long*)&acc[5], (unsigned long long)d[5]);
        if (d[6]) atomicAdd((unsigned long long*)&acc[6], (unsigned long long)d[6]);
      } else {
        atomicAdd((unsigned long long*)&acc[0], (unsigned long long)d[0]);
      }
    }
    if (u == U - 1 && tid == 0) ctl->n_seg = (u64)base + total;
  }
  grid_barrier(&ctl->barrier, G, epoch);

  // ---- phase 7: non-zero segments per tile
  const u64 S = *(volatile u64*)&ctl->n_seg;
  const u64 V = (S + FT - 1) / FT;
  for (u64 v = c; v < V; v += G) {
    const u64 s = v * FT + tid;
    u32 flag = 0;
    if (s < S) {
      u64 d[ND];
#pragma unroll
      for (int w = 0; w < ND; ++w) d[w] = *(volatile u64*)&a.seg_sums[s * ND + w];
      flag = diff_is_zero<ND>(d) ? 0u : 1u;
    }
    u32 total;
    block_exclusive_scan(flag, sm_scan, &total);
    if (tid == 0) a.tile_cnt[v] = total;
  }
  grid_barrier(&ctl->barrier, G, epoch);

  // ---- phase 8: emit surviving rows
  for (u64 v = c; v < V; v += G) {
    const u32 base = block_sum_prefix(a.tile_cnt, v, sm_scan);
    const u64 s = v * FT + tid;
    u32 flag = 0;
    u64 d[ND];
    if (s < S) {
#pragma unroll
      for (int w = 0; w < ND; ++w) d[w] = *(volatile u64*)&a.seg_sums[s * ND + w];
      flag = diff_is_zero<ND>(d) ? 0u : 1u;
    }
    u32 total;
    u32 ex = block_exclusive_scan(flag, sm_scan, &total);
    if (flag) {
      u64 r[NW];
      load_row<NW>(a.sorted, a.seg_first[s], r);
#pragma unroll
      for (int w = 0; w < ND; ++w) r[NK + w] = d[w];
      store_row<NW>(a.out, (u64)base + ex, r);
    }
    if (v == V - 1 && tid == 0) ctl->n_out = (u64)base + total;
  }
  if (a.table == nullptr) return;
  grid_barrier(&ctl->barrier, G, epoch);

  // ---- phase 9: hash index over the distinct keys of the output
  const u64 n_out = V == 0 ? 0 : *(volatile u64*)&ctl->n_out;
  for (u64 i = gtid; i < ((n_out + 31) / 32) * 32; i += gstride) {
    bool head = false;
    u64 key = 0;
    if (i < n_out) {
      key = a.out[i * NW];
      head = (i == 0) || a.out[(i - 1) * NW] != key;
    }
    u32 m = __ballot_sync(0xffffffffu, head);
    if (lane_id() == 0 && m) atomicAdd((unsigned long long*)&ctl->n_keys, (unsigned long long)__popc(m));
    if (head) {
      u64 h = mix64(key) & a.mask;
      while (true) {
        unsigned long long prev =
            atomicCAS((unsigned long long*)&a.table[h].meta, 0ull, (unsigned long long)(i + 1));
        if (prev == 0ull) {
          a.table[h].key = key;
          break;
        }
        h = (h + 1) & a.mask;
      }
    }
  }
}

template <int RB>
int32_t fused_t(mzgpu_ctx* ctx, const u64* rows, u64 n, bool want_index, FusedResult* res) {
  constexpr int ND = RowT<RB>::ND, NK = RowT<RB>::NK, TW = RowT<RB>::TW;
  static int max_ctas = 0;
  if (max_ctas == 0) {
    int per_sm = 0;
    MZ_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fused_sort_consolidate<RB>, FT, 0));
    max_ctas = per_sm * ctx->num_sms;
    if (max_ctas <= 0) {
      MZ_SET_ERR(ctx, "fused kernel cannot be made resident");
      return MZGPU_E_CUDA;
    }
  }
  FusedArgs a;
  memset(&a, 0, sizeof(a));
  a.rows = rows;
  a.n = n;
  a.T = (n + FTILE - 1) / FTILE;
  u64 slots = 0;
  if (want_index) {
    slots = 2;
    while (slots < 2 * n) slots <<= 1;
  }
  DevMem ctl, kv, state, sorted, tiles, sums, first;
  MZ_TRY(ctl.alloc(ctx, sizeof(FusedCtl)));
  MZ_TRY(kv.alloc(ctx, n * 24));
  MZ_TRY(state.alloc(ctx, 8ull * a.T * 256 * 4));
  MZ_TRY(sorted.alloc(ctx, n * RB));
  MZ_TRY(tiles.alloc(ctx, ((n + FT - 1) / FT) * 4));
  MZ_TRY(sums.alloc(ctx, n * ND * 8));
  MZ_TRY(first.alloc(ctx, n * 4));
  MZ_TRY(res->rows.alloc(ctx, n * RB));
  if (want_index) MZ_TRY(res->table.alloc(ctx, slots * sizeof(HashSlot)));
  // control block: zero, then the min/max identities
  FusedCtl* h = (FusedCtl*)ctx->h_fused;
  memset(h, 0, sizeof(FusedCtl));
  for (int k = 0; k < 6; ++k) h->minmax[2 * k] = ~0ull;
  MZ_CUDA(ctx, cudaMemcpyAsync(ctl.p, h, sizeof(FusedCtl), cudaMemcpyHostToDevice, ctx->stream));
  a.ctl = ctl.as<FusedCtl>();
  a.k0 = kv.as<u64>();
  a.k1 = a.k0 + n;
  a.v0 = (u32*)(a.k1 + n);
  a.v1 = a.v0 + n;
  a.tile_state = state.as<u32>();
  a.sorted = sorted.as<u64>();
  a.tile_cnt = tiles.as<u32>();
  a.seg_sums = sums.as<u64>();
  a.seg_first = first.as<u32>();
  a.out = res->rows.template as<u64>();
  a.table = want_index ? res->table.template as<HashSlot>() : nullptr;
  a.mask = want_index ? slots - 1 : 0;
  u64 want = a.T > ((n + FT - 1) / FT + 3) / 4 ? a.T : ((n + FT - 1) / FT + 3) / 4;
  unsigned grid = (unsigned)(want < (u64)max_ctas ? want : (u64)max_ctas);
  if (grid == 0) grid = 1;
  void* kargs[] = {(void*)&a};
  {
    MZ_BYTES(ctx, n * RB * 4);  // rows read (analyze, pack, gather) + sorted + out written; see DESIGN.md
    ProfScope prof(ctx, "k_fused_sort_consolidate");
    cudaError_t e = cudaLaunchCooperativeKernel((void*)k_fused_sort_consolidate<RB>, dim3(grid), dim3(FT), kargs,
                                                0, ctx->stream);
    if (e != cudaSuccess) {
      MZ_SET_ERR(ctx, "cooperative launch failed: %s", cudaGetErrorString(e));
      ctx->sticky = true;
      return MZGPU_E_CUDA;
    }
  }
  ctx->stats.kernel_launches++;
  MZ_CUDA(ctx, cudaMemcpyAsync(h, ctl.p, offsetof(FusedCtl, hist), cudaMemcpyDeviceToHost, ctx->stream));
  MZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->stats.d2h_bytes += offsetof(FusedCtl, hist);
  res->fallback = h->fallback != 0;
  res->n_out = h->n_out;
  res->n_keys = h->n_keys;
  res->slots = slots;
  res->min_time = TW >= 0 ? h->minmax[2 * (TW >= 0 ? TW : 0)] : 0;
  res->max_time = TW >= 0 ? h->minmax[2 * (TW >= 0 ? TW : 0) + 1] : 0;
  (void)NK;
  return MZGPU_OK;
}

}  // namespace

int32_t mz_fused_sort_consolidate(mzgpu_ctx* ctx, int row_bytes, const void* d_rows, u64 n, bool want_index,
                                  FusedResult* res) {
  const u64* r = (const u64*)d_rows;
  switch (row_bytes) {
    case 16: return fused_t<16>(ctx, r, n, want_index, res);
    case 32: return fused_t<32>(ctx, r, n, want_index, res);
    case 40: return fused_t<40>(ctx, r, n, want_index, res);
    case 80: return fused_t<80>(ctx, r, n, want_index, res);
    case 64: return fused_t<64>(ctx, r, n, want_index, res);
    default:
      MZ_SET_ERR(ctx, "fused: unsupported row width %d", row_bytes);
      return MZGPU_E_UNSUPPORTED;
  }
}
