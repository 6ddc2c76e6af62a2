// column.cu — device side of the columnar wire format.
//
// Reference: `Column<C>` (src/timely-util/src/columnar.rs:54-222) serialized with
// `columnar::bytes::indexed` (crate columnar 0.12.1; layout pinned by columnar.rs:247-258) and
// produced by `ColumnBuilder` (src/timely-util/src/columnar/builder.rs:28-111).  A container is
// struct-of-arrays: an index of k + 1 byte offsets, then k word-padded slices (one per leaf column
// of `((K, V), T, R)`); the operators of this library work on packed rows (array-of-structs).  The
// kernels here are the transposition between the two, HBM-bound at 2 * row_bytes per row:
//   k_col_decode_fixed / k_col_encode_fixed   ((u64, u64), u64, i64) <-> R32, (u64, i64) <-> R16
//   k_col_decode_rows  / k_col_encode_rows    ((Row, Row), Timestamp, Diff) <-> R32 with Rows of at
//                                             most 7 bytes packed into one order-preserving word
//                                             (mzgpu_rowkey_pack; RowRef::cmp, src/repr/src/row.rs:704-722)
//   k_col_lens_reduce / k_col_scan_bsums / k_col_prefix   inclusive prefix sums of the Row byte
//                                             lengths (the `Rows` bounds, src/repr/src/row.rs:606-611)
//   k_col_cuts                                ColumnBuilder's ship points for Row containers
#include "common.cuh"

namespace {

struct ColOff {
  u64 w[6];  // word offset of each slice inside the container
};

// ---- fixed-width layouts: column c of row i is word off.w[c] + i
template <int NW>
__global__ void __launch_bounds__(256) k_col_decode_fixed(const u64* __restrict__ words, u64 n, ColOff off,
                                                          u64* __restrict__ dst, u64 base) {
  for (u64 i = (u64)blockIdx.x * 256 + threadIdx.x; i < n; i += (u64)gridDim.x * 256) {
    u64 r[NW];
#pragma unroll
    for (int c = 0; c < NW; ++c) r[c] = words[off.w[c] + i];
    store_row<NW>(dst, base + i, r);
  }
}
template <int NW>
__global__ void __launch_bounds__(256) k_col_encode_fixed(const u64* __restrict__ rows, u64 first, u64 n, ColOff off,
                                                          u64* __restrict__ words) {
  const u64 gtid = (u64)blockIdx.x * 256 + threadIdx.x;
  if (gtid == 0) {
    // the index: 8 * (k + 1), then the END of every slice in bytes (all slices are whole words)
    words[0] = 8 * (u64)(NW + 1);
#pragma unroll
    for (int c = 0; c < NW; ++c) words[1 + c] = 8 * (off.w[c] + n);
  }
  for (u64 i = gtid; i < n; i += (u64)gridDim.x * 256) {
    u64 r[NW];
    load_row<NW>(rows, first + i, r);
#pragma unroll
    for (int c = 0; c < NW; ++c) words[off.w[c] + i] = r[c];
  }
}

// ---- Row layout
// flag word: 1 = a Row longer than 7 bytes, 2 = bounds not monotone / outside the bytes slice
__global__ void __launch_bounds__(256) k_col_decode_rows(const u64* __restrict__ words, u64 n, ColOff off,
                                                         u64 key_bytes, u64 val_bytes, u64* __restrict__ dst, u64 base,
                                                         u64* __restrict__ flag) {
  const u64* kb = words + off.w[0];
  const unsigned char* kv = (const unsigned char*)(words + off.w[1]);
  const u64* vb = words + off.w[2];
  const unsigned char* vv = (const unsigned char*)(words + off.w[3]);
  for (u64 i = (u64)blockIdx.x * 256 + threadIdx.x; i < n; i += (u64)gridDim.x * 256) {
    const u64 klo = i ? kb[i - 1] : 0, khi = kb[i], vlo = i ? vb[i - 1] : 0, vhi = vb[i];
    u64 r[4] = {0, 0, words[off.w[4] + i], words[off.w[5] + i]};
    if (khi < klo || khi > key_bytes || vhi < vlo || vhi > val_bytes) {
      atomicMax((unsigned long long*)flag, 2ull);
    } else if (khi - klo > 7 || vhi - vlo > 7) {
      atomicMax((unsigned long long*)flag, 1ull);
    } else {
      u64 k = (khi - klo) << 56, v = (vhi - vlo) << 56;
      for (u64 b = klo; b < khi; ++b) k |= (u64)kv[b] << (8 * (6 - (b - klo)));
      for (u64 b = vlo; b < vhi; ++b) v |= (u64)vv[b] << (8 * (6 - (b - vlo)));
      r[0] = k;
      r[1] = v;
    }
    store_row<4>(dst, base + i, r);
  }
}

// Row byte lengths of rows [first, first + n): per-block sums (2048 rows per block)
#define COL_ROWS_PER_BLOCK 2048
__device__ __forceinline__ u64 block_sum_256(u64 v, u64* sh) {
  for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  u64 t = 0;
  if (threadIdx.x < 8) t = sh[threadIdx.x];
  if (threadIdx.x < 32)
    for (int o = 4; o; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
  __syncthreads();
  return t;  // valid in thread 0
}
__global__ void __launch_bounds__(256) k_col_lens_reduce(const u64* __restrict__ rows, u64 first, u64 n,
                                                         u64* __restrict__ bsum) {
  __shared__ u64 sh[8];
  const u64 lo = (u64)blockIdx.x * COL_ROWS_PER_BLOCK;
  u64 k = 0, v = 0;
  for (int j = 0; j < COL_ROWS_PER_BLOCK / 256; ++j) {
    const u64 i = lo + (u64)j * 256 + threadIdx.x;
    if (i < n) {
      k += rows[(first + i) * 4] >> 56;
      v += rows[(first + i) * 4 + 1] >> 56;
    }
  }
  k = block_sum_256(k, sh);
  v = block_sum_256(v, sh);
  if (threadIdx.x == 0) {
    bsum[2 * (u64)blockIdx.x] = k;
    bsum[2 * (u64)blockIdx.x + 1] = v;
  }
}
// exclusive scan of the per-block sums in place (one CTA); totals to tot[0..1]
__global__ void __launch_bounds__(1024) k_col_scan_bsums(u64* __restrict__ bsum, u64 nb, u64* __restrict__ tot) {
  __shared__ u64 sk[1024], sv[1024];
  const u64 per = (nb + 1023) / 1024, lo = threadIdx.x * per, hi = lo + per < nb ? lo + per : nb;
  u64 k = 0, v = 0;
  for (u64 b = lo; b < hi; ++b) k += bsum[2 * b], v += bsum[2 * b + 1];
  sk[threadIdx.x] = k, sv[threadIdx.x] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    u64 ak = 0, av = 0;
    for (int t = 0; t < 1024; ++t) {
      const u64 tk = sk[t], tv = sv[t];
      sk[t] = ak, sv[t] = av;
      ak += tk, av += tv;
    }
    tot[0] = ak, tot[1] = av;
  }
  __syncthreads();
  k = sk[threadIdx.x], v = sv[threadIdx.x];
  for (u64 b = lo; b < hi; ++b) {
    const u64 tk = bsum[2 * b], tv = bsum[2 * b + 1];
    bsum[2 * b] = k, bsum[2 * b + 1] = v;
    k += tk, v += tv;
  }
}
// inclusive prefix sums pk[i], pv[i] of the key / val byte lengths (row i of the range)
__global__ void __launch_bounds__(256) k_col_prefix(const u64* __restrict__ rows, u64 first, u64 n,
                                                    const u64* __restrict__ bsum, u64* __restrict__ pk,
                                                    u64* __restrict__ pv) {
  __shared__ u64 shk[256], shv[256];
  const u64 lo = (u64)blockIdx.x * COL_ROWS_PER_BLOCK + (u64)threadIdx.x * (COL_ROWS_PER_BLOCK / 256);
  u64 lk[COL_ROWS_PER_BLOCK / 256], lv[COL_ROWS_PER_BLOCK / 256];
  u64 k = 0, v = 0;
#pragma unroll
  for (int j = 0; j < COL_ROWS_PER_BLOCK / 256; ++j) {
    const u64 i = lo + j;
    lk[j] = i < n ? rows[(first + i) * 4] >> 56 : 0;
    lv[j] = i < n ? rows[(first + i) * 4 + 1] >> 56 : 0;
    k += lk[j], v += lv[j];
  }
  shk[threadIdx.x] = k, shv[threadIdx.x] = v;
  __syncthreads();
  // Hillis-Steele over the 256 thread totals
  for (int o = 1; o < 256; o <<= 1) {
    u64 ak = 0, av = 0;
    if ((int)threadIdx.x >= o) ak = shk[threadIdx.x - o], av = shv[threadIdx.x - o];
    __syncthreads();
    shk[threadIdx.x] += ak, shv[threadIdx.x] += av;
    __syncthreads();
  }
  u64 bk = bsum[2 * (u64)blockIdx.x] + shk[threadIdx.x] - k, bv = bsum[2 * (u64)blockIdx.x + 1] + shv[threadIdx.x] - v;
#pragma unroll
  for (int j = 0; j < COL_ROWS_PER_BLOCK / 256; ++j) {
    const u64 i = lo + j;
    bk += lk[j], bv += lv[j];
    if (i < n) pk[i] = bk, pv[i] = bv;
  }
}
// container of rows [s, s + n) of the range whose prefix sums are pk / pv (relative to the range)
__global__ void __launch_bounds__(256) k_col_encode_rows(const u64* __restrict__ rows, u64 first, u64 s, u64 n,
                                                         const u64* __restrict__ pk, const u64* __restrict__ pv,
                                                         ColOff off, u64* __restrict__ words) {
  const u64 gtid = (u64)blockIdx.x * 256 + threadIdx.x;
  const u64 k0 = s ? pk[s - 1] : 0, v0 = s ? pv[s - 1] : 0;
  if (gtid == 0) {
    const u64 kbytes = n ? pk[s + n - 1] - k0 : 0, vbytes = n ? pv[s + n - 1] - v0 : 0;
    words[0] = 8 * 7;
    words[1] = 8 * (off.w[0] + n);
    words[2] = 8 * off.w[1] + kbytes;
    words[3] = 8 * (off.w[2] + n);
    words[4] = 8 * off.w[3] + vbytes;
    words[5] = 8 * (off.w[4] + n);
    words[6] = 8 * (off.w[5] + n);
  }
  unsigned char* kv = (unsigned char*)(words + off.w[1]);
  unsigned char* vv = (unsigned char*)(words + off.w[3]);
  for (u64 i = gtid; i < n; i += (u64)gridDim.x * 256) {
    u64 r[4];
    load_row<4>(rows, first + s + i, r);
    const u64 khi = pk[s + i] - k0, vhi = pv[s + i] - v0;
    const u64 kl = r[0] >> 56, vl = r[1] >> 56;
    words[off.w[0] + i] = khi;
    words[off.w[2] + i] = vhi;
    for (u64 b = 0; b < kl; ++b) kv[khi - kl + b] = (unsigned char)(r[0] >> (8 * (6 - b)));
    for (u64 b = 0; b < vl; ++b) vv[vhi - vl + b] = (unsigned char)(r[1] >> (8 * (6 - b)));
    words[off.w[4] + i] = r[2];
    words[off.w[5] + i] = r[3];
  }
}
// ColumnBuilder::push_into's ship points for Row containers (builder.rs:44-52): a container that
// starts at row s ends with the first row e at which the serialized size reaches the ship window
// (within 10 % of the next multiple of 2^18 words).  The size grows by at most 6 words per push, far
// less than the window's 26213 words, so the first e with words(s, e) >= 2^18 - 2^18 / 10 + 1 is the
// row at which at_capacity first holds: a binary search per container over the prefix sums.
// cuts[3j] = end (exclusive) of container j, cuts[3j + 1 .. 3j + 2] = the prefix sums at that end;
// *n_cuts = containers.
__device__ __forceinline__ u64 col_rowrow_words(u64 rows, u64 kbytes, u64 vbytes) {
  return 7 + 4 * rows + (kbytes + 7) / 8 + (vbytes + 7) / 8;
}
__global__ void k_col_cuts(const u64* __restrict__ pk, const u64* __restrict__ pv, u64 n, u64* __restrict__ cuts,
                           u64 cap, u64* __restrict__ n_cuts) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const u64 ship = (1ull << 18) - (1ull << 18) / 10 + 1;
  u64 s = 0, c = 0;
  while (s < n) {
    const u64 k0 = s ? pk[s - 1] : 0, v0 = s ? pv[s - 1] : 0;
    u64 lo = s, hi = n;  // first e in [s, n) with words(s .. e inclusive) >= ship, else n
    while (lo < hi) {
      const u64 mid = (lo + hi) >> 1;
      if (col_rowrow_words(mid - s + 1, pk[mid] - k0, pv[mid] - v0) >= ship)
        hi = mid;
      else
        lo = mid + 1;
    }
    const u64 e = lo < n ? lo + 1 : n;
    if (c < cap) cuts[3 * c] = e, cuts[3 * c + 1] = pk[e - 1], cuts[3 * c + 2] = pv[e - 1];
    ++c;
    s = e;
  }
  *n_cuts = c;
}

unsigned col_grid(mzgpu_ctx* ctx, u64 n) {
  u64 g = (n + 255) / 256;
  const u64 maxg = (u64)ctx->num_sms * 8;
  if (g > maxg) g = maxg;
  return (unsigned)(g ? g : 1);
}
ColOff col_off(const u64* w) {
  ColOff o;
  for (int i = 0; i < 6; ++i) o.w[i] = w[i];
  return o;
}

}  // namespace

int32_t mz_col_decode_fixed(mzgpu_ctx* ctx, int nw, const u64* d_words, u64 n, const u64* off_words, u64* d_dst,
                            u64 base) {
  if (n == 0) return MZGPU_OK;
  MZ_BYTES(ctx, n * (u64)nw * 16);
  if (nw == 4)
    MZ_LAUNCH(ctx, k_col_decode_fixed<4>, col_grid(ctx, n), 256, 0, d_words, n, col_off(off_words), d_dst, base);
  else
    MZ_LAUNCH(ctx, k_col_decode_fixed<2>, col_grid(ctx, n), 256, 0, d_words, n, col_off(off_words), d_dst, base);
  return MZGPU_OK;
}
int32_t mz_col_encode_fixed(mzgpu_ctx* ctx, int nw, const u64* d_rows, u64 first, u64 n, const u64* off_words,
                            u64* d_words) {
  MZ_BYTES(ctx, n * (u64)nw * 16);
  if (nw == 4)
    MZ_LAUNCH(ctx, k_col_encode_fixed<4>, col_grid(ctx, n), 256, 0, d_rows, first, n, col_off(off_words), d_words);
  else
    MZ_LAUNCH(ctx, k_col_encode_fixed<2>, col_grid(ctx, n), 256, 0, d_rows, first, n, col_off(off_words), d_words);
  return MZGPU_OK;
}
int32_t mz_col_decode_rows(mzgpu_ctx* ctx, const u64* d_words, u64 n, const u64* off_words, u64 key_bytes,
                           u64 val_bytes, u64* d_dst, u64 base, u64* d_flag) {
  if (n == 0) return MZGPU_OK;
  MZ_BYTES(ctx, n * 64 + key_bytes + val_bytes);
  MZ_LAUNCH(ctx, k_col_decode_rows, col_grid(ctx, n), 256, 0, d_words, n, col_off(off_words), key_bytes, val_bytes,
            d_dst, base, d_flag);
  return MZGPU_OK;
}
// pk / pv (n words each) = inclusive prefix sums of the Row byte lengths of rows [first, first + n);
// d_tot[0..1] = totals.  d_bsum: 2 * ceil(n / 2048) words of scratch.
int32_t mz_col_row_prefix(mzgpu_ctx* ctx, const u64* d_rows, u64 first, u64 n, u64* d_bsum, u64* d_pk, u64* d_pv,
                          u64* d_tot) {
  const u64 nb = (n + COL_ROWS_PER_BLOCK - 1) / COL_ROWS_PER_BLOCK;
  if (nb == 0) {
    MZ_CUDA(ctx, cudaMemsetAsync(d_tot, 0, 16, ctx->stream));
    return MZGPU_OK;
  }
  MZ_BYTES(ctx, n * 32);
  MZ_LAUNCH(ctx, k_col_lens_reduce, (unsigned)nb, 256, 0, d_rows, first, n, d_bsum);
  MZ_LAUNCH(ctx, k_col_scan_bsums, 1, 1024, 0, d_bsum, nb, d_tot);
  MZ_BYTES(ctx, n * 48);
  MZ_LAUNCH(ctx, k_col_prefix, (unsigned)nb, 256, 0, d_rows, first, n, d_bsum, d_pk, d_pv);
  return MZGPU_OK;
}
int32_t mz_col_encode_rows(mzgpu_ctx* ctx, const u64* d_rows, u64 first, u64 s, u64 n, const u64* d_pk,
                           const u64* d_pv, const u64* off_words, u64* d_words) {
  MZ_BYTES(ctx, n * 96);
  MZ_LAUNCH(ctx, k_col_encode_rows, col_grid(ctx, n), 256, 0, d_rows, first, s, n, d_pk, d_pv, col_off(off_words),
            d_words);
  return MZGPU_OK;
}
int32_t mz_col_cuts(mzgpu_ctx* ctx, const u64* d_pk, const u64* d_pv, u64 n, u64* d_cuts, u64 cap, u64* d_n_cuts) {
  MZ_LAUNCH(ctx, k_col_cuts, 1, 32, 0, d_pk, d_pv, n, d_cuts, cap, d_n_cuts);
  return MZGPU_OK;
}
