// correction.cu — device side of the MV sink's correction buffer.
//
// Reference: src/compute/src/sink/correction_v2.rs (CorrectionV2: insert / insert_negated :213-277,
// updates_before / consolidate_before :280-374, advance_since / consolidate_at_since :377-390,
// consolidate by (time, data) :1285-1321).  The buffer holds the difference between the desired
// and the persisted contents of a materialized view; the sink reads "all updates before `upper`"
// consolidated, with times advanced to `since`.
//
// The reference keeps chains of chunks sorted by (time, data) with a geometric size invariant so
// that a CPU never re-merges much.  On the GPU the same contract is met the way the Batcher meets
// its own (host.cu): inserts are stashed (time-major rows: (time, key, val | diff), so that the
// generic consolidate kernel sorts by (time, data) with no change), and a read consolidates
// everything buffered in one pass; the updates before `upper` are then a PREFIX of the sorted
// buffer, found by one binary search on the device.
#include "common.cuh"

namespace {

// R32 (key, val, time | diff) -> time-major (max(time, since), key, val | +-diff), appended at
// dst[base ...]; *out_len = base + n
__global__ void __launch_bounds__(256) k_corr_to_td(const u64* __restrict__ src, const DLen dn, u64 since, int negate,
                                                    u64* __restrict__ dst, const DLen dbase, u64 cap_rows,
                                                    u64* __restrict__ out_len, u64* __restrict__ status) {
  const u64 n = dlen_get(dn), base = dlen_get(dbase);
  const u64 gtid = (u64)blockIdx.x * 256 + threadIdx.x;
  u64 m = n;
  if (base + n > cap_rows) {
    if (gtid == 0) atomicMax((unsigned long long*)status, (unsigned long long)(base + n));
    m = cap_rows > base ? cap_rows - base : 0;
  }
  for (u64 i = gtid; i < m; i += (u64)gridDim.x * 256) {
    u64 r[4], o[4];
    load_row<4>(src, i, r);
    o[0] = r[2] < since ? since : r[2];
    o[1] = r[0];
    o[2] = r[1];
    o[3] = negate ? (u64)0 - r[3] : r[3];
    store_row<4>(dst, base + i, o);
  }
  if (gtid == 0) *out_len = base + n;
}
// times of stored rows advanced in place
__global__ void __launch_bounds__(256) k_corr_advance(u64* __restrict__ td, const DLen dn, u64 since) {
  const u64 n = dlen_get(dn);
  for (u64 i = (u64)blockIdx.x * 256 + threadIdx.x; i < n; i += (u64)gridDim.x * 256)
    if (td[i * 4] < since) td[i * 4] = since;
}
// first row of the (time, data)-sorted buffer whose time is >= upper
__global__ void k_corr_split(const u64* __restrict__ td, const DLen dn, u64 upper, u64* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const u64 n = dlen_get(dn);
  u64 lo = 0, hi = n;
  if (upper != MZGPU_FRONTIER_EMPTY) {
    while (lo < hi) {
      const u64 mid = (lo + hi) >> 1;
      if (td[mid * 4] < upper)
        lo = mid + 1;
      else
        hi = mid;
    }
  } else {
    lo = n;
  }
  *out = lo;
}
// the first dn time-major rows back to R32, appended at dst[base ...]
__global__ void __launch_bounds__(256) k_corr_from_td(const u64* __restrict__ td, const DLen dn, u64* __restrict__ dst,
                                                      const DLen dbase, u64 cap_rows, u64* __restrict__ out_len,
                                                      u64* __restrict__ status) {
  const u64 n = dlen_get(dn), base = dlen_get(dbase);
  const u64 gtid = (u64)blockIdx.x * 256 + threadIdx.x;
  u64 m = n;
  if (base + n > cap_rows) {
    if (gtid == 0) atomicMax((unsigned long long*)status, (unsigned long long)(base + n));
    m = cap_rows > base ? cap_rows - base : 0;
  }
  for (u64 i = gtid; i < m; i += (u64)gridDim.x * 256) {
    u64 r[4], o[4];
    load_row<4>(td, i, r);
    o[0] = r[1];
    o[1] = r[2];
    o[2] = r[0];
    o[3] = r[3];
    store_row<4>(dst, base + i, o);
  }
  if (gtid == 0) *out_len = base + n;
}

unsigned corr_grid(mzgpu_ctx* ctx, u64 n_ub) {
  u64 g = (n_ub + 255) / 256;
  const u64 maxg = (u64)ctx->num_sms * 8;
  if (g > maxg) g = maxg;
  return (unsigned)(g ? g : 1);
}

}  // namespace

int32_t mz_corr_to_td(mzgpu_ctx* ctx, const u64* d_rows, DLen n, u64 n_ub, u64 since, bool negate, u64* d_td, DLen base,
                      u64 cap_rows, u64* d_out_len) {
  MZ_LAUNCH(ctx, k_corr_to_td, corr_grid(ctx, n_ub), 256, 0, d_rows, n, since, negate ? 1 : 0, d_td, base, cap_rows,
            d_out_len, ctx->d_status);
  return MZGPU_OK;
}
int32_t mz_corr_advance(mzgpu_ctx* ctx, u64* d_td, DLen n, u64 n_ub, u64 since) {
  if (n_ub == 0) return MZGPU_OK;
  MZ_LAUNCH(ctx, k_corr_advance, corr_grid(ctx, n_ub), 256, 0, d_td, n, since);
  return MZGPU_OK;
}
int32_t mz_corr_split(mzgpu_ctx* ctx, const u64* d_td, DLen n, u64 upper, u64* d_out) {
  MZ_LAUNCH(ctx, k_corr_split, 1, 32, 0, d_td, n, upper, d_out);
  return MZGPU_OK;
}
int32_t mz_corr_from_td(mzgpu_ctx* ctx, const u64* d_td, DLen n, u64 n_ub, u64* d_dst, DLen base, u64 cap_rows,
                        u64* d_out_len) {
  MZ_LAUNCH(ctx, k_corr_from_td, corr_grid(ctx, n_ub), 256, 0, d_td, n, d_dst, base, cap_rows, d_out_len,
            ctx->d_status);
  return MZGPU_OK;
}
