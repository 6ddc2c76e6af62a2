// exchange.cu — the Exchange pact as a partition kernel.
//
// Reference: timely `Exchange(|k| k.hashed())` before every stateful operator
// (src/compute/src/extensions/arrange.rs:116, src/timely-util/src/columnar.rs:227-237,
// half_join's internal exchange).  Routing = FNV-1a 64 of the key's 8 LE bytes
// (Hashable::hashed in DD 0.23) modulo peers; output collections do not depend
// on the routing function.
//
// Rows are bucketed by destination on the device (count -> offsets -> scatter
// with one global atomic per (CTA, destination)); mzgpu_exchange (host.cu) then
// moves the buckets with one grouped ncclSend/ncclRecv all-to-all over NVLink.
#include "common.cuh"

namespace {

constexpr int XT = 256;
constexpr int MAX_PEERS = 64;

__host__ __device__ __forceinline__ u64 fnv1a64(u64 key) {
  u64 h = 0xcbf29ce484222325ull;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    h ^= (key >> (8 * i)) & 0xff;
    h *= 0x100000001b3ull;
  }
  return h;
}

template <int NW>
__global__ void __launch_bounds__(XT) k_part_count(const u64* __restrict__ rows, const DLen dn, u32 peers,
                                                   unsigned long long* __restrict__ counts) {
  const u64 n = dlen_get(dn);
  __shared__ u32 sh[MAX_PEERS];
  if (threadIdx.x < MAX_PEERS) sh[threadIdx.x] = 0;
  __syncthreads();
  for (u64 i = (u64)blockIdx.x * XT + threadIdx.x; i < n; i += (u64)gridDim.x * XT)
    atomicAdd(&sh[(u32)(fnv1a64(rows[i * NW]) % peers)], 1u);
  __syncthreads();
  if (threadIdx.x < peers && sh[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

template <int NW>
__global__ void __launch_bounds__(XT) k_part_scatter(const u64* __restrict__ rows, const DLen dn, u32 peers,
                                                     unsigned long long* __restrict__ cursors,
                                                     u64* __restrict__ out) {
  const u64 n = dlen_get(dn);
  __shared__ u32 sh_count[MAX_PEERS];
  __shared__ u64 sh_base[MAX_PEERS];
  for (u64 i0 = (u64)blockIdx.x * XT; i0 < n; i0 += (u64)gridDim.x * XT) {
    __syncthreads();
    if (threadIdx.x < MAX_PEERS) sh_count[threadIdx.x] = 0;
    __syncthreads();
    const u64 i = i0 + threadIdx.x;
    u32 dest = 0, rank = 0;
    u64 r[NW];
    if (i < n) {
      load_row<NW>(rows, i, r);
      dest = (u32)(fnv1a64(r[0]) % peers);
      rank = atomicAdd(&sh_count[dest], 1u);
    }
    __syncthreads();
    if (threadIdx.x < peers && sh_count[threadIdx.x])
      sh_base[threadIdx.x] = atomicAdd(&cursors[threadIdx.x], (unsigned long long)sh_count[threadIdx.x]);
    __syncthreads();
    if (i < n) store_row<NW>(out, sh_base[dest] + rank, r);
  }
}

// counts[0..P) -> cursors[0..P) = exclusive offsets (counts stay for the count exchange)
__global__ void k_part_offsets(const unsigned long long* __restrict__ counts, u32 peers,
                               unsigned long long* __restrict__ cursors) {
  if (threadIdx.x == 0) {
    unsigned long long off = 0;
    for (u32 p = 0; p < peers; ++p) {
      cursors[p] = off;
      off += counts[p];
    }
  }
}

// ---- several buffers in one launch (blockIdx.y = buffer): the exchange points of one round
struct PartJobs {
  const u64* rows[MZ_MAX_EXCHANGE];
  u64* out[MZ_MAX_EXCHANGE];
  DLen n[MZ_MAX_EXCHANGE];
  int nw[MZ_MAX_EXCHANGE];
};
__global__ void __launch_bounds__(XT) k_part_count_many(const __grid_constant__ PartJobs jobs, u32 peers,
                                                        unsigned long long* __restrict__ counts /* [job][64] */) {
  const int j = blockIdx.y;
  const u64 n = dlen_get(jobs.n[j]);
  const u64* rows = jobs.rows[j];
  const int nw = jobs.nw[j];
  __shared__ u32 sh[MAX_PEERS];
  if (threadIdx.x < MAX_PEERS) sh[threadIdx.x] = 0;
  __syncthreads();
  for (u64 i = (u64)blockIdx.x * XT + threadIdx.x; i < n; i += (u64)gridDim.x * XT)
    atomicAdd(&sh[(u32)(fnv1a64(rows[i * nw]) % peers)], 1u);
  __syncthreads();
  if (threadIdx.x < peers && sh[threadIdx.x])
    atomicAdd(&counts[j * 64 + threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}
__global__ void k_part_offsets_many(const unsigned long long* __restrict__ counts, u32 peers, u32 k,
                                    unsigned long long* __restrict__ cursors,
                                    unsigned long long* __restrict__ send_by_peer /* [peer][k] */) {
  const u32 j = threadIdx.x;
  if (j < k) {
    unsigned long long off = 0;
    for (u32 p = 0; p < peers; ++p) {
      const unsigned long long cnt = counts[j * 64 + p];
      cursors[j * 64 + p] = off;
      send_by_peer[(size_t)p * k + j] = cnt;  // one message per peer carries all k counts
      off += cnt;
    }
  }
}
__global__ void __launch_bounds__(XT) k_part_scatter_many(const __grid_constant__ PartJobs jobs, u32 peers,
                                                          unsigned long long* __restrict__ cursors) {
  const int j = blockIdx.y;
  const u64 n = dlen_get(jobs.n[j]);
  const u64* rows = jobs.rows[j];
  u64* out = jobs.out[j];
  const int nw = jobs.nw[j];
  __shared__ u32 sh_count[MAX_PEERS];
  __shared__ u64 sh_base[MAX_PEERS];
  for (u64 i0 = (u64)blockIdx.x * XT; i0 < n; i0 += (u64)gridDim.x * XT) {
    __syncthreads();
    if (threadIdx.x < MAX_PEERS) sh_count[threadIdx.x] = 0;
    __syncthreads();
    const u64 i = i0 + threadIdx.x;
    u32 dest = 0, rank = 0;
    if (i < n) {
      dest = (u32)(fnv1a64(rows[i * nw]) % peers);
      rank = atomicAdd(&sh_count[dest], 1u);
    }
    __syncthreads();
    if (threadIdx.x < peers && sh_count[threadIdx.x])
      sh_base[threadIdx.x] = atomicAdd(&cursors[j * 64 + threadIdx.x], (unsigned long long)sh_count[threadIdx.x]);
    __syncthreads();
    if (i < n) {
      const u64* src = rows + i * nw;
      u64* dst = out + (sh_base[dest] + rank) * nw;
      for (int w = 0; w < nw; w += 2) {  // rows are 16-byte multiples (nw even: 4 or 10)
        const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(src + w);
        *reinterpret_cast<ulonglong2*>(dst + w) = v;
      }
    }
  }
}

}  // namespace

uint32_t mzgpu_route(uint64_t key, uint32_t peers) { return (uint32_t)(fnv1a64(key) % peers); }

// Bucket rows by destination entirely on the device: d_counts[p] = rows for peer
// p, d_out = rows grouped by destination in peer order.  No host round trip.
int32_t mz_partition(mzgpu_ctx* ctx, int row_bytes, const void* d_rows, DLen n, u64 n_ub, u32 peers, void* d_out,
                     u64* d_counts /* MAX_PEERS words */, u64* d_cursors /* MAX_PEERS words */) {
  if (peers > MAX_PEERS) {
    MZ_SET_ERR(ctx, "exchange: %u peers exceed the supported maximum %d", peers, MAX_PEERS);
    return MZGPU_E_UNSUPPORTED;
  }
  MZ_CUDA(ctx, cudaMemsetAsync(d_counts, 0, MAX_PEERS * 8, ctx->stream));
  if (n_ub == 0) return MZGPU_OK;
  const u64* r = (const u64*)d_rows;
  unsigned long long* c = (unsigned long long*)d_counts;
  unsigned long long* cur = (unsigned long long*)d_cursors;
  u64 blocks = (n_ub + XT - 1) / XT;
  unsigned grid = (unsigned)(blocks < (u64)ctx->num_sms * 8 ? blocks : (u64)ctx->num_sms * 8);
  switch (row_bytes) {
    case 32: MZ_LAUNCH(ctx, k_part_count<4>, grid, XT, 0, r, n, peers, c); break;
    case 80: MZ_LAUNCH(ctx, k_part_count<10>, grid, XT, 0, r, n, peers, c); break;
    default: MZ_SET_ERR(ctx, "exchange: unsupported row width %d", row_bytes); return MZGPU_E_UNSUPPORTED;
  }
  MZ_LAUNCH(ctx, k_part_offsets, 1, 32, 0, c, peers, cur);
  switch (row_bytes) {
    case 32: MZ_LAUNCH(ctx, k_part_scatter<4>, grid, XT, 0, r, n, peers, cur, (u64*)d_out); break;
    case 80: MZ_LAUNCH(ctx, k_part_scatter<10>, grid, XT, 0, r, n, peers, cur, (u64*)d_out); break;
    default: return MZGPU_E_UNSUPPORTED;
  }
  return MZGPU_OK;
}

// All buffers of one exchange round in three launches (count, offsets, scatter).
// d_counts / d_cursors: [k][64] words.
int32_t mz_partition_many(mzgpu_ctx* ctx, u32 k, const int* row_bytes, const void* const* d_rows, const DLen* n,
                          const u64* n_ub, u32 peers, void* const* d_out, u64* d_counts, u64* d_cursors,
                          u64* d_send_by_peer) {
  if (peers > MAX_PEERS || k > MZ_MAX_EXCHANGE) {
    MZ_SET_ERR(ctx, "exchange: %u peers / %u buffers exceed the supported maximum", peers, k);
    return MZGPU_E_UNSUPPORTED;
  }
  MZ_CUDA(ctx, cudaMemsetAsync(d_counts, 0, (size_t)k * 64 * 8, ctx->stream));
  PartJobs jobs;
  memset(&jobs, 0, sizeof(jobs));
  u64 max_ub = 0;
  for (u32 j = 0; j < k; ++j) {
    if (row_bytes[j] != 32 && row_bytes[j] != 80) {
      MZ_SET_ERR(ctx, "exchange: unsupported row width %d", row_bytes[j]);
      return MZGPU_E_UNSUPPORTED;
    }
    jobs.rows[j] = (const u64*)d_rows[j];
    jobs.out[j] = (u64*)d_out[j];
    jobs.n[j] = n[j];
    jobs.nw[j] = row_bytes[j] / 8;
    max_ub = n_ub[j] > max_ub ? n_ub[j] : max_ub;
  }
  u64 blocks = (max_ub + XT - 1) / XT;
  const u64 maxb = (u64)ctx->num_sms * 4;
  if (blocks > maxb) blocks = maxb;
  if (blocks == 0) blocks = 1;
  dim3 grid((unsigned)blocks, k);
  unsigned long long* c = (unsigned long long*)d_counts;
  unsigned long long* cur = (unsigned long long*)d_cursors;
  MZ_LAUNCH(ctx, k_part_count_many, grid, XT, 0, jobs, peers, c);
  MZ_LAUNCH(ctx, k_part_offsets_many, 1, 32, 0, c, peers, k, cur, (unsigned long long*)d_send_by_peer);
  MZ_LAUNCH(ctx, k_part_scatter_many, grid, XT, 0, jobs, peers, cur);
  return MZGPU_OK;
}
