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
__global__ void __launch_bounds__(XT) k_part_count(const u64* __restrict__ rows, u64 n, u32 peers,
                                                   unsigned long long* __restrict__ counts) {
  __shared__ u32 sh[MAX_PEERS];
  if (threadIdx.x < MAX_PEERS) sh[threadIdx.x] = 0;
  __syncthreads();
  for (u64 i = (u64)blockIdx.x * XT + threadIdx.x; i < n; i += (u64)gridDim.x * XT)
    atomicAdd(&sh[(u32)(fnv1a64(rows[i * NW]) % peers)], 1u);
  __syncthreads();
  if (threadIdx.x < peers && sh[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

template <int NW>
__global__ void __launch_bounds__(XT) k_part_scatter(const u64* __restrict__ rows, u64 n, u32 peers,
                                                     unsigned long long* __restrict__ cursors,
                                                     u64* __restrict__ out) {
  __shared__ u32 sh_count[MAX_PEERS];
  __shared__ u64 sh_base[MAX_PEERS];
  if (threadIdx.x < MAX_PEERS) sh_count[threadIdx.x] = 0;
  __syncthreads();
  const u64 i = (u64)blockIdx.x * XT + threadIdx.x;
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

}  // namespace

uint32_t mzgpu_route(uint64_t key, uint32_t peers) { return (uint32_t)(fnv1a64(key) % peers); }

int32_t mz_partition(mzgpu_ctx* ctx, int row_bytes, const void* d_rows, u64 n, u32 peers, void* d_out,
                     u64* h_counts) {
  for (u32 p = 0; p < peers; ++p) h_counts[p] = 0;
  if (peers > MAX_PEERS) {
    MZ_SET_ERR(ctx, "exchange: %u peers exceed the supported maximum %d", peers, MAX_PEERS);
    return MZGPU_E_UNSUPPORTED;
  }
  if (n == 0) return MZGPU_OK;
  DevMem counts;
  MZ_TRY(counts.alloc(ctx, MAX_PEERS * 8));
  MZ_CUDA(ctx, cudaMemsetAsync(counts.p, 0, MAX_PEERS * 8, ctx->stream));
  const u64* r = (const u64*)d_rows;
  unsigned long long* c = counts.as<unsigned long long>();
  u64 blocks = (n + XT - 1) / XT;
  unsigned cgrid = (unsigned)(blocks < (u64)ctx->num_sms * 8 ? blocks : (u64)ctx->num_sms * 8);
  switch (row_bytes) {
    case 32: MZ_LAUNCH(ctx, k_part_count<4>, cgrid, XT, 0, r, n, peers, c); break;
    case 80: MZ_LAUNCH(ctx, k_part_count<10>, cgrid, XT, 0, r, n, peers, c); break;
    default: MZ_SET_ERR(ctx, "exchange: unsupported row width %d", row_bytes); return MZGPU_E_UNSUPPORTED;
  }
  MZ_CUDA(ctx, cudaMemcpyAsync(ctx->h_scratch + 32, counts.p, peers * 8, cudaMemcpyDeviceToHost,
                               ctx->stream));
  MZ_SYNC(ctx);
  ctx->stats.d2h_bytes += peers * 8;
  u64 off = 0;
  u64 offsets[MAX_PEERS];
  for (u32 p = 0; p < peers; ++p) {
    h_counts[p] = ctx->h_scratch[32 + p];
    offsets[p] = off;
    off += h_counts[p];
  }
  MZ_CUDA(ctx, cudaMemcpyAsync(counts.p, offsets, peers * 8, cudaMemcpyHostToDevice, ctx->stream));
  switch (row_bytes) {
    case 32: MZ_LAUNCH(ctx, k_part_scatter<4>, (unsigned)blocks, XT, 0, r, n, peers, c, (u64*)d_out); break;
    case 80: MZ_LAUNCH(ctx, k_part_scatter<10>, (unsigned)blocks, XT, 0, r, n, peers, c, (u64*)d_out); break;
    default: return MZGPU_E_UNSUPPORTED;
  }
  // `offsets` is a stack array: make sure the H2D copy has consumed it
  MZ_SYNC(ctx);
  return MZGPU_OK;
}
