// Single-pass exclusive prefix sum with decoupled look-back (self-written; no CUB/Thrust).
//
// Used by the transpose pipeline (row counts -> new column pointers p', the "exclusive scan"
// step of the counting sort that R's Matrix::t performs for reference RcppSparse.h:381-383) and
// by the synthetic generator (column lengths -> p).
//
// Each CTA takes the next tile id from a ticket (so every predecessor tile is already running
// or done — no deadlock whatever the hardware's block order), scans its 4096 counts locally,
// publishes its aggregate in a 64-bit status word {state:2, value:62}, then a warp walks the
// predecessors' status words backwards, 32 at a time, until it meets an inclusive prefix.
// One read and one write of the data: 8 B per element, HBM/L2-bound and tiny next to the sweeps.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace sb200 {

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_IPT = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_IPT;

constexpr unsigned long long ST_AGG = 1ull << 62;
constexpr unsigned long long ST_PREFIX = 2ull << 62;
constexpr unsigned long long ST_VALUE_MASK = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(SCAN_THREADS)
    scan_lookback_kernel(const uint32_t* __restrict__ in, int32_t* __restrict__ out, int64_t n,
                         unsigned long long* __restrict__ status, unsigned int* __restrict__ ticket,
                         unsigned long long* __restrict__ total_out) {
  __shared__ unsigned int s_tile;
  __shared__ unsigned long long s_warp[SCAN_THREADS / 32];
  __shared__ unsigned long long s_prefix;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(ticket, 1u);
  __syncthreads();
  const int64_t tile = s_tile;
  const int64_t base = tile * SCAN_TILE + static_cast<int64_t>(tid) * SCAN_IPT;

  uint32_t v[SCAN_IPT];
  unsigned long long tsum = 0;
#pragma unroll
  for (int k = 0; k < SCAN_IPT; ++k) {
    v[k] = (base + k < n) ? in[base + k] : 0u;
    tsum += v[k];
  }
  // block-exclusive scan of the per-thread sums
  unsigned long long inc = tsum;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const unsigned long long up = __shfl_up_sync(0xffffffffu, inc, off);
    if (lane >= off) inc += up;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  unsigned long long warp_base = 0, block_total = 0;
#pragma unroll
  for (int w = 0; w < SCAN_THREADS / 32; ++w) {
    if (w < warp) warp_base += s_warp[w];
    block_total += s_warp[w];
  }
  const unsigned long long thread_excl = warp_base + inc - tsum;

  // ---- decoupled look-back (warp 0) ---------------------------------------------------------------
  if (warp == 0) {
    unsigned long long exclusive = 0;
    if (tile == 0) {
      if (lane == 0) st_status(status, ST_PREFIX | block_total);
    } else {
      if (lane == 0) st_status(status + tile, ST_AGG | block_total);
      int64_t look = tile - 1;  // lane L inspects tile look - L
      while (true) {
        const int64_t mine = look - lane;
        unsigned long long w = ST_PREFIX;  // tiles before 0 behave as an empty inclusive prefix
        if (mine >= 0) {
          do {
            w = ld_status(status + mine);
          } while ((w >> 62) == 0ull);
        }
        const unsigned is_prefix = __ballot_sync(0xffffffffu, (w >> 62) == 2ull);
        // lanes up to and including the nearest prefix contribute
        const int first = is_prefix ? (__ffs(is_prefix) - 1) : 32;
        unsigned long long contrib = (lane <= first) ? (w & ST_VALUE_MASK) : 0ull;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) contrib += __shfl_down_sync(0xffffffffu, contrib, off);
        exclusive += __shfl_sync(0xffffffffu, contrib, 0);
        if (is_prefix) break;
        look -= 32;
      }
      if (lane == 0) st_status(status + tile, ST_PREFIX | (exclusive + block_total));
    }
    if (lane == 0) s_prefix = exclusive;
  }
  __syncthreads();
  const unsigned long long run0 = s_prefix + thread_excl;
  unsigned long long run = run0;
#pragma unroll
  for (int k = 0; k < SCAN_IPT; ++k) {
    if (base + k < n) out[base + k] = static_cast<int32_t>(run);
    run += v[k];
  }
  // the tile holding element n-1 also writes the grand total into out[n]
  const int64_t last_tile = (n - 1) / SCAN_TILE;
  if (tile == last_tile && tid == SCAN_THREADS - 1) {
    const unsigned long long grand = s_prefix + block_total;
    out[n] = static_cast<int32_t>(grand);
    if (total_out) *total_out = grand;
  }
}

__global__ void scan_empty_kernel(int32_t* out, unsigned long long* total_out) {
  out[0] = 0;
  if (total_out) *total_out = 0ull;
}

}  // namespace

size_t scan_workspace_bytes(int64_t n) {
  const int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  return 16 + sizeof(unsigned long long) * static_cast<size_t>(tiles > 0 ? tiles : 1);
}

int exclusive_scan_u32(cudaStream_t s, const uint32_t* d_in, int32_t* d_out, int64_t n, unsigned long long* d_total,
                       void* d_ws, size_t ws_bytes) {
  if (n < 0) return fail(SB200_E_INVALID, "scan: negative length");
  if (n == 0) {
    scan_empty_kernel<<<1, 1, 0, s>>>(d_out, d_total);
    count_launch();
    SB_CUDA(cudaGetLastError());
    return SB200_OK;
  }
  const size_t need = scan_workspace_bytes(n);
  if (ws_bytes < need) return fail(SB200_E_INVALID, "scan: workspace too small");
  const int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  SB_CUDA(cudaMemsetAsync(d_ws, 0, need, s));
  unsigned int* ticket = static_cast<unsigned int*>(d_ws);
  unsigned long long* status = reinterpret_cast<unsigned long long*>(static_cast<unsigned char*>(d_ws) + 16);
  scan_lookback_kernel<<<static_cast<unsigned>(tiles), SCAN_THREADS, 0, s>>>(d_in, d_out, n, status, ticket, d_total);
  count_launch();
  SB_CUDA(cudaGetLastError());
  return SB200_OK;
}

}  // namespace sb200
