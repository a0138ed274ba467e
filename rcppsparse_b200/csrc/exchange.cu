// Cross-GPU exchange for the column-sharded sweeps: result vectors live in a window of
// device memory that every rank of the node maps (cudaIpc), and the two exchange steps of the path are
// this library's own kernels over NVLink peer memory instead of library collectives:
//
//   * column-indexed results (colSums, colMeans, A^T v): each rank's sweep writes its slice straight into its
//     own window; xg_push_kernel copies that slice into the same place of every peer's window (P2P stores);
//     one flag barrier says "every slice has landed".
//   * row-indexed results (rowSums, rowMeans, A v): each rank's sweep leaves a full-length partial in its
//     window; after a flag barrier, rank r sums rows [r*nrow/N, (r+1)*nrow/N) of all N partials with P2P
//     loads IN RANK ORDER (the result is the same on every rank, run to run), divides for the means, and
//     stores the finished rows into every rank's result buffer; a second barrier ends the op.
//
// The barrier is an epoch counter per peer in each window: rank r stores epoch e into slot r of every peer
// (st.release.sys after a system fence), then waits until all of its own slots have reached e
// (ld.acquire.sys).  Epochs only grow, every rank issues the same sequence of barriers (SPMD), so no slot is
// ever reset.  A wait that sees no signal for 10 minutes (SB200_EXCHANGE_TIMEOUT_S) raises the window's error word and TRAPS:
// the stream fails loudly; it never goes on with data that has not arrived.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "pushrows.cuh"

namespace sb200 {

constexpr int XG_MAX_RANKS = PUSH_MAX_RANKS;
constexpr size_t XG_HEADER_BYTES = 4096;  // epoch slots, error word, CTA counter, gate word (see XG_W_*); data starts at 4096
constexpr uint32_t XG_MAGIC = 0x5B2000E8u;

struct XgPeers {
  unsigned char* base[XG_MAX_RANKS];
};

}  // namespace sb200

struct sb200_exchange {
  uint32_t magic;
  int device, rank, world;
  size_t bytes;  // whole window, header included
  unsigned char* window;
  bool connected;
  bool local;  // peers are windows of this same process (sb200_sharded): plain device pointers, nothing to unmap
  sb200::XgPeers peers;
  uint32_t epoch;
};

namespace sb200 {
namespace {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// header words of a window (uint32 index): [0,16) epoch slots, 16 error word, 32 CTA counter, 33 go word
constexpr int XG_W_ERROR = 16, XG_W_COUNT = 32, XG_W_GO = 33;

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// A peer that never signals (its process died, or it skipped a collective call) must not let this rank go on with
// data that has not arrived: the wait has no early exit.  After `timeout_ns` (default 10 minutes,
// SB200_EXCHANGE_TIMEOUT_S; legitimate skew between ranks — one rank's host busy with I/O, a checker, a debugger —
// is seconds) the window's error word is raised and the kernel TRAPS: the stream and every later CUDA call of
// this process fail loudly instead of returning sums of partial data.
__device__ __forceinline__ void xg_give_up(const XgPeers& peers, int rank, uint32_t code) {
  reinterpret_cast<volatile uint32_t*>(peers.base[rank])[XG_W_ERROR] = code;
  __threadfence_system();
  __trap();
}

// Signal epoch to every peer and wait for every peer's signal; called by threads q < world of ONE CTA.
// Everything this rank's earlier work wrote (locally or into peers) is ordered before the signal by the system
// fence; everything after the wait sees what the peers wrote before their signals.
__device__ __forceinline__ void xg_signal_and_wait(const XgPeers& peers, int rank, int q, uint32_t epoch,
                                                   unsigned long long timeout_ns) {
  __threadfence_system();
  st_release_sys(reinterpret_cast<uint32_t*>(peers.base[q]) + rank, epoch);
  const uint32_t* mine = reinterpret_cast<const uint32_t*>(peers.base[rank]) + q;
  const unsigned long long t0 = global_ns();
  unsigned spins = 0;
  while (static_cast<int32_t>(ld_acquire_sys(mine) - epoch) < 0) {
    if ((++spins & 1023u) == 0u && global_ns() - t0 > timeout_ns) xg_give_up(peers, rank, 0x80000000u | static_cast<uint32_t>(q));
    __nanosleep(64);
  }
}

__global__ void xg_barrier_kernel(XgPeers peers, int rank, int world, uint32_t epoch, unsigned long long timeout_ns) {
  if (threadIdx.x < world) xg_signal_and_wait(peers, rank, threadIdx.x, epoch, timeout_ns);
}

// Tail of a multi-CTA exchange kernel: the last CTA to get here runs the barrier, so the kernel (and with it
// the stream) completes only when every rank's data has landed.
__device__ __forceinline__ void xg_tail_barrier(const XgPeers& peers, int rank, int world, uint32_t epoch,
                                                unsigned long long timeout_ns) {
  __shared__ int is_last;
  __threadfence_system();  // my stores (peer windows included) before the count
  __syncthreads();
  uint32_t* hdr = reinterpret_cast<uint32_t*>(peers.base[rank]);
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(hdr + XG_W_COUNT, 1u);
    is_last = (t == gridDim.x - 1);
    if (is_last) hdr[XG_W_COUNT] = 0;  // ready for the next launch (stream-ordered)
  }
  __syncthreads();
  if (is_last && threadIdx.x < world) xg_signal_and_wait(peers, rank, threadIdx.x, epoch, timeout_ns);
}

// my slice of a column-indexed result -> the same offset in every peer's window, then the barrier
__global__ void __launch_bounds__(256) xg_push_kernel(XgPeers peers, int rank, int world, size_t off_bytes, int64_t n,
                                                       uint32_t epoch, unsigned long long timeout_ns) {
  const double* __restrict__ src = reinterpret_cast<const double*>(peers.base[rank] + off_bytes);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; k < n; k += stride) {
    const double v = src[k];
    for (int q = 0; q < world; ++q)
      if (q != rank) reinterpret_cast<double*>(peers.base[q] + off_bytes)[k] = v;
  }
  xg_tail_barrier(peers, rank, world, epoch, timeout_ns);
}

// Sharded transpose, the exchange step: every rank has transposed its own column block; output row r lives on the rank
// that owns r's row block and is the concatenation, in rank (= column) order, of every rank's segment of row r.  This
// kernel writes MY segment of every row straight to its final place in the owner's window (P2P stores): a warp takes 32
// consecutive rows and walks their concatenated entries 32 at a time (coalesced reads; the writes are contiguous per
// segment).  dst_off[r] = where my segment of row r starts in the owner's output (entries); owner of r = the q with
// rb[q] <= r < rb[q+1]; column ids leave as GLOBAL ids (local id + col_offset).  The barrier at the end = all landed.
struct XgPushRows {
  PushRowsSrc src;
  size_t cols_off[XG_MAX_RANKS];  // byte offset of the int32 output region inside peer q's window
  size_t vals_off[XG_MAX_RANKS];  // byte offset of the double output region
};

struct XgWindowDest {
  const XgPeers& peers;
  const XgPushRows& a;
  __device__ __forceinline__ int32_t* cols(int q) const { return reinterpret_cast<int32_t*>(peers.base[q] + a.cols_off[q]); }
  __device__ __forceinline__ double* vals(int q) const { return reinterpret_cast<double*>(peers.base[q] + a.vals_off[q]); }
};

__global__ void __launch_bounds__(256) xg_push_rows_kernel(XgPeers peers, int rank, int world, XgPushRows a, uint32_t epoch,
                                                            unsigned long long timeout_ns) {
  push_rows_body(a.src, world, XgWindowDest{peers, a});
  xg_tail_barrier(peers, rank, world, epoch, timeout_ns);
}

// The whole row exchange in one launch.  CTA 0 runs barrier `epoch` (every rank's partial is complete) and
// opens the gate for the other CTAs of this launch; rows [r0, r1) of the result = sum over ranks, in rank
// order, of their partials (/ divisor), stored into every rank's window; the last CTA runs barrier epoch+1.
// r0 is even, so pairs of rows move as 16-byte accesses — one load per peer in flight per thread.
template <int WORLD>
__global__ void __launch_bounds__(256) xg_reduce_kernel(XgPeers peers, int rank, int world, size_t partial_off,
                                                         size_t result_off, int64_t r0, int64_t r1, double divisor,
                                                         uint32_t epoch, unsigned long long timeout_ns) {
  uint32_t* hdr = reinterpret_cast<uint32_t*>(peers.base[rank]);
  if (blockIdx.x == 0) {
    if (threadIdx.x < world) xg_signal_and_wait(peers, rank, threadIdx.x, epoch, timeout_ns);
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(hdr + XG_W_GO), "r"(epoch) : "memory");
  } else {
    if (threadIdx.x == 0) {
      // CTA 0 of this launch opens the gate once every rank's partial is complete.  It is resident (launched
      // first) or becomes resident as soon as a slot frees up; like the barrier itself the wait has no early exit —
      // it gives up by trapping, never by going on with partials that may be incomplete.
      uint32_t v;
      const unsigned long long t0 = global_ns();
      unsigned spins = 0;
      for (;;) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(hdr + XG_W_GO) : "memory");
        if (static_cast<int32_t>(v - epoch) >= 0) break;
        if ((++spins & 255u) == 0u && global_ns() - t0 > timeout_ns + timeout_ns / 4) xg_give_up(peers, rank, 0xC0000000u);
        __nanosleep(200);  // a sweep may be running beside this kernel: do not hammer the L2 slice of the gate word
      }
    }
    __syncthreads();
  }
  const int nw = WORLD > 0 ? WORLD : world;
  const int64_t pairs = (r1 - r0) >> 1;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  auto reduce_pair = [&](int64_t k, double2& acc) {
    acc = *reinterpret_cast<const double2*>(peers.base[0] + partial_off + 8 * k);
    for (int q = 1; q < nw; ++q) {
      const double2 t = *reinterpret_cast<const double2*>(peers.base[q] + partial_off + 8 * k);
      acc.x = __dadd_rn(acc.x, t.x);
      acc.y = __dadd_rn(acc.y, t.y);
    }
    if (divisor != 0.0) {
      acc.x = __ddiv_rn(acc.x, divisor);
      acc.y = __ddiv_rn(acc.y, divisor);
    }
  };
  int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (WORLD > 0) {
    // two pairs of rows per step: 2 x WORLD independent 16-byte loads in flight per thread
    for (; j + stride < pairs; j += 2 * stride) {
      const int64_t ka = r0 + 2 * j, kb = r0 + 2 * (j + stride);
      double2 pa[WORLD > 0 ? WORLD : 1], pb[WORLD > 0 ? WORLD : 1];
#pragma unroll
      for (int q = 0; q < WORLD; ++q) {
        pa[q] = *reinterpret_cast<const double2*>(peers.base[q] + partial_off + 8 * ka);
        pb[q] = *reinterpret_cast<const double2*>(peers.base[q] + partial_off + 8 * kb);
      }
      double2 a = pa[0], b = pb[0];
#pragma unroll
      for (int q = 1; q < WORLD; ++q) {
        a.x = __dadd_rn(a.x, pa[q].x);
        a.y = __dadd_rn(a.y, pa[q].y);
        b.x = __dadd_rn(b.x, pb[q].x);
        b.y = __dadd_rn(b.y, pb[q].y);
      }
      if (divisor != 0.0) {
        a.x = __ddiv_rn(a.x, divisor);
        a.y = __ddiv_rn(a.y, divisor);
        b.x = __ddiv_rn(b.x, divisor);
        b.y = __ddiv_rn(b.y, divisor);
      }
#pragma unroll
      for (int q = 0; q < WORLD; ++q) {
        *reinterpret_cast<double2*>(peers.base[q] + result_off + 8 * ka) = a;
        *reinterpret_cast<double2*>(peers.base[q] + result_off + 8 * kb) = b;
      }
    }
  }
  for (; j < pairs; j += stride) {
    const int64_t k = r0 + 2 * j;
    double2 acc;
    reduce_pair(k, acc);
    for (int q = 0; q < nw; ++q) *reinterpret_cast<double2*>(peers.base[q] + result_off + 8 * k) = acc;
  }
  if (((r1 - r0) & 1) && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {  // odd tail row
    const int64_t k = r1 - 1;
    double acc = reinterpret_cast<const double*>(peers.base[0] + partial_off)[k];
    for (int q = 1; q < nw; ++q) acc = __dadd_rn(acc, reinterpret_cast<const double*>(peers.base[q] + partial_off)[k]);
    if (divisor != 0.0) acc = __ddiv_rn(acc, divisor);
    for (int q = 0; q < nw; ++q) reinterpret_cast<double*>(peers.base[q] + result_off)[k] = acc;
  }
  xg_tail_barrier(peers, rank, world, epoch + 1, timeout_ns);
}

int check_xg(const sb200_exchange* x, bool need_connected) {
  if (!x || x->magic != XG_MAGIC) return fail(SB200_E_INVALID, "not a live sb200_exchange handle");
  if (need_connected && !x->connected) return fail(SB200_E_INVALID, "exchange window is not connected to its peers yet");
  return SB200_OK;
}

int check_range(const sb200_exchange* x, size_t off, int64_t n, const char* what) {
  if (n < 0 || (off & 15) || off < XG_HEADER_BYTES || off + static_cast<size_t>(n) * 8 > x->bytes)
    return fail(SB200_E_INVALID, std::string(what) + ": range outside the exchange window (or not 16-byte aligned)");
  return SB200_OK;
}

// The exchange kernels run beside a sweep whose CTAs need the SM's largest shared-memory carveout; a kernel that
// asks for (almost) no shared memory would otherwise pull the SMs it lands on to a small carveout, and the
// sweep's CTAs could not be placed there until it left (measured: the sweep beside an exchange was slower by
// the exchange's whole duration).  Ask for the same carveout.
template <typename K>
void prefer_max_shared(K kernel) {
  cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaGetLastError();
}
void configure_kernels() {
  static bool done = false;
  if (done) return;
  done = true;
  prefer_max_shared(xg_barrier_kernel);
  prefer_max_shared(xg_push_kernel);
  prefer_max_shared(xg_reduce_kernel<0>);
  prefer_max_shared(xg_reduce_kernel<2>);
  prefer_max_shared(xg_reduce_kernel<4>);
  prefer_max_shared(xg_reduce_kernel<8>);
}

unsigned long long exchange_timeout_ns() {
  static const unsigned long long ns = [] {
    double sec = 600.0;
    if (const char* e = getenv("SB200_EXCHANGE_TIMEOUT_S")) {
      const double v = atof(e);
      if (v > 0.0) sec = v;
    }
    return static_cast<unsigned long long>(sec * 1e9);
  }();
  return ns;
}

// CTAs of an exchange kernel: they run beside the next op's sweep, so they must leave it its SM slots.  Measured at
// N = 2 (profiles/r02, C2 step of four sweeps): 296 CTAs 0.581 ms/step (the sweep beside a reduction takes 0.155 instead
// of 0.134 ms), 148: 0.552, 74: 0.548, 32: 0.548 — half a CTA per SM is enough to move 8 MB over NVLink inside one
// sweep.  SB200_XG_CTAS overrides.
int exchange_max_ctas() {
  static const int n = [] {
    int v = 74;
    if (const char* e = getenv("SB200_XG_CTAS")) {
      const int w = atoi(e);
      if (w >= 1 && w <= 1184) v = w;
    }
    return v;
  }();
  return n;
}

int launch_barrier(sb200_exchange* x, cudaStream_t st) {
  x->epoch += 1;
  xg_barrier_kernel<<<1, 32, 0, st>>>(x->peers, x->rank, x->world, x->epoch, exchange_timeout_ns());
  count_launch();
  SB_CUDA(cudaGetLastError());
  return SB200_OK;
}

}  // namespace
}  // namespace sb200

using namespace sb200;

extern "C" {

int sb200_exchange_create(int device, int64_t data_bytes, sb200_exchange** out, void* ipc_handle_out /* 64 bytes */) {
  if (!out || !ipc_handle_out || data_bytes < 0) return fail(SB200_E_INVALID, "sb200_exchange_create: bad argument");
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) {
    cudaGetLastError();
    return fail(SB200_E_NODEVICE, "sb200_exchange_create: CUDA device not present");
  }
  SB_CUDA(cudaSetDevice(device));
  sb200_exchange* x = new (std::nothrow) sb200_exchange();
  if (!x) return fail(SB200_E_NOMEM, "host allocation failed");
  memset(x, 0, sizeof(*x));
  x->magic = XG_MAGIC;
  x->device = device;
  x->rank = -1;
  x->bytes = XG_HEADER_BYTES + ((static_cast<size_t>(data_bytes) + 255) & ~static_cast<size_t>(255));
  // plain cudaMalloc: stream-ordered pool memory cannot be exported with cudaIpcGetMemHandle
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&x->window), x->bytes);
  if (e == cudaSuccess) e = cudaMemset(x->window, 0, x->bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, x->window);
  if (e != cudaSuccess) {
    if (x->window) cudaFree(x->window);
    delete x;
    return cuda_fail(e, "exchange window (cudaMalloc / cudaIpcGetMemHandle)", __FILE__, __LINE__);
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
  memcpy(ipc_handle_out, &h, 64);
  *out = x;
  return SB200_OK;
}

int sb200_exchange_connect(sb200_exchange* x, int rank, int world, const void* all_handles /* world x 64 bytes */) {
  SB_TRY(check_xg(x, false));
  if (x->connected) return fail(SB200_E_INVALID, "exchange window already connected");
  if (!all_handles || world < 1 || world > XG_MAX_RANKS || rank < 0 || rank >= world)
    return fail(SB200_E_INVALID, "sb200_exchange_connect: bad rank/world (at most 16 ranks of one node)");
  SB_CUDA(cudaSetDevice(x->device));
  for (int q = 0; q < world; ++q) {
    if (q == rank) {
      x->peers.base[q] = x->window;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const unsigned char*>(all_handles) + 64 * static_cast<size_t>(q), 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      for (int k = 0; k < q; ++k)
        if (k != rank && x->peers.base[k]) cudaIpcCloseMemHandle(x->peers.base[k]), x->peers.base[k] = nullptr;
      return cuda_fail(e, "cudaIpcOpenMemHandle (peer window; the ranks must be GPUs of one node with P2P access)", __FILE__, __LINE__);
    }
    x->peers.base[q] = static_cast<unsigned char*>(p);
  }
  configure_kernels();
  x->rank = rank;
  x->world = world;
  x->connected = true;
  return SB200_OK;
}

int sb200_exchange_destroy(sb200_exchange* x) {
  if (!x) return SB200_OK;
  SB_TRY(check_xg(x, false));
  cudaSetDevice(x->device);
  cudaDeviceSynchronize();
  if (x->connected && !x->local)
    for (int q = 0; q < x->world; ++q)
      if (q != x->rank && x->peers.base[q]) cudaIpcCloseMemHandle(x->peers.base[q]);
  cudaFree(x->window);
  cudaGetLastError();
  x->magic = 0;
  delete x;
  return SB200_OK;
}

}  // extern "C"

namespace sb200 {
// Single-process form (sharded.cu): the windows of xs[0..world) belong to this process, one per device; every device
// gets peer access to every other and the windows are addressed directly (unified addressing), no cudaIpc.
int exchange_connect_local(sb200_exchange** xs, int world) {
  if (world < 1 || world > XG_MAX_RANKS) return fail(SB200_E_INVALID, "exchange: at most 16 devices");
  for (int r = 0; r < world; ++r) SB_TRY(check_xg(xs[r], false));
  for (int r = 0; r < world; ++r) {
    SB_CUDA(cudaSetDevice(xs[r]->device));
    for (int q = 0; q < world; ++q) {
      if (q == r || xs[q]->device == xs[r]->device) continue;
      int can = 0;
      SB_CUDA(cudaDeviceCanAccessPeer(&can, xs[r]->device, xs[q]->device));
      if (!can) return fail(SB200_E_UNSUPPORTED, "exchange: devices " + std::to_string(xs[r]->device) + " and " +
                                                     std::to_string(xs[q]->device) + " have no peer access");
      const cudaError_t e = cudaDeviceEnablePeerAccess(xs[q]->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
      cudaGetLastError();
    }
  }
  configure_kernels();
  for (int r = 0; r < world; ++r) {
    for (int q = 0; q < world; ++q) xs[r]->peers.base[q] = xs[q]->window;
    xs[r]->rank = r;
    xs[r]->world = world;
    xs[r]->connected = true;
    xs[r]->local = true;
  }
  return SB200_OK;
}
unsigned char* exchange_window_base(const sb200_exchange* x) { return x->window; }
}  // namespace sb200

extern "C" {

int sb200_exchange_window(const sb200_exchange* x, void** base, int64_t* data_offset, int64_t* bytes) {
  SB_TRY(check_xg(x, false));
  if (base) *base = x->window;
  if (data_offset) *data_offset = static_cast<int64_t>(XG_HEADER_BYTES);
  if (bytes) *bytes = static_cast<int64_t>(x->bytes);
  return SB200_OK;
}

int sb200_exchange_barrier(sb200_exchange* x, void* cuda_stream) {
  SB_TRY(check_xg(x, true));
  SB_CUDA(cudaSetDevice(x->device));
  return launch_barrier(x, static_cast<cudaStream_t>(cuda_stream));
}

int sb200_exchange_gather(sb200_exchange* x, void* cuda_stream, int64_t full_offset, int64_t slice_begin, int64_t slice_len) {
  SB_TRY(check_xg(x, true));
  SB_CUDA(cudaSetDevice(x->device));
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  if (full_offset < 0 || slice_begin < 0) return fail(SB200_E_INVALID, "sb200_exchange_gather: negative offset");
  const size_t off = static_cast<size_t>(full_offset) + 8 * static_cast<size_t>(slice_begin);
  if (slice_len > 0) {
    if (off & 7) return fail(SB200_E_INVALID, "sb200_exchange_gather: misaligned slice");
    if (off < XG_HEADER_BYTES || off + 8 * static_cast<size_t>(slice_len) > x->bytes)
      return fail(SB200_E_INVALID, "sb200_exchange_gather: slice outside the exchange window");
  }
  if (x->world > 1) {
    int64_t blocks = (slice_len + 255) / 256;
    if (blocks > exchange_max_ctas()) blocks = exchange_max_ctas();
    if (blocks < 1) blocks = 1;  // an empty slice still takes part in the barrier
    x->epoch += 1;
    xg_push_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(x->peers, x->rank, x->world, off, slice_len, x->epoch, exchange_timeout_ns());
    count_launch();
    SB_CUDA(cudaGetLastError());
  }
  return SB200_OK;
}

int sb200_exchange_reduce(sb200_exchange* x, void* cuda_stream, int64_t partial_offset, int64_t result_offset, int64_t n,
                          double divisor) {
  SB_TRY(check_xg(x, true));
  SB_CUDA(cudaSetDevice(x->device));
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  if (partial_offset < 0 || result_offset < 0) return fail(SB200_E_INVALID, "sb200_exchange_reduce: negative offset");
  SB_TRY(check_range(x, static_cast<size_t>(partial_offset), n, "sb200_exchange_reduce partial"));
  SB_TRY(check_range(x, static_cast<size_t>(result_offset), n, "sb200_exchange_reduce result"));
  // row blocks start on even rows (16-byte accesses); the last one runs to n
  auto cut = [&](int q) { return q >= x->world ? n : ((n * q / x->world) & ~static_cast<int64_t>(1)); };
  const int64_t r0 = cut(x->rank), r1 = cut(x->rank + 1);
  int64_t blocks = ((r1 - r0) / 4 + 255) / 256;
  if (blocks > exchange_max_ctas()) blocks = exchange_max_ctas();  // at most two CTAs per SM beside a running sweep
  if (blocks < 1) blocks = 1;
  const unsigned g = static_cast<unsigned>(blocks);
  const size_t po = static_cast<size_t>(partial_offset), ro = static_cast<size_t>(result_offset);
  const uint32_t e = x->epoch + 1;
  x->epoch += 2;
  const unsigned long long tmo = exchange_timeout_ns();
  switch (x->world) {
    case 2: xg_reduce_kernel<2><<<g, 256, 0, st>>>(x->peers, x->rank, 2, po, ro, r0, r1, divisor, e, tmo); break;
    case 4: xg_reduce_kernel<4><<<g, 256, 0, st>>>(x->peers, x->rank, 4, po, ro, r0, r1, divisor, e, tmo); break;
    case 8: xg_reduce_kernel<8><<<g, 256, 0, st>>>(x->peers, x->rank, 8, po, ro, r0, r1, divisor, e, tmo); break;
    default: xg_reduce_kernel<0><<<g, 256, 0, st>>>(x->peers, x->rank, x->world, po, ro, r0, r1, divisor, e, tmo); break;
  }
  count_launch();
  SB_CUDA(cudaGetLastError());
  return SB200_OK;
}

int sb200_exchange_push_rows(sb200_exchange* x, void* cuda_stream, const int32_t* d_p_loc, const int32_t* d_cols,
                             const double* d_vals, const int64_t* d_dst_off, int32_t nrow, int32_t col_offset,
                             const int32_t* row_bounds, const int64_t* cols_offsets, const int64_t* vals_offsets) {
  SB_TRY(check_xg(x, true));
  SB_CUDA(cudaSetDevice(x->device));
  if (nrow < 0 || !row_bounds || !cols_offsets || !vals_offsets || (nrow > 0 && (!d_p_loc || !d_dst_off)))
    return fail(SB200_E_INVALID, "sb200_exchange_push_rows: bad argument");
  XgPushRows a;
  a.src.p_loc = d_p_loc;
  a.src.cols = d_cols;
  a.src.vals = d_vals;
  a.src.dst_off = d_dst_off;
  a.src.nrow = nrow;
  a.src.col_offset = col_offset;
  for (int q = 0; q <= x->world; ++q) a.src.rb[q] = row_bounds[q];
  for (int q = 0; q < x->world; ++q) {
    if (cols_offsets[q] < static_cast<int64_t>(XG_HEADER_BYTES) || vals_offsets[q] < static_cast<int64_t>(XG_HEADER_BYTES) ||
        (cols_offsets[q] & 3) || (vals_offsets[q] & 7))
      return fail(SB200_E_INVALID, "sb200_exchange_push_rows: output regions must lie behind the window header, aligned");
    a.cols_off[q] = static_cast<size_t>(cols_offsets[q]);
    a.vals_off[q] = static_cast<size_t>(vals_offsets[q]);
  }
  int64_t blocks = ((static_cast<int64_t>(nrow) + 31) / 32 + 7) / 8;
  if (blocks > 592) blocks = 592;
  if (blocks < 1) blocks = 1;
  x->epoch += 1;
  xg_push_rows_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(x->peers, x->rank, x->world, a,
                                                                                                        x->epoch, exchange_timeout_ns());
  count_launch();
  SB_CUDA(cudaGetLastError());
  return SB200_OK;
}

int sb200_exchange_status(sb200_exchange* x) {
  SB_TRY(check_xg(x, false));
  SB_CUDA(cudaSetDevice(x->device));
  uint32_t err = 0;
  const cudaError_t ce = cudaMemcpy(&err, x->window + 4 * XG_W_ERROR, sizeof(err), cudaMemcpyDeviceToHost);
  if (ce != cudaSuccess)  // a barrier that gave up traps its kernel: the context reports it from then on
    return cuda_fail(ce, "exchange status (an exchange barrier may have given up waiting for a peer and trapped)", __FILE__, __LINE__);
  if (err != 0)
    return fail(SB200_E_CUDA, "exchange barrier timed out waiting for rank " + std::to_string(err & 0xffffu) +
                                  " (a peer process died or skipped a collective call)");
  return SB200_OK;
}

}  // extern "C"
