// One process, several GPUs: the column-sharding layer behind the C ABI.
//
// The reference is one R process (zdebruine/RcppSparse; every method of RcppSparse::Matrix, RcppSparse.h:131-156,
// runs on the calling thread), so a drop-in that wants more than one GPU cannot ask for one process per device.
// sb200_sharded_create cuts the dgCMatrix into nnz-balanced contiguous column blocks (binary search in p; a block
// is itself a dgCMatrix with the full row count and p rebased to 0), uploads block k to devices[k] as an ordinary
// mirror (worker thread per device: each device has its own PCIe link), and gives every device a window of
// peer-mapped memory.  The ops then run as the same kernels as on one GPU:
//   column-indexed results (colSums, colMeans, A^T v)  disjoint slices: every device sweeps its block and its slice
//                                                      goes straight to its place in the caller's host vector;
//   row-indexed results (rowSums, rowMeans, A v)       full-length partials in the windows, summed in rank order by
//                                                      the library's own P2P reduction kernel (exchange.cu) running
//                                                      on every device at once; the mean's division by the GLOBAL
//                                                      ncol rides in that kernel (RcppSparse.h:154).
// All launches come from the calling thread, device after device, nothing blocks until the final synchronisation.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"
#include "pushrows.cuh"

struct sb200_sharded {
  uint32_t magic;
  int world;
  int32_t nrow, ncol;
  int64_t nnz;
  std::vector<int> devices;
  std::vector<int64_t> bounds;  // [world + 1] first column of every block
  std::vector<sb200_matrix*> blocks;
  std::vector<sb200_exchange*> windows;
  int64_t partial_off, result_off;  // byte offsets of the row partial / the row result inside every window
};

namespace sb200 {
namespace {

constexpr uint32_t SHARDED_MAGIC = 0x5B2005A4u;

int check_sharded(const sb200_sharded* s) {
  if (!s || s->magic != SHARDED_MAGIC) return fail(SB200_E_INVALID, "not a live sb200_sharded handle");
  return SB200_OK;
}

struct WallTrace {  // SB200_TRACE=1: host wall clock of the phases of an entry point on stderr
  bool on = getenv("SB200_TRACE") != nullptr;
  const char* title;
  std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
  explicit WallTrace(const char* what) : title(what) {}
  void phase(const char* what) {
    if (!on) return;
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[sb200 trace] %s: %s %.2f ms\n", title, what, std::chrono::duration<double, std::milli>(now - t).count());
    t = now;
  }
};

void destroy_sharded(sb200_sharded* s) {
  if (!s) return;
  WallTrace tr("sharded destroy");
  for (sb200_matrix* m : s->blocks)
    if (m) sb200_matrix_destroy(m);
  tr.phase("blocks");
  for (sb200_exchange* x : s->windows)
    if (x) sb200_exchange_destroy(x);
  tr.phase("windows");
  s->magic = 0;
  delete s;
}

// column-indexed op: out[bounds[k] .. bounds[k+1]) from device k
int run_columns(sb200_sharded* s, SweepMode mode, const double* v_host, double divisor, double* out) {
  if (s->ncol > 0 && !out) return fail(SB200_E_INVALID, "output buffer is NULL");
  DeviceGuard restore(s->blocks[0]->device);  // the caller's current device is put back on return
  int rc = SB200_OK;
  for (int k = 0; k < s->world && rc == SB200_OK; ++k) {
    sb200_matrix* m = s->blocks[k];
    DeviceGuard guard(m->device);
    if (!guard.ok) return fail(SB200_E_CUDA, "cudaSetDevice failed");
    if (mode == SWEEP_SPMV_T && s->nrow > 0)
      SB_CUDA(cudaMemcpyAsync(m->d_stage_in, v_host, sizeof(double) * static_cast<size_t>(s->nrow), cudaMemcpyHostToDevice, m->stream));
    rc = launch_sweep(m, mode, m->d_stage_in, divisor, m->d_stage_out);
    if (rc == SB200_OK && m->ncol > 0)
      SB_CUDA(cudaMemcpyAsync(out + s->bounds[k], m->d_stage_out, sizeof(double) * static_cast<size_t>(m->ncol), cudaMemcpyDeviceToHost,
                              m->stream));
  }
  for (int k = 0; k < s->world; ++k) {  // every device is waited for, whatever happened on another
    DeviceGuard guard(s->blocks[k]->device);
    const cudaError_t e = cudaStreamSynchronize(s->blocks[k]->stream);
    if (e != cudaSuccess && rc == SB200_OK) rc = cuda_fail(e, "sharded column sweep", __FILE__, __LINE__);
  }
  return rc;
}

// row-indexed op: partials in the windows, rank-ordered P2P reduction on every device, result read from device 0
int run_rows(sb200_sharded* s, SweepMode mode, const double* v_host, double divisor, double* out) {
  if (s->nrow > 0 && !out) return fail(SB200_E_INVALID, "output buffer is NULL");
  DeviceGuard restore(s->blocks[0]->device);  // the exchange entry points switch devices; put the caller's back on return
  int rc = SB200_OK;
  for (int k = 0; k < s->world && rc == SB200_OK; ++k) {
    sb200_matrix* m = s->blocks[k];
    DeviceGuard guard(m->device);
    if (!guard.ok) return fail(SB200_E_CUDA, "cudaSetDevice failed");
    double* partial = reinterpret_cast<double*>(exchange_window_base(s->windows[k]) + s->partial_off);
    if (mode == SWEEP_SPMV && m->ncol > 0)
      SB_CUDA(cudaMemcpyAsync(m->d_stage_in, v_host + s->bounds[k], sizeof(double) * static_cast<size_t>(m->ncol), cudaMemcpyHostToDevice,
                              m->stream));
    rc = launch_sweep(m, mode, m->d_stage_in, 0.0, partial);
  }
  if (rc == SB200_OK) {
    if (s->world > 1) {
      for (int k = 0; k < s->world && rc == SB200_OK; ++k)
        rc = sb200_exchange_reduce(s->windows[k], s->blocks[k]->stream, s->partial_off, s->result_off, s->nrow, divisor);
    } else if (divisor != 0.0 && s->nrow > 0) {
      DeviceGuard guard(s->blocks[0]->device);
      rc = launch_vec_div(s->blocks[0]->stream, reinterpret_cast<double*>(exchange_window_base(s->windows[0]) + s->partial_off), s->nrow,
                          divisor);
    }
  }
  if (rc == SB200_OK && s->nrow > 0) {
    DeviceGuard guard(s->blocks[0]->device);
    const int64_t off = s->world > 1 ? s->result_off : s->partial_off;
    SB_CUDA(cudaMemcpyAsync(out, exchange_window_base(s->windows[0]) + off, sizeof(double) * static_cast<size_t>(s->nrow),
                            cudaMemcpyDeviceToHost, s->blocks[0]->stream));
  }
  for (int k = 0; k < s->world; ++k) {
    DeviceGuard guard(s->blocks[k]->device);
    const cudaError_t e = cudaStreamSynchronize(s->blocks[k]->stream);
    if (e != cudaSuccess && rc == SB200_OK) rc = cuda_fail(e, "sharded row sweep", __FILE__, __LINE__);
  }
  return rc;
}


// ---- sharded transpose: local transposes, row segments pushed to the devices that own the rows -------
struct TpPtrs {
  const int32_t* p[PUSH_MAX_RANKS];  // row pointers of the devices' local transposes (peer-readable)
};
struct PtrDest {
  int32_t* c[PUSH_MAX_RANKS];
  double* v[PUSH_MAX_RANKS];
  __device__ __forceinline__ int32_t* cols(int q) const { return c[q]; }
  __device__ __forceinline__ double* vals(int q) const { return v[q]; }
};

// cnt[r] = entries of row r over all column blocks
__global__ void shard_row_counts_kernel(const TpPtrs tp, int world, int32_t nrow, uint32_t* __restrict__ cnt) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; r < nrow; r += stride) {
    uint32_t n = 0;
    for (int k = 0; k < world; ++k) n += static_cast<uint32_t>(tp.p[k][r + 1] - tp.p[k][r]);
    cnt[r] = n;
  }
}

struct RowBounds {
  int32_t rb[PUSH_MAX_RANKS + 1];
};

// dst[k * nrow + r] = where block k's piece of row r starts inside the arrays of the row's owner: the row's place in
// the result (P[r], relative to the owner's first row) plus the pieces of the blocks before k — block order = column order
__global__ void shard_row_dst_kernel(const TpPtrs tp, int world, int32_t nrow, const int32_t* __restrict__ P, const RowBounds b,
                                     int64_t* __restrict__ dst) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; r < nrow; r += stride) {
    int q = 0;
    while (q + 1 < world && b.rb[q + 1] <= r) ++q;
    int64_t run = static_cast<int64_t>(P[r]) - P[b.rb[q]];
    for (int k = 0; k < world; ++k) {
      dst[static_cast<int64_t>(k) * nrow + r] = run;
      run += tp.p[k][r + 1] - tp.p[k][r];
    }
  }
}

__global__ void __launch_bounds__(256) push_rows_local_kernel(const PushRowsSrc a, int world, const PtrDest d) {
  push_rows_body(a, world, d);
}

}  // namespace
}  // namespace sb200

using namespace sb200;

extern "C" {

int sb200_sharded_create(const int32_t* i, const int32_t* p, const double* x, int32_t nrow, int32_t ncol, int64_t nnz, int n_gpus,
                         const int* devices, unsigned flags, sb200_sharded** out) {
  if (!out) return fail(SB200_E_INVALID, "out is NULL");
  *out = nullptr;
  if (!p || (nnz > 0 && (!i || !x))) return fail(SB200_E_INVALID, "NULL slot array");
  if (nrow < 0 || ncol < 0 || nnz < 0) return fail(SB200_E_INVALID, "negative dimension");
  if (n_gpus < 1 || n_gpus > 16) return fail(SB200_E_INVALID, "n_gpus must be 1..16 (the GPUs of one node)");
  int have = 0;
  if (cudaGetDeviceCount(&have) != cudaSuccess || have <= 0) {
    cudaGetLastError();
    return fail(SB200_E_NODEVICE, "no CUDA device available: libsparse_b200 has no CPU fallback");
  }
  // the host arrays are validated on the device block by block; the split itself needs p[0] = 0, p monotone, p[ncol] = nnz
  if (p[0] != 0 || p[ncol] != nnz) return fail(SB200_E_STRUCTURE, "p[0] must be 0 and p[ncol] must be nnz");
  sb200_sharded* s = new (std::nothrow) sb200_sharded();
  if (!s) return fail(SB200_E_NOMEM, "host allocation failed");
  s->magic = SHARDED_MAGIC;
  s->world = n_gpus;
  s->nrow = nrow;
  s->ncol = ncol;
  s->nnz = nnz;
  s->devices.resize(n_gpus);
  for (int k = 0; k < n_gpus; ++k) {
    s->devices[k] = devices ? devices[k] : k;
    if (s->devices[k] < 0 || s->devices[k] >= have) {
      destroy_sharded(s);
      return fail(SB200_E_NODEVICE, "requested CUDA device " + std::to_string(devices ? devices[k] : k) + " not present");
    }
  }
  // nnz-balanced contiguous column blocks: first column whose start offset reaches k * nnz / n
  s->bounds.assign(n_gpus + 1, 0);
  for (int k = 1; k < n_gpus; ++k) {
    const int64_t target = (nnz * k) / n_gpus;
    int64_t lo = s->bounds[k - 1], hi = ncol;
    while (lo < hi) {
      const int64_t mid = lo + ((hi - lo) >> 1);
      if (p[mid] < target)
        lo = mid + 1;
      else
        hi = mid;
    }
    s->bounds[k] = lo;
  }
  s->bounds[n_gpus] = ncol;
  s->blocks.assign(n_gpus, nullptr);
  s->windows.assign(n_gpus, nullptr);
  // upload the blocks concurrently: one worker per device
  WallTrace tr("sharded create");
  std::vector<int> rcs(n_gpus, SB200_OK);
  std::vector<std::string> errs(n_gpus);
  std::vector<std::thread> workers;
  for (int k = 0; k < n_gpus; ++k) {
    workers.emplace_back([&, k] {
      const int64_t c0 = s->bounds[k], c1 = s->bounds[k + 1];
      const int64_t k0 = p[c0], k1 = p[c1];
      std::vector<int32_t> pk(static_cast<size_t>(c1 - c0) + 1);
      bool ok = true;
      for (int64_t c = c0; c <= c1; ++c) {
        pk[c - c0] = static_cast<int32_t>(p[c] - k0);
        if (c > c0 && p[c] < p[c - 1]) ok = false;
      }
      if (!ok) {
        rcs[k] = SB200_E_STRUCTURE;
        errs[k] = "p is not monotone";
        return;
      }
      rcs[k] = sb200_matrix_create(i + k0, pk.data(), x + k0, nrow, static_cast<int32_t>(c1 - c0), k1 - k0, s->devices[k], flags & ~SB200_LAZY_ROWS,
                                   &s->blocks[k]);
      if (rcs[k] != SB200_OK) errs[k] = sb200_last_error();
    });
  }
  for (auto& w : workers) w.join();
  tr.phase("blocks uploaded");
  for (int k = 0; k < n_gpus; ++k)
    if (rcs[k] != SB200_OK) {
      const int rc = rcs[k];
      const std::string msg = "column block " + std::to_string(k) + " (device " + std::to_string(s->devices[k]) + "): " + errs[k];
      destroy_sharded(s);
      return fail(rc, msg);
    }
  // windows: [row partial | row result], 256-byte aligned, behind the exchange header
  const int64_t vec = ((static_cast<int64_t>(nrow > 0 ? nrow : 1) * 8 + 255) / 256) * 256;
  int rc = SB200_OK;
  int prev = 0;
  cudaGetDevice(&prev);
  for (int k = 0; k < n_gpus && rc == SB200_OK; ++k) {
    unsigned char handle[64];
    rc = sb200_exchange_create(s->devices[k], 2 * vec, &s->windows[k], handle);
  }
  tr.phase("windows allocated");
  if (rc == SB200_OK) rc = exchange_connect_local(s->windows.data(), n_gpus);
  cudaSetDevice(prev);
  tr.phase("peers connected");
  if (rc != SB200_OK) {
    const std::string msg = sb200_last_error();
    destroy_sharded(s);
    return fail(rc, msg);
  }
  void* base = nullptr;
  int64_t data_off = 0, bytes = 0;
  sb200_exchange_window(s->windows[0], &base, &data_off, &bytes);
  s->partial_off = data_off;
  s->result_off = data_off + vec;
  *out = s;
  return SB200_OK;
}

int sb200_sharded_destroy(sb200_sharded* s) {
  if (!s) return SB200_OK;
  SB_TRY(check_sharded(s));
  int prev = 0;
  cudaGetDevice(&prev);
  destroy_sharded(s);
  cudaSetDevice(prev);
  return SB200_OK;
}

int sb200_sharded_info(const sb200_sharded* s, int* n_gpus, int64_t* bounds) {
  SB_TRY(check_sharded(s));
  if (n_gpus) *n_gpus = s->world;
  if (bounds)
    for (int k = 0; k <= s->world; ++k) bounds[k] = s->bounds[k];
  return SB200_OK;
}

int sb200_sharded_block(const sb200_sharded* s, int k, sb200_matrix** block) {
  SB_TRY(check_sharded(s));
  if (k < 0 || k >= s->world || !block) return fail(SB200_E_INVALID, "sb200_sharded_block: bad block index");
  *block = s->blocks[k];
  return SB200_OK;
}

int sb200_sharded_col_sums(sb200_sharded* s, double* out) {
  SB_TRY(check_sharded(s));
  return run_columns(s, SWEEP_COLSUM, nullptr, 0.0, out);
}
int sb200_sharded_col_means(sb200_sharded* s, double* out) {
  SB_TRY(check_sharded(s));
  if (s->nrow == 0) {  // RcppSparse.h:148: sums / Dim[0]; 0/0 = NaN there and here
    for (int32_t c = 0; c < s->ncol; ++c) out[c] = 0.0 / static_cast<double>(s->nrow);
    return SB200_OK;
  }
  return run_columns(s, SWEEP_COLSUM, nullptr, static_cast<double>(s->nrow), out);
}
int sb200_sharded_spmv_t(sb200_sharded* s, const double* v, double* y) {
  SB_TRY(check_sharded(s));
  if (s->nrow > 0 && !v) return fail(SB200_E_INVALID, "operand vector is NULL");
  return run_columns(s, SWEEP_SPMV_T, v, 0.0, y);
}
int sb200_sharded_row_sums(sb200_sharded* s, double* out) {
  SB_TRY(check_sharded(s));
  return run_rows(s, SWEEP_ROWSUM, nullptr, 0.0, out);
}
int sb200_sharded_row_means(sb200_sharded* s, double* out) {
  SB_TRY(check_sharded(s));
  if (s->ncol == 0) {
    for (int32_t r = 0; r < s->nrow; ++r) out[r] = 0.0 / static_cast<double>(s->ncol);
    return SB200_OK;
  }
  return run_rows(s, SWEEP_ROWSUM, nullptr, static_cast<double>(s->ncol), out);
}
int sb200_sharded_spmv(sb200_sharded* s, const double* v, double* y) {
  SB_TRY(check_sharded(s));
  if (s->ncol > 0 && !v) return fail(SB200_E_INVALID, "operand vector is NULL");
  return run_rows(s, SWEEP_SPMV, v, 0.0, y);
}


// CSC(A^T) of the whole matrix into host arrays (RcppSparse.h:375-385): every device transposes its column block, the
// row counts are summed and scanned on device 0 (p_out), rows are dealt to the devices in nnz-balanced contiguous
// ranges, every device pushes its piece of every row straight to its place in the owner's arrays over peer memory
// (block order = column order inside a row: bit-exact), and every owner copies its range of i_out / x_out to the host
// over its own PCIe link.
int sb200_sharded_transpose(sb200_sharded* s, int32_t* p_out, int32_t* i_out, double* x_out) {
  SB_TRY(check_sharded(s));
  if (!p_out || (s->nnz > 0 && (!i_out || !x_out))) return fail(SB200_E_INVALID, "NULL output array");
  const int W = s->world;
  const int32_t nrow = s->nrow;
  if (s->nnz == 0 || nrow == 0) {
    for (int64_t r = 0; r <= nrow; ++r) p_out[r] = 0;
    return SB200_OK;
  }
  DeviceGuard restore(s->blocks[0]->device);
  const bool trace = getenv("SB200_TRACE") != nullptr;
  auto t_prev = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!trace) return;
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[sb200 trace] sharded transpose: %s %.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
    t_prev = now;
  };
  // 1. local transposes, concurrently (the entry point blocks)
  std::vector<sb200_matrix*> T(W, nullptr);
  std::vector<int> rcs(W, SB200_OK);
  std::vector<std::string> errs(W);
  {
    std::vector<std::thread> workers;
    for (int k = 0; k < W; ++k)
      workers.emplace_back([&, k] {
        rcs[k] = sb200_transpose_dev(s->blocks[k], &T[k]);
        if (rcs[k] != SB200_OK) errs[k] = sb200_last_error();
      });
    for (auto& w : workers) w.join();
  }
  struct Scratch {  // everything below is released on every way out
    std::vector<sb200_matrix*>& T;
    sb200_sharded* s;
    uint32_t* cnt = nullptr;
    int32_t* P = nullptr;
    void* ws = nullptr;
    int64_t* dst = nullptr;
    int32_t* tpc = nullptr;          // device 0: copies of every block's row pointer
    std::vector<int64_t*> dstk;      // device k: its slice of dst
    std::vector<int32_t*> oc;        // the owners' arrays: plain cudaMalloc, visible to the peers (pool memory is not)
    std::vector<double*> ov;
    ~Scratch() {
      for (int k = 0; k < s->world; ++k) {
        DeviceGuard g(s->blocks[k]->device);
        cudaStreamSynchronize(s->blocks[k]->stream);
        if (k < static_cast<int>(oc.size()) && oc[k]) cudaFree(oc[k]);
        if (k < static_cast<int>(ov.size()) && ov[k]) cudaFree(ov[k]);
        if (k > 0 && k < static_cast<int>(dstk.size())) pool_free(dstk[k], s->blocks[k]->stream);
        if (k == 0) {
          pool_free(tpc, s->blocks[0]->stream);
          pool_free(cnt, s->blocks[0]->stream);
          pool_free(P, s->blocks[0]->stream);
          pool_free(ws, s->blocks[0]->stream);
          pool_free(dst, s->blocks[0]->stream);
        }
        if (T[k]) sb200_matrix_destroy(T[k]);
      }
      cudaGetLastError();
    }
  } sc{T, s};
  lap("local transposes");
  for (int k = 0; k < W; ++k)
    if (rcs[k] != SB200_OK) return fail(rcs[k], "sharded transpose, column block " + std::to_string(k) + ": " + errs[k]);
  // 2. row counts over the blocks -> p_out (device 0)
  sb200_matrix* m0 = s->blocks[0];
  cudaStream_t st0 = m0->stream;
  TpPtrs tp;
  const size_t ws_bytes = scan_workspace_bytes(nrow);
  {
    DeviceGuard g(m0->device);
    // the library's arrays are pool memory, which peers cannot address: bring the (small) row pointers to device 0
    const size_t pb = sizeof(int32_t) * (static_cast<size_t>(nrow) + 1);
    SB_TRY(pool_alloc(reinterpret_cast<void**>(&sc.tpc), pb * W, st0));
    for (int k = 0; k < PUSH_MAX_RANKS; ++k) tp.p[k] = k < W ? sc.tpc + static_cast<size_t>(k) * (nrow + 1) : nullptr;
    for (int k = 0; k < W; ++k)
      SB_CUDA(cudaMemcpyPeerAsync(sc.tpc + static_cast<size_t>(k) * (nrow + 1), m0->device, T[k]->d_p, s->blocks[k]->device, pb, st0));
    SB_TRY(pool_alloc(reinterpret_cast<void**>(&sc.cnt), sizeof(uint32_t) * static_cast<size_t>(nrow), st0));
    SB_TRY(pool_alloc(reinterpret_cast<void**>(&sc.P), sizeof(int32_t) * (static_cast<size_t>(nrow) + 1), st0));
    SB_TRY(pool_alloc(&sc.ws, ws_bytes, st0));
    SB_TRY(pool_alloc(reinterpret_cast<void**>(&sc.dst), sizeof(int64_t) * static_cast<size_t>(W) * nrow, st0));
    const unsigned grid = static_cast<unsigned>(std::min<int64_t>((static_cast<int64_t>(nrow) + 255) / 256, 4096));
    shard_row_counts_kernel<<<grid, 256, 0, st0>>>(tp, W, nrow, sc.cnt);
    count_launch();
    SB_CUDA(cudaGetLastError());
    SB_TRY(exclusive_scan_u32(st0, sc.cnt, sc.P, nrow, nullptr, sc.ws, ws_bytes));
    SB_CUDA(cudaMemcpyAsync(p_out, sc.P, sizeof(int32_t) * (static_cast<size_t>(nrow) + 1), cudaMemcpyDeviceToHost, st0));
    SB_CUDA(cudaStreamSynchronize(st0));
    lap("row counts, scan, p_out");
    if (p_out[nrow] != s->nnz) return fail(SB200_E_CUDA, "sharded transpose: the blocks' row counts do not add up to nnz");
    // 3. rows dealt to the devices: nnz-balanced contiguous ranges
    RowBounds rb;
    rb.rb[0] = 0;
    for (int q = 1; q < W; ++q) {
      const int64_t target = (s->nnz * q) / W;
      int32_t lo = rb.rb[q - 1], hi = nrow;
      while (lo < hi) {
        const int32_t mid = lo + ((hi - lo) >> 1);
        if (p_out[mid] < target)
          lo = mid + 1;
        else
          hi = mid;
      }
      rb.rb[q] = lo;
    }
    for (int q = W; q <= PUSH_MAX_RANKS; ++q) rb.rb[q] = nrow;
    shard_row_dst_kernel<<<grid, 256, 0, st0>>>(tp, W, nrow, sc.P, rb, sc.dst);
    count_launch();
    SB_CUDA(cudaGetLastError());
    SB_CUDA(cudaStreamSynchronize(st0));
    // every device gets its slice of dst
    sc.dstk.assign(W, nullptr);
    sc.dstk[0] = sc.dst;
    for (int k = 1; k < W; ++k) {
      DeviceGuard gk(s->blocks[k]->device);
      SB_TRY(pool_alloc(reinterpret_cast<void**>(&sc.dstk[k]), sizeof(int64_t) * static_cast<size_t>(nrow), s->blocks[k]->stream));
      SB_CUDA(cudaMemcpyPeerAsync(sc.dstk[k], s->blocks[k]->device, sc.dst + static_cast<int64_t>(k) * nrow, m0->device,
                                  sizeof(int64_t) * static_cast<size_t>(nrow), s->blocks[k]->stream));
    }
    // 4. the owners' arrays
    sc.oc.assign(W, nullptr);
    sc.ov.assign(W, nullptr);
    for (int q = 0; q < W; ++q) {
      DeviceGuard gq(s->blocks[q]->device);
      const size_t n = static_cast<size_t>(p_out[rb.rb[q + 1]] - p_out[rb.rb[q]]);
      SB_CUDA(cudaMalloc(reinterpret_cast<void**>(&sc.oc[q]), padded_bytes(sizeof(int32_t) * (n + 1))));
      SB_CUDA(cudaMalloc(reinterpret_cast<void**>(&sc.ov[q]), padded_bytes(sizeof(double) * (n + 1))));
    }
    // 5. every device pushes its pieces
    PtrDest pd;
    for (int q = 0; q < PUSH_MAX_RANKS; ++q) {
      pd.c[q] = q < W ? sc.oc[q] : nullptr;
      pd.v[q] = q < W ? sc.ov[q] : nullptr;
    }
    for (int k = 0; k < W; ++k) {
      DeviceGuard gk(s->blocks[k]->device);
      PushRowsSrc a;
      a.p_loc = T[k]->d_p;
      a.cols = T[k]->d_i;
      a.vals = T[k]->d_x;
      a.dst_off = sc.dstk[k];
      a.nrow = nrow;
      a.col_offset = static_cast<int32_t>(s->bounds[k]);
      for (int q = 0; q <= PUSH_MAX_RANKS; ++q) a.rb[q] = rb.rb[q];
      int64_t blocks = ((static_cast<int64_t>(nrow) + 31) / 32 + 7) / 8;
      if (blocks > 592) blocks = 592;
      push_rows_local_kernel<<<static_cast<unsigned>(blocks), 256, 0, s->blocks[k]->stream>>>(a, W, pd);
      count_launch();
      SB_CUDA(cudaGetLastError());
    }
    for (int k = 0; k < W; ++k) {
      DeviceGuard gk(s->blocks[k]->device);
      SB_CUDA(cudaStreamSynchronize(s->blocks[k]->stream));
    }
    lap("push");
    // 6. every owner brings its range home over its own link
    std::vector<int> drc(W, SB200_OK);
    std::vector<std::thread> workers;
    for (int q = 0; q < W; ++q)
      workers.emplace_back([&, q] {
        const int64_t o = p_out[rb.rb[q]];
        const size_t n = static_cast<size_t>(p_out[rb.rb[q + 1]] - o);
        if (n == 0) return;
        const int dev = s->blocks[q]->device;
        if (cudaSetDevice(dev) != cudaSuccess) {
          drc[q] = SB200_E_CUDA;
          return;
        }
        const size_t bi = sizeof(int32_t) * n, bx = sizeof(double) * n;
        if (bi >= STAGED_COPY_MIN_BYTES && host_is_pageable(i_out + o))
          drc[q] = staged_d2h(dev, i_out + o, sc.oc[q], bi);
        else if (cudaMemcpy(i_out + o, sc.oc[q], bi, cudaMemcpyDeviceToHost) != cudaSuccess)
          drc[q] = SB200_E_CUDA;
        if (drc[q] != SB200_OK) return;
        if (bx >= STAGED_COPY_MIN_BYTES && host_is_pageable(x_out + o))
          drc[q] = staged_d2h(dev, x_out + o, sc.ov[q], bx);
        else if (cudaMemcpy(x_out + o, sc.ov[q], bx, cudaMemcpyDeviceToHost) != cudaSuccess)
          drc[q] = SB200_E_CUDA;
      });
    for (auto& w : workers) w.join();
    lap("copies to the host");
    for (int q = 0; q < W; ++q)
      if (drc[q] != SB200_OK) return fail(drc[q], "sharded transpose: copying rows of device " + std::to_string(q) + " to the host failed");
  }
  return SB200_OK;
}

}  // extern "C"
