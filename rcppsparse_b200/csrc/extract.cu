// the reference's range cursors and dense extraction as batched device ops.
//
//   sb200_col_sums_in_rows   for EVERY column at once, the sum a user loop gets from
//                            InnerIteratorInRange(A, col, s) / InnerIteratorNotInRange(A, col, s)
//                            (reference RcppSparse.h:238-264, 270-321): entries whose row is / is not in the index set s.
//                            = the A^T v sweep with an indicator operand and SKIP semantics (a skipped Inf must not
//                            become Inf * 0 = NaN): sweep_kernel<SPMV_T> in mask mode.
//   sb200_gather_block       A(rows, cols) as a dense column-major block (reference :76-92 operator()(IntegerVector,
//                            IntegerVector), :95-107 col(...), :110-128 row(...)): a thread per output element,
//                            binary search in the column's sorted rows instead of the reference's linear scan per element.
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "common.cuh"

namespace sb200 {
int launch_masked_col_sums(sb200_matrix* m, const double* d_mask, double* d_out);  // sweep.cu

namespace {

__global__ void mask_fill_kernel(double* __restrict__ mask, int64_t n, double v) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; k < n; k += stride) mask[k] = v;
}
// indices outside [0, nrow) select nothing (the reference's cursor would never meet them)
__global__ void mask_set_kernel(double* __restrict__ mask, int32_t nrow, const int32_t* __restrict__ rows, int64_t n, double v) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; k < n; k += stride) {
    const int32_t r = rows[k];
    if (r >= 0 && r < nrow) mask[r] = v;
  }
}

// out[jc * nr + ir] = A(rows[ir], cols[jc]); rows / cols null = all rows / all columns
__global__ void gather_block_kernel(const int32_t* __restrict__ gi, const int32_t* __restrict__ gp, const double* __restrict__ gx,
                                    int32_t nrow, int32_t ncol, const int32_t* __restrict__ rows, int64_t nr,
                                    const int32_t* __restrict__ cols, int64_t nc, double* __restrict__ out) {
  const int64_t total = nr * nc;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int64_t jc = e / nr, ir = e - jc * nr;
    const int64_t c = cols ? cols[jc] : jc;
    const int64_t r = rows ? rows[ir] : ir;
    double v = 0.0;
    if (c >= 0 && c < ncol && r >= 0 && r < nrow) {
      int32_t lo = gp[c], hi = gp[c + 1];
      while (lo < hi) {
        const int32_t mid = lo + ((hi - lo) >> 1);
        if (gi[mid] < r)
          lo = mid + 1;
        else
          hi = mid;
      }
      if (lo < gp[c + 1] && gi[lo] == r) v = gx[lo];
    }
    out[e] = v;
  }
}

int blocks_for(const sb200_matrix* m, int64_t n) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = static_cast<int64_t>(m->sm_count) * 16;
  if (b > cap) b = cap;
  return static_cast<int>(b < 1 ? 1 : b);
}

}  // namespace
}  // namespace sb200

using namespace sb200;

extern "C" {

int sb200_col_sums_in_rows(sb200_matrix* m, const int32_t* rows, int64_t n, int negate, double* out) {
  SB_TRY(check_handle(m));
  DeviceGuard guard(m->device);
  if (!guard.ok) return fail(SB200_E_CUDA, "cudaSetDevice failed");
  SB_TRY(ensure_rows(m));  // a mirror created with SB200_LAZY_ROWS gets its row indices now
  if (n < 0 || (n > 0 && !rows)) return fail(SB200_E_INVALID, "sb200_col_sums_in_rows: bad index set");
  if (m->ncol > 0 && !out) return fail(SB200_E_INVALID, "output buffer is NULL");
  cudaStream_t st = m->stream;
  double* mask = m->d_stage_in;  // nrow entries fit: the staging vectors hold max(nrow, ncol)
  if (m->nrow > 0) {
    mask_fill_kernel<<<blocks_for(m, m->nrow), 256, 0, st>>>(mask, m->nrow, negate ? 1.0 : 0.0);
    count_launch();
    SB_CUDA(cudaGetLastError());
  }
  int32_t* d_rows = nullptr;
  if (n > 0 && m->nrow > 0) {
    SB_TRY(pool_alloc(reinterpret_cast<void**>(&d_rows), sizeof(int32_t) * static_cast<size_t>(n), st));
    cudaError_t e = cudaMemcpyAsync(d_rows, rows, sizeof(int32_t) * static_cast<size_t>(n), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
      mask_set_kernel<<<blocks_for(m, n), 256, 0, st>>>(mask, m->nrow, d_rows, n, negate ? 0.0 : 1.0);
      count_launch();
      e = cudaGetLastError();
    }
    pool_free(d_rows, st);
    if (e != cudaSuccess) return cuda_fail(e, "sb200_col_sums_in_rows: index set", __FILE__, __LINE__);
  }
  if (m->ncol > 0) SB_CUDA(cudaMemsetAsync(m->d_stage_out, 0, sizeof(double) * static_cast<size_t>(m->ncol), st));
  if (m->nnz > 0 && m->nrow > 0) SB_TRY(launch_masked_col_sums(m, mask, m->d_stage_out));
  if (m->ncol > 0)
    SB_CUDA(cudaMemcpyAsync(out, m->d_stage_out, sizeof(double) * static_cast<size_t>(m->ncol), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  return SB200_OK;
}

int sb200_gather_block(sb200_matrix* m, const int32_t* rows, int64_t nr, const int32_t* cols, int64_t nc, double* out) {
  SB_TRY(check_handle(m));
  DeviceGuard guard(m->device);
  if (!guard.ok) return fail(SB200_E_CUDA, "cudaSetDevice failed");
  SB_TRY(ensure_rows(m));  // a mirror created with SB200_LAZY_ROWS gets its row indices now
  if (!rows) nr = m->nrow;
  if (!cols) nc = m->ncol;
  if (nr < 0 || nc < 0) return fail(SB200_E_INVALID, "sb200_gather_block: negative size");
  const int64_t total = nr * nc;
  if (total == 0) return SB200_OK;
  if (!out) return fail(SB200_E_INVALID, "output buffer is NULL");
  if (total > (1LL << 31)) return fail(SB200_E_UNSUPPORTED, "sb200_gather_block: block of more than 2^31 elements");
  cudaStream_t st = m->stream;
  double* d_out = nullptr;
  int32_t *d_rows = nullptr, *d_cols = nullptr;
  int rc = pool_alloc(reinterpret_cast<void**>(&d_out), sizeof(double) * static_cast<size_t>(total), st);
  if (rc == SB200_OK && rows) rc = pool_alloc(reinterpret_cast<void**>(&d_rows), sizeof(int32_t) * static_cast<size_t>(nr), st);
  if (rc == SB200_OK && cols) rc = pool_alloc(reinterpret_cast<void**>(&d_cols), sizeof(int32_t) * static_cast<size_t>(nc), st);
  cudaError_t e = cudaSuccess;
  if (rc == SB200_OK) {
    if (rows) e = cudaMemcpyAsync(d_rows, rows, sizeof(int32_t) * static_cast<size_t>(nr), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && cols) e = cudaMemcpyAsync(d_cols, cols, sizeof(int32_t) * static_cast<size_t>(nc), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
      gather_block_kernel<<<blocks_for(m, total), 256, 0, st>>>(m->d_i, m->d_p, m->d_x, m->nrow, m->ncol, d_rows, nr, d_cols, nc, d_out);
      count_launch();
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, sizeof(double) * static_cast<size_t>(total), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  }
  pool_free(d_out, st);
  pool_free(d_rows, st);
  pool_free(d_cols, st);
  if (rc != SB200_OK) return rc;
  if (e != cudaSuccess) return cuda_fail(e, "sb200_gather_block", __FILE__, __LINE__);
  return SB200_OK;
}

}  // extern "C"
