// Structure validation at upload time.
//
// The reference validates nothing but slot presence (RcppSparse.h:34-41,407-415); a corrupt row
// index surfaces later as Rcpp::index_out_of_bounds from the bounds-checked sums(i[j]) at
// RcppSparse.h:142, or as silent undefined behaviour through the unchecked [] accessors.  On the
// device an out-of-range index would be a wild atomic into HBM, so the mirror checks the
// dgCMatrix invariants once, when it is created: p[0]=0, p non-decreasing, p[ncol]=nnz,
// 0 <= i < nrow, and i strictly ascending inside every column (at() relies on it,
// RcppSparse.h:67-68; the transpose's duplicate-free ranking relies on it too).
// Cost: one read of i and p (4 B/nnz) — inside an upload that is PCIe-bound anyway.
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "common.cuh"

namespace sb200 {

namespace {

enum : unsigned {
  BAD_P0 = 1u,
  BAD_PN = 2u,
  BAD_P_ORDER = 4u,
  BAD_I_RANGE = 8u,
  BAD_I_ORDER = 16u,
};

__global__ void validate_p_kernel(const int32_t* __restrict__ p, int64_t ncol, int64_t nnz, unsigned* __restrict__ err) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  unsigned bad = 0;
  for (int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; c <= ncol; c += stride) {
    const int32_t v = p[c];
    if (c == 0 && v != 0) bad |= BAD_P0;
    if (c == ncol && v != nnz) bad |= BAD_PN;
    if (c < ncol && p[c + 1] < v) bad |= BAD_P_ORDER;
    if (v < 0 || v > nnz) bad |= BAD_P_ORDER;
  }
  if (bad) atomicOr(err, bad);
}

__global__ void validate_i_kernel(const int32_t* __restrict__ i, const int32_t* __restrict__ p, int64_t ncol,
                                  int64_t nnz, int32_t nrow, unsigned* __restrict__ err) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  unsigned bad = 0;
  for (int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; k < nnz; k += stride) {
    const int32_t r = i[k];
    if (r < 0 || r >= nrow) bad |= BAD_I_RANGE;
    if (k + 1 < nnz && i[k + 1] <= r) {
      // a descent is legal only where a new column starts: k+1 must be one of the p values
      int64_t lo = 0, hi = ncol + 1;
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (p[mid] < k + 1)
          lo = mid + 1;
        else
          hi = mid;
      }
      if (lo > ncol || p[lo] != k + 1) bad |= BAD_I_ORDER;
    }
  }
  if (bad) atomicOr(err, bad);
}

}  // namespace

namespace {

int structure_error(unsigned h_err) {
  std::string msg = "dgCMatrix structure invalid:";
  if (h_err & BAD_P0) msg += " p[0] != 0;";
  if (h_err & BAD_PN) msg += " p[ncol] != nnz;";
  if (h_err & BAD_P_ORDER) msg += " p not non-decreasing within [0, nnz];";
  if (h_err & BAD_I_RANGE) msg += " row index outside [0, nrow) (reference: Rcpp::index_out_of_bounds at RcppSparse.h:142);";
  if (h_err & BAD_I_ORDER) msg += " row indices not strictly ascending inside a column;";
  return fail(SB200_E_STRUCTURE, msg);
}

}  // namespace

// i alone; p has been checked (binary searches over it are meaningful)
int validate_rows(sb200_matrix* m) {
  if (m->nnz <= 0) return SB200_OK;
  unsigned* d_err = static_cast<unsigned*>(m->d_ws);  // first word of the workspace; re-zeroed below
  SB_CUDA(cudaMemsetAsync(d_err, 0, sizeof(unsigned), m->stream));
  int64_t blocks = (m->nnz + 255) / 256;
  if (blocks > m->sm_count * 16) blocks = m->sm_count * 16;
  validate_i_kernel<<<static_cast<unsigned>(blocks), 256, 0, m->stream>>>(m->d_i, m->d_p, m->ncol, m->nnz, m->nrow, d_err);
  count_launch();
  SB_CUDA(cudaGetLastError());
  unsigned h_err = 0;
  SB_CUDA(cudaMemcpyAsync(&h_err, d_err, sizeof(unsigned), cudaMemcpyDeviceToHost, m->stream));
  SB_CUDA(cudaStreamSynchronize(m->stream));
  SB_CUDA(cudaMemsetAsync(d_err, 0, sizeof(unsigned), m->stream));  // the sweep's ticket lives here
  return h_err ? structure_error(h_err) : SB200_OK;
}

int validate_structure(sb200_matrix* m, bool rows) {
  unsigned* d_err = static_cast<unsigned*>(m->d_ws);
  SB_CUDA(cudaMemsetAsync(d_err, 0, sizeof(unsigned), m->stream));
  {
    int64_t blocks = (static_cast<int64_t>(m->ncol) + 1 + 255) / 256;
    if (blocks > m->sm_count * 8) blocks = m->sm_count * 8;
    validate_p_kernel<<<static_cast<unsigned>(blocks), 256, 0, m->stream>>>(m->d_p, m->ncol, m->nnz, d_err);
    count_launch();
    SB_CUDA(cudaGetLastError());
  }
  unsigned h_err = 0;
  SB_CUDA(cudaMemcpyAsync(&h_err, d_err, sizeof(unsigned), cudaMemcpyDeviceToHost, m->stream));
  SB_CUDA(cudaStreamSynchronize(m->stream));
  SB_CUDA(cudaMemsetAsync(d_err, 0, sizeof(unsigned), m->stream));
  if (h_err != 0) return structure_error(h_err);
  return rows ? validate_rows(m) : SB200_OK;
}

}  // namespace sb200
