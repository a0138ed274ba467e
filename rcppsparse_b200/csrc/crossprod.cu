// crossprod(): dense A^T A (reference inst/include/RcppSparse.h:158-194).
//
// The reference merges the sorted row lists of every PAIR of columns — n(n+1)/2 two-pointer merges, OpenMP over
// the first column.  On the device the sum is turned inside out: A^T A = sum over rows r of (row r)^T (row r),
// and the row-ordered copy of the mirror (capi.cu, build_row_companion) has every row as a contiguous list of
// (column id ascending, value).  A warp takes a row, stages it in shared memory, and for every pair a <= b of its
// entries fires one FP64 reduction at L2 into res(c_a, c_b) — sum_r k_r^2 / 2 reductions in all, nothing per
// column pair.  Only the upper triangle is accumulated; a tiled mirror pass then copies it below the diagonal,
// so the result is exactly symmetric like the reference's (which assigns res(col2, col1) = res(col1, col2)).
// The order of the additions inside one entry is not the reference's ascending-row order: results agree within
// the 1e-12 * sum|terms| bar, not bitwise.
#include "common.cuh"
#include "ptx.cuh"

namespace sb200 {
namespace {

constexpr int CP_THREADS = 256;
constexpr int CP_WARPS = CP_THREADS / 32;
constexpr int CP_CAP = 480;  // entries of a row staged per warp (45 KB per CTA); longer rows are walked in pieces

__global__ void __launch_bounds__(CP_THREADS) crossprod_rows_kernel(const int32_t* __restrict__ pt, const int32_t* __restrict__ ct,
                                                                     const double* __restrict__ xt, int32_t nrow, int64_t n,
                                                                     double* __restrict__ res) {
  __shared__ int32_t s_c[CP_WARPS][CP_CAP];
  __shared__ double s_x[CP_WARPS][CP_CAP];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t warps_total = static_cast<int64_t>(gridDim.x) * CP_WARPS;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * CP_WARPS + warp; r < nrow; r += warps_total) {
    const int32_t k0 = __ldg(pt + r), k1 = __ldg(pt + r + 1);
    // pieces of the row: (piece A, piece B >= A); inside a diagonal piece only pairs a <= b
    for (int32_t a0 = k0; a0 < k1; a0 += CP_CAP) {
      const int32_t an = min(CP_CAP, k1 - a0);
      __syncwarp();
      for (int32_t t = lane; t < an; t += 32) {
        s_c[warp][t] = __ldg(ct + a0 + t);
        s_x[warp][t] = __ldg(xt + a0 + t);
      }
      __syncwarp();
      // diagonal piece
      for (int32_t a = 0; a < an; ++a) {
        const int64_t ca = s_c[warp][a];
        const double xa = s_x[warp][a];
        for (int32_t b = a + lane; b < an; b += 32)
          ptx::red_add_f64(res + ca + static_cast<int64_t>(s_c[warp][b]) * n, __dmul_rn(xa, s_x[warp][b]));
      }
      // pieces to the right (rows longer than CP_CAP entries): b from global memory
      for (int32_t b = a0 + an + lane; b < k1; b += 32) {
        const int64_t cb = __ldg(ct + b);
        const double xb = __ldg(xt + b);
        for (int32_t a = 0; a < an; ++a) ptx::red_add_f64(res + s_c[warp][a] + cb * n, __dmul_rn(s_x[warp][a], xb));
      }
    }
  }
}

// res(j, i) = res(i, j) for i < j, 32 x 32 tiles through shared memory (both sides coalesced)
__global__ void __launch_bounds__(256) mirror_upper_kernel(double* __restrict__ res, int64_t n) {
  __shared__ double tile[32][33];
  const int64_t bi = blockIdx.x, bj = blockIdx.y;
  if (bi > bj) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int c = ty; c < 32; c += 8) {
    const int64_t row = bi * 32 + tx, col = bj * 32 + c;
    tile[c][tx] = (row < n && col < n) ? res[row + col * n] : 0.0;  // element (row, col) of the upper part
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int64_t row = bi * 32 + r, col = bj * 32 + tx;  // write (col, row) = below the diagonal when row < col
    if (row < n && col < n && row < col) res[col + row * n] = tile[tx][r];
  }
}

}  // namespace

// T = the row-ordered copy of the mirror (its "columns" are the rows of A, its i the column ids of A)
int launch_crossprod(const sb200_matrix* T, int32_t ncol_a, double* d_res, cudaStream_t st) {
  const int64_t n = ncol_a;
  if (n == 0) return SB200_OK;
  SB_CUDA(cudaMemsetAsync(d_res, 0, sizeof(double) * static_cast<size_t>(n) * static_cast<size_t>(n), st));
  if (T->nnz > 0 && T->ncol > 0) {
    int64_t blocks = (static_cast<int64_t>(T->ncol) + CP_WARPS - 1) / CP_WARPS;
    const int64_t cap = static_cast<int64_t>(T->sm_count) * 8;
    if (blocks > cap) blocks = cap;
    crossprod_rows_kernel<<<static_cast<unsigned>(blocks), CP_THREADS, 0, st>>>(T->d_p, T->d_i, T->d_x, T->ncol, n, d_res);
    count_launch();
    SB_CUDA(cudaGetLastError());
  }
  const unsigned tiles = static_cast<unsigned>((n + 31) / 32);
  if (tiles > 65535) return fail(SB200_E_UNSUPPORTED, "crossprod: more than 2M columns");
  mirror_upper_kernel<<<dim3(tiles, tiles), 256, 0, st>>>(d_res, n);
  count_launch();
  SB_CUDA(cudaGetLastError());
  return SB200_OK;
}

}  // namespace sb200
