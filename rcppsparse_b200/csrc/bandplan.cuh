// Row-band plan shared by the banded kernels (bands.cu) and the transpose (transpose.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace sb200 {

struct BandView {
  const int32_t* i;
  const int32_t* p;
  const double* x;
  int32_t ncol;
  int nb;
  int S;
  const int32_t* rb;   // [nb+1]
  const int32_t* cs;   // [S+1]
  const int32_t* bpt;  // [(nb-1)*ncol]
};

__device__ __forceinline__ int32_t band_start(const BandView& a, int b, int64_t c) {
  if (b == 0) return __ldg(a.p + c);
  if (b == a.nb) return __ldg(a.p + c + 1);
  return __ldg(a.bpt + static_cast<int64_t>(b - 1) * a.ncol + c);
}

struct BandPlan {
  int nb = 0, S = 0, max_rows = 0;
  bool has_offsets = false;
  int kind = 0;                       // transpose plans: which placement kernel the geometry was cut for
  int env_bands = 0, env_splits = 0;  // the tuning overrides it was built under
  int32_t* d_rb = nullptr;
  int32_t* d_cs = nullptr;
  int32_t* d_bpt = nullptr;
  int32_t* d_rowptr = nullptr;  // [nrow+1]
  int32_t* d_off = nullptr;     // [nrow*S+1] scan over (row, split); == d_rowptr when S == 1
};

// rows_cap: most rows a band may hold (consumer's shared-memory budget); want_bands: preferred band count;
// S: column splits.  nnz > 0 and nrow > 0 required.  Runs on m->stream and synchronises it.
int build_band_plan(sb200_matrix* m, int rows_cap, int want_bands, int S, BandPlan** out);
void free_band_plan(BandPlan* bp, cudaStream_t s);
BandView make_view(const sb200_matrix* m, const BandPlan* bp);
// the banded two-pass placement kernel of round 1 (tall matrices, until they get their own path)
int launch_transpose_banded(sb200_matrix* m, const BandPlan* bp, int32_t* d_i_out, double* d_x_out);

}  // namespace sb200
