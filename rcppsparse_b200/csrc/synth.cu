// Synthetic dgCMatrix generator, straight into HBM (tests and benchmarks only; not on the hot path).
//
// Bit-identical twin of rcppsparse_b200/synth.py (see its docstring for the recipe): everything is
// integer arithmetic on a 64-bit mixing hash of (seed, column, k), plus one IEEE division for the
// value, so the CPU oracle can be fed exactly the matrix the GPU holds without a 24 GB PCIe copy.
// Stands in for Matrix::rsparsematrix, which the reference's examples and benchmarks use
// (reference README.md:33-38, vignettes/Documentation.Rmd:377,425) and which needs R.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace sb200 {

namespace {

constexpr int SYNTH_TABLE = 4096;
constexpr int SYNTH_MAX_BANDS = 32;

struct SynthParams {
  int32_t nrow;
  int64_t col_begin;
  int64_t ncols;
  uint64_t seed;
  const int64_t* len_table;  // device, SYNTH_TABLE + 1 entries
  int32_t empty_permille;
  int32_t n_bands;
  int64_t band_lo[SYNTH_MAX_BANDS];
  int64_t band_hi[SYNTH_MAX_BANDS];
  int64_t band_w[SYNTH_MAX_BANDS];
  int64_t wsum;
};

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t h3(uint64_t seed, uint64_t a, uint64_t b) {
  return mix64(mix64(seed + a * 0xD6E8FEB86659FD93ull) ^ (b * 0xA0761D6478BD642Full));
}
__host__ __device__ __forceinline__ double value_from_hash(uint64_t h) {
  const int64_t s = static_cast<int64_t>((h & 0xFF) + ((h >> 8) & 0xFF) + ((h >> 16) & 0xFF) + ((h >> 24) & 0xFF));
  return static_cast<double>(s - 510) / 100.0;
}

__device__ __forceinline__ int64_t raw_length(const SynthParams& sp, uint64_t col) {
  const uint64_t h = h3(sp.seed, col, 0);
  const int64_t j = static_cast<int64_t>(h >> 52);
  const int64_t f = static_cast<int64_t>((h >> 36) & 0xFFFF);
  const int64_t q0 = sp.len_table[j], q1 = sp.len_table[j + 1];
  int64_t raw = q0 + (((q1 - q0) * f) >> 16);
  if (sp.empty_permille > 0 && static_cast<int64_t>(h3(sp.seed, col, 1) % 1000ull) < sp.empty_permille) raw = 0;
  if (raw < 0) raw = 0;
  if (raw > sp.nrow) raw = sp.nrow;
  return raw;
}

__device__ __forceinline__ int64_t band_count(const SynthParams& sp, int64_t raw, int j) {
  const int64_t share = (raw * sp.band_w[j]) / sp.wsum;
  const int64_t size = sp.band_hi[j] - sp.band_lo[j];
  return share < size ? share : size;
}

__global__ void synth_lengths_kernel(const SynthParams sp, uint32_t* __restrict__ lens) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; c < sp.ncols; c += stride) {
    const int64_t raw = raw_length(sp, static_cast<uint64_t>(sp.col_begin + c));
    int64_t tot = 0;
    for (int j = 0; j < sp.n_bands; ++j) tot += band_count(sp, raw, j);
    lens[c] = static_cast<uint32_t>(tot);
  }
}

__global__ void __launch_bounds__(256)
    synth_fill_kernel(const SynthParams sp, const int32_t* __restrict__ p, int32_t* __restrict__ gi,
                      double* __restrict__ gx) {
  __shared__ int64_t s_prefix[SYNTH_MAX_BANDS + 1];
  __shared__ int64_t s_cnt[SYNTH_MAX_BANDS];
  for (int64_t c = blockIdx.x; c < sp.ncols; c += gridDim.x) {
    const uint64_t gcol = static_cast<uint64_t>(sp.col_begin + c);
    __syncthreads();
    if (threadIdx.x == 0) {
      const int64_t raw = raw_length(sp, gcol);
      int64_t run = 0;
      for (int j = 0; j < sp.n_bands; ++j) {
        s_prefix[j] = run;
        s_cnt[j] = band_count(sp, raw, j);
        run += s_cnt[j];
      }
      s_prefix[sp.n_bands] = run;
    }
    __syncthreads();
    const int64_t len = s_prefix[sp.n_bands];
    const int64_t base = p[c];
    for (int64_t k = threadIdx.x; k < len; k += blockDim.x) {
      int j = 0;
      while (j + 1 < sp.n_bands && k >= s_prefix[j + 1]) ++j;
      const int64_t q = k - s_prefix[j];
      const int64_t cj = s_cnt[j];
      const int64_t sj = sp.band_hi[j] - sp.band_lo[j];
      const int64_t b0 = (q * sj) / cj;
      const int64_t b1 = ((q + 1) * sj) / cj;
      const uint64_t hr = h3(sp.seed, gcol, static_cast<uint64_t>(2 + 2 * k));
      const int64_t off = static_cast<int64_t>(hr % static_cast<uint64_t>(b1 - b0));
      gi[base + k] = static_cast<int32_t>(sp.band_lo[j] + b0 + off);
      gx[base + k] = value_from_hash(h3(sp.seed, gcol, static_cast<uint64_t>(3 + 2 * k)));
    }
  }
}

__global__ void synth_vector_kernel(uint64_t seed, int64_t begin, int64_t n, double* __restrict__ out) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; k < n; k += stride)
    out[k] = value_from_hash(h3(seed + 1, static_cast<uint64_t>(begin + k), 0));
}

}  // namespace
}  // namespace sb200

using namespace sb200;

extern "C" int sb200_synth_create(int32_t nrow, int64_t col_begin, int64_t col_end, uint64_t seed,
                                  const int64_t* len_table, int32_t empty_permille, int32_t n_bands,
                                  const int64_t* band_lo, const int64_t* band_hi, const int64_t* band_w, int device,
                                  sb200_matrix** out) {
  if (!out) return fail(SB200_E_INVALID, "out is NULL");
  *out = nullptr;
  if (!len_table || !band_lo || !band_hi || !band_w) return fail(SB200_E_INVALID, "NULL table");
  if (n_bands < 1 || n_bands > SYNTH_MAX_BANDS) return fail(SB200_E_INVALID, "n_bands must be in [1, 32]");
  if (nrow < 0 || col_begin < 0 || col_end < col_begin || col_end - col_begin > 2147483646LL)
    return fail(SB200_E_INVALID, "bad shape");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    return fail(SB200_E_NODEVICE, "no CUDA device available");
  }
  if (device < 0 || device >= ndev) return fail(SB200_E_NODEVICE, "requested CUDA device not present");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(SB200_E_CUDA, "cudaSetDevice failed");

  const int64_t ncols = col_end - col_begin;
  SynthParams sp;
  sp.nrow = nrow;
  sp.col_begin = col_begin;
  sp.ncols = ncols;
  sp.seed = seed;
  sp.empty_permille = empty_permille;
  sp.n_bands = n_bands;
  sp.wsum = 0;
  for (int j = 0; j < SYNTH_MAX_BANDS; ++j) {
    sp.band_lo[j] = j < n_bands ? band_lo[j] : 0;
    sp.band_hi[j] = j < n_bands ? band_hi[j] : 0;
    sp.band_w[j] = j < n_bands ? band_w[j] : 0;
    sp.wsum += sp.band_w[j];
  }
  if (sp.wsum <= 0) return fail(SB200_E_INVALID, "band weights must be positive");

  int64_t* d_table = nullptr;
  uint32_t* d_lens = nullptr;
  int32_t* d_ptmp = nullptr;
  void* d_scan_ws = nullptr;
  unsigned long long* d_total = nullptr;
  sb200_matrix* m = nullptr;
  int rc = SB200_OK;
  cudaStream_t st = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_table);
    cudaFree(d_lens);
    cudaFree(d_ptmp);
    cudaFree(d_scan_ws);
    cudaFree(d_total);
    if (st) cudaStreamDestroy(st);
  };
#define SY(expr)                                                      \
  do {                                                                \
    cudaError_t e_ = (expr);                                          \
    if (e_ != cudaSuccess) {                                          \
      rc = cuda_fail(e_, #expr, __FILE__, __LINE__);                  \
      cleanup();                                                      \
      if (m) sb200_matrix_destroy(m);                                 \
      return rc;                                                      \
    }                                                                 \
  } while (0)
  SY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  const size_t scan_ws = scan_workspace_bytes(ncols);
  SY(cudaMalloc(&d_table, sizeof(int64_t) * (SYNTH_TABLE + 1)));
  SY(cudaMalloc(&d_lens, sizeof(uint32_t) * static_cast<size_t>(ncols > 0 ? ncols : 1)));
  SY(cudaMalloc(&d_ptmp, sizeof(int32_t) * static_cast<size_t>(ncols + 1)));
  SY(cudaMalloc(&d_scan_ws, scan_ws));
  SY(cudaMalloc(&d_total, sizeof(unsigned long long)));
  SY(cudaMemcpyAsync(d_table, len_table, sizeof(int64_t) * (SYNTH_TABLE + 1), cudaMemcpyHostToDevice, st));
  sp.len_table = d_table;
  if (ncols > 0) {
    int64_t blocks = (ncols + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    synth_lengths_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(sp, d_lens);
    count_launch();
    SY(cudaGetLastError());
  }
  rc = exclusive_scan_u32(st, d_lens, d_ptmp, ncols, d_total, d_scan_ws, scan_ws);
  if (rc != SB200_OK) {
    cleanup();
    return rc;
  }
  unsigned long long h_total = 0;
  SY(cudaMemcpyAsync(&h_total, d_total, sizeof(h_total), cudaMemcpyDeviceToHost, st));
  SY(cudaStreamSynchronize(st));
  if (h_total > 2147483647ull) {
    cleanup();
    return fail(SB200_E_INVALID, "synthetic matrix would exceed int32 nnz (dgCMatrix limit)");
  }
  rc = alloc_matrix(device, nrow, static_cast<int32_t>(ncols), static_cast<int64_t>(h_total), &m);
  if (rc != SB200_OK) {
    cleanup();
    return rc;
  }
  SY(cudaMemcpyAsync(m->d_p, d_ptmp, sizeof(int32_t) * static_cast<size_t>(ncols + 1), cudaMemcpyDeviceToDevice, st));
  if (h_total > 0) {
    int64_t blocks = ncols;
    if (blocks > 148 * 16) blocks = 148 * 16;
    synth_fill_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(sp, m->d_p, m->d_i, m->d_x);
    count_launch();
    SY(cudaGetLastError());
  }
  SY(cudaStreamSynchronize(st));
#undef SY
  cleanup();
  rc = finish_matrix(m, 0);
  if (rc != SB200_OK) {
    sb200_matrix_destroy(m);
    return rc;
  }
  *out = m;
  return SB200_OK;
}

extern "C" int sb200_synth_vector_dev(sb200_matrix* m, uint64_t seed, int64_t begin, int64_t n, double* d_out) {
  SB_TRY(check_handle(m));
  DeviceGuard guard(m->device);
  if (!guard.ok) return fail(SB200_E_CUDA, "cudaSetDevice failed");
  if (n <= 0) return SB200_OK;
  if (!d_out) return fail(SB200_E_INVALID, "d_out is NULL");
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  synth_vector_kernel<<<static_cast<unsigned>(blocks), 256, 0, m->stream>>>(seed, begin, n, d_out);
  count_launch();
  SB_CUDA(cudaGetLastError());
  return SB200_OK;
}
