// Band-major companion of a resident mirror and the sweep that runs on it: A^T v (and, on the row-ordered copy,
// A v) with the operand in shared memory.
//
// Replaces, from the reference (zdebruine/RcppSparse), the InnerIterator sweep
//   y[col] += it.value() * v[it.row()]      (idiom of src/example.cpp:28-30, shape of RcppSparse.h:133-135)
// and, applied to the row-ordered copy of the mirror (A v = (A^T)^T v),
//   y[it.row()] += it.value() * v[col]      (shape of RcppSparse.h:140-142).
//
// Why.  sweep_kernel<SPMV_T> gathers v[i[k]] from L2, one 8-byte request per stored entry; B200's L2 serves
// ~245 G such gathers a second (profiles/r01), which caps the product at 0.37-0.55 of the HBM roofline.  The
// operand has to sit in shared memory, and 8 MB of it does not fit.  So the entries are regrouped ONCE, by
// (row band, column): band b holds rows [b*bw, (b+1)*bw), bw <= 12288 (96 KB of operand).
//
// Layout (round 2, third version; the history is in profiles/r02/README.md).  The run of column c inside band b
// is cut into BLOCKS of B = 8 entries; the last block of a run is padded with (x = 0, row = a dummy slot that holds
// 0.0 in the shared-memory slice, so a NaN or Inf in v never meets a padding entry).  32 consecutive blocks form a
// SLICE stored entry-pair-major — x as double2[B/2][32], the in-band row ids as ushort2[B/2][32], the column of
// every block as int32[32] — so that lane l of a warp owns block l of the slice and every load instruction of the
// warp reads one contiguous 512-byte (x) / 128-byte (rows, columns) line (B = 4 when runs are short: less padding).  A thread then has NOTHING to do per
// run: no bounds, no search, no merge path, no divergence — nine independent vector loads, eight shared-memory
// gathers, eight FMAs, one red.global.add.f64 on y[column].  ncu on the two earlier versions of this kernel
// (lane group per run, then thread per run on TMA-staged merge-path tiles) showed 100 and 47 thread instructions per
// entry and an issue-bound kernel; this one needs ~6.  The price is the padding: 13.6 instead of 10 bytes per entry
// at ~12 entries per run (C2, C4), 10.6 for long runs (C3) — against 12 algorithmic bytes.
//
// The layout is structure + values of the mirror, nothing about any result; it is built after the mirror has
// been asked for A^T v more than SB200_ROW_COMPANION_AFTER times (or on request, sb200_matrix_band_companion),
// costs one pass over the matrix and ~14 B per entry of HBM, and is dropped by sb200_matrix_refresh_values.
//
// Roofline: HBM.  Algorithmic bytes of the op stay SURVEY.md 8(d)'s 12N + 4(n+1) + 8n + 8m.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <new>
#include <string>
#include <vector>

#include "common.cuh"
#include "ptx.cuh"

namespace sb200 {

struct BandCompanion {
  int bw = 0;             // rows per band
  int nb = 0;             // bands
  int B = 8;              // entries per block (4 or 8)
  double mean_run = 0.0;  // entries per (band, column) run
  int64_t n_slices = 0;   // slices of 32 blocks over all bands (every band starts a new slice)
  double2* d_x = nullptr;      // [n_slices][B/2][32]
  ushort2* d_ri = nullptr;     // [n_slices][B/2][32] row inside the band; bw = the dummy slot of padding entries
  int32_t* d_col = nullptr;    // [n_slices][32] column of every block, -1 for the padding blocks at a band's end
  int64_t* d_sstart = nullptr; // [nb + 1] first slice of every band
  int64_t entries_padded = 0;
};

namespace {

constexpr int BS_THREADS = 512;
constexpr int BS_WARPS = BS_THREADS / 32;
constexpr int BS_MAX_BAND_ROWS = 12288;  // 96 KB of operand per band: two CTAs per SM

struct BsParams {
  const double2* x;
  const ushort2* ri;
  const int32_t* col;
  const int64_t* sstart;
  const double* v;
  double* y;
  int32_t nrow;
  int nb, bw;
  int64_t n_slices;
};

// One warp, one slice of 32 blocks: lane l sums block l.  All loads of the slice are issued before the first use.
template <int B>
__device__ __forceinline__ void sum_slice(const BsParams& prm, const double* __restrict__ vs, int64_t s, int lane) {
  constexpr int H = B / 2;
  const double2* __restrict__ xp = prm.x + s * (H * 32) + lane;
  const ushort2* __restrict__ rp = prm.ri + s * (H * 32) + lane;
  double2 xv[H];
  ushort2 rv[H];
#pragma unroll
  for (int j = 0; j < H; ++j) {
    xv[j] = ptx::ld_stream_v2f64(xp + j * 32);
    rv[j] = __ldg(rp + j * 32);
  }
  const int32_t c = __ldg(prm.col + s * 32 + lane);
  double a0 = 0.0, a1 = 0.0;
#pragma unroll
  for (int j = 0; j < H; ++j) {
    a0 = __fma_rn(xv[j].x, vs[rv[j].x], a0);
    a1 = __fma_rn(xv[j].y, vs[rv[j].y], a1);
  }
  double acc = __dadd_rn(a0, a1);
  // Long runs fill whole slices with blocks of ONE column (a dense row of A in the A v layout: 12288 entries per
  // band): 32 reductions on one address would queue up at one L2 slice.  Fold them in the warp first.
  const int32_t c0 = __shfl_sync(0xffffffffu, c, 0);
  if (__all_sync(0xffffffffu, c == c0)) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, off));
    if (lane == 0 && c >= 0) ptx::red_add_f64(prm.y + c, acc);
  } else if (c >= 0) {
    ptx::red_add_f64(prm.y + c, acc);
  }
}

template <int B>
__global__ void __launch_bounds__(BS_THREADS, 2) bandsweep_kernel(const BsParams prm) {
  extern __shared__ __align__(16) double vs[];  // [bw + 1]: the band's slice of the operand, then the dummy 0.0
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t s_begin = (prm.n_slices * blockIdx.x) / gridDim.x;
  const int64_t s_end = (prm.n_slices * (blockIdx.x + 1)) / gridDim.x;
  if (s_begin >= s_end) return;
  int b = 0;  // band of my first slice: largest b with sstart[b] <= s_begin
  {
    int lo = 0, hi = prm.nb - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (__ldg(prm.sstart + mid) <= s_begin)
        lo = mid;
      else
        hi = mid - 1;
    }
    b = lo;
  }
  int64_t s = s_begin;
  while (s < s_end) {
    while (b + 1 < prm.nb && __ldg(prm.sstart + b + 1) <= s) ++b;
    int64_t s_hi = __ldg(prm.sstart + b + 1);
    if (s_hi > s_end) s_hi = s_end;
    const int64_t row0 = static_cast<int64_t>(b) * prm.bw;
    int R = prm.bw;
    if (row0 + R > prm.nrow) R = static_cast<int>(prm.nrow - row0);
    __syncthreads();  // every warp is done with the previous band's slice
    for (int r = tid; r < R; r += BS_THREADS) vs[r] = __ldg(prm.v + row0 + r);
    if (tid == 0) vs[prm.bw] = 0.0;
    __syncthreads();
    // warps take the band's slices round-robin, two per step (18 vector loads in flight per thread)
    int64_t q = s + warp;
    for (; q + BS_WARPS < s_hi; q += 2 * BS_WARPS) {
      sum_slice<B>(prm, vs, q, lane);
      sum_slice<B>(prm, vs, q + BS_WARPS, lane);
    }
    if (q < s_hi) sum_slice<B>(prm, vs, q, lane);
    s = s_hi;
  }
}

// ---- build ---------------------------------------------------------------------------------------------------
// blocks per run, band-major: cnt[b*ncol + c] = ceil(entries of column c with row in band b / B)
__global__ void bbm_blocks_kernel(const int32_t* __restrict__ gp, const int32_t* __restrict__ bpt, int32_t ncol, int nb, int lgB,
                                  uint32_t* __restrict__ cnt) {
  const int64_t total = static_cast<int64_t>(nb) * ncol;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; g < total; g += stride) {
    const int b = static_cast<int>(g / ncol);
    const int64_t c = g - static_cast<int64_t>(b) * ncol;
    const int32_t s = (b == 0) ? __ldg(gp + c) : __ldg(bpt + static_cast<int64_t>(b - 1) * ncol + c);
    const int32_t e = (b == nb - 1) ? __ldg(gp + c + 1) : __ldg(bpt + static_cast<int64_t>(b) * ncol + c);
    cnt[g] = static_cast<uint32_t>((e - s + (1 << lgB) - 1) >> lgB);
  }
}

// bb[b] = blocks before band b (bo is the exclusive scan of the block counts), b = 0..nb-1
__global__ void bbm_band_blocks_kernel(const int32_t* __restrict__ bo, int32_t ncol, int nb, int32_t* __restrict__ bb) {
  for (int b = threadIdx.x; b < nb; b += blockDim.x) bb[b] = bo[static_cast<int64_t>(b) * ncol];
}

__global__ void bbm_fill_kernel(double2* __restrict__ x, ushort2* __restrict__ ri, int32_t* __restrict__ col, int64_t n_pairs,
                                int64_t n_blocks, unsigned short dummy) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t t0 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (int64_t k = t0; k < n_pairs; k += stride) {
    x[k] = make_double2(0.0, 0.0);
    ri[k] = make_ushort2(dummy, dummy);
  }
  for (int64_t k = t0; k < n_blocks; k += stride) col[k] = -1;
}

// Copy every run into its blocks.  A warp takes 32 consecutive runs and walks their concatenated entries 32 at a
// time, four steps in flight (coalesced reads); the writes land in the slice layout.  The owner lane of a run
// writes the run's column into its blocks.
template <int LGB>
__global__ void __launch_bounds__(256) bbm_copy_kernel(const int32_t* __restrict__ gi, const int32_t* __restrict__ gp,
                                                      const double* __restrict__ gx, const int32_t* __restrict__ bpt,
                                                      const int32_t* __restrict__ bo, const int32_t* __restrict__ shift,
                                                      int32_t ncol, int nb, int bw, double* __restrict__ xo,
                                                      unsigned short* __restrict__ ro, int32_t* __restrict__ colo) {
  constexpr int U = 4;
  constexpr int B = 1 << LGB, H = B / 2;
  const int lane = threadIdx.x & 31;
  const int64_t total = static_cast<int64_t>(nb) * ncol;
  const int64_t n_groups = (total + 31) / 32;
  const int64_t warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t w = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; w < n_groups; w += warps) {
    const int64_t g = w * 32 + lane;
    int32_t s = 0, len = 0, row0 = 0, blk0 = 0;
    if (g < total) {
      const int b = static_cast<int>(g / ncol);
      const int64_t c = g - static_cast<int64_t>(b) * ncol;
      s = (b == 0) ? __ldg(gp + c) : __ldg(bpt + static_cast<int64_t>(b - 1) * ncol + c);
      const int32_t e = (b == nb - 1) ? __ldg(gp + c + 1) : __ldg(bpt + static_cast<int64_t>(b) * ncol + c);
      len = e - s;
      row0 = b * bw;
      blk0 = __ldg(bo + g) + __ldg(shift + b);
      const int nblk = (len + B - 1) >> LGB;
      for (int q = 0; q < nblk; ++q) colo[blk0 + q] = static_cast<int32_t>(c);
    }
    int32_t incl = len;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int32_t up = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += up;
    }
    const int32_t excl = incl - len;
    const int32_t tot = __shfl_sync(0xffffffffu, incl, 31);
    for (int32_t base = 0; base < tot; base += 32 * U) {
      int32_t rr[U];
      double xx[U];
      int64_t dst[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int32_t q = base + u * 32 + lane;
        int l = 0;  // largest l with excl[l] <= q (skips empty runs)
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
          const int cand = l + step;
          const int32_t e = __shfl_sync(0xffffffffu, excl, cand & 31);
          if (cand < 32 && e <= q) l = cand;
        }
        const int32_t rs = __shfl_sync(0xffffffffu, s, l);
        const int32_t re = __shfl_sync(0xffffffffu, excl, l);
        const int32_t r0 = __shfl_sync(0xffffffffu, row0, l);
        const int32_t bq = __shfl_sync(0xffffffffu, blk0, l);
        rr[u] = 0;
        xx[u] = 0.0;
        dst[u] = -1;
        if (q < tot) {
          const int32_t j = q - re;  // position inside the run
          const int32_t k = rs + j;
          rr[u] = ptx::ld_stream_s32(gi + k) - r0;
          xx[u] = ptx::ld_stream_f64(gx + k);
          const int64_t blk = static_cast<int64_t>(bq) + (j >> LGB);
          const int jj = j & (B - 1);
          // slice layout: element jj of block blk sits at ((slice * H + jj / 2) * 32 + block-in-slice) * 2 + (jj & 1)
          dst[u] = (((blk >> 5) * H + (jj >> 1)) * 32 + (blk & 31)) * 2 + (jj & 1);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (dst[u] >= 0) {
          ro[dst[u]] = static_cast<unsigned short>(rr[u]);
          xo[dst[u]] = xx[u];
        }
      }
    }
  }
}

void free_companion(BandCompanion* bc, cudaStream_t s) {
  if (!bc) return;
  pool_free(bc->d_x, s);
  pool_free(bc->d_ri, s);
  pool_free(bc->d_col, s);
  pool_free(bc->d_sstart, s);
  delete bc;
}

int build_companion(sb200_matrix* m, BandCompanion** out) {
  cudaStream_t st = m->stream;
  const int32_t nrow = m->nrow, ncol = m->ncol;
  const int64_t nnz = m->nnz;
  int nb = static_cast<int>((static_cast<int64_t>(nrow) + BS_MAX_BAND_ROWS - 1) / BS_MAX_BAND_ROWS);
  if (nb < 1) nb = 1;
  int bw = static_cast<int>((static_cast<int64_t>(nrow) + nb - 1) / nb);
  bw = (bw + 15) & ~15;  // equal bands; the last one may be a few rows shorter
  if (bw > BS_MAX_BAND_ROWS) bw = BS_MAX_BAND_ROWS;
  if (const char* e = getenv("SB200_BMC_ROWS")) {
    const int v = atoi(e);
    if (v >= 16 && v <= BS_MAX_BAND_ROWS) bw = v & ~15;
  }
  nb = static_cast<int>((static_cast<int64_t>(nrow) + bw - 1) / bw);
  if (nb < 1) nb = 1;
  const int64_t G = static_cast<int64_t>(nb) * ncol;
  if (G > 2147483647LL - 8) return fail(SB200_E_UNSUPPORTED, "band companion: more than 2^31 (band, column) runs");
  BandCompanion* bc = new (std::nothrow) BandCompanion();
  if (!bc) return fail(SB200_E_NOMEM, "host allocation failed");
  bc->bw = bw;
  bc->nb = nb;
  bc->mean_run = static_cast<double>(nnz) / static_cast<double>(G > 0 ? G : 1);
  bc->B = bc->mean_run < 24.0 ? 4 : 8;  // short runs: less padding; long runs: half the blocks, half the reductions
  if (const char* e = getenv("SB200_BS_BLOCK")) {
    const int v = atoi(e);
    if (v == 4 || v == 8) bc->B = v;
  }
  const int lgB = bc->B == 4 ? 2 : 3;
  int32_t* d_rb = nullptr;
  int32_t* d_bpt = nullptr;
  uint32_t* d_cnt = nullptr;
  int32_t* d_bo = nullptr;
  int32_t* d_bb = nullptr;
  int32_t* d_shift = nullptr;
  unsigned long long* d_total = nullptr;
  void* d_scan_ws = nullptr;
  struct Guard {
    BandCompanion*& bc;
    void* tmp[8];
    cudaStream_t s;
    bool armed = true;
    ~Guard() {
      for (void* q : tmp) pool_free(q, s);
      if (armed) free_companion(bc, s);
    }
  } guard{bc, {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}, st};

  SB_TRY(pool_alloc(reinterpret_cast<void**>(&d_cnt), sizeof(uint32_t) * static_cast<size_t>(G > 0 ? G : 1), st));
  guard.tmp[0] = d_cnt;
  SB_TRY(pool_alloc(reinterpret_cast<void**>(&d_bo), sizeof(int32_t) * static_cast<size_t>(G + 1), st));
  guard.tmp[1] = d_bo;
  const size_t scan_ws = scan_workspace_bytes(G);
  SB_TRY(pool_alloc(&d_scan_ws, scan_ws, st));
  guard.tmp[2] = d_scan_ws;
  SB_TRY(pool_alloc(reinterpret_cast<void**>(&d_bb), sizeof(int32_t) * (static_cast<size_t>(nb) + 1), st));
  guard.tmp[3] = d_bb;
  SB_TRY(pool_alloc(reinterpret_cast<void**>(&d_shift), sizeof(int32_t) * (static_cast<size_t>(nb) + 1), st));
  guard.tmp[4] = d_shift;
  SB_TRY(pool_alloc(reinterpret_cast<void**>(&d_total), sizeof(unsigned long long), st));
  guard.tmp[7] = d_total;
  SB_TRY(pool_alloc(reinterpret_cast<void**>(&bc->d_sstart), sizeof(int64_t) * (static_cast<size_t>(nb) + 1), st));
  if (nb > 1) {
    std::vector<int32_t> rb(static_cast<size_t>(nb) + 1);
    for (int b = 0; b <= nb; ++b) {
      const int64_t r = static_cast<int64_t>(b) * bw;
      rb[b] = static_cast<int32_t>(r < nrow ? r : nrow);
    }
    SB_TRY(pool_alloc(reinterpret_cast<void**>(&d_rb), sizeof(int32_t) * rb.size(), st));
    guard.tmp[5] = d_rb;
    SB_TRY(pool_alloc(reinterpret_cast<void**>(&d_bpt), sizeof(int32_t) * static_cast<size_t>(nb - 1) * static_cast<size_t>(ncol), st));
    guard.tmp[6] = d_bpt;
    SB_CUDA(cudaMemcpyAsync(d_rb, rb.data(), sizeof(int32_t) * rb.size(), cudaMemcpyHostToDevice, st));
    SB_CUDA(cudaStreamSynchronize(st));  // rb is a host temporary
    SB_TRY(launch_band_ptr(m, d_rb, nb, d_bpt));
  }
  const int64_t cap_blocks = static_cast<int64_t>(m->sm_count) * 16;
  {
    int64_t blocks = (G + 255) / 256;
    if (blocks > cap_blocks) blocks = cap_blocks;
    if (blocks < 1) blocks = 1;
    bbm_blocks_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(m->d_p, d_bpt, ncol, nb, lgB, d_cnt);
    count_launch();
    SB_CUDA(cudaGetLastError());
  }
  SB_TRY(exclusive_scan_u32(st, d_cnt, d_bo, G, d_total, d_scan_ws, scan_ws));
  bbm_band_blocks_kernel<<<1, 256, 0, st>>>(d_bo, ncol, nb, d_bb);
  count_launch();
  SB_CUDA(cudaGetLastError());
  std::vector<int32_t> bb(static_cast<size_t>(nb) + 1);
  unsigned long long total_blocks = 0;
  SB_CUDA(cudaMemcpyAsync(bb.data(), d_bb, sizeof(int32_t) * static_cast<size_t>(nb), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaMemcpyAsync(&total_blocks, d_total, sizeof(total_blocks), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  if (total_blocks > 2000000000ull) return fail(SB200_E_UNSUPPORTED, "band companion: more than 2e9 blocks");
  bb[nb] = static_cast<int32_t>(total_blocks);
  // every band starts a new slice of 32 blocks: shift[b] = padding blocks inserted before band b
  std::vector<int32_t> shift(static_cast<size_t>(nb) + 1);
  std::vector<int64_t> sstart(static_cast<size_t>(nb) + 1);
  int64_t at = 0;  // blocks so far, padded
  for (int b = 0; b < nb; ++b) {
    shift[b] = static_cast<int32_t>(at - bb[b]);
    sstart[b] = at >> 5;
    at += bb[b + 1] - bb[b];
    at = (at + 31) & ~static_cast<int64_t>(31);
  }
  shift[nb] = 0;
  sstart[nb] = at >> 5;
  const int64_t n_blocks = at;
  if (n_blocks > 2147483647LL - 64) return fail(SB200_E_UNSUPPORTED, "band companion: block index exceeds int32");
  bc->n_slices = n_blocks >> 5;
  bc->entries_padded = n_blocks * bc->B;
  SB_CUDA(cudaMemcpyAsync(d_shift, shift.data(), sizeof(int32_t) * shift.size(), cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(bc->d_sstart, sstart.data(), sizeof(int64_t) * sstart.size(), cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaStreamSynchronize(st));  // host temporaries
  const int64_t n_pairs = n_blocks * (bc->B / 2);
  SB_TRY(pool_alloc(reinterpret_cast<void**>(&bc->d_x), sizeof(double2) * static_cast<size_t>(n_pairs > 0 ? n_pairs : 1), st));
  SB_TRY(pool_alloc(reinterpret_cast<void**>(&bc->d_ri), sizeof(ushort2) * static_cast<size_t>(n_pairs > 0 ? n_pairs : 1), st));
  SB_TRY(pool_alloc(reinterpret_cast<void**>(&bc->d_col), sizeof(int32_t) * static_cast<size_t>(n_blocks > 0 ? n_blocks : 1), st));
  if (n_blocks > 0) {
    int64_t blocks = (n_pairs + 255) / 256;
    if (blocks > cap_blocks) blocks = cap_blocks;
    bbm_fill_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(bc->d_x, bc->d_ri, bc->d_col, n_pairs, n_blocks,
                                                                   static_cast<unsigned short>(bw));
    count_launch();
    SB_CUDA(cudaGetLastError());
    blocks = ((G + 31) / 32 + 7) / 8;  // 8 warps per block
    if (blocks > cap_blocks) blocks = cap_blocks;
    if (blocks < 1) blocks = 1;
    double* xo = reinterpret_cast<double*>(bc->d_x);
    unsigned short* ro = reinterpret_cast<unsigned short*>(bc->d_ri);
    if (lgB == 2)
      bbm_copy_kernel<2><<<static_cast<unsigned>(blocks), 256, 0, st>>>(m->d_i, m->d_p, m->d_x, d_bpt, d_bo, d_shift, ncol, nb, bw, xo, ro,
                                                                        bc->d_col);
    else
      bbm_copy_kernel<3><<<static_cast<unsigned>(blocks), 256, 0, st>>>(m->d_i, m->d_p, m->d_x, d_bpt, d_bo, d_shift, ncol, nb, bw, xo, ro,
                                                                        bc->d_col);
    count_launch();
    SB_CUDA(cudaGetLastError());
  }
  SB_CUDA(cudaStreamSynchronize(st));
  guard.armed = false;
  *out = bc;
  return SB200_OK;
}

template <int B>
int launch_bandsweep_t(const sb200_matrix* m, const BsParams& prm) {
  const size_t smem = sizeof(double) * (static_cast<size_t>(prm.bw) + 2);
  auto kern = bandsweep_kernel<B>;
  SB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  int per_sm = static_cast<int>((224 * 1024) / (smem + 1024));
  if (per_sm > 2) per_sm = 2;
  if (per_sm < 1) per_sm = 1;
  int64_t grid = static_cast<int64_t>(m->sm_count) * per_sm;
  if (grid > prm.n_slices) grid = prm.n_slices;
  kern<<<static_cast<unsigned>(grid), BS_THREADS, smem, m->stream>>>(prm);
  count_launch();
  SB_CUDA(cudaGetLastError());
  return SB200_OK;
}

}  // namespace

void drop_band_companion(sb200_matrix* m, cudaStream_t s) {
  if (m->bmc) {
    free_companion(m->bmc, s);
    m->bmc = nullptr;
  }
  if (m->bmc_state == 1) m->bmc_state = 0;
  m->spmv_t_calls = 0;
}

// Non-fatal: a shape or a memory budget the companion cannot serve leaves bmc_state = -1 and the L2-gather
// sweep keeps serving the mirror.  The caller's last-error text is preserved.
int build_band_companion(sb200_matrix* m) {
  if (m->bmc_state == 1) return SB200_OK;
  m->bmc_state = -1;
  if (m->nnz == 0 || m->nrow == 0 || m->ncol == 0) return SB200_OK;
  const std::string saved = sb200_last_error();
  BandCompanion* bc = nullptr;
  const int rc = build_companion(m, &bc);
  if (rc != SB200_OK) {
    cudaGetLastError();
    set_error(saved);
    return SB200_OK;
  }
  m->bmc = bc;
  m->bmc_state = 1;
  return SB200_OK;
}

int64_t band_companion_bytes(const sb200_matrix* m) {
  const BandCompanion* bc = m->bmc;
  if (!bc) return 0;
  return bc->entries_padded * 10 + (bc->entries_padded / bc->B) * 4 + static_cast<int64_t>(bc->nb + 1) * 8;
}

int launch_bandsweep(sb200_matrix* m, const double* d_v, double* d_out) {
  const BandCompanion* bc = m->bmc;
  if (!bc) return fail(SB200_E_INVALID, "no band-major companion on this mirror");
  if (m->ncol > 0) SB_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * static_cast<size_t>(m->ncol), m->stream));
  if (bc->n_slices == 0) return SB200_OK;
  BsParams prm;
  prm.x = bc->d_x;
  prm.ri = bc->d_ri;
  prm.col = bc->d_col;
  prm.sstart = bc->d_sstart;
  prm.v = d_v;
  prm.y = d_out;
  prm.nrow = m->nrow;
  prm.nb = bc->nb;
  prm.bw = bc->bw;
  prm.n_slices = bc->n_slices;
  return bc->B == 4 ? launch_bandsweep_t<4>(m, prm) : launch_bandsweep_t<8>(m, prm);
}

}  // namespace sb200
