// CSC -> CSR transpose (= CSC of A^T) on B200: histogram / scan / band pointers / banded stable scatter.
//
// Replaces Matrix::transpose(), reference RcppSparse.h:375-385, whose arithmetic is R's Matrix::t
// (a serial counting sort in the Matrix package's C code).  Output contract: Dim swapped,
// p'[nrow+1], i' = source column ids ASCENDING inside every new column, x' permuted bit-exactly.
//
// The hard part is "ascending": positions inside an output row must follow source-column order,
// i.e. the scatter has to be STABLE, while 148 SMs append to the same rows concurrently.
// Design (all integer work, HBM-bound; algorithmic bytes 24N + 4(n+1) + 4(m+1)):
//
//  T1  row histogram: stream i once with 128-bit loads; counts privatised in shared memory when
//      the row range fits (u32[nrow] <= 192 KB), flushed with one global add per touched row.
//  T2  p' = exclusive scan of the counts (scan.cu, decoupled look-back).
//  T3  cut the ROW range into nb bands of ~equal nnz (binary searches in p'), bounded in rows so a
//      band's cursors fit in shared memory.
//  T4  band pointers: for every column c and band b the offset of the first entry of c whose row
//      is >= rb[b] (rows are sorted inside a column, so a band's entries are one contiguous run).
//      One more streaming read of i over merge-path tiles; writes (nb-1)*ncol offsets.
//  T5  banded scatter: CTA b owns band b's output rows EXCLUSIVELY and sweeps all columns in
//      order, reading only its runs.  Per round it takes up to 32*KW consecutive columns:
//        phase 1  every entry sets bit (column-in-round) in its row's bitmask   (order-free atomicOr)
//        phase 2  rank = popcount of lower bits = number of earlier columns in the round holding
//                 the same row; slot = cursor[row] + rank; store column id and value
//        phase 3  the entry owning the top bit advances cursor[row] by the row's popcount, clears it
//      No two entries of a round race for a slot, the order inside a row is source-column order by
//      construction, and each output row is appended to by exactly one CTA, so its write frontier
//      (one open 32-byte sector per row and array) stays in L2 until the sector is complete.
//      Neighbouring bands read neighbouring runs of the same columns at about the same time, so
//      partially used sectors are served from L2, not fetched from HBM twice.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace sb200 {

namespace {

// ------------------------------------------------------------------------------------------------
// T1: row histogram
// ------------------------------------------------------------------------------------------------
constexpr int HIST_THREADS = 512;
constexpr int HIST_SMEM_MAX_ROWS = 48 * 1024;  // u32 counters, 192 KB

template <bool PRIVATE>
__global__ void __launch_bounds__(HIST_THREADS)
    row_hist_kernel(const int32_t* __restrict__ gi, int64_t nnz, int32_t nrow, uint32_t* __restrict__ cnt) {
  extern __shared__ uint32_t h[];
  if (PRIVATE) {
    for (int r = threadIdx.x; r < nrow; r += HIST_THREADS) h[r] = 0u;
    __syncthreads();
  }
  const int64_t n4 = nnz >> 2;
  const int64_t g_begin = (n4 * blockIdx.x) / gridDim.x;
  const int64_t g_end = (n4 * (blockIdx.x + 1)) / gridDim.x;
  const int4* __restrict__ i4 = reinterpret_cast<const int4*>(gi);
  for (int64_t g = g_begin + threadIdx.x; g < g_end; g += HIST_THREADS) {
    const int4 r = ptx::ld_stream_v4s32(i4 + g);
    if (PRIVATE) {
      atomicAdd(&h[r.x], 1u);
      atomicAdd(&h[r.y], 1u);
      atomicAdd(&h[r.z], 1u);
      atomicAdd(&h[r.w], 1u);
    } else {
      ptx::red_add_u32(cnt + r.x, 1u);
      ptx::red_add_u32(cnt + r.y, 1u);
      ptx::red_add_u32(cnt + r.z, 1u);
      ptx::red_add_u32(cnt + r.w, 1u);
    }
  }
  if (blockIdx.x == gridDim.x - 1) {
    const int64_t k = (n4 << 2) + threadIdx.x;
    if (k < nnz) {
      if (PRIVATE)
        atomicAdd(&h[gi[k]], 1u);
      else
        ptx::red_add_u32(cnt + gi[k], 1u);
    }
  }
  if (PRIVATE) {
    __syncthreads();
    for (int r = threadIdx.x; r < nrow; r += HIST_THREADS) {
      const uint32_t c = h[r];
      if (c) ptx::red_add_u32(cnt + r, c);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// T3: band boundaries.  key(r) = (1-eps) * p'[r]/nnz + eps * r/nrow is non-decreasing; band b starts
// at the first row whose key reaches b/nb.  eps > 0 bounds the rows of a band (shared-memory cursors).
// ------------------------------------------------------------------------------------------------
__global__ void band_bounds_kernel(const int32_t* __restrict__ p_out, int32_t nrow, int64_t nnz, int nb, double eps,
                                   int32_t* __restrict__ rb, int32_t* __restrict__ max_rows) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > nb) return;
  int32_t r;
  if (b == 0) {
    r = 0;
  } else if (b == nb) {
    r = nrow;
  } else {
    const double wa = (1.0 - eps) / static_cast<double>(nnz);
    const double wb = eps / static_cast<double>(nrow);
    const double target = static_cast<double>(b) / static_cast<double>(nb);
    int32_t lo = 0, hi = nrow;
    while (lo < hi) {
      const int32_t mid = lo + ((hi - lo) >> 1);
      const double key = static_cast<double>(p_out[mid]) * wa + static_cast<double>(mid) * wb;
      if (key < target)
        lo = mid + 1;
      else
        hi = mid;
    }
    r = lo;
  }
  rb[b] = r;
}

__global__ void band_max_rows_kernel(const int32_t* __restrict__ rb, int nb, int32_t* __restrict__ max_rows) {
  int32_t mx = 0;
  for (int b = threadIdx.x; b < nb; b += blockDim.x) {
    const int32_t d = rb[b + 1] - rb[b];
    mx = d > mx ? d : mx;
  }
  atomicMax(max_rows, mx);
}

// ------------------------------------------------------------------------------------------------
// T4: band pointers over merge-path tiles (same plan as the sweeps: SWEEP_TILE items per tile)
// bpt[(b-1)*ncol + c] = first k in column c with i[k] >= rb[b], for b = 1..nb-1
// ------------------------------------------------------------------------------------------------
constexpr int BP_THREADS = 256;
constexpr int BP_IPT = SWEEP_TILE / BP_THREADS;  // 14
static_assert(BP_THREADS * BP_IPT == SWEEP_TILE, "band-pointer tiling must match the sweep plan");

__global__ void __launch_bounds__(BP_THREADS)
    band_ptr_kernel(const int32_t* __restrict__ gi, const int32_t* __restrict__ gp, const int32_t* __restrict__ plan,
                    int64_t n_tiles, int32_t ncol, int32_t nnz, const int32_t* __restrict__ rb, int nb,
                    int32_t* __restrict__ bpt) {
  extern __shared__ int32_t sm[];
  int32_t* as = sm;                     // as[j] = p[c0 + j], j = 0..nc+1   (SWEEP_TILE + 2)
  int32_t* is = sm + (SWEEP_TILE + 4);  // is[j] = i[k0 - 1 + j]            (SWEEP_TILE + 1)
  int32_t* rbs = is + (SWEEP_TILE + 4); // rbs[b] = rb[b], b = 0..nb
  for (int b = threadIdx.x; b <= nb; b += BP_THREADS) rbs[b] = rb[b];
  const int64_t total_items = static_cast<int64_t>(ncol) + nnz;

  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    __syncthreads();  // previous tile fully consumed (also covers the rbs fill)
    const int32_t c0 = plan[t], c1 = plan[t + 1];
    const int64_t d0 = t * SWEEP_TILE;
    int64_t d1 = d0 + SWEEP_TILE;
    if (d1 > total_items) d1 = total_items;
    const int32_t k0 = static_cast<int32_t>(d0 - c0), k1 = static_cast<int32_t>(d1 - c1);
    const int nc = c1 - c0, nk = k1 - k0;
    const int a_last = (c1 + 1 <= ncol) ? c1 + 1 : ncol;  // p index
    for (int j = threadIdx.x; c0 + j <= a_last; j += BP_THREADS) as[j] = gp[c0 + j];
    for (int j = threadIdx.x; j <= nk; j += BP_THREADS) {
      const int32_t k = k0 - 1 + j;
      is[j] = (k >= 0 && k < nnz) ? ptx::ld_stream_s32(gi + k) : 0;
    }
    __syncthreads();

    const int items = nc + nk;
    int d_lo = threadIdx.x * BP_IPT;
    if (d_lo > items) d_lo = items;
    int d_hi = d_lo + BP_IPT;
    if (d_hi > items) d_hi = items;
    if (d_lo >= d_hi) continue;
    // column ends consumed before d_lo: ends are as[1..], end of column c0+j is as[j+1]
    int lo = d_lo > nk ? d_lo - nk : 0;
    int hi = d_lo < nc ? d_lo : nc;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (as[mid + 1] - k0 <= d_lo - mid - 1)
        lo = mid + 1;
      else
        hi = mid;
    }
    int ci = lo;
    int ki = d_lo - ci;                // relative to k0; entry k0+ki is is[ki+1]
    int col_end = as[ci + 1] - k0;     // relative end of column c0+ci (stale past the last column: no items remain)
    // band of the previous entry when I start inside a column
    int cur = 0;
    if (k0 + ki > as[ci]) {  // column c0+ci started before my first item
      const int32_t prev_row = is[ki];
      int blo = 0, bhi = nb;  // largest b with rbs[b] <= prev_row
      while (blo < bhi) {
        const int mid = (blo + bhi + 1) >> 1;
        if (rbs[mid] <= prev_row)
          blo = mid;
        else
          bhi = mid - 1;
      }
      cur = blo;
    }
    int32_t next_rb = rbs[cur + 1];  // cur <= nb-1 always (rows < nrow = rbs[nb])
    for (int it = d_lo; it < d_hi; ++it) {
      const int64_t c = c0 + ci;
      if (ki < col_end) {
        const int32_t row = is[ki + 1];
        while (row >= next_rb) {  // entering band cur+1 (possibly skipping empty bands)
          ++cur;
          bpt[static_cast<int64_t>(cur - 1) * ncol + c] = k0 + ki;
          next_rb = rbs[cur + 1];
        }
        ++ki;
      } else {
        const int32_t e = k0 + col_end;
        while (cur < nb - 1) {  // bands after the column's last entry start at its end
          ++cur;
          bpt[static_cast<int64_t>(cur - 1) * ncol + c] = e;
        }
        cur = 0;
        next_rb = rbs[1];
        ++ci;
        col_end = as[ci + 1] - k0;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// T5: banded stable scatter
// ------------------------------------------------------------------------------------------------
constexpr int TB_THREADS = 512;
constexpr int TB_CH = TB_THREADS;  // columns per chunk (one run descriptor per thread)
constexpr int TB_EPT = 2;          // entries per thread per round
constexpr int TB_EMAX = TB_THREADS * TB_EPT;

struct BandArgs {
  const int32_t* i;
  const int32_t* p;
  const double* x;
  int32_t ncol;
  int nb;
  const int32_t* rb;
  const int32_t* bpt;
  const int32_t* p_out;
  int32_t* i_out;
  double* x_out;
  int max_rows;  // shared-memory capacity in rows
};

__device__ __forceinline__ int32_t band_start(const BandArgs& a, int b, int64_t c) {
  if (b == 0) return __ldg(a.p + c);
  if (b == a.nb) return __ldg(a.p + c + 1);
  return __ldg(a.bpt + static_cast<int64_t>(b - 1) * a.ncol + c);
}

template <int KW>
__global__ void __launch_bounds__(TB_THREADS) transpose_band_kernel(const BandArgs a) {
  extern __shared__ uint32_t tsm[];
  uint32_t* cursor = tsm;                      // [max_rows] next free slot of each row of the band
  uint32_t* bits = cursor + a.max_rows;        // [max_rows * KW] column-in-round bitmask per row
  __shared__ int32_t run_start[TB_CH];
  __shared__ int32_t run_off[TB_CH + 1];
  __shared__ int32_t warp_tot[TB_THREADS / 32];
  constexpr int KCOLS = 32 * KW;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int b = blockIdx.x; b < a.nb; b += gridDim.x) {
    const int32_t row0 = a.rb[b];
    const int32_t R = a.rb[b + 1] - row0;
    if (R <= 0) continue;
    __syncthreads();
    for (int r = tid; r < R; r += TB_THREADS) cursor[r] = static_cast<uint32_t>(a.p_out[row0 + r]);
    for (int r = tid; r < R * KW; r += TB_THREADS) bits[r] = 0u;

    // run descriptor of my column in the first chunk
    int32_t nxt_s = 0, nxt_e = 0;
    if (tid < a.ncol) {
      nxt_s = band_start(a, b, tid);
      nxt_e = band_start(a, b + 1, tid);
    }
    for (int64_t cbase = 0; cbase < a.ncol; cbase += TB_CH) {
      const int32_t s = nxt_s, e = nxt_e;
      // fetch the next chunk's descriptor early and pull its run towards L2
      {
        const int64_t cn = cbase + TB_CH + tid;
        nxt_s = nxt_e = 0;
        if (cn < a.ncol) {
          nxt_s = band_start(a, b, cn);
          nxt_e = band_start(a, b + 1, cn);
          for (int32_t k = nxt_s & ~31; k < nxt_e; k += 32) ptx::prefetch_l2(a.i + k);
          for (int32_t k = nxt_s & ~15; k < nxt_e; k += 16) ptx::prefetch_l2(a.x + k);
        }
      }
      // exclusive scan of run lengths over the chunk
      const int32_t len = e - s;
      int32_t inc = len;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int32_t up = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += up;
      }
      __syncthreads();  // previous chunk's rounds are over: run_* and warp_tot may be rewritten
      if (lane == 31) warp_tot[warp] = inc;
      __syncthreads();
      int32_t wbase = 0, chunk_total = 0;
#pragma unroll
      for (int w = 0; w < TB_THREADS / 32; ++w) {
        if (w < warp) wbase += warp_tot[w];
        chunk_total += warp_tot[w];
      }
      run_start[tid] = s;
      run_off[tid] = wbase + inc - len;
      if (tid == 0) run_off[TB_CH] = chunk_total;
      __syncthreads();

      int32_t q0 = 0;
      while (q0 < chunk_total) {  // uniform across the CTA
        // first column of the round: largest j with run_off[j] <= q0 (skips exhausted / empty runs)
        int jl = 0, jh = TB_CH - 1;
        while (jl < jh) {
          const int mid = (jl + jh + 1) >> 1;
          if (run_off[mid] <= q0)
            jl = mid;
          else
            jh = mid - 1;
        }
        const int j0 = jl;
        const int jend = (j0 + KCOLS < TB_CH) ? j0 + KCOLS : TB_CH;
        int32_t qend = run_off[jend];
        if (qend > q0 + TB_EMAX) qend = q0 + TB_EMAX;

        int32_t lr[TB_EPT];   // row inside the band, -1 = no entry
        int32_t jj[TB_EPT];   // column inside the round
        double val[TB_EPT];
        // ---- phase 1: load the round's entries, set (row, column-in-round) bits ----------------
#pragma unroll
        for (int u = 0; u < TB_EPT; ++u) {
          const int32_t q = q0 + tid + u * TB_THREADS;
          lr[u] = -1;
          jj[u] = 0;
          val[u] = 0.0;
          if (q < qend) {
            int l = j0, h = jend - 1;
            while (l < h) {
              const int mid = (l + h + 1) >> 1;
              if (run_off[mid] <= q)
                l = mid;
              else
                h = mid - 1;
            }
            const int32_t k = run_start[l] + (q - run_off[l]);
            lr[u] = ptx::ld_stream_s32(a.i + k) - row0;
            val[u] = ptx::ld_stream_f64(a.x + k);
            jj[u] = l - j0;
          }
        }
#pragma unroll
        for (int u = 0; u < TB_EPT; ++u)
          if (lr[u] >= 0) atomicOr(&bits[lr[u] * KW + (jj[u] >> 5)], 1u << (jj[u] & 31));
        __syncthreads();
        // ---- phase 2: rank inside the round = set bits below mine; store -------------------------
        bool top[TB_EPT];
        uint32_t rowcnt[TB_EPT];
#pragma unroll
        for (int u = 0; u < TB_EPT; ++u) {
          top[u] = false;
          rowcnt[u] = 0;
          if (lr[u] >= 0) {
            const int wi = jj[u] >> 5;
            const uint32_t below_mask = (1u << (jj[u] & 31)) - 1u;
            uint32_t rank = 0, total = 0;
            bool higher = false;
#pragma unroll
            for (int w = 0; w < KW; ++w) {
              const uint32_t word = bits[lr[u] * KW + w];
              total += __popc(word);
              if (w < wi) rank += __popc(word);
              if (w == wi) {
                rank += __popc(word & below_mask);
                higher = higher || ((word >> (jj[u] & 31)) >> 1) != 0u;
              }
              if (w > wi) higher = higher || word != 0u;
            }
            top[u] = !higher;
            rowcnt[u] = total;
            const uint32_t pos = cursor[lr[u]] + rank;
            a.i_out[pos] = static_cast<int32_t>(cbase + j0 + jj[u]);
            a.x_out[pos] = val[u];
          }
        }
        __syncthreads();
        // ---- phase 3: the top entry of each row advances the cursor and clears the row's bits -----
#pragma unroll
        for (int u = 0; u < TB_EPT; ++u) {
          if (lr[u] >= 0 && top[u]) {
            cursor[lr[u]] += rowcnt[u];
#pragma unroll
            for (int w = 0; w < KW; ++w) bits[lr[u] * KW + w] = 0u;
          }
        }
        __syncthreads();
        q0 = qend;
      }
    }
  }
}

__global__ void zero_i32_kernel(int32_t* d, int64_t n) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; k < n; k += stride) d[k] = 0;
}

struct Scratch {  // freed on every exit path
  void* ptr[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int n = 0;
  int alloc(void** out, size_t bytes) {
    cudaError_t e = cudaMalloc(out, bytes ? bytes : 16);
    if (e != cudaSuccess) return fail(SB200_E_NOMEM, std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e));
    ptr[n++] = *out;
    return SB200_OK;
  }
  ~Scratch() {
    for (int k = 0; k < n; ++k) cudaFree(ptr[k]);
  }
};

}  // namespace

int transpose_device(sb200_matrix* m, int32_t* d_p_out, int32_t* d_i_out, double* d_x_out) {
  cudaStream_t st = m->stream;
  const int32_t nrow = m->nrow, ncol = m->ncol;
  const int64_t nnz = m->nnz;
  if (nnz == 0 || nrow == 0) {
    int64_t blocks = (static_cast<int64_t>(nrow) + 1 + 255) / 256;
    if (blocks > 1024) blocks = 1024;
    zero_i32_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(d_p_out, static_cast<int64_t>(nrow) + 1);
    count_launch();
    SB_CUDA(cudaGetLastError());
    return SB200_OK;
  }

  // ---- geometry ------------------------------------------------------------------------------------
  int ctas_per_sm = 2;
  if (const char* e = getenv("SB200_TRANSPOSE_CTAS")) {
    const int v = atoi(e);
    if (v >= 1 && v <= 2) ctas_per_sm = v;
  }
  const int grid = m->sm_count * ctas_per_sm;
  const size_t smem_budget = (ctas_per_sm == 1 ? 196 : 98) * 1024;  // dynamic part per CTA
  const int rmax2 = static_cast<int>(smem_budget / (4 * (1 + 2)));  // rows a band may hold with KW = 2
  const int rmax4 = static_cast<int>(smem_budget / (4 * (1 + 4)));
  int passes = 1;
  double eps = 0.0;
  while (true) {
    const double need = static_cast<double>(nrow) / (static_cast<double>(grid) * passes * 0.9 * rmax2);
    if (need <= 0.5) {
      eps = need < 1e-3 ? 1e-3 : need;
      break;
    }
    ++passes;
    if (passes > 4096) return fail(SB200_E_UNSUPPORTED, "transpose: row count too large for the banded path");
  }
  int nb = grid * passes;
  // never more bands than useful: a band should own >= ~4 entries per column on average
  {
    const int64_t by_density = nnz / (static_cast<int64_t>(ncol > 0 ? ncol : 1) * 4);
    const int64_t floor_rows = (static_cast<int64_t>(nrow) + (rmax2 * 9 / 10) - 1) / (rmax2 * 9 / 10);
    int64_t want = by_density;
    if (want < floor_rows) want = floor_rows;
    if (want < 1) want = 1;
    if (want < nb) {
      nb = static_cast<int>(want);
      const double need = static_cast<double>(nrow) / (static_cast<double>(nb) * 0.9 * rmax2);
      eps = need < 1e-3 ? 1e-3 : (need > 1.0 ? 1.0 : need);
    }
  }
  if (const char* e = getenv("SB200_TRANSPOSE_BANDS")) {
    const int v = atoi(e);
    if (v >= 1) {
      nb = v;
      const double need = static_cast<double>(nrow) / (static_cast<double>(nb) * 0.9 * rmax2);
      if (need > 1.0) return fail(SB200_E_INVALID, "SB200_TRANSPOSE_BANDS too small for this row count");
      eps = need < 1e-3 ? 1e-3 : need;
    }
  }

  // ---- scratch ----------------------------------------------------------------------------------------
  Scratch sc;
  uint32_t* d_cnt = nullptr;
  void* d_scan_ws = nullptr;
  int32_t* d_rb = nullptr;
  int32_t* d_bpt = nullptr;
  int32_t* d_maxrows = nullptr;
  const size_t scan_ws = scan_workspace_bytes(nrow);
  SB_TRY(sc.alloc(reinterpret_cast<void**>(&d_cnt), sizeof(uint32_t) * static_cast<size_t>(nrow)));
  SB_TRY(sc.alloc(&d_scan_ws, scan_ws));
  SB_TRY(sc.alloc(reinterpret_cast<void**>(&d_rb), sizeof(int32_t) * static_cast<size_t>(nb + 1)));
  SB_TRY(sc.alloc(reinterpret_cast<void**>(&d_bpt),
                  sizeof(int32_t) * static_cast<size_t>(nb > 1 ? nb - 1 : 1) * static_cast<size_t>(ncol > 0 ? ncol : 1)));
  SB_TRY(sc.alloc(reinterpret_cast<void**>(&d_maxrows), sizeof(int32_t)));

  // ---- T1 histogram ---------------------------------------------------------------------------------------
  SB_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(uint32_t) * static_cast<size_t>(nrow), st));
  {
    int64_t blocks = (nnz / 4 + HIST_THREADS * 8 - 1) / (HIST_THREADS * 8);
    if (blocks < 1) blocks = 1;
    if (nrow <= HIST_SMEM_MAX_ROWS) {
      if (blocks > m->sm_count) blocks = m->sm_count;
      const size_t smem = sizeof(uint32_t) * static_cast<size_t>(nrow);
      SB_CUDA(cudaFuncSetAttribute(row_hist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(smem)));
      row_hist_kernel<true><<<static_cast<unsigned>(blocks), HIST_THREADS, smem, st>>>(m->d_i, nnz, nrow, d_cnt);
    } else {
      if (blocks > m->sm_count * 4) blocks = m->sm_count * 4;
      row_hist_kernel<false><<<static_cast<unsigned>(blocks), HIST_THREADS, 0, st>>>(m->d_i, nnz, nrow, d_cnt);
    }
    count_launch();
    SB_CUDA(cudaGetLastError());
  }
  // ---- T2 scan ------------------------------------------------------------------------------------------
  SB_TRY(exclusive_scan_u32(st, d_cnt, d_p_out, nrow, nullptr, d_scan_ws, scan_ws));
  // ---- T3 bands -----------------------------------------------------------------------------------------
  band_bounds_kernel<<<(nb + 1 + 127) / 128, 128, 0, st>>>(d_p_out, nrow, nnz, nb, eps, d_rb, d_maxrows);
  count_launch();
  SB_CUDA(cudaGetLastError());
  SB_CUDA(cudaMemsetAsync(d_maxrows, 0, sizeof(int32_t), st));
  band_max_rows_kernel<<<1, 256, 0, st>>>(d_rb, nb, d_maxrows);
  count_launch();
  SB_CUDA(cudaGetLastError());
  int32_t h_maxrows = 0;
  SB_CUDA(cudaMemcpyAsync(&h_maxrows, d_maxrows, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  // ---- T4 band pointers (overlaps the small copy above) -------------------------------------------------
  if (nb > 1) {
    const size_t smem = sizeof(int32_t) * (2 * (SWEEP_TILE + 4) + static_cast<size_t>(nb) + 1);
    SB_CUDA(cudaFuncSetAttribute(band_ptr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    int64_t blocks = m->n_tiles;
    const int64_t cap = static_cast<int64_t>(m->sm_count) * 6;
    if (blocks > cap) blocks = cap;
    band_ptr_kernel<<<static_cast<unsigned>(blocks), BP_THREADS, smem, st>>>(
        m->d_i, m->d_p, m->d_plan, m->n_tiles, ncol, static_cast<int32_t>(nnz), d_rb, nb, d_bpt);
    count_launch();
    SB_CUDA(cudaGetLastError());
  }
  SB_CUDA(cudaStreamSynchronize(st));
  if (h_maxrows > rmax2) return fail(SB200_E_UNSUPPORTED, "transpose: a row band exceeds shared memory (internal bound violated)");
  // ---- T5 banded scatter ------------------------------------------------------------------------------------
  BandArgs a;
  a.i = m->d_i;
  a.p = m->d_p;
  a.x = m->d_x;
  a.ncol = ncol;
  a.nb = nb;
  a.rb = d_rb;
  a.bpt = d_bpt;
  a.p_out = d_p_out;
  a.i_out = d_i_out;
  a.x_out = d_x_out;
  a.max_rows = h_maxrows;
  const int launch_grid = nb < grid ? nb : grid;
  if (h_maxrows <= rmax4) {
    const size_t smem = sizeof(uint32_t) * static_cast<size_t>(h_maxrows) * (1 + 4);
    SB_CUDA(cudaFuncSetAttribute(transpose_band_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(smem)));
    transpose_band_kernel<4><<<launch_grid, TB_THREADS, smem, st>>>(a);
  } else {
    const size_t smem = sizeof(uint32_t) * static_cast<size_t>(h_maxrows) * (1 + 2);
    SB_CUDA(cudaFuncSetAttribute(transpose_band_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(smem)));
    transpose_band_kernel<2><<<launch_grid, TB_THREADS, smem, st>>>(a);
  }
  count_launch();
  SB_CUDA(cudaGetLastError());
  SB_CUDA(cudaStreamSynchronize(st));  // scratch is freed on return
  return SB200_OK;
}

}  // namespace sb200
