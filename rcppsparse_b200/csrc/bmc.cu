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
// (row band, column): band b holds rows [b*bw, (b+1)*bw), bw <= 12288, and inside a band the entries keep their
// column-major order.  In that layout
//   * a band is a CSC matrix of its own: "virtual column" g = b*ncol + c is the run of column c inside band b,
//     vp[g] its first entry (one int32 per run), ri the row inside the band as uint16, x the value — 10 bytes
//     per entry instead of 12;
//   * a CTA keeps v[b*bw ..] (96 KB) in shared memory and streams the band's entries through 1-D bulk copies
//     (cp.async.bulk, SASS UBLKCP) on mbarriers, exactly like the column sweep; every gather is an LDS;
//   * a run's products are summed by a group of 4 / 8 / 32 lanes and leave as ONE red.global.add.f64 on
//     y[c] — 8 bytes per (column, band) at L2, nothing per entry.  No carries: a run cut by a tile boundary
//     simply sends two reductions.
// The work list is the merge path of (run ends, entries) of every band, cut into tiles of 1024 items, so
// power-law columns, empty runs and dense columns balance alike.  Four consumer groups of 128 threads per CTA
// share the operand slice and run their own two-stage rings (eight tiles in flight per SM), synchronising only
// inside the group.
//
// The layout is structure + values of the mirror, nothing about any result; it is built after the mirror has
// been asked for A^T v more than SB200_ROW_COMPANION_AFTER times (or on request, sb200_matrix_band_companion),
// costs one pass over the matrix and 10 B per entry of HBM, and is dropped by sb200_matrix_refresh_values.
//
// Roofline: HBM.  Algorithmic bytes of the op stay SURVEY.md 8(d)'s 12N + 4(n+1) + 8n + 8m; the kernel itself
// moves 10N + 4*nb*n + 8*nb*min(bw, m).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <new>
#include <string>
#include <vector>

#include "common.cuh"
#include "ptx.cuh"

namespace sb200 {

struct BandCompanion {
  int bw = 0;            // rows per band
  int nb = 0;            // bands
  int lanes = 8;         // lanes per run in the sweep (4, 8 or 32), from the mean run length
  int64_t n_tiles = 0;   // tiles over all bands
  int32_t* d_vp = nullptr;   // [nb*ncol + 1] first entry of every (band, column) run, band-major
  uint16_t* d_ri = nullptr;  // [nnz] row inside the band
  double* d_x = nullptr;     // [nnz]
  int64_t* d_ts = nullptr;   // [nb + 1] first tile of every band
  int32_t* d_plan = nullptr; // [n_tiles + nb] column coordinate of the band's merge path at every tile boundary
};

namespace {

constexpr int BS_GROUP = 128;  // threads per consumer group
constexpr int BS_GROUPS = 4;
constexpr int BS_THREADS = BS_GROUP * BS_GROUPS;
constexpr int BS_TILE = 1024;  // merge-path items per tile
constexpr int BS_STAGES = 2;
constexpr int BS_MAX_BAND_ROWS = 12288;
constexpr int BS_LONG_LIST = 16;

constexpr int BS_X_ELEMS = BS_TILE + 2;   // +1 align-down slack, +1 round-up
constexpr int BS_A_ELEMS = BS_TILE + 8;   // nc+1 values, +3 align-down, +3 round-up, +1 spare
constexpr int BS_R_ELEMS = BS_TILE + 16;  // +7 align-down, +7 round-up
constexpr size_t BS_X_BYTES = ((BS_X_ELEMS * 8 + 15) / 16) * 16;
constexpr size_t BS_A_BYTES = ((BS_A_ELEMS * 4 + 15) / 16) * 16;
constexpr size_t BS_R_BYTES = ((BS_R_ELEMS * 2 + 15) / 16) * 16;
constexpr size_t BS_STAGE_BYTES = BS_X_BYTES + BS_A_BYTES + BS_R_BYTES;

struct BsParams {
  const int32_t* vp;
  const uint16_t* ri;
  const double* x;
  const int64_t* ts;
  const int32_t* plan;
  const double* v;
  double* y;
  int32_t nrow, ncol;
  int nb, bw;
  int64_t n_tiles;
};

struct BsMeta {
  int32_t c0;     // column in progress at the tile's start
  int32_t nc;     // run ends inside the tile
  int32_t k0;     // first entry of the tile (global position in the band-major arrays)
  int32_t nk;     // entries inside the tile
  int32_t a_off;  // where vp[gbase + c0 + 1] sits in the staged window
  int32_t x_off;  // where x[k0] sits
  int32_t r_off;  // where ri[k0] sits
  int32_t pad;
};

__device__ __forceinline__ void group_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(BS_GROUP) : "memory"); }
// barrier over the group that also tells every thread whether any of them passed a true predicate
__device__ __forceinline__ bool group_sync_or(int id, bool pred) {
  uint32_t r;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.u32 q, %2, 0;\n\t"
      "bar.red.or.pred p, %1, %3, q;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(r)
      : "r"(id), "r"(static_cast<uint32_t>(pred ? 1 : 0)), "n"(BS_GROUP)
      : "memory");
  return r != 0;
}

// L lanes per run.  Runs longer than LONG_CAP entries are summed by the whole group afterwards.
template <int L>
__global__ void __launch_bounds__(BS_THREADS, 1) bandsweep_kernel(const BsParams prm) {
  constexpr int NG = BS_GROUP / L;  // runs a group sums at a time
  constexpr int LONG_CAP = (32 * L < 256) ? 32 * L : 256;
  extern __shared__ __align__(128) unsigned char bsm[];
  __shared__ uint64_t full_bar[BS_GROUPS][BS_STAGES];
  __shared__ BsMeta meta[BS_GROUPS][BS_STAGES];
  __shared__ int long_cnt[BS_GROUPS][BS_STAGES];
  __shared__ int long_list[BS_GROUPS][BS_STAGES][BS_LONG_LIST][3];

  const int tid = threadIdx.x, g = tid / BS_GROUP, tg = tid % BS_GROUP, lane = tid & 31;
  const int lg = tg / L, gl = tg % L;
  double* vs = reinterpret_cast<double*>(bsm);
  const size_t vs_bytes = (static_cast<size_t>(prm.bw) * 8 + 15) & ~static_cast<size_t>(15);
  unsigned char* my_stages = bsm + vs_bytes + static_cast<size_t>(g) * BS_STAGES * BS_STAGE_BYTES;

  if (tg == 0) {
    for (int s = 0; s < BS_STAGES; ++s) {
      ptx::mbar_init(&full_bar[g][s], 1);
      long_cnt[g][s] = 0;
    }
    ptx::fence_mbar_init();
  }
  __syncthreads();

  const int64_t t_begin = (prm.n_tiles * blockIdx.x) / gridDim.x;
  const int64_t t_end = (prm.n_tiles * (blockIdx.x + 1)) / gridDim.x;
  // band of my first tile: largest b with ts[b] <= t_begin
  int b = 0;
  {
    int lo = 0, hi = prm.nb - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (__ldg(prm.ts + mid) <= t_begin)
        lo = mid;
      else
        hi = mid - 1;
    }
    b = lo;
  }
  uint32_t n_used = 0;  // tiles this group has consumed: stage = n_used % STAGES, parity = (n_used / STAGES) & 1
  int64_t t = t_begin;
  while (t < t_end) {
    while (b + 1 < prm.nb && __ldg(prm.ts + b + 1) <= t) ++b;
    const int64_t ts_b = __ldg(prm.ts + b);
    int64_t t_hi = __ldg(prm.ts + b + 1);
    if (t_hi > t_end) t_hi = t_end;
    const int64_t gbase = static_cast<int64_t>(b) * prm.ncol;
    const int32_t Eb = __ldg(prm.vp + gbase);
    const int32_t Nb = __ldg(prm.vp + gbase + prm.ncol) - Eb;
    const int64_t items_b = static_cast<int64_t>(Nb) + prm.ncol;
    const int64_t row0 = static_cast<int64_t>(b) * prm.bw;
    int R = prm.bw;
    if (row0 + R > prm.nrow) R = static_cast<int>(prm.nrow - row0);

    __syncthreads();  // every group is done with the previous band's slice
    for (int r = tid; r < R; r += BS_THREADS) vs[r] = __ldg(prm.v + row0 + r);
    __syncthreads();

    // this group's tiles of the band: t + g, t + g + GROUPS, ...
    const int64_t first = t + g;
    const int n_my = first < t_hi ? static_cast<int>((t_hi - first + BS_GROUPS - 1) / BS_GROUPS) : 0;

    auto issue = [&](int idx, uint32_t seq) {
      const int s = static_cast<int>(seq % BS_STAGES);
      const int64_t j = first + static_cast<int64_t>(idx) * BS_GROUPS - ts_b;  // tile inside the band
      const int64_t pidx = ts_b + b + j;
      const int32_t c0 = __ldg(prm.plan + pidx), c1 = __ldg(prm.plan + pidx + 1);
      const int64_t d0 = j * BS_TILE;
      int64_t d1 = d0 + BS_TILE;
      if (d1 > items_b) d1 = items_b;
      const int32_t k0 = Eb + static_cast<int32_t>(d0 - c0), k1 = Eb + static_cast<int32_t>(d1 - c1);
      BsMeta mt;
      mt.c0 = c0;
      mt.nc = c1 - c0;
      mt.k0 = k0;
      mt.nk = k1 - k0;
      // run-end window: vp[gbase + c0 + 1 .. gbase + min(c1 + 1, ncol)]
      const int64_t a_first = gbase + c0 + 1;
      const int64_t a_last = gbase + ((c1 + 1 <= prm.ncol) ? c1 + 1 : prm.ncol);
      const int64_t a_al = a_first & ~static_cast<int64_t>(3);
      const int32_t a_cnt = (a_last >= a_first) ? static_cast<int32_t>(((a_last - a_al + 1) + 3) & ~static_cast<int64_t>(3)) : 0;
      mt.a_off = static_cast<int32_t>(a_first - a_al);
      const int32_t x_al = k0 & ~1;
      const int32_t x_cnt = (mt.nk > 0) ? (((k1 - x_al) + 1) & ~1) : 0;
      mt.x_off = k0 - x_al;
      const int32_t r_al = k0 & ~7;
      const int32_t r_cnt = (mt.nk > 0) ? (((k1 - r_al) + 7) & ~7) : 0;
      mt.r_off = k0 - r_al;
      mt.pad = 0;
      meta[g][s] = mt;
      unsigned char* st = my_stages + static_cast<size_t>(s) * BS_STAGE_BYTES;
      const uint32_t bytes = static_cast<uint32_t>(a_cnt) * 4u + static_cast<uint32_t>(x_cnt) * 8u + static_cast<uint32_t>(r_cnt) * 2u;
      ptx::mbar_arrive_expect_tx(&full_bar[g][s], bytes);
      if (x_cnt > 0) ptx::bulk_g2s(st, prm.x + x_al, static_cast<uint32_t>(x_cnt) * 8u, &full_bar[g][s]);
      if (a_cnt > 0) ptx::bulk_g2s(st + BS_X_BYTES, prm.vp + a_al, static_cast<uint32_t>(a_cnt) * 4u, &full_bar[g][s]);
      if (r_cnt > 0) ptx::bulk_g2s(st + BS_X_BYTES + BS_A_BYTES, prm.ri + r_al, static_cast<uint32_t>(r_cnt) * 2u, &full_bar[g][s]);
    };

    if (tg == 0) {
      for (int q = 0; q < BS_STAGES && q < n_my; ++q) issue(q, n_used + q);
    }
    __syncwarp();

    for (int idx = 0; idx < n_my; ++idx) {
      const uint32_t seq = n_used + idx;
      const int s = static_cast<int>(seq % BS_STAGES);
      ptx::mbar_wait(&full_bar[g][s], (seq / BS_STAGES) & 1u);
      const BsMeta mt = meta[g][s];
      const unsigned char* st = my_stages + static_cast<size_t>(s) * BS_STAGE_BYTES;
      const double* __restrict__ xs = reinterpret_cast<const double*>(st) + mt.x_off;
      const int32_t* __restrict__ as = reinterpret_cast<const int32_t*>(st + BS_X_BYTES) + mt.a_off;
      const uint16_t* __restrict__ rs = reinterpret_cast<const uint16_t*>(st + BS_X_BYTES + BS_A_BYTES) + mt.r_off;
      double* __restrict__ yb = prm.y + mt.c0;
      const int nseg = mt.nc + 1;  // runs c0 .. c0+nc-1 end in the tile, the last one stays open (may be empty)
      bool saw_long = false;
      for (int base = 0; base < nseg; base += NG) {  // trip count uniform over the group
        const int seg = base + lg;
        int beg = 0, end = 0;
        if (seg < nseg) {
          beg = (seg == 0) ? 0 : as[seg - 1] - mt.k0;
          end = (seg == mt.nc) ? mt.nk : as[seg] - mt.k0;
        }
        const int len = end - beg;
        const bool is_long = len > LONG_CAP;
        if (is_long) {
          if (gl == 0) {
            const int slot = atomicAdd(&long_cnt[g][s], 1);
            if (slot < BS_LONG_LIST) {
              long_list[g][s][slot][0] = beg;
              long_list[g][s][slot][1] = end;
              long_list[g][s][slot][2] = seg;
            }
          }
          saw_long = true;
          end = beg;
        }
        double a0 = 0.0, a1 = 0.0;
        int k = beg + gl;
        for (; k + L < end; k += 2 * L) {
          a0 = __fma_rn(xs[k], vs[rs[k]], a0);
          a1 = __fma_rn(xs[k + L], vs[rs[k + L]], a1);
        }
        if (k < end) a0 = __fma_rn(xs[k], vs[rs[k]], a0);
        double acc = __dadd_rn(a0, a1);
#pragma unroll
        for (int off = L / 2; off > 0; off >>= 1) acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, off));
        if (gl == 0 && len > 0 && !is_long) ptx::red_add_f64(yb + seg, acc);
      }
      if (group_sync_or(1 + g, saw_long)) {
        // rare: runs longer than LONG_CAP — all 128 threads on one run, one reduction per warp
        int nl = long_cnt[g][s];
        if (nl > BS_LONG_LIST) nl = BS_LONG_LIST;  // cannot happen: a tile holds at most TILE / LONG_CAP such runs
        for (int q = 0; q < nl; ++q) {
          const int beg = long_list[g][s][q][0], end = long_list[g][s][q][1], seg = long_list[g][s][q][2];
          double a0 = 0.0;
          for (int k = beg + tg; k < end; k += BS_GROUP) a0 = __fma_rn(xs[k], vs[rs[k]], a0);
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) a0 = __dadd_rn(a0, __shfl_xor_sync(0xffffffffu, a0, off));
          if (lane == 0) ptx::red_add_f64(yb + seg, a0);
        }
        group_sync(1 + g);
      }
      // every thread of the group is past its last read of stage s: refill it
      if (tg == 0) {
        long_cnt[g][s] = 0;
        if (idx + BS_STAGES < n_my) issue(idx + BS_STAGES, seq + BS_STAGES);
      }
      __syncwarp();
    }
    n_used += static_cast<uint32_t>(n_my);
    t = t_hi;
  }
}

// ---- build ---------------------------------------------------------------------------------------------------
// run lengths, band-major: cnt[b*ncol + c] = entries of column c with row in band b
__global__ void bmc_len_kernel(const int32_t* __restrict__ gp, const int32_t* __restrict__ bpt, int32_t ncol, int nb,
                               uint32_t* __restrict__ cnt) {
  const int64_t total = static_cast<int64_t>(nb) * ncol;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; g < total; g += stride) {
    const int b = static_cast<int>(g / ncol);
    const int64_t c = g - static_cast<int64_t>(b) * ncol;
    const int32_t s = (b == 0) ? __ldg(gp + c) : __ldg(bpt + static_cast<int64_t>(b - 1) * ncol + c);
    const int32_t e = (b == nb - 1) ? __ldg(gp + c + 1) : __ldg(bpt + static_cast<int64_t>(b) * ncol + c);
    cnt[g] = static_cast<uint32_t>(e - s);
  }
}

// Copy every run to its band-major place.  A warp takes 32 consecutive runs (consecutive destinations: its
// output is one contiguous stretch) and walks their concatenated entries 32 at a time, four steps in flight.
__global__ void __launch_bounds__(256) bmc_copy_kernel(const int32_t* __restrict__ gi, const int32_t* __restrict__ gp,
                                                      const double* __restrict__ gx, const int32_t* __restrict__ bpt,
                                                      const int32_t* __restrict__ vp, int32_t ncol, int nb, int bw,
                                                      uint16_t* __restrict__ ri, double* __restrict__ xb) {
  constexpr int U = 4;
  const int lane = threadIdx.x & 31;
  const int64_t total = static_cast<int64_t>(nb) * ncol;
  const int64_t n_groups = (total + 31) / 32;
  const int64_t warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t w = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; w < n_groups; w += warps) {
    const int64_t g = w * 32 + lane;
    int32_t s = 0, len = 0, row0 = 0, dst = 0;
    if (g < total) {
      const int b = static_cast<int>(g / ncol);
      const int64_t c = g - static_cast<int64_t>(b) * ncol;
      s = (b == 0) ? __ldg(gp + c) : __ldg(bpt + static_cast<int64_t>(b - 1) * ncol + c);
      dst = __ldg(vp + g);
      len = __ldg(vp + g + 1) - dst;
      row0 = b * bw;
    }
    const int32_t dst0 = __shfl_sync(0xffffffffu, dst, 0);
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
        rr[u] = 0;
        xx[u] = 0.0;
        if (q < tot) {
          const int32_t k = rs + (q - re);
          rr[u] = ptx::ld_stream_s32(gi + k) - r0;
          xx[u] = ptx::ld_stream_f64(gx + k);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int32_t q = base + u * 32 + lane;
        if (q < tot) {
          ri[dst0 + q] = static_cast<uint16_t>(rr[u]);
          xb[dst0 + q] = xx[u];
        }
      }
    }
  }
}

// tiles per band, then their running sum (one block; nb is small)
__global__ void bmc_tiles_kernel(const int32_t* __restrict__ vp, int32_t ncol, int nb, int64_t* __restrict__ ts) {
  for (int b = threadIdx.x; b < nb; b += blockDim.x) {
    const int64_t gb = static_cast<int64_t>(b) * ncol;
    const int64_t items = static_cast<int64_t>(__ldg(vp + gb + ncol) - __ldg(vp + gb)) + ncol;
    ts[b + 1] = (items + BS_TILE - 1) / BS_TILE;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int64_t run = 0;
    ts[0] = 0;
    for (int b = 0; b < nb; ++b) {
      run += ts[b + 1];
      ts[b + 1] = run;
    }
  }
}

// plan[ts[b] + b + j] = run ends of band b before diagonal j*TILE of its merge path (ends win ties), j = 0..T_b
__global__ void bmc_plan_kernel(const int32_t* __restrict__ vp, const int64_t* __restrict__ ts, int nb, int32_t ncol,
                                int64_t n_plan, int32_t* __restrict__ plan) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n_plan) return;
  int lo_b = 0, hi_b = nb - 1;  // largest b with ts[b] + b <= idx
  while (lo_b < hi_b) {
    const int mid = (lo_b + hi_b + 1) >> 1;
    if (ts[mid] + mid <= idx)
      lo_b = mid;
    else
      hi_b = mid - 1;
  }
  const int b = lo_b;
  const int64_t j = idx - ts[b] - b;
  const int64_t gb = static_cast<int64_t>(b) * ncol;
  const int32_t Eb = vp[gb];
  const int64_t Nb = static_cast<int64_t>(vp[gb + ncol]) - Eb;
  int64_t d = j * BS_TILE;
  if (d > Nb + ncol) d = Nb + ncol;
  int64_t lo = d > Nb ? d - Nb : 0;
  int64_t hi = d < ncol ? d : ncol;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (static_cast<int64_t>(vp[gb + mid + 1]) - Eb <= d - mid - 1)
      lo = mid + 1;
    else
      hi = mid;
  }
  plan[idx] = static_cast<int32_t>(lo);
}

void free_companion(BandCompanion* bc, cudaStream_t s) {
  if (!bc) return;
  pool_free(bc->d_vp, s);
  pool_free(bc->d_ri, s);
  pool_free(bc->d_x, s);
  pool_free(bc->d_ts, s);
  pool_free(bc->d_plan, s);
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
  const double mean_run = static_cast<double>(nnz) / static_cast<double>(G > 0 ? G : 1);
  bc->lanes = mean_run <= 12.0 ? 4 : (mean_run <= 96.0 ? 8 : 32);
  int32_t* d_rb = nullptr;
  int32_t* d_bpt = nullptr;
  uint32_t* d_cnt = nullptr;
  void* d_scan_ws = nullptr;
  struct Guard {
    BandCompanion*& bc;
    int32_t*& rb;
    int32_t*& bpt;
    uint32_t*& cnt;
    void*& ws;
    cudaStream_t s;
    bool armed = true;
    ~Guard() {
      pool_free(rb, s);
      pool_free(bpt, s);
      pool_free(cnt, s);
      pool_free(ws, s);
      if (armed) free_companion(bc, s);
    }
  } guard{bc, d_rb, d_bpt, d_cnt, d_scan_ws, st};

  SB_TRY(pool_alloc(reinterpret_cast<void**>(&bc->d_vp), padded_bytes(sizeof(int32_t) * static_cast<size_t>(G + 1)), st));
  SB_TRY(pool_alloc(reinterpret_cast<void**>(&bc->d_ri), padded_bytes(sizeof(uint16_t) * static_cast<size_t>(nnz)), st));
  SB_TRY(pool_alloc(reinterpret_cast<void**>(&bc->d_x), padded_bytes(sizeof(double) * static_cast<size_t>(nnz)), st));
  SB_TRY(pool_alloc(reinterpret_cast<void**>(&bc->d_ts), sizeof(int64_t) * (static_cast<size_t>(nb) + 1), st));
  SB_TRY(pool_alloc(reinterpret_cast<void**>(&d_cnt), sizeof(uint32_t) * static_cast<size_t>(G > 0 ? G : 1), st));
  const size_t scan_ws = scan_workspace_bytes(G);
  SB_TRY(pool_alloc(&d_scan_ws, scan_ws, st));
  if (nb > 1) {
    std::vector<int32_t> rb(static_cast<size_t>(nb) + 1);
    for (int b = 0; b <= nb; ++b) {
      const int64_t r = static_cast<int64_t>(b) * bw;
      rb[b] = static_cast<int32_t>(r < nrow ? r : nrow);
    }
    SB_TRY(pool_alloc(reinterpret_cast<void**>(&d_rb), sizeof(int32_t) * rb.size(), st));
    SB_TRY(pool_alloc(reinterpret_cast<void**>(&d_bpt), sizeof(int32_t) * static_cast<size_t>(nb - 1) * static_cast<size_t>(ncol), st));
    SB_CUDA(cudaMemcpyAsync(d_rb, rb.data(), sizeof(int32_t) * rb.size(), cudaMemcpyHostToDevice, st));
    SB_CUDA(cudaStreamSynchronize(st));  // rb is a host temporary
    SB_TRY(launch_band_ptr(m, d_rb, nb, d_bpt));
  }
  const int64_t cap_blocks = static_cast<int64_t>(m->sm_count) * 16;
  {
    int64_t blocks = (G + 255) / 256;
    if (blocks > cap_blocks) blocks = cap_blocks;
    if (blocks < 1) blocks = 1;
    bmc_len_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(m->d_p, d_bpt, ncol, nb, d_cnt);
    count_launch();
    SB_CUDA(cudaGetLastError());
  }
  SB_TRY(exclusive_scan_u32(st, d_cnt, bc->d_vp, G, nullptr, d_scan_ws, scan_ws));
  {
    int64_t blocks = ((G + 31) / 32 + 7) / 8;  // 8 warps per block
    if (blocks > cap_blocks) blocks = cap_blocks;
    if (blocks < 1) blocks = 1;
    bmc_copy_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(m->d_i, m->d_p, m->d_x, d_bpt, bc->d_vp, ncol, nb, bw, bc->d_ri,
                                                                   bc->d_x);
    count_launch();
    SB_CUDA(cudaGetLastError());
  }
  bmc_tiles_kernel<<<1, 256, 0, st>>>(bc->d_vp, ncol, nb, bc->d_ts);
  count_launch();
  SB_CUDA(cudaGetLastError());
  int64_t n_tiles = 0;
  SB_CUDA(cudaMemcpyAsync(&n_tiles, bc->d_ts + nb, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  bc->n_tiles = n_tiles;
  const int64_t n_plan = n_tiles + nb;
  SB_TRY(pool_alloc(reinterpret_cast<void**>(&bc->d_plan), sizeof(int32_t) * static_cast<size_t>(n_plan + 2), st));
  bmc_plan_kernel<<<static_cast<unsigned>((n_plan + 255) / 256), 256, 0, st>>>(bc->d_vp, bc->d_ts, nb, ncol, n_plan, bc->d_plan);
  count_launch();
  SB_CUDA(cudaGetLastError());
  SB_CUDA(cudaStreamSynchronize(st));
  guard.armed = false;
  *out = bc;
  return SB200_OK;
}

template <int L>
int launch_bandsweep_t(const sb200_matrix* m, const BsParams& prm, size_t smem) {
  auto kern = bandsweep_kernel<L>;
  SB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  int64_t grid = m->sm_count;
  if (grid > prm.n_tiles) grid = prm.n_tiles;
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

int launch_bandsweep(sb200_matrix* m, const double* d_v, double* d_out) {
  const BandCompanion* bc = m->bmc;
  if (!bc) return fail(SB200_E_INVALID, "no band-major companion on this mirror");
  if (m->ncol > 0) SB_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * static_cast<size_t>(m->ncol), m->stream));
  if (bc->n_tiles == 0) return SB200_OK;
  BsParams prm;
  prm.vp = bc->d_vp;
  prm.ri = bc->d_ri;
  prm.x = bc->d_x;
  prm.ts = bc->d_ts;
  prm.plan = bc->d_plan;
  prm.v = d_v;
  prm.y = d_out;
  prm.nrow = m->nrow;
  prm.ncol = m->ncol;
  prm.nb = bc->nb;
  prm.bw = bc->bw;
  prm.n_tiles = bc->n_tiles;
  const size_t smem = ((static_cast<size_t>(bc->bw) * 8 + 15) & ~static_cast<size_t>(15)) + BS_GROUPS * BS_STAGES * BS_STAGE_BYTES;
  int lanes = bc->lanes;
  if (const char* e = getenv("SB200_BS_LANES")) {
    const int v = atoi(e);
    if (v == 4 || v == 8 || v == 32) lanes = v;
  }
  switch (lanes) {
    case 4: return launch_bandsweep_t<4>(m, prm, smem);
    case 32: return launch_bandsweep_t<32>(m, prm, smem);
    default: return launch_bandsweep_t<8>(m, prm, smem);
  }
}

}  // namespace sb200
