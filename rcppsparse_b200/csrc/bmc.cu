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
//   * a run is cut into interleaved pieces of <= 16 entries, a THREAD sums a piece (~7 instructions per entry)
//     and sends ONE red.global.add.f64 to y[c] — 8 bytes per piece at L2, nothing per entry.  No carries: a
//     run cut by a tile boundary simply sends two reductions.
// The work list is the merge path of (run ends, entries) of every band, cut into tiles of 1024 items, so
// power-law columns, empty runs and dense columns balance alike.  Four consumer groups of 128 threads per CTA
// share the operand slice; each has its own two-stage ring driven by its own producer warp (eight tiles in
// flight or in work per SM) and synchronises only inside the group.
//
// The layout is structure + values of the mirror, nothing about any result; it is built after the mirror has
// been asked for A^T v more than SB200_ROW_COMPANION_AFTER times (or on request, sb200_matrix_band_companion),
// costs one pass over the matrix and 10 B per entry of HBM, and is dropped by sb200_matrix_refresh_values.
//
// Roofline: HBM.  Algorithmic bytes of the op stay SURVEY.md 8(d)'s 12N + 4(n+1) + 8n + 8m; the kernel itself
// moves 10N + 4*nb*n + 8*nb*min(bw, m).
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
  int bw = 0;            // rows per band
  int nb = 0;            // bands
  double mean_run = 0.0; // entries per (band, column) run
  int tile = 1024;       // merge-path items per tile (the tile plan below is cut for it)
  int stages = 2;        // ring depth per consumer group in the sweep
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
constexpr int BS_CONSUMERS = BS_GROUP * BS_GROUPS;
constexpr int BS_THREADS = BS_CONSUMERS + 32 * BS_GROUPS;  // + one producer warp per group (its lane 0 drives the ring)
constexpr int BS_MAX_BAND_ROWS = 12288;  // 96 KB of operand per band
constexpr int BS_LONG_LIST = 36;         // a tile holds at most TILE / (4 * 8 + 1) runs of more than 4 pieces
constexpr int BS_AHEAD = 0;  // entries the optional L2 prefetch runs ahead of the bulk copies (SB200_BS_AHEAD)

// Ring geometry: TILE merge-path items per tile, STAGES tiles per consumer group.  The tile size is part of the
// companion (its tile plan); both are picked at build time (default 1024 x 2; SB200_BS_CFG=tile,stages for tuning).
template <int TILE>
struct BsGeom {
  static constexpr int X_ELEMS = TILE + 2;   // +1 align-down slack, +1 round-up
  static constexpr int A_ELEMS = TILE + 8;   // nc+1 values, +3 align-down, +3 round-up, +1 spare
  static constexpr int R_ELEMS = TILE + 16;  // +7 align-down, +7 round-up
  static constexpr size_t X_BYTES = ((X_ELEMS * 8 + 15) / 16) * 16;
  static constexpr size_t A_BYTES = ((A_ELEMS * 4 + 15) / 16) * 16;
  static constexpr size_t R_BYTES = ((R_ELEMS * 2 + 15) / 16) * 16;
  static constexpr size_t STAGE_BYTES = X_BYTES + A_BYTES + R_BYTES;
};

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
  int64_t nnz;     // entries in ri / x
  int64_t n_runs;  // nb * ncol (vp has n_runs + 1 entries)
  int ahead;       // entries the L2 prefetch runs ahead of the bulk copies (0 = no prefetch)
  int cap_shift;   // a thread sums at most 1 << cap_shift entries (3 or 4)
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
// try_wait with a suspend-time hint: the thread sleeps in hardware (no issue slots) until the phase completes or
// the hint (ns) runs out
__device__ __forceinline__ bool mbar_try_wait_suspend(uint64_t* bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(ptx::smem_addr(bar)), "r"(parity), "r"(hint_ns)
      : "memory");
  return ok != 0;
}

// L2 prefetch of a byte range (cp.async.bulk.prefetch.L2: no shared memory, no completion to wait for)
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// One tile, one consumer group of 128 threads:
//   a THREAD per run: up to cap (16) entries are summed on the spot — x and the row id from the stage, the operand
//   from the band slice, one FMA per entry, one red.global.add.f64 on y[column] per run, nothing else per run but
//   its two bounds.  A run of up to 2 / 4 cap entries is cut into 2 / 4 interleaved PIECES (piece p takes entries
//   p, p + np, ...: neighbouring lanes read neighbouring entries): its thread sums piece 0 and queues the others as
//   packed 32-bit descriptors, summed a thread per piece after the group's barrier.  Longer runs go to a short
//   list and are summed by all 128 threads together, one reduction per warp.
// (Round 2's first version summed a run with a group of 4-32 lanes: ~100 instructions per entry at 12 entries per
// run — shuffles, bounds and loop overhead per run per lane; ncu: 317 M warp instructions for 1e8 entries.)
template <int TILE, int STAGES>
__global__ void __launch_bounds__(BS_THREADS, 1) bandsweep_kernel(const BsParams prm) {
  using G = BsGeom<TILE>;
  constexpr int PIECES = TILE / 2 + TILE / 8 + 8;  // non-empty runs of a tile + the extra pieces of split runs (<= nk / 8)
  extern __shared__ __align__(128) unsigned char bsm[];
  __shared__ uint64_t full_bar[BS_GROUPS][STAGES];   // producer -> group: the stage's bytes have landed
  __shared__ uint64_t empty_bar[BS_GROUPS][STAGES];  // group -> producer: every thread is done with the stage
  __shared__ BsMeta meta[BS_GROUPS][STAGES];
  __shared__ int piece_cnt[BS_GROUPS][2];  // by tile parity
  __shared__ int long_cnt[BS_GROUPS][2];
  __shared__ int long_list[BS_GROUPS][2][BS_LONG_LIST][3];  // runs of more than 4 pieces: (first entry, length, run)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool producer = tid >= BS_CONSUMERS;
  const int g = producer ? warp - BS_CONSUMERS / 32 : tid / BS_GROUP;
  const int tg = tid % BS_GROUP;
  double* vs = reinterpret_cast<double*>(bsm);
  const size_t vs_bytes = (static_cast<size_t>(prm.bw) * 8 + 15) & ~static_cast<size_t>(15);
  unsigned char* my_stages = bsm + vs_bytes + static_cast<size_t>(g) * STAGES * G::STAGE_BYTES;
  uint32_t* pieces = reinterpret_cast<uint32_t*>(bsm + vs_bytes + static_cast<size_t>(BS_GROUPS) * STAGES * G::STAGE_BYTES) + g * PIECES;
  const int cap_shift = prm.cap_shift, cap = 1 << cap_shift;

  if (tid < BS_GROUPS) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[tid][s], 1);
      ptx::mbar_init(&empty_bar[tid][s], 1);
    }
    piece_cnt[tid][0] = piece_cnt[tid][1] = 0;
    long_cnt[tid][0] = long_cnt[tid][1] = 0;
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
  uint32_t n_used = 0;  // tiles this group has been through: stage = n_used % STAGES, parity = (n_used / STAGES) & 1
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

    if (producer) {
      // ===================== producer warp of group g: lane 0 runs the group's ring ==========================
      if (lane == 0) {
        int64_t pidx = ts_b + b + (first - ts_b);  // plan entry of the tile being issued; += GROUPS per tile
        int32_t c0 = 0, c1 = 0;
        if (n_my > 0) {
          c0 = __ldg(prm.plan + pidx);
          c1 = __ldg(prm.plan + pidx + 1);
        }
        for (int idx = 0; idx < n_my; ++idx) {
          const uint32_t seq = n_used + idx;
          const int s = static_cast<int>(seq % STAGES);
          // next tile's plan entries: in flight while this tile is issued (and while the stage is waited for)
          int32_t nc0 = 0, nc1 = 0;
          if (idx + 1 < n_my) {
            nc0 = __ldg(prm.plan + pidx + BS_GROUPS);
            nc1 = __ldg(prm.plan + pidx + BS_GROUPS + 1);
          }
          const int64_t j = first + static_cast<int64_t>(idx) * BS_GROUPS - ts_b;  // tile inside the band
          const int64_t d0 = j * TILE;
          int64_t d1 = d0 + TILE;
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
          if (seq >= STAGES) {  // the stage is still being read: sleep in hardware until the group hands it back
            while (!mbar_try_wait_suspend(&empty_bar[g][s], ((seq / STAGES) - 1u) & 1u, 1000000u)) {
            }
          }
          meta[g][s] = mt;
          unsigned char* st = my_stages + static_cast<size_t>(s) * G::STAGE_BYTES;
          const uint32_t bytes = static_cast<uint32_t>(a_cnt) * 4u + static_cast<uint32_t>(x_cnt) * 8u + static_cast<uint32_t>(r_cnt) * 2u;
          ptx::mbar_arrive_expect_tx(&full_bar[g][s], bytes);
          if (x_cnt > 0) ptx::bulk_g2s(st, prm.x + x_al, static_cast<uint32_t>(x_cnt) * 8u, &full_bar[g][s]);
          if (a_cnt > 0) ptx::bulk_g2s(st + G::X_BYTES, prm.vp + a_al, static_cast<uint32_t>(a_cnt) * 4u, &full_bar[g][s]);
          if (r_cnt > 0) ptx::bulk_g2s(st + G::X_BYTES + G::A_BYTES, prm.ri + r_al, static_cast<uint32_t>(r_cnt) * 2u, &full_bar[g][s]);
          if (prm.ahead > 0) {
            // optional: pull the stretch `ahead` entries past this tile from HBM into L2 (the band's entries and
            // run ends are linear streams; the four producers cover them between them)
            int64_t pk0 = (static_cast<int64_t>(k1) + prm.ahead) & ~static_cast<int64_t>(7);
            int64_t pk1 = (pk0 + mt.nk + (mt.nk >> 2) + 15) & ~static_cast<int64_t>(7);
            if (pk1 > prm.nnz) pk1 = prm.nnz & ~static_cast<int64_t>(7);
            if (pk1 > pk0) {
              prefetch_l2_bulk(prm.x + pk0, static_cast<uint32_t>(pk1 - pk0) * 8u);
              prefetch_l2_bulk(prm.ri + pk0, static_cast<uint32_t>(pk1 - pk0) * 2u);
            }
            const int64_t ahead_runs = (mt.nk > 0) ? (static_cast<int64_t>(prm.ahead) * (mt.nc + 1)) / mt.nk : static_cast<int64_t>(prm.ahead);
            int64_t pa0 = (a_last + ahead_runs) & ~static_cast<int64_t>(3);
            int64_t pa1 = (pa0 + mt.nc + (mt.nc >> 2) + 11) & ~static_cast<int64_t>(3);
            if (pa1 > prm.n_runs) pa1 = prm.n_runs & ~static_cast<int64_t>(3);
            if (pa1 > pa0) prefetch_l2_bulk(prm.vp + pa0, static_cast<uint32_t>(pa1 - pa0) * 4u);
          }
          c0 = nc0;
          c1 = nc1;
          pidx += BS_GROUPS;
        }
      }
      __syncwarp();
    } else {
      // ======================================= consumer group g ==============================================
      for (int idx = 0; idx < n_my; ++idx) {
        const uint32_t seq = n_used + idx;
        const int s = static_cast<int>(seq % STAGES);
        const int par = static_cast<int>(seq & 1u);
        ptx::mbar_wait(&full_bar[g][s], (seq / STAGES) & 1u);
        const BsMeta mt = meta[g][s];
        const unsigned char* st = my_stages + static_cast<size_t>(s) * G::STAGE_BYTES;
        const double* __restrict__ xs = reinterpret_cast<const double*>(st) + mt.x_off;
        const int32_t* __restrict__ as = reinterpret_cast<const int32_t*>(st + G::X_BYTES) + mt.a_off;
        const uint16_t* __restrict__ rs = reinterpret_cast<const uint16_t*>(st + G::X_BYTES + G::A_BYTES) + mt.r_off;
        double* __restrict__ yb = prm.y + mt.c0;
        const int nseg = mt.nc + 1;  // runs c0 .. c0+nc-1 end in the tile, the last one stays open (may be empty)
        // ---- a thread per run: runs of up to `cap` entries are summed on the spot by their thread; a run of up to
        //      2 / 4 cap is cut into 2 / 4 interleaved pieces — the thread sums piece 0 and queues the others;
        //      longer runs go to the long list.  Nothing per run but its two bounds and one reduction. -------------
        for (int base = 0; base < nseg; base += BS_GROUP) {  // trip count uniform over the group
          const int seg = base + tg;
          int beg = 0, len = 0;
          if (seg < nseg) {
            beg = (seg == 0) ? 0 : as[seg - 1] - mt.k0;
            len = ((seg == mt.nc) ? mt.nk : as[seg] - mt.k0) - beg;
          }
          if (len > (cap << 2)) {
            const int slot = atomicAdd(&long_cnt[g][par], 1);
            long_list[g][par][slot][0] = beg;
            long_list[g][par][slot][1] = len;
            long_list[g][par][slot][2] = seg;
          } else if (len > 0) {
            const int lg = (len > (cap << 1)) ? 2 : ((len > cap) ? 1 : 0);
            const int np = 1 << lg;
            if (lg) {
              uint32_t* dst = pieces + atomicAdd(&piece_cnt[g][par], np - 1);
              const uint32_t common = (static_cast<uint32_t>(lg) << 10) | (static_cast<uint32_t>(seg) << 21);
              for (int p = 1; p < np; ++p)
                dst[p - 1] = static_cast<uint32_t>(beg + p) | common | (static_cast<uint32_t>((len - 1 - p) >> lg) << 17);  // count - 1
            }
            int cnt = ((len - 1) >> lg) + 1;
            int k = beg;
            double a0 = 0.0, a1 = 0.0;
            for (; cnt >= 2; cnt -= 2, k += 2 * np) {
              a0 = __fma_rn(xs[k], vs[rs[k]], a0);
              a1 = __fma_rn(xs[k + np], vs[rs[k + np]], a1);
            }
            if (cnt) a0 = __fma_rn(xs[k], vs[rs[k]], a0);
            ptx::red_add_f64(yb + seg, __dadd_rn(a0, a1));
          }
        }
        group_sync(1 + g);
        // ---- the queued pieces: a thread per piece ------------------------------------------------------------
        const int P = piece_cnt[g][par];
        for (int w = tg; w < P; w += BS_GROUP) {
          const uint32_t d = pieces[w];
          int k = static_cast<int>(d & 1023u);
          const int np = 1 << ((d >> 10) & 7u);
          int cnt = static_cast<int>((d >> 17) & 15u) + 1;
          const int seg = static_cast<int>(d >> 21);
          double a0 = 0.0, a1 = 0.0;
          for (; cnt >= 2; cnt -= 2, k += 2 * np) {
            a0 = __fma_rn(xs[k], vs[rs[k]], a0);
            a1 = __fma_rn(xs[k + np], vs[rs[k + np]], a1);
          }
          if (cnt) a0 = __fma_rn(xs[k], vs[rs[k]], a0);
          ptx::red_add_f64(yb + seg, __dadd_rn(a0, a1));
        }
        // ---- long runs (more than 4 pieces): all 128 threads on one run, one reduction per warp ---------------
        const int nl = long_cnt[g][par];
        for (int q = 0; q < nl; ++q) {
          const int beg = long_list[g][par][q][0], end = beg + long_list[g][par][q][1], seg = long_list[g][par][q][2];
          double a0 = 0.0, a1 = 0.0;
          int k = beg + tg;
          for (; k + BS_GROUP < end; k += 2 * BS_GROUP) {
            a0 = __fma_rn(xs[k], vs[rs[k]], a0);
            a1 = __fma_rn(xs[k + BS_GROUP], vs[rs[k + BS_GROUP]], a1);
          }
          if (k < end) a0 = __fma_rn(xs[k], vs[rs[k]], a0);
          a0 = __dadd_rn(a0, a1);
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) a0 = __dadd_rn(a0, __shfl_xor_sync(0xffffffffu, a0, off));
          if (lane == 0) ptx::red_add_f64(yb + seg, a0);
        }
        group_sync(1 + g);  // every thread of the group is past its last read of stage s and of the piece list
        if (tg == 0) {
          piece_cnt[g][par] = 0;
          long_cnt[g][par] = 0;
          ptx::mbar_arrive(&empty_bar[g][s]);
        }
      }
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
__global__ void bmc_tiles_kernel(const int32_t* __restrict__ vp, int32_t ncol, int nb, int tile, int64_t* __restrict__ ts) {
  for (int b = threadIdx.x; b < nb; b += blockDim.x) {
    const int64_t gb = static_cast<int64_t>(b) * ncol;
    const int64_t items = static_cast<int64_t>(__ldg(vp + gb + ncol) - __ldg(vp + gb)) + ncol;
    ts[b + 1] = (items + tile - 1) / tile;
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
                                int tile, int64_t n_plan, int32_t* __restrict__ plan) {
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
  int64_t d = j * tile;
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
  bc->mean_run = mean_run;
  if (const char* e = getenv("SB200_BS_CFG")) {  // tuning: tile,stages out of the instantiated set
    int tl = 0, sg = 0;
    if (sscanf(e, "%d,%d", &tl, &sg) == 2 && ((tl == 1024 && sg == 2) || (tl == 640 && sg == 3) || (tl == 512 && sg == 4))) {
      bc->tile = tl;
      bc->stages = sg;
    }
  }
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
  bmc_tiles_kernel<<<1, 256, 0, st>>>(bc->d_vp, ncol, nb, bc->tile, bc->d_ts);
  count_launch();
  SB_CUDA(cudaGetLastError());
  int64_t n_tiles = 0;
  SB_CUDA(cudaMemcpyAsync(&n_tiles, bc->d_ts + nb, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  bc->n_tiles = n_tiles;
  const int64_t n_plan = n_tiles + nb;
  SB_TRY(pool_alloc(reinterpret_cast<void**>(&bc->d_plan), sizeof(int32_t) * static_cast<size_t>(n_plan + 2), st));
  bmc_plan_kernel<<<static_cast<unsigned>((n_plan + 255) / 256), 256, 0, st>>>(bc->d_vp, bc->d_ts, nb, ncol, bc->tile, n_plan, bc->d_plan);
  count_launch();
  SB_CUDA(cudaGetLastError());
  SB_CUDA(cudaStreamSynchronize(st));
  guard.armed = false;
  *out = bc;
  return SB200_OK;
}

template <int TILE, int STAGES>
int launch_bandsweep_t(const sb200_matrix* m, const BsParams& prm) {
  constexpr int PIECES = TILE / 2 + TILE / 8 + 8;
  const size_t smem = ((static_cast<size_t>(prm.bw) * 8 + 15) & ~static_cast<size_t>(15)) + BS_GROUPS * STAGES * BsGeom<TILE>::STAGE_BYTES +
                      BS_GROUPS * PIECES * sizeof(uint32_t);
  auto kern = bandsweep_kernel<TILE, STAGES>;
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
  prm.nnz = m->nnz;
  prm.n_runs = static_cast<int64_t>(bc->nb) * m->ncol;
  prm.ahead = BS_AHEAD;
  if (const char* e = getenv("SB200_BS_AHEAD")) {
    const int v = atoi(e);
    if (v >= 0 && v <= (1 << 22)) prm.ahead = v;
  }
  // entries per thread: 16 when runs are short (one piece per run, half as many reductions), 8 when they are
  // long (a 1024-entry tile then cuts into ~128 pieces: one per thread)
  prm.cap_shift = bc->mean_run >= 48.0 ? 3 : 4;
  if (const char* e = getenv("SB200_BS_CAP")) {
    const int v = atoi(e);
    if (v == 8) prm.cap_shift = 3;
    if (v == 16) prm.cap_shift = 4;
  }
  if (bc->tile == 640 && bc->stages == 3) return launch_bandsweep_t<640, 3>(m, prm);
  if (bc->tile == 512 && bc->stages == 4) return launch_bandsweep_t<512, 4>(m, prm);
  return launch_bandsweep_t<1024, 2>(m, prm);
}

}  // namespace sb200
