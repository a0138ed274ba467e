// Row segments of a local transpose pushed to the devices that own the rows (sharded transpose): shared by
// the cross-process exchange (exchange.cu: destinations inside peer-mapped windows) and the single-process sharding
// layer (sharded.cu: destinations are ordinary allocations of peer-accessible devices).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace sb200 {

constexpr int PUSH_MAX_RANKS = 16;

struct PushRowsSrc {
  const int32_t* p_loc;    // [nrow + 1] my local transpose: row pointer,
  const int32_t* cols;     // column ids (local to my column block),
  const double* vals;      // values
  const int64_t* dst_off;  // [nrow] where my segment of row r starts inside its owner's arrays
  int32_t nrow;
  int32_t col_offset;      // first column of my block
  int32_t rb[PUSH_MAX_RANKS + 1];  // rows [rb[q], rb[q+1]) belong to owner q
};

// A warp takes 32 consecutive rows and walks their concatenated segments 128 entries a step: consecutive lanes move
// consecutive entries of a segment (coalesced loads, coalesced peer stores).  Dest: int32_t* cols(int q), double* vals(int q).
template <class Dest>
__device__ __forceinline__ void push_rows_body(const PushRowsSrc& a, const int world, const Dest& dest) {
  constexpr int U = 4;
  const int lane = threadIdx.x & 31;
  const int64_t n_groups = (static_cast<int64_t>(a.nrow) + 31) / 32;
  const int64_t warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t w = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; w < n_groups; w += warps) {
    const int64_t r = w * 32 + lane;
    int32_t s = 0, len = 0;
    int q = 0;
    int64_t d0 = 0;
    if (r < a.nrow) {
      s = a.p_loc[r];
      len = a.p_loc[r + 1] - s;
      d0 = a.dst_off[r];
      while (q + 1 < world && a.rb[q + 1] <= r) ++q;
    }
    int32_t incl = len;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int32_t up = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += up;
    }
    const int32_t excl = incl - len;
    const int32_t tot = __shfl_sync(0xffffffffu, incl, 31);
    const uint32_t d0_lo = static_cast<uint32_t>(d0), d0_hi = static_cast<uint32_t>(static_cast<uint64_t>(d0) >> 32);
    for (int32_t base = 0; base < tot; base += 32 * U) {
      int32_t cc[U];
      double xx[U];
      int32_t* cdst[U];
      double* xdst[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int32_t t = base + u * 32 + lane;
        int l = 0;  // largest l with excl[l] <= t (skips empty segments)
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
          const int cand = l + step;
          const int32_t e = __shfl_sync(0xffffffffu, excl, cand & 31);
          if (cand < 32 && e <= t) l = cand;
        }
        const int32_t rs = __shfl_sync(0xffffffffu, s, l);
        const int32_t re = __shfl_sync(0xffffffffu, excl, l);
        const int ql = __shfl_sync(0xffffffffu, q, l);
        const int64_t dl = static_cast<int64_t>((static_cast<uint64_t>(__shfl_sync(0xffffffffu, d0_hi, l)) << 32) |
                                                 __shfl_sync(0xffffffffu, d0_lo, l));
        cdst[u] = nullptr;
        xdst[u] = nullptr;
        cc[u] = 0;
        xx[u] = 0.0;
        if (t < tot) {
          const int32_t j = t - re;
          cc[u] = a.cols[rs + j] + a.col_offset;
          xx[u] = a.vals[rs + j];
          cdst[u] = dest.cols(ql) + dl + j;
          xdst[u] = dest.vals(ql) + dl + j;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (cdst[u]) {
          *cdst[u] = cc[u];
          *xdst[u] = xx[u];
        }
      }
    }
  }
}

}  // namespace sb200
