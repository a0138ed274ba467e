// Warp-cooperative walk over 32 runs of entries (shared by the banded kernels in bands.cu and the band-major
// companion builder in bmc.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace sb200 {

// One warp, 32 runs (start s, length len per lane).  The concatenated entries are walked 32*U at a
// time in lane order = (run, position) order: for every group of U steps first load(k, l, valid) is
// called U times (all loads of the group are in flight together — the kernels are latency-bound
// otherwise), then use(payload) U times in step order.  k = entry index, l = lane owning the run.
// Every lane calls load/use in every step (valid = false past the end), so they may use warp-wide
// primitives.
template <int U, typename Payload, typename Load, typename Use>
__device__ __forceinline__ void warp_walk_runs(int32_t s, int32_t len, int lane, Load load, Use use) {
  int32_t incl = len;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int32_t up = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += up;
  }
  const int32_t excl = incl - len;
  const int32_t total = __shfl_sync(0xffffffffu, incl, 31);
  for (int32_t base = 0; base < total; base += 32 * U) {
    Payload pl[U];
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
      pl[u] = load(rs + (q - re), l, q < total);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (base + u * 32 < total) use(pl[u]);  // warp-uniform
    }
  }
}

}  // namespace sb200
