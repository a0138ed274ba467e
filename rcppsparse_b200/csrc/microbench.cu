// Stand-alone micro-benchmarks that size the design decisions of the sweeps (not part of the library):
//   read-only stream, copy, random FP64 RED into an L2-resident vector, coalesced RED, random 8-byte
//   gather from an L2-resident vector, shared-memory FP64 atomics, u32 RED, strided short runs.
// Usage: microbench [n_elements=100000000] [target_rows=1000000]     prints one JSON line per test.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>
#include <stdlib.h>

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess) {                                                                   \
      fprintf(stderr, "CUDA %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e_)); \
      exit(1);                                                                                 \
    }                                                                                          \
  } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__global__ void init_kernel(int32_t* idx, double* val, int64_t n, int32_t rows, int sorted_runs) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
    idx[k] = (int32_t)(mix64(k) % (uint64_t)rows);
    val[k] = (double)((int)(mix64(k ^ 0x55) & 1023) - 512) / 64.0;
  }
}

__global__ void read_kernel(const double2* __restrict__ x, int64_t n2, double* out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double acc = 0;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n2; k += stride) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(x + k));
    acc += v.x + v.y;
  }
  if (acc == 1.2345e-300) out[0] = acc;
}

__global__ void copy_kernel(const double2* __restrict__ x, double2* __restrict__ y, int64_t n2) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n2; k += stride) y[k] = x[k];
}

// mode 0: random rows (idx), mode 1: coalesced (row = k % rows)
template <int MODE>
__global__ void red_f64_kernel(const int4* __restrict__ idx, const double2* __restrict__ val, int64_t n4, int32_t rows,
                               double* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n4; g += stride) {
    int4 r = idx[g];
    const double2 a = val[2 * g], b = val[2 * g + 1];
    if (MODE == 1) {
      const int32_t base = (int32_t)((4 * g) % rows);
      r.x = base; r.y = (base + 1) % rows; r.z = (base + 2) % rows; r.w = (base + 3) % rows;
    }
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(out + r.x), "d"(a.x) : "memory");
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(out + r.y), "d"(a.y) : "memory");
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(out + r.z), "d"(b.x) : "memory");
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(out + r.w), "d"(b.y) : "memory");
  }
}

// coalesced-by-lane RED: consecutive lanes hit consecutive rows (the flush pattern of a privatised band)
__global__ void red_f64_lane_kernel(const double* __restrict__ val, int64_t n, int32_t rows, double* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride)
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(out + (k % rows)), "d"(val[k]) : "memory");
}

__global__ void red_u32_kernel(const int4* __restrict__ idx, int64_t n4, uint32_t* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n4; g += stride) {
    const int4 r = idx[g];
    asm volatile("red.global.add.u32 [%0], 1;" ::"l"(out + r.x) : "memory");
    asm volatile("red.global.add.u32 [%0], 1;" ::"l"(out + r.y) : "memory");
    asm volatile("red.global.add.u32 [%0], 1;" ::"l"(out + r.z) : "memory");
    asm volatile("red.global.add.u32 [%0], 1;" ::"l"(out + r.w) : "memory");
  }
}

__global__ void gather_kernel(const int4* __restrict__ idx, const double2* __restrict__ val, int64_t n4,
                              const double* __restrict__ v, double* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double acc = 0;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n4; g += stride) {
    const int4 r = idx[g];
    const double2 a = val[2 * g], b = val[2 * g + 1];
    acc += a.x * __ldg(v + r.x) + a.y * __ldg(v + r.y) + b.x * __ldg(v + r.z) + b.y * __ldg(v + r.w);
  }
  if (acc == 1.2345e-300) out[0] = acc;
}

// shared-memory FP64 atomics on a private band of `band` rows, then a coalesced RED flush
__global__ void smem_atomic_kernel(const int4* __restrict__ idx, const double2* __restrict__ val, int64_t n4, int band,
                                   double* __restrict__ out) {
  extern __shared__ double acc[];
  for (int r = threadIdx.x; r < band; r += blockDim.x) acc[r] = 0.0;
  __syncthreads();
  const int64_t per = (n4 + gridDim.x - 1) / gridDim.x;
  const int64_t g0 = per * blockIdx.x, g1 = (g0 + per < n4) ? g0 + per : n4;
  for (int64_t g = g0 + threadIdx.x; g < g1; g += blockDim.x) {
    const int4 r = idx[g];
    const double2 a = val[2 * g], b = val[2 * g + 1];
    atomicAdd(&acc[(unsigned)r.x % band], a.x);
    atomicAdd(&acc[(unsigned)r.y % band], a.y);
    atomicAdd(&acc[(unsigned)r.z % band], b.x);
    atomicAdd(&acc[(unsigned)r.w % band], b.y);
  }
  __syncthreads();
  for (int r = threadIdx.x; r < band; r += blockDim.x)
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(out + r), "d"(acc[r]) : "memory");
}

// short strided runs: CTA b reads `run` consecutive entries out of every `period` (the banded access pattern)
__global__ void runs_kernel(const int32_t* __restrict__ idx, const double* __restrict__ val, int64_t n, int run,
                            int period, int nbands, double* __restrict__ out) {
  const int band = blockIdx.x % nbands;
  const int64_t ncolumns = n / period;
  double acc = 0;
  const int lanes_per_run = run;  // one thread per entry of a run
  const int runs_per_iter = blockDim.x / lanes_per_run;
  const int my_run = threadIdx.x / lanes_per_run, my_k = threadIdx.x % lanes_per_run;
  const int64_t c_begin = (ncolumns * (blockIdx.x / nbands)) / (gridDim.x / nbands);
  const int64_t c_end = (ncolumns * (blockIdx.x / nbands + 1)) / (gridDim.x / nbands);
  if (my_run < runs_per_iter)
    for (int64_t c = c_begin + my_run; c < c_end; c += runs_per_iter) {
      const int64_t k = c * period + (int64_t)band * run + my_k;
      acc += val[k] * (double)idx[k];
    }
  if (acc == 1.2345e-300) out[0] = acc;
}

template <typename F>
static float time_it(F f, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  f();
  f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    f();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

// G consecutive doubles at a pseudo-random G-aligned place of an n-double region: what a scattered store of
// 8 / 32 / 128 bytes costs per byte (the transpose writes 4 + 8 bytes per entry to unrelated sectors)
template <int G>
__global__ void scatter_store_kernel(double* __restrict__ out, int64_t n) {
  const int64_t groups = n / G;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; k < groups; k += stride) {
    const int64_t g = static_cast<int64_t>(mix64(static_cast<uint64_t>(k)) % static_cast<uint64_t>(groups));
    double* dst = out + g * G;
#pragma unroll
    for (int t = 0; t < G; t += 2) *reinterpret_cast<double2*>(dst + t) = make_double2(1.0 + k, 2.0);
  }
}
template <>
__global__ void scatter_store_kernel<1>(double* __restrict__ out, int64_t n) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; k < n; k += stride)
    out[static_cast<int64_t>(mix64(static_cast<uint64_t>(k)) % static_cast<uint64_t>(n))] = 1.0 + k;
}

// F append fronts (output rows), every entry goes to the next free slot of a pseudo-random front: each 32-byte
// sector receives its 4 doubles at 4 far-apart times, like the rows of a transposed matrix
__global__ void front_store_kernel(double* __restrict__ out, int64_t n, int64_t fronts) {
  const int64_t per = n / fronts;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; k < per * fronts; k += stride) {
    const int64_t round = k / fronts;  // all fronts advance together: slot `round` of front f
    const int64_t f = static_cast<int64_t>(mix64(static_cast<uint64_t>(k)) % static_cast<uint64_t>(fronts));
    out[f * per + round] = 1.0 + k;
  }
}

int main(int argc, char** argv) {
  const int64_t n = argc > 1 ? atoll(argv[1]) : 100000000LL;
  const int32_t rows = argc > 2 ? atoi(argv[2]) : 1000000;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("{\"test\":\"device\",\"name\":\"%s\",\"sms\":%d,\"l2_mb\":%.1f,\"n\":%lld,\"rows\":%d}\n", prop.name, sms,
         prop.l2CacheSize / 1048576.0, (long long)n, rows);
  int32_t* idx;
  double *val, *val2, *out;
  uint32_t* cnt;
  CK(cudaMalloc(&idx, n * 4 + 64));
  CK(cudaMalloc(&val, n * 8 + 64));
  CK(cudaMalloc(&val2, n * 8 + 64));
  CK(cudaMalloc(&out, (size_t)rows * 8 + 64));
  CK(cudaMalloc(&cnt, (size_t)rows * 4 + 64));
  CK(cudaMemset(out, 0, (size_t)rows * 8));
  CK(cudaMemset(cnt, 0, (size_t)rows * 4));
  init_kernel<<<sms * 8, 256>>>(idx, val, n, rows, 0);
  CK(cudaDeviceSynchronize());
  const int64_t n4 = n / 4, n2 = n / 2;
  const int reps = 5;
  if (argc > 3 && std::string(argv[3]) == "stores") {
    const int grid = sms * 8;
    float ms = time_it([&] { scatter_store_kernel<1><<<grid, 256>>>(val2, n); }, reps);
    printf("{\"test\":\"scatter_store\",\"bytes_per_store\":8,\"ms\":%.4f,\"GBps\":%.1f}\n", ms, n * 8.0 / ms / 1e6);
    ms = time_it([&] { scatter_store_kernel<2><<<grid, 256>>>(val2, n); }, reps);
    printf("{\"test\":\"scatter_store\",\"bytes_per_store\":16,\"ms\":%.4f,\"GBps\":%.1f}\n", ms, n * 8.0 / ms / 1e6);
    ms = time_it([&] { scatter_store_kernel<4><<<grid, 256>>>(val2, n); }, reps);
    printf("{\"test\":\"scatter_store\",\"bytes_per_store\":32,\"ms\":%.4f,\"GBps\":%.1f}\n", ms, n * 8.0 / ms / 1e6);
    ms = time_it([&] { scatter_store_kernel<16><<<grid, 256>>>(val2, n); }, reps);
    printf("{\"test\":\"scatter_store\",\"bytes_per_store\":128,\"ms\":%.4f,\"GBps\":%.1f}\n", ms, n * 8.0 / ms / 1e6);
    for (int64_t fronts : {1000LL, 30000LL, 180000LL, 1000000LL, 10000000LL}) {
      ms = time_it([&] { front_store_kernel<<<grid, 256>>>(val2, n, fronts); }, reps);
      printf("{\"test\":\"front_store\",\"fronts\":%lld,\"ms\":%.4f,\"GBps\":%.1f}\n", (long long)fronts, ms, n * 8.0 / ms / 1e6);
    }
    return 0;
  }
  for (int mult = 4; mult <= 16; mult *= 2) {
    const int grid = sms * mult;
    float ms = time_it([&] { read_kernel<<<grid, 256>>>((const double2*)val, n2, out); }, reps);
    printf("{\"test\":\"read_f64\",\"grid\":%d,\"ms\":%.4f,\"GBps\":%.1f}\n", grid, ms, n * 8.0 / ms / 1e6);
  }
  {
    float ms = time_it([&] { copy_kernel<<<sms * 8, 256>>>((const double2*)val, (double2*)val2, n2); }, reps);
    printf("{\"test\":\"copy_f64\",\"ms\":%.4f,\"GBps\":%.1f}\n", ms, n * 16.0 / ms / 1e6);
    ms = time_it([&] { CK(cudaMemcpyAsync(val2, val, n * 8, cudaMemcpyDeviceToDevice)); }, reps);
    printf("{\"test\":\"memcpy_d2d\",\"ms\":%.4f,\"GBps\":%.1f}\n", ms, n * 16.0 / ms / 1e6);
  }
  for (int mult = 4; mult <= 16; mult *= 2) {
    const int grid = sms * mult;
    float ms = time_it([&] { red_f64_kernel<0><<<grid, 256>>>((const int4*)idx, (const double2*)val, n4, rows, out); }, reps);
    printf("{\"test\":\"red_f64_random\",\"grid\":%d,\"rows\":%d,\"ms\":%.4f,\"Gops\":%.2f,\"GBps_12B\":%.1f}\n", grid, rows, ms,
           n / ms / 1e6, n * 12.0 / ms / 1e6);
  }
  {
    float ms = time_it([&] { red_f64_kernel<1><<<sms * 8, 256>>>((const int4*)idx, (const double2*)val, n4, rows, out); }, reps);
    printf("{\"test\":\"red_f64_seq4\",\"ms\":%.4f,\"Gops\":%.2f}\n", ms, n / ms / 1e6);
    ms = time_it([&] { red_f64_lane_kernel<<<sms * 8, 256>>>(val, n, rows, out); }, reps);
    printf("{\"test\":\"red_f64_lane_coalesced\",\"ms\":%.4f,\"Gops\":%.2f}\n", ms, n / ms / 1e6);
    ms = time_it([&] { red_f64_kernel<0><<<sms * 8, 256>>>((const int4*)idx, (const double2*)val, n4, 30000, out); }, reps);
    printf("{\"test\":\"red_f64_random\",\"grid\":%d,\"rows\":30000,\"ms\":%.4f,\"Gops\":%.2f}\n", sms * 8, ms, n / ms / 1e6);
  }
  {
    float ms = time_it([&] { red_u32_kernel<<<sms * 8, 256>>>((const int4*)idx, n4, cnt); }, reps);
    printf("{\"test\":\"red_u32_random\",\"rows\":%d,\"ms\":%.4f,\"Gops\":%.2f}\n", rows, ms, n / ms / 1e6);
  }
  for (int mult = 4; mult <= 16; mult *= 2) {
    const int grid = sms * mult;
    float ms = time_it([&] { gather_kernel<<<grid, 256>>>((const int4*)idx, (const double2*)val, n4, out, val2); }, reps);
    printf("{\"test\":\"gather_f64_random\",\"grid\":%d,\"rows\":%d,\"ms\":%.4f,\"Gops\":%.2f,\"GBps_12B\":%.1f}\n", grid, rows, ms,
           n / ms / 1e6, n * 12.0 / ms / 1e6);
  }
  for (int band = 4096; band <= 24576; band *= 2) {
    if (band > 24576) break;
    CK(cudaFuncSetAttribute(smem_atomic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, band * 8));
    float ms = time_it([&] { smem_atomic_kernel<<<sms, 1024, band * 8>>>((const int4*)idx, (const double2*)val, n4, band, out); }, reps);
    printf("{\"test\":\"smem_atomic_f64\",\"band\":%d,\"ms\":%.4f,\"Gops\":%.2f,\"GBps_12B\":%.1f}\n", band, ms, n / ms / 1e6,
           n * 12.0 / ms / 1e6);
  }
  for (int run = 4; run <= 32; run *= 2) {
    const int nbands = 148, period = run * nbands;
    const int col_groups = 4;
    float ms = time_it([&] { runs_kernel<<<nbands * col_groups, 256>>>(idx, val, n, run, period, nbands, out); }, reps);
    const double touched = (double)(n / period) * period;
    printf("{\"test\":\"banded_runs\",\"run\":%d,\"ms\":%.4f,\"GBps_12B\":%.1f}\n", run, ms, touched * 12.0 / ms / 1e6);
  }
  return 0;
}
