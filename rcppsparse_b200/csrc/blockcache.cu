// Device memory for everything the library allocates: pool_alloc / pool_free.
//
// Underneath is the device's stream-ordered pool (cudaMallocAsync, release threshold = max so the driver keeps what it
// has mapped).  On top of it sits a per-device cache of freed blocks keyed by capacity.  Why: the call pattern of the
// reference's API is create -> op -> destroy per .Call (reference src/RcppExports.cpp:20: the Exporter builds a fresh
// Matrix every call), i.e. the same block sizes again and again, but each time on a NEW stream.  Measured on B200
// (tools/e2e_transpose_probe.py, SB200_TRACE=1): cudaMallocAsync serving such a request from memory the pool already
// holds, freed on another (since destroyed) stream, takes anything from 1 ms to 950 ms (91 MB block, pool reserved
// 4.0 GB, used 2.5 GB) -- a one-shot transpose of C2 swung between 73 and 1390 ms on it.  A block handed back from the
// cache costs a map lookup and one cudaStreamWaitEvent.
//
// Ordering: pool_free(ptr, s) records the block's event on s; the next owner's stream waits for that event before the
// block is handed out, so the block is reused only after everything that was enqueued on s at the time of the free.
// Capacity classes: 512 B steps below 1 MiB, 2 MiB steps above (the driver maps in 2 MiB pages); a request takes the
// smallest idle block that is at most 1/8 (+ one step) larger.  The cache holds at most a third of the device's memory
// (SB200_CACHE_MB overrides), oldest blocks go back to the driver pool first; when the driver reports out of memory
// everything idle goes back and the allocation is tried again.  sb200_trim() empties cache and pool.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <chrono>
#include <map>
#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace sb200 {
namespace {

constexpr int MAX_DEVICES = 32;
constexpr size_t SMALL_STEP = 512, LARGE_STEP = 2u << 20, LARGE_FROM = 1u << 20;

struct Block {
  void* ptr = nullptr;
  size_t cap = 0;
  cudaEvent_t ev = nullptr;  // recorded at the last free
  int device = 0;
  uint64_t tick = 0;  // order of the frees: the oldest idle block is evicted first
};

struct DeviceCache {
  std::multimap<size_t, Block> idle;
  size_t idle_bytes = 0;
  size_t limit = 0;
  bool configured = false;
};

// The cache lives for the whole process and is never destroyed: handles may be torn down by a host language's
// finalisers after static destructors have started.
struct Cache {
  std::mutex mu;
  std::unordered_map<void*, Block> live;  // blocks in use, by address
  DeviceCache dev[MAX_DEVICES];
  uint64_t tick = 0;
};
Cache& cache() {
  static Cache* c = new Cache();
  return *c;
}
size_t capacity_class(size_t bytes) {
  if (bytes < 16) bytes = 16;
  const size_t step = bytes >= LARGE_FROM ? LARGE_STEP : SMALL_STEP;
  return (bytes + step - 1) / step * step;
}

// cache().mu held
void configure(int dev) {
  DeviceCache& dc = cache().dev[dev];
  if (dc.configured) return;
  dc.configured = true;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    uint64_t keep = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) == cudaSuccess) dc.limit = prop.totalGlobalMem / 3;
  if (const char* e = getenv("SB200_CACHE_MB")) dc.limit = static_cast<size_t>(atoll(e)) << 20;
  cudaGetLastError();
}

// cache().mu held; the block leaves the cache for the driver pool, ordered after its last use
void release_to_driver(Block& b, cudaStream_t s) {
  if (b.ev) {
    cudaStreamWaitEvent(s, b.ev, 0);
    cudaEventDestroy(b.ev);
  }
  cudaFreeAsync(b.ptr, s);
  cudaGetLastError();
}

// cache().mu held
void flush_device(int dev, cudaStream_t s) {
  DeviceCache& dc = cache().dev[dev];
  for (auto& kv : dc.idle) release_to_driver(kv.second, s);
  dc.idle.clear();
  dc.idle_bytes = 0;
}

}  // namespace

int pool_alloc(void** out, size_t bytes, cudaStream_t s) {
  static const bool trace = getenv("SB200_TRACE") != nullptr;
  *out = nullptr;
  int dev = 0;
  SB_CUDA(cudaGetDevice(&dev));
  // cudaMallocAsync takes the pool of the STREAM's device; the legacy and per-thread streams belong to the current one
  if (s != nullptr && s != cudaStreamLegacy && s != cudaStreamPerThread && cudaStreamGetDevice(s, &dev) != cudaSuccess) {
    cudaGetLastError();
    SB_CUDA(cudaGetDevice(&dev));
  }
  if (dev < 0 || dev >= MAX_DEVICES) return fail(SB200_E_INVALID, "device index out of range");
  const size_t cap = capacity_class(bytes);
  size_t idle_now = 0;
  {
    std::lock_guard<std::mutex> lock(cache().mu);
    configure(dev);
    DeviceCache& dc = cache().dev[dev];
    auto it = dc.idle.lower_bound(cap);
    if (it != dc.idle.end() && it->first <= cap + cap / 8 + (cap >= LARGE_FROM ? LARGE_STEP : SMALL_STEP)) {
      Block b = it->second;
      dc.idle.erase(it);
      dc.idle_bytes -= b.cap;
      cudaError_t e = cudaStreamWaitEvent(s, b.ev, 0);
      if (e != cudaSuccess) {  // keep the block out of circulation rather than hand it out unordered
        release_to_driver(b, static_cast<cudaStream_t>(0));
        return cuda_fail(e, "cudaStreamWaitEvent (block cache)", __FILE__, __LINE__);
      }
      cache().live[b.ptr] = b;
      *out = b.ptr;
      return SB200_OK;
    }
    idle_now = dc.idle_bytes;
  }
  // miss: the driver pool serves it (outside the lock: growing the pool can take milliseconds, and the workers of the
  // one-process sharding layer allocate on several devices at once)
  Block b;
  b.cap = cap;
  b.device = dev;
  const auto t0 = std::chrono::steady_clock::now();
  cudaError_t e = cudaMallocAsync(&b.ptr, cap, s);
  if (e == cudaErrorMemoryAllocation && idle_now > 0) {  // idle blocks of other sizes are in the way
    cudaGetLastError();
    DeviceGuard guard(dev);
    cudaDeviceSynchronize();
    {
      std::lock_guard<std::mutex> lock(cache().mu);
      flush_device(dev, static_cast<cudaStream_t>(0));
    }
    cudaDeviceSynchronize();
    e = cudaMallocAsync(&b.ptr, cap, s);
  }
  if (trace) {  // a request the driver pool was slow to serve shows up as milliseconds here
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (ms > 0.5) fprintf(stderr, "[sb200 trace] cudaMallocAsync of %.1f MB took %.2f ms (cache idle %.1f MB)\n", cap / 1048576.0, ms, idle_now / 1048576.0);
  }
  if (e != cudaSuccess) return cuda_fail(e, "cudaMallocAsync", __FILE__, __LINE__);
  e = cudaEventCreateWithFlags(&b.ev, cudaEventDisableTiming);
  if (e != cudaSuccess) {
    cudaFreeAsync(b.ptr, s);
    return cuda_fail(e, "cudaEventCreate (block cache)", __FILE__, __LINE__);
  }
  {
    std::lock_guard<std::mutex> lock(cache().mu);
    cache().live[b.ptr] = b;
  }
  *out = b.ptr;
  return SB200_OK;
}

void pool_free(void* ptr, cudaStream_t s) {
  if (!ptr) return;
  std::lock_guard<std::mutex> lock(cache().mu);
  auto it = cache().live.find(ptr);
  if (it == cache().live.end()) {  // not ours (never happens inside the library): plain stream-ordered free
    cudaFreeAsync(ptr, s);
    cudaGetLastError();
    return;
  }
  Block b = it->second;
  cache().live.erase(it);
  DeviceCache& dc = cache().dev[b.device];
  if (cudaEventRecord(b.ev, s) != cudaSuccess) {  // a stream that is gone: nothing can be ordered after it any more
    cudaGetLastError();
    cudaEventDestroy(b.ev);
    cudaFree(b.ptr);  // synchronising free, valid for pool memory
    cudaGetLastError();
    return;
  }
  if (b.cap > dc.limit) {  // larger than the whole cache (or the cache is off): straight back to the driver pool
    cudaEventDestroy(b.ev);
    b.ev = nullptr;
    release_to_driver(b, s);
    return;
  }
  b.tick = ++cache().tick;
  dc.idle.emplace(b.cap, b);
  dc.idle_bytes += b.cap;
  while (dc.idle_bytes > dc.limit && !dc.idle.empty()) {  // oldest first
    auto old = dc.idle.begin();
    for (auto j = dc.idle.begin(); j != dc.idle.end(); ++j)
      if (j->second.tick < old->second.tick) old = j;
    dc.idle_bytes -= old->second.cap;
    release_to_driver(old->second, s);
    dc.idle.erase(old);
  }
}

size_t pool_idle_bytes(int device) {
  if (device < 0 || device >= MAX_DEVICES) return 0;
  std::lock_guard<std::mutex> lock(cache().mu);
  return cache().dev[device].idle_bytes;
}

// the device must be current and idle (sb200_trim synchronises it first)
void pool_release_idle(int device) {
  if (device < 0 || device >= MAX_DEVICES) return;
  std::lock_guard<std::mutex> lock(cache().mu);
  flush_device(device, static_cast<cudaStream_t>(0));
}

}  // namespace sb200
