// scripts/l2_bench.cu -- microbenchmark 2: how much of the gathered vector can L2 keep, and do
// per-load L2 eviction-priority hints (createpolicy.range) protect a hot prefix from the cold tail
// and from the index stream?  (exploration tool for DESIGN.md; not part of the product)
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e = (x);                                                           \
    if (e != cudaSuccess) {                                                        \
      printf("CUDA error %s at line %d: %s\n", cudaGetErrorName(e), __LINE__, #x); \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

constexpr int U = 8;

__device__ __forceinline__ uint64_t mix(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// hot_pct percent of the indices fall uniformly in [0, hot), the rest uniformly in [hot, table)
__global__ void fill_idx(uint32_t *idx, uint64_t count, uint32_t hot, uint32_t table, uint32_t hot_permille) {
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
    uint64_t h = mix(i);
    uint32_t sel = (uint32_t)(h % 1000);
    uint64_t r = mix(h);
    idx[i] = (sel < hot_permille || hot >= table) ? (uint32_t)(r % hot) : hot + (uint32_t)(r % (table - hot));
  }
}
__global__ void fill_tab(double *t, uint64_t n) {
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) t[i] = 1.0 / (double)(i + 1);
}

enum { PLAIN = 0, RANGE_LAST_FIRST, RANGE_LAST_NORMAL, RANGE_LAST_NOALLOC, FRAC_LAST };

template <int MODE>
__global__ void __launch_bounds__(256) gather(const uint32_t *__restrict__ idx, uint64_t count,
                                              const double *__restrict__ table, uint32_t hot_bytes,
                                              uint32_t total_bytes, double *out) {
  const int lane = threadIdx.x & 31;
  const uint64_t warps = (uint64_t)gridDim.x * (blockDim.x >> 5);
  const uint64_t w0 = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  uint64_t pol = 0, pol_idx;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_idx));
  if (MODE == RANGE_LAST_FIRST)
    asm volatile("createpolicy.range.global.L2::evict_last.L2::evict_first.b64 %0, [%1], %2, %3;"
                 : "=l"(pol) : "l"(table), "r"(hot_bytes), "r"(total_bytes));
  if (MODE == RANGE_LAST_NORMAL)
    asm volatile("createpolicy.range.global.L2::evict_normal.L2::evict_first.b64 %0, [%1], %2, %3;"
                 : "=l"(pol) : "l"(table), "r"(hot_bytes), "r"(total_bytes));
  if (MODE == RANGE_LAST_NOALLOC)
    asm volatile("createpolicy.range.global.L2::evict_last.L2::evict_unchanged.b64 %0, [%1], %2, %3;"
                 : "=l"(pol) : "l"(table), "r"(hot_bytes), "r"(total_bytes));
  if (MODE == FRAC_LAST) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  double acc = 0;
  const uint64_t chunks = count / (32 * U);
  for (uint64_t c = w0; c < chunks; c += warps) {
    const uint32_t *p = idx + c * 32 * U + lane;
    uint32_t s[U];
    double v[U];
#pragma unroll
    for (int j = 0; j < U; ++j)
      asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(s[j]) : "l"(p + j * 32), "l"(pol_idx));
#pragma unroll
    for (int j = 0; j < U; ++j) {
      if (MODE == PLAIN)
        v[j] = __ldg(table + s[j]);
      else
        asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v[j]) : "l"(table + s[j]), "l"(pol));
    }
#pragma unroll
    for (int j = 0; j < U; ++j) acc += v[j];
  }
  if (acc == 123.456) out[0] = acc;
}

template <typename F>
double time_ms(F f, int reps = 4) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  f();
  f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(a));
    f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  const uint64_t count = 1ull << 28;
  uint32_t *idx;
  double *table, *out;
  const uint64_t max_table = 1ull << 26;
  CK(cudaMalloc(&idx, count * 4));
  CK(cudaMalloc(&table, (max_table + 2) * 8));
  CK(cudaMalloc(&out, 64));
  fill_tab<<<148 * 8, 256>>>(table, max_table + 2);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s, %d SMs, L2 %d MiB, persistingL2CacheMaxSize %d MiB\n", prop.name, prop.multiProcessorCount,
         prop.l2CacheSize >> 20, prop.persistingL2CacheMaxSize >> 20);
  const int grid = prop.multiProcessorCount * 8;
  auto report = [&](const char *name, double hot_mb, double tab_mb, int permille, double ms) {
    printf("%-34s hot=%6.1f MiB table=%6.1f MiB hot%%=%5.1f : %7.3f ms %7.2f Ggather/s\n", name, hot_mb, tab_mb,
           permille / 10.0, ms, count / ms / 1e6);
    fflush(stdout);
  };
  // 1. effective L2 capacity: uniform random gathers over tables of growing size (index stream evict-first)
  for (int mb : {16, 32, 48, 56, 64, 72, 80, 96, 112, 128, 160, 256}) {
    const uint32_t ent = (uint32_t)((uint64_t)mb * 1048576 / 8);
    fill_idx<<<148 * 8, 256>>>(idx, count, ent, ent, 1000);
    report("uniform plain", mb, mb, 1000, time_ms([&] { gather<PLAIN><<<grid, 256>>>(idx, count, table, 0, 0, out); }));
    report("uniform frac evict_last", mb, mb, 1000,
           time_ms([&] { gather<FRAC_LAST><<<grid, 256>>>(idx, count, table, 0, 0, out); }));
  }
  // 2. hot prefix (hot% of the accesses) + cold tail over a 512 MiB vector
  const uint32_t total_bytes = (uint32_t)(max_table * 8);
  for (int permille : {970, 900}) {
    for (int mb : {32, 48, 64, 88, 112}) {
      const uint32_t hot = (uint32_t)((uint64_t)mb * 1048576 / 8);
      const uint32_t hot_bytes = hot * 8u;
      fill_idx<<<148 * 8, 256>>>(idx, count, hot, (uint32_t)max_table, permille);
      report("split plain", mb, 512, permille,
             time_ms([&] { gather<PLAIN><<<grid, 256>>>(idx, count, table, hot_bytes, total_bytes, out); }));
      report("split range last/first", mb, 512, permille,
             time_ms([&] { gather<RANGE_LAST_FIRST><<<grid, 256>>>(idx, count, table, hot_bytes, total_bytes, out); }));
      report("split range normal/first", mb, 512, permille,
             time_ms([&] { gather<RANGE_LAST_NORMAL><<<grid, 256>>>(idx, count, table, hot_bytes, total_bytes, out); }));
      report("split range last/unchanged", mb, 512, permille,
             time_ms([&] { gather<RANGE_LAST_NOALLOC><<<grid, 256>>>(idx, count, table, hot_bytes, total_bytes, out); }));
    }
  }
  return 0;
}
