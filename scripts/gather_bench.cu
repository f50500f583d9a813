// scripts/gather_bench.cu -- microbenchmark: how fast can one B200 do random 8-byte gathers, by path?
// (exploration tool for DESIGN.md "gather rate"; not part of the product)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/gather_bench scripts/gather_bench.cu
// Each variant streams a uint32 index array (coalesced, like the SELL kernel) and sums table[idx].
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace cg = cooperative_groups;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e = (x);                                                               \
    if (e != cudaSuccess) {                                                            \
      printf("CUDA error %s at line %d: %s\n", cudaGetErrorName(e), __LINE__, #x);     \
      exit(1);                                                                         \
    }                                                                                  \
  } while (0)

constexpr int U = 8;

__device__ __forceinline__ uint64_t mix(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__global__ void fill_idx(uint32_t *idx, uint64_t count, uint32_t table, int skew) {
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
    uint64_t h = mix(i);
    uint32_t v;
    if (skew) {  // RMAT-like: each bit is 1 with probability 0.24 -> hot set = few-ones ids... then sorted-by-heat label
      // approximate a degree-sorted label: draw geometric-ish skew by squaring a uniform
      double u = (double)(h >> 11) * (1.0 / 9007199254740992.0);
      v = (uint32_t)(u * u * u * u * table);
    } else {
      v = (uint32_t)(h % table);
    }
    idx[i] = v < table ? v : table - 1;
  }
}
__global__ void fill_tab(double *t, uint64_t n) {
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) t[i] = 1.0 / (double)(i + 1);
}

enum { LD_NC = 0, LD_PLAIN, LD_CG, LD_NOALLOC, LD_VOLATILE, LD_F32, LD_V2 };

template <int MODE>
__device__ __forceinline__ double ld(const double *p) {
  double v;
  if (MODE == LD_NC) {
    v = __ldg(p);
  } else if (MODE == LD_PLAIN) {
    asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
  } else if (MODE == LD_CG) {
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
  } else if (MODE == LD_NOALLOC) {
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  } else if (MODE == LD_VOLATILE) {
    asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
  } else if (MODE == LD_F32) {
    float f;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(f) : "l"(reinterpret_cast<const float *>(p)));
    v = f;
  } else {  // 16-byte gather
    double a, b;
    asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "l"((const double *)((uintptr_t)p & ~15ull)));
    v = a + b;
  }
  return v;
}

template <int MODE>
__global__ void __launch_bounds__(256) gather_global(const uint32_t *__restrict__ idx, uint64_t count,
                                                     const double *__restrict__ table, double *out) {
  const int lane = threadIdx.x & 31;
  const uint64_t warps = (uint64_t)gridDim.x * (blockDim.x >> 5);
  const uint64_t w0 = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  double acc = 0;
  const uint64_t chunks = count / (32 * U);
  for (uint64_t c = w0; c < chunks; c += warps) {
    const uint32_t *p = idx + c * 32 * U + lane;
    uint32_t s[U];
    double v[U];
#pragma unroll
    for (int j = 0; j < U; ++j) s[j] = __ldcs(p + j * 32);
#pragma unroll
    for (int j = 0; j < U; ++j) v[j] = ld<MODE>(table + s[j]);
#pragma unroll
    for (int j = 0; j < U; ++j) acc += v[j];
  }
  if (acc == 123.456) out[0] = acc;
}

// table (first `hot` entries) staged in shared memory; indices >= hot go to global
__global__ void __launch_bounds__(1024) gather_smem(const uint32_t *__restrict__ idx, uint64_t count,
                                                    const double *__restrict__ table, uint32_t hot, double *out) {
  extern __shared__ double sm[];
  for (uint32_t i = threadIdx.x; i < hot; i += blockDim.x) sm[i] = table[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const uint64_t warps = (uint64_t)gridDim.x * (blockDim.x >> 5);
  const uint64_t w0 = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  double acc = 0;
  const uint64_t chunks = count / (32 * U);
  for (uint64_t c = w0; c < chunks; c += warps) {
    const uint32_t *p = idx + c * 32 * U + lane;
    uint32_t s[U];
    double v[U];
#pragma unroll
    for (int j = 0; j < U; ++j) s[j] = __ldcs(p + j * 32);
#pragma unroll
    for (int j = 0; j < U; ++j) v[j] = s[j] < hot ? sm[s[j]] : __ldg(table + s[j]);
#pragma unroll
    for (int j = 0; j < U; ++j) acc += v[j];
  }
  if (acc == 123.456) out[0] = acc;
}

// hot entries spread over the shared memories of a thread-block cluster (DSMEM gathers)
__global__ void __launch_bounds__(1024) gather_dsmem(const uint32_t *__restrict__ idx, uint64_t count,
                                                     const double *__restrict__ table, uint32_t per_cta,
                                                     double *out) {
  extern __shared__ double sm[];
  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t crank = cluster.block_rank();
  const uint32_t csize = cluster.num_blocks();
  for (uint32_t i = threadIdx.x; i < per_cta; i += blockDim.x) sm[i] = table[crank * per_cta + i];
  cluster.sync();
  const uint32_t hot = per_cta * csize;
  const int lane = threadIdx.x & 31;
  const uint64_t warps = (uint64_t)gridDim.x * (blockDim.x >> 5);
  const uint64_t w0 = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  double acc = 0;
  const uint64_t chunks = count / (32 * U);
  uint32_t sm_base;
  asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(sm_base) : "l"(sm));
  for (uint64_t c = w0; c < chunks; c += warps) {
    const uint32_t *p = idx + c * 32 * U + lane;
    uint32_t s[U];
    double v[U];
#pragma unroll
    for (int j = 0; j < U; ++j) s[j] = __ldcs(p + j * 32);
#pragma unroll
    for (int j = 0; j < U; ++j) {
      if (s[j] < hot) {
        const uint32_t r = s[j] / per_cta, off = s[j] - r * per_cta;
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(sm_base + off * 8), "r"(r));
        asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v[j]) : "r"(remote));
      } else {
        v[j] = __ldg(table + s[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < U; ++j) acc += v[j];
  }
  if (acc == 123.456) out[0] = acc;
  cluster.sync();
}

// scatter-add (RED.F64) for comparison with the gather direction
__global__ void __launch_bounds__(256) scatter_red(const uint32_t *__restrict__ idx, uint64_t count, double *table) {
  const int lane = threadIdx.x & 31;
  const uint64_t warps = (uint64_t)gridDim.x * (blockDim.x >> 5);
  const uint64_t w0 = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const uint64_t chunks = count / (32 * U);
  for (uint64_t c = w0; c < chunks; c += warps) {
    const uint32_t *p = idx + c * 32 * U + lane;
#pragma unroll
    for (int j = 0; j < U; ++j) atomicAdd(table + __ldcs(p + j * 32), 1.0);
  }
}

// per-lane 16-byte cp.async.bulk (TMA engine, no L1tex tag stage) into shared memory, mbarrier per warp
__global__ void __launch_bounds__(256) gather_bulk(const uint32_t *__restrict__ idx, uint64_t count,
                                                   const double *__restrict__ table, double *out) {
  __shared__ __align__(16) double stage[8][U][32][2];
  __shared__ __align__(8) unsigned long long bars[8];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint64_t warps = (uint64_t)gridDim.x * (blockDim.x >> 5);
  const uint64_t w0 = (uint64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  uint32_t bar = (uint32_t)__cvta_generic_to_shared(&bars[wib]);
  if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
  __syncwarp();
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  double acc = 0;
  uint32_t phase = 0;
  const uint64_t chunks = count / (32 * U);
  for (uint64_t c = w0; c < chunks; c += warps) {
    const uint32_t *p = idx + c * 32 * U + lane;
    uint32_t s[U];
#pragma unroll
    for (int j = 0; j < U; ++j) s[j] = __ldcs(p + j * 32);
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(32 * U * 16));
    __syncwarp();
#pragma unroll
    for (int j = 0; j < U; ++j) {
      uint32_t dst = (uint32_t)__cvta_generic_to_shared(&stage[wib][j][lane][0]);
      const double *src = (const double *)((uintptr_t)(table + s[j]) & ~15ull);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];" ::"r"(dst),
                   "l"(src), "r"(bar)
                   : "memory");
    }
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(ok)
                   : "r"(bar), "r"(phase)
                   : "memory");
    }
    phase ^= 1;
#pragma unroll
    for (int j = 0; j < U; ++j) acc += stage[wib][j][lane][s[j] & 1];
    __syncwarp();
  }
  if (acc == 123.456) out[0] = acc;
}

template <typename F>
double time_ms(F f, int reps = 3) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
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
  const uint64_t count = 1ull << 28;  // 268M gathers per launch (1 GiB of indices, larger than L2)
  uint32_t *idx;
  double *table, *out;
  const uint64_t max_table = 1ull << 26;  // 512 MiB of doubles
  CK(cudaMalloc(&idx, count * 4));
  CK(cudaMalloc(&table, (max_table + 2) * 8));
  CK(cudaMalloc(&out, 64));
  fill_tab<<<148 * 8, 256>>>(table, max_table + 2);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs\n", prop.name, sms);
  auto report = [&](const char *name, uint64_t tab, int skew, double ms) {
    printf("%-28s table=%9llu entries (%7.1f MiB) skew=%d : %8.3f ms  %7.2f Ggather/s  %6.3f gather/cycle/SM@1.965GHz\n",
           name, (unsigned long long)tab, tab * 8.0 / 1048576.0, skew, ms, count / ms / 1e6,
           count / (ms * 1e-3) / sms / 1.965e9);
    fflush(stdout);
  };
  const int grid = sms * 4;
  for (int skew = 0; skew < 2; ++skew) {
    for (uint64_t tab : {1ull << 13, 1ull << 22, 1ull << 26}) {
      fill_idx<<<148 * 8, 256>>>(idx, count, (uint32_t)tab, skew);
      CK(cudaDeviceSynchronize());
      report("ld.global.nc (ldg)", tab, skew, time_ms([&] { gather_global<LD_NC><<<grid, 256>>>(idx, count, table, out); }));
      report("ld.global", tab, skew, time_ms([&] { gather_global<LD_PLAIN><<<grid, 256>>>(idx, count, table, out); }));
      report("ld.global.cg", tab, skew, time_ms([&] { gather_global<LD_CG><<<grid, 256>>>(idx, count, table, out); }));
      report("ld.nc.L1::no_allocate", tab, skew,
             time_ms([&] { gather_global<LD_NOALLOC><<<grid, 256>>>(idx, count, table, out); }));
      report("ld.volatile", tab, skew,
             time_ms([&] { gather_global<LD_VOLATILE><<<grid, 256>>>(idx, count, table, out); }));
      report("ld.nc.f32 (4B)", tab, skew, time_ms([&] { gather_global<LD_F32><<<grid, 256>>>(idx, count, table, out); }));
      report("ld.nc.v2.f64 (16B)", tab, skew, time_ms([&] { gather_global<LD_V2><<<grid, 256>>>(idx, count, table, out); }));
      report("cp.async.bulk 16B/lane", tab, skew, time_ms([&] { gather_bulk<<<grid, 256>>>(idx, count, table, out); }));
      if (skew == 0 && tab == (1ull << 26))
        report("red.add.f64 scatter", tab, skew, time_ms([&] { scatter_red<<<grid, 256>>>(idx, count, table); }));
    }
  }
  // occupancy variants of the plain gather (more warps in flight)
  fill_idx<<<148 * 8, 256>>>(idx, count, 1u << 26, 0);
  for (int g : {sms * 2, sms * 4, sms * 8}) {
    char nm[64];
    snprintf(nm, sizeof nm, "ldg grid=%d", g);
    report(nm, 1ull << 26, 0, time_ms([&] { gather_global<LD_NC><<<g, 256>>>(idx, count, table, out); }));
  }
  // shared-memory hot set: all indices inside the hot set (pure smem gather rate), 1 CTA/SM
  {
    const uint32_t hot = 24 * 1024;  // 192 KiB of doubles
    CK(cudaFuncSetAttribute(gather_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, hot * 8));
    fill_idx<<<148 * 8, 256>>>(idx, count, hot, 0);
    report("smem 24K hot (all hit)", hot, 0, time_ms([&] { gather_smem<<<sms, 1024, hot * 8>>>(idx, count, table, hot, out); }));
    fill_idx<<<148 * 8, 256>>>(idx, count, 1u << 26, 1);
    report("smem 24K hot + global (skew)", 1ull << 26, 1,
           time_ms([&] { gather_smem<<<sms, 1024, hot * 8>>>(idx, count, table, hot, out); }));
    report("ldg only (skew), 1024thr", 1ull << 26, 1,
           time_ms([&] { gather_smem<<<sms, 1024, hot * 8>>>(idx, count, table, 0, out); }));
  }
  // DSMEM hot set over clusters of 2/4/8 CTAs
  for (int cs : {2, 4, 8}) {
    const uint32_t per_cta = 24 * 1024;
    CK(cudaFuncSetAttribute(gather_dsmem, cudaFuncAttributeMaxDynamicSharedMemorySize, per_cta * 8));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((sms / cs) * cs);
    cfg.blockDim = dim3(1024);
    cfg.dynamicSmemBytes = per_cta * 8;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const uint32_t hot = per_cta * cs;
    fill_idx<<<148 * 8, 256>>>(idx, count, hot, 0);
    char nm[64];
    snprintf(nm, sizeof nm, "dsmem cluster=%d (all hit)", cs);
    report(nm, hot, 0, time_ms([&] { CK(cudaLaunchKernelEx(&cfg, gather_dsmem, (const uint32_t *)idx, count, (const double *)table, per_cta, out)); }));
    fill_idx<<<148 * 8, 256>>>(idx, count, 1u << 26, 1);
    snprintf(nm, sizeof nm, "dsmem cluster=%d + global skew", cs);
    report(nm, 1ull << 26, 1, time_ms([&] { CK(cudaLaunchKernelEx(&cfg, gather_dsmem, (const uint32_t *)idx, count, (const double *)table, per_cta, out)); }));
  }
  return 0;
}
