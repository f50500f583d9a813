// scripts/epilogue_bench.cu -- microbenchmark of the SELL row epilogue (exploration tool, not product code).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_bin/epilogue_bench scripts/epilogue_bench.cu
// Question (profiles/r02_epilogue.md): the epilogue of a 1/8 partition (3.4 M rows, 67 MB of reads) takes 0.26 ms under
// ncu -- the time the single-GPU epilogue needs for 27 M rows.  Which part of it is the fixed cost?
// Variants: V0 = the product's code shape (asm-volatile cache-hinted loads, one row per thread per trip),
// V1 = same with plain loads/stores, V2 = R rows per thread per trip with all loads issued first (__ldcs/__stcs),
// V3 = V2 + no rank RMW hazard (read old from rank, write new), grid sweep.
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

__device__ __forceinline__ uint64_t pol_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ double ldh_f64(const double *p, uint64_t pol) {
  double v;
  asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ uint32_t ldh_u32(const uint32_t *p, uint64_t pol) {
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void sth_f64(double *p, double v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
}

struct Args {
  const double *sums;
  double *rank;
  const uint32_t *outdeg;
  double *contrib;
  uint64_t rows;
  double base, damping;
  unsigned long long *diff_bits;
};

__device__ __forceinline__ void block_max(double m, unsigned long long *out) {
  for (int o = 16; o > 0; o >>= 1) {
    double t = __shfl_xor_sync(0xffffffffu, m, o);
    if (t > m) m = t;
  }
  __shared__ double wm[32];
  if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (unsigned w = 1; w < blockDim.x / 32; ++w)
      if (wm[w] > m) m = wm[w];
    if (m > 0.0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
  }
}

// V0: the product's shape
__global__ void __launch_bounds__(256) epi_v0(Args a) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t pol = pol_evict_first();
  double mx = 0.0;
  for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < a.rows; r += stride) {
    const double acc = ldh_f64(a.sums + r, pol);
    const double next = __dadd_rn(a.base, __dmul_rn(a.damping, acc));
    const double prev = ldh_f64(a.rank + r, pol);
    sth_f64(a.rank + r, next, pol);
    const uint32_t od = ldh_u32(a.outdeg + r, pol);
    if (od != 0) sth_f64(a.contrib + r, __ddiv_rn(next, (double)od), pol);
    const double d = fabs(next - prev);
    if (d > mx) mx = d;
  }
  block_max(mx, a.diff_bits);
}

// V1: plain C++ loads / stores, same shape
__global__ void __launch_bounds__(256) epi_v1(Args a) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  double mx = 0.0;
  for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < a.rows; r += stride) {
    const double acc = a.sums[r];
    const double next = __dadd_rn(a.base, __dmul_rn(a.damping, acc));
    const double prev = a.rank[r];
    a.rank[r] = next;
    const uint32_t od = a.outdeg[r];
    if (od != 0) a.contrib[r] = __ddiv_rn(next, (double)od);
    const double d = fabs(next - prev);
    if (d > mx) mx = d;
  }
  block_max(mx, a.diff_bits);
}

// V2: R rows per thread per trip, every load issued before the first use; streaming cache operators
template <int R>
__global__ void __launch_bounds__(256) epi_v2(Args a) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  double mx = 0.0;
  for (uint64_t r0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r0 < a.rows; r0 += stride * R) {
    double acc[R], prev[R];
    uint32_t od[R];
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const uint64_t r = r0 + j * stride;
      const bool ok = r < a.rows;
      acc[j] = ok ? __ldcs(a.sums + r) : 0.0;
      prev[j] = ok ? __ldcs(a.rank + r) : 0.0;
      od[j] = ok ? __ldcs(a.outdeg + r) : 0u;
    }
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const uint64_t r = r0 + j * stride;
      if (r < a.rows) {
        const double next = __dadd_rn(a.base, __dmul_rn(a.damping, acc[j]));
        __stcs(a.rank + r, next);
        if (od[j] != 0) __stcs(a.contrib + r, __ddiv_rn(next, (double)od[j]));
        const double d = fabs(next - prev[j]);
        if (d > mx) mx = d;
      }
    }
  }
  block_max(mx, a.diff_bits);
}

// V4: V0's hinted asm loads but hoisted: all three loads issued before the first use (asm non-volatile)
__device__ __forceinline__ double ldn_f64(const double *p, uint64_t pol) {
  double v;
  asm("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ uint32_t ldn_u32(const uint32_t *p, uint64_t pol) {
  uint32_t v;
  asm("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}
template <int R>
__global__ void __launch_bounds__(256) epi_v4(Args a) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t pol = pol_evict_first();
  double mx = 0.0;
  for (uint64_t r0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r0 < a.rows; r0 += stride * R) {
    double acc[R], prev[R];
    uint32_t od[R];
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const uint64_t r = r0 + j * stride;
      const bool ok = r < a.rows;
      acc[j] = ok ? ldn_f64(a.sums + r, pol) : 0.0;
      prev[j] = ok ? ldn_f64(a.rank + r, pol) : 0.0;
      od[j] = ok ? ldn_u32(a.outdeg + r, pol) : 0u;
    }
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const uint64_t r = r0 + j * stride;
      if (r < a.rows) {
        const double next = __dadd_rn(a.base, __dmul_rn(a.damping, acc[j]));
        sth_f64(a.rank + r, next, pol);
        if (od[j] != 0) sth_f64(a.contrib + r, __ddiv_rn(next, (double)od[j]), pol);
        const double d = fabs(next - prev[j]);
        if (d > mx) mx = d;
      }
    }
  }
  block_max(mx, a.diff_bits);
}

__global__ void fill(double *sums, double *rank, uint32_t *od, uint64_t n) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    sums[i] = 1e-8 * (double)(i % 1000 + 1);
    rank[i] = 1.5e-8;
    od[i] = (i % 5 == 0) ? 0u : (uint32_t)(i % 37 + 1);
  }
}
__global__ void flush(double *buf, uint64_t n) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) buf[i] = (double)i;
}

template <typename F>
void run(const char *name, F launch, Args a, double *flush_buf, uint64_t flush_n, int grid) {
  float cold = 0, warm = 0;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int rep = 0; rep < 5; ++rep) {
    flush<<<148 * 8, 256>>>(flush_buf, flush_n);
    CK(cudaEventRecord(e0));
    launch(grid, a);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep == 0 || ms < cold) cold = ms;
  }
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(e0));
    launch(grid, a);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep == 0 || ms < warm) warm = ms;
  }
  CK(cudaGetLastError());
  const double bytes = a.rows * (8.0 + 8 + 4 + 8 + 8 * 0.8);
  printf("%-14s rows %9llu grid %5d  cold %.4f ms (%.0f GB/s)  back-to-back %.4f ms\n", name,
         (unsigned long long)a.rows, grid, cold, bytes / cold / 1e6, warm);
}

int main() {
  const uint64_t big = 27023736, small = 3377984;
  double *sums, *rank, *contrib, *fl;
  uint32_t *od;
  unsigned long long *diff;
  const uint64_t flush_n = 64ull << 20;  // 512 MB > L2
  CK(cudaMalloc(&sums, big * 8));
  CK(cudaMalloc(&rank, big * 8));
  CK(cudaMalloc(&contrib, big * 8));
  CK(cudaMalloc(&od, big * 4));
  CK(cudaMalloc(&fl, flush_n * 8));
  CK(cudaMalloc(&diff, 8));
  CK(cudaMemset(diff, 0, 8));
  fill<<<148 * 8, 256>>>(sums, rank, od, big);
  CK(cudaDeviceSynchronize());
  for (uint64_t rows : {small, big}) {
    Args a{sums, rank, od, contrib, rows, 0.15 / 6.7e7, 0.85, diff};
    for (int per_sm : {6, 2, 1, 12, 24}) {
      const int grid = 148 * per_sm;
      run("v0 hinted", [](int g, Args x) { epi_v0<<<g, 256>>>(x); }, a, fl, flush_n, grid);
    }
    const int grid = 148 * 6;
    run("v1 plain", [](int g, Args x) { epi_v1<<<g, 256>>>(x); }, a, fl, flush_n, grid);
    run("v2 R=2 cs", [](int g, Args x) { epi_v2<2><<<g, 256>>>(x); }, a, fl, flush_n, grid);
    run("v2 R=4 cs", [](int g, Args x) { epi_v2<4><<<g, 256>>>(x); }, a, fl, flush_n, grid);
    run("v2 R=4 cs", [](int g, Args x) { epi_v2<4><<<g, 256>>>(x); }, a, fl, flush_n, 148 * 2);
    run("v2 R=8 cs", [](int g, Args x) { epi_v2<8><<<g, 256>>>(x); }, a, fl, flush_n, 148 * 2);
    run("v4 R=1 hint", [](int g, Args x) { epi_v4<1><<<g, 256>>>(x); }, a, fl, flush_n, grid);
    run("v4 R=4 hint", [](int g, Args x) { epi_v4<4><<<g, 256>>>(x); }, a, fl, flush_n, grid);
    run("v4 R=4 hint", [](int g, Args x) { epi_v4<4><<<g, 256>>>(x); }, a, fl, flush_n, 148 * 2);
  }
  return 0;
}
