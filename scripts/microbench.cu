// Device ceilings that bound the merge kernels: FP64 FMA rate, streaming copy with 256-bit accesses, fp64 RED rate.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/microbench scripts/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma(double* out, int iters) {
  double a[8];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-9 + i;
  const double b = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
  double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
struct __align__(32) d4 { double v[4]; };
__global__ void copy256(const d4* __restrict__ in, d4* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    d4 r; asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3]) : "l"(in + i));
    asm volatile("st.global.v4.f64 [%4], {%0,%1,%2,%3};" :: "d"(r.v[0]), "d"(r.v[1]), "d"(r.v[2]), "d"(r.v[3]), "l"(out + i) : "memory");
  }
}
__global__ void write256(d4* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    asm volatile("st.global.v4.f64 [%4], {%0,%1,%2,%3};" :: "d"(1.0), "d"(2.0), "d"(3.0), "d"(4.0), "l"(out + i) : "memory");
  }
}
__global__ void red64(double* g, size_t n, int spread) {
  size_t i = (blockIdx.x * (size_t)blockDim.x + threadIdx.x);
  atomicAdd(g + (spread ? i % n : (i % 4096)), 1.0);
}
template <typename F> float timeit(F f, int rep) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); for (int i = 0; i < rep; ++i) f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms / rep;
}
int main() {
  double* out; cudaMalloc(&out, 148 * 16 * 1024 * 8);
  const int iters = 4096;
  float ms = timeit([&] { dfma<<<148 * 8, 256>>>(out, iters); }, 5);
  double fma_s = 148.0 * 8 * 256 * 8 * iters / (ms * 1e-3);
  printf("fp64 FMA: %.2f T FMA/s (%.1f TFLOP/s), %.2f FMA/clk/SM at 1.965 GHz\n", fma_s / 1e12, 2 * fma_s / 1e12, fma_s / 148 / 1.965e9);
  size_t n = (size_t)1 << 28;  // 8 GiB per buffer
  d4 *x, *y; cudaMalloc(&x, n * 32); cudaMalloc(&y, n * 32); cudaMemset(x, 0, n * 32);
  ms = timeit([&] { copy256<<<148 * 16, 256>>>(x, y, n); }, 5);
  printf("copy 256-bit: %.1f GB/s (read+write)\n", 2.0 * n * 32 / ms / 1e6);
  ms = timeit([&] { write256<<<148 * 16, 256>>>(y, n); }, 5);
  printf("write-only 256-bit: %.1f GB/s\n", 1.0 * n * 32 / ms / 1e6);
  size_t na = (size_t)1 << 26;
  ms = timeit([&] { red64<<<(unsigned)(na / 256), 256>>>((double*)x, na, 1); }, 5);
  printf("fp64 RED spread over 512 MiB: %.1f G atomics/s\n", na / ms / 1e6);
  ms = timeit([&] { red64<<<(unsigned)(na / 256), 256>>>((double*)x, na, 0); }, 5);
  printf("fp64 RED onto 4096 hot addresses: %.1f G atomics/s\n", na / ms / 1e6);
  return 0;
}
