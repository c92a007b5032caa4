// FP64 tensor-core (mma.sync m8n8k4.f64) issue rate on B200, next to the vector DFMA rate (scripts/microbench.cu).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/microbench_dmma scripts/microbench_dmma.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dmma(double* out, int iters) {
  double c[8][2];
  for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x * 1e-9 + i; c[i][1] = i; }
  double a = 1.0000001 + threadIdx.x * 1e-12, b = 0.9999999;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  double s = 0; for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
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
template <typename F> float timeit(F f, int rep) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); for (int i = 0; i < rep; ++i) f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms / rep;
}
int main() {
  double* out; cudaMalloc(&out, 148 * 16 * 1024 * 8);
  const int iters = 4096;
  for (int warps = 4; warps <= 64; warps *= 2) {
    const int ctas = 148 * (warps >= 8 ? warps / 8 : 1), thr = warps >= 8 ? 256 : warps * 32;
    float ms = timeit([&] { dmma<<<ctas, thr>>>(out, iters); }, 3);
    double fma_s = (double)ctas * (thr / 32) * 8 * iters * 256.0 / (ms * 1e-3);
    float ms2 = timeit([&] { dfma<<<ctas, thr>>>(out, iters); }, 3);
    double fma2 = (double)ctas * thr * 8 * iters / (ms2 * 1e-3);
    printf("%2d warps/SM: DMMA %.2f T FMA/s (%.1f TFLOP/s)   DFMA %.2f T FMA/s\n", warps, fma_s / 1e12, 2 * fma_s / 1e12, fma2 / 1e12);
  }
  return 0;
}
