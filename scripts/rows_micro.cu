// What the FP64 pipe sustains on the instruction mix of the rows kernel's inner loop (two particles x four sites per trip:
// 32 DFMA + 8 DMUL), at the occupancies the kernel can have.  Steps: m0 registers only, m1 + the broadcast row loads from
// shared memory, m2 + accumulators in shared memory and the mantissa/exponent fold (= the kernel's loop).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/rows_micro scripts/rows_micro.cu && scripts/rows_micro
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kCoef = 20, kR = 32, kT = 256;

__device__ __forceinline__ void unit_trip(const double* rowp, double* pp, const double (&Lb)[4][4], const double (&x0)[4]) {
  const double2 a0 = *reinterpret_cast<const double2*>(rowp), a1 = *reinterpret_cast<const double2*>(rowp + 2);
  const double2 b0 = *reinterpret_cast<const double2*>(rowp + kCoef), b1 = *reinterpret_cast<const double2*>(rowp + kCoef + 2);
  const double pa = pp[0], pb = pp[kT];
  double xa[4], xb[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) { xa[q] = fma(a0.x, Lb[q][0], x0[q]); xb[q] = fma(b0.x, Lb[q][0], x0[q]); }
#pragma unroll
  for (int q = 0; q < 4; ++q) { xa[q] = fma(a0.y, Lb[q][1], xa[q]); xb[q] = fma(b0.y, Lb[q][1], xb[q]); }
#pragma unroll
  for (int q = 0; q < 4; ++q) { xa[q] = fma(a1.x, Lb[q][2], xa[q]); xb[q] = fma(b1.x, Lb[q][2], xb[q]); }
#pragma unroll
  for (int q = 0; q < 4; ++q) { xa[q] = fma(a1.y, Lb[q][3], xa[q]); xb[q] = fma(b1.y, Lb[q][3], xb[q]); }
  pp[0] = pa * ((xa[0] * xa[1]) * (xa[2] * xa[3]));
  pp[kT] = pb * ((xb[0] * xb[1]) * (xb[2] * xb[3]));
}

template <int MODE>
__global__ void __launch_bounds__(256, 2) micro(const double* __restrict__ in, double* __restrict__ out, int trips, int reps) {
  extern __shared__ __align__(16) unsigned char smem[];
  double* sC = reinterpret_cast<double*>(smem);
  double* s_prod = sC + kR * kCoef;
  int* s_exp = reinterpret_cast<int*>(s_prod + kR * kT);
  const int tid = threadIdx.x;
  for (int i = tid; i < kR * kCoef; i += kT) sC[i] = in[i] + 0.5;
  for (int j = 0; j < kR; ++j) { s_prod[j * kT + tid] = 1.0; s_exp[j * kT + tid] = 0; }
  double Lb[4][4], x0[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    x0[q] = in[1000 + q] * 1e-30;
#pragma unroll
    for (int m = 0; m < 4; ++m) Lb[q][m] = in[tid * 16 + q * 4 + m] + 0.25;
  }
  __syncthreads();
  double ra = 1.0, rb = 1.0;
  int ea = 0, eb = 0;
  for (int r = 0; r < reps; ++r) {
    const double* rowp = sC + 4 * (r & 3);
    double* pp = s_prod + tid;
    int* pe = s_exp + tid;
    double2 a0 = *reinterpret_cast<const double2*>(rowp), a1 = *reinterpret_cast<const double2*>(rowp + 2);
    double2 b0 = *reinterpret_cast<const double2*>(rowp + kCoef), b1 = *reinterpret_cast<const double2*>(rowp + kCoef + 2);
    if (MODE == 4) {
#pragma unroll 2
      for (int i = 0; i < trips; ++i) unit_trip(rowp + 2 * kCoef * i, pp + 2 * kT * i, Lb, x0);
    } else
    for (int i = 0; i < trips; ++i, pp += 2 * kT, pe += 2 * kT, rowp += 2 * kCoef) {
      if (MODE >= 1) {
        a0 = *reinterpret_cast<const double2*>(rowp); a1 = *reinterpret_cast<const double2*>(rowp + 2);
        b0 = *reinterpret_cast<const double2*>(rowp + kCoef); b1 = *reinterpret_cast<const double2*>(rowp + kCoef + 2);
      } else {
        a0.x = ra * 0.999; b0.x = rb * 0.999;   // (depends on the previous trip: nothing can be hoisted)
      }
      double pa = ra, pb = rb;
      int xa_e = ea, xb_e = eb;
      if (MODE >= 2) { pa = pp[0]; pb = pp[kT]; }
      if (MODE == 2) { xa_e = pe[0]; xb_e = pe[kT]; }
      double xa[4], xb[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) { xa[q] = fma(a0.x, Lb[q][0], x0[q]); xb[q] = fma(b0.x, Lb[q][0], x0[q]); }
#pragma unroll
      for (int q = 0; q < 4; ++q) { xa[q] = fma(a0.y, Lb[q][1], xa[q]); xb[q] = fma(b0.y, Lb[q][1], xb[q]); }
#pragma unroll
      for (int q = 0; q < 4; ++q) { xa[q] = fma(a1.x, Lb[q][2], xa[q]); xb[q] = fma(b1.x, Lb[q][2], xb[q]); }
#pragma unroll
      for (int q = 0; q < 4; ++q) { xa[q] = fma(a1.y, Lb[q][3], xa[q]); xb[q] = fma(b1.y, Lb[q][3], xb[q]); }
      const double va = (xa[0] * xa[1]) * (xa[2] * xa[3]), vb = (xb[0] * xb[1]) * (xb[2] * xb[3]);
      if (MODE >= 3) {
        pp[0] = pa * va; pp[kT] = pb * vb;
      } else if (MODE >= 2) {
        const int ha = __double2hiint(va), hb = __double2hiint(vb);
        const unsigned sa = (unsigned)ha >> 20, sb = (unsigned)hb >> 20;
        double na = pa * __hiloint2double((ha & 0x000fffff) | 0x3ff00000, __double2loint(va));
        double nb = pb * __hiloint2double((hb & 0x000fffff) | 0x3ff00000, __double2loint(vb));
        if ((sa - 1u) >= 0x7feu) na = __longlong_as_double(0x7ff8000000000000ll);
        if ((sb - 1u) >= 0x7feu) nb = __longlong_as_double(0x7ff8000000000000ll);
        pp[0] = na; pp[kT] = nb;
        pe[0] = xa_e + ((int)sa - 1023); pe[kT] = xb_e + ((int)sb - 1023);
      } else {
        ra = pa * va + 1.0; rb = pb * vb + 1.0;   // stays near 1
        ra = ra > 2.0 ? 1.0 : ra; rb = rb > 2.0 ? 1.0 : rb;
      }
    }
  }
  double s = ra + rb + ea + eb;
  for (int j = 0; j < kR; ++j) s += s_prod[j * kT + tid] + s_exp[j * kT + tid];
  out[blockIdx.x * kT + tid] = s;
}

template <int MODE>
void run(const double* in, double* out, int ctas_per_sm) {
  // shared memory sized so that exactly ctas_per_sm CTAs fit (227 KB per SM); registers allow two
  const int base = kR * kCoef * 8 + kR * kT * 12;
  int smem = ctas_per_sm == 1 ? 120 * 1024 : base;
  cudaFuncSetAttribute(micro<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int trips = 16, reps = 400;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  micro<MODE><<<148 * ctas_per_sm, kT, smem>>>(in, out, trips, 10);
  cudaEventRecord(a);
  micro<MODE><<<148 * ctas_per_sm, kT, smem>>>(in, out, trips, reps);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double ops = 148.0 * ctas_per_sm * kT * (double)reps * trips * 40;
  printf("mode %d, %d CTA/SM (%d warps/scheduler): %.2f T FP64 op/s = %.3f of 18.2   [%s]\n", MODE, ctas_per_sm, 2 * ctas_per_sm,
         ops / ms / 1e9, ops / ms / 1e9 / 18.2, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  double *in, *out;
  cudaMalloc(&in, 1 << 20); cudaMemset(in, 0, 1 << 20); cudaMalloc(&out, 148 * 4 * kT * 8);
  run<0>(in, out, 1); run<0>(in, out, 2);
  run<1>(in, out, 1); run<1>(in, out, 2);
  run<2>(in, out, 1); run<2>(in, out, 2);
  run<3>(in, out, 1); run<3>(in, out, 2);
  run<4>(in, out, 1); run<4>(in, out, 2);
  return 0;
}
