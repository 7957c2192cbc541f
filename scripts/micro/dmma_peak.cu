// microbenchmark: FP64 tensor-core (DMMA) throughput on B200 next to the DFMA pipe, and whether the two overlap.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 dmma_peak.cu -o dmma_peak
// Variants: mma.sync m8n8k4 and m16n8k16 (f64), NACC independent accumulator tiles per warp; DFMA chains alone;
// both in the same warp (interleaved); both in the same CTA (alternate warps).
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, "
      "{%0,%1,%2,%3};"
      : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]),
        "d"(b[2]), "d"(b[3]));
}

// MODE 0: m8n8k4 only, 1: m16n8k16 only, 2: DFMA only, 3: m8n8k4 + DFMA in every warp, 4: even warps DMMA, odd warps DFMA
template <int MODE, int NACC>
__global__ void __launch_bounds__(1024) k(double* out, int iters, double a0, double b0) {
  double c2[NACC][2], c4[NACC][4], v[NACC];
  double a8[8], b4[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) a8[i] = a0 + i * 1e-9;
#pragma unroll
  for (int i = 0; i < 4; ++i) b4[i] = b0 + i * 1e-9;
#pragma unroll
  for (int kk = 0; kk < NACC; ++kk) {
    c2[kk][0] = c2[kk][1] = threadIdx.x * 1e-3 + kk;
#pragma unroll
    for (int i = 0; i < 4; ++i) c4[kk][i] = threadIdx.x * 1e-3 + kk + i;
    v[kk] = threadIdx.x + kk;
  }
  const bool dm = MODE == 0 || MODE == 1 || MODE == 3 || (MODE == 4 && ((threadIdx.x >> 5) & 1) == 0);
  const bool df = MODE == 2 || MODE == 3 || (MODE == 4 && ((threadIdx.x >> 5) & 1) == 1);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (dm) {
#pragma unroll
        for (int kk = 0; kk < NACC; ++kk) {
          if (MODE == 1)
            dmma16816(c4[kk], a8, b4);
          else
            dmma884(c2[kk], a8[0], b4[0]);
        }
      }
      if (df) {
#pragma unroll
        for (int kk = 0; kk < NACC; ++kk) v[kk] = fma(v[kk], a0, b0);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int kk = 0; kk < NACC; ++kk) s += c2[kk][0] + c2[kk][1] + c4[kk][0] + c4[kk][1] + c4[kk][2] + c4[kk][3] + v[kk];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int NACC>
void run(int threads, const char* name) {
  double* out;
  cudaMalloc(&out, 148 * 1024 * 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 2048;
  k<MODE, NACC><<<148, threads>>>(out, 64, 1.0000001, 1e-9);
  cudaEventRecord(e0);
  k<MODE, NACC><<<148, threads>>>(out, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double warps = 148.0 * threads / 32.0;
  const double mma_fl = (MODE == 1 ? 4096.0 : 512.0) * 8 * NACC * iters;  // per warp
  const double fma_fl = 2.0 * 32 * 8 * NACC * iters;                      // per warp
  double dm_w = (MODE == 0 || MODE == 1 || MODE == 3) ? warps : (MODE == 4 ? warps / 2 : 0);
  double df_w = (MODE == 2 || MODE == 3) ? warps : (MODE == 4 ? warps / 2 : 0);
  printf("%-28s threads/SM %4d acc %d: DMMA %.2f TFLOP/s  DFMA %.2f TFLOP/s  (%.3f ms) %s\n", name, threads, NACC,
         dm_w * mma_fl / ms / 1e9, df_w * fma_fl / ms / 1e9, ms, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  for (int threads : {128, 256, 512, 1024}) {
    run<0, 1>(threads, "m8n8k4");
    run<0, 4>(threads, "m8n8k4");
    run<1, 1>(threads, "m16n8k16");
    run<1, 4>(threads, "m16n8k16");
    run<2, 4>(threads, "dfma");
    run<2, 8>(threads, "dfma");
    run<3, 4>(threads, "m8n8k4+dfma same warp");
    run<4, 4>(threads, "m8n8k4 / dfma alt warps");
  }
  return 0;
}
