// microbenchmark: FP64 FMA throughput at 12 warps per SM (what the C4 cell kernel has) as a function of the
// number of independent DFMA chains per thread.  nvcc -arch=sm_100a -O3 dfma_ilp.cu -o dfma_ilp
#include <cstdio>
#include <cuda_runtime.h>
template <int NCH>
__global__ void __launch_bounds__(1024) chains(double* out, int iters, double a, double b) {
  double v[NCH];
#pragma unroll
  for (int k = 0; k < NCH; ++k) v[k] = threadIdx.x + k;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u)
#pragma unroll
      for (int k = 0; k < NCH; ++k) v[k] = fma(v[k], a, b);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < NCH; ++k) s += v[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NCH>
void run(int threads) {
  double* out;
  cudaMalloc(&out, 148 * 1024 * 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 4096;
  chains<NCH><<<148, threads>>>(out, 64, 1.0000001, 1e-9);
  cudaEventRecord(e0);
  chains<NCH><<<148, threads>>>(out, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double fl = 2.0 * 16 * NCH * iters * 148.0 * threads;
  printf("threads/SM %4d chains/thread %2d: %.2f TFLOP/s\n", threads, NCH, fl / ms / 1e9);
  cudaFree(out);
}
int main() {
  for (int threads : {384, 512, 768, 1024}) {
    if (threads <= 384) {
      run<1>(threads); run<2>(threads); run<3>(threads); run<4>(threads); run<6>(threads); run<8>(threads);
    } else {
      run<1>(threads); run<2>(threads); run<4>(threads);
    }
  }
  return 0;
}
