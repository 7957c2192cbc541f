// microbenchmark: every CTA streams its own L2-resident buffer repeatedly (what an assembled
// per-point operator kept in L2 would cost).  nvcc -arch=sm_100a -O3 l2bw.cu -o l2bw
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(384) stream(const double2* __restrict__ buf, size_t per_cta, int reps, double* out) {
  const double2* b = buf + (size_t)blockIdx.x * per_cta;
  double acc = 0;
  for (int r = 0; r < reps; ++r)
    for (size_t i = threadIdx.x; i < per_cta; i += blockDim.x * 4) {
      double2 v0 = b[i], v1 = (i + blockDim.x < per_cta) ? b[i + blockDim.x] : double2{0, 0};
      double2 v2 = (i + 2 * blockDim.x < per_cta) ? b[i + 2 * blockDim.x] : double2{0, 0};
      double2 v3 = (i + 3 * blockDim.x < per_cta) ? b[i + 3 * blockDim.x] : double2{0, 0};
      acc += v0.x + v0.y + v1.x + v1.y + v2.x + v2.y + v3.x + v3.y;
    }
  if (acc == 12345.678) out[0] = acc;
}
int main() {
  for (int kb : {283, 553}) {
    for (int ctas_per_sm : {1, 2}) {
      int nb = 148 * ctas_per_sm;
      size_t per = (size_t)kb * 1024 / 16;
      double2* buf; double* out;
      cudaMalloc(&buf, per * 16 * nb); cudaMalloc(&out, 8); cudaMemset(buf, 0, per * 16 * nb);
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      int reps = 200;
      stream<<<nb, 384>>>(buf, per, 5, out);
      cudaEventRecord(e0); stream<<<nb, 384>>>(buf, per, reps, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double bytes = (double)per * 16 * nb * reps;
      printf("%d KB/CTA x %d CTAs (%.0f MB total): %.2f TB/s, %.2f us per pass\n", kb, nb, per * 16.0 * nb / 1e6, bytes / ms / 1e9, ms * 1e3 / reps);
      cudaFree(buf); cudaFree(out);
    }
  }
  return 0;
}
