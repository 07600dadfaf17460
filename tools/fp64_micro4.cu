// Does alternating DFMA and DMMA instructions cost more FP64-pipe time than the same instructions in blocks?
// (independent instructions, W warps per sub-partition, cycles per iteration of 16 DFMA + 16 DMMA)
#include <cstdio>
#include <cuda_runtime.h>
#define DMMA(c, a, b) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b))
#define DFMA(x, a, b) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(x) : "d"(a), "d"(b))
template <int MODE> __global__ void k(double *out, long long *cyc, int iters) {
  double c[16][2], x[16];
  for (int i = 0; i < 16; i++) { c[i][0] = c[i][1] = 0; x[i] = threadIdx.x; }
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    if (MODE == 0) { // alternate
#pragma unroll
      for (int i = 0; i < 16; i++) { DMMA(c[i], a, b); DFMA(x[i], a, b); }
    } else if (MODE == 1) { // blocks
#pragma unroll
      for (int i = 0; i < 16; i++) DMMA(c[i], a, b);
#pragma unroll
      for (int i = 0; i < 16; i++) DFMA(x[i], a, b);
    } else if (MODE == 2) { // DMMA only
#pragma unroll
      for (int i = 0; i < 16; i++) DMMA(c[i], a, b);
    } else { // DFMA only
#pragma unroll
      for (int i = 0; i < 16; i++) DFMA(x[i], a, b);
    }
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < 16; i++) s += c[i][0] + c[i][1] + x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double *out; long long *cyc, h;
  cudaMalloc(&out, sizeof(double) * 148 * 1024 * 2);
  cudaMalloc(&cyc, 64);
  const int iters = 2048;
#define RUN(MODE, WARPS, NAME)                                     \
  k<MODE><<<148, 32 * 4 * WARPS>>>(out, cyc, iters);               \
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);                  \
  printf("%-28s %d warps/sub-partition: %.1f cycles per iteration (16 DFMA + 16 DMMA) per sub-partition-warp\n", NAME, WARPS, (double)h / iters / WARPS);
  for (int w : {1, 3}) {
    if (w == 1) { RUN(0, 1, "alternating") RUN(1, 1, "blocks of 16") RUN(2, 1, "DMMA only") RUN(3, 1, "DFMA only") }
    else { RUN(0, 3, "alternating") RUN(1, 3, "blocks of 16") RUN(2, 3, "DMMA only") RUN(3, 3, "DFMA only") }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
