// FP64 pipe microbenchmarks, round 2: what a dependent DMMA chain costs, and what a dependent DFMA chain costs while other
// warps of the same SM sub-partition stream DMMAs (the situation of k_seg_eliminate: one warp in its 15-pivot factorisation,
// the others in their mma phase).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/fp64_micro2 tools/fp64_micro2.cu
#include <cstdio>
#include <cuda_runtime.h>
#define DMMA(c, a, b) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b))

template <int ILP> __global__ void k_dmma_chain(double *out, long long *cyc, int iters) {
  double c[8][2];
  for (int i = 0; i < 8; i++) c[i][0] = c[i][1] = 0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) DMMA(c[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
// warps 0..3 of the CTA (one per sub-partition) run a dependent DFMA chain, the others stream DMMAs with 4 independent accumulators
__global__ void k_mixed(double *out, long long *cyc, int iters, int dmma_on) {
  const int wid = threadIdx.x >> 5;
  double x = 1.5 + threadIdx.x * 1e-6;
  double c[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  long long t0 = clock64();
  if (wid < 4) {
    for (int i = 0; i < iters; i++) x = fma(x, 0.999999, 1e-7);
  } else if (dmma_on) {
    for (int it = 0; it < iters / 2; it++) {
#pragma unroll
      for (int i = 0; i < 4; i++) DMMA(c[i], a, b);
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x + c[0][0] + c[1][1] + c[2][0] + c[3][1];
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double *out; long long *cyc, h;
  cudaMalloc(&out, sizeof(double) * 148 * 1024 * 2);
  cudaMalloc(&cyc, 64);
  const int iters = 4096;
#define RUN(ILP, WARPS)                                                                                   \
  k_dmma_chain<ILP><<<148, 32 * 4 * WARPS>>>(out, cyc, iters);                                            \
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);                                                         \
  printf("dmma chain: %d independent accumulators, %d warps/sub-partition: %.1f cycles per DMMA per warp, %.1f cycles per DMMA per sub-partition\n", ILP, WARPS, \
         (double)h / iters / ILP, (double)h / iters / ILP / WARPS);
  RUN(1, 1) RUN(2, 1) RUN(4, 1) RUN(8, 1) RUN(1, 2) RUN(1, 3) RUN(1, 4) RUN(2, 3) RUN(4, 3) RUN(4, 4)
  for (int group : {4, 8, 12, 16}) // one DFMA-chain warp per sub-partition among `group / 4` warps per sub-partition
    for (int on : {0, 1}) {
      k_mixed<<<148, 32 * group>>>(out, cyc, iters, on);
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      printf("dependent DFMA chain, %d warps per sub-partition, others %s: %.1f cycles per DFMA\n", group / 4, on ? "stream DMMA" : "idle", (double)h / iters);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
