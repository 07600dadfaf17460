// DMMA m8n8k4 issue rate with DISTINCT operand registers (the elimination's pattern: S tile (t,u) += W[kk][t] W[kk][u]),
// against the same-operand stream of fp64_micro2.  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
#define DMMA(c, a, b) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b))

template <int MODE> __global__ void k(double *out, long long *cyc, int iters) {
  double c[15][2], w[4][5];
  for (int i = 0; i < 15; i++) c[i][0] = c[i][1] = 0;
  for (int kk = 0; kk < 4; kk++)
    for (int t = 0; t < 5; t++) w[kk][t] = 1.0 + 1e-9 * (threadIdx.x + 7 * kk + 3 * t);
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    if (MODE == 0) { // 60 DMMAs, same operands
#pragma unroll
      for (int i = 0; i < 15; i++)
#pragma unroll
        for (int kk = 0; kk < 4; kk++) DMMA(c[i], w[0][0], w[0][1]);
    } else if (MODE == 1) { // the Schur pattern: 15 tile pairs x 4 k-steps, distinct operands, kk innermost (dependent chains of 4)
      int i = 0;
#pragma unroll
      for (int ta = 0; ta < 5; ta++)
#pragma unroll
        for (int tb = ta; tb < 5; tb++) {
#pragma unroll
          for (int kk = 0; kk < 4; kk++) DMMA(c[i], w[kk][ta], w[kk][tb]);
          i++;
        }
    } else { // same work, kk outermost: 15 independent accumulators between dependent uses
#pragma unroll
      for (int kk = 0; kk < 4; kk++) {
        int i = 0;
#pragma unroll
        for (int ta = 0; ta < 5; ta++)
#pragma unroll
          for (int tb = ta; tb < 5; tb++) {
            DMMA(c[i], w[kk][ta], w[kk][tb]);
            i++;
          }
      }
    }
    // keep the operands changing so that nothing is hoisted
#pragma unroll
    for (int t = 0; t < 5; t++) w[it & 3][t] += 1e-12;
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < 15; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double *out; long long *cyc, h;
  cudaMalloc(&out, sizeof(double) * 148 * 1024 * 2);
  cudaMalloc(&cyc, 64);
  const int iters = 512;
#define RUN(MODE, WARPS, NAME)                                                                 \
  k<MODE><<<148, 32 * 4 * WARPS>>>(out, cyc, iters);                                           \
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);                                              \
  printf("%-44s %d warps/sub-partition: %.1f cycles per DMMA per sub-partition\n", NAME, WARPS, (double)h / iters / 60 / WARPS);
  RUN(0, 1, "same operands") RUN(0, 3, "same operands")
  RUN(1, 1, "Schur pattern, dependent k-steps adjacent") RUN(1, 3, "Schur pattern, dependent k-steps adjacent")
  RUN(2, 1, "Schur pattern, k-step outermost") RUN(2, 3, "Schur pattern, k-step outermost")
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
