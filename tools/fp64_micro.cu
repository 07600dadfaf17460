// Microbenchmark (design input, not product): FP64 issue rates on B200 that decide how the 15x15 block
// kernels are mapped -- DFMA from registers, DFMA fed by broadcast LDS.64 / LDS.128, and the legacy
// FP64 tensor path (mma.sync m8n8k4 f64 -> DMMA).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/fp64_micro tools/fp64_micro.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double *out, int iters) {
  double a[8], x = threadIdx.x * 1e-9, y = 1.0000001;
  for (int i = 0; i < 8; i++) a[i] = i + x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = fma(a[i], y, x);
  }
  double s = 0;
  for (int i = 0; i < 8; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DFMA with one operand from a broadcast shared-memory load (8 B / 16 B per load)
template <int VEC>
__global__ void k_dfma_lds(double *out, int iters) {
  __shared__ double sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = 1.0 + i * 1e-9;
  __syncthreads();
  double a[8], x = threadIdx.x * 1e-9;
  for (int i = 0; i < 8; i++) a[i] = i + x;
  for (int it = 0; it < iters; it++) {
    const int base = (it * 8) & 1016;
    if (VEC == 1) {
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = fma(a[i], sm[base + i], x);
    } else {
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        const double2 v = *reinterpret_cast<const double2 *>(&sm[base + i]);
        a[i] = fma(a[i], v.x, x);
        a[i + 1] = fma(a[i + 1], v.y, x);
      }
    }
  }
  double s = 0;
  for (int i = 0; i < 8; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 4x4 register tile fed by per-lane LDS.128 (the Schur-update inner loop shape)
__global__ void k_tile44(double *out, int iters) {
  __shared__ __align__(16) double sm[16 * 40];
  for (int i = threadIdx.x; i < 640; i += blockDim.x) sm[i] = 1.0 + i * 1e-9;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int ta = lane % 10, tb = (lane * 7) % 10;
  double acc[16];
  for (int i = 0; i < 16; i++) acc[i] = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < 15; k++) {
      const double2 a0 = *reinterpret_cast<const double2 *>(&sm[40 * k + 4 * ta]);
      const double2 a1 = *reinterpret_cast<const double2 *>(&sm[40 * k + 4 * ta + 2]);
      const double2 b0 = *reinterpret_cast<const double2 *>(&sm[40 * k + 4 * tb]);
      const double2 b1 = *reinterpret_cast<const double2 *>(&sm[40 * k + 4 * tb + 2]);
      const double av[4] = {a0.x, a0.y, a1.x, a1.y}, bv[4] = {b0.x, b0.y, b1.x, b1.y};
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[4 * i + j] = fma(av[i], bv[j], acc[4 * i + j]);
    }
  }
  double s = 0;
  for (int i = 0; i < 16; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma(double *out, int iters) {
  double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; it++) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[0]), "+d"(c0[1]) : "d"(a), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[0]), "+d"(c1[1]) : "d"(a), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c2[0]), "+d"(c2[1]) : "d"(a), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c3[0]), "+d"(c3[1]) : "d"(a), "d"(b));
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c0[1] + c1[0] + c1[1] + c2[0] + c2[1] + c3[0] + c3[1];
}

// dependent chain latencies: DFMA, sqrt, divide, rsqrt (cycles per op, one warp)
__global__ void k_latency(double *out, long long *cyc) {
  double x = 1.5 + threadIdx.x * 1e-6;
  long long t0 = clock64();
  for (int i = 0; i < 256; i++) x = fma(x, 0.999999, 1e-7);
  long long t1 = clock64();
  for (int i = 0; i < 256; i++) x = sqrt(x + 1.0);
  long long t2 = clock64();
  for (int i = 0; i < 256; i++) x = 1.0 / (x + 0.5);
  long long t3 = clock64();
  for (int i = 0; i < 256; i++) x = rsqrt(x + 1.0);
  long long t4 = clock64();
  for (int i = 0; i < 256; i++) x = __shfl_sync(0xffffffffu, x, (i + 1) & 31) + 1e-9;
  long long t5 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) {
    cyc[0] = (t1 - t0) / 256; cyc[1] = (t2 - t1) / 256; cyc[2] = (t3 - t2) / 256; cyc[3] = (t4 - t3) / 256; cyc[4] = (t5 - t4) / 256;
  }
}

template <typename F> float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 4; r++) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  double *out; long long *cyc;
  const int blocks = 148 * 8, threads = 256, iters = 20000;
  cudaMalloc(&out, sizeof(double) * blocks * threads);
  cudaMalloc(&cyc, 64);
  const double lanes = (double)blocks * threads;
  float ms = time_ms([&] { k_dfma<<<blocks, threads>>>(out, iters); });
  printf("dfma_reg      %.2f TFLOP/s\n", lanes * iters * 8 * 2 / (ms * 1e-3) / 1e12);
  ms = time_ms([&] { k_dfma_lds<1><<<blocks, threads>>>(out, iters); });
  printf("dfma_lds64    %.2f TFLOP/s (1 broadcast LDS.64 per DFMA)\n", lanes * iters * 8 * 2 / (ms * 1e-3) / 1e12);
  ms = time_ms([&] { k_dfma_lds<2><<<blocks, threads>>>(out, iters); });
  printf("dfma_lds128   %.2f TFLOP/s (1 broadcast LDS.128 per 2 DFMA)\n", lanes * iters * 8 * 2 / (ms * 1e-3) / 1e12);
  ms = time_ms([&] { k_tile44<<<blocks, threads>>>(out, iters / 20); });
  printf("tile4x4       %.2f TFLOP/s (4 per-lane LDS.128 per 16 DFMA)\n", lanes * (iters / 20) * 15 * 16 * 2 / (ms * 1e-3) / 1e12);
  ms = time_ms([&] { k_dmma<<<blocks, threads>>>(out, iters); });
  printf("dmma_m8n8k4   %.2f TFLOP/s\n", (lanes / 32) * iters * 4 * 512.0 / (ms * 1e-3) / 1e12);
  for (int w : {1, 2, 4, 8}) {
    ms = time_ms([&] { k_dfma<<<148, 32 * 4 * w>>>(out, iters); });
    printf("dfma_reg %d warps/quadrant: %.2f TFLOP/s\n", w, 148.0 * 128 * w * iters * 16 / (ms * 1e-3) / 1e12);
  }
  k_latency<<<1, 32>>>(out, cyc);
  long long h[5];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("latency cycles: dfma %lld sqrt %lld div %lld rsqrt %lld shfl+dadd %lld\n", h[0], h[1], h[2], h[3], h[4]);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
