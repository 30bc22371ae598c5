// FP64 tensor-pipe issue rate on this GPU for the mma.sync f64 shapes (independent accumulators, no memory traffic):
// decides whether the block SpTRSV kernels are bound by the DMMA pipe.   nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
template <int SHAPE>
__global__ void k(double *out, int iters) {
  double c[8][4];
  for (int q = 0; q < 8; ++q)
    for (int i = 0; i < 4; ++i) c[q][i] = 0.0;
  double a[8], b[4];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  for (int i = 0; i < 4; ++i) b[i] = threadIdx.x * 1e-4 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (SHAPE == 0)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[q][0]), "+d"(c[q][1]) : "d"(a[0]), "d"(b[0]));
      else if (SHAPE == 1)
        asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+d"(c[q][0]), "+d"(c[q][1]), "+d"(c[q][2]), "+d"(c[q][3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
      else if (SHAPE == 2)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+d"(c[q][0]), "+d"(c[q][1]), "+d"(c[q][2]), "+d"(c[q][3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
      else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                     : "+d"(c[q][0]), "+d"(c[q][1]), "+d"(c[q][2]), "+d"(c[q][3])
                     : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    }
  }
  double s = 0;
  for (int q = 0; q < 8; ++q)
    for (int i = 0; i < 4; ++i) s += c[q][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// plain FP64 FMA rate for comparison
__global__ void kf(double *out, int iters) {
  double c[16];
  for (int q = 0; q < 16; ++q) c[q] = q;
  const double a = threadIdx.x * 1e-3, b = 1.0000001;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int q = 0; q < 16; ++q) c[q] = fma(c[q], b, a);
  double s = 0;
  for (int q = 0; q < 16; ++q) s += c[q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F>
float timeit(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  f();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}
int main() {
  double *out;
  cudaMalloc(&out, 148 * 8 * 512 * sizeof(double));
  const int iters = 20000, blocks = 148 * 4, threads = 256;
  const double flops[4] = {512, 1024, 2048, 4096};
  const char *names[4] = {"m8n8k4", "m16n8k4", "m16n8k8", "m16n8k16"};
  float ms;
  ms = timeit([&] { k<0><<<blocks, threads>>>(out, iters); });
  printf("%-9s %8.3f ms  %7.2f TFLOP/s\n", names[0], ms, flops[0] * 8.0 * iters * blocks * (threads / 32) / ms * 1e-9);
  ms = timeit([&] { k<1><<<blocks, threads>>>(out, iters); });
  printf("%-9s %8.3f ms  %7.2f TFLOP/s\n", names[1], ms, flops[1] * 8.0 * iters * blocks * (threads / 32) / ms * 1e-9);
  ms = timeit([&] { k<2><<<blocks, threads>>>(out, iters); });
  printf("%-9s %8.3f ms  %7.2f TFLOP/s\n", names[2], ms, flops[2] * 8.0 * iters * blocks * (threads / 32) / ms * 1e-9);
  ms = timeit([&] { k<3><<<blocks, threads>>>(out, iters); });
  printf("%-9s %8.3f ms  %7.2f TFLOP/s\n", names[3], ms, flops[3] * 8.0 * iters * blocks * (threads / 32) / ms * 1e-9);
  ms = timeit([&] { kf<<<blocks, threads>>>(out, iters); });
  printf("%-9s %8.3f ms  %7.2f TFLOP/s\n", "DFMA", ms, 2.0 * 16 * iters * (double)blocks * threads / ms * 1e-9);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
