// FP64 tensor-core MMA issue rates on sm_100a: m8n8k4 vs m16n8k4 / k8 / k16 (and DFMA for scale).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_rate dmma_rate.cu && ./dmma_rate
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
template <int SHAPE, int NACC>
__global__ void k(double* out, double a0, double b0) {
  double acc[NACC][4];
  for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  double a[8], b[4];
  for (int i = 0; i < 8; ++i) a[i] = a0 + threadIdx.x * 1e-3 + i;
  for (int i = 0; i < 4; ++i) b[i] = b0 + threadIdx.x * 1e-4 + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      if (SHAPE == 0) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(acc[i][0]), "+d"(acc[i][1]) : "d"(a[0]), "d"(b[0]));
      } else if (SHAPE == 1) {
        asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+d"(acc[i][0]), "+d"(acc[i][1]), "+d"(acc[i][2]), "+d"(acc[i][3])
                     : "d"(a[0]), "d"(a[1]), "d"(b[0]));
      } else if (SHAPE == 2) {
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+d"(acc[i][0]), "+d"(acc[i][1]), "+d"(acc[i][2]), "+d"(acc[i][3])
                     : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
      } else if (SHAPE == 3) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                     : "+d"(acc[i][0]), "+d"(acc[i][1]), "+d"(acc[i][2]), "+d"(acc[i][3])
                     : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                       "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
      } else {
        asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(acc[i][0]) : "d"(a[0]), "d"(b[0]));
        asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(acc[i][1]) : "d"(a[1]), "d"(b[1]));
      }
    }
  }
  double s = 0;
  for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int SHAPE, int NACC>
void run(const char* name, double fma_per_inst, int warps_per_sm) {
  double* out;
  cudaMalloc(&out, 148 * 1024 * 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int threads = warps_per_sm * 32;
  k<SHAPE, NACC><<<148, threads>>>(out, 1.0, 2.0);
  cudaEventRecord(e0);
  k<SHAPE, NACC><<<148, threads>>>(out, 1.0, 2.0);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t err = cudaGetLastError();
  const double inst = 148.0 * warps_per_sm * ITERS * NACC;
  const double tf = 2.0 * inst * fma_per_inst / (ms * 1e-3) / 1e12;
  const double cyc = ms * 1e-3 * 1.965e9 / (ITERS * NACC * (warps_per_sm / 4.0));
  printf("%-10s acc=%d warps/SM=%2d  %.3f ms  %.2f TFLOP/s  %.1f cycles per warp-instruction per sub-partition  (%s)\n",
         name, NACC, warps_per_sm, ms, tf, cyc, cudaGetErrorString(err));
  cudaFree(out);
}

int main() {
  run<0, 4>("m8n8k4", 256, 16);
  run<0, 8>("m8n8k4", 256, 16);
  run<0, 4>("m8n8k4", 256, 4);
  run<1, 4>("m16n8k4", 512, 16);
  run<2, 4>("m16n8k8", 1024, 16);
  run<3, 4>("m16n8k16", 2048, 16);
  run<3, 2>("m16n8k16", 2048, 8);
  run<4, 8>("dfma x2", 64, 16);
  return 0;
}
