// micro-benchmark (development aid): FP64 pipe latency / issue behaviour on the box
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP> __global__ void chain(double* out, int iters, long long* cyc) {
  double a[ILP];
  const double x = 1.0 + 1e-9 * threadIdx.x, y = 1e-9;
#pragma unroll
  for (int k = 0; k < ILP; ++k) a[k] = k * 0.5 + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < ILP; ++k) a[k] = fma(a[k], x, y);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += a[k];
  if (s == 1234.5) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void lds_chain(double* out, int iters, long long* cyc) {
  __shared__ int idx[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) idx[i] = (i + 33) & 1023;
  __syncthreads();
  int j = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) j = idx[j];
  long long t1 = clock64();
  if (j == -1) out[0] = j;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void sincos_chain(double* out, int iters, long long* cyc) {
  double a = 0.3 + threadIdx.x * 0.01, s, c;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) { sincos(a, &s, &c); a = s + c; }
  long long t1 = clock64();
  if (a == 1234.5) out[0] = a;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void div_chain(double* out, int iters, long long* cyc) {
  double a = 1.3 + threadIdx.x * 0.01;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) { a = 1.0 / a + 0.5; }
  long long t1 = clock64();
  if (a == 1234.5) out[0] = a;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int ILP> void run(const char* name, int warps_per_block, int blocks) {
  double* d; long long* c; cudaMalloc(&d, 8); cudaMalloc(&c, 8);
  const int iters = 4096;
  chain<ILP><<<blocks, 32 * warps_per_block>>>(d, iters, c);
  chain<ILP><<<blocks, 32 * warps_per_block>>>(d, iters, c);
  long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("%s ILP=%d warps/block=%d blocks=%d: %.2f cycles per DFMA per warp (%.2f per iteration)\n", name, ILP, warps_per_block, blocks, (double)h / iters / ILP, (double)h / iters);
}
int main() {
  run<1>("dfma", 1, 1); run<2>("dfma", 1, 1); run<4>("dfma", 1, 1); run<8>("dfma", 1, 1); run<16>("dfma", 1, 1);
  run<1>("dfma", 4, 1); run<4>("dfma", 4, 1); run<8>("dfma", 4, 1);
  run<1>("dfma", 8, 1); run<4>("dfma", 8, 1); run<1>("dfma", 16, 1); run<2>("dfma", 16, 1); run<1>("dfma", 32, 1);
  double* d; long long* c; cudaMalloc(&d, 8); cudaMalloc(&c, 8); long long h;
  lds_chain<<<1, 32>>>(d, 4096, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("lds dependent: %.1f cycles\n", h / 4096.0);
  sincos_chain<<<1, 32>>>(d, 4096, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("sincos+add dependent: %.1f cycles\n", h / 4096.0);
  div_chain<<<1, 32>>>(d, 4096, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("1/x+add dependent: %.1f cycles\n", h / 4096.0);
  return 0;
}
