// micro-benchmark: fp64 / fp32 FMA issue rate per SM on this GPU (development aid)
#include <cstdio>
#include <cuda_runtime.h>
template <typename T, int ILP>
__global__ void fma_kernel(T* out, int iters) {
  T a[ILP], b = (T)1.000001, c = (T)0.5;
#pragma unroll
  for (int i = 0; i < ILP; ++i) a[i] = (T)(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = a[i] * b + c;
  }
  T s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename T, int ILP>
void run(const char* name, int blocks_per_sm, int threads) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  T* out; cudaMalloc(&out, sizeof(T) * sms * blocks_per_sm * threads);
  const int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  fma_kernel<T, ILP><<<sms * blocks_per_sm, threads>>>(out, 100);
  cudaEventRecord(e0);
  fma_kernel<T, ILP><<<sms * blocks_per_sm, threads>>>(out, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double fmas = (double)sms * blocks_per_sm * threads * iters * ILP;
  printf("%s ILP %d, %d x %d threads/SM: %.2f TFLOP/s, %.1f FMA lanes/clk/SM (at %.0f MHz nominal)\n", name, ILP,
         blocks_per_sm, threads, 2 * fmas / ms / 1e9, fmas / (ms * 1e-3) / sms / (clk * 1e3), clk / 1e3);
  cudaFree(out);
}
int main() {
  run<double, 1>("fp64", 1, 128); run<double, 4>("fp64", 1, 128); run<double, 4>("fp64", 4, 128);
  run<double, 4>("fp64", 4, 256); run<double, 8>("fp64", 2, 1024);
  run<float, 4>("fp32", 4, 256); run<float, 8>("fp32", 2, 1024);
  return 0;
}
