// FP64 / FP32 FMA throughput and dependent-chain latency on the device (nvcc -arch=sm_100a -O3 fp64_rate.cu -o fp64_rate)
#include <cstdio>
#include <cuda_runtime.h>
template <class T>
__global__ void k_thr(T* out, T a, T b, int n) {
  T x[8];
  for (int k = 0; k < 8; k++) x[k] = a + (T)(threadIdx.x + k);
  for (int i = 0; i < n; i++)
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = x[k] * b + a;
  T s = 0;
  for (int k = 0; k < 8; k++) s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class T>
__global__ void k_lat(T* out, T a, T b, int n, long long* cyc) {
  T x = a;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) x = x * b + a;
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <class T>
void run(const char* name) {
  T* d; cudaMalloc(&d, 148 * 8 * 256 * sizeof(T));
  long long* dc; cudaMalloc(&dc, 8);
  const int n = 4096;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_thr<T><<<148 * 8, 256>>>(d, (T)1.0001, (T)0.9999, n);
  cudaEventRecord(e0);
  k_thr<T><<<148 * 8, 256>>>(d, (T)1.0001, (T)0.9999, n);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double flops = 2.0 * 148 * 8 * 256 * 8.0 * n;
  k_lat<T><<<1, 32>>>(d, (T)1.0001, (T)0.9999, 4096, dc);
  long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
  printf("%s: %.2f TFLOP/s (FMA), dependent FMA latency %.1f cycles\n", name, flops / ms / 1e9, (double)c / 4096);
}
int main() { run<float>("fp32"); run<double>("fp64"); return 0; }
