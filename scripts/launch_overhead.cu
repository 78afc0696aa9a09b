// How much of a short kernel's CUDA-event time is launch / shared-memory carve-out reconfiguration?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void empty_k(int* p) { extern __shared__ int s[]; if (p && threadIdx.x == 9999) p[0] = s[0]; }
__global__ void spin_k(long long cycles) { long long t0 = clock64(); while (clock64() - t0 < cycles) {} }
int main() {
  const int big = 225 * 1024, mid = 154 * 1024;
  cudaFuncSetAttribute(empty_k, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  char* flush; cudaMalloc(&flush, 512u << 20);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  auto timeit = [&](const char* name, int smem_prev, int smem, int threads, bool do_flush) {
    float best = 1e9, sum = 0; int n = 20;
    for (int i = 0; i < n; ++i) {
      if (do_flush) cudaMemsetAsync(flush, 0, 512u << 20);
      if (smem_prev >= 0) empty_k<<<148, 256, smem_prev>>>(nullptr);
      cudaEventRecord(a);
      empty_k<<<148, threads, smem>>>(nullptr);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b); sum += ms; if (ms < best) best = ms;
    }
    printf("{\"case\":\"%s\",\"avg_us\":%.2f,\"min_us\":%.2f}\n", name, sum / n * 1e3, best * 1e3);
  };
  timeit("empty 225KB/640thr after memset flush", -1, big, 640, true);
  timeit("empty 225KB/640thr after 0-smem kernel", 0, big, 640, false);
  timeit("empty 225KB/640thr after same-carveout kernel", big, big, 640, false);
  timeit("empty 154KB/640thr after 225KB kernel", big, mid, 640, false);
  timeit("empty 0KB/256thr after 0KB kernel", 0, 0, 256, false);
  timeit("empty 0KB/256thr after memset flush", -1, 0, 256, true);
  // a 20 us spin kernel with the big config after flush: event time minus 20 us = overhead
  cudaFuncSetAttribute(spin_k, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  for (int rep = 0; rep < 2; ++rep) {
    float sum = 0; int n = 20;
    for (int i = 0; i < n; ++i) {
      cudaMemsetAsync(flush, 0, 512u << 20);
      cudaEventRecord(a);
      spin_k<<<148, 640, rep ? big : 0>>>(38000);  // ~20 us at 1.9 GHz
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b); sum += ms;
    }
    printf("{\"case\":\"spin 38000 cycles, smem %s, after flush\",\"avg_us\":%.2f}\n", rep ? "225KB" : "0", sum / n * 1e3);
  }
  return 0;
}
