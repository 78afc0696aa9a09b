// Measures the dense fp16 mma.sync (HMMA m16n8k16) rate of the legacy tensor path on this GPU, to
// bound what the attention kernels can expect from it.  nvcc -arch=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* out, int iters) {
  float d[8][4] = {};
  unsigned a[4] = {0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u}, b0 = 0x3c003c00u + threadIdx.x, b1 = 0x3c003c00u;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(d[j][0]), "+f"(d[j][1]), "+f"(d[j][2]), "+f"(d[j][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  float s = 0;
  for (int j = 0; j < 8; ++j) s += d[j][0] + d[j][1] + d[j][2] + d[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out;
  cudaMalloc(&out, 148 * 8 * 1024 * 4);
  for (int warps = 4; warps <= 32; warps *= 2) {
    int iters = 20000, blocks = 148 * (32 / warps > 2 ? 2 : 1);
    k<<<blocks, warps * 32>>>(out, 100);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<<<blocks, warps * 32>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 16 * 8 * 16 * 8.0 * iters * warps * blocks;
    printf("{\"bench\":\"mma_sync_m16n8k16_f16\",\"blocks\":%d,\"warps_per_block\":%d,\"ms\":%.3f,\"tflops\":%.1f}\n", blocks,
           warps, ms, flops / ms / 1e9);
  }
  return 0;
}
