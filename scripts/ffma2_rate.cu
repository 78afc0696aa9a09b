// Issue/throughput of packed fp32 FFMA2 vs scalar FFMA, MUFU.EX2 and FMNMX3 on one SM sub-partition mix (sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/bin/ffma2_rate scripts/ffma2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 4096
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1) {
  asm volatile("{\n\t.reg .b64 a, b, c, d;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\tmov.b64 c, {%6, %7};\n\t"
      "fma.rn.f32x2 d, a, b, c;\n\tmov.b64 {%0, %1}, d;\n\t}" : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}
template <int MODE>
__global__ void k(float* out, float s, long long* cyc) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 0.001f + i;
  long long t0 = clock64();
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      if (MODE == 0) { x[i] = fmaf(x[i], s, 0.5f); x[i + 1] = fmaf(x[i + 1], s, 0.5f); }
      if (MODE == 1) ffma2(x[i], x[i + 1], x[i], x[i + 1], s, s, 0.5f, 0.5f);
      if (MODE == 2) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i])); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i + 1])); }
      if (MODE == 3) { x[i] = fmaxf(fmaxf(x[i], x[i + 1]), s); x[i + 1] = fmaxf(fmaxf(x[i + 1], x[i]), -s); }
      if (MODE == 4) {  // packed half exp2: one instruction for the pair
        unsigned int h = __float_as_uint(x[i]);
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h));
        x[i] = __uint_as_float(h);
      }
      if (MODE == 5) {  // bf16x2
        unsigned int h = __float_as_uint(x[i]);
        asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h));
        x[i] = __uint_as_float(h);
      }
    }
  }
  long long t1 = clock64();
  float acc = 0; for (int i = 0; i < 16; ++i) acc += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 4 << 20); cudaMalloc(&cyc, 8);
  const char* names[6] = {"FFMA (2 per pair)", "FFMA2 (1 per pair)", "MUFU.EX2 (2 per pair)", "FMNMX3-ish (2 per pair)", "ex2.f16x2 (1 per pair)", "ex2.bf16x2 (1 per pair)"};
  for (int warps = 1; warps <= 8; warps *= 2) {  // warps per sub-partition (block = 4 SMSPs x warps)
    for (int m = 0; m < 6; ++m) {
      long long h;
      if (m == 0) k<0><<<1, 128 * warps>>>(out, 1.0001f, cyc); if (m == 1) k<1><<<1, 128 * warps>>>(out, 1.0001f, cyc);
      if (m == 2) k<2><<<1, 128 * warps>>>(out, 1.0001f, cyc); if (m == 3) k<3><<<1, 128 * warps>>>(out, 1.0001f, cyc);
      if (m == 4) k<4><<<1, 128 * warps>>>(out, 1.0001f, cyc); if (m == 5) k<5><<<1, 128 * warps>>>(out, 1.0001f, cyc);
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      printf("{\"warps_per_smsp\": %d, \"op\": \"%s\", \"cycles_per_pair_per_warp\": %.3f, \"cycles_per_pair_per_smsp\": %.3f}\n", warps, names[m],
             (double)h / (ITER * 8), (double)h / (ITER * 8) / warps);
    }
  }
  return 0;
}
