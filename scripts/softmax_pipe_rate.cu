// Which pipe does the fp32 -> 16-bit pack of the softmax sit on?  Per-warp cost of the M-phase instruction mixes on one
// SM sub-partition (sm_100a): MUFU.EX2, F2FP.PACK_AB, PRMT truncation pack, and their mixes, at 1..4 warps per sub-partition.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/bin/softmax_pipe_rate scripts/softmax_pipe_rate.cu
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#define ITER 2048
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1) {
  asm volatile("{\n\t.reg .b64 a, b, c, d;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\tmov.b64 c, {%6, %7};\n\t"
      "fma.rn.f32x2 d, a, b, c;\n\tmov.b64 {%0, %1}, d;\n\t}" : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned pack_h(float a, float b) { unsigned r; asm volatile("cvt.rn.f16x2.f32 %0, %2, %1;" : "=r"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ unsigned pack_b(float a, float b) { unsigned r; asm volatile("cvt.rn.bf16x2.f32 %0, %2, %1;" : "=r"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ unsigned pack_prmt(float a, float b) { unsigned r; asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(r) : "r"(__float_as_uint(a)), "r"(__float_as_uint(b))); return r; }
// fp32 in (2^-24, 1] -> fp16 bits by integer arithmetic (round to nearest up, flush below 2^-14): two values -> one word
__device__ __forceinline__ unsigned pack_int(float a, float b) {
  unsigned ua = __float_as_uint(a), ub = __float_as_uint(b);
  ua = (ua + 0x1000u) >> 13; ub = (ub + 0x1000u) >> 13;         // keep exponent+10 mantissa bits
  int ha = (int)ua - (112 << 10), hb = (int)ub - (112 << 10);   // rebias 127 -> 15
  ha = max(ha, 0); hb = max(hb, 0);
  return (unsigned)ha | ((unsigned)hb << 16);
}
template <int MODE>
__global__ void k(unsigned* out, float s, long long* cyc) {
  float x[16]; unsigned acc = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = -(threadIdx.x * 0.001f + i * 0.1f);
  long long t0 = clock64();
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      float a = x[i], b = x[i + 1];
      if (MODE == 0) acc ^= pack_h(a, b);
      if (MODE == 1) acc ^= pack_b(a, b);
      if (MODE == 2) acc ^= pack_prmt(a, b);
      if (MODE == 3) acc ^= pack_h(ex2(a), ex2(b));
      if (MODE == 4) acc ^= pack_prmt(ex2(a), ex2(b));
      if (MODE == 5) { float e0, e1; ffma2(e0, e1, a, b, s, s, -1.f, -1.f); acc ^= pack_h(ex2(e0), ex2(e1)); }
      if (MODE == 6) { float e0, e1; ffma2(e0, e1, a, b, s, s, -1.f, -1.f); acc ^= pack_int(ex2(e0), ex2(e1)); }
      if (MODE == 7) { float e0 = ex2(a), e1 = ex2(b); acc ^= __float_as_uint(e0) ^ __float_as_uint(e1); }
      if (MODE == 8) { float e0, e1; ffma2(e0, e1, a, b, s, s, -1.f, -1.f); acc ^= pack_b(ex2(e0), ex2(e1)); }
      if (MODE == 9) acc ^= pack_int(a, b);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] += 1e-7f * (float)(acc & 1);  // keep the loop body dependent on acc (not hoistable)
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int M> static void run(int warps, unsigned* out, long long* cyc, const char* name) {
  long long h;
  k<M><<<1, 128 * warps>>>(out, 1.0001f, cyc);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("{\"warps_per_smsp\": %d, \"mix\": \"%s\", \"cycles_per_pair_per_warp\": %.2f, \"cycles_per_pair_per_smsp\": %.2f}\n", warps, name,
         (double)h / (ITER * 8), (double)h / (ITER * 8) / warps);
}
int main() {
  unsigned* out; long long* cyc; cudaMalloc(&out, 4 << 20); cudaMalloc(&cyc, 8);
  for (int warps = 1; warps <= 4; ++warps) {
    run<0>(warps, out, cyc, "F2FP.F16 pack only");
    run<1>(warps, out, cyc, "F2FP.BF16 pack only");
    run<2>(warps, out, cyc, "PRMT pack only");
    run<9>(warps, out, cyc, "integer fp16 pack only");
    run<7>(warps, out, cyc, "2 MUFU.EX2");
    run<3>(warps, out, cyc, "2 MUFU.EX2 + F2FP.F16");
    run<4>(warps, out, cyc, "2 MUFU.EX2 + PRMT");
    run<5>(warps, out, cyc, "FFMA2 + 2 MUFU.EX2 + F2FP.F16");
    run<8>(warps, out, cyc, "FFMA2 + 2 MUFU.EX2 + F2FP.BF16");
    run<6>(warps, out, cyc, "FFMA2 + 2 MUFU.EX2 + integer fp16 pack");
  }
  return 0;
}
