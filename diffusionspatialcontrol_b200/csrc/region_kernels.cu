// Region-weight map builder on the device (companion of the attention path).
//
// Replaces the per-resolution body of encode_region_map_sp
// (reference source/modules/encode_region_map_function.py:49-69):
//   bin = (map < 255);  ds = cv2.resize(bin, (w_r, h_r), INTER_CUBIC);  m = (ds == ds.max());
//   m = m * S;  m[m == 0] = -S';  W[:, idx:idx+n] += m   (fp32 += fp64, one rounding per addition)
//
// The resize is OpenCV's separable 4-tap cubic (a = -0.75): float32 coefficients evaluated in
// interpolateCubic()'s operation order (no FMA contraction), source coordinate (d+0.5)*scale-0.5 in
// fp64 then cast to float32, replicate borders; for a 0/1 image the rounded uint8 result is
// 1 iff the exact weighted sum exceeds 0.5.  The sum is accumulated in fp64 in a fixed order
// (horizontal taps left to right, then vertical taps top to bottom) with explicit round-to-nearest
// mul/add so the result is bit-identical to the numpy oracle (oracle/region_map.py).
#include "dsc_device.cuh"
#include "dsc_internal.h"

namespace dsc {

struct Taps {
  int ofs;     // floor(source coordinate), unclamped
  float c[4];  // cubic coefficients for taps ofs-1 .. ofs+2
};

__device__ __forceinline__ Taps cubic_taps(int d, int src, int dst) {
  // OpenCV: inv_scale = (double)dst/src; scale = 1./inv_scale; fx = (float)((d+0.5)*scale - 0.5)
  const double inv_scale = __ddiv_rn(static_cast<double>(dst), static_cast<double>(src));
  const double scale = __ddiv_rn(1.0, inv_scale);
  const float f = __double2float_rn(__dsub_rn(__dmul_rn(static_cast<double>(d) + 0.5, scale), 0.5));
  const float fl = floorf(f);
  const float x = __fsub_rn(f, fl);
  const float A = -0.75f;
  Taps t;
  t.ofs = static_cast<int>(fl);
  const float x1 = __fadd_rn(x, 1.f);
  // ((A*(x+1) - 5A)*(x+1) + 8A)*(x+1) - 4A
  float v = __fsub_rn(__fmul_rn(A, x1), __fmul_rn(5.f, A));
  v = __fadd_rn(__fmul_rn(v, x1), __fmul_rn(8.f, A));
  t.c[0] = __fsub_rn(__fmul_rn(v, x1), __fmul_rn(4.f, A));
  // ((A+2)*x - (A+3))*x*x + 1
  v = __fsub_rn(__fmul_rn(__fadd_rn(A, 2.f), x), __fadd_rn(A, 3.f));
  t.c[1] = __fadd_rn(__fmul_rn(__fmul_rn(v, x), x), 1.f);
  const float xm = __fsub_rn(1.f, x);
  v = __fsub_rn(__fmul_rn(__fadd_rn(A, 2.f), xm), __fadd_rn(A, 3.f));
  t.c[2] = __fadd_rn(__fmul_rn(__fmul_rn(v, xm), xm), 1.f);
  t.c[3] = __fsub_rn(__fsub_rn(__fsub_rn(1.f, t.c[0]), t.c[1]), t.c[2]);
  return t;
}

__global__ void region_downsample_kernel(const uint8_t* __restrict__ maps, int R, int Hpx, int Wpx, int w_r, int h_r,
                                         uint8_t* __restrict__ ds, uint32_t* __restrict__ any_set) {
  const long long n_out = static_cast<long long>(R) * h_r * w_r;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_out) return;
  const int ox = static_cast<int>(i % w_r);
  const int oy = static_cast<int>((i / w_r) % h_r);
  const int r = static_cast<int>(i / (static_cast<long long>(w_r) * h_r));
  const uint8_t* src = maps + static_cast<long long>(r) * Hpx * Wpx;
  uint8_t bit;
  if (w_r == Wpx && h_r == Hpx) {
    bit = src[static_cast<long long>(oy) * Wpx + ox] < 255 ? 1 : 0;  // cv2 copies when the size is unchanged
  } else {
    const Taps tx = cubic_taps(ox, Wpx, w_r);
    const Taps ty = cubic_taps(oy, Hpx, h_r);
    double out = 0.0;
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
      const int yi = min(max(ty.ofs + ky - 1, 0), Hpx - 1);
      const uint8_t* row = src + static_cast<long long>(yi) * Wpx;
      double acc = 0.0;
#pragma unroll
      for (int kx = 0; kx < 4; ++kx) {
        const int xi = min(max(tx.ofs + kx - 1, 0), Wpx - 1);
        const double s = row[xi] < 255 ? 1.0 : 0.0;
        acc = __dadd_rn(acc, __dmul_rn(s, static_cast<double>(tx.c[kx])));
      }
      out = __dadd_rn(out, __dmul_rn(acc, static_cast<double>(ty.c[ky])));
    }
    bit = out > 0.5 ? 1 : 0;
  }
  ds[i] = bit;
  if (bit) atomicOr(any_set + r, 1u);
}

__global__ void region_accumulate_kernel(const uint8_t* __restrict__ ds, const uint32_t* __restrict__ any_set, int L_r,
                                         const double* __restrict__ weight, const double* __restrict__ mask_outsides,
                                         const int32_t* __restrict__ span_region,
                                         const int32_t* __restrict__ span_start, const int32_t* __restrict__ span_len,
                                         int n_spans, int n_tok, float* __restrict__ W_out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(L_r) * n_tok) return;
  const int tok = static_cast<int>(i % n_tok);
  const int l = static_cast<int>(i / n_tok);
  float w = 0.f;
  for (int s = 0; s < n_spans; ++s) {
    const int st = span_start[s];
    if (tok < st || tok >= st + span_len[s]) continue;
    const int r = span_region[s];
    // m = (ds == ds.max()): a region that vanished at this resolution (max 0) selects every pixel
    const bool m = ds[static_cast<long long>(r) * L_r + l] == (any_set[r] ? 1 : 0);
    const double prod = m ? weight[r] : __dmul_rn(0.0, weight[r]);     // m * S
    const double val = (prod == 0.0) ? -mask_outsides[r] : prod;       // m[m == 0] = -S'
    w = __double2float_rn(__dadd_rn(static_cast<double>(w), val));     // fp32 += fp64
  }
  W_out[i] = w;
}

cudaError_t run_region_downsample(const uint8_t* maps, int R, int Hpx, int Wpx, int w_r, int h_r, uint8_t* ds,
                                  uint32_t* any_set, cudaStream_t st) {
  const long long n = static_cast<long long>(R) * h_r * w_r;
  if (n == 0) return cudaSuccess;
  region_downsample_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(maps, R, Hpx, Wpx, w_r, h_r, ds,
                                                                                  any_set);
  return cudaGetLastError();
}

cudaError_t run_region_accumulate(const uint8_t* ds, const uint32_t* any_set, int R, int L_r, const double* weight,
                                  const double* mask_outsides, const int32_t* span_region, const int32_t* span_start,
                                  const int32_t* span_len, int n_spans, int n_tok, float* W_out, cudaStream_t st) {
  (void)R;
  const long long n = static_cast<long long>(L_r) * n_tok;
  if (n == 0) return cudaSuccess;
  region_accumulate_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(
      ds, any_set, L_r, weight, mask_outsides, span_region, span_start, span_len, n_spans, n_tok, W_out);
  return cudaGetLastError();
}

}  // namespace dsc
