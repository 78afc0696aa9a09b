// Internal (non-ABI) declarations shared by the kernels and the C-ABI translation unit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dsc_b200.h"

namespace dsc {

// Head of the workspace == dsc_xattn_stats_t (include/dsc_b200.h); per-CTA fp64 partials follow.
using Workspace = dsc_xattn_stats_t;
constexpr int kWorkspaceHeader = 64;
constexpr int kMaxPartials = 1024;
// Behind the per-CTA partials: the HANDOFF slots of the prepared-K/V call (xattn_x3.cu), kMaxPartials x {sum, sum of squares} as
// fp64 bit patterns with bit 0 forced to 1 (0 = "not published yet"; the buffer starts zeroed and the call leaves it zeroed),
// and in the header, behind the public fields, the count of pass-2 CTAs that have consumed them.
constexpr int kHandoffOffset = kWorkspaceHeader + 16 * kMaxPartials;
constexpr int kReadersOffset = 56;
static_assert(sizeof(Workspace) <= kWorkspaceHeader, "workspace header");

// Division by a launch-invariant divisor without the ~100-cycle integer-division sequence (role prologues, tile decode):
// q = x / d for every 32-bit x.  d >= 1.  (round-up magic number, 33-bit form: umulhi + shift-add)
struct FastDiv {
  uint32_t m, l, d;
};
inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f{0u, 0u, d};
  if (d <= 1u) return f;
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;  // ceil(log2 d)
  f.l = l;
  f.m = static_cast<uint32_t>(((1ull << 32) * ((1ull << l) - d)) / d + 1ull);
  return f;
}
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t fastdiv(uint32_t x, const FastDiv& f) {
  if (f.d <= 1u) return x;
  const uint32_t q = __umulhi(x, f.m);
  return (((x - q) >> 1) + q) >> (f.l - 1u);
}
#endif

struct XattnParams {
  const void* q;
  const void* k;
  const void* v;
  void* out;
  const float* W;
  const float* sigma_dev;
  float sigma_host;
  float scale;
  long long q_sb, q_sl;  // element strides: batch, row (heads are contiguous D-wide column groups)
  long long k_sb, k_ss;
  long long v_sb, v_ss;
  long long o_sb, o_sl;
  int B, H, L, S, Bw;
  int n_hg;         // head groups per batch row
  int n_sl;         // 16-row slices per (batch, head-group)
  long long total;  // B * n_hg * n_sl
  Workspace* ws;
  int w_pitch;          // floats between consecutive query rows of W
  // long prompts (more than DSC_MAX_KEYS keys) run as chunks of <= 80 keys: pass 1 accumulates its partials per
  // chunk and folds them all at the last one; pass 2 writes per-chunk outputs + log-sum-exp, merged afterwards
  int chunk;            // index of this key chunk (partials slot = chunk * gridDim.x + blockIdx.x)
  int fold_chunks;      // pass 1: number of chunk slots to fold into the std at the end of this launch (0 = none)
  double n_total;       // pass 1: number of scores of the WHOLE call (all chunks)
  int w_col0;           // pass 2: first W column of this chunk
  float* lse;           // pass 2: optional [B, H, L] log2-sum-exp of each row's logits (nullptr = not wanted)
  int fused_cps;        // fused single-launch kernel: CTAs per (batch, head-group) segment
  // compact region map (optional; tcgen05 pass 2): only the n_active <= 16 key columns that are non-zero anywhere,
  // fp32 [Bw, L, 20] (16 values + 4 pad floats per row); the kernel permutes the keys so that these columns come first
  const float* wc;
  int n_active;
  int active_cols[16];  // ascending key indices of the compact columns
  unsigned flags;       // 4-warpgroup tcgen05 kernel: launcher-set mode bits (see launch_tc5x4)
  const void* kv_image;  // 3-warpgroup tcgen05 kernels: prepared K / V^T images (dsc_xattn_prepare_kv), one per (batch, head group)
  // 3-warpgroup tcgen05 kernels: tile ranges and tile decode without divisions (set by the launcher): CTA b owns tiles
  // [b * tiles_q + min(b, tiles_r), +tiles_q + (b < tiles_r))
  uint32_t tiles_q, tiles_r;
  FastDiv div_nsl, div_nhg;
  int handoff;           // 3-warpgroup tcgen05 kernels, both passes of one call: the std goes from pass 1 to pass 2 through the handoff slots  // additive attention mask M (optional; mma.sync kernels, dsc_xattn_call_masked): fp32, element (b, h, l, s) at
  // mask[b * m_sb + h * m_sh + l * m_sl + m_col0 + s]; a zero stride broadcasts that dimension.  a = qk_scale * Q K^T + M:
  // pass 1 then sums a itself (and finalises with scale = 1), pass 2 adds M next to beta * W
  const float* mask;
  long long m_sb, m_sh, m_sl;
  int m_col0;      // first mask column of this key chunk (long prompts)
  float qk_scale;  // masked pass 1: the score scale (p.scale is 1 there: the partial sums are already scaled)
};

// Kernel-selection overrides (A/B runs, tests).  Filled ONCE from the environment when the library is loaded
// (DSC_XATTN_IMPL, DSC_XATTN_STATS_IMPL, DSC_NO_FUSED, DSC_TC5_FUSED, DSC_NO_PDL, DSC_TC5_FLAGS, DSC_TC5_VARIANT);
// afterwards only dsc_config_set changes it.  No attention call touches the environment.
enum Impl { kImplAuto = 0, kImplMma = 1, kImplTc5 = 2, kImplGram = 3 };
struct Config {
  int impl = kImplAuto;        // both passes
  int stats_impl = kImplAuto;  // pass 1 alone (kImplAuto: follow impl)
  bool no_fused = false;       // never take the single-launch forms
  bool tc5_fused = false;      // single-launch two-phase tcgen05 form (opt-in)
  bool no_pdl = false;         // no programmatic dependent launch
  bool tc5_x2 = false;         // 2-warpgroup tcgen05 variant instead of x4
  unsigned tc5_flags = 0;      // launch_tc5x4 mode bits
};
const Config& config();

int sm_count_cached();
int heads_per_group(int D);  // 0 if D is unsupported
int stats_grid(long long total);
cudaError_t run_stats(const XattnParams& p, int D, int dtype, cudaStream_t st);
cudaError_t run_forward(const XattnParams& p, int D, int dtype, cudaStream_t st);
// both passes in one cooperative launch when every CTA's share of Q fits in shared memory (small layers / batches)
bool fused_plan(int B, int H, int L, int D, int S, int* cps_out);
cudaError_t run_fused(const XattnParams& p, int D, int dtype, cudaStream_t st);
// pass 1 through the Gram identity (no scores formed); D = 40, S <= 80
bool gram_supports(int D, int S);
cudaError_t run_stats_gram(const XattnParams& p, int D, int dtype, cudaStream_t st);
// out[b, l, h*D + d] = sum_c w_c * chunk_out[c][b, l, h*D + d], w_c = 2^(lse[c][b,h,l] - max) / sum (chunk outputs dense)
cudaError_t run_merge_chunks(const void* chunk_out, const float* lse, int n_chunks, void* out, long long o_sb, long long o_sl,
                             int B, int H, int L, int D, int dtype, cudaStream_t st);

// tcgen05 / TMEM implementation of the same two passes (xattn_tc5.cu), D in {40, 80}
bool tc5_supports(int D);
cudaError_t run_stats_tc5(const XattnParams& p, int D, int dtype, cudaStream_t st);
cudaError_t run_forward_tc5(const XattnParams& p, int D, int dtype, cudaStream_t st);
bool tc5_fused_supports(int D);  // both passes in one cooperative launch
cudaError_t run_fused_tc5(const XattnParams& p, int D, int dtype, cudaStream_t st);

// decoupled-warpgroup tcgen05 kernels over a prepared K / V^T image (xattn_x3.cu): D in {40, 80, 160}, S = 77, whole
// 160-column head groups (H % 4 / H % 2 / any H), compact region map
bool x3_supports(int H, int D, int S);
size_t x3_image_bytes(int B, int H, int D);
cudaError_t run_prepare_kv_x3(const void* k, const void* v, long long k_sb, long long k_ss, long long v_sb, long long v_ss, int B,
                              int H, int D, int S, int n_active, const int* cols, int dtype, void* image, cudaStream_t st);
cudaError_t run_stats_x3(const XattnParams& p, int D, int dtype, cudaStream_t st);
cudaError_t run_forward_x3(const XattnParams& p, int D, int dtype, cudaStream_t st);
cudaError_t run_fused_x3(const XattnParams& p, int D, int dtype, cudaStream_t st);  // both passes, one cooperative launch

cudaError_t run_region_downsample(const uint8_t* maps, int R, int Hpx, int Wpx, int w_r, int h_r, uint8_t* ds,
                                  uint32_t* any_set, cudaStream_t st);
cudaError_t run_region_accumulate(const uint8_t* ds, const uint32_t* any_set, int R, int L_r, const double* weight,
                                  const double* mask_outsides, const int32_t* span_region, const int32_t* span_start,
                                  const int32_t* span_len, int n_spans, int n_tok, float* W_out, cudaStream_t st);

struct StepCoef {
  float c_x;       // sigma_next / sigma
  float c_d;       // -expm1(-h)
  float c_den;     // 1 + 1/(2r)   (1 for the first-order update)
  float c_prev;    // -1/(2r)      (0 for the first-order update)
  float sigma;     // den = x - sigma * eps
  float cfg;
  float c_in_next;  // 1/sqrt(sigma_next^2 + 1)
};
cudaError_t run_dpmpp2m_step(float* x, const void* eps_uc, float* den_prev, void* unet_in_next, long long n_elem,
                             const StepCoef& c, int dtype, cudaStream_t st);

}  // namespace dsc
