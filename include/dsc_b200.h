/* libdsc_b200.so -- C ABI of the B200-native region-masked cross-attention path.
 *
 * Drop-in boundary for the hot path of duongve13112002/DiffusionSpatialControl.  Every entry point
 * takes plain pointers and sizes (device pointers unless marked HOST), enqueues its work on the
 * given CUDA stream and returns immediately: 0 = ok, <0 = DSC_ERR_* (invalid argument /
 * unsupported configuration, nothing was launched), >0 = cudaError_t from the launch.  No entry
 * point synchronises the host or throws; dsc_last_error is thread-local and the only process-wide
 * state is the kernel-selection table of dsc_config_set (read-only for every other entry point).  `stream` is a cudaStream_t passed as void*.
 *
 * Reference interfaces replaced (paths relative to the reference repository root):
 *   dsc_xattn_stats + dsc_xattn_forward
 *        source/modules/attention_modify.py:74-103  scaled_dot_product_attention_regionstate
 *        source/app.py:1004                         weight_func = w * sigma * qk.std()
 *        (called from AttnProcessor2_0.__call__, source/modules/attention_modify.py:479-481)
 *   dsc_region_downsample + dsc_region_accumulate
 *        source/modules/encode_region_map_function.py:49-53 (resize / ==max / *S / -S')
 *        source/modules/encode_region_map_function.py:57-69 (+= into the token columns)
 *   dsc_dpmpp2m_step
 *        k_diffusion.sampling.sample_dpmpp_2m as driven by source/modules/model_k_diffusion.py:1091-1175
 *        with the denoiser scalings of source/modules/external_k_diffusion.py:95-98,109-114 and the
 *        CFG combine of model_k_diffusion.py:1162-1166
 */
#ifndef DSC_B200_H_
#define DSC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSC_VERSION 108 /* major*10000 + minor*100 + patch */

#define DSC_DTYPE_F16 0
#define DSC_DTYPE_BF16 1

#define DSC_OK 0
#define DSC_ERR_INVALID_ARGUMENT (-1) /* null pointer, non-positive size, bad dtype code */
#define DSC_ERR_UNSUPPORTED (-2)      /* valid request outside what the kernels implement */
#define DSC_ERR_LAYOUT (-3)           /* stride / alignment contract violated */
#define DSC_ERR_SHAPE (-4)            /* shape contract violated (e.g. B % Bw != 0), mirrors the reference raising */

#define DSC_MAX_KEYS 80 /* keys (text tokens) per kernel pass: the tuned case; SD-1.5 uses 77 */
#define DSC_MAX_KEYS_TOTAL 480 /* long prompts (77*k tokens, reference prompt_parser.py:161-194) run as chunks of 80 keys */

int dsc_version(void);

/* Message for the last non-zero return on the calling thread ("" if none). */
const char* dsc_last_error(void);

/* Kernel-selection overrides for A/B runs and tests; the defaults are the measured-fastest choice per shape.
 * The environment (DSC_XATTN_IMPL, DSC_XATTN_STATS_IMPL, DSC_NO_FUSED, DSC_TC5_FUSED, DSC_NO_PDL, DSC_TC5_VARIANT,
 * DSC_TC5_FLAGS) is read ONCE, when the library is first used; no attention call ever calls getenv.  Afterwards only
 * this function changes the table:  key in {"xattn_impl" (mma|tc5), "stats_impl" (mma|tc5|gram), "no_fused" (1),
 * "tc5_fused" (1), "no_pdl" (1), "tc5_variant" (x2), "tc5_flags" (int)}; value NULL or "" restores the default.
 * Process-wide: not to be called concurrently with attention calls. */
int dsc_config_set(const char* key /*HOST*/, const char* value_or_null /*HOST*/);

/* Number of SMs the persistent kernels will be sized for on the current device (148 on B200). */
int dsc_sm_count(void);

/* Bytes of device workspace one attention call needs (stats + per-CTA partials + ticket; for
 * S > DSC_MAX_KEYS also the per-key-chunk outputs and log-sum-exps that dsc_xattn_forward merges).
 * The first 64 + 2 * 16384 bytes (header | per-CTA partials | handoff slots of dsc_xattn_call_prepared) must be zero-filled
 * ONCE after allocation; calls leave it reusable.  A workspace serves one call at a time (one stream). */
int dsc_xattn_workspace_bytes(int B, int H, int L, int D, int S, size_t* out /*HOST*/);

/* Pass 1.  a = scale * Q K^T over the whole call; writes to `workspace`
 *   {float std (unbiased, N-1), float mean, double sum, double sumsq, double n}
 * at the offsets of dsc_xattn_stats_t below.  Deterministic (fixed-order fp64 combine).
 *
 * q: [B,H,L,D] addressed through q_str[4] (ELEMENT strides of B,H,L,D); k: [B,H,S,D] through k_str.
 * Layout contract (DSC_ERR_LAYOUT otherwise): stride(D)==1, stride(H)==D, stride(L) and stride(B)
 * multiples of 8 elements, base pointers 16-byte aligned -- i.e. the [B, L, H*D] projection output
 * viewed as heads, which is exactly what the reference processor produces
 * (attention_modify.py:471-474).  D in {40,64,80,128,160}; 1 <= S <= DSC_MAX_KEYS_TOTAL.
 * mask_or_null: additive attention mask M as a dense fp32 [B, H, L, S] device tensor, or NULL (every SD-1.5 call).  With a
 * mask the statistics are those of a = scale * Q K^T + M (the tensor the reference's weight_func sees: attention_modify.py:90-95,
 * baddbmm variant :39-70 with beta = 1); pad keys and rows beyond L carry no mask value.  Masked calls run on the mma.sync
 * kernels (xattn_kernels.cu).  See dsc_xattn_call_masked for strided / broadcast masks and for what the reference's two
 * processors do with a mask. */
int dsc_xattn_stats(const void* q, const void* k, const int64_t q_str[4] /*HOST*/, const int64_t k_str[4] /*HOST*/,
                    const void* mask_or_null, int B, int H, int L, int D, int S, float scale, int dtype,
                    void* workspace, void* stream);

/* Pass 2.  out = softmax(scale * Q K^T + beta * W) V with beta = sigma * std, std read from
 * `workspace` (written by dsc_xattn_stats earlier on the same stream).
 *
 * W: fp32 [Bw, L, S] region-weight map whose query rows are w_pitch floats apart (S <= w_pitch <=
 *    DSC_MAX_KEYS; w_pitch == S is the reference's dense tensor; batch stride = L * w_pitch; for
 *    S > DSC_MAX_KEYS any w_pitch >= S that is a multiple of 4, with a 16-byte aligned base); row b
 *    of the batch uses W[b / (B / Bw)] (the batch-major repeat_interleave of
 *    attention_modify.py:96-99); requires B % Bw == 0.  The padded form w_pitch == DSC_MAX_KEYS with
 *    a 16-byte aligned base is the fast path (rows fetched by TMA boxes and read as 128-bit words);
 *    the pad columns are never read as weights.
 * sigma: if sigma_dev_or_null != NULL it points to ONE fp32 on the device (no host sync),
 *    else sigma_host is used.
 * out: [B, L, H*D] addressed through o_str[3] (element strides of B, L, and 1 for the last dim).
 * v: [B,H,S,D] through v_str, same layout contract as k. */
int dsc_xattn_forward(const void* q, const void* k, const void* v, const int64_t q_str[4] /*HOST*/,
                      const int64_t k_str[4] /*HOST*/, const int64_t v_str[4] /*HOST*/, const float* W, int Bw,
                      int w_pitch, const float* sigma_dev_or_null, float sigma_host, const void* workspace, void* out,
                      const int64_t o_str[3] /*HOST*/, int B, int H, int L, int D, int S, float scale, int dtype,
                      void* stream);

/* One whole attention call: out = softmax(scale Q K^T + sigma * std(scale Q K^T) * W) V, i.e. dsc_xattn_stats followed
 * by dsc_xattn_forward with the same arguments -- and, when every CTA's share of Q fits in shared memory (the small
 * layers, small batches), ONE cooperative launch that reads Q once and keeps it on chip between the two passes.  This is
 * what the processor calls (replaces attention_modify.py:90-103 including the weight_func of app.py:1004).  The
 * statistics are left in the workspace exactly as by dsc_xattn_stats. */
int dsc_xattn_call(const void* q, const void* k, const void* v, const int64_t q_str[4] /*HOST*/,
                   const int64_t k_str[4] /*HOST*/, const int64_t v_str[4] /*HOST*/, const float* W, int Bw, int w_pitch,
                   const float* sigma_dev_or_null, float sigma_host, void* workspace, void* out,
                   const int64_t o_str[3] /*HOST*/, int B, int H, int L, int D, int S, float scale, int dtype,
                   void* stream);

/* dsc_xattn_call with the region map ALSO given in compact form: Wc is fp32 [Bw, L, DSC_COMPACT_PITCH] holding, for every
 * query row, the values of the n_active <= DSC_MAX_COMPACT_COLS key columns of W that are non-zero anywhere (column
 * active_cols[j], ascending, in slot j; remaining slots zero) -- with region prompts only the few tokens of the region
 * phrases carry weights (encode_region_map_function.py:57-63).  Where it applies (tcgen05 pass 2, S == 77) the kernels
 * then read 80 instead of 308 bytes of W per query row: the keys are permuted so that the active columns come first
 * (softmax and P V do not depend on the key order).  W itself must still be valid (other paths use it); Wc == NULL or
 * n_active == 0 means "no compact form".  Same result as dsc_xattn_call. */
#define DSC_MAX_COMPACT_COLS 16
#define DSC_COMPACT_PITCH 20
int dsc_xattn_call_cw(const void* q, const void* k, const void* v, const int64_t q_str[4] /*HOST*/,
                      const int64_t k_str[4] /*HOST*/, const int64_t v_str[4] /*HOST*/, const float* W, int Bw, int w_pitch,
                      const float* Wc, int n_active, const int32_t* active_cols /*HOST*/, const float* sigma_dev_or_null,
                      float sigma_host, void* workspace, void* out, const int64_t o_str[3] /*HOST*/, int B, int H, int L,
                      int D, int S, float scale, int dtype, void* stream);

/* dsc_xattn_call with an additive attention mask: out = softmax(a + sigma * std(a) * W) V, a = scale Q K^T + M.
 * mask: fp32 device pointer, element (b, h, l, s) at mask[b * mask_str[0] + h * mask_str[1] + l * mask_str[2] + s] (element
 * strides >= 0; 0 broadcasts the dimension: {0, 0, 0} is one [S] key bias, {H*S, S, 0} the [B*H, 1, S] tensor diffusers'
 * prepare_attention_mask returns, {H*L*S, L*S, S} a dense [B, H, L, S]); 4-byte aligned; values may be -inf (a fully masked
 * row is NaN, as in the reference).  The std is taken over a INCLUDING M, exactly as the reference does: its weight_func
 * receives the masked scores (attention_modify.py:90-95; baddbmm variant :39-70, :166).
 * What the reference's processors do with a mask (probed on the unmodified module, tests/test_oracle_attention.py):
 *   AttnProcessor (baddbmm, :107-207): M is the baddbmm input, beta = 1 -- this entry point.
 *   AttnProcessor2_0 (:414-503): a BOOL mask is never applied (:86-87 only rewrites the mask itself), a float mask is added
 *   in place into a [L, S] bias (:89), which raises for the 4-D tensor the processor builds (:452) and works only for masks
 *   that broadcast into [L, S] through the bare function (:74) -- mask_str = {0, 0, S or 0}.
 * Two launches (pass 1, pass 2) on the mma.sync kernels, any supported D and S; M is read per element through the strides
 * (a rarely used path: SD-1.5 passes no cross-attention mask). */
int dsc_xattn_call_masked(const void* q, const void* k, const void* v, const int64_t q_str[4] /*HOST*/,
                          const int64_t k_str[4] /*HOST*/, const int64_t v_str[4] /*HOST*/, const float* W, int Bw, int w_pitch,
                          const float* mask, const int64_t mask_str[3] /*HOST*/, const float* sigma_dev_or_null, float sigma_host,
                          void* workspace, void* out, const int64_t o_str[3] /*HOST*/, int B, int H, int L, int D, int S,
                          float scale, int dtype, void* stream);

/* ---- prepared K / V: the fast form of the call for the SD-1.5 shapes ---------------------------------------------------
 * K and V of a cross-attention layer are projections of the text embeddings: they do not change during a generation
 * (attention_modify.py:465-466 recomputes the same to_k / to_v on each of the 25 steps).  dsc_xattn_prepare_kv lays
 * them out ONCE as the shared-memory image the tcgen05 kernels multiply from (UMMA K-major chunks of K, V^T with a ones
 * row, keys permuted so that the active columns of the compact region map come first); dsc_xattn_call_prepared then
 * runs pass 1 and pass 2 on that image: each CTA fetches K / V^T with one bulk copy instead of re-gathering them.
 * Supported (dsc_xattn_prepared_supported != 0): D in {40, 80, 160} with H * D a multiple of 160, S == 77,
 * 1 <= n_active <= 16.  The image depends on k, v AND the active column list: prepare again when any of them changes. */
int dsc_xattn_prepared_supported(int H, int D, int S);
int dsc_xattn_kv_image_bytes(int B, int H, int D, int S, size_t* out /*HOST*/);
int dsc_xattn_prepare_kv(const void* k, const void* v, const int64_t k_str[4] /*HOST*/, const int64_t v_str[4] /*HOST*/,
                         int n_active, const int32_t* active_cols /*HOST*/, int B, int H, int D, int S, int dtype,
                         void* kv_image, void* stream);
/* passes: 1 = pass 1 only (std -> workspace), 2 = pass 2 only (std read from the workspace), 3 = the whole call.
 * Wc / n_active / Bw as in dsc_xattn_call_cw (the column list itself is baked into the image); same result as
 * dsc_xattn_call_cw on the k, v the image was prepared from.
 * passes == 3 is ONE cooperative launch (pass 1 and pass 2 as two phases of the same persistent CTAs; the std goes from
 * one to the other through per-CTA slots in the workspace that pass 2 polls and folds in a fixed order); when the device
 * cannot hold the grid (another context occupies SMs) or after dsc_config_set("no_fused", "1") it is two launches, pass 2
 * a programmatic dependent launch of pass 1.  The statistics header of the workspace is filled either way. */
#define DSC_PASS_STATS 1
#define DSC_PASS_FORWARD 2
/* Number of kernel launches dsc_xattn_call_prepared(passes = 3) issues under the current configuration: 1 or 2; -1 for an
 * unsupported shape.  Introspection only. */
int dsc_xattn_call_prepared_launches(int B, int H, int L, int D, int S);
int dsc_xattn_call_prepared(const void* q, const int64_t q_str[4] /*HOST*/, const void* kv_image, const float* Wc, int Bw,
                            int n_active, const float* sigma_dev_or_null, float sigma_host, void* workspace, void* out,
                            const int64_t o_str[3] /*HOST*/, int B, int H, int L, int D, int S, float scale, int dtype,
                            int passes, void* stream);

/* Number of kernel launches dsc_xattn_call will issue for this shape on the current device: 1 (fused), 2 (pass 1 +
 * pass 2) or 2 * chunks + 1 (long prompts); -1 for an unsupported shape.  Introspection only. */
int dsc_xattn_call_launches(int B, int H, int L, int D, int S);

/* Layout of the head of the workspace (read-only for callers). */
typedef struct dsc_xattn_stats_t {
  uint32_t ticket; /* internal, returns to 0 */
  uint32_t n_partials;
  float std_unbiased;
  float mean;
  double sum;
  double sumsq;
  double n;
} dsc_xattn_stats_t;

/* Region map, step 1 (encode_region_map_function.py:49-51): for each of R region maps
 * (uint8 [R, Hpx, Wpx], 255 = outside) compute bin = (map < 255), the OpenCV INTER_CUBIC resize of
 * bin to (w_r, h_r) and write ds[R, h_r*w_r] uint8 in {0,1}, plus any_set[R] (uint32, 1 if any
 * output pixel is 1 -- the `== max` rule needs it).  any_set must be zero on entry.  Bit-exact
 * against cv2 for integer scale factors (see oracle/region_map.py). */
int dsc_region_downsample(const uint8_t* maps, int R, int Hpx, int Wpx, int w_r, int h_r, uint8_t* ds,
                          uint32_t* any_set, void* stream);

/* Region map, step 2 (encode_region_map_function.py:51-69): W_out[L_r, n_tok] fp32 (zeroed by this
 * call), then for each span i in order:  W[:, start_i : start_i+len_i] = fp32(fp64(W) + val) with
 * val = (ds[region_i] == max(ds[region_i]) && weight != 0) ? weight : -mask_outsides  (fp64).
 * span_* are int32 device arrays of length n_spans, weight/mask_outsides fp64 device arrays [R]. */
int dsc_region_accumulate(const uint8_t* ds, const uint32_t* any_set, int R, int L_r, const double* weight,
                          const double* mask_outsides, const int32_t* span_region, const int32_t* span_start,
                          const int32_t* span_len, int n_spans, int n_tok, float* W_out, void* stream);

/* One DPM-Solver++(2M) step in VE (k-diffusion) space, fused with the denoiser scalings and CFG:
 *   eps   = eps_u + cfg * (eps_c - eps_u)            eps_uc = [2, n_elem] (uncond rows first), dtype
 *   den   = x - sigma * eps
 *   d     = first ? den : (1 + 1/(2r)) * den - 1/(2r) * den_prev
 *   x     = (sigma_next / sigma) * x - expm1(-h) * d          (x, den_prev: fp32, updated in place)
 *   unet_in_next[2, n_elem] = x / sqrt(sigma_next^2 + 1)       (dtype; both CFG halves; may be NULL)
 * with h = ln(sigma/sigma_next), r = ln(sigma_prev/sigma)/h; first!=0 or sigma_next==0 selects the
 * first-order update.  Scalars are HOST values; coefficients are formed in fp64 on the host. */
int dsc_dpmpp2m_step(float* x, const void* eps_uc, float* den_prev, void* unet_in_next_or_null, int64_t n_elem,
                     double sigma_prev, double sigma, double sigma_next, double cfg, int first, int dtype,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DSC_B200_H_ */
