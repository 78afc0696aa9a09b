// extern "C" surface of libdsc_b200.so (see include/dsc_b200.h for the contract of every entry point).
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "dsc_internal.h"

namespace dsc {

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
static int cuda_fail(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
  return static_cast<int>(e);
}

int sm_count_cached() {
  static thread_local int dev_cached = -1, sms = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != dev_cached) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    sms = v;
    dev_cached = dev;
  }
  return sms;
}

static int parse_impl(const char* v) {
  if (!v || !v[0]) return kImplAuto;
  if (strcmp(v, "mma") == 0) return kImplMma;
  if (strcmp(v, "tc5") == 0) return kImplTc5;
  if (strcmp(v, "gram") == 0) return kImplGram;
  return kImplAuto;
}
static bool parse_flag(const char* v) { return v && v[0] == '1'; }

static bool config_apply(Config& c, const char* key, const char* v) {
  if (strcmp(key, "xattn_impl") == 0) c.impl = parse_impl(v);
  else if (strcmp(key, "stats_impl") == 0) c.stats_impl = parse_impl(v);
  else if (strcmp(key, "no_fused") == 0) c.no_fused = parse_flag(v);
  else if (strcmp(key, "tc5_fused") == 0) c.tc5_fused = parse_flag(v);
  else if (strcmp(key, "no_pdl") == 0) c.no_pdl = parse_flag(v);
  else if (strcmp(key, "tc5_variant") == 0) c.tc5_x2 = v && v[0] == 'x' && v[1] == '2';
  else if (strcmp(key, "tc5_flags") == 0) c.tc5_flags = v ? static_cast<unsigned>(atoi(v)) : 0u;
  else return false;
  return true;
}

static Config& config_mut() {
  static Config c = [] {  // the one place the environment is read: once, at first use
    Config e;
    config_apply(e, "xattn_impl", getenv("DSC_XATTN_IMPL"));
    config_apply(e, "stats_impl", getenv("DSC_XATTN_STATS_IMPL"));
    config_apply(e, "no_fused", getenv("DSC_NO_FUSED"));
    config_apply(e, "tc5_fused", getenv("DSC_TC5_FUSED"));
    config_apply(e, "no_pdl", getenv("DSC_NO_PDL"));
    config_apply(e, "tc5_variant", getenv("DSC_TC5_VARIANT"));
    config_apply(e, "tc5_flags", getenv("DSC_TC5_FLAGS"));
    return e;
  }();
  return c;
}
const Config& config() { return config_mut(); }

// Kernel family per call.  xattn_impl=mma forces the legacy mma.sync kernels, =tc5 forces tcgen05/TMEM
// wherever it is implemented (D = 40, 80); default "auto" = whichever measured faster on B200 for the head dim
// (profiles/): tcgen05 at D = 40 and D = 80 (both passes; since pass 1 stages K by TMA it is level with mma.sync at
// D = 80 too), mma.sync elsewhere.  The two passes only share the std in the workspace, so the families mix freely.
static bool use_tc5(int D, bool stats) {
  const Config& c = config();
  int e = c.impl;
  if (stats && c.stats_impl != kImplAuto) e = c.stats_impl;  // pass 1 alone (A/B runs)
  if (e == kImplMma) return false;
  if (e == kImplTc5) return tc5_supports(D);
  return D == 40 || D == 80;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Shared validation of one [B,H,rows,D] operand viewed inside [B,rows,H*D].
static int check_bhxd(const char* name, const void* ptr, const int64_t s[4], int D) {
  if (!ptr || !s) return fail(DSC_ERR_INVALID_ARGUMENT, "%s: null pointer", name);
  if (!aligned16(ptr)) return fail(DSC_ERR_LAYOUT, "%s: base pointer must be 16-byte aligned", name);
  if (s[3] != 1) return fail(DSC_ERR_LAYOUT, "%s: stride(D) must be 1 (got %lld)", name, (long long)s[3]);
  if (s[1] != D) return fail(DSC_ERR_LAYOUT, "%s: stride(H) must equal D=%d (got %lld)", name, D, (long long)s[1]);
  if (s[2] % 8 != 0 || s[0] % 8 != 0)
    return fail(DSC_ERR_LAYOUT, "%s: stride(B)=%lld and stride(rows)=%lld must be multiples of 8 elements", name,
                (long long)s[0], (long long)s[2]);
  return DSC_OK;
}

static int check_dims(int B, int H, int L, int D, int S, int dtype) {
  if (B <= 0 || H <= 0 || L <= 0 || D <= 0 || S <= 0)
    return fail(DSC_ERR_INVALID_ARGUMENT, "non-positive dimension (B=%d H=%d L=%d D=%d S=%d)", B, H, L, D, S);
  if (dtype != DSC_DTYPE_F16 && dtype != DSC_DTYPE_BF16) return fail(DSC_ERR_INVALID_ARGUMENT, "bad dtype code %d", dtype);
  if (heads_per_group(D) == 0) return fail(DSC_ERR_UNSUPPORTED, "head dim %d not in {40,64,80,128,160}", D);
  if (S > DSC_MAX_KEYS_TOTAL) return fail(DSC_ERR_UNSUPPORTED, "S=%d keys > %d", S, DSC_MAX_KEYS_TOTAL);
  return DSC_OK;
}

static void fill_partition(XattnParams& p, int B, int H, int L, int D, int S) {
  const int G = heads_per_group(D);
  p.B = B;
  p.H = H;
  p.L = L;
  p.S = S;
  p.n_hg = (H + G - 1) / G;
  p.n_sl = (L + 15) / 16;
  p.total = static_cast<long long>(B) * p.n_hg * p.n_sl;
}

// Long prompts (S > DSC_MAX_KEYS; the reference concatenates 77-token windows, prompt_parser.py:161-194): the keys are
// processed as chunks of <= 80.  The workspace then also holds the per-chunk outputs and log-sum-exps.
static int n_chunks(int S) { return (S + DSC_MAX_KEYS - 1) / DSC_MAX_KEYS; }
static size_t stats_bytes() { return static_cast<size_t>(kHandoffOffset) + sizeof(double) * 2 * kMaxPartials; }  // header | partials | handoff slots
static size_t chunk_out_bytes(int B, int H, int L, int D) { return static_cast<size_t>(B) * L * H * D * 2; }
static size_t chunk_lse_bytes(int B, int H, int L) { return (static_cast<size_t>(B) * H * L * 4 + 15) / 16 * 16; }

}  // namespace dsc

using namespace dsc;

extern "C" {

int dsc_version(void) { return DSC_VERSION; }

const char* dsc_last_error(void) { return g_err; }

int dsc_sm_count(void) { return sm_count_cached(); }

int dsc_config_set(const char* key, const char* value_or_null) {
  if (!key) return fail(DSC_ERR_INVALID_ARGUMENT, "key is null");
  if (!config_apply(config_mut(), key, value_or_null)) return fail(DSC_ERR_INVALID_ARGUMENT, "unknown configuration key '%s'", key);
  return DSC_OK;
}

int dsc_xattn_workspace_bytes(int B, int H, int L, int D, int S, size_t* out) {
  if (!out) return fail(DSC_ERR_INVALID_ARGUMENT, "out is null");
  if (B <= 0 || H <= 0 || L <= 0 || D <= 0 || S <= 0) return fail(DSC_ERR_INVALID_ARGUMENT, "non-positive dimension");
  *out = stats_bytes();
  if (S > DSC_MAX_KEYS) *out += n_chunks(S) * (chunk_out_bytes(B, H, L, D) + chunk_lse_bytes(B, H, L));
  return DSC_OK;
}

}  // extern "C"

namespace dsc {
// Additive attention mask of a call: fp32, element (b, h, l, s) at m[b * sb + h * sh + l * sl + s] (0 = broadcast).
struct MaskArg {
  const float* m = nullptr;
  long long sb = 0, sh = 0, sl = 0;
};
static int check_mask(const MaskArg& mk) {
  if (!mk.m) return DSC_OK;
  if ((reinterpret_cast<uintptr_t>(mk.m) & 3) != 0) return fail(DSC_ERR_LAYOUT, "mask must be 4-byte aligned");
  if (mk.sb < 0 || mk.sh < 0 || mk.sl < 0) return fail(DSC_ERR_LAYOUT, "mask strides must be >= 0 (0 broadcasts a dimension)");
  return DSC_OK;
}
static int stats_impl(const void* q, const void* k, const int64_t q_str[4], const int64_t k_str[4], const MaskArg& mk, int B, int H,
                      int L, int D, int S, float scale, int dtype, void* workspace, void* stream);
}  // namespace dsc

extern "C" {

int dsc_xattn_stats(const void* q, const void* k, const int64_t q_str[4], const int64_t k_str[4],
                    const void* mask_or_null, int B, int H, int L, int D, int S, float scale, int dtype,
                    void* workspace, void* stream) {
  MaskArg mk;
  if (mask_or_null) {  // dense fp32 [B, H, L, S]
    mk.m = static_cast<const float*>(mask_or_null);
    mk.sl = S;
    mk.sh = static_cast<long long>(L) * S;
    mk.sb = static_cast<long long>(H) * L * S;
  }
  return stats_impl(q, k, q_str, k_str, mk, B, H, L, D, S, scale, dtype, workspace, stream);
}

}  // extern "C"

namespace dsc {
static int stats_impl(const void* q, const void* k, const int64_t q_str[4], const int64_t k_str[4], const MaskArg& mk, int B, int H,
                      int L, int D, int S, float scale, int dtype, void* workspace, void* stream) {
  int rc = check_dims(B, H, L, D, S, dtype);
  if (rc) return rc;
  if ((rc = check_mask(mk))) return rc;
  if (!workspace) return fail(DSC_ERR_INVALID_ARGUMENT, "workspace is null");
  if ((rc = check_bhxd("q", q, q_str, D))) return rc;
  if ((rc = check_bhxd("k", k, k_str, D))) return rc;
  XattnParams p{};
  fill_partition(p, B, H, L, D, S);
  p.q = q;
  p.k = k;
  p.q_sb = q_str[0];
  p.q_sl = q_str[2];
  p.k_sb = k_str[0];
  p.k_ss = k_str[2];
  p.scale = scale;
  p.mask = mk.m;  // with a mask: the mma.sync kernels, a = scale * Q K^T + M summed per element
  p.m_sb = mk.sb;
  p.m_sh = mk.sh;
  p.m_sl = mk.sl;
  p.ws = static_cast<Workspace*>(workspace);
  const int C = n_chunks(S);
  if (static_cast<long long>(stats_grid(p.total)) * C > kMaxPartials)
    return fail(DSC_ERR_UNSUPPORTED, "grid x key chunks exceeds the workspace's partial slots");
  p.n_total = static_cast<double>(B) * H * static_cast<double>(L) * S;
  p.fold_chunks = 1;
  cudaError_t e = cudaSuccess;
  if (C == 1) {
    const bool gram = config().stats_impl == kImplGram;
    if (mk.m) e = run_stats(p, D, dtype, static_cast<cudaStream_t>(stream));
    else if (gram && gram_supports(D, S)) e = run_stats_gram(p, D, dtype, static_cast<cudaStream_t>(stream));
    else
      e = use_tc5(D, true) ? run_stats_tc5(p, D, dtype, static_cast<cudaStream_t>(stream))
                           : run_stats(p, D, dtype, static_cast<cudaStream_t>(stream));
  } else {  // sums over key chunks; the last launch folds every chunk's partials into the std of the WHOLE call
    const size_t esz = 2;
    for (int c = 0; c < C && e == cudaSuccess; ++c) {
      p.k = static_cast<const char*>(k) + static_cast<size_t>(c) * DSC_MAX_KEYS * k_str[2] * esz;
      p.S = (c == C - 1) ? S - c * DSC_MAX_KEYS : DSC_MAX_KEYS;
      p.chunk = c;
      p.m_col0 = c * DSC_MAX_KEYS;
      p.fold_chunks = (c == C - 1) ? C : 0;
      e = run_stats(p, D, dtype, static_cast<cudaStream_t>(stream));
    }
  }
  return e == cudaSuccess ? DSC_OK : cuda_fail(e, "dsc_xattn_stats");
}
}  // namespace dsc

extern "C" {

// Which form dsc_xattn_call takes: 0 = two launches, 1 = single launch with Q resident in shared memory (mma.sync family,
// small problems), 2 = single launch, two phases over the tile list (tcgen05 family, D = 40 / 80).
static int call_form(int B, int H, int L, int D, int S) {
  const Config& c = config();
  if (n_chunks(S) > 1 || c.no_fused) return 0;
  const bool want_mma = c.impl == kImplMma, want_tc5 = c.impl == kImplTc5;
  if (!want_tc5 && fused_plan(B, H, L, D, S, nullptr)) return 1;
  // Form 2 is opt-in (DSC_TC5_FUSED=1): measured on B200 it only ties the two-launch form at D = 40 (62.3 vs 62.5 us --
  // programmatic dependent launch already hides pass 2's prologue behind the tail of pass 1) and loses at D = 80
  // (45 vs 42 us, where pass 1 is faster on the mma.sync kernel).
  if (!want_mma && c.tc5_fused && tc5_fused_supports(D)) return 2;
  return 0;
}

struct CompactW {
  const float* wc = nullptr;
  int n = 0;
  const int32_t* cols = nullptr;
};

static int forward_impl(const void* q, const void* k, const void* v, const int64_t q_str[4], const int64_t k_str[4],
                        const int64_t v_str[4], const float* W, int Bw, int w_pitch, const float* sigma_dev_or_null,
                        float sigma_host, const void* workspace, void* out, const int64_t o_str[3], int B, int H, int L, int D,
                        int S, float scale, int dtype, void* stream, bool with_stats, const char* who,
                        const CompactW& cw = CompactW(), const MaskArg& mk = MaskArg()) {
  int rc = check_dims(B, H, L, D, S, dtype);
  if (!rc) rc = check_mask(mk);
  if (cw.wc != nullptr && cw.n > 0) {
    if (cw.n > DSC_MAX_COMPACT_COLS || !cw.cols || !aligned16(cw.wc))
      return fail(DSC_ERR_INVALID_ARGUMENT, "compact map: need 1..%d columns, a column list and a 16-byte aligned base", DSC_MAX_COMPACT_COLS);
    for (int j = 0; j < cw.n; ++j)
      if (cw.cols[j] < 0 || cw.cols[j] >= S || (j > 0 && cw.cols[j] <= cw.cols[j - 1]))
        return fail(DSC_ERR_INVALID_ARGUMENT, "compact map: column list must be ascending and inside [0, S)");
  }
  if (rc) return rc;
  if (!workspace || !W || !out || !o_str) return fail(DSC_ERR_INVALID_ARGUMENT, "null pointer");
  if (Bw <= 0 || B % Bw != 0)
    return fail(DSC_ERR_SHAPE, "region map batch Bw=%d must divide the attention batch B=%d", Bw, B);
  if ((rc = check_bhxd("q", q, q_str, D))) return rc;
  if ((rc = check_bhxd("k", k, k_str, D))) return rc;
  if ((rc = check_bhxd("v", v, v_str, D))) return rc;
  if (!aligned16(out) || o_str[2] != 1 || o_str[1] % 8 != 0 || o_str[0] % 8 != 0)
    return fail(DSC_ERR_LAYOUT, "out: need 16-byte base, unit inner stride, row/batch strides multiple of 8");
  if ((reinterpret_cast<uintptr_t>(W) & 3) != 0) return fail(DSC_ERR_LAYOUT, "W must be 4-byte aligned");
  const int C = n_chunks(S);
  if (w_pitch < S || (C == 1 && w_pitch > DSC_MAX_KEYS))
    return fail(DSC_ERR_LAYOUT, "w_pitch=%d must be in [S=%d, %d]", w_pitch, S, DSC_MAX_KEYS);
  if (C > 1 && (w_pitch % 4 != 0 || !aligned16(W)))
    return fail(DSC_ERR_LAYOUT, "S=%d > %d keys: W rows must be 16-byte aligned (w_pitch %% 4 == 0, got %d)", S,
                DSC_MAX_KEYS, w_pitch);
  XattnParams p{};
  fill_partition(p, B, H, L, D, S);
  p.q = q;
  p.k = k;
  p.v = v;
  p.out = out;
  p.W = W;
  p.Bw = Bw;
  p.w_pitch = w_pitch;
  p.sigma_dev = sigma_dev_or_null;
  p.sigma_host = sigma_host;
  p.scale = scale;
  p.q_sb = q_str[0];
  p.q_sl = q_str[2];
  p.k_sb = k_str[0];
  p.k_ss = k_str[2];
  p.v_sb = v_str[0];
  p.v_ss = v_str[2];
  p.o_sb = o_str[0];
  p.o_sl = o_str[1];
  p.ws = const_cast<Workspace*>(static_cast<const Workspace*>(workspace));
  p.mask = mk.m;
  p.m_sb = mk.sb;
  p.m_sh = mk.sh;
  p.m_sl = mk.sl;
  if (cw.wc != nullptr && cw.n > 0) {
    p.wc = cw.wc;
    p.n_active = cw.n;
    for (int j = 0; j < cw.n; ++j) p.active_cols[j] = cw.cols[j];
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaSuccess;
  if (with_stats) {
    // one attention call: a single cooperative launch when the problem fits on chip, else pass 1 then pass 2
    const int form = mk.m ? 0 : call_form(B, H, L, D, S);  // a mask: pass 1 then pass 2 on the mma.sync kernels
    if (form != 0) {
      p.n_total = static_cast<double>(B) * H * static_cast<double>(L) * S;
      p.fold_chunks = 1;
      e = form == 1 ? run_fused(p, D, dtype, st) : run_fused_tc5(p, D, dtype, st);
      if (e != cudaErrorCooperativeLaunchTooLarge) return e == cudaSuccess ? DSC_OK : cuda_fail(e, who);
      (void)cudaGetLastError();  // the device cannot hold the whole grid right now (e.g. shared with another context):
      e = cudaSuccess;           // nothing was launched -- take the two-launch form below
    }
    rc = stats_impl(q, k, q_str, k_str, mk, B, H, L, D, S, scale, dtype, const_cast<void*>(workspace), stream);
    if (rc) return rc;
  }
  if (C == 1) {
    e = (!mk.m && use_tc5(D, false)) ? run_forward_tc5(p, D, dtype, st) : run_forward(p, D, dtype, st);
  } else {
    // per key chunk: softmax over the chunk's keys (with the std of the WHOLE call) -> dense chunk output + log2-sum-exp
    // in the workspace (layout: stats | C chunk outputs | C lse planes; see dsc_xattn_workspace_bytes); then one merge
    char* wsb = static_cast<char*>(const_cast<void*>(workspace));
    char* chunk_out = wsb + stats_bytes();
    float* lse = reinterpret_cast<float*>(chunk_out + C * chunk_out_bytes(B, H, L, D));
    const size_t lse_stride = chunk_lse_bytes(B, H, L) / 4;
    for (int c = 0; c < C && e == cudaSuccess; ++c) {
      p.k = static_cast<const char*>(k) + static_cast<size_t>(c) * DSC_MAX_KEYS * k_str[2] * 2;
      p.v = static_cast<const char*>(v) + static_cast<size_t>(c) * DSC_MAX_KEYS * v_str[2] * 2;
      p.S = (c == C - 1) ? S - c * DSC_MAX_KEYS : DSC_MAX_KEYS;
      p.w_col0 = c * DSC_MAX_KEYS;
      p.m_col0 = c * DSC_MAX_KEYS;
      p.out = chunk_out + c * chunk_out_bytes(B, H, L, D);
      p.o_sb = static_cast<long long>(L) * H * D;
      p.o_sl = static_cast<long long>(H) * D;
      p.lse = lse + c * lse_stride;
      e = run_forward(p, D, dtype, st);
    }
    if (e == cudaSuccess)
      e = run_merge_chunks(chunk_out, lse, C, out, o_str[0], o_str[1], B, H, L, D, dtype, st);
  }
  return e == cudaSuccess ? DSC_OK : cuda_fail(e, who);
}

int dsc_xattn_call_launches(int B, int H, int L, int D, int S) {
  if (B <= 0 || H <= 0 || L <= 0 || S <= 0 || heads_per_group(D) == 0 || S > DSC_MAX_KEYS_TOTAL) return -1;
  const int C = n_chunks(S);
  if (C > 1) return 2 * C + 1;
  return call_form(B, H, L, D, S) != 0 ? 1 : 2;
}

int dsc_xattn_forward(const void* q, const void* k, const void* v, const int64_t q_str[4], const int64_t k_str[4],
                      const int64_t v_str[4], const float* W, int Bw, int w_pitch, const float* sigma_dev_or_null,
                      float sigma_host, const void* workspace, void* out, const int64_t o_str[3], int B, int H, int L, int D,
                      int S, float scale, int dtype, void* stream) {
  return forward_impl(q, k, v, q_str, k_str, v_str, W, Bw, w_pitch, sigma_dev_or_null, sigma_host, workspace, out, o_str, B, H,
                      L, D, S, scale, dtype, stream, false, "dsc_xattn_forward");
}

int dsc_xattn_call(const void* q, const void* k, const void* v, const int64_t q_str[4], const int64_t k_str[4],
                   const int64_t v_str[4], const float* W, int Bw, int w_pitch, const float* sigma_dev_or_null, float sigma_host,
                   void* workspace, void* out, const int64_t o_str[3], int B, int H, int L, int D, int S, float scale, int dtype,
                   void* stream) {
  return forward_impl(q, k, v, q_str, k_str, v_str, W, Bw, w_pitch, sigma_dev_or_null, sigma_host, workspace, out, o_str, B, H,
                      L, D, S, scale, dtype, stream, true, "dsc_xattn_call");
}

int dsc_xattn_call_cw(const void* q, const void* k, const void* v, const int64_t q_str[4], const int64_t k_str[4],
                      const int64_t v_str[4], const float* W, int Bw, int w_pitch, const float* Wc, int n_active,
                      const int32_t* active_cols, const float* sigma_dev_or_null, float sigma_host, void* workspace, void* out,
                      const int64_t o_str[3], int B, int H, int L, int D, int S, float scale, int dtype, void* stream) {
  CompactW cw;
  cw.wc = Wc;
  cw.n = n_active;
  cw.cols = active_cols;
  return forward_impl(q, k, v, q_str, k_str, v_str, W, Bw, w_pitch, sigma_dev_or_null, sigma_host, workspace, out, o_str, B, H,
                      L, D, S, scale, dtype, stream, true, "dsc_xattn_call_cw", cw);
}

int dsc_xattn_call_masked(const void* q, const void* k, const void* v, const int64_t q_str[4], const int64_t k_str[4],
                          const int64_t v_str[4], const float* W, int Bw, int w_pitch, const float* mask, const int64_t mask_str[3],
                          const float* sigma_dev_or_null, float sigma_host, void* workspace, void* out, const int64_t o_str[3], int B,
                          int H, int L, int D, int S, float scale, int dtype, void* stream) {
  if (!mask || !mask_str) return fail(DSC_ERR_INVALID_ARGUMENT, "dsc_xattn_call_masked: mask / mask_str is null (use dsc_xattn_call)");
  MaskArg mk;
  mk.m = mask;
  mk.sb = mask_str[0];
  mk.sh = mask_str[1];
  mk.sl = mask_str[2];
  return forward_impl(q, k, v, q_str, k_str, v_str, W, Bw, w_pitch, sigma_dev_or_null, sigma_host, workspace, out, o_str, B, H, L, D,
                      S, scale, dtype, stream, true, "dsc_xattn_call_masked", CompactW(), mk);
}

int dsc_xattn_prepared_supported(int H, int D, int S) { return x3_supports(H, D, S) ? 1 : 0; }

int dsc_xattn_kv_image_bytes(int B, int H, int D, int S, size_t* out) {
  if (!out) return fail(DSC_ERR_INVALID_ARGUMENT, "out is null");
  if (B <= 0) return fail(DSC_ERR_INVALID_ARGUMENT, "non-positive batch");
  if (!x3_supports(H, D, S)) return fail(DSC_ERR_UNSUPPORTED, "prepared K/V: need D in {40, 80, 160}, S == 77, whole 160-column head groups (H=%d D=%d S=%d)", H, D, S);
  *out = x3_image_bytes(B, H, D);
  return DSC_OK;
}

static int check_cols(int n_active, const int32_t* cols, int S) {
  if (n_active < 0 || n_active > DSC_MAX_COMPACT_COLS || (n_active > 0 && !cols))
    return fail(DSC_ERR_INVALID_ARGUMENT, "compact map: need 0..%d columns and a column list", DSC_MAX_COMPACT_COLS);
  for (int j = 0; j < n_active; ++j)
    if (cols[j] < 0 || cols[j] >= S || (j > 0 && cols[j] <= cols[j - 1]))
      return fail(DSC_ERR_INVALID_ARGUMENT, "compact map: column list must be ascending and inside [0, S)");
  return DSC_OK;
}

int dsc_xattn_prepare_kv(const void* k, const void* v, const int64_t k_str[4], const int64_t v_str[4], int n_active,
                         const int32_t* active_cols, int B, int H, int D, int S, int dtype, void* kv_image, void* stream) {
  int rc = check_dims(B, H, 1, D, S, dtype);
  if (rc) return rc;
  if (!x3_supports(H, D, S)) return fail(DSC_ERR_UNSUPPORTED, "prepared K/V: need D in {40, 80, 160}, S == 77, whole 160-column head groups (H=%d D=%d S=%d)", H, D, S);
  if (!kv_image || !aligned16(kv_image)) return fail(DSC_ERR_INVALID_ARGUMENT, "kv_image must be a 16-byte aligned device buffer");
  if ((rc = check_bhxd("k", k, k_str, D))) return rc;
  if ((rc = check_bhxd("v", v, v_str, D))) return rc;
  if ((rc = check_cols(n_active, active_cols, S))) return rc;
  cudaError_t e = run_prepare_kv_x3(k, v, k_str[0], k_str[2], v_str[0], v_str[2], B, H, D, S, n_active, active_cols, dtype, kv_image,
                                    static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? DSC_OK : cuda_fail(e, "dsc_xattn_prepare_kv");
}

int dsc_xattn_call_prepared_launches(int B, int H, int L, int D, int S) {
  if (B <= 0 || H <= 0 || L <= 0 || !x3_supports(H, D, S)) return -1;
  return config().no_fused ? 2 : 1;
}

int dsc_xattn_call_prepared(const void* q, const int64_t q_str[4], const void* kv_image, const float* Wc, int Bw, int n_active,
                            const float* sigma_dev_or_null, float sigma_host, void* workspace, void* out, const int64_t o_str[3],
                            int B, int H, int L, int D, int S, float scale, int dtype, int passes, void* stream) {
  int rc = check_dims(B, H, L, D, S, dtype);
  if (rc) return rc;
  if (!x3_supports(H, D, S)) return fail(DSC_ERR_UNSUPPORTED, "prepared K/V: need D in {40, 80, 160}, S == 77, whole 160-column head groups (H=%d D=%d S=%d)", H, D, S);
  if (passes < 1 || passes > 3) return fail(DSC_ERR_INVALID_ARGUMENT, "passes must be 1, 2 or 3");
  if (!workspace || !kv_image || !aligned16(kv_image)) return fail(DSC_ERR_INVALID_ARGUMENT, "null / misaligned workspace or kv_image");
  if ((rc = check_bhxd("q", q, q_str, D))) return rc;
  if (!(scale > 0.f)) return fail(DSC_ERR_UNSUPPORTED, "prepared K/V path needs scale > 0");
  XattnParams p{};
  p.B = B;
  p.H = H;
  p.L = L;
  p.S = S;
  p.q = q;
  p.q_sb = q_str[0];
  p.q_sl = q_str[2];
  p.scale = scale;
  p.kv_image = kv_image;
  p.ws = static_cast<Workspace*>(workspace);
  p.handoff = passes == (DSC_PASS_STATS | DSC_PASS_FORWARD) ? 1 : 0;  // one call: pass 2 folds pass 1's partials itself
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaSuccess;
  if (passes & DSC_PASS_STATS) {
    if (stats_grid(static_cast<long long>(B) * (H * D / 160) * ((L + 127) / 128)) > kMaxPartials)
      return fail(DSC_ERR_UNSUPPORTED, "grid exceeds the workspace's partial slots");
  }
  if (passes & DSC_PASS_FORWARD) {
    if (!Wc || !aligned16(Wc) || !out || !o_str) return fail(DSC_ERR_INVALID_ARGUMENT, "pass 2 needs Wc (16-byte aligned) and out");
    if (n_active < 1 || n_active > DSC_MAX_COMPACT_COLS) return fail(DSC_ERR_INVALID_ARGUMENT, "n_active must be in 1..%d", DSC_MAX_COMPACT_COLS);
    if (Bw <= 0 || B % Bw != 0) return fail(DSC_ERR_SHAPE, "region map batch Bw=%d must divide the attention batch B=%d", Bw, B);
    if (!aligned16(out) || o_str[2] != 1 || o_str[1] % 8 != 0 || o_str[0] % 8 != 0)
      return fail(DSC_ERR_LAYOUT, "out: need 16-byte base, unit inner stride, row/batch strides multiple of 8");
    p.out = out;
    p.o_sb = o_str[0];
    p.o_sl = o_str[1];
    p.wc = Wc;
    p.n_active = n_active;
    p.Bw = Bw;
    p.sigma_dev = sigma_dev_or_null;
    p.sigma_host = sigma_host;
  }
  if (p.handoff && !config().no_fused) {
    // one call: both passes in ONE cooperative launch; if the device cannot hold the grid (it is sized to the SM count: only
    // when another context occupies SMs), fall back to pass 1 + pass 2 as two launches
    e = run_fused_x3(p, D, dtype, st);
    if (e == cudaSuccess) return DSC_OK;
    if (e != cudaErrorCooperativeLaunchTooLarge) return cuda_fail(e, "dsc_xattn_call_prepared (single launch)");
    (void)cudaGetLastError();
  }
  if (passes & DSC_PASS_STATS) {
    e = run_stats_x3(p, D, dtype, st);
    if (e != cudaSuccess) return cuda_fail(e, "dsc_xattn_call_prepared (pass 1)");
  }
  if (passes & DSC_PASS_FORWARD) {
    e = run_forward_x3(p, D, dtype, st);
    if (e != cudaSuccess) return cuda_fail(e, "dsc_xattn_call_prepared (pass 2)");
  }
  return DSC_OK;
}

int dsc_region_downsample(const uint8_t* maps, int R, int Hpx, int Wpx, int w_r, int h_r, uint8_t* ds,
                          uint32_t* any_set, void* stream) {
  if (R < 0 || Hpx <= 0 || Wpx <= 0 || w_r <= 0 || h_r <= 0) return fail(DSC_ERR_INVALID_ARGUMENT, "bad size");
  if (R == 0) return DSC_OK;
  if (!maps || !ds || !any_set) return fail(DSC_ERR_INVALID_ARGUMENT, "null pointer");
  cudaError_t e = run_region_downsample(maps, R, Hpx, Wpx, w_r, h_r, ds, any_set, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? DSC_OK : cuda_fail(e, "dsc_region_downsample");
}

int dsc_region_accumulate(const uint8_t* ds, const uint32_t* any_set, int R, int L_r, const double* weight,
                          const double* mask_outsides, const int32_t* span_region, const int32_t* span_start,
                          const int32_t* span_len, int n_spans, int n_tok, float* W_out, void* stream) {
  if (R < 0 || L_r <= 0 || n_tok <= 0 || n_spans < 0) return fail(DSC_ERR_INVALID_ARGUMENT, "bad size");
  if (!W_out) return fail(DSC_ERR_INVALID_ARGUMENT, "W_out is null");
  if (n_spans > 0 && (!ds || !any_set || !weight || !mask_outsides || !span_region || !span_start || !span_len))
    return fail(DSC_ERR_INVALID_ARGUMENT, "null pointer");
  cudaError_t e = run_region_accumulate(ds, any_set, R, L_r, weight, mask_outsides, span_region, span_start, span_len,
                                        n_spans, n_tok, W_out, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? DSC_OK : cuda_fail(e, "dsc_region_accumulate");
}

int dsc_dpmpp2m_step(float* x, const void* eps_uc, float* den_prev, void* unet_in_next_or_null, int64_t n_elem,
                     double sigma_prev, double sigma, double sigma_next, double cfg, int first, int dtype,
                     void* stream) {
  if (!x || !eps_uc || !den_prev) return fail(DSC_ERR_INVALID_ARGUMENT, "null pointer");
  if (n_elem < 0) return fail(DSC_ERR_INVALID_ARGUMENT, "negative n_elem");
  if (dtype != DSC_DTYPE_F16 && dtype != DSC_DTYPE_BF16) return fail(DSC_ERR_INVALID_ARGUMENT, "bad dtype code %d", dtype);
  if (!(sigma > 0.0) || sigma_next < 0.0) return fail(DSC_ERR_INVALID_ARGUMENT, "need sigma > 0 and sigma_next >= 0");
  StepCoef c{};
  const bool first_order = first != 0 || sigma_next == 0.0;
  if (!first_order && !(sigma_prev > sigma)) return fail(DSC_ERR_INVALID_ARGUMENT, "need sigma_prev > sigma");
  // t = -ln(sigma); h = t_next - t = ln(sigma / sigma_next)
  double c_x, c_d;
  if (sigma_next == 0.0) {
    c_x = 0.0;  // sigma_next / sigma
    c_d = 1.0;  // -expm1(-inf)
  } else {
    const double h = log(sigma / sigma_next);
    c_x = sigma_next / sigma;
    c_d = -expm1(-h);
  }
  double c_den = 1.0, c_prev = 0.0;
  if (!first_order) {
    const double h = log(sigma / sigma_next);
    const double h_last = log(sigma_prev / sigma);
    const double r = h_last / h;
    c_den = 1.0 + 1.0 / (2.0 * r);
    c_prev = -1.0 / (2.0 * r);
  }
  c.c_x = static_cast<float>(c_x);
  c.c_d = static_cast<float>(c_d);
  c.c_den = static_cast<float>(c_den);
  c.c_prev = static_cast<float>(c_prev);
  c.sigma = static_cast<float>(sigma);
  c.cfg = static_cast<float>(cfg);
  c.c_in_next = static_cast<float>(1.0 / sqrt(sigma_next * sigma_next + 1.0));
  cudaError_t e = run_dpmpp2m_step(x, eps_uc, den_prev, unet_in_next_or_null, n_elem, c, dtype,
                                   static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? DSC_OK : cuda_fail(e, "dsc_dpmpp2m_step");
}

}  // extern "C"
