// tcgen05 / TMA plumbing shared by the tcgen05 kernel files (xattn_tc5.cu, xattn_x3.cu): fences, commit, tensor-map box
// loads / stores, UMMA issue + descriptors, the tile-list decode, the deterministic fold of pass 1, host-side tensor maps.
#pragma once
#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

#include "dsc_device.cuh"
#include "dsc_internal.h"

#include <type_traits>

namespace dsc {

// ---------------------------------------------------------------- tcgen05 plumbing
// programmatic dependent launch (PDL): pass 2 is launched while pass 1 still runs; it may do everything that does
// not need the std (barrier init, TMEM alloc, K/V staging, first Q tiles, first QK^T) and blocks here before beta.
__device__ __forceinline__ void pdl_wait_prior_grid() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// TMA tensor-map (3-D: columns, rows, batch) box load / store
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, int c0, int c1, int c2, uint32_t bar,
                                            uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, "
      "%5}], [%2], %6;" ::"r"(dst),
      "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, int c0, int c1, int c2, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(m), "r"(src),
               "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]^T   (kind::f16: fp16 or bf16 inputs, fp32 accumulate)
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// shared-memory matrix descriptor, K-major, no swizzle: 8x16B core matrices, LBO between the two K chunks
// of a k16 step, SBO between 8-row groups
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return static_cast<uint64_t>((addr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(lbo >> 4) << 16) |
         (static_cast<uint64_t>(sbo >> 4) << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor with independent A / B formats (0 = F16, 1 = BF16)
__host__ __device__ constexpr uint32_t idesc_ab(uint32_t afmt, uint32_t bfmt, int n) {
  return (1u << 4) | (afmt << 7) | (bfmt << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
}
template <typename T>
__host__ __device__ constexpr uint32_t idesc_f16(int n) {
  const uint32_t fmt = std::is_same<T, __half>::value ? 0u : 1u;  // 0 = F16, 1 = BF16
  // [4,6) D format = F32 | [7,10) A format | [10,13) B format | bits 15/16 = 0: A, B K-major | [17,23) N>>3 | [24,29) M>>4
  return (1u << 4) | (fmt << 7) | (fmt << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
}

struct Item {
  int b, hg, nheads, l0, rows, tile;
};
// Last step of pass 1, executed by ONE WARP of the last-arriving CTA: fold the per-CTA fp64 partials in a fixed
// order (lane-strided loads, then a fixed shuffle tree: deterministic, and ~5 loads deep instead of a 148-long
// serial chain of L2 round trips) and publish std / mean.
__device__ __forceinline__ void finalize_stats(const XattnParams& p, const double* partials, int lane) {
  double sa = 0.0, sb = 0.0;
  for (unsigned int c = lane; c < gridDim.x; c += 32) {
    sa += __ldcg(partials + 2 * c);
    sb += __ldcg(partials + 2 * c + 1);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sa += __shfl_xor_sync(0xffffffffu, sa, o);
    sb += __shfl_xor_sync(0xffffffffu, sb, o);
  }
  if (lane == 0) {
    const double scl = static_cast<double>(p.scale);
    const double n = static_cast<double>(p.B) * p.H * static_cast<double>(p.L) * p.S;
    const double sum = sa * scl, sumsq = sb * scl * scl, mean = sum / n;
    double var = (n > 1.0) ? (sumsq - sum * mean) / (n - 1.0) : nan("");
    if (var < 0.0) var = 0.0;
    p.ws->std_unbiased = static_cast<float>(sqrt(var));
    p.ws->mean = static_cast<float>(mean);
    p.ws->sum = sum;
    p.ws->sumsq = sumsq;
    p.ws->n = n;
    p.ws->n_partials = gridDim.x;
    __threadfence();
    p.ws->ticket = 0u;  // reusable without a memset
  }
}

// ---- host: tensor maps -------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}
// [B, L, cols] 16-bit tensor with element strides (sb, sl, 1) -> boxes of 32 columns x 128 rows, 64B swizzle
inline bool make_map(CUtensorMap* m, const void* base, int cols, int L, int B, long long sl, long long sb) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  cuuint64_t gdim[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(L), static_cast<cuuint64_t>(B)};
  cuuint64_t gstr[2] = {static_cast<cuuint64_t>(sl) * 2, static_cast<cuuint64_t>(B > 1 ? sb : sl * L) * 2};
  cuuint32_t box[3] = {32, 128, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<void*>(base), gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// padded region-weight map fp32 [Bw, L, 80] -> boxes of 84 columns (4 out of range: zero-filled) x 128 rows
inline bool make_map_w(CUtensorMap* m, const float* base, int L, int Bw) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  cuuint64_t gdim[3] = {DSC_MAX_KEYS, static_cast<cuuint64_t>(L), static_cast<cuuint64_t>(Bw)};
  cuuint64_t gstr[2] = {DSC_MAX_KEYS * 4ull, static_cast<cuuint64_t>(L) * DSC_MAX_KEYS * 4ull};
  cuuint32_t box[3] = {84, 128, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// K [B, S, cols] 16-bit (element strides sb, ss, 1) -> boxes of 8 columns (16 B) x 80 keys: one UMMA K-major chunk column
inline bool make_map_k(CUtensorMap* m, const void* base, int cols, int S, int B, long long ss, long long sb) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  cuuint64_t gdim[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(S), static_cast<cuuint64_t>(B)};
  cuuint64_t gstr[2] = {static_cast<cuuint64_t>(ss) * 2, static_cast<cuuint64_t>(B > 1 ? sb : ss * S) * 2};
  cuuint32_t box[3] = {8, DSC_MAX_KEYS, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<void*>(base), gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// compact region map fp32 [Bw, L, 20] -> boxes of 20 columns x 128 rows
inline bool make_map_wc(CUtensorMap* m, const float* base, int L, int Bw) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  cuuint64_t gdim[3] = {DSC_COMPACT_PITCH, static_cast<cuuint64_t>(L), static_cast<cuuint64_t>(Bw)};
  cuuint64_t gstr[2] = {DSC_COMPACT_PITCH * 4ull, static_cast<cuuint64_t>(L) * DSC_COMPACT_PITCH * 4ull};
  cuuint32_t box[3] = {DSC_COMPACT_PITCH, 128, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace dsc
