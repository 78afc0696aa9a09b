// Region-masked cross-attention on the 5th-generation tensor cores (tcgen05.mma + TMEM), D in {40, 80}.
//
// Same two passes as xattn_kernels.cu (pass 1: std of scale*QK^T, pass 2: softmax(scale*QK^T + beta*W) V),
// replacing scaled_dot_product_attention_regionstate (reference source/modules/attention_modify.py:74-103)
// and weight_func (reference source/app.py:1004).  Why a second implementation: the legacy mma.sync path
// is tensor-pipe / issue bound on B200 (HMMA measured at 554 TFLOP/s, about the rate the whole call needs);
// tcgen05 is ~4x faster and takes the operand traffic off the LSU, which leaves HBM as the bound.
//
// Work unit = tile of 128 query rows x one head group (G*D = 160 columns: 4 heads at D=40, 2 at D=80).
// One persistent CTA per SM owns a contiguous range of the (batch, head-group)-major tile list.
//
// Two kernel organisations live in this file:
//   * xattn_tc5x4_* (default): FOUR consumer warpgroups, one head each, 640 threads -- see the block comment above
//     x4_phase() for the roles, the TMEM column budget, the compact region map (keys permuted so that the weighted
//     columns come first, 80-byte W rows, 3-stage ring), K staged by TMA in pass 1, the two-part publication of P, the
//     piecewise O read, programmatic dependent launch, and the opt-in single-launch form of both passes.
//   * xattn_tc5_kernel (DSC_TC5_VARIANT=x2, kept for A/B runs): TWO consumer warpgroups that alternate heads with
//     software pipelining:
//       warp 8        producer : TMA tensor-map loads (cp.async.bulk.tensor, 64B-swizzled 32-column boxes) of the Q
//                                tile + one bulk copy of the W tile into a 2/3-stage smem ring, tensor-map stores of
//                                finished O tiles.  (Measured, profiles/r1_tma_copy_rate.jsonl: a 1-D bulk copy costs
//                                ~30 ns of TMA issue per SM whatever its size, so per-row copies cap at 1.6-3 TB/s;
//                                boxes stream at > 5 TB/s.)
//       warps 9, 10   MMA      : one elected thread per warpgroup issues tcgen05.mma (M=128, N=80 for S=QK^T; N=48/96
//                                for O=PV), tcgen05.commit -> mbarrier
//       warps 0-3/4-7 consumers: ONE THREAD PER QUERY ROW (TMEM lane = row); warpgroup g takes the heads h = g (mod 2)
//                     of the tile, so QK^T/PV of one head overlaps the softmax of the other:
//                       Q_h row: smem -> registers -> tcgen05.st (A operand lives in TMEM)
//                       S row  : tcgen05.ld 80 fp32 -> + beta*W row (registers, shared by the heads) -> max / exp2 /
//                                sum entirely in-thread (no shuffles) -> P (fp16/bf16) -> tcgen05.st (A of PV)
//                       O row  : tcgen05.ld -> * 1/sum -> overwrite the Q_h columns of the row in smem
// In both, K_h and V_h^T of the head group stay resident in shared memory in the UMMA canonical K-major no-swizzle
// layout (8-row x 16-byte core matrices; chunk pitch = LBO, 128 B between 8-row groups = SBO); the odd half
// k-step of D=40 multiplies an explicit zero chunk.
#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

#include "tc5_common.cuh"
#include "tc5_tmem.cuh"

#include <stdio.h>
#include <stdlib.h>

#include <type_traits>

namespace dsc {

constexpr float kLog2eT = 1.4426950408889634f;

template <int D>
struct TC {
  static_assert(D == 40 || D == 80, "tcgen05 path: head dim 40 or 80");
  static constexpr int G = 160 / D;
  static constexpr int GW = 160;
  static constexpr int ROWS = 128;
  // Q/O tile in smem = 5 TMA boxes of [128 rows x 32 columns (64 B)], SWIZZLE_64B: the 16-byte chunk c of
  // row r of a box sits at r*64 + ((c ^ ((r>>1)&3)) << 4), so 8 consecutive rows (one quarter-warp of
  // row-per-thread accesses) always hit 8 distinct 16-byte bank groups.  Dense: no padding bytes.
  static constexpr int BOX_COLS = 32;
  static constexpr int NBOX = GW / BOX_COLS;
  static constexpr int BOX_BYTES = ROWS * BOX_COLS * 2;
  static constexpr int QT_BYTES = NBOX * BOX_BYTES;
  static constexpr int WT_BYTES = ROWS * DSC_MAX_KEYS * 4;
  static constexpr int DCH = D / 8;                      // 16-byte chunks per head row
  static constexpr int KSTEPS = (D + 15) / 16;           // k16 steps of QK^T
  static constexpr int KCH = KSTEPS * 2;                 // chunks per head incl. the zero pad chunk
  static constexpr int K_CH_BYTES = DSC_MAX_KEYS * 16;   // one chunk column: 80 keys x 16 B
  static constexpr int K_HEAD_BYTES = KCH * K_CH_BYTES;
  // O = P [V | 1]: row D of V^T is all ones, so column D of O is the softmax row sum (of the rounded P that the
  // tensor core actually multiplies); UMMA N must be a multiple of 16 -> 48 / 96 rows
  static constexpr int N_PV = (D == 40) ? 48 : 96;
  static constexpr int VT_CH_BYTES = N_PV * 16;          // one chunk column of V^T: N_PV rows x (8 keys) 16 B
  static constexpr int VT_HEAD_BYTES = 10 * VT_CH_BYTES;
  static constexpr int K_BYTES = G * K_HEAD_BYTES;
  static constexpr int VT_BYTES_RAW = G * VT_HEAD_BYTES;
  // round K + V^T up so that the stage ring (swizzled boxes) starts 1024-byte aligned
  static constexpr int VT_BYTES = ((K_BYTES + VT_BYTES_RAW + 1023) / 1024) * 1024 - K_BYTES;
  static constexpr int K_BYTES_PAD = ((K_BYTES + 1023) / 1024) * 1024;  // stats pass: K only
  // TMEM columns of one warpgroup
  static constexpr int S_COL = 0;    // S = Q K^T (fp32, 80 columns)
  static constexpr int P_COL = 80;   // P (16-bit pairs, 40 columns): separate from S so QK^T of the next head can start early
  static constexpr int O_COL = 120;  // O (fp32, N_PV columns)
  static constexpr int QA_COL = 120 + N_PV;
  static constexpr int QA_COLS = KCH * 4;
  static constexpr int WG_COLS = 256;
  static_assert(QA_COL + QA_COLS <= WG_COLS, "TMEM budget");
  static constexpr int BAR_BYTES = 256;
  static constexpr int FWD_STAGES = 2;
  static constexpr int STATS_STAGES = 3;
  static constexpr int FWD_SMEM = K_BYTES + VT_BYTES + FWD_STAGES * (QT_BYTES + WT_BYTES) + BAR_BYTES;
  static constexpr int STATS_SMEM = K_BYTES_PAD + STATS_STAGES * QT_BYTES + BAR_BYTES;
};

// byte offset of the 16-byte chunk `cg` (0 .. GW/8-1) of tile row `r` inside the swizzled Q/O tile
template <int D>
__device__ __forceinline__ uint32_t tile_chunk_off(int r, int cg) {
  return (cg >> 2) * TC<D>::BOX_BYTES + r * 64 + (((cg & 3) ^ ((r >> 1) & 3)) << 4);
}

// item index (32-bit: the launcher guarantees total < 2^31) -> (batch, head group, 128-row tile)
template <int D>
__device__ __forceinline__ Item decode(int idx, const XattnParams& p) {
  Item it;
  const int seg = idx / p.n_sl;
  it.tile = idx - seg * p.n_sl;
  it.b = seg / p.n_hg;
  it.hg = seg - it.b * p.n_hg;
  it.nheads = min(TC<D>::G, p.H - it.hg * TC<D>::G);
  it.l0 = it.tile * TC<D>::ROWS;
  it.rows = min(TC<D>::ROWS, p.L - it.l0);
  return it;
}

#ifdef DSC_WATCHDOG
// Debug build only: a barrier wait that gives up after ~1 s, records who was waiting on what in the
// workspace debug words (offset 48: {tag | block << 8 | warp << 24, parity}) and lets the kernel drain.
__device__ unsigned int g_wd_abort = 0;
__device__ unsigned int g_wd_info[2] = {0, 0};
__device__ __forceinline__ void mbar_wait_wd(uint32_t bar, uint32_t parity, uint32_t tag, Workspace* ws) {
  const long long t0 = clock64();
  while (true) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
    if (*reinterpret_cast<volatile unsigned int*>(&g_wd_abort)) return;
    if (clock64() - t0 > (1ll << 31)) {
      if (atomicCAS(&g_wd_abort, 0u, 1u) == 0u) {
        unsigned int* d = reinterpret_cast<unsigned int*>(reinterpret_cast<unsigned char*>(ws) + 48);
        d[0] = tag | (blockIdx.x << 8) | ((threadIdx.x >> 5) << 24);
        d[1] = parity | 0x100u;
      }
      return;
    }
  }
}
#define MBAR_WAIT(bar, parity, tag) mbar_wait_wd(bar, parity, tag, p.ws)
#else
#define MBAR_WAIT(bar, parity, tag) mbar_wait(bar, parity)
#endif
// Service warps (producer, MMA issuers) poll with a short sleep between probes: a hot spin loop would
// compete with the two consumer warps of the same sub-partition for issue slots and the barrier unit.
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
#ifdef DSC_WATCHDOG
  const long long t0 = clock64();
#endif
  while (true) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
#ifdef DSC_WATCHDOG
    if (*reinterpret_cast<volatile unsigned int*>(&g_wd_abort)) return;
    if (clock64() - t0 > (1ll << 31)) {
      if (atomicCAS(&g_wd_abort, 0u, 1u) == 0u) {
        g_wd_info[0] = 100u | (blockIdx.x << 8) | ((threadIdx.x >> 5) << 24);
        g_wd_info[1] = (bar & 0xffffu) | (parity << 16) | ((threadIdx.x & 31) << 20);
      }
      return;
    }
#endif
    __nanosleep(64);
  }
}

#ifdef DSC_TRACE
// Debug build only: clock64 timeline of block 0 (lane 0 of warps 0, 4, 9, 10) -> g_trace[warp][slot] = {tag, clock}.
// The slot counter lives in a register (tr_n) so that a trace point costs one clock read and one store.
__device__ long long g_trace[4][512][2];
__device__ int g_trace_n[4];
#define TRACE_DECL                                                                                        \
  int tr_n = 0;                                                                                           \
  const int tr_k = (blockIdx.x == 0 && (threadIdx.x & 31) == 0)                                           \
                       ? ((threadIdx.x >> 5) == 0 ? 0 : (threadIdx.x >> 5) == 4 ? 1 : (threadIdx.x >> 5) == 9 ? 2 \
                          : (threadIdx.x >> 5) == 10 ? 3 : -1)                                            \
                       : -1;
#define TRACE(tag)                                  \
  do {                                              \
    if (tr_k >= 0 && tr_n < 512) {                  \
      g_trace[tr_k][tr_n][0] = (tag);               \
      g_trace[tr_k][tr_n][1] = clock64();           \
      g_trace_n[tr_k] = ++tr_n;                     \
    }                                               \
  } while (0)
#define TRACE_DECL_X4                                                                                     \
  int tr_n = 0;                                                                                           \
  const int tr_k = blockIdx.x != 0 ? -1                                                                  \
                   : threadIdx.x == 0 ? 0 : threadIdx.x == 128 ? 1 : threadIdx.x == 512 ? 2                \
                   : threadIdx.x == 19 * 32 + 16 ? 3 : -1;
#else
#define TRACE_DECL
#define TRACE_DECL_X4
#define TRACE(tag) do {} while (0)
#endif

constexpr int kConsumerThreads = 256;
constexpr int kThreads = 384;       // warps 0-7 consumers | 8 producer | 9, 10 MMA issuers | 11 idle (fills the warpgroup)
constexpr int kConsumerRegs = 224;  // setmaxnreg: 8 x 32 x 224 + 4 x 32 x 56 = 64512 <= 65536
constexpr int kServiceRegs = 56;


// Pull the K / V head group of (batch, head group) `it` towards L2 (one 128-byte line per request): issued at kernel
// start for the first run and right after each restage for the following one, so the staging loads hit L2.
template <typename T, int D, bool STATS, int NTHR>
__device__ __forceinline__ void prefetch_kv(const XattnParams& p, const Item& it, int ctid) {
  using C = TC<D>;
  const int row_bytes = it.nheads * D * 2;
  const char* kb = reinterpret_cast<const char*>(reinterpret_cast<const T*>(p.k) + it.b * p.k_sb + it.hg * C::GW);
  const char* vb = reinterpret_cast<const char*>(reinterpret_cast<const T*>(p.v) + it.b * p.v_sb + it.hg * C::GW);
  for (int key = ctid; key < p.S; key += NTHR) {  // one thread per key row (<= 3 lines of 128 B), no divisions
    for (int l = 0; l < row_bytes; l += 128) {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(kb + static_cast<long long>(key) * p.k_ss * 2 + l));
      if constexpr (!STATS) asm volatile("prefetch.global.L2 [%0];" ::"l"(vb + static_cast<long long>(key) * p.v_ss * 2 + l));
    }
  }
}

// K_h -> canonical [chunk][key][16 B]; V_h -> V^T canonical [key chunk][d][8 keys] (+ the ones row), by NTHR
// consumer threads.  All loads of a thread are issued before their first use (clamped addresses, no predicated
// loads), and the thread->piece maps are chosen so that every shared-memory store is bank-conflict free:
//   K  : piece = (head, chunk, key), key fastest  -> consecutive lanes store 16 B apart
//   V^T: item  = (head, key chunk, d), d fastest  -> a thread gathers the 8 keys of one d (8 two-byte loads, lanes
//        coalesce along d) and stores one 16-byte row; consecutive lanes store 16 B apart.  (Scattering two-byte
//        elements of a row-major piece instead costs 8-way conflicts: measured 7.4k cycles per restage.)
#ifdef DSC_TRACE
__device__ long long g_kv_trace[8];
__device__ unsigned long long g_cta_times[160][2];  // globaltimer (ns) at CTA start / end
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define CTA_TIME(k) do { if (threadIdx.x == 0 && blockIdx.x < 160) g_cta_times[blockIdx.x][k] = gtimer(); } while (0)
#define KV_TRACE(tag) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_kv_trace[(tag) - 31] = clock64(); } while (0)
#else
#define CTA_TIME(k) do {} while (0)
#define KV_TRACE(tag) do {} while (0)
#endif
template <typename T, int D, bool STATS, int NTHR = 256>
__device__ __forceinline__ void stage_kv(unsigned char* smem, const XattnParams& p, const Item& it, int ctid,
                                         const unsigned char* perm = nullptr) {
  using C = TC<D>;
  constexpr int MAXK = (C::G * DSC_MAX_KEYS * C::DCH + NTHR - 1) / NTHR;  // pieces incl. the pad keys (stored as zeros)
  constexpr int NPAIR = D / 2 + 1;  // column pairs of V incl. the (ones, zero) pair that forms row D
  constexpr int MAXV = (C::G * 10 * NPAIR + NTHR - 1) / NTHR;
  KV_TRACE(31);
  // 32-bit element offsets from the (batch, head group) base: the index math must stay cheap, this runs on every
  // thread for every piece (64-bit multiplies here cost more issue slots than the loads themselves)
  const T* __restrict__ kg = reinterpret_cast<const T*>(p.k) + it.b * p.k_sb + it.hg * C::GW;
  const int kss = static_cast<int>(p.k_ss), vss = static_cast<int>(p.v_ss);
  const int n_k = it.nheads * C::DCH * DSC_MAX_KEYS;  // pieces (head, chunk, key slot 0..79), key fastest
  uint4 kv[MAXK];
  int ksoff[MAXK];
#pragma unroll
  for (int u = 0; u < MAXK; ++u) {
    const int e = min(ctid + u * NTHR, n_k - 1);
    const int hc = e / DSC_MAX_KEYS, key = e - hc * DSC_MAX_KEYS;  // constant divisors only
    const int h = hc / C::DCH, c = hc - h * C::DCH;
    ksoff[u] = h * C::K_HEAD_BYTES + c * C::K_CH_BYTES + key * 16;
    const int slot = min(key, p.S - 1);
    const int krow = perm ? perm[slot] : slot;  // key slot -> row of K (compact region map: weighted columns first)
    kv[u] = __ldg(reinterpret_cast<const uint4*>(kg + (krow * kss + h * D + c * 8)));
    if (key >= p.S) kv[u] = make_uint4(0, 0, 0, 0);  // pad keys: exact zero scores
  }
  uint32_t ve[STATS ? 1 : MAXV][8];  // ve[u][j] = V[key 8*kc+j][d, d+1] (two 16-bit values)
  int vsoff[STATS ? 1 : MAXV];
  if constexpr (!STATS) {
    const T* __restrict__ vg = reinterpret_cast<const T*>(p.v) + it.b * p.v_sb + it.hg * C::GW;
    const int n_v = it.nheads * 10 * NPAIR;
    const uint32_t one = std::is_same<T, __half>::value ? 0x3C00u : 0x3F80u;
#pragma unroll
    for (int u = 0; u < MAXV; ++u) {
      const int e = min(ctid + u * NTHR, n_v - 1);
      const int hk = e / NPAIR, dp = e - hk * NPAIR;  // constant divisors
      const int h = hk / 10, kc = hk - h * 10;
      vsoff[u] = h * C::VT_HEAD_BYTES + kc * C::VT_CH_BYTES + dp * 32;
      const int col = h * D + min(dp, D / 2 - 1) * 2;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int key = kc * 8 + j;
        const uint32_t x = __ldg(reinterpret_cast<const uint32_t*>(vg + (min(key, p.S - 1) * vss + col)));
        ve[u][j] = key >= p.S ? 0u : (dp == D / 2 ? one : x);  // pad keys multiply nothing; pair D/2 = (1, 0)
      }
    }
  }
  KV_TRACE(32);
#pragma unroll
  for (int u = 0; u < MAXK; ++u)
    if (ctid + u * NTHR < n_k) *reinterpret_cast<uint4*>(smem + ksoff[u]) = kv[u];
  KV_TRACE(33);
  if constexpr (!STATS) {
    const int n_v = it.nheads * 10 * NPAIR;
#pragma unroll
    for (int u = 0; u < MAXV; ++u) {
      if (ctid + u * NTHR < n_v) {
        uint4 lo, hi;  // row d (low halves) and row d+1 (high halves), 8 keys each
        lo.x = __byte_perm(ve[u][0], ve[u][1], 0x5410); hi.x = __byte_perm(ve[u][0], ve[u][1], 0x7632);
        lo.y = __byte_perm(ve[u][2], ve[u][3], 0x5410); hi.y = __byte_perm(ve[u][2], ve[u][3], 0x7632);
        lo.z = __byte_perm(ve[u][4], ve[u][5], 0x5410); hi.z = __byte_perm(ve[u][4], ve[u][5], 0x7632);
        lo.w = __byte_perm(ve[u][6], ve[u][7], 0x5410); hi.w = __byte_perm(ve[u][6], ve[u][7], 0x7632);
        unsigned char* dst = smem + C::K_BYTES + vsoff[u];
        *reinterpret_cast<uint4*>(dst) = lo;
        *reinterpret_cast<uint4*>(dst + 16) = hi;
      }
    }
  }
  KV_TRACE(34);
  fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
  KV_TRACE(35);
}

template <typename T, int D, bool STATS>
__global__ void __launch_bounds__(kThreads, 1)
xattn_tc5_kernel(const XattnParams p, const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_o) {
  using C = TC<D>;
  constexpr int NST = STATS ? C::STATS_STAGES : C::FWD_STAGES;
  constexpr int STAGE_BYTES = STATS ? C::QT_BYTES : (C::QT_BYTES + C::WT_BYTES);
  constexpr int KV_BYTES = STATS ? C::K_BYTES_PAD : (C::K_BYTES + C::VT_BYTES);
  static_assert(KV_BYTES % 1024 == 0 && STAGE_BYTES % 1024 == 0, "swizzled boxes need aligned bases");
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  TRACE_DECL
  if constexpr (STATS) pdl_launch_dependents();
  {  // first thing: start pulling this CTA's first K / V head group into L2
    const int begin0 = static_cast<int>(p.total * blockIdx.x / gridDim.x);
    if (tid < kConsumerThreads && begin0 < p.total) prefetch_kv<T, D, STATS, kConsumerThreads>(p, decode<D>(begin0, p), tid);
  }
  const uint32_t s0 = smem_u32(smem);
  const uint32_t sStage = s0 + KV_BYTES;
  const uint32_t bars = sStage + NST * STAGE_BYTES;
  // barrier map (8 B each): full[NST] | odone[NST] | qrdy[2] | srdy[2] | prdy[2] | ordy[2] ; tmem ptr after
  const uint32_t b_full = bars, b_odone = bars + 8 * NST, b_qrdy = bars + 16 * NST, b_srdy = b_qrdy + 16,
                 b_prdy = b_qrdy + 32, b_ordy = b_qrdy + 48;
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + KV_BYTES + NST * STAGE_BYTES + 192);

  TRACE(1);
  // one-time init: zero K / V^T (pad keys, pad chunk), barriers, TMEM allocation
  for (int i = tid; i < KV_BYTES / 16; i += kThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(b_full + 8 * s, 1);
      mbar_init(b_odone + 8 * s, kConsumerThreads);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(b_qrdy + 8 * g, 128);
      mbar_init(b_srdy + 8 * g, 1);
      mbar_init(b_prdy + 8 * g, 128);
      mbar_init(b_ordy + 8 * g, 1);
    }
    fence_mbar_init();
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
                     smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  TRACE(2);

  const int begin = static_cast<int>(p.total * blockIdx.x / gridDim.x);
  const int n_items = static_cast<int>(p.total * (blockIdx.x + 1) / gridDim.x) - begin;

  if (warp >= 8) {
    // registers go to the consumer warpgroups: the service warps need very few
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kServiceRegs));
    if (warp == 8) {
      // ============================== producer: TMA loads and stores ===============================
      const uint64_t pol = STATS ? policy_evict_last() : policy_evict_first();
      auto store_tile = [&](int i) {  // O tile of item i leaves through the TMA (pass 2 only)
        if constexpr (!STATS) {
          if (lane == 0) {
            const Item it = decode<D>(begin + i, p);
            const uint32_t sQ = sStage + (i % NST) * STAGE_BYTES;
#pragma unroll
            for (int j = 0; j < C::NBOX; ++j)  // rows >= L and columns >= H*D are clipped by the TMA
              tma_store_3d(&tm_o, it.hg * C::GW + j * C::BOX_COLS, it.l0, it.b, sQ + j * C::BOX_BYTES);
            bulk_commit();
            bulk_wait_read0();
          }
          __syncwarp();
        }
      };
      for (int i = 0; i < n_items; ++i) {
        const int s = i % NST;
        if (i >= NST) {
          if (lane == 0) mbar_wait_relaxed(b_odone + 8 * s, ((i / NST) - 1) & 1);
          __syncwarp();
          store_tile(i - NST);
        }
        if (lane == 0) {
          const Item it = decode<D>(begin + i, p);
          const uint32_t sQ = sStage + s * STAGE_BYTES;
          uint32_t tx = C::QT_BYTES;  // out-of-bounds parts of a box are zero-filled and still counted
          const float* wsrc = nullptr;
          uint32_t wbytes = 0;
          if constexpr (!STATS) {
            wsrc = p.W + (static_cast<long long>(it.b / (p.B / p.Bw)) * p.L + it.l0) * p.w_pitch;
            wbytes = it.rows * p.w_pitch * 4;
            if (((reinterpret_cast<uintptr_t>(wsrc) | wbytes) & 15) == 0) tx += wbytes; else wbytes = 0;
          }
          mbar_arrive_expect_tx(b_full + 8 * s, tx);
#pragma unroll
          for (int j = 0; j < C::NBOX; ++j)
            tma_load_3d(sQ + j * C::BOX_BYTES, &tm_q, it.hg * C::GW + j * C::BOX_COLS, it.l0, it.b, b_full + 8 * s, pol);
          if (wbytes != 0) bulk_g2s_hint(sQ + C::QT_BYTES, wsrc, wbytes, b_full + 8 * s, pol);
        }
        __syncwarp();
      }
      for (int i = (n_items > NST ? n_items - NST : 0); i < n_items; ++i) {
        if (lane == 0) mbar_wait_relaxed(b_odone + 8 * (i % NST), (i / NST) & 1);
        __syncwarp();
        store_tile(i);
      }
      if (lane == 0) bulk_wait0();
    } else if (warp <= 10) {
      // ============================== MMA issuer of warpgroup g (one thread) =======================
      // Mirrors the consumer pipeline: S(n+1) = Q K^T is issued while the consumers work on head n.
      const int g = warp - 9;
      if (lane == 0) {
        constexpr uint32_t idesc_qk = idesc_f16<T>(80);
        constexpr uint32_t idesc_pv = idesc_f16<T>(C::N_PV);
        const uint32_t tw = tmem_base + g * C::WG_COLS;
        uint32_t nq = 0, np = 0;
        auto issue_qk = [&](int h) {
          TRACE(20);
          mbar_wait_relaxed(b_qrdy + 8 * g, nq & 1);
          TRACE(21);
          ++nq;
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < C::KSTEPS; ++ks)
            umma_ts(tw + C::S_COL, tw + C::QA_COL + ks * 8,
                    smem_desc(s0 + h * C::K_HEAD_BYTES + ks * 2 * C::K_CH_BYTES, C::K_CH_BYTES, 128), idesc_qk, ks);
          tc_commit(b_srdy + 8 * g);
          TRACE(22);
        };
        for (int r0 = 0; r0 < n_items;) {  // runs of consecutive tiles that share one (batch, head group)
          const Item it0 = decode<D>(begin + r0, p);
          const int r1 = min(n_items, r0 + p.n_sl - it0.tile);
          const int hpw = it0.nheads > g ? (it0.nheads - g + 1) >> 1 : 0;  // heads of this warpgroup per tile: 0, 1 or 2
          const int n_pairs = (r1 - r0) * hpw;
          if (n_pairs > 0) issue_qk(g);
          for (int n = 0; n < n_pairs; ++n) {
            if (n + 1 < n_pairs) issue_qk(hpw == 2 ? g + 2 * ((n + 1) & 1) : g);
            if constexpr (!STATS) {
              const int h = hpw == 2 ? g + 2 * (n & 1) : g;
              TRACE(23);
              mbar_wait_relaxed(b_prdy + 8 * g, np & 1);
              TRACE(24);
              ++np;
              tc_fence_after();
#pragma unroll
              for (int kk = 0; kk < 5; ++kk)
                umma_ts(tw + C::O_COL, tw + C::P_COL + kk * 8,
                        smem_desc(s0 + C::K_BYTES + h * C::VT_HEAD_BYTES + kk * 2 * C::VT_CH_BYTES, C::VT_CH_BYTES, 128),
                        idesc_pv, kk);
              tc_commit(b_ordy + 8 * g);
              TRACE(25);
            }
          }
          r0 = r1;
        }
      }
      __syncwarp();
    }
  } else {
    // ============================== consumers: one thread per query row =============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kConsumerRegs));
    const int g = warp >> 2;                       // warpgroup
    const int row = (warp & 3) * 32 + lane;        // tile row == TMEM lane
    const uint32_t row_off = row * 64, row_sw = (row >> 1) & 3;  // swizzled position of this row's chunks
    auto chunk_off = [&](int cg) -> uint32_t {
      return (cg >> 2) * C::BOX_BYTES + row_off + ((static_cast<uint32_t>(cg & 3) ^ row_sw) << 4);
    };
    const uint32_t tw = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + g * C::WG_COLS;
    float beta_l2 = 0.f;  // read after pass 1 has completed (PDL), right before first use
    bool have_beta = STATS;
    const float scale_l2 = p.scale * kLog2eT;
    if constexpr (D % 16 == 8) {  // zero the K-padding columns of the A operand once
      uint32_t z[4] = {0, 0, 0, 0};
      tmem_st_x4(tw + C::QA_COL + D / 2, z);
      tc_wait_st();
    }
    double dsum = 0.0, dsq = 0.0;
    uint32_t n_s = 0, n_o = 0;

    // Q_h row of (item i, head h): swizzled smem -> registers -> TMEM (A operand of S = Q K^T); signals the MMA warp
    auto stage_q = [&](int i, int h) {
      const unsigned char* qtile = smem + KV_BYTES + (i % NST) * STAGE_BYTES;
      uint32_t qw[D / 2];
#pragma unroll
      for (int c = 0; c < C::DCH; ++c) {
        const uint4 v = *reinterpret_cast<const uint4*>(qtile + chunk_off(h * C::DCH + c));
        qw[4 * c] = v.x; qw[4 * c + 1] = v.y; qw[4 * c + 2] = v.z; qw[4 * c + 3] = v.w;
      }
      if constexpr (D == 40) {
        tmem_st_x16(tw + C::QA_COL, qw);
        tmem_st_x4(tw + C::QA_COL + 16, qw + 16);
      } else {
        tmem_st_x32(tw + C::QA_COL, qw);
        tmem_st_x8(tw + C::QA_COL + 32, qw + 32);
      }
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(b_qrdy + 8 * g);
    };
    // O row of the previous head: TMEM -> * 1/rowsum (the ones-row of V^T put the sum in column D) -> smem tile
    auto drain_o = [&](int i, int h) {
      if constexpr (!STATS) {
        MBAR_WAIT(b_ordy + 8 * g, n_o & 1, 5);
        ++n_o;
        tc_fence_after();
        float o[D + 8];
        if constexpr (D == 40) {
          tmem_ld_x32(tw + C::O_COL, reinterpret_cast<uint32_t*>(o));
          tmem_ld_x16(tw + C::O_COL + 32, reinterpret_cast<uint32_t*>(o) + 32);
        } else {
          tmem_ld_x64(tw + C::O_COL, reinterpret_cast<uint32_t*>(o));
          tmem_ld_x16(tw + C::O_COL + 64, reinterpret_cast<uint32_t*>(o) + 64);
          tmem_ld_x8(tw + C::O_COL + 80, reinterpret_cast<uint32_t*>(o) + 80);
        }
        tc_wait_ld();
        const float inv = 1.f / o[D];
        unsigned char* qtile = smem + KV_BYTES + (i % NST) * STAGE_BYTES;
#pragma unroll
        for (int c = 0; c < C::DCH; ++c) {  // O_h overwrites Q_h of this row (same swizzled chunks)
          uint4 v;
          v.x = Mma<T>::pack(o[8 * c] * inv, o[8 * c + 1] * inv);
          v.y = Mma<T>::pack(o[8 * c + 2] * inv, o[8 * c + 3] * inv);
          v.z = Mma<T>::pack(o[8 * c + 4] * inv, o[8 * c + 5] * inv);
          v.w = Mma<T>::pack(o[8 * c + 6] * inv, o[8 * c + 7] * inv);
          *reinterpret_cast<uint4*>(qtile + chunk_off(h * C::DCH + c)) = v;
        }
      }
    };
    auto release_item = [&](int i) {  // this thread is done with the tile of item i
      if constexpr (!STATS) fence_proxy_async();  // O rows -> visible to the TMA store
      mbar_arrive(b_odone + 8 * (i % NST));
    };

    for (int r0 = 0; r0 < n_items;) {  // runs of consecutive tiles that share one (batch, head group)
      const Item it0 = decode<D>(begin + r0, p);
      const int r1 = min(n_items, r0 + p.n_sl - it0.tile);
      // new (batch, head group): restage K / V^T (every MMA on the old ones has been consumed)
      TRACE(3);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      stage_kv<T, D, STATS>(smem, p, it0, tid);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      TRACE(4);
      if (r1 < n_items) prefetch_kv<T, D, STATS, kConsumerThreads>(p, decode<D>(begin + r1, p), tid);
      const int hpw = it0.nheads > g ? (it0.nheads - g + 1) >> 1 : 0;  // heads of this warpgroup per tile: 0, 1 or 2
      if (hpw == 0) {  // this warpgroup has no head in these tiles: just hand them back
        for (int i = r0; i < r1; ++i) {
          MBAR_WAIT(b_full + 8 * (i % NST), (i / NST) & 1, 6);
          release_item(i);
        }
        r0 = r1;
        continue;
      }
      const float* wbase = STATS ? nullptr : p.W + (static_cast<long long>(it0.b / (p.B / p.Bw)) * p.L) * p.w_pitch;
      MBAR_WAIT(b_full + 8 * (r0 % NST), (r0 / NST) & 1, 7);
      TRACE(5);
      stage_q(r0, g);
      TRACE(6);
      if (STATS && hpw == 1) release_item(r0);
      float bw[STATS ? 1 : 80];
      for (int i = r0; i < r1; ++i) {
        const int l0 = (it0.tile + (i - r0)) * C::ROWS;
        const int rows = min(C::ROWS, p.L - l0);
        for (int hi = 0; hi < hpw; ++hi) {
          const int h = g + 2 * hi;
          const bool first = (i == r0) && (hi == 0);
          // ---- S row of this head
          TRACE(10);
          MBAR_WAIT(b_srdy + 8 * g, n_s & 1, 8);
          TRACE(11);
          ++n_s;
          tc_fence_after();
          float sc[80];
          tmem_ld_x64(tw + C::S_COL, reinterpret_cast<uint32_t*>(sc));
          tmem_ld_x16(tw + C::S_COL + 64, reinterpret_cast<uint32_t*>(sc) + 64);
          tc_wait_ld();
          TRACE(12);
          // The previous head's O normally leaves TMEM after this head's softmax (its PV has long finished by
          // then).  If that head was the last one of the PREVIOUS tile, its stage must be handed back before
          // we wait for the tile after this one (2-stage ring), so it is drained first.
          bool o_pending = !STATS && !first;
          if (o_pending && hi == 0) {
            drain_o(i - 1, g + 2 * (hpw - 1));
            release_item(i - 1);
            o_pending = false;
          }
          // ---- look ahead: Q of the next head goes to TMEM now, its QK^T runs under this head's softmax
          if (hi + 1 < hpw) {
            stage_q(i, h + 2);
            if (STATS) release_item(i);  // pass 1 only reads Q; h + 2 is the last head of this warpgroup
          } else if (i + 1 < r1) {
            MBAR_WAIT(b_full + 8 * ((i + 1) % NST), ((i + 1) / NST) & 1, 9);
            stage_q(i + 1, g);
            if (STATS && hpw == 1) release_item(i + 1);
          }
          TRACE(13);
          if constexpr (STATS) {
            float fs[4] = {0.f, 0.f, 0.f, 0.f}, fq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 80; ++j) {
              fs[j & 3] += sc[j];
              fq[j & 3] = fmaf(sc[j], sc[j], fq[j & 3]);
            }
            if (row < rows) {  // pad keys contribute exact zeros; rows beyond L were zero-filled by the TMA
              dsum += static_cast<double>((fs[0] + fs[1]) + (fs[2] + fs[3]));
              dsq += static_cast<double>((fq[0] + fq[1]) + (fq[2] + fq[3]));
            }
          } else {
            if (!have_beta) {
              pdl_wait_prior_grid();
              const float sigma = p.sigma_dev ? __ldcg(p.sigma_dev) : p.sigma_host;
              beta_l2 = sigma * __ldcg(&p.ws->std_unbiased) * kLog2eT;
              have_beta = true;
            }
            if (hi == 0) {  // beta*W row (log2 domain), shared by this row's heads; keys >= S get -inf
              const float* wsrc = wbase + static_cast<long long>(l0) * p.w_pitch;
              const bool bulk = ((reinterpret_cast<uintptr_t>(wsrc) | static_cast<uintptr_t>(rows * p.w_pitch * 4)) & 15) == 0;
              if (bulk) {  // all loads first (independent, conflict-free: row pitch = S words), then the scaling
                const float* wt = reinterpret_cast<const float*>(smem + KV_BYTES + (i % NST) * STAGE_BYTES + C::QT_BYTES) + row * p.w_pitch;
                if (p.S == 77) {
#pragma unroll
                  for (int j = 0; j < 77; ++j) bw[j] = wt[j];
                  bw[77] = bw[78] = bw[79] = -INFINITY;
                } else {
#pragma unroll
                  for (int j = 0; j < 80; ++j) bw[j] = (j < p.S) ? wt[j] : -INFINITY;
                }
              } else {
                const float* wr = wsrc + static_cast<long long>(row < rows ? row : 0) * p.w_pitch;
#pragma unroll
                for (int j = 0; j < 80; ++j) bw[j] = (j < p.S) ? __ldg(wr + j) : -INFINITY;
              }
#pragma unroll
              for (int j = 0; j < 80; ++j) bw[j] *= beta_l2;  // -inf stays -inf (beta >= 0); beta = 0 handled below
              if (beta_l2 == 0.f) {
#pragma unroll
                for (int j = 0; j < 80; ++j) bw[j] = (j < p.S) ? 0.f : -INFINITY;
              }
            }
            float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int j = 0; j < 80; ++j) {
              sc[j] = fmaf(sc[j], scale_l2, bw[j]);
              mx[j & 3] = fmaxf(mx[j & 3], sc[j]);
            }
            const float m = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
            uint32_t pw[40];
#pragma unroll
            for (int j = 0; j < 40; ++j) pw[j] = Mma<T>::pack(ex2_approx(sc[2 * j] - m), ex2_approx(sc[2 * j + 1] - m));
            TRACE(14);
            // ---- O of the previous head (same tile) must leave TMEM before P lets the next PV overwrite it
            if (o_pending) drain_o(i, h - 2);
            TRACE(15);
            tmem_st_x32(tw + C::P_COL, pw);
            tmem_st_x8(tw + C::P_COL + 32, pw + 32);
            tc_wait_st();
            tc_fence_before();
            mbar_arrive(b_prdy + 8 * g);
            TRACE(16);
          }
        }
      }
      TRACE(7);
      if constexpr (!STATS) {
        drain_o(r1 - 1, g + 2 * (hpw - 1));
        release_item(r1 - 1);
      }
      TRACE(8);
      r0 = r1;
    }
    if constexpr (STATS) {
      // CTA partial in a fixed order (warp shuffle tree, then warps 0..7 serially) -> workspace
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
        dsq += __shfl_xor_sync(0xffffffffu, dsq, o);
      }
      double* red = reinterpret_cast<double*>(smem);  // K region is dead: every MMA has been consumed
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (lane == 0) {
        red[warp] = dsum;
        red[8 + warp] = dsq;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      double* partials = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(p.ws) + kWorkspaceHeader);
      if (warp == 0) {
        unsigned int last = 0;
        if (lane == 0) {
          double a = 0.0, b = 0.0;
          for (int w = 0; w < 8; ++w) {
            a += red[w];
            b += red[8 + w];
          }
          partials[2 * blockIdx.x] = a;
          partials[2 * blockIdx.x + 1] = b;
          __threadfence();
          last = atomicAdd(&p.ws->ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {  // last CTA: fold all partials (deterministic order)
          __threadfence();
          finalize_stats(p, partials, lane);
        }
      }
    }
  }

  // teardown: everyone is done with TMEM before it is released
  TRACE(9);
  tc_fence_before();
  __syncthreads();
  TRACE(30);
  tc_fence_after();
  if (warp == 8) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// =============================================================================================
// Four consumer warpgroups, ONE HEAD EACH: xattn_tc5x4_kernel (D = 40 and D = 80).
//
// Same producer / TMA ring / resident K, V^T / TMEM-operand scheme as xattn_tc5_kernel, but 16 consumer warps
// (4 per SM sub-partition instead of 2) hide the TMEM and mbarrier round trips of the per-head chain
//   Q row -> TMEM | S = Q K^T | S row -> softmax -> P -> TMEM | O = P [V|1] | O row -> smem
// behind each other.  TMEM: 128 columns per warpgroup: S (80 fp32; P overwrites its first 40 columns once the row is
// in registers) and O (48 fp32; the Q operand occupies its first 24 / 40 columns until S is done).  The beta*W row is
// streamed from the shared W tile (96 registers per thread here).
//   D = 40: a 160-column tile holds 4 heads -> warpgroup g = head g, all four on the same tile.
//   D = 80: a tile holds 2 heads -> warpgroups {0,1} take the even tiles of the CTA's list, {2,3} the odd ones (two
//           tiles in flight, one per ring stage); O = P V is issued as two N = 48 halves (40 V columns + the ones
//           column) so that the O region stays 48 columns wide.  V^T is staged as 40-column "virtual heads" in both
//           cases.
// Service warps 16..19: lane 0 of warp 16+g issues the tensor-core work of warpgroup g; lane 16 of warp 19 is the TMA
// producer (its first ring fill is issued before the CTA-wide start-up barrier).
// (A fifth service warp does not fit: 21 warps x 96 registers cannot be launched -- registers are handed out per
//  4-warp group -- and tensor-core issue from the consumer warps themselves measured 10 % slower: five tcgen05.mma
//  plus the commit keep the issuing warp busy for 400-600 cycles per batch.)
// Region map W, three forms: compact (CW = true; what the processor passes: only the <= 16 weighted key columns, 80 B
// per query row, keys permuted at K / V^T staging so that they are slots 0..15, 3-stage ring), padded dense (rows 80
// floats apart, one 84 x 128 TMA box per tile, 128-bit reads), dense (the reference's pitch-77 tensor: bulk copy, scalar
// reads).  Pass 1 stages K by TMA (one 8-column x 80-key box per UMMA chunk column); pass 2 stages K and V^T with the
// consumer threads (transpose + key permutation).  P is published in two parts (keys 0..47 | 48..79) so that P V starts
// under the last exponentials; at D = 40 the next tile's Q row is fetched into registers during P V and goes to TMEM
// after a partial O read.  Pass 1 -> pass 2 (and predecessor -> pass 1) are programmatic dependent launches.
#ifndef DSC_POLY_PATTERN
#define DSC_POLY_PATTERN 0x00  // bit i: key pairs with (pair & 7) == i take the polynomial 2^x below (0 = all on MUFU)
#endif
constexpr int kEpochOffsetT = 48;  // epoch word of the single-launch grid barrier, behind the public workspace header
constexpr int kX4Consumers = 512;
constexpr int kX4Threads = 640;
// 640 threads -> 96 registers per thread from __launch_bounds__; the consumer path fits, so no setmaxnreg here (a
// CTA's register pool is what it was launched with: 640 x 96 leaves nothing to hand over).

template <int D>
struct X4 {
  using C = TC<D>;
  static constexpr int HPT = 160 / D;  // heads per tile
  static constexpr int PAR = 4 / HPT;  // tile-parity groups of warpgroups
  static constexpr int NH = D / 40;    // O = P V halves (40 V columns each)
  static constexpr int QW = D / 2;     // 32-bit words of one Q row-head
  static constexpr int VH_BYTES = TC<40>::VT_HEAD_BYTES;  // V^T of one 40-column virtual head (10 chunks x 48 rows x 16 B)
  static constexpr int VT_OFF = C::K_BYTES;
  static constexpr int KV_FWD = ((C::K_BYTES + 4 * VH_BYTES + 1023) / 1024) * 1024;
  // W tile: 128 rows at a pitch of 84 floats (336 B = 21 x 16 B, 21 odd: a quarter-warp's 128-bit reads of 8
  // consecutive rows hit 8 distinct bank groups); the tensor-map box is 84 wide over rows of 80 floats, the 4
  // out-of-range columns arrive as zeros.  Dense (unpadded) W tiles are bulk-copied at their own pitch <= 80.
  static constexpr int W_SMEM_PITCH = 84;
  static constexpr int WT_BYTES = C::ROWS * W_SMEM_PITCH * 4;
  static_assert(WT_BYTES % 1024 == 0, "stage alignment");
  static constexpr int FWD_SMEM = KV_FWD + C::FWD_STAGES * (C::QT_BYTES + WT_BYTES) + C::BAR_BYTES;
  // compact region map (only the <= 16 key columns that carry weights; keys permuted so that they come first): the W tile
  // shrinks from 42 KB to 10 KB (128 rows x 20 floats: 16 values + 4 pad, pitch 80 B = 5 x 16 B, conflict-free for
  // 128-bit reads) and a THIRD ring stage fits
  static constexpr int CW_PITCH = DSC_COMPACT_PITCH;
  static constexpr int CWT_BYTES = C::ROWS * CW_PITCH * 4;
  static constexpr int CW_STAGES = 3;
  static constexpr int CW_PERM_OFF = KV_FWD + CW_STAGES * (C::QT_BYTES + CWT_BYTES) + C::BAR_BYTES;  // 80-byte slot -> key table
  static constexpr int CW_SMEM = CW_PERM_OFF + 128;
  static_assert((C::QT_BYTES + CWT_BYTES) % 1024 == 0 && CW_SMEM <= 227 * 1024, "compact-W stage");
  static constexpr int STATS_SMEM = C::K_BYTES_PAD + C::STATS_STAGES * C::QT_BYTES + C::BAR_BYTES;
};

// V (160 columns of the head group = n_vh virtual heads of 40) -> V^T canonical [key chunk][row d][8 keys], row 40 = ones
template <typename T, int NTHR>
__device__ __forceinline__ void stage_vt40(unsigned char* sVt, const XattnParams& p, const Item& it, int gw, int n_vh,
                                           int ctid, const unsigned char* perm = nullptr) {
  constexpr int NPAIR = 21;  // 20 column pairs + the (ones, zero) pair
  constexpr int MAXV = (4 * 10 * NPAIR + NTHR - 1) / NTHR;
  const T* __restrict__ vg = reinterpret_cast<const T*>(p.v) + it.b * p.v_sb + it.hg * gw;
  const int vss = static_cast<int>(p.v_ss);
  const int n_v = n_vh * 10 * NPAIR;
  const uint32_t one = std::is_same<T, __half>::value ? 0x3C00u : 0x3F80u;
  uint32_t ve[MAXV][8];
  int vsoff[MAXV];
#pragma unroll
  for (int u = 0; u < MAXV; ++u) {
    const int e = min(ctid + u * NTHR, n_v - 1);
    const int hk = e / NPAIR, dp = e - hk * NPAIR;
    const int vh = hk / 10, kc = hk - vh * 10;
    vsoff[u] = vh * TC<40>::VT_HEAD_BYTES + kc * TC<40>::VT_CH_BYTES + dp * 32;
    const int col = vh * 40 + min(dp, 19) * 2;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int key = kc * 8 + j;
      const int slot = min(key, p.S - 1);
      const int vrow = perm ? perm[slot] : slot;
      const uint32_t x = __ldg(reinterpret_cast<const uint32_t*>(vg + (vrow * vss + col)));
      ve[u][j] = key >= p.S ? 0u : (dp == 20 ? one : x);
    }
  }
#pragma unroll
  for (int u = 0; u < MAXV; ++u) {
    if (ctid + u * NTHR < n_v) {
      uint4 lo, hi;
      lo.x = __byte_perm(ve[u][0], ve[u][1], 0x5410); hi.x = __byte_perm(ve[u][0], ve[u][1], 0x7632);
      lo.y = __byte_perm(ve[u][2], ve[u][3], 0x5410); hi.y = __byte_perm(ve[u][2], ve[u][3], 0x7632);
      lo.z = __byte_perm(ve[u][4], ve[u][5], 0x5410); hi.z = __byte_perm(ve[u][4], ve[u][5], 0x7632);
      lo.w = __byte_perm(ve[u][6], ve[u][7], 0x5410); hi.w = __byte_perm(ve[u][6], ve[u][7], 0x7632);
      *reinterpret_cast<uint4*>(sVt + vsoff[u]) = lo;
      *reinterpret_cast<uint4*>(sVt + vsoff[u] + 16) = hi;
    }
  }
  fence_proxy_async();
}

// One pass (STATS: pass 1, else pass 2) as a device function, so that it can be a kernel of its own (MODE 0) or one of the
// two phases of the single-launch kernel below (MODE 1: pass 1 followed by a grid barrier, keeps the TMEM allocation and
// returns its base; MODE 2: pass 2 on that allocation, no programmatic-dependent-launch handshake).
template <typename T, int D, bool STATS, int MODE, bool CW = false>
__device__ __forceinline__ uint32_t x4_phase(const XattnParams& p, const CUtensorMap& tm_q, const CUtensorMap& tm_o,
                                             const CUtensorMap& tm_w, uint32_t tmem_in, unsigned int epoch0 = 0u) {
  using C = TC<D>;
  using X = X4<D>;
  constexpr int HPT = X::HPT, PAR = X::PAR, NH = X::NH;
  static_assert(!(CW && STATS), "the compact region map only concerns pass 2");
  constexpr int NST = STATS ? C::STATS_STAGES : (CW ? X::CW_STAGES : C::FWD_STAGES);
  constexpr int STAGE_BYTES = STATS ? C::QT_BYTES : (C::QT_BYTES + (CW ? X::CWT_BYTES : X::WT_BYTES));
  constexpr int KV_BYTES = STATS ? C::K_BYTES_PAD : X::KV_FWD;
  constexpr int S_COL = 0, O_COL = 80, WG_COLS = 128;  // P aliases S, the Q operand aliases O
  constexpr int STAGE_CONSUMERS = HPT * 128;           // threads that hand a ring stage back
  // pass 1, D = 40: the 48 columns behind S hold TWO Q-operand buffers (24 columns each), so the next tile's Q rows
  // are in TMEM before this tile's S has been read and its Q K^T starts the moment the S columns are free
  constexpr bool QDB = STATS && D == 40;
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  TRACE_DECL_X4
  const bool warp_arrive = (p.flags & 1u) != 0, issuer_spin = (p.flags & 2u) != 0;
  // padded W (rows 80 floats apart, 16-byte aligned, 77 keys): TMA box per tile, 128-bit reads, packed fp32 math
  const bool w_fast = !STATS && (p.flags & 4u) != 0;
  TRACE(1);
  CTA_TIME(0);
  if constexpr (STATS && MODE == 0) {
    // pass 1 is itself launched as a programmatic dependent of whatever precedes it in the stream: its CTAs may be
    // placed while that kernel drains; nothing of its inputs is touched before the predecessor has completed
    pdl_wait_prior_grid();
    pdl_launch_dependents();  // let pass 2 start its prologue on SMs as they free up
  }
  if constexpr (!STATS && MODE == 0) pdl_launch_dependents();  // a following pass 1 (next call) may be placed as SMs free up
  {  // first thing: start pulling this CTA's first K / V head group into L2
    const int begin0 = static_cast<int>(p.total * blockIdx.x / gridDim.x);
    if (tid < kX4Consumers && begin0 < p.total) prefetch_kv<T, D, STATS, kX4Consumers>(p, decode<D>(begin0, p), tid);
  }
  TRACE(50);
  const uint32_t s0 = smem_u32(smem);
  const uint32_t sStage = s0 + KV_BYTES;
  const uint32_t bars = sStage + NST * STAGE_BYTES;
  // barrier map (8 B each): full[NST] | odone[NST] | qrdy[4] | srdy[4] | prdy[4] | ordy[4] ; tmem ptr after
  const uint32_t b_full = bars, b_odone = bars + 8 * NST, b_qrdy = bars + 16 * NST, b_srdy = b_qrdy + 32,
                 b_prdy = b_qrdy + 64, b_ordy = b_qrdy + 96;
  const uint32_t b_beta = b_qrdy + 128;  // single-launch form: "the std of this call has been published" (grid barrier)
  const uint32_t b_kfull = b_qrdy + 136;  // pass 1: the K head group has landed (staged by TMA, see issue_k)
  const uint32_t b_prdy2 = b_qrdy + 144;  // pass 2: second part of P (keys 48..79) is in TMEM, one barrier per warpgroup
  static_assert(16 * NST + 128 + 16 + 32 <= 240, "barrier area");
  // Pass 1 stages K with the TMA itself: one box of 8 columns x 80 keys per (head, 16-byte chunk) IS the UMMA K-major
  // layout [chunk][key][16 B] (keys >= S arrive as zeros); 20 such boxes on one barrier stream at about a row per clock
  // (profiles/r1_tma_copy_rate.jsonl) with no thread work.  (Pass 2 cannot: V must be transposed and, with the compact
  // region map, the keys permuted.)  For pass 1 the tensor map of K travels in the tm_o slot.
  auto issue_k = [&](const Item& it) {
    const uint64_t pol_k = policy_evict_last();
    mbar_arrive_expect_tx(b_kfull, it.nheads * C::DCH * C::K_CH_BYTES);
    for (int h = 0; h < it.nheads; ++h)
#pragma unroll
      for (int c = 0; c < C::DCH; ++c)
        tma_load_3d(s0 + h * C::K_HEAD_BYTES + c * C::K_CH_BYTES, &tm_o, it.hg * C::GW + h * D + c * 8, 0, it.b, b_kfull, pol_k);
  };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + KV_BYTES + NST * STAGE_BYTES + 240);

  // Only the zero chunk that pads K's contraction dim (D = 40: the odd half k-step) has to be cleared: staging writes
  // every other byte the tensor core reads (pad keys as explicit zeros), and the V^T rows 42..47 it never writes only
  // feed the O columns 42..47 that nobody reads.
  if constexpr (C::KCH > C::DCH) {
    constexpr int ZCH = (C::KCH - C::DCH) * C::K_CH_BYTES / 16;  // uint4 per head
    for (int i = tid; i < C::G * ZCH; i += kX4Threads) {
      const int h = i / ZCH, r = i - h * ZCH;
      reinterpret_cast<uint4*>(smem + h * C::K_HEAD_BYTES + C::DCH * C::K_CH_BYTES)[r] = make_uint4(0, 0, 0, 0);
    }
  }
  const unsigned char* perm = nullptr;
  if constexpr (CW) {
    // key slot -> key: the n_active weighted columns first (ascending), then every other key in order
    unsigned char* pt = smem + X::CW_PERM_OFF;
    if (tid < DSC_MAX_KEYS) {
      int key = tid;
      if (tid < p.n_active) {
#pragma unroll
        for (int j = 0; j < DSC_MAX_COMPACT_COLS; ++j) key = (j == tid) ? p.active_cols[j] : key;
      } else {
        key = tid - p.n_active;
#pragma unroll
        for (int j = 0; j < DSC_MAX_COMPACT_COLS; ++j) key += (j < p.n_active && p.active_cols[j] <= key) ? 1 : 0;
      }
      pt[tid] = static_cast<unsigned char>(min(key, DSC_MAX_KEYS - 1));
    }
    perm = pt;  // visible to everybody after the start-up barrier below
  }
  TRACE(51);
  const int begin = static_cast<int>(p.total * blockIdx.x / gridDim.x);
  const int n_items = static_cast<int>(p.total * (blockIdx.x + 1) / gridDim.x) - begin;
  const uint64_t pol = STATS ? policy_evict_last() : policy_evict_first();
  // producer: TMA loads of tile i (Q boxes + W tile) into ring stage i % NST
  auto load_tile = [&](int i) {
    const int s = i % NST;
    const Item it = decode<D>(begin + i, p);
    const uint32_t sQ = sStage + s * STAGE_BYTES;
    uint32_t tx = C::QT_BYTES;
    const float* wsrc = nullptr;
    uint32_t wbytes = 0;
    if constexpr (CW) {
      tx += X::CWT_BYTES;
    } else if constexpr (!STATS) {
      if (w_fast) {
        tx += X::WT_BYTES;  // the whole box is counted, zero-filled parts included
      } else {
        wsrc = p.W + (static_cast<long long>(it.b / (p.B / p.Bw)) * p.L + it.l0) * p.w_pitch;
        wbytes = it.rows * p.w_pitch * 4;
        if (((reinterpret_cast<uintptr_t>(wsrc) | wbytes) & 15) == 0) tx += wbytes; else wbytes = 0;
      }
    }
    mbar_arrive_expect_tx(b_full + 8 * s, tx);
#pragma unroll
    for (int j = 0; j < C::NBOX; ++j)
      tma_load_3d(sQ + j * C::BOX_BYTES, &tm_q, it.hg * C::GW + j * C::BOX_COLS, it.l0, it.b, b_full + 8 * s, pol);
    if constexpr (CW) {
      tma_load_3d(sQ + C::QT_BYTES, &tm_w, 0, it.l0, it.b / (p.B / p.Bw), b_full + 8 * s, pol);  // 20 x 128 box of Wc
    } else if constexpr (!STATS) {
      if (w_fast) tma_load_3d(sQ + C::QT_BYTES, &tm_w, 0, it.l0, it.b / (p.B / p.Bw), b_full + 8 * s, pol);
      else if (wbytes != 0) bulk_g2s_hint(sQ + C::QT_BYTES, wsrc, wbytes, b_full + 8 * s, pol);
    }
    TRACE(40);
  };
  constexpr int kProducerTid = 19 * 32 + 16;
  if (tid == kProducerTid) {  // the producer owns the ring barriers and fills the ring right away
    for (int s = 0; s < NST; ++s) {
      mbar_init(b_full + 8 * s, 1);
      mbar_init(b_odone + 8 * s, warp_arrive ? STAGE_CONSUMERS / 32 : STAGE_CONSUMERS);
    }
    if constexpr (MODE == 2) mbar_init(b_beta, 1);
    fence_mbar_init();
    for (int i = 0; i < min(NST, n_items); ++i) load_tile(i);
  }
  if (tid == 0) {
    for (int g = 0; g < 4; ++g) {
      mbar_init(b_qrdy + 8 * g, warp_arrive ? 4 : 128);
      mbar_init(b_srdy + 8 * g, 1);
      mbar_init(b_prdy + 8 * g, warp_arrive ? 4 : 128);
      mbar_init(b_prdy2 + 8 * g, warp_arrive ? 4 : 128);
      mbar_init(b_ordy + 8 * g, 1);
    }
    if constexpr (STATS) mbar_init(b_kfull, 1);
    fence_mbar_init();
    if constexpr (STATS) {
      if (n_items > 0) {
        fence_proxy_async();  // the zero chunk cleared above (generic proxy) before the async-proxy writes next to it
        issue_k(decode<D>(begin, p));
      }
    }
  }
  if constexpr (MODE != 2) {
    if (warp == 16) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
                       smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)))
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  TRACE(52);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = MODE == 2 ? tmem_in : *tmem_ptr_smem;
  TRACE(2);

  if (warp >= 16) {
    const int wsvc = __shfl_sync(0xffffffffu, warp, 0) - 16;  // warp-uniform
    if (tid == kProducerTid) {
      // ============================== producer: TMA loads and stores ===============================
      if constexpr (MODE == 2) {
        // grid barrier of the single-launch form, waited for by ONE thread per CTA while everybody else already stages
        // K / V^T, Q rows and the first Q K^T of pass 2: the last CTA to publish its pass-1 partial bumps the epoch word
        volatile unsigned int* epoch =
            reinterpret_cast<volatile unsigned int*>(reinterpret_cast<unsigned char*>(p.ws) + kEpochOffsetT);
        unsigned int spins = 0;
        while (*epoch == epoch0) {
          __nanosleep(32);
          if (++spins > (1u << 25)) __trap();  // unreachable under a cooperative launch; never hang, never use a stale std
        }
        __threadfence();
        mbar_arrive(b_beta);
      }
      auto store_tile = [&](int i) {
        if constexpr (!STATS) {
          const Item it = decode<D>(begin + i, p);
          const uint32_t sQ = sStage + (i % NST) * STAGE_BYTES;
#pragma unroll
          for (int j = 0; j < C::NBOX; ++j)
            tma_store_3d(&tm_o, it.hg * C::GW + j * C::BOX_COLS, it.l0, it.b, sQ + j * C::BOX_BYTES);
          bulk_commit();
          bulk_wait_read0();
        }
      };
      for (int i = NST; i < n_items; ++i) {
        const int s = i % NST;
        mbar_wait_relaxed(b_odone + 8 * s, ((i / NST) - 1) & 1);
        TRACE(41);
        store_tile(i - NST);
        TRACE(42);
        load_tile(i);
      }
      for (int i = (n_items > NST ? n_items - NST : 0); i < n_items; ++i) {
        mbar_wait_relaxed(b_odone + 8 * (i % NST), (i / NST) & 1);
        store_tile(i);  // waits until the TMA has read the tile; the writes themselves complete before the grid does
      }
    } else if (lane == 0) {
      // ============================== MMA issuer of warpgroup g ====================================
      const int g = wsvc, h = g % HPT, par = g / HPT;
      constexpr uint32_t idesc_qk = idesc_f16<T>(80);
      constexpr uint32_t idesc_pv = idesc_f16<T>(48);
      const uint32_t tw = tmem_base + g * WG_COLS;
      const uint64_t kdesc = smem_desc(s0 + h * C::K_HEAD_BYTES, C::K_CH_BYTES, 128);
      uint32_t nq = 0, np = 0, np2 = 0;
      for (int i = 0; i < n_items; ++i) {
        if (PAR > 1 && (i % PAR) != par) continue;
        const Item it = decode<D>(begin + i, p);
        if (h >= it.nheads) continue;
        if (issuer_spin) mbar_wait(b_qrdy + 8 * g, nq & 1); else mbar_wait_relaxed(b_qrdy + 8 * g, nq & 1);
        const uint32_t qa = tw + O_COL + (QDB ? (nq & 1) * 24 : 0);
        ++nq;
        TRACE(21);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < C::KSTEPS; ++ks)  // descriptor start address advances by 2 chunks per k-step
          umma_ts(tw + S_COL, qa + ks * 8, kdesc + static_cast<uint64_t>((ks * 2 * C::K_CH_BYTES) >> 4), idesc_qk, ks);
        tc_commit(b_srdy + 8 * g);
        TRACE(22);
        if constexpr (!STATS) {
#pragma unroll
          for (int j = 0; j < NH; ++j) {  // O half j = P [V_half j | 1]
            const uint64_t vdesc = smem_desc(s0 + X::VT_OFF + (h * NH + j) * X::VH_BYTES, TC<40>::VT_CH_BYTES, 128);
            if (issuer_spin) mbar_wait(b_prdy + 8 * g, np & 1); else mbar_wait_relaxed(b_prdy + 8 * g, np & 1);
            ++np;
            TRACE(24);
            tc_fence_after();
            // P arrives in two parts (keys 0..47, then 48..79): the first three k-steps are multiplied while the consumers
            // still exponentiate the rest.  (The second V half of D = 80 finds all of P in place.)
#pragma unroll
            for (int kk = 0; kk < 3; ++kk)
              umma_ts(tw + O_COL, tw + S_COL + kk * 8, vdesc + static_cast<uint64_t>((kk * 2 * TC<40>::VT_CH_BYTES) >> 4), idesc_pv, kk);
            if (j == 0) {
              if (issuer_spin) mbar_wait(b_prdy2 + 8 * g, np2 & 1); else mbar_wait_relaxed(b_prdy2 + 8 * g, np2 & 1);
              ++np2;
              tc_fence_after();
            }
#pragma unroll
            for (int kk = 3; kk < 5; ++kk)
              umma_ts(tw + O_COL, tw + S_COL + kk * 8, vdesc + static_cast<uint64_t>((kk * 2 * TC<40>::VT_CH_BYTES) >> 4), idesc_pv, kk);
            tc_commit(b_ordy + 8 * g);
            TRACE(25);
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ============================== consumers: one head per warpgroup, one thread per query row ====
    const int g = warp >> 2, h = g % HPT, par = g / HPT;
    const int row = (warp & 3) * 32 + lane;
    auto arrive = [&](uint32_t bar) {  // whole (converged) warp
      if (warp_arrive) {
        __syncwarp();
        if (lane == 0) mbar_arrive(bar);
      } else {
        mbar_arrive(bar);
      }
    };
    const uint32_t row_off = row * 64, row_sw = (row >> 1) & 3;
    auto chunk_off = [&](int cg) -> uint32_t {
      return (cg >> 2) * C::BOX_BYTES + row_off + ((static_cast<uint32_t>(cg & 3) ^ row_sw) << 4);
    };
    const uint32_t tw = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + g * WG_COLS;
    float beta_l2 = 0.f;     // sigma * std * log2(e); read after pass 1 has completed (PDL), right before first use
    bool have_beta = STATS;
    const float scale_l2 = p.scale * kLog2eT;
    double dsum = 0.0, dsq = 0.0;
    uint32_t n_s = 0, n_o = 0, n_q = 0, n_run = 0;
    for (int r0 = 0; r0 < n_items;) {
      const Item it0 = decode<D>(begin + r0, p);
      const int r1 = min(n_items, r0 + p.n_sl - it0.tile);
      TRACE(3);
      if constexpr (STATS) {
        if (r0 > 0) {  // every warpgroup is done with the previous K: its successor may land
          asm volatile("bar.sync 1, 512;" ::: "memory");
          if (tid == 0) issue_k(it0);
        }
        MBAR_WAIT(b_kfull, n_run & 1, 4);
        ++n_run;
      } else {
        asm volatile("bar.sync 1, 512;" ::: "memory");
        stage_kv<T, D, true, kX4Consumers>(smem, p, it0, tid, perm);  // K of the head group
        stage_vt40<T, kX4Consumers>(smem + X::VT_OFF, p, it0, C::GW, it0.nheads * NH, tid, perm);
        asm volatile("bar.sync 1, 512;" ::: "memory");
      }
      TRACE(4);
      if (r1 < n_items) prefetch_kv<T, D, STATS, kX4Consumers>(p, decode<D>(begin + r1, p), tid);
      const bool active = h < it0.nheads;
      const int first = r0 + ((par - (r0 % PAR)) + PAR) % PAR;  // first tile of this warpgroup's parity in the run
      // Q row of head h of tile i -> TMEM (first D/2 (+4 zero) columns of the O region)
      constexpr int QWN = D == 40 ? 24 : 40;
      auto load_q = [&](int i, uint32_t (&qw)[QWN]) {  // Q row of head h of tile i: shared memory -> registers
        const int s = i % NST;
        const unsigned char* qtile = smem + KV_BYTES + s * STAGE_BYTES;
        MBAR_WAIT(b_full + 8 * s, (i / NST) & 1, 6);
        TRACE(5);
#pragma unroll
        for (int c = 0; c < C::DCH; ++c) {
          const uint4 v = *reinterpret_cast<const uint4*>(qtile + chunk_off(h * C::DCH + c));
          qw[4 * c] = v.x; qw[4 * c + 1] = v.y; qw[4 * c + 2] = v.z; qw[4 * c + 3] = v.w;
        }
      };
      auto put_q = [&](uint32_t (&qw)[QWN]) {  // registers -> tcgen05.st (not yet waited for)
        if constexpr (D == 40) {
          qw[20] = qw[21] = qw[22] = qw[23] = 0u;  // zero K padding of the odd half k-step
          const uint32_t qa = tw + O_COL + (QDB ? (n_q & 1) * 24 : 0);
          tmem_st_x16(qa, qw);
          tmem_st_x8(qa + 16, qw + 16);
        } else {
          tmem_st_x32(tw + O_COL, qw);
          tmem_st_x8(tw + O_COL + 32, qw + 32);
        }
        ++n_q;
      };
      auto publish_q = [&](int i) {  // operand rows (and, before them, this thread's S reads) are complete
        tc_wait_st();
        tc_fence_before();
        arrive(b_qrdy + 8 * g);
        TRACE(6);
        if constexpr (STATS) arrive(b_odone + 8 * (i % NST));  // pass 1 only reads Q
      };
      auto store_q = [&](int i) {
        uint32_t qw[QWN];
        load_q(i, qw);
        put_q(qw);
      };
      auto stage_q = [&](int i) {
        store_q(i);
        publish_q(i);
      };
      if (!active) {  // no head for this warpgroup in these tiles: just hand its tiles back
        for (int i = first; i < r1; i += PAR) {
          MBAR_WAIT(b_full + 8 * (i % NST), (i / NST) & 1, 6);
          if constexpr (!STATS) fence_proxy_async();
          arrive(b_odone + 8 * (i % NST));
        }
        r0 = r1;
        continue;
      }
      if (first < r1) stage_q(first);
      for (int i = first; i < r1; i += PAR) {
        const int s = i % NST;
        const int l0 = (it0.tile + (i - r0)) * C::ROWS;
        const int rows = min(C::ROWS, p.L - l0);
        unsigned char* qtile = smem + KV_BYTES + s * STAGE_BYTES;
        const bool has_next = i + PAR < r1;
        if constexpr (QDB) {
          if (has_next) store_q(i + PAR);  // into the other operand buffer, while this tile's Q K^T may still run
        }
        // ---- S row
        TRACE(10);
        MBAR_WAIT(b_srdy + 8 * g, n_s & 1, 8);
        ++n_s;
        TRACE(11);
        tc_fence_after();
        float sc[80];
        tmem_ld_x64(tw + S_COL, reinterpret_cast<uint32_t*>(sc));
        tmem_ld_x16(tw + S_COL + 64, reinterpret_cast<uint32_t*>(sc) + 64);
        tc_wait_ld();
        TRACE(12);
        if constexpr (STATS) {
          if (has_next) {  // S is in registers: the next tile's Q K^T may run under this tile's accumulation
            tc_fence_before();
            if constexpr (QDB) publish_q(i + PAR); else stage_q(i + PAR);
          }
          float fs[4] = {0.f, 0.f, 0.f, 0.f}, fq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < 80; j += 4) {  // element j -> accumulator j & 3, two fp32 lanes per instruction
            fadd2(fs[0], fs[1], fs[0], fs[1], sc[j], sc[j + 1]);
            fadd2(fs[2], fs[3], fs[2], fs[3], sc[j + 2], sc[j + 3]);
            ffma2(fq[0], fq[1], sc[j], sc[j + 1], sc[j], sc[j + 1], fq[0], fq[1]);
            ffma2(fq[2], fq[3], sc[j + 2], sc[j + 3], sc[j + 2], sc[j + 3], fq[2], fq[3]);
          }
          if (row < rows) {
            dsum += static_cast<double>((fs[0] + fs[1]) + (fs[2] + fs[3]));
            dsq += static_cast<double>((fq[0] + fq[1]) + (fq[2] + fq[3]));
          }
        } else {
          if (!have_beta) {
            if constexpr (MODE == 0) pdl_wait_prior_grid();  // pass 1 (same stream, launched just before) has published the std
            if constexpr (MODE == 2) mbar_wait(b_beta, 0);         // ... or, in the single-launch form, the grid barrier
            const float sigma = p.sigma_dev ? __ldcg(p.sigma_dev) : p.sigma_host;
            beta_l2 = sigma * __ldcg(&p.ws->std_unbiased) * kLog2eT;
            have_beta = true;
          }
          // logits in the log2 domain: s*scale*log2e + beta*log2e*W, W streamed from the shared tile
          uint32_t pw[40];
          if constexpr (CW) {
            // compact region map: the weighted key columns are slots 0..15 (keys permuted at staging), one 80-byte row
            // of W per query.  y = s * a + w * bw, 2^(e * y - e * max(y)): a = scale / beta, bw = 1, e = beta -- or, for a
            // vanishing beta, a = scale, bw = beta, e = 1
            const bool bpos = beta_l2 > 1e-20f;
            const float ca = bpos ? scale_l2 / beta_l2 : scale_l2, cbw = bpos ? 1.f : beta_l2, ce = bpos ? beta_l2 : 1.f;
            const float4* wt4 = reinterpret_cast<const float4*>(qtile + C::QT_BYTES + row * (X::CW_PITCH * 4));
            // slots 0..15: y formed explicitly; slots 16..79 carry no weight: their y is a * s, so the row max is taken on
            // the raw scores (a > 0) and the scaling folds into the single FFMA that forms the exponent
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float4 w = wt4[j];
              if (!bpos) {
                fmul2(w.x, w.y, w.x, w.y, cbw, cbw);
                fmul2(w.z, w.w, w.z, w.w, cbw, cbw);
              }
              ffma2(sc[4 * j], sc[4 * j + 1], sc[4 * j], sc[4 * j + 1], ca, ca, w.x, w.y);
              ffma2(sc[4 * j + 2], sc[4 * j + 3], sc[4 * j + 2], sc[4 * j + 3], ca, ca, w.z, w.w);
            }
            sc[77] = sc[78] = sc[79] = -INFINITY;  // pad keys (the launcher takes this path only for S == 77)
            float my[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, mr[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int j = 0; j < 16; ++j) my[j & 3] = fmaxf(my[j & 3], sc[j]);
#pragma unroll
            for (int j = 16; j < 80; ++j) mr[j & 3] = fmaxf(mr[j & 3], sc[j]);
            const float m = fmaxf(fmaxf(fmaxf(my[0], my[1]), fmaxf(my[2], my[3])),
                                  ca * fmaxf(fmaxf(mr[0], mr[1]), fmaxf(mr[2], mr[3])));
            const float nb = -ce * m, k2 = ce * ca;
#pragma unroll
            for (int j = 0; j < 39; ++j) {
              float e0, e1;
              if (j < 8) ffma2(e0, e1, sc[2 * j], sc[2 * j + 1], ce, ce, nb, nb);
              else ffma2(e0, e1, sc[2 * j], sc[2 * j + 1], k2, k2, nb, nb);
              pw[j] = j < 38 ? Mma<T>::pack(ex2_approx(e0), ex2_approx(e1)) : Mma<T>::pack(ex2_approx(e0), 0.f);
              if (j == 23) {  // keys 0..47 are done: publish them, the tensor core starts on P V
                TRACE(14);
                tmem_st_x16(tw + S_COL, pw);
                tmem_st_x8(tw + S_COL + 16, pw + 16);
                tc_wait_st();
                tc_fence_before();
                arrive(b_prdy + 8 * g);
              }
            }
            pw[39] = 0u;
            tmem_st_x16(tw + S_COL + 24, pw + 24);
            tc_wait_st();
            tc_fence_before();
            arrive(b_prdy2 + 8 * g);
            TRACE(16);
          } else if (w_fast && beta_l2 > 1e-20f) {
            // x = beta * y with y = s * (scale/beta) + W: the row max is taken on y, 2^(x - max) = 2^(beta*y - beta*ymax)
            const float4* wt4 = reinterpret_cast<const float4*>(qtile + C::QT_BYTES + row * (X::W_SMEM_PITCH * 4));
            const float cy = scale_l2 / beta_l2;
            float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int j = 0; j < 20; ++j) {
              const float4 w = wt4[j];
              ffma2(sc[4 * j], sc[4 * j + 1], sc[4 * j], sc[4 * j + 1], cy, cy, w.x, w.y);
              ffma2(sc[4 * j + 2], sc[4 * j + 3], sc[4 * j + 2], sc[4 * j + 3], cy, cy, w.z, w.w);
            }
            sc[77] = sc[78] = sc[79] = -INFINITY;  // pad keys
#pragma unroll
            for (int j = 0; j < 80; ++j) mx[j & 3] = fmaxf(mx[j & 3], sc[j]);
            const float nb = -beta_l2 * fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
            // 2^e for 77 keys.  Optional (-DDSC_POLY_PATTERN=0x52 ...): the selected key pairs go through the FMA pipe
            // instead of MUFU.EX2: e = j + f, j = round(e) taken from the mantissa of e + 1.5*2^23, 2^f on [-0.5, 0.5]
            // as a cubic (max. relative error 7.5e-5), exponent added with an integer shift-add.  Measured on B200 at
            // 2/8, 3/8, 4/8 of the pairs: no change of the kernel time (6/8: +2 us) -- the softmax phase is bound by
            // issue slots and latency, not by the 4/clk MUFU -- so the default keeps every exp on MUFU.
#pragma unroll
            for (int j = 0; j < 39; ++j) {
              float e0, e1;
              ffma2(e0, e1, sc[2 * j], sc[2 * j + 1], beta_l2, beta_l2, nb, nb);
              if (((DSC_POLY_PATTERN >> (j & 7)) & 1) && j < 38) {
                constexpr float kMagic = 12582912.f;
                e0 = fmaxf(e0, -126.f);
                e1 = fmaxf(e1, -126.f);
                float t0, t1, r0_, r1_, f0, f1, q0, q1;
                fadd2(t0, t1, e0, e1, kMagic, kMagic);
                fadd2(r0_, r1_, t0, t1, -kMagic, -kMagic);
                ffma2(f0, f1, r0_, r1_, -1.f, -1.f, e0, e1);
                ffma2(q0, q1, f0, f1, 0.0551716648f, 0.0551716648f, 0.2426111251f, 0.2426111251f);
                ffma2(q0, q1, q0, q1, f0, f1, 0.6932609677f, 0.6932609677f);
                ffma2(q0, q1, q0, q1, f0, f1, 0.9999280572f, 0.9999280572f);
                q0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(t0) << 23));
                q1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(t1) << 23));
                pw[j] = Mma<T>::pack(q0, q1);
              } else if (j < 38) {
                pw[j] = Mma<T>::pack(ex2_approx(e0), ex2_approx(e1));
              } else {
                pw[j] = Mma<T>::pack(ex2_approx(e0), 0.f);  // key 77 is a pad key
              }
            }
            pw[39] = 0u;  // pad keys 78, 79
          } else {
            const float* wsrc = p.W + ((static_cast<long long>(it0.b / (p.B / p.Bw)) * p.L + l0) * p.w_pitch);
            const bool bulk = w_fast || ((reinterpret_cast<uintptr_t>(wsrc) | static_cast<uintptr_t>(rows * p.w_pitch * 4)) & 15) == 0;
            const int wp = w_fast ? X::W_SMEM_PITCH : p.w_pitch;
            float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            if (bulk) {
              const float* wt = reinterpret_cast<const float*>(qtile + C::QT_BYTES) + row * wp;
              if (p.S == 77) {
#pragma unroll
                for (int j = 0; j < 77; ++j) {
                  sc[j] = fmaf(sc[j], scale_l2, wt[j] * beta_l2);
                  mx[j & 3] = fmaxf(mx[j & 3], sc[j]);
                }
                sc[77] = sc[78] = sc[79] = -INFINITY;
              } else {
#pragma unroll
                for (int j = 0; j < 80; ++j) {
                  sc[j] = (j < p.S) ? fmaf(sc[j], scale_l2, wt[j] * beta_l2) : -INFINITY;
                  mx[j & 3] = fmaxf(mx[j & 3], sc[j]);
                }
              }
            } else {
              const float* wr = wsrc + static_cast<long long>(row < rows ? row : 0) * p.w_pitch;
#pragma unroll
              for (int j = 0; j < 80; ++j) {
                sc[j] = (j < p.S) ? fmaf(sc[j], scale_l2, __ldg(wr + j) * beta_l2) : -INFINITY;
                mx[j & 3] = fmaxf(mx[j & 3], sc[j]);
              }
            }
            const float m = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
#pragma unroll
            for (int j = 0; j < 40; ++j) pw[j] = Mma<T>::pack(ex2_approx(sc[2 * j] - m), ex2_approx(sc[2 * j + 1] - m));
          }
          // (ex2.approx.f16x2 was tried to halve the MUFU work: on sm_100a it lowers to two MUFU.EX2.F16 plus
          //  repacking, i.e. more issue slots for the same MUFU count -- measured slower)
          if constexpr (!CW) {
            TRACE(14);
            tmem_st_x32(tw + S_COL, pw);  // P over the first 40 columns of S (the whole S row is in registers)
            tmem_st_x8(tw + S_COL + 32, pw + 32);
            tc_wait_st();
            tc_fence_before();
            arrive(b_prdy + 8 * g);
            arrive(b_prdy2 + 8 * g);
            TRACE(16);
          }
          auto o_chunk = [&](const float* oc8, float inv_, int chunk) {  // 8 O columns -> 16 bytes of the row in smem
            float t[8];
            fmul2(t[0], t[1], oc8[0], oc8[1], inv_, inv_);
            fmul2(t[2], t[3], oc8[2], oc8[3], inv_, inv_);
            fmul2(t[4], t[5], oc8[4], oc8[5], inv_, inv_);
            fmul2(t[6], t[7], oc8[6], oc8[7], inv_, inv_);
            uint4 v;
            v.x = Mma<T>::pack(t[0], t[1]);
            v.y = Mma<T>::pack(t[2], t[3]);
            v.z = Mma<T>::pack(t[4], t[5]);
            v.w = Mma<T>::pack(t[6], t[7]);
            *reinterpret_cast<uint4*>(qtile + chunk_off(chunk)) = v;
          };
          if constexpr (PAR == 1) {
            // D = 40.  The next tile's Q row leaves shared memory while the tensor core multiplies P V (the scores and P
            // are dead: registers are free).  O is read in pieces: columns 0..23 (where the Q operand lives) and the ones
            // column first, then the Q row goes to TMEM -- its Q K^T starts -- and only then columns 24..39.
            uint32_t qnext[QWN];
            load_q(has_next ? i + 1 : i, qnext);  // (unconditional: this tile's own stage when there is no next tile)
            MBAR_WAIT(b_ordy + 8 * g, n_o & 1, 5);
            ++n_o;
            TRACE(17);
            tc_fence_after();
            float oa[24], oz[4];
            tmem_ld_x16(tw + O_COL, reinterpret_cast<uint32_t*>(oa));
            tmem_ld_x8(tw + O_COL + 16, reinterpret_cast<uint32_t*>(oa) + 16);
            tmem_ld_x4(tw + O_COL + 40, reinterpret_cast<uint32_t*>(oz));
            tc_wait_ld();
            const float inv1 = 1.f / oz[0];  // the ones column: softmax row sum of the rounded P
            tc_fence_before();
            if (has_next) {
              put_q(qnext);
              publish_q(i + 1);
            }
            o_chunk(oa, inv1, h * C::DCH);
            o_chunk(oa + 8, inv1, h * C::DCH + 1);
            o_chunk(oa + 16, inv1, h * C::DCH + 2);
            float ob[16];
            tmem_ld_x16(tw + O_COL + 24, reinterpret_cast<uint32_t*>(ob));
            tc_wait_ld();
            tc_fence_before();
            o_chunk(ob, inv1, h * C::DCH + 3);
            o_chunk(ob + 8, inv1, h * C::DCH + 4);
          } else {
          float inv = 0.f;
#pragma unroll
          for (int j = 0; j < NH; ++j) {
            MBAR_WAIT(b_ordy + 8 * g, n_o & 1, 5);
            ++n_o;
            TRACE(17);
            tc_fence_after();
            float o[48];
            tmem_ld_x32(tw + O_COL, reinterpret_cast<uint32_t*>(o));
            tmem_ld_x16(tw + O_COL + 32, reinterpret_cast<uint32_t*>(o) + 32);
            tc_wait_ld();
            if (j == 0) inv = 1.f / o[40];  // the ones column: softmax row sum of the rounded P
            tc_fence_before();
            if (j + 1 < NH) {
              arrive(b_prdy + 8 * g);  // the O columns are free: the second V half may be multiplied
            } else if (PAR == 1 && has_next) {
              stage_q(i + PAR);  // ... or the next tile's Q goes to TMEM now; its Q K^T runs under the O store
            }
#pragma unroll
            for (int c = 0; c < 5; ++c) {
              uint4 v;
              float* oc = o + 8 * c;
              fmul2(oc[0], oc[1], oc[0], oc[1], inv, inv);
              fmul2(oc[2], oc[3], oc[2], oc[3], inv, inv);
              fmul2(oc[4], oc[5], oc[4], oc[5], inv, inv);
              fmul2(oc[6], oc[7], oc[6], oc[7], inv, inv);
              v.x = Mma<T>::pack(oc[0], oc[1]);
              v.y = Mma<T>::pack(oc[2], oc[3]);
              v.z = Mma<T>::pack(oc[4], oc[5]);
              v.w = Mma<T>::pack(oc[6], oc[7]);
              *reinterpret_cast<uint4*>(qtile + chunk_off(h * C::DCH + j * 5 + c)) = v;
            }
          }
          }
          TRACE(18);
          fence_proxy_async();
          arrive(b_odone + 8 * s);
          if (PAR > 1 && has_next) stage_q(i + PAR);  // same ring stage: only after it has been handed back and refilled
        }
      }
      r0 = r1;
    }
    if constexpr (STATS) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
        dsq += __shfl_xor_sync(0xffffffffu, dsq, o);
      }
      double* red = reinterpret_cast<double*>(smem);
      asm volatile("bar.sync 1, 512;" ::: "memory");
      if (lane == 0) {
        red[warp] = dsum;
        red[16 + warp] = dsq;
      }
      asm volatile("bar.sync 1, 512;" ::: "memory");
      double* partials = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(p.ws) + kWorkspaceHeader);
      if (warp == 0) {
        unsigned int last = 0;
        if (lane == 0) {
          double a = 0.0, b = 0.0;
          for (int w = 0; w < 16; ++w) {
            a += red[w];
            b += red[16 + w];
          }
          partials[2 * blockIdx.x] = a;
          partials[2 * blockIdx.x + 1] = b;
          __threadfence();
          last = atomicAdd(&p.ws->ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {  // last CTA: fold all partials (deterministic order)
          __threadfence();
          finalize_stats(p, partials, lane);
          if constexpr (MODE == 1) {  // single-launch form: the publication is the grid barrier -- bump the epoch word
            if (lane == 0) {
              __threadfence();
              atomicAdd(reinterpret_cast<unsigned int*>(reinterpret_cast<unsigned char*>(p.ws) + kEpochOffsetT), 1u);
            }
          }
        }
      }
    }
  }

  TRACE(9);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  TRACE(30);
  CTA_TIME(1);
  if constexpr (MODE == 1) {
    if (tid == 0)  // this phase's barrier words become ring-stage bytes of the next phase
      for (int i = 0; i < 2 * NST + 22; ++i) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(bars + 8 * i) : "memory");
  } else {
    if (warp == 16) {
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
  }
  return tmem_base;
}

template <typename T, int D, bool STATS>
__global__ void __launch_bounds__(kX4Threads, 1)
xattn_tc5x4_kernel(const XattnParams p, const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_o,
                   const __grid_constant__ CUtensorMap tm_w) {
  x4_phase<T, D, STATS, 0>(p, tm_q, tm_o, tm_w, 0u);
}

// pass 2 with the compact region map (3-stage ring, 80-byte W rows, keys permuted)
template <typename T, int D>
__global__ void __launch_bounds__(kX4Threads, 1)
xattn_tc5x4_cw_kernel(const XattnParams p, const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_o,
                      const __grid_constant__ CUtensorMap tm_w) {
  x4_phase<T, D, false, 0, true>(p, tm_q, tm_o, tm_w, 0u);
}

// Both passes in ONE cooperative launch (grid <= SM count, one CTA per SM): pass 1 over the CTA's tile range, grid
// barrier (the last CTA to publish its partial folds them all, writes the std and bumps an epoch word), pass 2 over the
// same range.  K stays in place conceptually (it is restaged together with V^T while the slower CTAs still arrive), Q
// is read again while much of it is still in L2 (pass 1 loads it evict-last), and there is no second launch, no second
// TMEM allocation and no pass-2 prologue behind the end of pass 1.
template <typename T, int D>
__global__ void __launch_bounds__(kX4Threads, 1)
xattn_tc5x4_fused_kernel(const XattnParams p, const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_o,
                         const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_k) {
  volatile unsigned int* epoch = reinterpret_cast<volatile unsigned int*>(reinterpret_cast<unsigned char*>(p.ws) + kEpochOffsetT);
  unsigned int e0 = 0;
  if (threadIdx.x == 19 * 32 + 16) e0 = *epoch;  // the producer thread, which waits for the bump in pass 2; read before
                                                 // this CTA has arrived at the barrier: the bump cannot have happened yet
  const uint32_t tmem_base = x4_phase<T, D, true, 1>(p, tm_q, tm_k, tm_q, 0u);
  x4_phase<T, D, false, 2>(p, tm_q, tm_o, tm_w, tmem_base, e0);
}



template <typename T, int D, bool STATS>
static cudaError_t launch_tc5(XattnParams p, cudaStream_t st) {
  using C = TC<D>;
  constexpr int smem = STATS ? C::STATS_SMEM : C::FWD_SMEM;
  static thread_local int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(xattn_tc5_kernel<T, D, STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured_dev = dev;
  }
  CUtensorMap tm_q, tm_o;
  if (!make_map(&tm_q, p.q, p.H * D, p.L, p.B, p.q_sl, p.q_sb)) return cudaErrorInvalidValue;
  if (STATS) tm_o = tm_q;
  else if (!make_map(&tm_o, p.out, p.H * D, p.L, p.B, p.o_sl, p.o_sb)) return cudaErrorInvalidValue;
  // re-partition for 128-row tiles and 160-column head groups
  p.n_hg = (p.H + C::G - 1) / C::G;
  p.n_sl = (p.L + C::ROWS - 1) / C::ROWS;
  p.total = static_cast<long long>(p.B) * p.n_hg * p.n_sl;
  if (p.total >= (1ll << 31)) return cudaErrorInvalidValue;
  const int sms = sm_count_cached();
  const int grid = static_cast<int>(p.total < sms ? p.total : sms);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (!STATS && !config().no_pdl) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, xattn_tc5_kernel<T, D, STATS>, p, tm_q, tm_o);
}

#ifdef DSC_WATCHDOG
extern "C" int dsc_debug_watchdog(unsigned int* out /*HOST 3*/) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_wd_info, 8);
  cudaMemcpyFromSymbol(out + 2, g_wd_abort, 4);
  return 0;
}
#endif

#ifdef DSC_TRACE
extern "C" int dsc_debug_trace(long long* out /*HOST 4*512*2*/, int* counts /*HOST 4*/) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_trace, sizeof(long long) * 4 * 512 * 2);
  cudaMemcpyFromSymbol(counts, g_trace_n, sizeof(int) * 4);
  { static unsigned long long ct[160][2]; cudaMemcpyFromSymbol(ct, g_cta_times, sizeof(ct));
    unsigned long long t0 = ~0ull, t1 = 0; for (int i = 0; i < 148; ++i) { if (ct[i][0] && ct[i][0] < t0) t0 = ct[i][0]; if (ct[i][1] > t1) t1 = ct[i][1]; }
    printf("cta times (ns, last kernel): span %llu;", t1 - t0);
    unsigned long long smin = ~0ull, smax = 0, dmin = ~0ull, dmax = 0;
    for (int i = 0; i < 148; ++i) { unsigned long long st = ct[i][0] - t0, d = ct[i][1] - ct[i][0]; if (st < smin) smin = st; if (st > smax) smax = st; if (d < dmin) dmin = d; if (d > dmax) dmax = d; }
    printf(" start offsets %llu..%llu; durations %llu..%llu\n", smin, smax, dmin, dmax);
    printf("cta durations ns:"); for (int i = 0; i < 148; ++i) printf(" %llu", ct[i][1] - ct[i][0]); printf("\n"); }
  { long long kv[8]; cudaMemcpyFromSymbol(kv, g_kv_trace, sizeof(kv));
    printf("kv staging (block 0, thread 0, last call): enter->loads_issued %lld, ->k_stored %lld, ->v_stored %lld, ->fenced %lld cycles\n",
           kv[1] - kv[0], kv[2] - kv[0], kv[3] - kv[0], kv[4] - kv[0]); }
  int z[4] = {0, 0, 0, 0};
  cudaMemcpyToSymbol(g_trace_n, z, sizeof(z));
  return 0;
}
#endif

template <typename T, int D>
static cudaError_t launch_tc5x4_fused(XattnParams p, cudaStream_t st) {
  using C = TC<D>;
  constexpr int smem = X4<D>::FWD_SMEM > X4<D>::STATS_SMEM ? X4<D>::FWD_SMEM : X4<D>::STATS_SMEM;
  static thread_local int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(xattn_tc5x4_fused_kernel<T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured_dev = dev;
  }
  CUtensorMap tm_q, tm_o;
  if (!make_map(&tm_q, p.q, p.H * D, p.L, p.B, p.q_sl, p.q_sb)) return cudaErrorInvalidValue;
  if (!make_map(&tm_o, p.out, p.H * D, p.L, p.B, p.o_sl, p.o_sb)) return cudaErrorInvalidValue;
  p.n_hg = (p.H + C::G - 1) / C::G;
  p.n_sl = (p.L + C::ROWS - 1) / C::ROWS;
  p.total = static_cast<long long>(p.B) * p.n_hg * p.n_sl;
  if (p.total >= (1ll << 31)) return cudaErrorInvalidValue;
  const int sms = sm_count_cached();
  const int grid = static_cast<int>(p.total < sms ? p.total : sms);
  const unsigned env_flags = config().tc5_flags;
  p.flags = env_flags & 3u;
  CUtensorMap tm_w = tm_q;
  if (!(env_flags & 8u) && p.w_pitch == DSC_MAX_KEYS && p.S == 77 && (reinterpret_cast<uintptr_t>(p.W) & 15) == 0) {
    if (!make_map_w(&tm_w, p.W, p.L, p.Bw)) return cudaErrorInvalidValue;
    p.flags |= 4u;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kX4Threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;  // the grid barrier between the passes needs every CTA resident
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CUtensorMap tm_k;
  if (!make_map_k(&tm_k, p.k, p.H * D, p.S, p.B, p.k_ss, p.k_sb)) return cudaErrorInvalidValue;
  return cudaLaunchKernelEx(&cfg, xattn_tc5x4_fused_kernel<T, D>, p, tm_q, tm_o, tm_w, tm_k);
}

template <typename T, int D, bool STATS>
static cudaError_t launch_tc5x4(XattnParams p, cudaStream_t st) {
  using C = TC<D>;
  constexpr int smem = STATS ? X4<D>::STATS_SMEM : X4<D>::FWD_SMEM;
  static thread_local int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(xattn_tc5x4_kernel<T, D, STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured_dev = dev;
  }
  CUtensorMap tm_q, tm_o;
  if (!make_map(&tm_q, p.q, p.H * D, p.L, p.B, p.q_sl, p.q_sb)) return cudaErrorInvalidValue;
  if (STATS) {
    if (!make_map_k(&tm_o, p.k, p.H * D, p.S, p.B, p.k_ss, p.k_sb)) return cudaErrorInvalidValue;  // pass 1: K's map
  } else if (!make_map(&tm_o, p.out, p.H * D, p.L, p.B, p.o_sl, p.o_sb)) return cudaErrorInvalidValue;
  p.n_hg = (p.H + C::G - 1) / C::G;
  p.n_sl = (p.L + C::ROWS - 1) / C::ROWS;
  p.total = static_cast<long long>(p.B) * p.n_hg * p.n_sl;
  if (p.total >= (1ll << 31)) return cudaErrorInvalidValue;
  const int sms = sm_count_cached();
  const int grid = static_cast<int>(p.total < sms ? p.total : sms);
  // experiment knobs (A/B runs): bit 0 one mbarrier arrival per warp, bit 1 MMA issuers spin instead of
  // nanosleep-polling, bit 3 do not use the padded-W fast path
  const unsigned env_flags = config().tc5_flags;
  p.flags = env_flags & 3u;
  CUtensorMap tm_w = tm_q;
  bool compact = false;
  if constexpr (!STATS) {
    compact = p.wc != nullptr && p.n_active > 0 && p.S == 77 && p.scale > 0.f && !(env_flags & 16u);  // bit 4: ignore the compact map
    if (compact) {
      if (!make_map_wc(&tm_w, p.wc, p.L, p.Bw)) return cudaErrorInvalidValue;
    } else if (!(env_flags & 8u) && p.w_pitch == DSC_MAX_KEYS && p.S == 77 && (reinterpret_cast<uintptr_t>(p.W) & 15) == 0) {
      if (!make_map_w(&tm_w, p.W, p.L, p.Bw)) return cudaErrorInvalidValue;
      p.flags |= 4u;  // W tiles arrive as TMA boxes at a pitch of 84 floats
    }
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kX4Threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = config().no_pdl ? 0 : 1;  // pass 2 may overlap the tail of pass 1 (and pass 1 its predecessor's)
  if constexpr (!STATS) {
    if (compact) {
      static thread_local int cw_dev = -1;
      if (cw_dev != dev) {
        cudaError_t e = cudaFuncSetAttribute(xattn_tc5x4_cw_kernel<T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, X4<D>::CW_SMEM);
        if (e != cudaSuccess) return e;
        cw_dev = dev;
      }
      cfg.dynamicSmemBytes = X4<D>::CW_SMEM;
      return cudaLaunchKernelEx(&cfg, xattn_tc5x4_cw_kernel<T, D>, p, tm_q, tm_o, tm_w);
    }
  }
  return cudaLaunchKernelEx(&cfg, xattn_tc5x4_kernel<T, D, STATS>, p, tm_q, tm_o, tm_w);
}

// Two tcgen05 variants: "x4" (4 consumer warpgroups, one head each; default) and "x2" (2 warpgroups, software-
// pipelined heads).  DSC_TC5_VARIANT=x2 selects the latter for A/B runs.
static bool use_x4(int D) {
  if (D != 40 && D != 80) return false;
  return !config().tc5_x2;
}

bool tc5_supports(int D) { return D == 40 || D == 80; }

cudaError_t run_stats_tc5(const XattnParams& p, int D, int dtype, cudaStream_t st) {
  if (use_x4(D)) {
    if (dtype == DSC_DTYPE_F16) return D == 40 ? launch_tc5x4<__half, 40, true>(p, st) : launch_tc5x4<__half, 80, true>(p, st);
    return D == 40 ? launch_tc5x4<__nv_bfloat16, 40, true>(p, st) : launch_tc5x4<__nv_bfloat16, 80, true>(p, st);
  }
  if (dtype == DSC_DTYPE_F16)
    return D == 40 ? launch_tc5<__half, 40, true>(p, st) : launch_tc5<__half, 80, true>(p, st);
  return D == 40 ? launch_tc5<__nv_bfloat16, 40, true>(p, st) : launch_tc5<__nv_bfloat16, 80, true>(p, st);
}

// both passes in one cooperative launch (4-warpgroup kernels, D = 40 / 80)
bool tc5_fused_supports(int D) { return use_x4(D); }
cudaError_t run_fused_tc5(const XattnParams& p, int D, int dtype, cudaStream_t st) {
  if (!use_x4(D)) return cudaErrorInvalidValue;
  if (dtype == DSC_DTYPE_F16) return D == 40 ? launch_tc5x4_fused<__half, 40>(p, st) : launch_tc5x4_fused<__half, 80>(p, st);
  return D == 40 ? launch_tc5x4_fused<__nv_bfloat16, 40>(p, st) : launch_tc5x4_fused<__nv_bfloat16, 80>(p, st);
}

cudaError_t run_forward_tc5(const XattnParams& p, int D, int dtype, cudaStream_t st) {
  if (use_x4(D)) {
    if (dtype == DSC_DTYPE_F16) return D == 40 ? launch_tc5x4<__half, 40, false>(p, st) : launch_tc5x4<__half, 80, false>(p, st);
    return D == 40 ? launch_tc5x4<__nv_bfloat16, 40, false>(p, st) : launch_tc5x4<__nv_bfloat16, 80, false>(p, st);
  }
  if (dtype == DSC_DTYPE_F16)
    return D == 40 ? launch_tc5<__half, 40, false>(p, st) : launch_tc5<__half, 80, false>(p, st);
  return D == 40 ? launch_tc5<__nv_bfloat16, 40, false>(p, st) : launch_tc5<__nv_bfloat16, 80, false>(p, st);
}

}  // namespace dsc
