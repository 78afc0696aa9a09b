// Region-masked cross-attention on a prepared K / V^T image: warp-specialised tcgen05 + TMEM + TMA kernels ("x3"),
// head dims 40 / 80 / 160, S = 77.  What RegionAttnProcessor runs for every SD-1.5 cross-attention layer.
//
// Same two passes as xattn_tc5.cu (pass 1: std of scale*QK^T over the whole call; pass 2: softmax(scale*QK^T + beta*W) V;
// reference source/modules/attention_modify.py:74-103 with the weight_func of source/app.py:1004), organised so that no
// thread ever waits for a tensor-core round trip it has just started, and -- for the whole call -- as ONE cooperative
// launch (xattn_x3_fused_kernel: x3_phase<pass 1> then x3_phase<pass 2> in the same persistent CTAs):
//
//   * Work item = (128-row tile, head).  Items of a CTA's tile range are dealt round-robin to 3 / 2 / 1 consumer
//     warpgroups (one thread per query row; TMEM lane = row).  A warpgroup owns S (80 fp32 columns) | P (40, 16-bit pairs)
//     | O (48 / 96 / 176 fp32) as SEPARATE TMEM regions (3 x 168 = 504 <= 512 columns at D = 40), so
//         Q K^T of item n+1 is issued the moment S(n) has been read into registers,
//         P V   of item n   runs while the warpgroup already exponentiates item n+1.
//     The x4 kernel (xattn_tc5.cu) aliases P on S and the Q operand on O: its four warpgroups sit through the serial chain
//     O -> Q rows to TMEM -> QK^T -> S -> softmax -> P -> PV together (2.8k of 5.7k cycles per tile idle, profiles/r1_*).
//   * The consumers never touch O: a DRAIN warpgroup (pass 2; warp d <-> TMEM lanes 32d..32d+31 of every consumer
//     warpgroup) takes the O rows out of TMEM, scales them by 1 / rowsum (ones row of V^T), and stores them over the rows'
//     own Q columns of the ring stage; the register file is re-divided with setmaxnreg (consumers up, drain / service down).
//   * Q is the A operand STRAIGHT FROM SHARED MEMORY (no smem -> registers -> TMEM hop, no consumer thread involved).
//     The TMA engine costs about one cycle per box ROW whatever its width (profiles/r1_tma_copy_rate.jsonl), so the
//     160-column tile arrives as just three boxes: columns [0,64) and [64,128) with SWIZZLE_128B and [128,160) with
//     SWIZZLE_64B -- the UMMA K-major canonical layouts.  A k16 step must lie inside one swizzle row, and head h starts
//     at column 40h, so the contraction of head h runs over the three 16-column blocks (32-byte aligned in the row) that
//     cover its 40 columns, and the K image holds ZEROS where a block's columns belong to a neighbouring head (h = 0:
//     blocks 0-2, h = 1: 2-4, h = 2: 5-7, h = 3: 7-9): 3 MMAs per head, as many as 40 columns need anyway.  Finished O
//     rows leave through the same three box shapes; the compact W tile (128 rows x 80 B, contiguous) is ONE bulk copy.
//   * K and V^T|1 of a (batch, head group) arrive as bulk copies of head RECORDS of the prepared image
//     (dsc_xattn_prepare_kv: UMMA K-major chunks, keys permuted so that the weighted columns of the compact region map are
//     slots 0..15, ones row appended to V^T so that column D of O is the softmax row sum) into a ring of record slots
//     (5 for 4 heads at D = 40: the next group's first record loads while all heads of the old one are in use).  K / V never
//     change during a generation (attention_modify.py:465-466 recomputes the same projections on each of the 25 steps).
//   * Producer warp and tensor-core issuer warps run CONVERGED: every TMA / MMA operand is warp-uniform to the compiler and
//     lives in uniform registers; the instruction is issued by the elected lane.
//   * pass 1 (STATS): same skeleton without the drain warpgroup, S double-buffered in TMEM (2 x 80 columns per
//     warpgroup), 4-stage Q ring, every thread accumulates sum / sum of squares of its S rows; the CTA's partial goes
//     to the workspace: ticket + deterministic fold by the last CTA (pass 1 on its own), or -- whole call -- two 8-byte
//     stores into the HANDOFF slots that drain warp 0 of every pass-2 CTA polls and folds in the same fixed order.
// Launch forms: the whole call = one cooperative launch (fallback: two launches, pass 2 a programmatic dependent launch of
// pass 1 that never waits for pass 1's completion); pass 1 / pass 2 on their own = xattn_x3_kernel<.., STATS>.
// Debug builds: -DDSC_TRACE (clock64 timelines), -DDSC_PHASE (cycles per consumer phase), -DDSC_CTATIME (globaltimer span
// of every CTA and pass), -DDSC_WATCHDOG (who waited on what); A/B switches X3_TURNS, X3_POLY, X3_NSLOT40.
#include "tc5_common.cuh"
#include "tc5_tmem.cuh"

#include <stdio.h>
#include <stdlib.h>

namespace dsc {

#ifndef X3_LATE_PVWAIT
#define X3_LATE_PVWAIT 0  // pass 2: wait for P V of the previous item right before the first P store instead of before the exponentials (A/B: no change, 37.7 us)
#endif
#ifndef X3_TURNS
#define X3_TURNS 2      // pass 2, 3 warpgroups: at most this many of the three warps that share an SM sub-partition exponentiate at a time (0 = off)
#endif
#ifndef X3_L2_PREFETCH
#define X3_L2_PREFETCH 1  // tile i + NST is pulled towards L2 when tile i is loaded
#endif

#ifndef X3_POLY
#define X3_POLY 0  // pass 2: every X3_POLY-th pair of exponentials on the FMA / ALU pipes (0 = all on the MUFU pipe; measured slower: see DESIGN.md)
#endif

namespace x3 {
constexpr int X3_TURNS_DEFAULT = X3_TURNS;
constexpr float kLog2e = 1.4426950408889634f;
constexpr int GW = 160, ROWS = 128;  // a tile: 128 query rows x one 160-column head group (4 / 2 / 1 heads of 40 / 80 / 160)
// ring stage: Q / O tile = boxes [0,64) | [64,128) (SW128, 16 KB each) | [128,160) (SW64, 8 KB), then the compact W tile
constexpr int BOX128_BYTES = ROWS * 128, BOX64_BYTES = ROWS * 64;
constexpr int QT_BYTES = 2 * BOX128_BYTES + BOX64_BYTES;               // 40960
constexpr int CW_BYTES = ROWS * DSC_COMPACT_PITCH * 4;                 // 10240
constexpr int FWD_STAGE = QT_BYTES + CW_BYTES;                         // 51200
constexpr int STATS_STAGE = QT_BYTES, STATS_NST = 4;
constexpr int BAR_BYTES = 512;
constexpr int S_COL = 0, P_COL = 80, O_COL = 120, S1_COL = 80;         // TMEM columns of one warpgroup: S | P | O (pass 1: S | S)
constexpr int K_CH = DSC_MAX_KEYS * 16;                                // one 8-column chunk of K: 80 key slots x 16 B
constexpr int round_1k(int x) { return (x + 1023) / 1024 * 1024; }

// Per head dim.  Head h of a group starts at column HD * h; its contraction runs over the KSTEPS 16-column blocks that cover
// it (HD = 40: 3 blocks, partly shared with a neighbour -- the K image is zero there; HD = 80 / 160: 5 / 10 whole blocks).
// P V has N = ON columns: the head's HD value columns, the ones column (softmax row sum), zero padding to a multiple of 16.
// A warpgroup owns 120 + ON TMEM columns, so 3 / 2 / 1 warpgroups fit the 512 columns.
template <int HD>
struct Cfg {
  static_assert(HD == 40 || HD == 80 || HD == 160, "head dim");
  static constexpr int D = HD, HPT = GW / HD, LOG_HPT = HD == 40 ? 2 : HD == 80 ? 1 : 0;
  static constexpr int NWG = HD == 40 ? 3 : HD == 80 ? 2 : 1;
  // pass 1: consumer warpgroups + 4 service warps.  pass 2: consumer warpgroups + one DRAIN warpgroup (O rows: TMEM -> x 1/rowsum
  // -> ring stage; warp d serves TMEM lanes 32d..32d+31 of every consumer warpgroup) + 4 service warps; the register file is
  // re-divided with setmaxnreg (consumers up, drain / service warps down)
  static constexpr int CONSUMERS = NWG * 128;
  static constexpr int threads(bool stats) { return CONSUMERS + (stats ? 128 : 256); }
  static constexpr int REGS_CONSUMER = HD == 40 ? 128 : HD == 80 ? 184 : 232;  // x CONSUMERS
  static constexpr int REGS_DRAIN = HD == 40 ? 56 : HD == 80 ? 80 : 96;        // x 128
  static constexpr int REGS_SERVICE = HD == 40 ? 40 : 56;                      // x 128
  static constexpr int REGS_LAUNCH = (65536 / threads(false)) / 8 * 8;         // what __launch_bounds__(threads, 1) grants: 96 / 128 / 168
  static_assert(CONSUMERS * REGS_CONSUMER + 128 * (REGS_DRAIN + REGS_SERVICE) <= threads(false) * REGS_LAUNCH, "register pool");
  static constexpr int KSTEPS = HD == 40 ? 3 : HD / 16, NKC = 2 * KSTEPS;
  static constexpr int ON = HD == 40 ? 48 : HD == 80 ? 96 : 176;
  static constexpr int WG_COLS = O_COL + ON;
  static constexpr int CPH = HD / 8;                                   // 16-byte chunks of a head's O row
  // prepared image of one (batch, head group) = HPT head RECORDS [K_h | V^T_h]; shared memory holds NSLOT records, used as a
  // ring in record order (run R, head h) -> record R * HPT + h -> slot record % NSLOT: a head's slot is refilled with the next
  // (batch, head group)'s record as soon as the last item that reads it has finished, while the other heads of the old
  // group are still being worked on (HD = 160: one head per group, so two whole records alternate and the ring of Q
  // stages shrinks to 2)
  static constexpr int K_HEAD = NKC * K_CH;
  static constexpr int VT_CH = ON * 16, VT_HEAD = 10 * VT_CH;
  static constexpr int REC_BYTES = K_HEAD + VT_HEAD;                   // 15360 / 28160 / 53760
  static constexpr int IMG_BYTES = HPT * REC_BYTES;                    // 61440 / 56320 / 53760
#ifndef X3_NSLOT40
#define X3_NSLOT40 5
#endif
  // HD = 40: one slot more than a (batch, head group) has heads -- at a run boundary the first record of the next run is
  // loaded while all four heads of the old run are still in use (a CTA whose tiles cross a boundary pays ~1.4 us for the
  // drained pipeline with 4 slots: profiles/r2_x3_span_detail*.txt)
  static constexpr int NSLOT = HD == 40 ? X3_NSLOT40 : 2;
  static constexpr int FWD_NST = HD == 160 ? 2 : 3;
  static constexpr int FWD_SMEM = round_1k(NSLOT * REC_BYTES) + FWD_NST * FWD_STAGE + BAR_BYTES;
  static constexpr int STATS_SMEM = round_1k(NSLOT * K_HEAD) + STATS_NST * STATS_STAGE + BAR_BYTES;
  static constexpr int TURNS = NWG == 3 ? X3_TURNS_DEFAULT : 0;
  static_assert(FWD_STAGE % 1024 == 0 && STATS_STAGE % 1024 == 0 &&
                    REC_BYTES % 16 == 0 && K_HEAD % 16 == 0, "alignment");
  static_assert(FWD_SMEM <= 227 * 1024 && STATS_SMEM <= 227 * 1024, "shared memory budget");
  static_assert(NWG * WG_COLS <= 512 && 2 * S1_COL * NWG <= 512, "TMEM budget");
};

struct Tile {
  int b, hg, tile, l0;
};
__device__ __forceinline__ Tile decode(int idx, const XattnParams& p) {
  Tile t;
  const int seg = static_cast<int>(fastdiv(static_cast<uint32_t>(idx), p.div_nsl));
  t.tile = idx - seg * p.n_sl;
  t.b = static_cast<int>(fastdiv(static_cast<uint32_t>(seg), p.div_nhg));
  t.hg = seg - t.b * p.n_hg;
  t.l0 = t.tile * ROWS;
  return t;
}

// D[tmem] (+)= A[smem desc] * B[smem desc]^T
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major SWIZZLE_64B descriptor: rows 64 B apart, 8-row groups 512 B apart (LBO unused: both chunks of a k-step lie in the row)
__device__ __forceinline__ uint64_t smem_desc_sw64(uint32_t addr) {
  return static_cast<uint64_t>((addr & 0x3FFFFu) >> 4) | (1ull << 16) | (static_cast<uint64_t>(512 >> 4) << 32) | (1ull << 46) |
         (4ull << 61);
}

// K-major SWIZZLE_128B descriptor: rows 128 B apart, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr) {
  return static_cast<uint64_t>((addr & 0x3FFFFu) >> 4) | (1ull << 16) | (static_cast<uint64_t>(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}

#ifdef DSC_WATCHDOG
__device__ unsigned int g_x3_abort = 0;
__device__ unsigned int g_x3_info[4] = {0, 0, 0, 0};
#endif
// Every barrier wait is bounded: a protocol error must end in a trap (an error the host sees), never in a hung GPU.
// (-DDSC_WATCHDOG: record who waited on what, let the kernel drain; read back with dsc_debug_x3_watchdog.)
__device__ __forceinline__ bool test_bar(uint32_t bar, uint32_t parity) {  // non-blocking probe of a barrier phase
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// event loops (producer, issuers): back off when nothing could be done; a loop that makes no progress for ~2 s traps
__device__ __forceinline__ void idle_or_trap(bool progress, uint32_t& spins, long long& t0, unsigned ns) {
  if (progress) {
    spins = 0;
    return;
  }
  __nanosleep(ns);
  if (++spins == 4096) t0 = clock64();
  if (spins > 4096 && (spins & 1023) == 0 && clock64() - t0 > (1ll << 32)) {
#ifdef DSC_WATCHDOG
    if (atomicCAS(&g_x3_abort, 0u, 1u) == 0u) {
      g_x3_info[0] = 99;
      g_x3_info[1] = blockIdx.x;
      g_x3_info[2] = threadIdx.x;
    }
#endif
    __trap();
  }
}
// warp-converged variants (tensor-core issuer warps): every lane probes, the vote makes the answer warp-uniform
__device__ __forceinline__ bool test_bar_u(uint32_t bar, uint32_t parity) { return __all_sync(0xffffffffu, test_bar(bar, parity)); }
__device__ __forceinline__ void wait_bar_u(uint32_t bar, uint32_t parity) {
  // one asm block (see wait_bar): nothing lane-dependent leaves it, so the caller's control flow stays warp-uniform
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .u32 n;\n"
      "mov.u32 n, 0;\n"
      "X3_WAIT_U:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra X3_DONE_U;\n"
      "nanosleep.u32 40;\n"
      "add.u32 n, n, 1;\n"
      "setp.lt.u32 p, n, 0x4000000;\n"
      "@p bra X3_WAIT_U;\n"
      "trap;\n"  // unreachable unless the barrier protocol is broken: never hang the GPU
      "X3_DONE_U:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
// explicit shared-space accesses (32-bit addresses derived from the opaque base: no generic-pointer conversion in the loops)
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t lds32_volatile(uint32_t a) {
  uint32_t v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void red_add_shared(uint32_t a, uint32_t x) {
  asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(x) : "memory");
}
// 2^t for two arguments t <= 0 WITHOUT the MUFU pipe (the softmax's bottleneck: 16 exponentials per clock and SM): Cody-Waite
// split t = n + f with n = floor(t) through the round-down add of 1.5 * 2^23 (the integer lands in the low mantissa bits),
// 2^f on [0, 1) as a cubic (max relative error 7.5e-5, a sixth of the half-ulp of the fp16 the result is rounded to), n added
// to the exponent field.  Arguments below -126 are clamped (the result is < 2^-126 either way: it rounds to zero in P).
__device__ __forceinline__ void exp2_fma2(float& p0, float& p1, float t0, float t1) {
  constexpr float kMagic = 12582912.f;  // 1.5 * 2^23
  t0 = fmaxf(t0, -126.f);
  t1 = fmaxf(t1, -126.f);
  const float r0 = __fadd_rd(t0, kMagic), r1 = __fadd_rd(t1, kMagic);
  float n0, n1, f0, f1, q0, q1;
  fadd2(n0, n1, r0, r1, -kMagic, -kMagic);  // floor(t), exact
  ffma2(f0, f1, n0, n1, -1.f, -1.f, t0, t1);  // t - floor(t), exact
  ffma2(q0, q1, f0, f1, 0.07802421599626541f, 0.07802421599626541f, 0.22606723010540009f, 0.22606723010540009f);
  ffma2(q0, q1, q0, q1, f0, f1, 0.6958337426185608f, 0.6958337426185608f);
  ffma2(q0, q1, q0, q1, f0, f1, 0.9999251961708069f, 0.9999251961708069f);
  p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(r0) << 23));
  p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(r1) << 23));
}
// x, through a shuffle with the caller's own lane: a value ptxas keeps in a register instead of recomputing it
__device__ __forceinline__ uint32_t opaque(uint32_t x) {
  uint32_t lane, y;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
  asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=r"(y) : "r"(x), "r"(lane));
  return y;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(pred));
  return pred != 0;
}
// Blocking wait on a barrier phase.  The whole loop is ONE asm block: the satisfied case is try_wait + one short forward
// branch.  (Written in C++, with the time-out logic inline, the compiler wraps every wait in BSSY / BSYNC, three register
// clears and a far taken branch over the slow path -- ~10 instructions and an instruction-fetch bubble per wait, six waits
// per item.)  A wait that polls 2^26 times (>= 1 s) traps: a protocol error must end in an error the host sees, never in a hung GPU.
template <bool RELAXED>
__device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t parity, uint32_t tag) {
#ifdef DSC_WATCHDOG
  long long t0 = 0;
  uint32_t spins = 0;
  while (true) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
    if (RELAXED) __nanosleep(40);
    if (++spins == 64) t0 = clock64();
    if (spins > 64) {
      if (*reinterpret_cast<volatile unsigned int*>(&g_x3_abort)) return;
      if (clock64() - t0 > (1ll << 30)) {
        if (atomicCAS(&g_x3_abort, 0u, 1u) == 0u) {
          g_x3_info[0] = tag;
          g_x3_info[1] = blockIdx.x;
          g_x3_info[2] = threadIdx.x;
          g_x3_info[3] = parity;
        }
        return;
      }
    }
  }
#else
  (void)tag;
  if constexpr (RELAXED) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .u32 n;\n"
        "mov.u32 n, 0;\n"
        "X3_WAIT_R:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra X3_DONE_R;\n"
        "nanosleep.u32 40;\n"
        "add.u32 n, n, 1;\n"
        "setp.lt.u32 p, n, 0x4000000;\n"
        "@p bra X3_WAIT_R;\n"
        "trap;\n"
        "X3_DONE_R:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
  } else {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .u32 n;\n"
        "mov.u32 n, 0;\n"
        "X3_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra X3_DONE;\n"
        "add.u32 n, n, 1;\n"
        "setp.lt.u32 p, n, 0x4000000;\n"
        "@p bra X3_WAIT;\n"
        "trap;\n"
        "X3_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
  }
#endif
}

#ifdef DSC_TRACE
// Debug build only: clock64 timeline of block 0 -- thread 0 of each consumer warpgroup, the producer, the three issuers
// -> g_x3_trace[pass][who][slot] = {tag, clock}; globaltimer at start / end of every CTA.
__device__ long long g_x3_trace[2][8][1024][2];
__device__ int g_x3_trace_n[2][8];
__device__ unsigned long long g_x3_cta[2][160][2];
__device__ __forceinline__ unsigned long long gtimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define X3_TRACE_DECL                                                                                   \
  int tr_n = 0;                                                                                         \
  const int tr_c = Cfg<HD>::CONSUMERS, tr_s = tr_c + ((STATS && !FUSED) ? 0 : 128);                                 \
  const int tr_k = blockIdx.x != 0 ? -1                                                                 \
                   : (threadIdx.x & 127) == 0 && (int)threadIdx.x < tr_c ? (int)(threadIdx.x >> 7)      \
                   : (int)threadIdx.x == tr_s ? 3                                                       \
                   : ((int)threadIdx.x >= tr_s + 32 && (int)threadIdx.x < tr_s + 32 * (1 + Cfg<HD>::NWG) && (threadIdx.x & 31) == 0) \
                       ? 4 + (int)((threadIdx.x - tr_s - 32) >> 5)                                      \
                   : (!STATS && (int)threadIdx.x == tr_c) ? 7 : -1;
#define X3_TRACE(tag)                                            \
  do {                                                           \
    if (tr_k >= 0 && tr_n < 1024) {                              \
      g_x3_trace[STATS ? 0 : 1][tr_k][tr_n][0] = (tag);          \
      g_x3_trace[STATS ? 0 : 1][tr_k][tr_n][1] = clock64();      \
      g_x3_trace_n[STATS ? 0 : 1][tr_k] = ++tr_n;                \
    }                                                            \
  } while (0)
#define X3_CTA_TIME(k) do { if (threadIdx.x == 0 && blockIdx.x < 160) g_x3_cta[STATS ? 0 : 1][blockIdx.x][k] = gtimer_ns(); } while (0)
#elif defined(DSC_CTATIME)
// Debug build only: globaltimer at the start / end of every CTA and pass, nothing else (the in-kernel span of a call without
// the event-timer quantisation and the launch overhead: scripts/x3_span.py)
__device__ unsigned long long g_x3_cta[2][160][2];
__device__ __forceinline__ unsigned long long gtimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define X3_TRACE_DECL
#define X3_TRACE(tag) do {} while (0)
#define X3_CTA_TIME(k) do { if (threadIdx.x == 0 && blockIdx.x < 160) g_x3_cta[STATS ? 0 : 1][blockIdx.x][k] = gtimer_ns(); } while (0)
#else
#define X3_TRACE_DECL
#define X3_TRACE(tag) do {} while (0)
#define X3_CTA_TIME(k) do {} while (0)
#endif

#ifdef DSC_PHASE
// Debug build only: cycles per phase of the pass-2 consumer loop, accumulated in registers (no stores inside the loop), written
// once at the end by lane 0 of every consumer warp of blocks 0..3 -> g_x3_phase[block][warp][phase]
__device__ unsigned int g_x3_phase[4][12][8];
// (-DDSC_PHASE_SKIP_FIRST: the warp's first item -- pass 2's start-up: first tile, first record, the std -- is left out)
#ifdef DSC_PHASE_SKIP_FIRST
#define X3_PH_DECL long long ph_t = clock64(); unsigned int ph[8] = {0, 0, 0, 0, 0, 0, 0, 0}; bool ph_on = false;
#define X3_PH(k) do { const long long ph_c = clock64(); if (ph_on) ph[k] += static_cast<unsigned int>(ph_c - ph_t); ph_t = ph_c; if ((k) == 7) ph_on = true; } while (0)
#else
#define X3_PH_DECL long long ph_t = clock64(); unsigned int ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define X3_PH(k) do { const long long ph_c = clock64(); ph[k] += static_cast<unsigned int>(ph_c - ph_t); ph_t = ph_c; } while (0)
#endif
#define X3_PH_FLUSH do { if (!STATS && lane == 0 && blockIdx.x < 4 && warp < 12) { for (int k = 0; k < 8; ++k) g_x3_phase[blockIdx.x][warp][k] = ph[k]; } } while (0)
#else
#define X3_PH_DECL
#define X3_PH(k) do {} while (0)
#define X3_PH_FLUSH do {} while (0)
#endif

// One pass over the CTA's tile range.  FUSED: both passes run in ONE (cooperative) launch, x3_phase<STATS> then
// x3_phase<!STATS> with the pass-2 thread layout (the drain warpgroup idles through pass 1), TMEM allocated once (returned by
// pass 1, handed to pass 2), pass 1's barriers invalidated before pass 2 lays out its own; the std goes through the handoff
// slots.  Returns the TMEM base address.
template <typename T, int HD, bool STATS, bool FUSED>
__device__ __forceinline__ uint32_t x3_phase(const XattnParams& p, const CUtensorMap& tm_qa, const CUtensorMap& tm_qb,
                                             const CUtensorMap& tm_qp, const CUtensorMap& tm_oa, const CUtensorMap& tm_ob,
                                             uint32_t tmem_in) {
  using C = Cfg<HD>;
  constexpr int D = C::D, HPT = C::HPT, LOG_HPT = C::LOG_HPT, NWG = C::NWG, CONSUMERS = C::CONSUMERS, WG_COLS = C::WG_COLS;
  constexpr int K_HEAD = C::K_HEAD, VT_CH = C::VT_CH, IMG_BYTES = C::IMG_BYTES, NSLOT = C::NSLOT;
  constexpr int DW0 = 4 * NWG;                    // pass 2: first warp of the drain warpgroup
  constexpr int SW0 = (STATS && !FUSED) ? 4 * NWG : 4 * NWG + 4;  // first service warp
  constexpr int NST = STATS ? STATS_NST : C::FWD_NST;
  constexpr int STAGE = STATS ? STATS_STAGE : FWD_STAGE;
  constexpr int RECB = STATS ? K_HEAD : C::REC_BYTES;  // bytes of a record this pass needs (pass 1: its K part) = slot pitch
  constexpr int KV = (NSLOT * RECB + 1023) / 1024 * 1024;  // the ring stages behind it hold swizzled boxes: 1024-byte aligned
  extern __shared__ __align__(1024) unsigned char smem[];
  // thread index and shared-memory base are made opaque (a shuffle with the thread's own lane): left alone, ptxas
  // re-materialises them (S2R SR_TID.X / SR_CgaCtaId + LEA, long-latency special-register reads) in front of every
  // barrier operation of the role loops
  const int tid = opaque(threadIdx.x);
  const int warp = tid >> 5, lane = tid & 31;
  X3_TRACE_DECL
  X3_CTA_TIME(0);
  X3_TRACE(1);
  if (warp == SW0 + 3 && lane < 5) {  // hide the descriptor fetches behind the rest of the prologue
    const CUtensorMap* m = lane == 0 ? &tm_qa : lane == 1 ? &tm_qb : lane == 2 ? &tm_qp : lane == 3 ? &tm_oa : &tm_ob;
    asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
  }
  if constexpr (STATS) {
    // pass 1 is itself a programmatic dependent of whatever precedes it in the stream: nothing of its inputs is touched
    // before that kernel has completed; pass 2 may be placed as SMs free up
    pdl_wait_prior_grid();
    pdl_launch_dependents();
  } else if constexpr (!FUSED) {
    pdl_launch_dependents();  // a following pass 1 (next call) may be placed early; it waits for our completion itself
  }
  X3_TRACE(3);
  const uint32_t s0 = opaque(smem_u32(smem));
  const uint32_t sStage = s0 + KV;
  const uint32_t bars = sStage + NST * STAGE;
  // barrier map (8 B each): full[4] | odone[4] | kvfull[8] | kvfree[8] | srdy[3][2] | sfree[3][2] | prdy[3] | ordy[3] | ofree[3] |
  // std[1] (pass 2, handoff: the folding warp has put the std at std_addr)
  const uint32_t b_full = bars, b_odone = bars + 32, b_kvfull = bars + 64, b_kvfree = bars + 128, b_srdy = bars + 192,
                 b_sfree = bars + 240, b_prdy = bars + 288, b_ordy = bars + 312, b_ofree = bars + 336, b_std = bars + 360;
  constexpr int N_BARS = 46;
  static_assert(NSLOT <= 8, "kvfull / kvfree barrier arrays");
  const uint32_t turn_addr = bars + 384;
  const uint32_t std_addr = bars + 408;
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + KV + NST * STAGE + 400);

  const int begin = static_cast<int>(blockIdx.x * p.tiles_q + min(blockIdx.x, p.tiles_r));
  const int n_items = static_cast<int>(p.tiles_q + (blockIdx.x < p.tiles_r ? 1u : 0u));
  // the CTA's tiles span the (batch, head group) segments seg0 .. seg0 + n_runs - 1 ("runs"); records = n_runs * HPT
  const int seg0 = static_cast<int>(fastdiv(static_cast<uint32_t>(begin), p.div_nsl));
  const int n_rec = n_items > 0 ? (static_cast<int>(fastdiv(static_cast<uint32_t>(begin + n_items - 1), p.div_nsl)) - seg0 + 1) * HPT : 0;
  const uint64_t pol_q = STATS ? policy_evict_last() : policy_evict_first();  // pass 2 reads Q again: keep it in L2
  const unsigned char* img = reinterpret_cast<const unsigned char*>(p.kv_image);

  // The producer WARP runs converged (see the issuer warps below): addresses / coordinates are warp-uniform, the TMA
  // instructions are issued by the elected lane (always the same one: bulk groups are per thread)
  auto load_record = [&](int rec) {  // head record (K_h | V^T_h; pass 1: K_h) of run rec / HPT -> slot rec % NSLOT: one bulk copy
    const int slot = rec % NSLOT;
    const unsigned char* src = img + static_cast<size_t>(seg0 + (rec >> LOG_HPT)) * IMG_BYTES + (rec & (HPT - 1)) * C::REC_BYTES;
    if (elect_one()) {
      mbar_arrive_expect_tx(b_kvfull + 8 * slot, RECB);
      bulk_g2s_hint(s0 + slot * RECB, src, RECB, b_kvfull + 8 * slot, policy_evict_last());
    }
    __syncwarp();
  };
  auto load_tile = [&](int i) {  // Q boxes (+ compact W tile) of tile i -> ring stage i % NST
    const int s = i % NST;
    const Tile t = decode(begin + i, p);
    const uint32_t sQ = sStage + s * STAGE, bar = b_full + 8 * s;
    const int c0 = t.hg * GW;
    uint32_t wbytes = 0;
    const float* wsrc = nullptr;
    if constexpr (!STATS) {
      wbytes = static_cast<uint32_t>(min(ROWS, p.L - t.l0)) * (DSC_COMPACT_PITCH * 4);
      wsrc = p.wc + (static_cast<size_t>(t.b / (p.B / p.Bw)) * p.L + t.l0) * DSC_COMPACT_PITCH;
    }
    const bool run_ends = t.tile == p.n_sl - 1 && i + 1 < n_items;  // pull the next image towards L2
    const bool pre = X3_L2_PREFETCH && i + NST < n_items;           // the tile that will reuse this stage: towards L2
    const Tile n1 = decode(begin + min(i + 1, n_items - 1), p);
    const Tile n2 = decode(begin + min(i + NST, n_items - 1), p);
    if (elect_one()) {
      mbar_arrive_expect_tx(bar, QT_BYTES + wbytes);  // out-of-range parts of a box are zero-filled and still counted
      tma_load_3d(sQ, &tm_qa, c0, t.l0, t.b, bar, pol_q);
      tma_load_3d(sQ + BOX128_BYTES, &tm_qa, c0 + 64, t.l0, t.b, bar, pol_q);
      tma_load_3d(sQ + 2 * BOX128_BYTES, &tm_qb, c0 + 128, t.l0, t.b, bar, pol_q);
      if constexpr (!STATS)  // the compact W rows of a tile are contiguous: one bulk copy
        bulk_g2s_hint(sQ + QT_BYTES, wsrc, wbytes, bar, pol_q);
      if (run_ends)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(img + (static_cast<size_t>(n1.b) * p.n_hg + n1.hg) * IMG_BYTES),
                     "r"(IMG_BYTES)
                     : "memory");
      if (pre) {  // one 160-column box (+ its W rows)
        asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(&tm_qp), "r"(n2.hg * GW), "r"(n2.l0),
                     "r"(n2.b)
                     : "memory");
        if constexpr (!STATS)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(
                           p.wc + (static_cast<size_t>(n2.b / (p.B / p.Bw)) * p.L + n2.l0) * DSC_COMPACT_PITCH),
                       "r"(static_cast<uint32_t>(min(ROWS, p.L - n2.l0)) * (DSC_COMPACT_PITCH * 4))
                       : "memory");
      }
    }
    __syncwarp();
  };

  // barrier init is spread over the service warps so that the first loads leave as early as possible: the producer
  // warp initialises only what those loads signal (full[], kvfull[]), service warp 2 the rest
  if (warp == SW0) {
    if (lane < 12) mbar_init((lane < 4 ? b_full : b_kvfull - 32) + 8 * lane, 1);  // full[4], kvfull[8]
    fence_mbar_init();
    __syncwarp();
    X3_TRACE(4);
    if (n_items > 0) {  // only what the first Q K^T needs is issued ahead of the CTA-wide barrier
      load_record(0);
      load_tile(0);
      X3_TRACE(6);
    }
  }
  if (warp == SW0 + 2) {
    for (int idx = 4 + lane; idx < N_BARS; idx += 32) {
      if (idx >= 8 && idx < 16) continue;  // kvfull[]: the producer's
      // odone[]: 128 rows of every head of the tile | kvfree[]: every consumer | srdy, ordy: one commit | sfree, prdy: a warpgroup |
      // ofree[]: the drain warpgroup
      const uint32_t cnt = idx < 8 ? HPT * 128u : idx < 24 ? static_cast<uint32_t>(CONSUMERS) : idx < 30 ? 1u : idx < 39 ? 128u : idx < 42 ? 1u : idx < 45 ? 128u : 1u;
      mbar_init(bars + 8 * idx, cnt);
    }
    fence_mbar_init();
    if (lane >= 28) asm volatile("st.shared.u32 [%0], %1;" ::"r"(turn_addr + 4 * (lane - 28)), "r"(0u) : "memory");  // whose turn it is on each SM sub-partition (pass 2)
  }
  constexpr bool kAllocTmem = STATS || !FUSED;  // fused pass 2 inherits pass 1's allocation
  if (kAllocTmem && warp == SW0 + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
                     smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    X3_TRACE(7);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = kAllocTmem ? *tmem_ptr_smem : tmem_in;
  X3_TRACE(2);

  const int n_jobs = n_items * HPT;  // (tile, head) work items of this CTA, in order; item J -> warpgroup J % NWG

  if (warp >= SW0) {
    if constexpr (!STATS) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(C::REGS_SERVICE));
    if (warp == SW0) {
      // ============================== producer warp: TMA loads and stores ==========================
      // It never blocks on one condition while another could make progress:
      //   * head record rec goes into slot rec % NSLOT as soon as every warpgroup has released record rec - NSLOT
      //   * tile ld goes into ring stage ld % NST as soon as tile ld - NST has been retired and its store has read the stage
      //   * tile dn is retired (pass 2: its O rows leave through three tensor-map stores) when every row has been handed back
      // While a record refill may become possible (the consumers are within a tile of the end of the run that still
      // uses the slot) both conditions are polled; otherwise the warp sleeps on the hand-back barrier.
      int ld = n_items > 0 ? 1 : 0, dn = 0, rec = n_rec > 0 ? 1 : 0;
      uint32_t idle_spins = 0;
      long long idle_t0 = 0;
      X3_TRACE(8);
      while (dn < n_items) {
        while (rec < n_rec && (rec < NSLOT || test_bar_u(b_kvfree + 8 * (rec % NSLOT), (rec / NSLOT - 1) & 1))) {
          load_record(rec++);
          X3_TRACE(33);
        }
        while (ld < n_items && ld < dn + NST) {
          if constexpr (!STATS) {
            if (ld >= NST) {  // the store of tile ld - NST (bulk group ld - NST of dn committed so far) must have read the stage
              const int later = dn - 1 - (ld - NST);  // groups committed after it: may stay pending
              if (elect_one()) {
                if (later <= 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                else if (later == 1) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                else asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
              }
              __syncwarp();
            }
          }
          load_tile(ld++);
          X3_TRACE(34);
        }
        bool poll = false;
        if (rec < n_rec) {  // last tile of the run whose record occupies the slot record `rec` is waiting for
          const int e_old = min(n_items, (seg0 + ((rec - NSLOT) >> LOG_HPT) + 1) * p.n_sl - begin) - 1;
          poll = dn + 1 >= e_old;
        }
        if (poll) {
          if (!test_bar_u(b_odone + 8 * (dn % NST), (dn / NST) & 1)) {
            idle_or_trap(false, idle_spins, idle_t0, 32);
            continue;
          }
          idle_spins = 0;
        } else {
          wait_bar_u(b_odone + 8 * (dn % NST), (dn / NST) & 1);
        }
        X3_TRACE(31);
        if constexpr (!STATS) {
          const Tile t = decode(begin + dn, p);
          const uint32_t sQ = sStage + (dn % NST) * STAGE;
          const int c0 = t.hg * GW;
          if (elect_one()) {
            tma_store_3d(&tm_oa, c0, t.l0, t.b, sQ);  // rows >= L and columns >= H*D are clipped by the TMA
            tma_store_3d(&tm_oa, c0 + 64, t.l0, t.b, sQ + BOX128_BYTES);
            tma_store_3d(&tm_ob, c0 + 128, t.l0, t.b, sQ + 2 * BOX128_BYTES);
            bulk_commit();
          }
          __syncwarp();
          X3_TRACE(32);
        }
        ++dn;
      }
      if constexpr (!STATS) {
        // the CTA must outlive the stores' READS of shared memory; their global writes complete with the grid
        if (elect_one()) bulk_wait_read0();
        __syncwarp();
      } else if constexpr (FUSED) {
        // single launch: the Q tiles are on their way; pull towards L2 what pass 2's first tiles need and pass 1 never
        // touched (the compact W rows, the V^T halves of the first image) while the consumers finish pass 1
        const int npre = min(n_items, C::FWD_NST);
        for (int i = 0; i < npre; ++i) {
          const Tile t = decode(begin + i, p);
          const float* w = p.wc + (static_cast<size_t>(t.b / (p.B / p.Bw)) * p.L + t.l0) * DSC_COMPACT_PITCH;
          const uint32_t wb = static_cast<uint32_t>(min(ROWS, p.L - t.l0)) * (DSC_COMPACT_PITCH * 4);
          if (elect_one()) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(w), "r"(wb) : "memory");
        }
        if (n_items > 0 && elect_one())
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(img + static_cast<size_t>(seg0) * IMG_BYTES), "r"(IMG_BYTES) : "memory");
        __syncwarp();
      }
    } else if (warp > SW0 && warp - SW0 - 1 < NWG) {
      // ============================== tensor-core issuer of warpgroup g ============================
      // The WHOLE warp runs this loop converged: every value that feeds an MMA (descriptors, TMEM addresses) is
      // warp-uniform to the compiler (warp index through a shuffle, barrier probes through votes), so it lives in uniform
      // registers and an MMA is ONE predicated UTCHMMA issued by the elected lane -- a single-lane loop gets every operand
      // through an ELECT / R2UR.BROADCAST waterfall, ~120 cycles per MMA (profiles/r2_x3_trace_*: 350 / 700 cycles to issue
      // the 3 / 5 MMAs of an item), and that latency sits on the consumers' critical path (P published -> P V done).
      // Steady state: Q K^T of item j + NWG when the S columns are free (the consumers have read S(j)), then P V of item j
      // when its P is published -- each a sleeping wait on one barrier.  The exception is a Q K^T whose head record has not
      // arrived yet: its slot is released only after earlier P Vs have completed, so the issuer never blocks on a record
      // while a P V is outstanding; it issues that P V first and comes back.
      // Invariant that keeps the consumers' record release alive: before blocking for Q K^T(jq), P V(jq - 2 NWG) has
      // been issued (jq - jp <= NWG).
      const int g = __shfl_sync(0xffffffffu, warp, 0) - SW0 - 1;
      constexpr uint32_t idesc_qk = idesc_f16<T>(80);
      constexpr uint32_t idesc_pv = idesc_f16<T>(C::ON);
      constexpr int AHEAD = STATS ? 2 * NWG : NWG;  // pass 1: S is double-buffered
      const uint32_t tw = __shfl_sync(0xffffffffu, tmem_base, 0) + g * WG_COLS;
      uint32_t nqk = 0, npv = 0;
      int jq = g, jp = g;
      // (batch, head group) run of a tile, tracked without divisions: tiles only move forward
      const int run_end0 = (seg0 + 1) * p.n_sl - begin;  // first local tile of run 1
      int q_run = 0, q_next = run_end0, p_run = 0, p_next = run_end0;
      // records this issuer has SEEN loaded, in record order.  Every record is observed, used by this warpgroup or not: a
      // parity test is only meaningful against the phase right after the last one observed (testing the phase of a later
      // record while an earlier load into the same slot is still in flight would read as "complete")
      int rec_seen = 0;
      while (STATS ? jq < n_jobs : jp < n_jobs) {
        bool qk_now = false;
        int rc = 0;
        if (jq < n_jobs && (STATS || jq - jp <= AHEAD)) {
          const int iq = jq >> LOG_HPT;
          while (iq >= q_next) {
            ++q_run;
            q_next += p.n_sl;
          }
          rc = q_run * HPT + (jq & (HPT - 1));
          while (rec_seen <= rc && test_bar_u(b_kvfull + 8 * (rec_seen % NSLOT), (rec_seen / NSLOT) & 1)) ++rec_seen;
          qk_now = rec_seen > rc;
          if (!qk_now && (STATS || jp == jq)) {  // nothing else to do: sleep on the record that is next in line
            wait_bar_u(b_kvfull + 8 * (rec_seen % NSLOT), (rec_seen / NSLOT) & 1);
            ++rec_seen;
            continue;
          }
        }
        if (qk_now) {
          const int i = jq >> LOG_HPT, h = jq & (HPT - 1), s = i % NST, slot = rc % NSLOT;
          X3_TRACE(42);
          wait_bar_u(b_full + 8 * s, (i / NST) & 1);
          uint32_t buf = 0;
          if constexpr (STATS) {
            buf = nqk & 1;
            if (nqk >= 2) wait_bar_u(b_sfree + 16 * g + 8 * buf, ((nqk >> 1) - 1) & 1);
          } else {
            if (nqk >= 1) wait_bar_u(b_sfree + 16 * g, (nqk - 1) & 1);
          }
          X3_TRACE(43);
          ++nqk;
          tc_fence_after();
          const uint32_t sQ = sStage + s * STAGE;
          const uint32_t d = tw + (STATS ? buf * S1_COL : S_COL);
          const uint32_t kb = s0 + slot * RECB;
          // head h = the KSTEPS 16-column blocks that cover columns HD*h .. HD*h + HD-1 (HD = 40: the K image is zero where
          // a block's columns belong to a neighbour); block t: boxes of 4 blocks (SW128) for t < 8, the SW64 box for t = 8, 9
          const int t0b = (h * D) >> 4;
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < C::KSTEPS; ++ks) {
              const int t = t0b + ks;
              const uint64_t ad = t < 8 ? smem_desc_sw128(sQ + (t >> 2) * BOX128_BYTES + (t & 3) * 32)
                                        : smem_desc_sw64(sQ + 2 * BOX128_BYTES + (t - 8) * 32);
              umma_ss(d, ad, smem_desc(kb + ks * 2 * K_CH, K_CH, 128), idesc_qk, ks);
            }
            tc_commit(b_srdy + 16 * g + 8 * buf);
          }
          __syncwarp();
          X3_TRACE(44);
          jq += NWG;
          if constexpr (!STATS) {
            if (jq - jp <= AHEAD && jq < n_jobs) continue;  // start-up: the first item's Q K^T pair before any P V
          }
        }
        if constexpr (!STATS) {
          if (jp < jq) {
            X3_TRACE(45);
            wait_bar_u(b_prdy + 8 * g, npv & 1);
            if (npv >= 1) wait_bar_u(b_ofree + 8 * g, (npv - 1) & 1);  // the drain warpgroup holds O of the previous item in registers
            X3_TRACE(46);
            ++npv;
            tc_fence_after();
            const int ip = jp >> LOG_HPT;
            while (ip >= p_next) {
              ++p_run;
              p_next += p.n_sl;
            }
            const uint32_t vb = s0 + ((p_run * HPT + (jp & (HPT - 1))) % NSLOT) * RECB + K_HEAD;
            const uint64_t vdesc = smem_desc(vb, VT_CH, 128);
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < 5; ++kk)
                umma_ts(tw + O_COL, tw + P_COL + kk * 8, vdesc + static_cast<uint64_t>((kk * 2 * VT_CH) >> 4), idesc_pv, kk);
              tc_commit(b_ordy + 8 * g);
            }
            __syncwarp();
            X3_TRACE(47);
            jp += NWG;
          }
        }
      }
    }
    if constexpr (!STATS) {
      // formal ordering: pass 2 complete => pass 1 complete (it is, long ago: returns at once)
      if (!FUSED && warp == SW0 && p.handoff) pdl_wait_prior_grid();
    }
    __syncwarp();
  } else if (STATS && warp >= DW0) {
    // (fused launch, pass 1: the drain warpgroup has nothing to do)
  } else if (!STATS && warp >= DW0) {
    // ============================== drain warpgroup (pass 2): O rows out of TMEM ====================
    // Warp d serves TMEM lanes 32d .. 32d+31 (tile rows) of every consumer warpgroup.  Items in CTA order: wait for P V(J)
    // (ordy of warpgroup J % NWG), O row -> registers, O columns handed back to the issuer (ofree: P V of the warpgroup's
    // next item may overwrite them), x 1/rowsum (ones row of V^T: column HD) -> over the row's own Q columns in the ring
    // stage, stage handed back to the producer (odone).  The consumer warps never touch O: their loop is S -> softmax -> P.
    if constexpr (!STATS) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(C::REGS_DRAIN));
      if (warp == DW0 && (FUSED || p.handoff)) {
        // ------------------------------ first: drain warp 0 folds pass 1's partials (one call) -------
        // pass 1's CTAs publish {sum, sum of squares} into the handoff slots; this warp polls them (volatile loads: L2),
        // folds them in the fixed order of finalize_stats and hands the std to the consumers -- no ticket, no fold in pass 1's
        // last CTA, no wait for pass 1's completion and the memory flush behind it (griddepcontrol.wait).  Every CTA folds
        // the same values in the same order: bit-identical std everywhere.  The CTA that reads last zeroes the slots again
        // and fills the public statistics header.
        const volatile unsigned long long* slots =
            reinterpret_cast<const volatile unsigned long long*>(reinterpret_cast<const unsigned char*>(p.ws) + kHandoffOffset);
        const unsigned int n_part = gridDim.x;  // pass 1 runs the same grid (launch())
        double sa, sb;
        long long t0 = 0;
        uint32_t spins = 0;
        while (true) {
          bool ok = true;
          sa = 0.0;
          sb = 0.0;
          for (unsigned int c = lane; c < n_part; c += 32) {
            const unsigned long long va = slots[2 * c], vb = slots[2 * c + 1];
            ok = ok && va != 0ull && vb != 0ull;
            sa += __longlong_as_double(static_cast<long long>(va));
            sb += __longlong_as_double(static_cast<long long>(vb));
          }
          if (__all_sync(0xffffffffu, ok)) break;
          __nanosleep(100);
          if (++spins == 4096) t0 = clock64();
          if (spins > 4096 && (spins & 255) == 0 && clock64() - t0 > (1ll << 33)) __trap();
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          sa += __shfl_xor_sync(0xffffffffu, sa, o);
          sb += __shfl_xor_sync(0xffffffffu, sb, o);
        }
        const double scl = static_cast<double>(p.scale);
        const double n = static_cast<double>(p.B) * p.H * static_cast<double>(p.L) * p.S;
        const double sum = sa * scl, sumsq = sb * scl * scl, mean = sum / n;
        double var = (n > 1.0) ? (sumsq - sum * mean) / (n - 1.0) : nan("");
        if (var < 0.0) var = 0.0;
        const float std_f = static_cast<float>(sqrt(var));
        if (lane == 0) {
          asm volatile("st.shared.u32 [%0], %1;" ::"r"(std_addr), "r"(__float_as_uint(std_f)) : "memory");
          mbar_arrive(b_std);  // (release: the store above is visible to the waiting consumers)
        }
        unsigned int last = 0;
        if (lane == 0) last = atomicAdd(reinterpret_cast<unsigned int*>(reinterpret_cast<unsigned char*>(p.ws) + kReadersOffset), 1u) == gridDim.x - 1 ? 1u : 0u;
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {  // every CTA has read the slots: leave them empty for the next call, publish the statistics
          unsigned long long* w = reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(p.ws) + kHandoffOffset);
          for (unsigned int c = lane; c < 2 * n_part; c += 32) w[c] = 0ull;
          if (lane == 0) {
            p.ws->std_unbiased = std_f;
            p.ws->mean = static_cast<float>(mean);
            p.ws->sum = sum;
            p.ws->sumsq = sumsq;
            p.ws->n = n;
            p.ws->n_partials = n_part;
            *reinterpret_cast<unsigned int*>(reinterpret_cast<unsigned char*>(p.ws) + kReadersOffset) = 0u;
          }
        }
      }
      const int row = (warp & 3) * 32 + lane;
      const uint32_t tl = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + O_COL;
      auto pack_store = [&](const float* o8, float inv, uint32_t d) {
        float t[8];
        fmul2(t[0], t[1], o8[0], o8[1], inv, inv);
        fmul2(t[2], t[3], o8[2], o8[3], inv, inv);
        fmul2(t[4], t[5], o8[4], o8[5], inv, inv);
        fmul2(t[6], t[7], o8[6], o8[7], inv, inv);
        uint4 v;
        v.x = Mma<T>::pack(t[0], t[1]);
        v.y = Mma<T>::pack(t[2], t[3]);
        v.z = Mma<T>::pack(t[4], t[5]);
        v.w = Mma<T>::pack(t[6], t[7]);
        sts128(d, v);
      };
      int g = 0, k = 0;  // item J = NWG * k + g
      for (int J = 0; J < n_jobs; ++J) {
        const int i = J >> LOG_HPT, h = J & (HPT - 1), s = i % NST;
        const uint32_t tw = tl + g * WG_COLS;
        const uint32_t st = sStage + s * STAGE;
        // 16-byte chunk c of the head's O row = global chunk G = (HD/8) h + c of the 160-column tile row: the place its Q
        // columns had (swizzled: 8 consecutive rows hit 8 distinct bank groups)
        auto dst_of = [&](int c) -> uint32_t {
          const int G = h * C::CPH + c;
          return G < 16 ? st + (G >> 3) * BOX128_BYTES + row * 128 + (((G & 7) ^ (row & 7)) << 4)
                        : st + 2 * BOX128_BYTES + row * 64 + ((((G - 16) & 3) ^ ((row >> 1) & 3)) << 4);
        };
        X3_TRACE(13);
        wait_bar<false>(b_ordy + 8 * g, k & 1, 7);
        X3_TRACE(14);
        tc_fence_after();
        if constexpr (HD == 40) {
          float o[40], oz[4];
          tmem_ld_x16(tw, reinterpret_cast<uint32_t*>(o));
          tmem_ld_x16(tw + 16, reinterpret_cast<uint32_t*>(o + 16));
          tmem_ld_x8(tw + 32, reinterpret_cast<uint32_t*>(o + 32));
          tmem_ld_x4(tw + 40, reinterpret_cast<uint32_t*>(oz));
          tc_wait_ld();
          tc_fence_before();
          mbar_arrive(b_ofree + 8 * g);
          const float inv = 1.f / oz[0];
#pragma unroll
          for (int c = 0; c < 5; ++c) pack_store(o + 8 * c, inv, dst_of(c));
        } else {
          // HD = 80 / 160: the row sum (column HD) first, then PIECE columns at a time, the next piece in flight while the
          // current one is scaled, packed and stored
          constexpr int PIECE = HD == 160 ? 32 : 16, NP = HD / PIECE;
          float oz[4], oa[PIECE], ob[PIECE];
          auto ld_piece = [&](int pc, float* dst) {
            if constexpr (PIECE == 32) tmem_ld_x32(tw + PIECE * pc, reinterpret_cast<uint32_t*>(dst));
            else tmem_ld_x16(tw + PIECE * pc, reinterpret_cast<uint32_t*>(dst));
          };
          tmem_ld_x4(tw + HD, reinterpret_cast<uint32_t*>(oz));
          ld_piece(0, oa);
          tc_wait_ld();
          const float inv = 1.f / oz[0];
#pragma unroll
          for (int pc = 0; pc < NP; ++pc) {
            float* cur = (pc & 1) ? ob : oa;
            if (pc + 1 < NP) ld_piece(pc + 1, (pc & 1) ? oa : ob);
#pragma unroll
            for (int q = 0; q < PIECE / 8; ++q) pack_store(cur + 8 * q, inv, dst_of((PIECE / 8) * pc + q));
            if (pc + 1 < NP) tc_wait_ld();
          }
          tc_fence_before();
          mbar_arrive(b_ofree + 8 * g);
        }
        fence_proxy_async();  // O rows -> visible to the TMA store
        mbar_arrive(b_odone + 8 * s);
        X3_TRACE(15);
        if (++g == NWG) {
          g = 0;
          ++k;
        }
      }
    }
  } else {
    // ============================== consumers: one thread per query row ============================
    if constexpr (!STATS) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(C::REGS_CONSUMER));
    const int g = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t tw = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + g * WG_COLS;
    float beta_l2 = 0.f;  // sigma * std * log2(e); read after pass 1 has completed (PDL), right before first use
    bool have_beta = STATS;
    const float scale_l2 = p.scale * kLog2e;
    // logits in the log2 domain.  Compact region map: the weighted key columns are slots 0..15 (keys permuted in the
    // prepared image), one 80-byte row of W per query.  y = s * a + w * bw, 2^(e * y - e * max(y)): a = scale / beta,
    // bw = 1, e = beta -- or, for a vanishing beta, a = scale, bw = beta, e = 1   (set once, when beta is known)
    bool bpos = true;
    float ca = 0.f, cbw = 1.f, ce = 1.f;
    // sigma does not depend on pass 1 (whatever wrote it completed before pass 1 released its dependents): fetched now, so
    // that a DRAM-cold line is not waited for after the std has arrived
    float sigma_v = 0.f;
    if constexpr (!STATS) sigma_v = p.sigma_dev ? __ldcg(p.sigma_dev) : p.sigma_host;
    double dsum = 0.0, dsq = 0.0;
    uint32_t n_s = 0, n_o = 0;
    bool pend = false;  // P V of this warpgroup's previous item may still be reading P (and the head record's V^T)
    // 3 warpgroups (HD = 40): the exponentials of an item ("M phase": 77 MUFU.EX2 per row, the bottleneck pipe: 4 lanes per
    // clock and SM sub-partition) are SERIALISED per sub-partition in item order: warp (g, quarter) waits until
    // turn[quarter] is within TURNS of its item's sequence number.  Left alone, the three warps of a sub-partition run in
    // lock-step -- they share the MUFU pipe fairly, so they finish their M phases together and then all do their MUFU-free
    // work (S row out of TMEM, W, row max, barriers) while the pipe idles.  With turns, one or two warps exponentiate at
    // the full pipe rate while the others do their MUFU-free part.
    const uint32_t my_turn = turn_addr + 4 * (warp & 3);

    // P V of the warpgroup's previous item has completed (ordy: its commit): P may be overwritten, the record's V^T released
    auto pv_done = [&]() {
      if (pend) {
        X3_TRACE(13);
        wait_bar<false>(b_ordy + 8 * g, n_o & 1, 7);
        X3_TRACE(14);
        ++n_o;
        pend = false;
      }
    };

    // Head records are released in record order: this warpgroup arrives on kvfree[record] once it has no item left
    // that is <= the record's last item -- checked at the top of every item, after the pending O row (whose P V read the
    // record's V^T) has been drained.  (The check must not sit behind the wait for the item's own S: its Q K^T may need
    // the very slot this arrival frees.)
    int arr_next = 0;
    auto last_job_of = [&](int rc) {  // last (tile, head) item that reads record rc
      const int e = min(n_items, (seg0 + (rc >> LOG_HPT) + 1) * p.n_sl - begin) - 1;
      return e * HPT + (rc & (HPT - 1));
    };
    int arr_last = n_rec > 0 ? last_job_of(0) : 0x7fffffff;
    {
      X3_PH_DECL
      for (int j = g; j < n_jobs; j += NWG) {
        const int i = j >> LOG_HPT, s = i % NST;
        if (arr_last < j) {
          if constexpr (!STATS) pv_done();
          do {
            mbar_arrive(b_kvfree + 8 * (arr_next % NSLOT));
            ++arr_next;
            arr_last = arr_next < n_rec ? last_job_of(arr_next) : 0x7fffffff;
          } while (arr_last < j);
        }
        float sc[80];
        if constexpr (STATS) {
          const uint32_t buf = n_s & 1;
          X3_TRACE(10);
          wait_bar<false>(b_srdy + 16 * g + 8 * buf, (n_s >> 1) & 1, 8);
          X3_TRACE(11);
          ++n_s;
          tc_fence_after();
          tmem_ld_x64(tw + buf * S1_COL, reinterpret_cast<uint32_t*>(sc));
          tmem_ld_x16(tw + buf * S1_COL + 64, reinterpret_cast<uint32_t*>(sc) + 64);
          tc_wait_ld();
          tc_fence_before();
          mbar_arrive(b_sfree + 16 * g + 8 * buf);
          mbar_arrive(b_odone + 8 * s);  // pass 1 only reads Q: this item's Q K^T is complete
          X3_TRACE(12);
          // rows beyond L were zero-filled by the TMA and pad keys are zero rows of K: they add exact zeros
          float fs[4] = {0.f, 0.f, 0.f, 0.f}, fq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int c = 0; c < 80; c += 4) {
            fadd2(fs[0], fs[1], fs[0], fs[1], sc[c], sc[c + 1]);
            fadd2(fs[2], fs[3], fs[2], fs[3], sc[c + 2], sc[c + 3]);
            ffma2(fq[0], fq[1], sc[c], sc[c + 1], sc[c], sc[c + 1], fq[0], fq[1]);
            ffma2(fq[2], fq[3], sc[c + 2], sc[c + 3], sc[c + 2], sc[c + 3], fq[2], fq[3]);
          }
          dsum += static_cast<double>((fs[0] + fs[1]) + (fs[2] + fs[3]));
          dsq += static_cast<double>((fq[0] + fq[1]) + (fq[2] + fq[3]));
        } else {
          X3_TRACE(10);
          X3_PH(0);
          wait_bar<false>(b_srdy + 16 * g, n_s & 1, 8);
          ++n_s;
          X3_TRACE(11);
          X3_PH(1);
          tc_fence_after();
          tmem_ld_x64(tw + S_COL, reinterpret_cast<uint32_t*>(sc));
          tmem_ld_x16(tw + S_COL + 64, reinterpret_cast<uint32_t*>(sc) + 64);
          tc_wait_ld();
          tc_fence_before();
          mbar_arrive(b_sfree + 16 * g);  // the next item's Q K^T may overwrite S
          X3_TRACE(12);
          X3_PH(2);
          if (!have_beta) {
            X3_TRACE(17);
            float std_v;
            if (FUSED || p.handoff) {  // the folding warp of this CTA has summed pass 1's per-CTA partials
              wait_bar<false>(b_std, 0, 11);
              std_v = __uint_as_float(lds32_volatile(std_addr));
            } else {
              pdl_wait_prior_grid();  // pass 1 (same stream, launched just before) has published the std
              std_v = __ldcg(&p.ws->std_unbiased);
            }
            X3_TRACE(18);
            beta_l2 = sigma_v * std_v * kLog2e;
            have_beta = true;
            bpos = beta_l2 > 1e-20f;
            ca = bpos ? scale_l2 / beta_l2 : scale_l2;
            cbw = bpos ? 1.f : beta_l2;
            ce = bpos ? beta_l2 : 1.f;
          }
          wait_bar<false>(b_full + 8 * s, (i / NST) & 1, 9);  // (long complete: the Q K^T needed the stage)
          const uint32_t wt4 = sStage + s * STAGE + QT_BYTES + row * (DSC_COMPACT_PITCH * 4);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float4 w = lds128(wt4 + 16 * c);
            if (!bpos) {
              fmul2(w.x, w.y, w.x, w.y, cbw, cbw);
              fmul2(w.z, w.w, w.z, w.w, cbw, cbw);
            }
            ffma2(sc[4 * c], sc[4 * c + 1], sc[4 * c], sc[4 * c + 1], ca, ca, w.x, w.y);
            ffma2(sc[4 * c + 2], sc[4 * c + 3], sc[4 * c + 2], sc[4 * c + 3], ca, ca, w.z, w.w);
          }
          sc[77] = sc[78] = sc[79] = -INFINITY;  // pad keys
          // slots 16..79 carry no weight: their y is a * s, so the row max is taken on the raw scores (a > 0) and the
          // scaling folds into the single FFMA that forms the exponent
          float my[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, mr[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int c = 0; c < 16; ++c) my[c & 3] = fmaxf(my[c & 3], sc[c]);
#pragma unroll
          for (int c = 16; c < 80; ++c) mr[c & 3] = fmaxf(mr[c & 3], sc[c]);
          const float m = fmaxf(fmaxf(fmaxf(my[0], my[1]), fmaxf(my[2], my[3])), ca * fmaxf(fmaxf(mr[0], mr[1]), fmaxf(mr[2], mr[3])));
          const float nb = -ce * m, k2 = ce * ca;
          // P V of the previous item was issued when that item published P: long finished after the S read, W and row max
          // above; its completion proves P(previous) has been consumed, so this item's P may go in
          X3_PH(3);
          if constexpr (!X3_LATE_PVWAIT) pv_done();
          X3_PH(4);
          if constexpr (C::TURNS > 0) {
            const uint32_t seq = static_cast<uint32_t>(j);
            if (lane == 0) {
              long long t0 = 0;
              uint32_t spins = 0;
              while (static_cast<int>(seq - lds32_volatile(my_turn)) >= C::TURNS) {  // at most TURNS items ahead of the oldest unfinished one
                __nanosleep(20);
                if (++spins == 256) t0 = clock64();
                if (spins > 256 && clock64() - t0 > (1ll << 32)) __trap();
              }
            }
            __syncwarp();
          }
          X3_TRACE(19);
          X3_PH(5);
          uint32_t pw[40];
#pragma unroll
          for (int c = 0; c < 39; ++c) {
            float e0, e1;
            if (c < 8) ffma2(e0, e1, sc[2 * c], sc[2 * c + 1], ce, ce, nb, nb);
            else ffma2(e0, e1, sc[2 * c], sc[2 * c + 1], k2, k2, nb, nb);
            if (X3_POLY > 0 && c % (X3_POLY > 0 ? X3_POLY : 1) == X3_POLY - 1 && c < 38) {
              float p0, p1;
              exp2_fma2(p0, p1, e0, e1);  // this pair on the FMA / ALU pipes instead of the MUFU pipe
              pw[c] = Mma<T>::pack(p0, p1);
            } else {
              pw[c] = c < 38 ? Mma<T>::pack(ex2_approx(e0), ex2_approx(e1)) : Mma<T>::pack(ex2_approx(e0), 0.f);
            }
            if (c == 23) {  // keys 0..47 are done: first part of P
              if constexpr (X3_LATE_PVWAIT) pv_done();  // (only now is P(previous) about to be overwritten)
              tmem_st_x16(tw + P_COL, pw);
              tmem_st_x8(tw + P_COL + 16, pw + 16);
            }
          }
          pw[39] = 0u;  // pad keys 78, 79
          tmem_st_x16(tw + P_COL + 24, pw + 24);  // (volatile, reads pw[24..39]: every exponential above precedes it)
          if constexpr (C::TURNS > 0) {
            __syncwarp();
            if (lane == 0) red_add_shared(my_turn, 1u);  // one more M phase complete: the next warp in line may start
          }
          X3_PH(6);
          tc_wait_st();
          tc_fence_before();
          mbar_arrive(b_prdy + 8 * g);
          X3_TRACE(16);
          X3_PH(7);
          pend = true;
        }
      }
      X3_PH_FLUSH;
    }
    if constexpr (STATS) {
      // CTA partial in a fixed order (warp shuffle tree, then the consumer warps serially) -> workspace; the last CTA folds
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
        dsq += __shfl_xor_sync(0xffffffffu, dsq, o);
      }
      double* red = reinterpret_cast<double*>(smem);  // the K image is dead: every Q K^T of this CTA has been consumed
      asm volatile("bar.sync 1, %0;" ::"n"(CONSUMERS) : "memory");
      if (lane == 0) {
        red[warp] = dsum;
        red[4 * NWG + warp] = dsq;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(CONSUMERS) : "memory");
      double* partials = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(p.ws) + kWorkspaceHeader);
      if (warp == 0) {
        unsigned int last = 0;
        if (lane == 0) {
          double a = 0.0, b = 0.0;
          for (int w = 0; w < 4 * NWG; ++w) {
            a += red[w];
            b += red[4 * NWG + w];
          }
          if (FUSED || p.handoff) {
            // the value IS the message: two 8-byte stores (atomic each), bit 0 set so that a published value is never the
            // all-zero "empty" pattern (1 ulp of fp64); no fence, no ticket -- pass 2 polls the slots and folds them
            unsigned long long* slots = reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(p.ws) + kHandoffOffset);
            __stcg(slots + 2 * blockIdx.x, static_cast<unsigned long long>(__double_as_longlong(a)) | 1ull);
            __stcg(slots + 2 * blockIdx.x + 1, static_cast<unsigned long long>(__double_as_longlong(b)) | 1ull);
          } else {
            // (same bit 0 as the handoff form: both protocols fold the same values in the same order -> the same std)
            partials[2 * blockIdx.x] = __longlong_as_double(__double_as_longlong(a) | 1ll);
            partials[2 * blockIdx.x + 1] = __longlong_as_double(__double_as_longlong(b) | 1ll);
            __threadfence();
            last = atomicAdd(&p.ws->ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
          }
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
          __threadfence();
          finalize_stats(p, partials, lane);
        }
      }
    }
  }

  X3_TRACE(9);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  X3_CTA_TIME(1);
  if constexpr (FUSED && STATS) {
    // pass 2 re-uses this shared memory with another layout: the barrier objects must be invalidated before their bytes
    // become tile data (pass 2 writes there only after its own CTA-wide barrier, i.e. after this warp is done)
    if (warp == SW0 + 2) {
      for (int idx = lane; idx < N_BARS; idx += 32) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(bars + 8 * idx) : "memory");
    }
  } else {
    if (warp == SW0 + 1) {
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
  }
  return tmem_base;
}

template <typename T, int HD, bool STATS>
__global__ void __launch_bounds__(Cfg<HD>::threads(STATS), 1)
xattn_x3_kernel(const XattnParams p, const __grid_constant__ CUtensorMap tm_qa, const __grid_constant__ CUtensorMap tm_qb,
                const __grid_constant__ CUtensorMap tm_qp, const __grid_constant__ CUtensorMap tm_oa,
                const __grid_constant__ CUtensorMap tm_ob) {
  x3_phase<T, HD, STATS, false>(p, tm_qa, tm_qb, tm_qp, tm_oa, tm_ob, 0u);
}

// Both passes of one call in ONE cooperative launch (every CTA resident: pass 2 polls the handoff slots that all pass-1 phases
// fill): no second launch, no second prologue (TMEM allocation, descriptor fetches), pass 2's first loads leave the moment
// the CTA's own pass 1 is done while the other CTAs are still streaming.
template <typename T, int HD>
__global__ void __launch_bounds__(Cfg<HD>::threads(false), 1)
xattn_x3_fused_kernel(const XattnParams p, const __grid_constant__ CUtensorMap tm_qa, const __grid_constant__ CUtensorMap tm_qb,
                      const __grid_constant__ CUtensorMap tm_qp, const __grid_constant__ CUtensorMap tm_oa,
                      const __grid_constant__ CUtensorMap tm_ob) {
  const uint32_t tmem = x3_phase<T, HD, true, true>(p, tm_qa, tm_qb, tm_qp, tm_qa, tm_qb, 0u);
  x3_phase<T, HD, false, true>(p, tm_qa, tm_qb, tm_qp, tm_oa, tm_ob, tmem);
}

// ---- K / V^T image ------------------------------------------------------------------------------
// One block per (batch, head group) writes the image the kernels above bulk-copy into shared memory record by record: HPT
// head records of REC_BYTES = K_HEAD + VT_HEAD (HPT heads of HD columns; NKC = 2 * KSTEPS; ON = HD + 1 rounded up to 16):
//   K_h  : [chunk 0..NKC-1][key slot 0..79][16 B]   chunk c = columns 16*floor(HD*h/16) + 8c .. +7 of the head GROUP (the
//          16-column blocks that cover the head), zeros where those columns belong to a neighbouring head (HD = 40 only);
//          key slots >= S = zeros (exact zero scores)
//   V^T_h: [key chunk 0..9][row d 0..ON-1][8 key slots x 2 B]   row HD = ones (softmax row sum), rows above = zeros
// key slot -> key: the n_active weighted columns of the compact region map first (ascending), then every other key in
// order (softmax and P V do not depend on the key order; pass 1 sums over all keys).
struct ActiveCols {
  int n;
  int col[DSC_MAX_COMPACT_COLS];
};

template <typename T, int HD>
__global__ void __launch_bounds__(256) x3_prepare_kv_kernel(const T* __restrict__ k, const T* __restrict__ v, long long k_sb,
                                                            long long k_ss, long long v_sb, long long v_ss, int S, int n_hg,
                                                            const ActiveCols ac, unsigned char* __restrict__ image) {
  using C = Cfg<HD>;
  constexpr int D = C::D, HPT = C::HPT, NKC = C::NKC, ON = C::ON;
  __shared__ int perm[DSC_MAX_KEYS];
  const int tid = threadIdx.x;
  const int b = blockIdx.x / n_hg, hg = blockIdx.x - b * n_hg;
  if (tid < DSC_MAX_KEYS) {
    int key = tid;
    if (tid < ac.n) {
      key = ac.col[tid];
    } else {
      key = tid - ac.n;
      for (int j = 0; j < ac.n; ++j) key += (ac.col[j] <= key) ? 1 : 0;
    }
    perm[tid] = min(key, DSC_MAX_KEYS - 1);
  }
  __syncthreads();
  unsigned char* img = image + static_cast<size_t>(blockIdx.x) * C::IMG_BYTES;
  const T* kb = k + b * k_sb + hg * GW;
  const T* vb = v + b * v_sb + hg * GW;
  for (int e = tid; e < HPT * NKC * DSC_MAX_KEYS; e += 256) {
    const int slot = e % DSC_MAX_KEYS, hc = e / DSC_MAX_KEYS, c = hc % NKC, h = hc / NKC;
    // chunk c of head h = columns 16 * floor(HD h / 16) + 8c .. +7 of the head group; zero outside the head's own columns
    const int col = ((h * D) >> 4) * 16 + c * 8;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (col >= h * D && col < (h + 1) * D && slot < S) val = *reinterpret_cast<const uint4*>(kb + perm[slot] * k_ss + col);
    *reinterpret_cast<uint4*>(img + static_cast<size_t>(h) * C::REC_BYTES + static_cast<size_t>(c * DSC_MAX_KEYS + slot) * 16) = val;
  }
  const unsigned short one = std::is_same<T, __half>::value ? 0x3C00u : 0x3F80u;
  for (int e = tid; e < HPT * 10 * ON; e += 256) {
    const int d = e % ON, hk = e / ON, kc = hk % 10, vh = hk / 10;
    unsigned short w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int slot = kc * 8 + j;
      unsigned short x = 0;
      if (slot < S) {
        if (d < D) x = *reinterpret_cast<const unsigned short*>(vb + perm[slot] * v_ss + vh * D + d);
        else if (d == D) x = one;
      }
      w[j] = x;
    }
    uint4 val;
    val.x = w[0] | (static_cast<uint32_t>(w[1]) << 16);
    val.y = w[2] | (static_cast<uint32_t>(w[3]) << 16);
    val.z = w[4] | (static_cast<uint32_t>(w[5]) << 16);
    val.w = w[6] | (static_cast<uint32_t>(w[7]) << 16);
    *reinterpret_cast<uint4*>(img + static_cast<size_t>(vh) * C::REC_BYTES + C::K_HEAD + static_cast<size_t>(kc * ON + d) * 16) = val;
  }
}

// [B, L, cols] 16-bit tensor with element strides (sb, sl, 1) -> boxes of box_cols columns x 128 rows, no swizzle
static bool make_map_plain(CUtensorMap* m, const void* base, int cols, int L, int B, long long sl, long long sb, int box_cols) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  cuuint64_t gdim[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(L), static_cast<cuuint64_t>(B)};
  cuuint64_t gstr[2] = {static_cast<cuuint64_t>(sl) * 2, static_cast<cuuint64_t>(B > 1 ? sb : sl * L) * 2};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(box_cols), ROWS, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<void*>(base), gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// [B, L, cols] 16-bit tensor -> boxes of 64 columns (SWIZZLE_128B) or 32 columns (SWIZZLE_64B) x 128 rows
static bool make_map_sw(CUtensorMap* m, const void* base, int cols, int L, int B, long long sl, long long sb, int box_cols) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  cuuint64_t gdim[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(L), static_cast<cuuint64_t>(B)};
  cuuint64_t gstr[2] = {static_cast<cuuint64_t>(sl) * 2, static_cast<cuuint64_t>(B > 1 ? sb : sl * L) * 2};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(box_cols), ROWS, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <typename T, int HD, bool STATS>
static cudaError_t launch(XattnParams p, cudaStream_t st) {
  using C = Cfg<HD>;
  constexpr int smem = STATS ? C::STATS_SMEM : C::FWD_SMEM;
  static thread_local int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(xattn_x3_kernel<T, HD, STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured_dev = dev;
  }
  CUtensorMap tm_qa, tm_qb, tm_qp, tm_oa, tm_ob;
  if (!make_map_sw(&tm_qa, p.q, p.H * HD, p.L, p.B, p.q_sl, p.q_sb, 64)) return cudaErrorInvalidValue;
  if (!make_map_sw(&tm_qb, p.q, p.H * HD, p.L, p.B, p.q_sl, p.q_sb, 32)) return cudaErrorInvalidValue;
  if (!make_map_plain(&tm_qp, p.q, p.H * HD, p.L, p.B, p.q_sl, p.q_sb, GW)) return cudaErrorInvalidValue;  // L2 prefetch only
  if (STATS) {
    tm_oa = tm_qa;
    tm_ob = tm_qb;
  } else {
    if (!make_map_sw(&tm_oa, p.out, p.H * HD, p.L, p.B, p.o_sl, p.o_sb, 64)) return cudaErrorInvalidValue;
    if (!make_map_sw(&tm_ob, p.out, p.H * HD, p.L, p.B, p.o_sl, p.o_sb, 32)) return cudaErrorInvalidValue;
  }
  p.n_hg = p.H / C::HPT;
  p.n_sl = (p.L + ROWS - 1) / ROWS;
  p.total = static_cast<long long>(p.B) * p.n_hg * p.n_sl;
  if (p.total >= (1ll << 31)) return cudaErrorInvalidValue;
  const int sms = sm_count_cached();
  const int grid = static_cast<int>(p.total < sms ? p.total : sms);
  p.tiles_q = static_cast<uint32_t>(p.total / grid);
  p.tiles_r = static_cast<uint32_t>(p.total % grid);
  p.div_nsl = make_fastdiv(static_cast<uint32_t>(p.n_sl));
  p.div_nhg = make_fastdiv(static_cast<uint32_t>(p.n_hg));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(C::threads(STATS));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = config().no_pdl ? 0 : 1;  // pass 2 may overlap the tail of pass 1 (and pass 1 its predecessor's)
  return cudaLaunchKernelEx(&cfg, xattn_x3_kernel<T, HD, STATS>, p, tm_qa, tm_qb, tm_qp, tm_oa, tm_ob);
}

template <typename T, int HD>
static cudaError_t launch_fused(XattnParams p, cudaStream_t st) {
  using C = Cfg<HD>;
  constexpr int smem = C::FWD_SMEM > C::STATS_SMEM ? C::FWD_SMEM : C::STATS_SMEM;
  static thread_local int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(xattn_x3_fused_kernel<T, HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured_dev = dev;
  }
  CUtensorMap tm_qa, tm_qb, tm_qp, tm_oa, tm_ob;
  if (!make_map_sw(&tm_qa, p.q, p.H * HD, p.L, p.B, p.q_sl, p.q_sb, 64)) return cudaErrorInvalidValue;
  if (!make_map_sw(&tm_qb, p.q, p.H * HD, p.L, p.B, p.q_sl, p.q_sb, 32)) return cudaErrorInvalidValue;
  if (!make_map_plain(&tm_qp, p.q, p.H * HD, p.L, p.B, p.q_sl, p.q_sb, GW)) return cudaErrorInvalidValue;  // L2 prefetch only
  if (!make_map_sw(&tm_oa, p.out, p.H * HD, p.L, p.B, p.o_sl, p.o_sb, 64)) return cudaErrorInvalidValue;
  if (!make_map_sw(&tm_ob, p.out, p.H * HD, p.L, p.B, p.o_sl, p.o_sb, 32)) return cudaErrorInvalidValue;
  p.n_hg = p.H / C::HPT;
  p.n_sl = (p.L + ROWS - 1) / ROWS;
  p.total = static_cast<long long>(p.B) * p.n_hg * p.n_sl;
  if (p.total >= (1ll << 31)) return cudaErrorInvalidValue;
  const int sms = sm_count_cached();
  const int grid = static_cast<int>(p.total < sms ? p.total : sms);
  p.tiles_q = static_cast<uint32_t>(p.total / grid);
  p.tiles_r = static_cast<uint32_t>(p.total % grid);
  p.div_nsl = make_fastdiv(static_cast<uint32_t>(p.n_sl));
  p.div_nhg = make_fastdiv(static_cast<uint32_t>(p.n_hg));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(C::threads(false));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeCooperative;  // pass 2 waits for values every pass-1 phase publishes: all CTAs resident
  attr[0].val.cooperative = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = config().no_pdl ? 1 : 2;
#ifdef DSC_CTATIME  // measurement builds only: DSC_X3_NOCOOP=1 drops the cooperative attribute (what does it cost per launch?
                    // nothing: 41.0 vs 41.0 us by events, 37.5 vs 37.5 us back to back in a graph, profiles/r2_call_nocoop_ab.jsonl)
  static const bool nocoop = getenv("DSC_X3_NOCOOP") != nullptr;
  if (nocoop) {
    cfg.attrs = attr + 1;
    cfg.numAttrs = 1;
  }
#endif
  cudaError_t e = cudaLaunchKernelEx(&cfg, xattn_x3_fused_kernel<T, HD>, p, tm_qa, tm_qb, tm_qp, tm_oa, tm_ob);
  if (e != cudaSuccess && cfg.numAttrs == 2) {  // a driver that refuses the combination: cooperative only
    (void)cudaGetLastError();
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, xattn_x3_fused_kernel<T, HD>, p, tm_qa, tm_qb, tm_qp, tm_oa, tm_ob);
  }
  return e;
}

template <typename T, int HD>
static cudaError_t prepare(const void* k, const void* v, long long k_sb, long long k_ss, long long v_sb, long long v_ss, int B, int H,
                           int S, const ActiveCols& ac, void* image, cudaStream_t st) {
  const int n_hg = H / Cfg<HD>::HPT;
  x3_prepare_kv_kernel<T, HD><<<B * n_hg, 256, 0, st>>>(static_cast<const T*>(k), static_cast<const T*>(v), k_sb, k_ss, v_sb, v_ss,
                                                        S, n_hg, ac, static_cast<unsigned char*>(image));
  return cudaGetLastError();
}

}  // namespace x3

// Prepared-K/V path: head dims 40 / 80 / 160 (4 / 2 / 1 heads per 160-column group), the 77 keys of one CLIP window
bool x3_supports(int H, int D, int S) {
  if (S != 77 || H <= 0) return false;
  return (D == 40 && H % 4 == 0) || (D == 80 && H % 2 == 0) || D == 160;
}

int x3_tile_count_per_batch_row(int D) { return D == 40 ? 4 : D == 80 ? 2 : 1; }

size_t x3_image_bytes(int B, int H, int D) {
  const size_t per = D == 40 ? x3::Cfg<40>::IMG_BYTES : D == 80 ? x3::Cfg<80>::IMG_BYTES : x3::Cfg<160>::IMG_BYTES;
  return static_cast<size_t>(B) * (H / x3_tile_count_per_batch_row(D)) * per;
}

#define X3_DISPATCH(FN, ...)                                                                                    \
  (dtype == DSC_DTYPE_F16                                                                                       \
       ? (D == 40 ? FN<__half, 40>(__VA_ARGS__) : D == 80 ? FN<__half, 80>(__VA_ARGS__) : FN<__half, 160>(__VA_ARGS__)) \
       : (D == 40 ? FN<__nv_bfloat16, 40>(__VA_ARGS__)                                                          \
                  : D == 80 ? FN<__nv_bfloat16, 80>(__VA_ARGS__) : FN<__nv_bfloat16, 160>(__VA_ARGS__)))

cudaError_t run_prepare_kv_x3(const void* k, const void* v, long long k_sb, long long k_ss, long long v_sb, long long v_ss, int B,
                              int H, int D, int S, int n_active, const int* cols, int dtype, void* image, cudaStream_t st) {
  x3::ActiveCols ac{};
  ac.n = n_active;
  for (int j = 0; j < n_active; ++j) ac.col[j] = cols[j];
  return X3_DISPATCH(x3::prepare, k, v, k_sb, k_ss, v_sb, v_ss, B, H, S, ac, image, st);
}

namespace x3 {
template <typename T, int HD>
static cudaError_t launch_stats(const XattnParams& p, cudaStream_t st) { return launch<T, HD, true>(p, st); }
template <typename T, int HD>
static cudaError_t launch_forward(const XattnParams& p, cudaStream_t st) { return launch<T, HD, false>(p, st); }
}  // namespace x3

cudaError_t run_stats_x3(const XattnParams& p, int D, int dtype, cudaStream_t st) { return X3_DISPATCH(x3::launch_stats, p, st); }
cudaError_t run_fused_x3(const XattnParams& p, int D, int dtype, cudaStream_t st) { return X3_DISPATCH(x3::launch_fused, p, st); }
cudaError_t run_forward_x3(const XattnParams& p, int D, int dtype, cudaStream_t st) { return X3_DISPATCH(x3::launch_forward, p, st); }

#ifdef DSC_TRACE
extern "C" int dsc_debug_x3_trace(long long* out /*HOST 2*8*1024*2*/, int* counts /*HOST 2*8*/, unsigned long long* cta /*HOST 2*160*2*/) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, x3::g_x3_trace, sizeof(long long) * 2 * 8 * 1024 * 2);
  cudaMemcpyFromSymbol(counts, x3::g_x3_trace_n, sizeof(int) * 16);
  cudaMemcpyFromSymbol(cta, x3::g_x3_cta, sizeof(unsigned long long) * 2 * 160 * 2);
  int z[16] = {0};
  cudaMemcpyToSymbol(x3::g_x3_trace_n, z, sizeof(z));
  return 0;
}
#endif

#ifdef DSC_CTATIME
extern "C" int dsc_debug_x3_cta(unsigned long long* cta /*HOST 2*160*2*/) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(cta, x3::g_x3_cta, sizeof(unsigned long long) * 2 * 160 * 2);
  return 0;
}
#endif

#ifdef DSC_PHASE
extern "C" int dsc_debug_x3_phase(unsigned int* out /*HOST 4*12*8*/) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, x3::g_x3_phase, sizeof(unsigned int) * 4 * 12 * 8);
  return 0;
}
#endif

#ifdef DSC_WATCHDOG
extern "C" int dsc_debug_x3_watchdog(unsigned int* out /*HOST 5*/) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, x3::g_x3_info, 16);
  cudaMemcpyFromSymbol(out + 4, x3::g_x3_abort, 4);
  unsigned int z[5] = {0, 0, 0, 0, 0};
  cudaMemcpyToSymbol(x3::g_x3_info, z, 16);
  cudaMemcpyToSymbol(x3::g_x3_abort, z, 4);
  return 0;
}
#endif

}  // namespace dsc
