// Region-masked cross-attention for SD-1.5 cross-attention layers on B200 (sm_100a).
//
//   pass 1  xattn_stats_kernel    a = scale*QK^T over the whole call -> sum / sum-of-squares -> std
//   pass 2  xattn_forward_kernel  recompute QK^T, + beta*W, softmax (registers + quad shuffles), PV
//
// Replaces scaled_dot_product_attention_regionstate (reference source/modules/attention_modify.py:74-103)
// with weight_func = w*sigma*qk.std() (reference source/app.py:1004).
//
// Data layout in HBM (no copies, exactly what the reference processor produces, attention_modify.py:471-474):
//   Q, O : [B, L, H*D]  (heads are column groups of one row)       K, V : [B, S, H*D]      W : fp32 [Bw, L, S]
//
// Decomposition.  The call is cut into "slices" of 16 query rows x one head-group (G heads, G*D = 320
// or 256 columns = 640/512 contiguous bytes per row).  A persistent CTA (one per SM, 8 warps) owns a
// contiguous range of slices of the (batch, head-group)-major slice list; the K and V head-group of the
// current batch row stay resident in shared memory, each warp streams its own slices:
//   TMA bulk copies (cp.async.bulk, one per row -> padded pitch, ldmatrix conflict-free) + mbarrier,
//   QK^T and PV on the tensor cores (mma.sync m16n8k16 / m16n8k8, fp32 accumulate),
//   the 77-key score rows live in registers (10 n-tiles x 4), softmax row reductions are quad shuffles,
//   the W tile is read once per slice and reused by all heads of the group,
//   O overwrites the warp's Q slice in shared memory and leaves through TMA bulk stores.
#include <cuda.h>  // CUtensorMap (types only; the encoder comes from cudaGetDriverEntryPoint)
#include <stdlib.h>

#include <type_traits>

#include "dsc_device.cuh"
#include "dsc_internal.h"

namespace dsc {

constexpr float kLog2e = 1.4426950408889634f;

// TMA tensor-map box load.  The Q / K / V tensors are described as [B, rows, H*D/2] arrays of 32-bit words so that one
// box of (G*D/2 + 4) words x R rows lands in shared memory as R rows of PITCH = G*D*2 + 16 bytes -- the padded,
// ldmatrix-conflict-free layout -- with ONE TMA instruction (a 1-D bulk copy per row costs ~30 ns of TMA issue each,
// profiles/r1_tma_copy_rate.jsonl: 154 row copies for K and V alone were 4.6 us of every CTA's prologue).  Rows beyond
// the tensor (keys >= S, queries >= L) and words beyond H*D are zero-filled by the TMA.
__device__ __forceinline__ void tma_box_load(uint32_t dst, const CUtensorMap* m, int c0, int c1, int c2, uint32_t bar,
                                             uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, "
      "%5}], [%2], %6;" ::"r"(dst),
      "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
      : "memory");
}

template <int D>
struct Tile {
  static_assert(D == 40 || D == 64 || D == 80 || D == 128 || D == 160, "unsupported head dim");
  static constexpr int G = (D == 40) ? 8 : (D == 64) ? 5 : (D == 80) ? 4 : 2;  // heads per group
  static constexpr int GW = G * D;                                             // columns per group
  static constexpr int PITCH = GW * 2 + 16;  // bytes; == 16 (mod 128) -> ldmatrix rows hit distinct banks
  static constexpr int KV_ROWS = 80;         // keys padded to 10 n-tiles
  static constexpr int NT = 10;
  static constexpr int WARPS = 8;
  static constexpr int ROWS = 16;  // query rows per slice (one m16 tile per warp)
  static constexpr int KV_BYTES = KV_ROWS * PITCH;
  static constexpr int QS_BYTES = ROWS * PITCH;
  static constexpr int WS_BYTES = ROWS * KV_ROWS * 4;
  static constexpr int DCH = (D > 80) ? D / 2 : D;  // PV is done in column chunks of <= 80
  static constexpr int NCH = D / DCH;
  static constexpr int ND = DCH / 8;
  static_assert((GW * 2) % 128 == 0, "pitch rule");
  // forward: K + V + 8 x (Q slice + W slice) + 9 mbarriers
  static constexpr int FWD_SMEM = 2 * KV_BYTES + WARPS * (QS_BYTES + WS_BYTES) + 128;
  // stats: K + 8 x 2 Q slices + 17 mbarriers
  static constexpr int STATS_SMEM = KV_BYTES + WARPS * 2 * QS_BYTES + 256;
};

// ---------------------------------------------------------------------------------------------
// S tile = Q_h K_h^T for one warp: 16 rows x 80 keys, fp32 accumulators in the mma C layout
//   acc[j][0..1] -> row g,   keys 8j+2t, 8j+2t+1        acc[j][2..3] -> row g+8      (g = lane/4, t = lane%4)
// q_base / k_base: shared addresses of row 0 of the slice / of K, already offset to the head's columns.
template <typename T, int D>
__device__ __forceinline__ void qk_tile(float (&acc)[10][4], uint32_t q_base, uint32_t k_base, int lane) {
  constexpr int PITCH = Tile<D>::PITCH;
#pragma unroll
  for (int j = 0; j < 10; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
  // A (x4): lanes 0-7 rows 0-7 | 8-15 rows 8-15 | 16-23 rows 0-7, +8 cols | 24-31 rows 8-15, +8 cols
  const uint32_t a_addr = q_base + ((lane & 7) + ((lane >> 3) & 1) * 8) * PITCH + (lane >> 4) * 16;
  // B (x4): lanes 0-7 keys 0-7 | 8-15 keys 0-7, +8 cols | 16-23 keys 8-15 | 24-31 keys 8-15, +8 cols
  const uint32_t b_addr = k_base + ((lane & 7) + ((lane >> 4) & 1) * 8) * PITCH + ((lane >> 3) & 1) * 16;
#pragma unroll
  for (int ks = 0; ks < D / 16; ++ks) {
    uint32_t a[4];
    ldsm_x4(a, a_addr + ks * 32);
#pragma unroll
    for (int jp = 0; jp < 5; ++jp) {
      uint32_t b[4];
      ldsm_x4(b, b_addr + jp * 16 * PITCH + ks * 32);
      Mma<T>::k16(acc[2 * jp], a, b[0], b[1]);
      Mma<T>::k16(acc[2 * jp + 1], a, b[2], b[3]);
    }
  }
  if constexpr (D % 16 == 8) {  // k8 tail (D = 40): no zero padding of the contraction dim
    constexpr int k0 = (D / 16) * 16;
    uint32_t a2[2];
    ldsm_x2(a2, q_base + (lane & 15) * PITCH + k0 * 2);
#pragma unroll
    for (int q4 = 0; q4 < 2; ++q4) {
      uint32_t b[4];
      ldsm_x4(b, k_base + (q4 * 32 + lane) * PITCH + k0 * 2);
#pragma unroll
      for (int i = 0; i < 4; ++i) Mma<T>::k8(acc[4 * q4 + i], a2[0], a2[1], b[i]);
    }
    uint32_t b2[2];
    ldsm_x2(b2, k_base + (64 + (lane & 15)) * PITCH + k0 * 2);
    Mma<T>::k8(acc[8], a2[0], a2[1], b2[0]);
    Mma<T>::k8(acc[9], a2[0], a2[1], b2[1]);
  }
}

// O chunk = P V_h[:, chunk] : 16 rows x (ND*8) columns.  pa = P in the mma A layout (5 k-steps of 16 keys).
template <typename T, int D>
__device__ __forceinline__ void pv_tile(float (&o)[Tile<D>::ND][4], const uint32_t (&pa)[5][4], uint32_t v_base,
                                        int lane) {
  constexpr int PITCH = Tile<D>::PITCH;
  constexpr int ND = Tile<D>::ND;
#pragma unroll
  for (int n = 0; n < ND; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
  // B (x4.trans): lanes 0-7 keys 0-7 | 8-15 keys 8-15 | 16-23 keys 0-7, +8 cols | 24-31 keys 8-15, +8 cols
  const uint32_t v_addr = v_base + ((lane & 7) + ((lane >> 3) & 1) * 8) * PITCH + (lane >> 4) * 16;
#pragma unroll
  for (int kk = 0; kk < 5; ++kk) {
#pragma unroll
    for (int np = 0; np < ND / 2; ++np) {
      uint32_t b[4];
      ldsm_x4_t(b, v_addr + kk * 16 * PITCH + np * 32);
      Mma<T>::k16(o[2 * np], pa[kk], b[0], b[1]);
      Mma<T>::k16(o[2 * np + 1], pa[kk], b[2], b[3]);
    }
    if constexpr (ND % 2 == 1) {
      uint32_t b2[2];
      ldsm_x2_t(b2, v_base + (kk * 16 + (lane & 15)) * PITCH + (ND - 1) * 16);
      Mma<T>::k16(o[ND - 1], pa[kk], b2[0], b2[1]);
    }
  }
}

// Segment bookkeeping shared by both passes: the slice list is (batch, head-group)-major.
struct Seg {
  int b, hg, nheads;
  long long end;  // first slice index after this segment, clipped to the CTA's range
};
template <int D>
__device__ __forceinline__ Seg seg_of(long long idx, long long cta_end, int n_sl, int n_hg, int H) {
  Seg s;
  const long long seg = idx / n_sl;
  s.b = static_cast<int>(seg / n_hg);
  s.hg = static_cast<int>(seg % n_hg);
  s.nheads = min(Tile<D>::G, H - s.hg * Tile<D>::G);
  const long long e = (seg + 1) * n_sl;
  s.end = e < cta_end ? e : cta_end;
  return s;
}

// CTA partial (raw, unscaled) in a fixed order -> workspace; the last CTA folds all partials.  Called by every thread
// of the CTA (NW warps).
// The fused single-launch kernel uses the publication as a grid barrier: the folding CTA bumps an epoch word behind the
// public header once the std is visible; everybody else waits for the bump (see xattn_fused_kernel).
constexpr int kEpochOffset = 48;  // bytes into the workspace header (the public dsc_xattn_stats_t ends at 40)
template <int NW>
__device__ __forceinline__ void publish_stats(const XattnParams& p, double dsum, double dsq, bool signal_epoch = false) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
    dsq += __shfl_xor_sync(0xffffffffu, dsq, o);
  }
  __shared__ double red[2 * NW];
  __shared__ unsigned int s_last;
  if (lane == 0) {
    red[warp] = dsum;
    red[NW + warp] = dsq;
  }
  __syncthreads();
  double* partials = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(p.ws) + kWorkspaceHeader);
  if (tid == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < NW; ++w) {
      a += red[w];
      b += red[NW + w];
    }
    const unsigned int slot = p.chunk * gridDim.x + blockIdx.x;
    partials[2 * slot] = a;
    partials[2 * slot + 1] = b;
    __threadfence();
    const unsigned int t = atomicAdd(&p.ws->ticket, 1u);
    s_last = (t == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (s_last && warp == 0) {
    __threadfence();
    double a = 0.0, b = 0.0;
    const unsigned int n_fold = p.fold_chunks * gridDim.x;  // 0: an earlier key chunk of a long prompt, nothing to publish yet
    for (unsigned int i = lane; i < n_fold; i += 32) {  // fixed assignment + fixed tree = deterministic
      a += __ldcg(partials + 2 * i);
      b += __ldcg(partials + 2 * i + 1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == 0 && n_fold == 0) {
      __threadfence();
      p.ws->ticket = 0u;
    }
    if (lane == 0 && n_fold != 0) {
      const double sc = static_cast<double>(p.scale);
      const double n = p.n_total;
      const double sum = a * sc, sumsq = b * sc * sc;
      const double mean = sum / n;
      double var = (n > 1.0) ? (sumsq - sum * mean) / (n - 1.0) : nan("");
      if (var < 0.0) var = 0.0;
      p.ws->std_unbiased = static_cast<float>(sqrt(var));
      p.ws->mean = static_cast<float>(mean);
      p.ws->sum = sum;
      p.ws->sumsq = sumsq;
      p.ws->n = n;
      p.ws->n_partials = n_fold;
      p.ws->ticket = 0u;  // reusable without a memset (its next use is in a later kernel: no ordering needed here)
      __threadfence();    // the statistics above are visible before ...
      if (signal_epoch)   // ... the barrier of the single-launch kernel opens
        atomicAdd(reinterpret_cast<unsigned int*>(reinterpret_cast<unsigned char*>(p.ws) + kEpochOffset), 1u);
    }
  }
}

// =============================================================================================
// pass 1
// =============================================================================================
// MASKED: a = qk_scale * Q K^T + M (additive attention mask, attention_modify.py:39-70 / :84-91) -- the sums are taken
// over a itself (p.scale is 1 for the fold), M read per element through its (b, h, l) strides.
template <typename T, int D, bool MASKED = false>
__global__ void __launch_bounds__(256, 1)
xattn_stats_kernel(const XattnParams p, const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k) {
  using TL = Tile<D>;
  constexpr int PITCH = TL::PITCH;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // PDL: pass 2 may start its prologue early
  const uint32_t sK = smem_u32(smem);
  const uint32_t sQ = sK + TL::KV_BYTES + warp * 2 * TL::QS_BYTES;  // two stages per warp
  const uint32_t bars = sK + TL::KV_BYTES + TL::WARPS * 2 * TL::QS_BYTES;
  const uint32_t kbar = bars;
  const uint32_t qbar = bars + 8 + warp * 16;  // [stage]

  // zero the K pad rows (keys S..79): their scores are exactly 0 and drop out of both sums
  for (int i = tid; i < (TL::KV_ROWS - p.S) * (PITCH / 16); i += 256)
    reinterpret_cast<uint4*>(smem + p.S * PITCH)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    mbar_init(kbar, 1);
    for (int w = 0; w < TL::WARPS; ++w) {
      mbar_init(bars + 8 + w * 16, 1);
      mbar_init(bars + 8 + w * 16 + 8, 1);
    }
    fence_mbar_init();
  }
  fence_proxy_async();
  __syncthreads();

  const long long total = p.total;
  const long long cta_begin = total * blockIdx.x / gridDim.x;
  const long long cta_end = total * (blockIdx.x + 1) / gridDim.x;
  const uint64_t pol_q = policy_evict_last();  // Q is read again by pass 2: ask L2 to keep it
  const uint64_t pol_kv = policy_evict_last();

  double dsum = 0.0, dsq = 0.0;
  uint32_t it = 0, kphase = 0;

  for (long long idx = cta_begin; idx < cta_end;) {
    const Seg sg = seg_of<D>(idx, cta_end, p.n_sl, p.n_hg, p.H);
    __syncthreads();  // every warp is done with the previous K
    if (tid == 0) {
      mbar_arrive_expect_tx(kbar, TL::KV_BYTES);  // the whole 80-row box counts (rows >= S arrive as zeros)
      tma_box_load(sK, &tm_k, sg.hg * (TL::GW / 2), 0, sg.b, kbar, pol_kv);
    }

    auto issue = [&](long long sl_idx, uint32_t stage) {
      const int l0 = static_cast<int>(sl_idx % p.n_sl) * TL::ROWS;
      const int rows = min(TL::ROWS, p.L - l0);
      const uint32_t bar = qbar + stage * 8;
      (void)rows;
      if (lane == 0) {
        mbar_arrive_expect_tx(bar, TL::QS_BYTES);
        tma_box_load(sQ + stage * TL::QS_BYTES, &tm_q, sg.hg * (TL::GW / 2), l0, sg.b, bar, pol_q);
      }
    };

    const long long first = idx + warp;
    if (first < sg.end) issue(first, it & 1);
    bool k_ready = false;
    for (long long sl = first; sl < sg.end; sl += TL::WARPS) {
      if (sl + TL::WARPS < sg.end) issue(sl + TL::WARPS, (it + 1) & 1);
      mbar_wait(qbar + (it & 1) * 8, (it >> 1) & 1);
      if (!k_ready) {
        mbar_wait(kbar, kphase);
        k_ready = true;
      }
      const int l0 = static_cast<int>(sl % p.n_sl) * TL::ROWS;
      const int rows = min(TL::ROWS, p.L - l0);
      const uint32_t qs = sQ + (it & 1) * TL::QS_BYTES;
      const float m0 = ((lane >> 2) < rows) ? 1.f : 0.f;      // row g valid
      const float m1 = ((lane >> 2) + 8 < rows) ? 1.f : 0.f;  // row g+8 valid
      for (int h = 0; h < sg.nheads; ++h) {
        float acc[10][4];
        qk_tile<T, D>(acc, qs + h * D * 2, sK + h * D * 2, lane);
        float fs = 0.f, fq = 0.f;
        if constexpr (MASKED) {
          // pad keys and rows beyond L are not scores: they carry no mask value and drop out here
          const float* mb = p.mask + sg.b * p.m_sb + static_cast<long long>(sg.hg * TL::G + h) * p.m_sh +
                            static_cast<long long>(l0) * p.m_sl + p.m_col0;
#pragma unroll
          for (int j = 0; j < 10; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int col = 8 * j + 2 * (lane & 3) + (i & 1), row = (lane >> 2) + 8 * (i >> 1);
              float v = 0.f;
              if (row < rows && col < p.S) v = fmaf(acc[j][i], p.qk_scale, __ldg(mb + row * p.m_sl + col));
              fs += v;
              fq = fmaf(v, v, fq);
            }
        } else if (rows == TL::ROWS) {
#pragma unroll
          for (int j = 0; j < 10; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              fs += acc[j][i];
              fq = fmaf(acc[j][i], acc[j][i], fq);
            }
        } else {  // tail slice: rows beyond L hold stale shared memory
#pragma unroll
          for (int j = 0; j < 10; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float v = (i < 2 ? m0 : m1) != 0.f ? acc[j][i] : 0.f;
              fs += v;
              fq = fmaf(v, v, fq);
            }
        }
        dsum += static_cast<double>(fs);
        dsq += static_cast<double>(fq);
      }
      __syncwarp();  // all lanes have consumed this stage before it is refilled
      ++it;
    }
    kphase ^= 1;
    idx = sg.end;
  }

  publish_stats<TL::WARPS>(p, dsum, dsq);
}

// =============================================================================================
// pass 1 through the Gram identity (D = 40):   sum a^2 = scale^2 <Q_h^T Q_h, K_h^T K_h>_F ,  sum a = scale (sum_l q_l).(sum_s k_s)
//
// The scores are never formed: per (batch, head) a warp accumulates the 40 x 40 Gram matrix of its head's Q columns over
// all query rows (A = Q^T and B = Q are the SAME ldmatrix.trans fragments of the row-major tile: 3 ldmatrix.x4 feed
// 15 + 3 mma.sync per 16 rows), the column sums ride along as a product with a ones fragment, and at the end of the
// (batch, head-group) run the Gram matrix of K (80 zero-padded keys) is formed in the same fragment layout and the two
// are contracted element by element.  No per-score work at all: pass 1 becomes a pure stream of Q.
// 8 consumer warps = the 8 heads of a 320-column tile, warp 8 = TMA producer (64-row tiles, 3-stage ring).
#ifdef DSC_TRACE
__device__ unsigned long long g_gram_times[160][8];  // globaltimer (ns) per CTA at the stamps below
__device__ __forceinline__ unsigned long long gram_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define GRAM_STAMP(k) do { if ((threadIdx.x & 31) == 0 && (threadIdx.x >> 5) == ((k) == 2 || (k) == 3 || (k) == 4 ? 0 : 8) && blockIdx.x < 160) g_gram_times[blockIdx.x][k] = gram_now(); } while (0)
#else
#define GRAM_STAMP(k) do {} while (0)
#endif

template <int D>
struct GramTile {
  using TL = Tile<D>;
  static constexpr int ROWS = 64;
  static constexpr int STAGES = 3;
  static constexpr int QT_BYTES = ROWS * TL::PITCH;
  static constexpr int SMEM = TL::KV_BYTES + STAGES * QT_BYTES + 128;
  static constexpr int MT = (D + 15) / 16;  // m-tiles; rows >= D of the last one are a neighbour's columns and are ignored
  static constexpr int NT = D / 8;
  static constexpr int THREADS = 288;
};

template <typename T, int D>
__global__ void __launch_bounds__(GramTile<D>::THREADS, 1)
xattn_gram_stats_kernel(const XattnParams p, const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k) {
  using TL = Tile<D>;
  using GT = GramTile<D>;
  constexpr int PITCH = TL::PITCH, MT = GT::MT, NT = GT::NT, NST = GT::STAGES;
  static_assert(TL::G == 8 && 2 * MT >= NT, "one warp per head; the A fragments must cover every n-tile");
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  GRAM_STAMP(0);
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // PDL: pass 2 may start its prologue early
  const uint32_t sK = smem_u32(smem);
  const uint32_t sQ = sK + TL::KV_BYTES;
  const uint32_t bars = sQ + NST * GT::QT_BYTES;
  const uint32_t b_kfull = bars, b_kempty = bars + 8, b_full = bars + 16, b_empty = bars + 16 + 8 * NST;
  if (tid == 0) {
    mbar_init(b_kfull, 1);
    mbar_init(b_kempty, TL::WARPS);
    for (int s = 0; s < NST; ++s) {
      mbar_init(b_full + 8 * s, 1);
      mbar_init(b_empty + 8 * s, TL::WARPS);
    }
    fence_mbar_init();
  }
  fence_proxy_async();
  __syncthreads();

  const long long cta_begin = p.total * blockIdx.x / gridDim.x;
  const long long cta_end = p.total * (blockIdx.x + 1) / gridDim.x;
  double dsum = 0.0, dsq = 0.0;
  GRAM_STAMP(1);

  if (warp == TL::WARPS) {
    if (lane == 0) {  // ---- producer
      const uint64_t pol = policy_evict_last();  // Q is read again by pass 2, K by both
      uint32_t i = 0, run = 0;
      for (long long idx = cta_begin; idx < cta_end; ++run) {
        const Seg sg = seg_of<D>(idx, cta_end, p.n_sl, p.n_hg, p.H);
        if (run > 0) mbar_wait(b_kempty, (run - 1) & 1);
        mbar_arrive_expect_tx(b_kfull, TL::KV_BYTES);  // 80-row box, keys >= S arrive as zeros
        tma_box_load(sK, &tm_k, sg.hg * (TL::GW / 2), 0, sg.b, b_kfull, pol);
        for (long long t = idx; t < sg.end; ++t, ++i) {
          const uint32_t s = i % NST;
          if (i >= NST) mbar_wait(b_empty + 8 * s, ((i / NST) - 1) & 1);
          mbar_arrive_expect_tx(b_full + 8 * s, GT::QT_BYTES);  // rows >= L arrive as zeros
          tma_box_load(sQ + s * GT::QT_BYTES, &tm_q, sg.hg * (TL::GW / 2), static_cast<int>(t % p.n_sl) * GT::ROWS, sg.b,
                       b_full + 8 * s, pol);
        }
        idx = sg.end;
      }
    }
  } else {
    // ---- consumers: warp = head of the group
    const int h = warp, g = lane >> 2, t4 = lane & 3;
    const uint32_t ones = std::is_same<T, __half>::value ? 0x3C003C00u : 0x3F803F80u;
    // ldmatrix.x4.trans of m-tile mi at k-step ks: matrix q = lane / 8 -> rows (q / 2) * 8 + lane % 8, column block
    // 2 mi + (q & 1); the four results are the mma A fragment of Q^T, and the B fragments of n-tiles 2 mi, 2 mi + 1
    const uint32_t lane_off = (((lane >> 4) & 1) * 8 + (lane & 7)) * PITCH + (h * D + ((lane >> 3) & 1) * 8) * 2;
    float cq[MT][NT][4], cs[MT][4];
#pragma unroll
    for (int mi = 0; mi < MT; ++mi) {
      cs[mi][0] = cs[mi][1] = cs[mi][2] = cs[mi][3] = 0.f;
#pragma unroll
      for (int nj = 0; nj < NT; ++nj) cq[mi][nj][0] = cq[mi][nj][1] = cq[mi][nj][2] = cq[mi][nj][3] = 0.f;
    }
    uint32_t i = 0, run = 0;
    for (long long idx = cta_begin; idx < cta_end; ++run) {
      const Seg sg = seg_of<D>(idx, cta_end, p.n_sl, p.n_hg, p.H);
      const bool active = h < sg.nheads;
      for (long long t = idx; t < sg.end; ++t, ++i) {
        const uint32_t s = i % NST;
        mbar_wait(b_full + 8 * s, (i / NST) & 1);
        if (i == 0) GRAM_STAMP(2);
        if (active) {
          const uint32_t tile = sQ + s * GT::QT_BYTES + lane_off;
#pragma unroll
          for (int ks = 0; ks < GT::ROWS / 16; ++ks) {
            uint32_t x[MT][4];
#pragma unroll
            for (int mi = 0; mi < MT; ++mi) ldsm_x4_t(x[mi], tile + ks * 16 * PITCH + mi * 32);
#pragma unroll
            for (int mi = 0; mi < MT; ++mi) {
#pragma unroll
              for (int nj = 0; nj < NT; ++nj) Mma<T>::k16(cq[mi][nj], x[mi], x[nj >> 1][nj & 1], x[nj >> 1][2 + (nj & 1)]);
              Mma<T>::k16(cs[mi], x[mi], ones, ones);  // column sums of Q (every n column holds the same sum)
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(b_empty + 8 * s);
      }
      // ---- end of the (batch, head-group) run: contract with the Gram matrix of K
      GRAM_STAMP(3);
      mbar_wait(b_kfull, run & 1);
      if (active) {
        float acc2 = 0.f, acc1 = 0.f;
#pragma unroll
        for (int mi = 0; mi < MT; ++mi) {
          float ck[NT][4], csk[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int nj = 0; nj < NT; ++nj) ck[nj][0] = ck[nj][1] = ck[nj][2] = ck[nj][3] = 0.f;
#pragma unroll
          for (int ks = 0; ks < TL::KV_ROWS / 16; ++ks) {
            uint32_t x[MT][4];
#pragma unroll
            for (int m2 = 0; m2 < MT; ++m2) ldsm_x4_t(x[m2], sK + lane_off + ks * 16 * PITCH + m2 * 32);
#pragma unroll
            for (int nj = 0; nj < NT; ++nj) Mma<T>::k16(ck[nj], x[mi], x[nj >> 1][nj & 1], x[nj >> 1][2 + (nj & 1)]);
            Mma<T>::k16(csk, x[mi], ones, ones);
          }
          const bool v_lo = 16 * mi + g < D, v_hi = 16 * mi + g + 8 < D;
#pragma unroll
          for (int nj = 0; nj < NT; ++nj) {
            if (v_lo) acc2 += cq[mi][nj][0] * ck[nj][0] + cq[mi][nj][1] * ck[nj][1];
            if (v_hi) acc2 += cq[mi][nj][2] * ck[nj][2] + cq[mi][nj][3] * ck[nj][3];
          }
          if (t4 == 0) {
            if (v_lo) acc1 += cs[mi][0] * csk[0];
            if (v_hi) acc1 += cs[mi][2] * csk[2];
          }
          cs[mi][0] = cs[mi][1] = cs[mi][2] = cs[mi][3] = 0.f;
#pragma unroll
          for (int nj = 0; nj < NT; ++nj) cq[mi][nj][0] = cq[mi][nj][1] = cq[mi][nj][2] = cq[mi][nj][3] = 0.f;
        }
        dsq += static_cast<double>(acc2);
        dsum += static_cast<double>(acc1);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(b_kempty);
      idx = sg.end;
    }
  }
  GRAM_STAMP(4);
  __syncthreads();
  GRAM_STAMP(5);
  publish_stats<TL::WARPS + 1>(p, dsum, dsq);
  GRAM_STAMP(6);
}

// One 16-row slice, all heads of the group: logits = S*scale + beta*W (log2 domain), softmax over the keys, O = P V.
// Q_h columns of the slice (shared memory, qsm / sQ) are overwritten by O_h.  lse0: optional log2-sum-exp output of
// (first head of the group, first row of the slice); heads are L floats apart.
// MASKED: the additive attention mask of (first head of the group, first row of the slice, first key of the chunk) is at
// m0; heads are m_sh, rows m_sl floats apart (0 = broadcast); logits = S*scale + M + beta*W.
template <typename T, int D, bool MASKED = false>
__device__ __forceinline__ void softmax_pv_slice(int nheads, int S, float beta_l2, float scale_l2, uint32_t sQ, uint32_t sK,
                                                 uint32_t sV, const float* wsm, int wp, unsigned char* qsm, float* lse0, int L,
                                                 int rows, int lane, const float* m0 = nullptr, long long m_sh = 0,
                                                 long long m_sl = 0) {
  using TL = Tile<D>;
  constexpr int PITCH = TL::PITCH;
  constexpr int ND = TL::ND;
  const int g = lane >> 2, t = lane & 3;
  // beta*W in the accumulator layout, shared by all heads of the group; columns >= S get -inf
  float bw[10][4];
#pragma unroll
  for (int j = 0; j < 10; ++j) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int col = 8 * j + 2 * t + (i & 1);
      const int row = g + 8 * (i >> 1);
      bw[j][i] = (col < S) ? wsm[row * wp + col] * beta_l2 : -INFINITY;
    }
  }

  for (int h = 0; h < nheads; ++h) {
    float acc[10][4];
    qk_tile<T, D>(acc, sQ + h * D * 2, sK + h * D * 2, lane);
    // logits (log2 domain), row max
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 10; ++j) {
      acc[j][0] = fmaf(acc[j][0], scale_l2, bw[j][0]);
      acc[j][1] = fmaf(acc[j][1], scale_l2, bw[j][1]);
      acc[j][2] = fmaf(acc[j][2], scale_l2, bw[j][2]);
      acc[j][3] = fmaf(acc[j][3], scale_l2, bw[j][3]);
      if constexpr (MASKED) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int col = 8 * j + 2 * t + (i & 1), row = g + 8 * (i >> 1);
          if (row < rows && col < S) acc[j][i] = fmaf(__ldg(m0 + h * m_sh + row * m_sl + col), kLog2e, acc[j][i]);
        }
      }
      mx0 = fmaxf(mx0, fmaxf(acc[j][0], acc[j][1]));
      mx1 = fmaxf(mx1, fmaxf(acc[j][2], acc[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    // p = 2^(s - max), row sums, pack to the input dtype as the A operand of PV
    float sum0 = 0.f, sum1 = 0.f;
    uint32_t pa[5][4];
#pragma unroll
    for (int j = 0; j < 10; ++j) {
      const float p0 = ex2_approx(acc[j][0] - mx0), p1 = ex2_approx(acc[j][1] - mx0);
      const float p2 = ex2_approx(acc[j][2] - mx1), p3 = ex2_approx(acc[j][3] - mx1);
      sum0 += p0 + p1;
      sum1 += p2 + p3;
      pa[j >> 1][(j & 1) * 2 + 0] = Mma<T>::pack(p0, p1);
      pa[j >> 1][(j & 1) * 2 + 1] = Mma<T>::pack(p2, p3);
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    const float inv0 = 1.f / sum0, inv1 = 1.f / sum1;
    if (lse0 != nullptr && t == 0) {  // log2-sum-exp of the row's logits (key chunks of a long prompt are merged with it)
      float* lrow = lse0 + static_cast<long long>(h) * L;
      if (g < rows) lrow[g] = mx0 + log2f(sum0);
      if (g + 8 < rows) lrow[g + 8] = mx1 + log2f(sum1);
    }

    __syncwarp();  // Q_h has been consumed by every lane: its columns may now be overwritten by O_h
#pragma unroll
    for (int c = 0; c < TL::NCH; ++c) {
      float o[ND][4];
      pv_tile<T, D>(o, pa, sV + (h * D + c * TL::DCH) * 2, lane);
      unsigned char* orow0 = qsm + g * PITCH + (h * D + c * TL::DCH + 2 * t) * 2;
      unsigned char* orow1 = orow0 + 8 * PITCH;
#pragma unroll
      for (int n = 0; n < ND; ++n) {
        *reinterpret_cast<uint32_t*>(orow0 + n * 16) = Mma<T>::pack(o[n][0] * inv0, o[n][1] * inv0);
        *reinterpret_cast<uint32_t*>(orow1 + n * 16) = Mma<T>::pack(o[n][2] * inv1, o[n][3] * inv1);
      }
    }
  }
}

// =============================================================================================
// pass 2
// =============================================================================================
template <typename T, int D, bool MASKED = false>
__global__ void __launch_bounds__(256, 1)
xattn_forward_kernel(const XattnParams p, const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                     const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_w) {
  using TL = Tile<D>;
  constexpr int PITCH = TL::PITCH;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sK = smem_u32(smem);
  const uint32_t sV = sK + TL::KV_BYTES;
  const uint32_t sQ = sV + TL::KV_BYTES + warp * (TL::QS_BYTES + TL::WS_BYTES);
  const uint32_t sW = sQ + TL::QS_BYTES;
  const float* wsm = reinterpret_cast<const float*>(smem + 2 * TL::KV_BYTES + warp * (TL::QS_BYTES + TL::WS_BYTES) +
                                                     TL::QS_BYTES);
  unsigned char* qsm = smem + 2 * TL::KV_BYTES + warp * (TL::QS_BYTES + TL::WS_BYTES);
  const uint32_t bars = sK + 2 * TL::KV_BYTES + TL::WARPS * (TL::QS_BYTES + TL::WS_BYTES);
  const uint32_t kvbar = bars;
  const uint32_t qbar = bars + 8 + warp * 8;

  // zero the pad rows (keys S..79) of K and V once: padded scores are exactly 0 (then -inf through the
  // bias), padded P columns multiply zeros.  The row copies below never touch them.
  for (int i = tid; i < (TL::KV_ROWS - p.S) * (PITCH / 16); i += 256) {
    reinterpret_cast<uint4*>(smem + p.S * PITCH)[i] = make_uint4(0, 0, 0, 0);
    reinterpret_cast<uint4*>(smem + TL::KV_BYTES + p.S * PITCH)[i] = make_uint4(0, 0, 0, 0);
  }
  if (tid == 0) {
    mbar_init(kvbar, 1);
    for (int w = 0; w < TL::WARPS; ++w) mbar_init(bars + 8 + w * 8, 1);
    fence_mbar_init();
  }
  fence_proxy_async();
  __syncthreads();

  const long long total = p.total;
  const long long cta_begin = total * blockIdx.x / gridDim.x;
  const long long cta_end = total * (blockIdx.x + 1) / gridDim.x;
  T* __restrict__ out = reinterpret_cast<T*>(p.out);
  const uint64_t pol_stream = policy_evict_first();  // Q and W are dead after this pass

  // beta = sigma * std(a), folded with log2(e): softmax runs in the exp2 domain.  Read lazily: with programmatic
  // dependent launch this kernel starts (barrier init, K/V and first Q/W copies) while pass 1 is still finishing.
  float beta_l2 = 0.f;
  bool have_beta = false;
  const float scale_l2 = p.scale * kLog2e;
  const int w_rep = p.B / p.Bw;
  const bool w_tmap = (p.flags & 1u) != 0;          // W rows 16-byte aligned: slices arrive as one 80 x 16 box
  const int wp = w_tmap ? TL::KV_ROWS : p.w_pitch;  // floats between W rows in shared memory

  uint32_t it = 0, kvphase = 0;

  for (long long idx = cta_begin; idx < cta_end;) {
    const Seg sg = seg_of<D>(idx, cta_end, p.n_sl, p.n_hg, p.H);
    const uint32_t row_bytes = sg.nheads * D * 2;
    const float* Wb = p.W + static_cast<long long>(sg.b / w_rep) * p.L * p.w_pitch;
    __syncthreads();  // every warp is done with the previous K/V
    if (tid == 0) {
      mbar_arrive_expect_tx(kvbar, 2 * TL::KV_BYTES);  // two 80-row boxes; rows >= S arrive as zeros
      const uint64_t pol_kv = policy_evict_last();
      tma_box_load(sK, &tm_k, sg.hg * (TL::GW / 2), 0, sg.b, kvbar, pol_kv);
      tma_box_load(sV, &tm_v, sg.hg * (TL::GW / 2), 0, sg.b, kvbar, pol_kv);
    }

    auto issue = [&](long long sl_idx) {
      const int l0 = static_cast<int>(sl_idx % p.n_sl) * TL::ROWS;
      const int rows = min(TL::ROWS, p.L - l0);
      const float* wsrc = Wb + static_cast<long long>(l0) * p.w_pitch;
      const uint32_t wbytes = rows * p.w_pitch * 4;
      const bool w_bulk = ((reinterpret_cast<uintptr_t>(wsrc) | wbytes) & 15) == 0;
      if (lane == 0) {
        mbar_arrive_expect_tx(qbar, TL::QS_BYTES + (w_tmap ? TL::WS_BYTES : w_bulk ? wbytes : 0));
        tma_box_load(sQ, &tm_q, sg.hg * (TL::GW / 2), l0, sg.b, qbar, pol_stream);
        if (w_tmap) tma_box_load(sW, &tm_w, p.w_col0, l0, sg.b / w_rep, qbar, pol_stream);  // columns >= pitch: zeros
      }
      if (w_tmap) {
      } else if (w_bulk) {
        if (lane == 16) bulk_g2s_hint(sW, wsrc, wbytes, qbar, pol_stream);
      } else {  // odd tail / unaligned W: plain loads (visible to this warp after the __syncwarp below)
        float* wdst = const_cast<float*>(wsm);
        for (int i = lane; i < rows * p.w_pitch; i += 32) wdst[i] = __ldg(wsrc + i);
      }
      __syncwarp();
    };

    const long long first = idx + warp;
    if (first < sg.end) issue(first);
    bool kv_ready = false;
    for (long long sl = first; sl < sg.end; sl += TL::WARPS) {
      mbar_wait(qbar, it & 1);
      ++it;
      if (!kv_ready) {
        mbar_wait(kvbar, kvphase);
        kv_ready = true;
      }
      const int l0 = static_cast<int>(sl % p.n_sl) * TL::ROWS;
      const int rows = min(TL::ROWS, p.L - l0);

      if (!have_beta) {
        asm volatile("griddepcontrol.wait;" ::: "memory");  // pass 1 has published the std
        const float sigma = p.sigma_dev ? __ldcg(p.sigma_dev) : p.sigma_host;
        beta_l2 = sigma * __ldcg(&p.ws->std_unbiased) * kLog2e;
        have_beta = true;
      }
      const float* m0 = nullptr;
      if constexpr (MASKED)
        m0 = p.mask + sg.b * p.m_sb + static_cast<long long>(sg.hg * TL::G) * p.m_sh + static_cast<long long>(l0) * p.m_sl + p.m_col0;
      softmax_pv_slice<T, D, MASKED>(sg.nheads, p.S, beta_l2, scale_l2, sQ, sK, sV, wsm, wp, qsm,
                                     p.lse ? p.lse + (static_cast<long long>(sg.b) * p.H + sg.hg * TL::G) * p.L + l0 : nullptr, p.L,
                                     rows, lane, m0, p.m_sh, p.m_sl);

      // O slice: shared -> global through the TMA, one row per lane
      fence_proxy_async();
      __syncwarp();
      if (lane < rows) {
        bulk_s2g(out + sg.b * p.o_sb + static_cast<long long>(l0 + lane) * p.o_sl + sg.hg * TL::GW, sQ + lane * PITCH,
                 row_bytes);
        bulk_commit();
        bulk_wait_read0();  // the slice buffer may be refilled once the TMA has read it
      }
      __syncwarp();
      if (sl + TL::WARPS < sg.end) issue(sl + TL::WARPS);
    }
    kvphase ^= 1;
    idx = sg.end;
  }
  bulk_wait0();  // global writes of this thread's bulk stores are complete before the CTA retires
}

// =============================================================================================
// both passes in ONE launch for problems that fit on chip (small layers, small batches)
//
// When every CTA can keep its share of Q resident in shared memory -- at most one 16-row slice per warp next to the K and
// V of its (batch, head group): the footprint of the pass-2 kernel -- the two passes need not be two launches with Q
// read twice: each warp loads its slice once, forms its scores for the statistics, the CTAs meet at a grid barrier
// (cooperative launch; the barrier IS the deterministic fold of the per-CTA partials: the last arriver publishes the
// std and bumps an epoch word), and the same warp then redoes Q K^T from shared memory for the softmax and P V.
// One launch floor instead of two and Q read once.  CTAs per (batch, head-group) segment: p.fused_cps.
template <typename T, int D>
__global__ void __launch_bounds__(256, 1)
xattn_fused_kernel(const XattnParams p, const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                   const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_w) {
  using TL = Tile<D>;
  constexpr int PITCH = TL::PITCH;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ unsigned int s_epoch0;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sK = smem_u32(smem);
  const uint32_t sV = sK + TL::KV_BYTES;
  const uint32_t sQ = sV + TL::KV_BYTES + warp * (TL::QS_BYTES + TL::WS_BYTES);
  const uint32_t sW = sQ + TL::QS_BYTES;
  const float* wsm = reinterpret_cast<const float*>(smem + 2 * TL::KV_BYTES + warp * (TL::QS_BYTES + TL::WS_BYTES) +
                                                     TL::QS_BYTES);
  unsigned char* qsm = smem + 2 * TL::KV_BYTES + warp * (TL::QS_BYTES + TL::WS_BYTES);
  const uint32_t bars = sK + 2 * TL::KV_BYTES + TL::WARPS * (TL::QS_BYTES + TL::WS_BYTES);
  const uint32_t kvbar = bars;
  const uint32_t qbar = bars + 8 + warp * 8;
  volatile unsigned int* epoch = reinterpret_cast<volatile unsigned int*>(reinterpret_cast<unsigned char*>(p.ws) + kEpochOffset);
  if (tid == 0) {
    s_epoch0 = *epoch;  // read before this CTA arrives: the bump cannot have happened yet
    mbar_init(kvbar, 1);
    for (int w = 0; w < TL::WARPS; ++w) mbar_init(bars + 8 + w * 8, 1);
    fence_mbar_init();
  }
  fence_proxy_async();
  __syncthreads();

  // this CTA: segment (batch, head group) and a run of <= 8 slices, one per warp
  const int seg = blockIdx.x / p.fused_cps, part = blockIdx.x % p.fused_cps;
  const int b = seg / p.n_hg, hg = seg % p.n_hg;
  const int nheads = min(TL::G, p.H - hg * TL::G);
  const int spc = (p.n_sl + p.fused_cps - 1) / p.fused_cps;
  const int sl = part * spc + warp;
  const bool have_slice = warp < spc && sl < p.n_sl;
  const int l0 = sl * TL::ROWS;
  const int rows = have_slice ? min(TL::ROWS, p.L - l0) : 0;
  const int w_rep = p.B / p.Bw;
  const bool w_tmap = (p.flags & 1u) != 0;
  const int wp = w_tmap ? TL::KV_ROWS : p.w_pitch;

  if (tid == 0) {
    mbar_arrive_expect_tx(kvbar, 2 * TL::KV_BYTES);  // two 80-row boxes; keys >= S arrive as zeros
    const uint64_t pol_kv = policy_evict_last();
    tma_box_load(sK, &tm_k, hg * (TL::GW / 2), 0, b, kvbar, pol_kv);
    tma_box_load(sV, &tm_v, hg * (TL::GW / 2), 0, b, kvbar, pol_kv);
  }
  if (have_slice) {
    const uint64_t pol_stream = policy_evict_first();
    const float* wsrc = p.W + (static_cast<long long>(b / w_rep) * p.L + l0) * p.w_pitch;
    const uint32_t wbytes = rows * p.w_pitch * 4;
    const bool w_bulk = ((reinterpret_cast<uintptr_t>(wsrc) | wbytes) & 15) == 0;
    if (lane == 0) {
      mbar_arrive_expect_tx(qbar, TL::QS_BYTES + (w_tmap ? TL::WS_BYTES : w_bulk ? wbytes : 0));
      tma_box_load(sQ, &tm_q, hg * (TL::GW / 2), l0, b, qbar, pol_stream);  // rows >= L arrive as zeros
      if (w_tmap) tma_box_load(sW, &tm_w, 0, l0, b / w_rep, qbar, pol_stream);
    }
    if (!w_tmap) {
      if (w_bulk) {
        if (lane == 16) bulk_g2s_hint(sW, wsrc, wbytes, qbar, pol_stream);
      } else {
        float* wdst = const_cast<float*>(wsm);
        for (int i = lane; i < rows * p.w_pitch; i += 32) wdst[i] = __ldg(wsrc + i);
      }
    }
    __syncwarp();
  }

  // ---- pass 1 on the resident slice
  double dsum = 0.0, dsq = 0.0;
  if (have_slice) {
    mbar_wait(qbar, 0);
    mbar_wait(kvbar, 0);
    for (int h = 0; h < nheads; ++h) {
      float acc[10][4];
      qk_tile<T, D>(acc, sQ + h * D * 2, sK + h * D * 2, lane);
      float fs = 0.f, fq = 0.f;  // rows >= L and keys >= S were zero-filled by the TMA: they add exact zeros
#pragma unroll
      for (int j = 0; j < 10; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          fs += acc[j][i];
          fq = fmaf(acc[j][i], acc[j][i], fq);
        }
      dsum += static_cast<double>(fs);
      dsq += static_cast<double>(fq);
    }
  }
  publish_stats<TL::WARPS>(p, dsum, dsq, true);

  // ---- grid barrier: wait until the last CTA has published the std
  if (tid == 0) {
    unsigned int spins = 0;
    while (*epoch == s_epoch0 && ++spins < (1u << 24)) __nanosleep(40);  // bounded: a launch that is not co-resident must not hang
    s_epoch0 = (*epoch == s_epoch0) ? 1u : 0u;  // reused as "barrier timed out" flag
    __threadfence();
  }
  __syncthreads();
  const bool timed_out = s_epoch0 != 0u;  // cannot happen under a cooperative launch; if it ever does, fail loudly (NaN output)

  // ---- pass 2 from shared memory
  if (have_slice) {
    const float sigma = p.sigma_dev ? __ldcg(p.sigma_dev) : p.sigma_host;
    const float beta_l2 = timed_out ? __int_as_float(0x7fc00000) : sigma * __ldcg(&p.ws->std_unbiased) * kLog2e;
    softmax_pv_slice<T, D>(nheads, p.S, beta_l2, p.scale * kLog2e, sQ, sK, sV, wsm, wp, qsm, nullptr, p.L, rows, lane);
    fence_proxy_async();
    __syncwarp();
    if (lane < rows) {
      T* out = reinterpret_cast<T*>(p.out);
      bulk_s2g(out + b * p.o_sb + static_cast<long long>(l0 + lane) * p.o_sl + hg * TL::GW, sQ + lane * PITCH, nheads * D * 2);
      bulk_commit();
      bulk_wait0();
    }
  }
}

// =============================================================================================
// host-side launchers
// =============================================================================================
static int grid_for(long long total) {
  const int sms = sm_count_cached();
  long long want = (total + 3) / 4;  // at least ~4 slices per CTA before spreading further
  if (want < 1) want = 1;
  return static_cast<int>(want < sms ? want : sms);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}
// [B, rows, cols] 16-bit tensor (element strides sb, sr, 1) seen as 32-bit words; box = (box_words, box_rows, 1)
static bool make_map32(CUtensorMap* m, const void* base, int cols, int rows, int B, long long sr, long long sb,
                       int box_words, int box_rows) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  cuuint64_t gdim[3] = {static_cast<cuuint64_t>(cols / 2), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(B)};
  cuuint64_t gstr[2] = {static_cast<cuuint64_t>(sr) * 2, static_cast<cuuint64_t>(B > 1 ? sb : sr * rows) * 2};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(box_words), static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<void*>(base), gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// fp32 [Bw, L, pitch] region map -> boxes of 80 columns x 16 rows, when the rows are 16-byte aligned (sets *flags |= 1)
static bool make_map_w16(CUtensorMap* m, const XattnParams& p, int box_cols, int box_rows, unsigned* flags) {
  if (p.w_pitch % 4 != 0 || (reinterpret_cast<uintptr_t>(p.W) & 15) != 0) return true;  // not eligible: plain copies
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  cuuint64_t gdim[3] = {static_cast<cuuint64_t>(p.w_pitch), static_cast<cuuint64_t>(p.L), static_cast<cuuint64_t>(p.Bw)};
  cuuint64_t gstr[2] = {static_cast<cuuint64_t>(p.w_pitch) * 4, static_cast<cuuint64_t>(p.L) * p.w_pitch * 4};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  if (enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(p.W), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  *flags |= 1u;
  return true;
}

template <typename T, int D, bool MASKED>
static cudaError_t launch_stats_m(const XattnParams& p_in, cudaStream_t st) {
  using TL = Tile<D>;
  XattnParams p = p_in;
  if (MASKED) {  // the kernel sums a = qk_scale * Q K^T + M itself: the fold must not scale again
    p.qk_scale = p_in.scale;
    p.scale = 1.f;
  }
  static thread_local int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(xattn_stats_kernel<T, D, MASKED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         TL::STATS_SMEM);
    if (e != cudaSuccess) return e;
    configured_dev = dev;
  }
  CUtensorMap tm_q, tm_k;
  if (!make_map32(&tm_q, p.q, p.H * D, p.L, p.B, p.q_sl, p.q_sb, TL::PITCH / 4, TL::ROWS) ||
      !make_map32(&tm_k, p.k, p.H * D, p.S, p.B, p.k_ss, p.k_sb, TL::PITCH / 4, TL::KV_ROWS))
    return cudaErrorInvalidValue;
  xattn_stats_kernel<T, D, MASKED><<<grid_for(p.total), 256, TL::STATS_SMEM, st>>>(p, tm_q, tm_k);
  return cudaGetLastError();
}
template <typename T, int D>
static cudaError_t launch_stats(const XattnParams& p, cudaStream_t st) {
  return p.mask ? launch_stats_m<T, D, true>(p, st) : launch_stats_m<T, D, false>(p, st);
}

template <typename T, int D>
static cudaError_t launch_gram_stats(const XattnParams& p_in, cudaStream_t st) {
  using TL = Tile<D>;
  using GT = GramTile<D>;
  static thread_local int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(xattn_gram_stats_kernel<T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, GT::SMEM);
    if (e != cudaSuccess) return e;
    configured_dev = dev;
  }
  XattnParams p = p_in;
  p.n_sl = (p.L + GT::ROWS - 1) / GT::ROWS;  // 64-row tiles here
  p.total = static_cast<long long>(p.B) * p.n_hg * p.n_sl;
  CUtensorMap tm_q, tm_k;
  if (!make_map32(&tm_q, p.q, p.H * D, p.L, p.B, p.q_sl, p.q_sb, TL::PITCH / 4, GT::ROWS) ||
      !make_map32(&tm_k, p.k, p.H * D, p.S, p.B, p.k_ss, p.k_sb, TL::PITCH / 4, TL::KV_ROWS))
    return cudaErrorInvalidValue;
  const int sms = sm_count_cached();
  const int grid = static_cast<int>(p.total < sms ? p.total : sms);
  xattn_gram_stats_kernel<T, D><<<grid, GT::THREADS, GT::SMEM, st>>>(p, tm_q, tm_k);
  return cudaGetLastError();
}

// slices per CTA <= 8 (one per warp) with whole CTAs per (batch, head-group) segment, all CTAs co-resident
bool fused_plan(int B, int H, int L, int D, int S, int* cps_out) {
  if (S > 80 || heads_per_group(D) == 0) return false;
  const int G = heads_per_group(D);
  const long long n_seg = static_cast<long long>(B) * ((H + G - 1) / G);
  const int sms = sm_count_cached();
  if (n_seg > sms) return false;
  const int n_sl = (L + 15) / 16;
  int cps = static_cast<int>(sms / n_seg);
  if (cps > n_sl) cps = n_sl;
  const int spc = (n_sl + cps - 1) / cps;  // slices per CTA with every SM in use
  if (spc > 8) return false;
  cps = (n_sl + spc - 1) / spc;            // fewest CTAs with that many slices each: a shorter barrier
  if (cps_out) *cps_out = cps;
  return true;
}

template <typename T, int D>
static cudaError_t launch_fused(const XattnParams& p_in, cudaStream_t st) {
  using TL = Tile<D>;
  XattnParams p = p_in;
  static thread_local int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(xattn_fused_kernel<T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, TL::FWD_SMEM);
    if (e != cudaSuccess) return e;
    configured_dev = dev;
  }
  int cps = 0;
  if (!fused_plan(p.B, p.H, p.L, D, p.S, &cps)) return cudaErrorInvalidValue;
  p.fused_cps = cps;
  CUtensorMap tm_q, tm_k, tm_v;
  if (!make_map32(&tm_q, p.q, p.H * D, p.L, p.B, p.q_sl, p.q_sb, TL::PITCH / 4, TL::ROWS) ||
      !make_map32(&tm_k, p.k, p.H * D, p.S, p.B, p.k_ss, p.k_sb, TL::PITCH / 4, TL::KV_ROWS) ||
      !make_map32(&tm_v, p.v, p.H * D, p.S, p.B, p.v_ss, p.v_sb, TL::PITCH / 4, TL::KV_ROWS))
    return cudaErrorInvalidValue;
  CUtensorMap tm_w = tm_q;
  p.flags = 0;
  if (!make_map_w16(&tm_w, p, TL::KV_ROWS, TL::ROWS, &p.flags)) return cudaErrorInvalidValue;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(p.B * p.n_hg * cps);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = TL::FWD_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;  // the grid barrier needs every CTA resident
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, xattn_fused_kernel<T, D>, p, tm_q, tm_k, tm_v, tm_w);
}

template <typename T, int D, bool MASKED>
static cudaError_t launch_forward_m(const XattnParams& p_in, cudaStream_t st) {
  XattnParams p = p_in;
  using TL = Tile<D>;
  static thread_local int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(xattn_forward_kernel<T, D, MASKED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         TL::FWD_SMEM);
    if (e != cudaSuccess) return e;
    configured_dev = dev;
  }
  CUtensorMap tm_q, tm_k, tm_v;
  if (!make_map32(&tm_q, p.q, p.H * D, p.L, p.B, p.q_sl, p.q_sb, TL::PITCH / 4, TL::ROWS) ||
      !make_map32(&tm_k, p.k, p.H * D, p.S, p.B, p.k_ss, p.k_sb, TL::PITCH / 4, TL::KV_ROWS) ||
      !make_map32(&tm_v, p.v, p.H * D, p.S, p.B, p.v_ss, p.v_sb, TL::PITCH / 4, TL::KV_ROWS))
    return cudaErrorInvalidValue;
  CUtensorMap tm_w = tm_q;
  p.flags = 0;
  if (!make_map_w16(&tm_w, p, TL::KV_ROWS, TL::ROWS, &p.flags)) return cudaErrorInvalidValue;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid_for(p.total));
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = TL::FWD_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = config().no_pdl ? 0 : 1;  // may overlap the tail of pass 1 (griddepcontrol.wait before beta)
  return cudaLaunchKernelEx(&cfg, xattn_forward_kernel<T, D, MASKED>, p, tm_q, tm_k, tm_v, tm_w);
}
template <typename T, int D>
static cudaError_t launch_forward(const XattnParams& p, cudaStream_t st) {
  return p.mask ? launch_forward_m<T, D, true>(p, st) : launch_forward_m<T, D, false>(p, st);
}

int heads_per_group(int D) {
  switch (D) {
    case 40: return Tile<40>::G;
    case 64: return Tile<64>::G;
    case 80: return Tile<80>::G;
    case 128: return Tile<128>::G;
    case 160: return Tile<160>::G;
    default: return 0;
  }
}

int stats_grid(long long total) { return grid_for(total); }

#define DSC_DISPATCH_D(FN, T)                         \
  switch (D) {                                        \
    case 40: return FN<T, 40>(p, st);                 \
    case 64: return FN<T, 64>(p, st);                 \
    case 80: return FN<T, 80>(p, st);                 \
    case 128: return FN<T, 128>(p, st);               \
    case 160: return FN<T, 160>(p, st);               \
    default: return cudaErrorInvalidValue;            \
  }

cudaError_t run_stats(const XattnParams& p, int D, int dtype, cudaStream_t st) {
  if (dtype == 0) {
    DSC_DISPATCH_D(launch_stats, __half)
  } else {
    DSC_DISPATCH_D(launch_stats, __nv_bfloat16)
  }
}

bool gram_supports(int D, int S) { return D == 40 && S <= Tile<40>::KV_ROWS; }

cudaError_t run_stats_gram(const XattnParams& p, int D, int dtype, cudaStream_t st) {
  if (!gram_supports(D, p.S)) return cudaErrorInvalidValue;
  return dtype == 0 ? launch_gram_stats<__half, 40>(p, st) : launch_gram_stats<__nv_bfloat16, 40>(p, st);
}

cudaError_t run_fused(const XattnParams& p, int D, int dtype, cudaStream_t st) {
  if (dtype == 0) {
    DSC_DISPATCH_D(launch_fused, __half)
  } else {
    DSC_DISPATCH_D(launch_fused, __nv_bfloat16)
  }
}

cudaError_t run_forward(const XattnParams& p, int D, int dtype, cudaStream_t st) {
  if (dtype == 0) {
    DSC_DISPATCH_D(launch_forward, __half)
  } else {
    DSC_DISPATCH_D(launch_forward, __nv_bfloat16)
  }
}

// =============================================================================================
// long prompts: merge of the per-chunk outputs (each normalised over its own <= 80 keys) with their log2-sum-exp
// =============================================================================================
template <typename T>
__global__ void __launch_bounds__(256) merge_chunks_kernel(const T* __restrict__ chunk_out, const float* __restrict__ lse,
                                                           int n_chunks, T* __restrict__ out, long long o_sb, long long o_sl,
                                                           int B, int H, int L, int D) {
  // one thread per (b, l, h, 8-column group); chunk_out: [C][B][L][H*D] dense, lse: [C][B][H][L]
  const int vec_per_head = D / 8;
  const long long n = static_cast<long long>(B) * L * H * vec_per_head;
  const long long chunk_elems = static_cast<long long>(B) * L * H * D;
  const long long lse_elems = static_cast<long long>(B) * H * L;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int vq = static_cast<int>(i % vec_per_head);
    const int h = static_cast<int>((i / vec_per_head) % H);
    const long long bl = i / (static_cast<long long>(vec_per_head) * H);
    const int l = static_cast<int>(bl % L);
    const int b = static_cast<int>(bl / L);
    const long long li = (static_cast<long long>(b) * H + h) * L + l;
    float m = -INFINITY;
    for (int c = 0; c < n_chunks; ++c) m = fmaxf(m, __ldg(lse + c * lse_elems + li));
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float wsum = 0.f;
    const long long src = (bl * H + h) * D + vq * 8;
    for (int c = 0; c < n_chunks; ++c) {
      const float w = exp2f(__ldg(lse + c * lse_elems + li) - m);
      wsum += w;
      const uint4 raw = __ldg(reinterpret_cast<const uint4*>(chunk_out + c * chunk_elems + src));
      const T* v = reinterpret_cast<const T*>(&raw);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(w, static_cast<float>(v[j]), acc[j]);
    }
    const float inv = 1.f / wsum;
    uint4 o;
    o.x = Mma<T>::pack(acc[0] * inv, acc[1] * inv);
    o.y = Mma<T>::pack(acc[2] * inv, acc[3] * inv);
    o.z = Mma<T>::pack(acc[4] * inv, acc[5] * inv);
    o.w = Mma<T>::pack(acc[6] * inv, acc[7] * inv);
    *reinterpret_cast<uint4*>(out + b * o_sb + l * o_sl + h * D + vq * 8) = o;
  }
}

cudaError_t run_merge_chunks(const void* chunk_out, const float* lse, int n_chunks, void* out, long long o_sb, long long o_sl,
                             int B, int H, int L, int D, int dtype, cudaStream_t st) {
  const long long n = static_cast<long long>(B) * L * H * (D / 8);
  const int grid = static_cast<int>(std::min<long long>((n + 255) / 256, 148 * 16));
  if (dtype == 0)
    merge_chunks_kernel<__half><<<grid, 256, 0, st>>>(static_cast<const __half*>(chunk_out), lse, n_chunks,
                                                     static_cast<__half*>(out), o_sb, o_sl, B, H, L, D);
  else
    merge_chunks_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(chunk_out), lse, n_chunks,
                                                            static_cast<__nv_bfloat16*>(out), o_sb, o_sl, B, H, L, D);
  return cudaGetLastError();
}

}  // namespace dsc

#ifdef DSC_TRACE
extern "C" int dsc_debug_gram_times(unsigned long long* out /*HOST 160*8*/) {
  cudaDeviceSynchronize();
  return static_cast<int>(cudaMemcpyFromSymbol(out, dsc::g_gram_times, sizeof(unsigned long long) * 160 * 8));
}
#endif
