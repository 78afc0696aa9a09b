// Microbenchmark: throughput of 1-D TMA bulk copies (cp.async.bulk global->shared) as a function of the
// copy size, and of 2-D tensor-map box loads, one producer warp per CTA, 148 CTAs, 3 stages in flight.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_rate tma_rate.cu   (no -lcuda: driver entry point)
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c)); }
__device__ __forceinline__ void expect(uint32_t b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n) : "memory"); }
__device__ __forceinline__ void wait(uint32_t b, uint32_t ph) {
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(b), "r"(ph) : "memory");
}
__device__ __forceinline__ void g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma2d(uint32_t dst, const CUtensorMap* m, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

constexpr int NST = 3;
// mode 0: each of the 32 lanes issues copies; mode 1: lane 0 issues all
__global__ void k_bulk(const char* src, size_t row_stride, int rows_per_tile, int copy_bytes, int tiles, int mode, unsigned* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bars[NST];
  const int lane = threadIdx.x;
  const uint32_t stage_bytes = rows_per_tile * copy_bytes;
  if (lane == 0) { for (int s = 0; s < NST; ++s) mbar_init(s32(&bars[s]), 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncwarp();
  const char* base = src + (size_t)blockIdx.x * tiles * rows_per_tile * row_stride;
  auto issue = [&](int t) {
    const int s = t % NST;
    if (lane == 0) expect(s32(&bars[s]), stage_bytes);
    __syncwarp();
    if (mode == 0) { for (int r = lane; r < rows_per_tile; r += 32) g2s(s32(smem) + s * stage_bytes + r * copy_bytes, base + ((size_t)t * rows_per_tile + r) * row_stride, copy_bytes, s32(&bars[s])); }
    else if (lane == 0) { for (int r = 0; r < rows_per_tile; ++r) g2s(s32(smem) + s * stage_bytes + r * copy_bytes, base + ((size_t)t * rows_per_tile + r) * row_stride, copy_bytes, s32(&bars[s])); }
  };
  for (int t = 0; t < NST && t < tiles; ++t) issue(t);
  unsigned acc = 0;
  for (int t = 0; t < tiles; ++t) {
    wait(s32(&bars[t % NST]), (t / NST) & 1);
    acc += smem[(t % NST) * stage_bytes + lane];
    __syncwarp();
    if (t + NST < tiles) issue(t + NST);
  }
  if (acc == 0xdeadbeef) sink[0] = acc;
}

__global__ void k_tensor(const __grid_constant__ CUtensorMap map, int box_rows, int box_bytes_row, int tiles, unsigned* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bars[NST];
  const int lane = threadIdx.x;
  const uint32_t stage_bytes = box_rows * box_bytes_row;
  if (lane == 0) { for (int s = 0; s < NST; ++s) mbar_init(s32(&bars[s]), 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncwarp();
  auto issue = [&](int t) {
    const int s = t % NST;
    if (lane == 0) { expect(s32(&bars[s]), stage_bytes); tma2d(s32(smem) + s * stage_bytes, &map, 0, (blockIdx.x * tiles + t) * box_rows, s32(&bars[s])); }
  };
  for (int t = 0; t < NST && t < tiles; ++t) issue(t);
  unsigned acc = 0;
  for (int t = 0; t < tiles; ++t) {
    wait(s32(&bars[t % NST]), (t / NST) & 1);
    acc += smem[(t % NST) * stage_bytes + lane];
    __syncwarp();
    if (t + NST < tiles) issue(t + NST);
  }
  if (acc == 0xdeadbeef) sink[0] = acc;
}

// NB boxes of box_cols x box_rows (adjacent column ranges of the same rows) per tile, all on one barrier, NST tiles in
// flight: what staging K / V into 16-byte-row UMMA layouts by TMA would look like (NB = 20, box_cols = 8, box_rows = 80)
__global__ void k_tensor_multi(const __grid_constant__ CUtensorMap map, int box_rows, int box_bytes_row, int nb, int tiles, unsigned* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bars[NST];
  const int lane = threadIdx.x;
  const uint32_t box_bytes = box_rows * box_bytes_row, stage_bytes = nb * box_bytes;
  if (lane == 0) { for (int s = 0; s < NST; ++s) mbar_init(s32(&bars[s]), 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncwarp();
  auto issue = [&](int t) {
    const int s = t % NST;
    if (lane == 0) {
      expect(s32(&bars[s]), stage_bytes);
      for (int j = 0; j < nb; ++j)
        tma2d(s32(smem) + s * stage_bytes + j * box_bytes, &map, j * (box_bytes_row / 2), (blockIdx.x * tiles + t) * box_rows, s32(&bars[s]));
    }
  };
  for (int t = 0; t < NST && t < tiles; ++t) issue(t);
  unsigned acc = 0;
  for (int t = 0; t < tiles; ++t) {
    wait(s32(&bars[t % NST]), (t / NST) & 1);
    acc += smem[(t % NST) * stage_bytes + lane];
    __syncwarp();
    if (t + NST < tiles) issue(t + NST);
  }
  if (acc == 0xdeadbeef) sink[0] = acc;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const size_t bufsz = (size_t)2 << 30;
  char* buf; unsigned* sink;
  cudaMalloc(&buf, bufsz); cudaMemset(buf, 1, bufsz); cudaMalloc(&sink, 4);
  cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_tensor, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int ctas = 148;
  struct Cfg { int copy_bytes, rows; size_t stride; int mode; };
  Cfg cfgs[] = {{320, 128, 640, 0}, {320, 128, 640, 1}, {640, 64, 640, 0}, {640, 64, 1280, 0}, {1280, 32, 1280, 0}, {2560, 16, 2560, 0},
                {5120, 8, 5120, 0}, {10240, 4, 10240, 0}, {40960, 1, 40960, 0}, {320, 128, 320, 0}, {160, 128, 640, 0}, {4928, 8, 4928, 0}};
  for (auto c : cfgs) {
    size_t tile_span = (size_t)c.rows * c.stride;
    int tiles = (int)(bufsz / ctas / tile_span); if (tiles > 64) tiles = 64;
    int smem = NST * c.rows * c.copy_bytes;
    k_bulk<<<ctas, 32, smem>>>(buf, c.stride, c.rows, c.copy_bytes, tiles, c.mode, sink);
    cudaEventRecord(e0);
    k_bulk<<<ctas, 32, smem>>>(buf, c.stride, c.rows, c.copy_bytes, tiles, c.mode, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double bytes = (double)ctas * tiles * c.rows * c.copy_bytes;
    printf("{\"bench\":\"bulk_1d\",\"copy_bytes\":%d,\"copies_per_tile\":%d,\"row_stride\":%zu,\"issuer\":\"%s\",\"tiles\":%d,\"ms\":%.4f,\"gbs\":%.1f,\"ns_per_copy_per_sm\":%.1f,\"err\":\"%s\"}\n",
           c.copy_bytes, c.rows, c.stride, c.mode ? "lane0" : "32lanes", tiles, ms, bytes / ms / 1e6, ms * 1e6 / (tiles * c.rows), cudaGetErrorString(cudaGetLastError()));
  }
  // 2-D tensor map: matrix [rows][640 B] of u16 (320 cols), box = box_cols x 128 rows
  EncodeFn encode = nullptr; cudaDriverEntryPointQueryResult qres;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres);
  if (!encode) { printf("{\"bench\":\"tensor_2d\",\"err\":\"no cuTensorMapEncodeTiled\"}\n"); return 0; }
  int boxcols[] = {160, 168, 320, 64, 32, 16, 8};
  for (int bc : boxcols) {
    cuuint64_t rows_total = bufsz / 640;
    cuuint64_t gdim[2] = {320, rows_total}; cuuint64_t gstr[1] = {640};
    cuuint32_t box[2] = {(cuuint32_t)bc, 128}; cuuint32_t estr[2] = {1, 1};
    CUtensorMap map;
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, buf, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("{\"bench\":\"tensor_2d\",\"box_cols\":%d,\"err\":\"encode %d\"}\n", bc, (int)r); continue; }
    int tiles = 64; int smem = NST * 128 * bc * 2;
    k_tensor<<<ctas, 32, smem>>>(map, 128, bc * 2, tiles, sink);
    cudaEventRecord(e0);
    k_tensor<<<ctas, 32, smem>>>(map, 128, bc * 2, tiles, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double bytes = (double)ctas * tiles * 128 * bc * 2;
    printf("{\"bench\":\"tensor_2d\",\"box_cols\":%d,\"box_rows\":128,\"tiles\":%d,\"ms\":%.4f,\"gbs\":%.1f,\"err\":\"%s\"}\n", bc, tiles, ms, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  }
  // many narrow boxes per tile
  struct { int bc, rows, nb; } multi[] = {{8, 80, 20}, {8, 80, 40}, {8, 128, 20}, {32, 128, 5}, {16, 80, 10}};
  for (auto m : multi) {
    cuuint64_t rows_total = bufsz / 640;
    cuuint64_t gdim[2] = {320, rows_total}; cuuint64_t gstr[1] = {640};
    cuuint32_t box[2] = {(cuuint32_t)m.bc, (cuuint32_t)m.rows}; cuuint32_t estr[2] = {1, 1};
    CUtensorMap map;
    if (encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, buf, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) continue;
    int tiles = 64; int smem = NST * m.nb * m.rows * m.bc * 2;
    cudaFuncSetAttribute(k_tensor_multi, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k_tensor_multi<<<ctas, 32, smem>>>(map, m.rows, m.bc * 2, m.nb, tiles, sink);
    cudaEventRecord(e0);
    k_tensor_multi<<<ctas, 32, smem>>>(map, m.rows, m.bc * 2, m.nb, tiles, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double bytes = (double)ctas * tiles * m.nb * m.rows * m.bc * 2;
    printf("{\"bench\":\"tensor_2d_multi\",\"box_cols\":%d,\"box_rows\":%d,\"boxes_per_tile\":%d,\"tiles\":%d,\"ms\":%.4f,\"us_per_tile\":%.3f,\"ns_per_box_row\":%.2f,\"gbs\":%.1f,\"err\":\"%s\"}\n",
           m.bc, m.rows, m.nb, tiles, ms, ms * 1e3 / tiles, ms * 1e6 / tiles / (m.nb * m.rows), bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
