// resnet_kernels.cu -- tcgen05 tensor-core path for the dense FC / residual blocks of the cost-to-go ResNet
// (utils/pytorch_models.py:45-86 with eval-mode BatchNorm folded into the Linear layers).
//
// One kernel per layer:   OUT = epilogue( sum_p A_p[M,K] * W_p[N,K]^T )
//   operands   fp16, K-major, staged by TMA (cp.async.bulk.tensor.2d, 128-byte swizzle) into shared memory
//   math       tcgen05.mma.cta_group::1.kind::f16, 128x256x16 per instruction, fp32 accumulators in TMEM
//              (2 x 256 columns, double buffered so the epilogue of tile i overlaps the MMAs of tile i+1)
//   epilogue   tcgen05.ld -> v = acc*scale + bias (+ skip) -> ReLU -> fp16 hi / lo split written for the next
//              layer (and optionally fp32)
// Precision: a layer is a list of up to three operand PAIRS swept over K one after the other into the same
// accumulator: [A_hi*W_hi] (plain fp16 GEMM), or [A_hi*W_lo, A_lo*W_hi, A_hi*W_hi] with x = hi + lo (two fp16 terms
// = 22 significand bits) -- the fp32-parity mode of the search -- or [A_hi*W_lo, A_hi*W_hi] when A is exact in
// fp16 (the one-hot first layer).  The SMALL products are swept FIRST: the tensor core's fp32 accumulation
// truncates (measured: error grows linearly with the number of accumulated MMAs), so adding 2^-11-sized terms into
// an already large accumulator would lose them; summed first they arrive intact, and the main sweep is kept to
// K <= 1024 per accumulator (longer K is split by the caller into launches chained through an fp32 partial sum).
// Measured on the trained cube3 network: max |error| 3.5e-5 vs fp64 (3.4e-4 with a single interleaved sweep).
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue.
// The production kernel is the CTA-PAIR variant (tcgen05 cta_group::2, 256 x 256 tile per cluster of two CTAs).
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdlib.h>
#include "dcb_internal.h"

namespace dcb {
namespace {

constexpr int BM = 128, BN = 256, BK = 64;            // CTA tile; BK fp16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int kThreads = 192;
constexpr uint32_t kTmemCols = 512;
constexpr int A_TILE_BYTES = BM * BK * 2;              // 16 KB
constexpr int W_TILE_BYTES = BN * BK * 2;              // 32 KB
constexpr int kMaxPairs = 3;
// Epilogue staging: per epilogue warp two slice buffers of {hi tile, lo tile}, a tile = 32 rows x 64 fp16 = 4 KB in the TMA
// 128-byte-swizzle layout.  The residual slice is TMA-loaded INTO the buffer, consumed, overwritten in place with the output
// slice and TMA-stored from there.
constexpr int EPI_TILE_BYTES = 32 * 128;
constexpr int EPI_BUF_BYTES = 2 * EPI_TILE_BYTES;      // hi + lo
constexpr int EPI_WARP_BYTES = 2 * EPI_BUF_BYTES;      // double buffered
constexpr int EPI_BYTES = 4 * EPI_WARP_BYTES;          // 64 KB per CTA

template <bool PAIR> struct Cfg;
template <> struct Cfg<false> {                        // one CTA computes a 128 x 256 tile
  static constexpr int STAGE_BYTES = A_TILE_BYTES + W_TILE_BYTES;    // 48 KB
  static constexpr int STAGES = 3;
};
template <> struct Cfg<true> {                         // a CTA pair computes a 256 x 256 tile; each CTA stages its 128 A rows and half of W
  static constexpr int STAGE_BYTES = A_TILE_BYTES + W_TILE_BYTES / 2;   // 32 KB
  static constexpr int STAGES = 5;
};
template <bool PAIR> constexpr int smem_bytes() { return Cfg<PAIR>::STAGES * Cfg<PAIR>::STAGE_BYTES + EPI_BYTES + 1024 /*align slack*/ + 256 /*barriers*/; }

// ---- PTX wrappers ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  const uint32_t addr = smem_addr(bar);
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_load_2d(const CUtensorMap *map, uint64_t *bar, void *dst, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_addr(dst)),
               "l"(map), "r"(smem_addr(bar)), "r"(c0), "r"(c1)
               : "memory");
}
// TMA tile store shared -> global (rows / columns outside the tensor are clipped)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, const void *src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_addr(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
// ---- CTA-pair (cta_group::2) variants -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t map_to_cta(const void *p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion bytes are credited to the mbarrier of the LEADER CTA of the pair
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap *map, uint32_t leader_bar_cluster_addr, void *dst, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_addr(dst)),
      "l"(map), "r"(leader_bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
template <bool PAIR> __device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst, uint32_t cols) {
  if constexpr (PAIR) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(smem_dst)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(smem_dst)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
}
template <bool PAIR> __device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  if constexpr (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
  else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
template <bool PAIR> __device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  if constexpr (PAIR) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// arrive on an mbarrier once all previously issued MMAs of this thread have completed (pair: on the barrier at this offset in
// BOTH CTAs of the pair)
template <bool PAIR> __device__ __forceinline__ void umma_commit(uint64_t *bar) {
  if constexpr (PAIR)
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_addr(bar)),
                 "h"((uint16_t)0x3)
                 : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128-byte-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // start address            bits [0,14)
  d |= (uint64_t)1 << 16;                           // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                 // stride byte offset = 1024 bits [32,46)
  d |= (uint64_t)1 << 46;                           // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                           // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: fp16 x fp16 -> fp32, both operands K-major, N = 256, M = 128 (one CTA) or 256 (CTA pair)
template <bool PAIR> constexpr uint32_t idesc() { return (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(((PAIR ? 2 : 1) * BM) >> 4) << 24); }

struct EpilogueArgs {
  const float *bias;          // [Np]
  float scale;                // acc * scale (undoes the power-of-two weight pre-scaling)
  const __half *skip_hi;      // [M][Np] or nullptr  (read through EpiMaps::skip_hi)
  const __half *skip_lo;      // [M][Np] or nullptr
  int relu;
  __half *out_hi;             // [M][Np] or nullptr  (written through EpiMaps::out_hi)
  __half *out_lo;             // [M][Np] or nullptr
  float *out_f32;             // [M][Np] or nullptr
  const float *partial_in;    // [M][Np] fp32 partial sum of earlier K chunks of ANOTHER launch (unscaled) or nullptr
  float *partial_out;         // if set: write acc (+ partial_in) here and skip the rest of the epilogue
  const float *dot_w;         // [Np] or nullptr: fused fc_out -- dot_partial[row][n_tile] = sum over the tile's columns of out * dot_w
  float *dot_partial;         // [M][Np/256]; the caller adds the n_tile partials in a fixed order (reproducible)
};

struct PairMaps {              // operand pairs, swept in order; the last one is the main product
  CUtensorMap a[kMaxPairs];
  CUtensorMap w[kMaxPairs];
};
struct EpiMaps {               // [M][Np] fp16 matrices as 64-column x 32-row boxes, 128-byte swizzle
  CUtensorMap out_hi, out_lo, skip_hi, skip_lo;
};
struct GemmParams {
  int n_pairs;
  int64_t M;                   // rows covered by the tensor maps (upper bound of the row count)
  const int32_t *m_dev;        // optional DEVICE-side row count: rows = clamp(*m_dev - m_off, 0, M) -- lets a search iteration be
  int32_t m_off;               //   launched without the host knowing how many children survived CLOSED
  int Np, k_blocks;
  int chunk_kb;                // k-blocks accumulated per TMEM accumulator; longer K is folded on chip through `scratch`
  float4 *scratch;             // [gridDim.x][64][128] float4, L2-resident per-CTA partial sums (only when k_blocks > chunk_kb)
};

struct EpiWarp {
  uint8_t *buf;                // this warp's two slice buffers
  uint64_t *bar;               // [2] "residual slice landed"
  uint32_t phase;              // bit b = parity to wait for on bar[b]
};

// Residual slices 0 and 1 of a tile -> this warp's two staging buffers (issued before the accumulator is awaited).
__device__ __forceinline__ void epilogue_prefetch_skip(const EpilogueArgs &ep, const EpiMaps &em, EpiWarp &ew, int32_t row0, int n0, int lane) {
  if (!ep.skip_hi || ep.partial_out) return;
  __syncwarp();                                           // every lane is done with the buffers
  if (lane == 0) {
    bulk_wait_read<0>();                                  // ... and so are the stores issued from them
    const uint32_t bytes = ep.skip_lo ? EPI_BUF_BYTES : EPI_TILE_BYTES;
#pragma unroll
    for (int b = 0; b < 2; b++) {
      mbar_expect_tx(&ew.bar[b], bytes);
      tma_load_2d(&em.skip_hi, &ew.bar[b], ew.buf + b * EPI_BUF_BYTES, n0 + b * 64, row0);
      if (ep.skip_lo) tma_load_2d(&em.skip_lo, &ew.bar[b], ew.buf + b * EPI_BUF_BYTES + EPI_TILE_BYTES, n0 + b * 64, row0);
    }
  }
}

// One accumulator (this warp's 32 TMEM lanes x 256 columns) of work unit (tile, K chunk).  A thread owns one output row.
//   non-final chunk : acc (+ earlier chunks) -> per-CTA fp32 scratch (coalesced float4 columns, stays in L2)
//   final chunk     : v = (acc + scratch) * scale + bias (+ residual) -> ReLU -> fp16 hi / lo, written in place over the residual
//                     slice in the swizzled staging tile and shipped with TMA stores; optional fp32 output / fused fc_out dot.
// The accumulator is handed back to the MMA warp right after the last tcgen05.ld, before the arithmetic of the last slice.
template <bool PAIR>
__device__ __forceinline__ void epilogue_unit(const EpilogueArgs &ep, const EpiMaps &em, uint32_t taddr, int64_t row0, int n0, int64_t M_eff, int Np,
                                              int lane, int row_in_tile, EpiWarp &ew, int chunk, int n_chunks, float4 *scr,
                                              uint64_t *empty_bar, uint32_t empty_bar_leader) {
  const int64_t row = row0 + lane;
  const bool row_ok = row < M_eff;
  const bool final_unit = chunk == n_chunks - 1;
  const bool has_skip = ep.skip_hi != nullptr && !ep.partial_out, has_skip_lo = ep.skip_lo != nullptr, tma_out = ep.out_hi != nullptr;
  float dot = 0.0f;
#pragma unroll 1
  for (int c = 0; c < 4; c++) {
    float v[64];
    {
      uint32_t acc0[32], acc1[32];
      tmem_ld_32x32(taddr + c * 64, acc0);
      tmem_ld_32x32(taddr + c * 64 + 32, acc1);
#pragma unroll
      for (int j = 0; j < 32; j++) { v[j] = __uint_as_float(acc0[j]); v[32 + j] = __uint_as_float(acc1[j]); }
    }
    if (c == 3) {                                  // accumulator drained: the MMA warp may overwrite it
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (PAIR) mbar_arrive_cluster(empty_bar_leader); else mbar_arrive(empty_bar);
      }
    }
    if (n_chunks > 1) {
      float4 *s = scr + (c * 16) * 128 + row_in_tile;
      if (chunk > 0) {
#pragma unroll
        for (int q = 0; q < 16; q++) { const float4 f = s[q * 128]; v[4 * q] += f.x; v[4 * q + 1] += f.y; v[4 * q + 2] += f.z; v[4 * q + 3] += f.w; }
      }
      if (!final_unit) {
#pragma unroll
        for (int q = 0; q < 16; q++) s[q * 128] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        continue;
      }
    }
    const int64_t off = row * Np + n0 + c * 64;
    if (row_ok && ep.partial_in) {
      const float4 *pi = reinterpret_cast<const float4 *>(ep.partial_in + off);
#pragma unroll
      for (int q = 0; q < 16; q++) { const float4 f = pi[q]; v[4 * q] += f.x; v[4 * q + 1] += f.y; v[4 * q + 2] += f.z; v[4 * q + 3] += f.w; }
    }
    if (ep.partial_out) {                          // cross-launch K-chunk chaining: raw fp32 sums only (warp-uniform branch)
      if (row_ok) {
        float4 *po = reinterpret_cast<float4 *>(ep.partial_out + off);
#pragma unroll
        for (int q = 0; q < 16; q++) po[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
      continue;
    }
#pragma unroll
    for (int j = 0; j < 64; j++) v[j] = fmaf(v[j], ep.scale, __ldg(ep.bias + n0 + c * 64 + j));
    const int b = c & 1;
    uint8_t *tile_hi = ew.buf + b * EPI_BUF_BYTES, *tile_lo = tile_hi + EPI_TILE_BYTES;
    if (has_skip) {                                // residual slice: chunk q of row `lane` sits at 16-byte slot q ^ (lane & 7)
      mbar_wait(&ew.bar[b], (ew.phase >> b) & 1u);
      ew.phase ^= 1u << b;
#pragma unroll
      for (int q = 0; q < 8; q++) {
        const uint4 u = *reinterpret_cast<const uint4 *>(tile_hi + lane * 128 + ((q ^ (lane & 7)) << 4));
        const __half2 *h = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
        for (int e = 0; e < 4; e++) { const float2 f = __half22float2(h[e]); v[q * 8 + 2 * e] += f.x; v[q * 8 + 2 * e + 1] += f.y; }
      }
      if (has_skip_lo) {
#pragma unroll
        for (int q = 0; q < 8; q++) {
          const uint4 u = *reinterpret_cast<const uint4 *>(tile_lo + lane * 128 + ((q ^ (lane & 7)) << 4));
          const __half2 *h = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
          for (int e = 0; e < 4; e++) { const float2 f = __half22float2(h[e]); v[q * 8 + 2 * e] += f.x; v[q * 8 + 2 * e + 1] += f.y; }
        }
      }
    }
    if (ep.relu) {
#pragma unroll
      for (int j = 0; j < 64; j++) v[j] = fmaxf(v[j], 0.0f);
    }
    if (row_ok && ep.out_f32) {
      float4 *o = reinterpret_cast<float4 *>(ep.out_f32 + off);
#pragma unroll
      for (int q = 0; q < 16; q++) o[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
    if (ep.dot_w) {                                // fused fc_out (pytorch_models.py:85): this row's share of the dot product
#pragma unroll
      for (int j = 0; j < 64; j++) dot = fmaf(v[j], __ldg(ep.dot_w + n0 + c * 64 + j), dot);
    }
    if (tma_out) {
      if (!has_skip) {                             // the store issued from this buffer two slices ago must have read it
        if (lane == 0) bulk_wait_read<1>();
        __syncwarp();
      }
#pragma unroll
      for (int q = 0; q < 8; q++) {                // fp16 hi / lo split, in place
        uint4 ph, pl;
        __half2 *hh = reinterpret_cast<__half2 *>(&ph), *ll = reinterpret_cast<__half2 *>(&pl);
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const float a = fminf(fmaxf(v[8 * q + 2 * e], -65504.0f), 65504.0f), bb = fminf(fmaxf(v[8 * q + 2 * e + 1], -65504.0f), 65504.0f);
          const __half2 h = __floats2half2_rn(a, bb);
          const float2 hf = __half22float2(h);
          hh[e] = h;
          ll[e] = __floats2half2_rn(a - hf.x, bb - hf.y);
        }
        const int slot = lane * 128 + ((q ^ (lane & 7)) << 4);
        *reinterpret_cast<uint4 *>(tile_hi + slot) = ph;
        *reinterpret_cast<uint4 *>(tile_lo + slot) = pl;
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&em.out_hi, tile_hi, n0 + c * 64, (int32_t)row0);
        if (ep.out_lo) tma_store_2d(&em.out_lo, tile_lo, n0 + c * 64, (int32_t)row0);
        bulk_commit_group();
      }
    }
    if (has_skip && c < 2) {                       // residual slice c + 2 -> the buffer just consumed
      if (!tma_out) __syncwarp();
      if (lane == 0) {
        if (tma_out) bulk_wait_read<0>();
        mbar_expect_tx(&ew.bar[b], has_skip_lo ? EPI_BUF_BYTES : EPI_TILE_BYTES);
        tma_load_2d(&em.skip_hi, &ew.bar[b], tile_hi, n0 + (c + 2) * 64, (int32_t)row0);
        if (has_skip_lo) tma_load_2d(&em.skip_lo, &ew.bar[b], tile_lo, n0 + (c + 2) * 64, (int32_t)row0);
      }
    }
  }
  if (final_unit && ep.dot_w && row_ok) ep.dot_partial[row * (Np / BN) + n0 / BN] = dot;
}

// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue.
// A work unit is (output tile, K chunk); units alternate between the two TMEM accumulators so the epilogue of one overlaps the
// MMAs of the next.  PAIR: two CTAs of a cluster compute one 256 x 256 tile with tcgen05 cta_group::2 -- each CTA stages its own
// 128 A rows and HALF of the W tile, the leader's MMA (M = 256) reads both halves, so a CTA moves 32 KB instead of 48 KB of
// shared memory per four MMAs (the single-CTA MMA is shared-memory-bandwidth bound: ~184 instead of 128 cycles).  Barriers of
// the pair: full[s] lives in the LEADER (its arrive.expect_tx covers both CTAs' bytes), empty[s] and tmem_full[b] are signalled
// in both CTAs by the leader's multicast commits, tmem_empty[b] (leader) collects the 2 x 4 epilogue warps.
template <bool PAIR>
__device__ __forceinline__ void gemm_body(const PairMaps &maps, const EpiMaps &emaps, const EpilogueArgs &ep, const GemmParams &gp) {
  constexpr int STAGES = Cfg<PAIR>::STAGES, STAGE_BYTES = Cfg<PAIR>::STAGE_BYTES;
  extern __shared__ uint8_t smem_dyn[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  uint8_t *ep_stage = smem + STAGES * STAGE_BYTES;
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(ep_stage + EPI_BYTES);
  uint64_t *empty_bar = full_bar + STAGES;
  uint64_t *tmem_full = empty_bar + STAGES;          // [2]
  uint64_t *tmem_empty = tmem_full + 2;              // [2]
  uint64_t *skip_bar = tmem_empty + 2;               // [4 warps][2]
  uint32_t *tmem_base_slot = reinterpret_cast<uint32_t *>(skip_bar + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  int64_t M_eff = gp.M;
  if (gp.m_dev) {
    const int64_t d = (int64_t)(*gp.m_dev) - gp.m_off;
    M_eff = d < 0 ? 0 : (d < gp.M ? d : gp.M);
  }
  const int Np = gp.Np, n_tiles = Np / BN;
  const int n_chunks = (gp.k_blocks + gp.chunk_kb - 1) / gp.chunk_kb;
  constexpr int ROWS_PER_TILE = PAIR ? 2 * BM : BM;
  const int64_t total_tiles = ((M_eff + ROWS_PER_TILE - 1) / ROWS_PER_TILE) * n_tiles;
  const int64_t tile_first = PAIR ? (blockIdx.x >> 1) : blockIdx.x, tile_step = PAIR ? (gridDim.x >> 1) : gridDim.x;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; b++) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], PAIR ? 8 : 4); }   // epilogue warps arrive
    for (int b = 0; b < 8; b++) mbar_init(&skip_bar[b], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<PAIR>(tmem_base_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const uint32_t leader_full0 = PAIR ? map_to_cta(&full_bar[0], 0) : 0u;
      for (int64_t t = tile_first; t < total_tiles; t += tile_step) {
        const int32_t m0 = PAIR ? (int32_t)((2 * (t / n_tiles) + rank) * BM) : (int32_t)((t / n_tiles) * BM);
        const int32_t n0 = PAIR ? (int32_t)((t % n_tiles) * BN + rank * (BN / 2)) : (int32_t)((t % n_tiles) * BN);
        for (int ch = 0; ch < n_chunks; ch++) {
          const int kb0 = ch * gp.chunk_kb, kb1 = min(kb0 + gp.chunk_kb, gp.k_blocks);
          for (int p = 0; p < gp.n_pairs; p++) {
            for (int kb = kb0; kb < kb1; kb++) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              uint8_t *st = smem + stage * STAGE_BYTES;
              if constexpr (PAIR) {
                // bytes of both CTAs land on the leader's barrier.  The peer does not arrive on it: a release-arrive across the
                // cluster costs its producer ~0.5 us per stage and the leader's own arrive.expect_tx keeps the phase open.
                if (leader) mbar_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
                tma_load_2d_pair(&maps.a[p], leader_full0 + 8u * (uint32_t)stage, st, kb * BK, m0);
                tma_load_2d_pair(&maps.w[p], leader_full0 + 8u * (uint32_t)stage, st + A_TILE_BYTES, kb * BK, n0);
              } else {
                mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
                tma_load_2d(&maps.a[p], &full_bar[stage], st, kb * BK, m0);
                tma_load_2d(&maps.w[p], &full_bar[stage], st + A_TILE_BYTES, kb * BK, n0);
              }
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: one thread (of the leader CTA) =====================
    if (leader && lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int buf = 0; uint32_t acc_phase = 0;
      for (int64_t t = tile_first; t < total_tiles; t += tile_step) {
        for (int ch = 0; ch < n_chunks; ch++) {
          const int kb0 = ch * gp.chunk_kb, kb1 = min(kb0 + gp.chunk_kb, gp.k_blocks);
          mbar_wait(&tmem_empty[buf], acc_phase ^ 1);            // the epilogue has drained this accumulator
          tc_fence_after();
          const uint32_t tmem_acc = tmem_base + (uint32_t)buf * BN;
          uint32_t first = 1;
          for (int p = 0; p < gp.n_pairs; p++) {
            for (int kb = kb0; kb < kb1; kb++) {
              mbar_wait(&full_bar[stage], phase);
              tc_fence_after();
              const uint32_t st = smem_addr(smem + stage * STAGE_BYTES);
              const uint64_t da = make_smem_desc(st), dw = make_smem_desc(st + A_TILE_BYTES);
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; k++) {
                const uint64_t koff = (uint64_t)((k * UMMA_K * 2) >> 4);      // 32 bytes per K step inside the swizzle row
                umma_f16<PAIR>(tmem_acc, da + koff, dw + koff, idesc<PAIR>(), first ? 0u : 1u);
                first = 0;
              }
              umma_commit<PAIR>(&empty_bar[stage]);               // frees the smem stage when these MMAs retire
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
          }
          umma_commit<PAIR>(&tmem_full[buf]);                     // accumulator complete
          if (++buf == 2) { buf = 0; acc_phase ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue: 4 warps, warp (w % 4) owns TMEM lanes 32*(w%4) .. +31 of its CTA =====================
    const int lane_grp = warp & 3;
    EpiWarp ew{ep_stage + (warp - 2) * EPI_WARP_BYTES, skip_bar + 2 * (warp - 2), 0u};
    float4 *scr = gp.scratch ? gp.scratch + (size_t)blockIdx.x * (64 * 128) : nullptr;
    const uint32_t leader_tmem_empty0 = PAIR ? map_to_cta(&tmem_empty[0], 0) : 0u;
    int buf = 0; uint32_t acc_phase = 0;
    for (int64_t t = tile_first; t < total_tiles; t += tile_step) {
      const int64_t row0 = (PAIR ? (2 * (t / n_tiles) + rank) : (t / n_tiles)) * BM + lane_grp * 32;
      const int n0 = (int)((t % n_tiles) * BN);
      for (int ch = 0; ch < n_chunks; ch++) {
        if (ch == n_chunks - 1) epilogue_prefetch_skip(ep, emaps, ew, (int32_t)row0, n0, lane);
        mbar_wait(&tmem_full[buf], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)buf * BN;
        epilogue_unit<PAIR>(ep, emaps, taddr, row0, n0, M_eff, Np, lane, lane_grp * 32 + lane, ew, ch, n_chunks, scr, &tmem_empty[buf],
                            leader_tmem_empty0 + 8u * (uint32_t)buf);
        if (++buf == 2) { buf = 0; acc_phase ^= 1; }
      }
    }
    if (lane == 0) bulk_wait_read<0>();               // staging tiles must outlive the stores that read them
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();
  if (warp == 1) tmem_dealloc<PAIR>(tmem_base, kTmemCols);
}

__global__ void __launch_bounds__(kThreads, 1)
resnet_gemm_kernel(const __grid_constant__ PairMaps maps, const __grid_constant__ EpiMaps emaps, EpilogueArgs ep, GemmParams gp) {
  gemm_body<false>(maps, emaps, ep, gp);
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
resnet_gemm_pair_kernel(const __grid_constant__ PairMaps maps, const __grid_constant__ EpiMaps emaps, EpilogueArgs ep, GemmParams gp) {
  gemm_body<true>(maps, emaps, ep, gp);
}

// ---- small kernels around the GEMMs ------------------------------------------------------------------
// one-hot encoding of the nnet input (pytorch_models.py:49-52) as an fp16 matrix [M][Kp], K index = s*depth + value
__global__ void __launch_bounds__(256) onehot_kernel(const uint8_t *__restrict__ x, int64_t M, int S, int depth, int Kp, __half *__restrict__ out) {
  const int64_t total = M * (int64_t)(Kp / 8);                 // one thread writes 8 halves (16 bytes)
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / (Kp / 8);
    const int k0 = (int)(i - m * (Kp / 8)) * 8;
    __half h[8];
#pragma unroll
    for (int e = 0; e < 8; e++) {
      const int k = k0 + e, s = k / depth, v = k - s * depth;
      h[e] = (s < S && x[m * S + s] == v) ? __float2half(1.0f) : __float2half(0.0f);
    }
    *reinterpret_cast<uint4 *>(out + m * Kp + k0) = *reinterpret_cast<const uint4 *>(h);
  }
}

// same, reading the states of the listed nodes straight from the search's node arena (state of node i at arena + i*S):
// state_to_nnet_input (cube3.py:77-85: sticker / 9 -> colour; cube4: sticker / 16) and F.one_hot in one pass, no intermediate u8 matrix
// DEPTH > 0: the one-hot depth as a compile-time constant (the division k / depth becomes a multiply-shift); 0: runtime depth.
template <int DIV, int DEPTH>
__global__ void __launch_bounds__(256) onehot_gather_kernel(const uint8_t *__restrict__ arena, const uint32_t *__restrict__ ids, int64_t M, int S,
                                                            int depth_rt, int Kp, __half *__restrict__ out, const int32_t *__restrict__ m_dev, int32_t m_off) {
  if (m_dev) {                                                  // device-side row count (see GemmParams::m_dev)
    const int64_t d = (int64_t)(*m_dev) - m_off;
    M = d < 0 ? 0 : (d < M ? d : M);
  }
  const int depth = DEPTH > 0 ? DEPTH : depth_rt;
  const int chunks = Kp / 8;                                    // one thread writes 8 halves (16 bytes)
  const int64_t total = M * (int64_t)chunks;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / chunks;
    const int k0 = (int)(i - m * chunks) * 8;
    const uint8_t *st = arena + (uint64_t)ids[m] * S;
    // the 8 columns k0..k0+7 belong to at most two positions when depth >= 8, to a handful otherwise: walk (s, v) incrementally
    int s = k0 / depth, v = k0 - s * depth;
    int x = -1;
    if (s < S) { x = st[s]; if (DIV == 9) x = (x * 57) >> 9; else if (DIV == 16) x >>= 4; }
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 8; e++) {
      const uint32_t bit = (x == v) ? 0x3C00u : 0u;             // fp16 1.0
      if (e & 1) w[e >> 1] |= bit << 16; else w[e >> 1] = bit;
      if (++v == depth) {
        v = 0; ++s;
        x = -1;
        if (s < S) { x = st[s]; if (DIV == 9) x = (x * 57) >> 9; else if (DIV == 16) x >>= 4; }
      }
    }
    *reinterpret_cast<uint4 *>(out + m * Kp + k0) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// fc_out (pytorch_models.py:85): out[m] = sum_n (hi+lo)[m][n] * w[n] + b, one warp per row, fp32.  Each lane reads 16 bytes
// (8 halves) of hi and lo per step; the summation order is fixed, so results are reproducible run to run.
__global__ void __launch_bounds__(256) rowdot_kernel(const __half *__restrict__ x_hi, const __half *__restrict__ x_lo, const float *__restrict__ w,
                                                     float bias, int64_t M, int n_valid, int ld, float *__restrict__ out) {
  const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  float acc = 0.0f;
  for (int n = lane * 8; n < n_valid; n += 256) {              // ld is a multiple of 8 and rows are 16-byte aligned
    const uint4 uh = *reinterpret_cast<const uint4 *>(x_hi + row * ld + n);
    uint4 ul = make_uint4(0, 0, 0, 0);
    if (x_lo) ul = *reinterpret_cast<const uint4 *>(x_lo + row * ld + n);
    const __half2 *hh = reinterpret_cast<const __half2 *>(&uh), *hl = reinterpret_cast<const __half2 *>(&ul);
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const float2 f = __half22float2(hh[e]), g = __half22float2(hl[e]);
      const int k = n + 2 * e;
      if (k < n_valid) acc = fmaf(f.x + g.x, w[k], acc);
      if (k + 1 < n_valid) acc = fmaf(f.y + g.y, w[k + 1], acc);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[row] = acc + bias;
}

// ---- host side ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// fp16 row-major [rows][cols] view with leading dimension ld (elements); box = [box_rows][64 cols], 128-byte swizzle,
// out-of-bounds rows read as zero
bool make_map(CUtensorMap *map, const void *ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

int64_t resnet_gemm_scratch_bytes() { return (int64_t)160 * 64 * 128 * sizeof(float4); }     // >= SM count CTAs x 128 KB

int resnet_gemm_device(const void *a_hi, const void *a_lo, int64_t lda, const void *w_hi, const void *w_lo, int64_t ldw, const float *bias,
                       float scale, const void *skip_hi, const void *skip_lo, int relu, void *out_hi, void *out_lo, float *out_f32,
                       const float *partial_in, float *partial_out, const float *dot_w, float *dot_partial, int64_t M, int Np, int Kp,
                       const int32_t *m_dev, int32_t m_off, int chunk_k, void *scratch, cudaStream_t st) {
  if (M == 0) return DCB_OK;
  if (Np % BN || Kp % BK) return DCB_ERR_BAD_ARG;
  if (chunk_k <= 0 || chunk_k > Kp) chunk_k = Kp;
  if (chunk_k % BK) return DCB_ERR_BAD_ARG;
  if (chunk_k < Kp && !scratch) return DCB_ERR_BAD_ARG;
  // CTA-pair (cta_group::2) kernel is the default; DCB_GEMM_PAIR=0 selects the single-CTA kernel (kept for A/B measurements).
  static int pair_mode = -1;
  if (pair_mode < 0) { const char *e = getenv("DCB_GEMM_PAIR"); pair_mode = (e && e[0] == '0') ? 0 : 1; }
  const int use_pair = pair_mode;
  const int w_box = use_pair ? BN / 2 : BN;
  CUtensorMap ma_hi, ma_lo, mw_hi, mw_lo;
  if (!make_map(&ma_hi, a_hi, M, Kp, lda, BM) || !make_map(&mw_hi, w_hi, Np, Kp, ldw, w_box)) return DCB_ERR_CUDA;
  if (a_lo && !make_map(&ma_lo, a_lo, M, Kp, lda, BM)) return DCB_ERR_CUDA;
  if (w_lo && !make_map(&mw_lo, w_lo, Np, Kp, ldw, w_box)) return DCB_ERR_CUDA;
  PairMaps maps;
  int n = 0;
  if (w_lo) { maps.a[n] = ma_hi; maps.w[n] = mw_lo; n++; }                 // small products first
  if (a_lo && w_lo) { maps.a[n] = ma_lo; maps.w[n] = mw_hi; n++; }
  maps.a[n] = ma_hi; maps.w[n] = mw_hi; n++;                               // main product last
  for (int i = n; i < kMaxPairs; i++) { maps.a[i] = ma_hi; maps.w[i] = mw_hi; }
  EpiMaps em;
  em.out_hi = em.out_lo = em.skip_hi = em.skip_lo = ma_hi;                 // placeholders for unused maps (never dereferenced)
  if (out_hi && !make_map(&em.out_hi, out_hi, M, Np, Np, 32)) return DCB_ERR_CUDA;
  if (out_lo && !make_map(&em.out_lo, out_lo, M, Np, Np, 32)) return DCB_ERR_CUDA;
  if (skip_hi && !make_map(&em.skip_hi, skip_hi, M, Np, Np, 32)) return DCB_ERR_CUDA;
  if (skip_lo && !make_map(&em.skip_lo, skip_lo, M, Np, Np, 32)) return DCB_ERR_CUDA;
  EpilogueArgs ep{bias, scale, (const __half *)skip_hi, (const __half *)skip_lo, relu, (__half *)out_hi, (__half *)out_lo, out_f32,
                  partial_in, partial_out, dot_w, dot_partial};
  GemmParams gp{n, M, m_dev, m_off, Np, Kp / BK, chunk_k / BK, chunk_k < Kp ? reinterpret_cast<float4 *>(scratch) : nullptr};
  static bool configured = false;
  static int sms = 148;
  if (!configured) {
    if (cudaFuncSetAttribute(resnet_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<false>()) != cudaSuccess) return dcb_cuda_fail();
    if (cudaFuncSetAttribute(resnet_gemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<true>()) != cudaSuccess) return dcb_cuda_fail();
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms > 160) sms = 160;                                              // scratch is sized for 160 CTAs
    configured = true;
  }
  if (use_pair) {
    const int64_t pair_tiles = ((M + 2 * BM - 1) / (2 * BM)) * (Np / BN);
    const int64_t pairs = pair_tiles < sms / 2 ? pair_tiles : sms / 2;
    resnet_gemm_pair_kernel<<<(unsigned)(2 * pairs), kThreads, smem_bytes<true>(), st>>>(maps, em, ep, gp);
    return dcb_check_launch();
  }
  const int64_t tiles = ((M + BM - 1) / BM) * (Np / BN);
  const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
  resnet_gemm_kernel<<<grid, kThreads, smem_bytes<false>(), st>>>(maps, em, ep, gp);
  return dcb_check_launch();
}

int onehot_device(const uint8_t *x, int64_t M, int S, int depth, int Kp, void *out, cudaStream_t st) {
  if (M == 0) return DCB_OK;
  int64_t blocks = (M * (Kp / 8) + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  onehot_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, M, S, depth, Kp, (__half *)out);
  return dcb_check_launch();
}

int onehot_gather_device(int env, const uint8_t *arena, const uint32_t *ids, int64_t M, int S, int depth, int Kp, void *out, const int32_t *m_dev,
                         int32_t m_off, cudaStream_t st) {
  if (M == 0) return DCB_OK;
  int64_t blocks = (M * (Kp / 8) + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
#define DCB_ONEHOT(DIV, DEPTH) onehot_gather_kernel<DIV, DEPTH><<<(unsigned)blocks, 256, 0, st>>>(arena, ids, M, S, depth, Kp, (__half *)out, m_dev, m_off)
  if (env == 0 && depth == 6) DCB_ONEHOT(9, 6);
  else if (env == 0) DCB_ONEHOT(9, 0);
  else if (env == DCB_ENV_CUBE4 && depth == 6) DCB_ONEHOT(16, 6);
  else if (env == DCB_ENV_CUBE4) DCB_ONEHOT(16, 0);
  else if (depth == 16) DCB_ONEHOT(0, 16);
  else if (depth == 25) DCB_ONEHOT(0, 25);
  else if (depth == 36) DCB_ONEHOT(0, 36);
  else if (depth == 49) DCB_ONEHOT(0, 49);
  else if (depth == 2) DCB_ONEHOT(0, 2);
  else DCB_ONEHOT(0, 0);
#undef DCB_ONEHOT
  return dcb_check_launch();
}

int rowdot_device(const void *x_hi, const void *x_lo, const float *w, float bias, int64_t M, int n_valid, int ld, float *out, cudaStream_t st) {
  if (M == 0) return DCB_OK;
  const unsigned blocks = (unsigned)((M * 32 + 255) / 256);
  rowdot_kernel<<<blocks, 256, 0, st>>>((const __half *)x_hi, (const __half *)x_lo, w, bias, M, n_valid, ld, out);
  return dcb_check_launch();
}

}  // namespace dcb
