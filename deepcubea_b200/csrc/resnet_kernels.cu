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

constexpr int STAGE_BYTES = A_TILE_BYTES + W_TILE_BYTES;   // one A tile + one W tile per pipeline stage (48 KB)
constexpr int STAGES = 4;
constexpr int EPI_STAGE_BYTES = 4 * 8192;               // per epilogue warp: 32 rows x 128 B, fp16 hi and lo
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int kMaxPairs = 3;

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
__device__ __forceinline__ void tma_load_2d(const CUtensorMap *map, uint64_t *bar, void *dst, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_addr(dst)),
               "l"(map), "r"(smem_addr(bar)), "r"(c0), "r"(c1)
               : "memory");
}
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
// TMA load whose completion bytes are credited to the mbarrier of the LEADER CTA of the pair (bit 24 of a shared::cluster
// address selects the CTA within the pair)
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap *map, uint32_t leader_bar_cluster_addr, void *dst, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_addr(dst)),
      "l"(map), "r"(leader_bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t *smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(smem_dst)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all MMAs issued so far have retired) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_addr(bar)),
               "h"((uint16_t)0x3)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(smem_dst)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}
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
// kind::f16 instruction descriptor: fp16 x fp16 -> fp32, both operands K-major, M=128, N=256
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

struct EpilogueArgs {
  const float *bias;          // [Np]
  float scale;                // acc * scale (undoes the power-of-two weight pre-scaling)
  const __half *skip_hi;      // [M][Np] or nullptr
  const __half *skip_lo;      // [M][Np] or nullptr
  int relu;
  __half *out_hi;             // [M][Np]
  __half *out_lo;             // [M][Np] or nullptr
  float *out_f32;             // [M][Np] or nullptr
  const float *partial_in;    // [M][Np] fp32 partial sum of earlier K chunks (unscaled) or nullptr
  float *partial_out;         // if set: write acc (+ partial_in) here and skip the rest of the epilogue
  const float *dot_w;         // [Np] or nullptr: fused fc_out -- dot_partial[row][n_tile] = sum over the tile's columns of out * dot_w
  float *dot_partial;         // [M][Np/256]; the caller adds the n_tile partials in a fixed order (reproducible)
};

struct PairMaps {              // operand pairs, swept in order; the last one is the main product
  CUtensorMap a[kMaxPairs];
  CUtensorMap w[kMaxPairs];
};

// While an epilogue warp waits for its accumulator it has nothing to do: pull the tile's residual input / fp32 partial sums
// (one row of 256 columns per thread) into L2, so the row-strided reads later hit L2 instead of DRAM.
__device__ __forceinline__ void prefetch_epilogue_inputs(const EpilogueArgs &ep, int64_t row, int n0, int64_t M, int Np) {
  if (row >= M) return;
  if (ep.skip_hi) {
    const char *h = reinterpret_cast<const char *>(ep.skip_hi + row * Np + n0);
#pragma unroll
    for (int q = 0; q < 4; q++) asm volatile("prefetch.global.L2 [%0];" ::"l"(h + 128 * q));
    if (ep.skip_lo) {
      const char *l = reinterpret_cast<const char *>(ep.skip_lo + row * Np + n0);
#pragma unroll
      for (int q = 0; q < 4; q++) asm volatile("prefetch.global.L2 [%0];" ::"l"(l + 128 * q));
    }
  }
  if (ep.partial_in) {
    const char *f = reinterpret_cast<const char *>(ep.partial_in + row * Np + n0);
#pragma unroll
    for (int q = 0; q < 8; q++) asm volatile("prefetch.global.L2 [%0];" ::"l"(f + 128 * q));
  }
}

// One accumulator tile (this warp's 32 TMEM lanes x 256 columns) -> global memory.  A thread owns one output row; results
// leave through a per-warp XOR-swizzled shared-memory tile so that eight lanes write one full 128-byte line of a row.
__device__ __forceinline__ void epilogue_tile(const EpilogueArgs &ep, uint32_t taddr, int64_t row0, int n0, int64_t M, int Np, int lane,
                                              uint8_t *stage_hi, uint8_t *stage_lo) {
  const int64_t row = row0 + lane;
  float dot = 0.0f;
      const bool row_ok = row < M;
#pragma unroll 1
for (int c = 0; c < BN; c += 64) {
  uint32_t acc0[32], acc1[32];
  tmem_ld_32x32(taddr + c, acc0);
  tmem_ld_32x32(taddr + c + 32, acc1);
  float v[64];
#pragma unroll
  for (int j = 0; j < 32; j++) { v[j] = __uint_as_float(acc0[j]); v[32 + j] = __uint_as_float(acc1[j]); }
  const int64_t off = row * Np + n0 + c;
  if (row_ok && ep.partial_in) {
    const float4 *pi = reinterpret_cast<const float4 *>(ep.partial_in + off);
#pragma unroll
    for (int q = 0; q < 16; q++) { const float4 f = pi[q]; v[4 * q] += f.x; v[4 * q + 1] += f.y; v[4 * q + 2] += f.z; v[4 * q + 3] += f.w; }
  }
  if (ep.partial_out) {                        // K-chunk chaining: raw fp32 sums only (warp-uniform branch)
    if (row_ok) {
      float4 *po = reinterpret_cast<float4 *>(ep.partial_out + off);
#pragma unroll
      for (int q = 0; q < 16; q++) po[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
    continue;
  }
#pragma unroll
  for (int j = 0; j < 64; j++) v[j] = fmaf(v[j], ep.scale, __ldg(ep.bias + n0 + c + j));
  if (row_ok && ep.skip_hi) {
    const uint4 *sh = reinterpret_cast<const uint4 *>(ep.skip_hi + off);
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const uint4 u = sh[q];
      const __half2 *h = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
      for (int e = 0; e < 4; e++) { const float2 f = __half22float2(h[e]); v[q * 8 + 2 * e] += f.x; v[q * 8 + 2 * e + 1] += f.y; }
    }
    if (ep.skip_lo) {
      const uint4 *sl = reinterpret_cast<const uint4 *>(ep.skip_lo + off);
#pragma unroll
      for (int q = 0; q < 8; q++) {
        const uint4 u = sl[q];
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
  if (ep.dot_w) {
#pragma unroll
    for (int j = 0; j < 64; j++) dot = fmaf(v[j], __ldg(ep.dot_w + n0 + c + j), dot);
  }
  if (!ep.out_hi) continue;
  // fp16 hi / lo split, staged (chunk q of row `lane` lives at 16-byte slot q ^ (lane & 7): conflict-free both ways)
  __syncwarp();                                  // the previous slice has been read out of the staging tile
#pragma unroll
  for (int q = 0; q < 8; q++) {
    uint4 ph, pl;
    __half2 *hh = reinterpret_cast<__half2 *>(&ph), *ll = reinterpret_cast<__half2 *>(&pl);
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const float a = fminf(fmaxf(v[8 * q + 2 * e], -65504.0f), 65504.0f), b = fminf(fmaxf(v[8 * q + 2 * e + 1], -65504.0f), 65504.0f);
      const __half2 h = __floats2half2_rn(a, b);
      const float2 hf = __half22float2(h);
      hh[e] = h;
      ll[e] = __floats2half2_rn(a - hf.x, b - hf.y);
    }
    const int slot = ((q ^ (lane & 7)) << 4) + lane * 128;
    *reinterpret_cast<uint4 *>(stage_hi + slot) = ph;
    *reinterpret_cast<uint4 *>(stage_lo + slot) = pl;
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int r = 4 * i + (lane >> 3), ch = lane & 7;
    const int slot = ((ch ^ (r & 7)) << 4) + r * 128;
    if (row0 + r < M) {
      const int64_t o = (row0 + r) * Np + n0 + c + ch * 8;
      *reinterpret_cast<uint4 *>(ep.out_hi + o) = *reinterpret_cast<const uint4 *>(stage_hi + slot);
      if (ep.out_lo) *reinterpret_cast<uint4 *>(ep.out_lo + o) = *reinterpret_cast<const uint4 *>(stage_lo + slot);
    }
  }
}
  if (ep.dot_w && row_ok) ep.dot_partial[row * (Np / BN) + n0 / BN] = dot;
}

__global__ void __launch_bounds__(kThreads, 1)
resnet_gemm_kernel(const __grid_constant__ PairMaps maps, int n_pairs, EpilogueArgs ep, int64_t M, int Np, int Kp) {
  extern __shared__ uint8_t smem_dyn[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  uint8_t *ep_stage = smem + STAGES * STAGE_BYTES;
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(ep_stage + EPI_STAGE_BYTES);
  uint64_t *empty_bar = full_bar + STAGES;
  uint64_t *tmem_full = empty_bar + STAGES;          // [2]
  uint64_t *tmem_empty = tmem_full + 2;              // [2]
  uint32_t *tmem_base_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = Np / BN, k_blocks = Kp / BK;
  const int64_t m_tiles = (M + BM - 1) / BM;
  const int64_t total_tiles = m_tiles * n_tiles;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; b++) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 4); }   // 4 epilogue warps arrive
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_base_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int64_t t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int32_t m0 = (int32_t)((t / n_tiles) * BM), n0 = (int32_t)((t % n_tiles) * BN);
        for (int p = 0; p < n_pairs; p++) {
          for (int kb = 0; kb < k_blocks; kb++) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t *st = smem + stage * STAGE_BYTES;
            mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
            tma_load_2d(&maps.a[p], &full_bar[stage], st, kb * BK, m0);
            tma_load_2d(&maps.w[p], &full_bar[stage], st + A_TILE_BYTES, kb * BK, n0);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int buf = 0; uint32_t acc_phase = 0;
      for (int64_t t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        mbar_wait(&tmem_empty[buf], acc_phase ^ 1);            // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + (uint32_t)buf * BN;
        uint32_t first = 1;
        for (int p = 0; p < n_pairs; p++) {
          for (int kb = 0; kb < k_blocks; kb++) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t st = smem_addr(smem + stage * STAGE_BYTES);
            const uint64_t da = make_smem_desc(st), dw = make_smem_desc(st + A_TILE_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; k++) {
              const uint64_t koff = (uint64_t)((k * UMMA_K * 2) >> 4);      // 32 bytes per K step inside the swizzle row
              umma_f16(tmem_acc, da + koff, dw + koff, kIdesc, first ? 0u : 1u);
              first = 0;
            }
            umma_commit(&empty_bar[stage]);                     // frees the smem stage when these MMAs retire
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
        umma_commit(&tmem_full[buf]);                           // accumulator complete
        if (++buf == 2) { buf = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue: 4 warps, warp (w % 4) owns TMEM lanes 32*(w%4) .. +31 =====================
    // A thread owns one output row (TMEM lane); results leave through a per-warp XOR-swizzled shared-memory tile so that
    // eight lanes write one full 128-byte line of a row (fp16 hi / lo) instead of 32 lanes writing 16-byte fragments.
    const int lane_grp = warp & 3;
    uint8_t *stage_hi = ep_stage + (warp - 2) * 8192, *stage_lo = stage_hi + 4096;
    int buf = 0; uint32_t acc_phase = 0;
    for (int64_t t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int64_t row0 = (t / n_tiles) * BM + lane_grp * 32;
      const int n0 = (int)((t % n_tiles) * BN);
      prefetch_epilogue_inputs(ep, row0 + lane, n0, M, Np);
      mbar_wait(&tmem_full[buf], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)buf * BN;
      const int64_t row = row0 + lane;
      const bool row_ok = row < M;
      float dot = 0.0f;
#pragma unroll 1
      for (int c = 0; c < BN; c += 64) {
        uint32_t acc0[32], acc1[32];
        tmem_ld_32x32(taddr + c, acc0);
        tmem_ld_32x32(taddr + c + 32, acc1);
        float v[64];
#pragma unroll
        for (int j = 0; j < 32; j++) { v[j] = __uint_as_float(acc0[j]); v[32 + j] = __uint_as_float(acc1[j]); }
        const int64_t off = row * Np + n0 + c;
        if (row_ok && ep.partial_in) {
          const float4 *pi = reinterpret_cast<const float4 *>(ep.partial_in + off);
#pragma unroll
          for (int q = 0; q < 16; q++) { const float4 f = pi[q]; v[4 * q] += f.x; v[4 * q + 1] += f.y; v[4 * q + 2] += f.z; v[4 * q + 3] += f.w; }
        }
        if (ep.partial_out) {                        // K-chunk chaining: raw fp32 sums only (warp-uniform branch)
          if (row_ok) {
            float4 *po = reinterpret_cast<float4 *>(ep.partial_out + off);
#pragma unroll
            for (int q = 0; q < 16; q++) po[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          }
          continue;
        }
#pragma unroll
        for (int j = 0; j < 64; j++) v[j] = fmaf(v[j], ep.scale, __ldg(ep.bias + n0 + c + j));
        if (row_ok && ep.skip_hi) {
          const uint4 *sh = reinterpret_cast<const uint4 *>(ep.skip_hi + off);
#pragma unroll
          for (int q = 0; q < 8; q++) {
            const uint4 u = sh[q];
            const __half2 *h = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
            for (int e = 0; e < 4; e++) { const float2 f = __half22float2(h[e]); v[q * 8 + 2 * e] += f.x; v[q * 8 + 2 * e + 1] += f.y; }
          }
          if (ep.skip_lo) {
            const uint4 *sl = reinterpret_cast<const uint4 *>(ep.skip_lo + off);
#pragma unroll
            for (int q = 0; q < 8; q++) {
              const uint4 u = sl[q];
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
          for (int j = 0; j < 64; j++) dot = fmaf(v[j], __ldg(ep.dot_w + n0 + c + j), dot);
        }
        if (!ep.out_hi) continue;                      // last layer: only the dot product leaves the kernel (warp-uniform)
        // fp16 hi / lo split, staged (chunk q of row `lane` lives at 16-byte slot q ^ (lane & 7): conflict-free both ways)
        __syncwarp();                                  // the previous slice has been read out of the staging tile
#pragma unroll
        for (int q = 0; q < 8; q++) {
          uint4 ph, pl;
          __half2 *hh = reinterpret_cast<__half2 *>(&ph), *ll = reinterpret_cast<__half2 *>(&pl);
#pragma unroll
          for (int e = 0; e < 4; e++) {
            const float a = fminf(fmaxf(v[8 * q + 2 * e], -65504.0f), 65504.0f), b = fminf(fmaxf(v[8 * q + 2 * e + 1], -65504.0f), 65504.0f);
            const __half2 h = __floats2half2_rn(a, b);
            const float2 hf = __half22float2(h);
            hh[e] = h;
            ll[e] = __floats2half2_rn(a - hf.x, b - hf.y);
          }
          const int slot = ((q ^ (lane & 7)) << 4) + lane * 128;
          *reinterpret_cast<uint4 *>(stage_hi + slot) = ph;
          *reinterpret_cast<uint4 *>(stage_lo + slot) = pl;
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const int r = 4 * i + (lane >> 3), ch = lane & 7;
          const int slot = ((ch ^ (r & 7)) << 4) + r * 128;
          if (row0 + r < M) {
            const int64_t o = (row0 + r) * Np + n0 + c + ch * 8;
            *reinterpret_cast<uint4 *>(ep.out_hi + o) = *reinterpret_cast<const uint4 *>(stage_hi + slot);
            if (ep.out_lo) *reinterpret_cast<uint4 *>(ep.out_lo + o) = *reinterpret_cast<const uint4 *>(stage_lo + slot);
          }
        }
      }
      if (ep.dot_w && row_ok) ep.dot_partial[row * n_tiles + (int)(t % n_tiles)] = dot;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);
      if (++buf == 2) { buf = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// =====================================================================================================================
// CTA-pair variant (tcgen05 cta_group::2): two CTAs of a cluster compute one 256 x 256 tile.  Each CTA stages its own 128
// A rows and HALF of the W tile (128 of the 256 N rows); the pair MMA (M = 256) issued by the leader reads both halves,
// so a CTA stages 32 KB instead of 48 KB per four MMAs -- a third more MMA time per staged byte, which is what bounds the
// single-CTA kernel (its 4 x 48 KB pipeline covers ~1 us of MMA work against ~1.4 us of load latency).
// Barriers: full[s] lives in the LEADER (its arrive.expect_tx covers both CTAs' bytes; both CTAs' TMA loads credit it), empty[s] and tmem_full[b] are signalled in both CTAs by the leader's multicast commits,
// tmem_empty[b] (leader) collects the 2 x 4 epilogue warps of both CTAs.
// =====================================================================================================================
constexpr int P_STAGE_BYTES = A_TILE_BYTES + W_TILE_BYTES / 2;      // 32 KB per CTA
constexpr int P_STAGES = 6;
constexpr int P_SMEM_BYTES = P_STAGES * P_STAGE_BYTES + EPI_STAGE_BYTES + 1024 + 256;
constexpr uint32_t kIdescPair = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
resnet_gemm_pair_kernel(const __grid_constant__ PairMaps maps, int n_pairs, EpilogueArgs ep, int64_t M, int Np, int Kp) {
  extern __shared__ uint8_t smem_dyn[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  uint8_t *ep_stage = smem + P_STAGES * P_STAGE_BYTES;
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(ep_stage + EPI_STAGE_BYTES);
  uint64_t *empty_bar = full_bar + P_STAGES;
  uint64_t *tmem_full = empty_bar + P_STAGES;        // [2]
  uint64_t *tmem_empty = tmem_full + 2;              // [2]
  uint32_t *tmem_base_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int n_tiles = Np / BN, k_blocks = Kp / BK;
  const int64_t m_pairs = (M + 2 * BM - 1) / (2 * BM);
  const int64_t total_tiles = m_pairs * n_tiles;
  const int64_t tile_first = blockIdx.x >> 1, tile_step = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < P_STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; b++) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 8); }   // 4 epilogue warps x 2 CTAs
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(tmem_base_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const uint32_t leader_full0 = map_to_cta(&full_bar[0], 0);
      for (int64_t t = tile_first; t < total_tiles; t += tile_step) {
        const int32_t m0 = (int32_t)((2 * (t / n_tiles) + rank) * BM), n0 = (int32_t)((t % n_tiles) * BN + rank * (BN / 2));
        for (int p = 0; p < n_pairs; p++) {
          for (int kb = 0; kb < k_blocks; kb++) {
            mbar_wait(&empty_bar[stage], phase ^ 1);                         // own stage buffer is free (leader's commit reaches both CTAs)
            uint8_t *st = smem + stage * P_STAGE_BYTES;
            // bytes of both CTAs land on the leader's barrier.  The peer does not arrive on it: a release-arrive across
            // the cluster costs its producer ~0.5 us per stage (ncu: 43% of that warp's samples in the fence) and the
            // leader's own arrive.expect_tx already keeps the phase open until every byte is in.
            if (leader) mbar_expect_tx(&full_bar[stage], 2 * P_STAGE_BYTES);
            tma_load_2d_pair(&maps.a[p], leader_full0 + 8u * (uint32_t)stage, st, kb * BK, m0);
            tma_load_2d_pair(&maps.w[p], leader_full0 + 8u * (uint32_t)stage, st + A_TILE_BYTES, kb * BK, n0);
            if (++stage == P_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: one thread of the leader CTA =====================
    if (leader && lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int buf = 0; uint32_t acc_phase = 0;
      for (int64_t t = tile_first; t < total_tiles; t += tile_step) {
        mbar_wait(&tmem_empty[buf], acc_phase ^ 1);            // both CTAs' epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + (uint32_t)buf * BN;
        uint32_t first = 1;
        for (int p = 0; p < n_pairs; p++) {
          for (int kb = 0; kb < k_blocks; kb++) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t st = smem_addr(smem + stage * P_STAGE_BYTES);
            const uint64_t da = make_smem_desc(st), dw = make_smem_desc(st + A_TILE_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; k++) {
              const uint64_t koff = (uint64_t)((k * UMMA_K * 2) >> 4);
              umma_f16_pair(tmem_acc, da + koff, dw + koff, kIdescPair, first ? 0u : 1u);
              first = 0;
            }
            umma_commit_pair(&empty_bar[stage]);                // stage free in both CTAs once these MMAs retire
            if (++stage == P_STAGES) { stage = 0; phase ^= 1; }
          }
        }
        umma_commit_pair(&tmem_full[buf]);                      // accumulators of both CTAs complete
        if (++buf == 2) { buf = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (both CTAs, own 128 rows each) =====================
    const int lane_grp = warp & 3;
    uint8_t *stage_hi = ep_stage + (warp - 2) * 8192, *stage_lo = stage_hi + 4096;
    const uint32_t leader_tmem_empty0 = map_to_cta(&tmem_empty[0], 0);
    int buf = 0; uint32_t acc_phase = 0;
    for (int64_t t = tile_first; t < total_tiles; t += tile_step) {
      const int64_t row0 = (2 * (t / n_tiles) + rank) * BM + lane_grp * 32;
      const int n0 = (int)((t % n_tiles) * BN);
      prefetch_epilogue_inputs(ep, row0 + lane, n0, M, Np);
      mbar_wait(&tmem_full[buf], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)buf * BN;
      epilogue_tile(ep, taddr, row0, n0, M, Np, lane, stage_hi, stage_lo);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(leader_tmem_empty0 + 8u * (uint32_t)buf);
      if (++buf == 2) { buf = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_pair(tmem_base, kTmemCols);
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
template <int DIV>
__global__ void __launch_bounds__(256) onehot_gather_kernel(const uint8_t *__restrict__ arena, const uint32_t *__restrict__ ids, int64_t M, int S,
                                                            int depth, int Kp, __half *__restrict__ out) {
  const int64_t total = M * (int64_t)(Kp / 8);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / (Kp / 8);
    const int k0 = (int)(i - m * (Kp / 8)) * 8;
    const uint8_t *st = arena + (uint64_t)ids[m] * S;
    __half h[8];
#pragma unroll
    for (int e = 0; e < 8; e++) {
      const int k = k0 + e, s = k / depth, v = k - s * depth;
      int x = -1;
      if (s < S) { x = st[s]; if (DIV == 9) x = (x * 57) >> 9; else if (DIV == 16) x >>= 4; }
      h[e] = (x == v) ? __float2half(1.0f) : __float2half(0.0f);
    }
    *reinterpret_cast<uint4 *>(out + m * Kp + k0) = *reinterpret_cast<const uint4 *>(h);
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

int resnet_gemm_device(const void *a_hi, const void *a_lo, int64_t lda, const void *w_hi, const void *w_lo, int64_t ldw, const float *bias,
                       float scale, const void *skip_hi, const void *skip_lo, int relu, void *out_hi, void *out_lo, float *out_f32,
                       const float *partial_in, float *partial_out, const float *dot_w, float *dot_partial, int64_t M, int Np, int Kp,
                       cudaStream_t st) {
  if (M == 0) return DCB_OK;
  if (Np % BN || Kp % BK) return DCB_ERR_BAD_ARG;
  // CTA-pair (cta_group::2) kernel: DCB_GEMM_PAIR=1 always, 2 = only for launches whose epilogue neither reads a residual
  // input nor chains fp32 partial sums, default 0 = never.  Measured (r01): the pair MMA itself is ~25% faster (tensor pipe 91.6%
  // active on a K=N=1024 layer without residual, 513 vs 594 us) but with the present epilogue the whole network is not
  // (10.4 vs 10.3 ms): the epilogue, not the MMA, is the next thing to fix.
  static int pair_mode = -1;
  if (pair_mode < 0) { const char *e = getenv("DCB_GEMM_PAIR"); pair_mode = !e ? 0 : (e[0] == '1' ? 1 : (e[0] == '2' ? 2 : 0)); }
  const int use_pair = pair_mode == 1 || (pair_mode == 2 && !skip_hi && !partial_in && !partial_out);
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
  EpilogueArgs ep{bias, scale, (const __half *)skip_hi, (const __half *)skip_lo, relu, (__half *)out_hi, (__half *)out_lo, out_f32,
                  partial_in, partial_out, dot_w, dot_partial};
  static bool configured = false;
  static int sms = 148;
  if (!configured) {
    if (cudaFuncSetAttribute(resnet_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) return dcb_cuda_fail();
    if (cudaFuncSetAttribute(resnet_gemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_BYTES) != cudaSuccess) return dcb_cuda_fail();
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    configured = true;
  }
  if (use_pair) {
    const int64_t pair_tiles = ((M + 2 * BM - 1) / (2 * BM)) * (Np / BN);
    const int64_t pairs = pair_tiles < sms / 2 ? pair_tiles : sms / 2;
    resnet_gemm_pair_kernel<<<(unsigned)(2 * pairs), kThreads, P_SMEM_BYTES, st>>>(maps, n, ep, M, Np, Kp);
    return dcb_check_launch();
  }
  const int64_t tiles = ((M + BM - 1) / BM) * (Np / BN);
  const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
  resnet_gemm_kernel<<<grid, kThreads, SMEM_BYTES, st>>>(maps, n, ep, M, Np, Kp);
  return dcb_check_launch();
}

int onehot_device(const uint8_t *x, int64_t M, int S, int depth, int Kp, void *out, cudaStream_t st) {
  if (M == 0) return DCB_OK;
  int64_t blocks = (M * (Kp / 8) + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  onehot_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, M, S, depth, Kp, (__half *)out);
  return dcb_check_launch();
}

int onehot_gather_device(int env, const uint8_t *arena, const uint32_t *ids, int64_t M, int S, int depth, int Kp, void *out, cudaStream_t st) {
  if (M == 0) return DCB_OK;
  int64_t blocks = (M * (Kp / 8) + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  if (env == 0) onehot_gather_kernel<9><<<(unsigned)blocks, 256, 0, st>>>(arena, ids, M, S, depth, Kp, (__half *)out);
  else if (env == DCB_ENV_CUBE4) onehot_gather_kernel<16><<<(unsigned)blocks, 256, 0, st>>>(arena, ids, M, S, depth, Kp, (__half *)out);
  else onehot_gather_kernel<0><<<(unsigned)blocks, 256, 0, st>>>(arena, ids, M, S, depth, Kp, (__half *)out);
  return dcb_check_launch();
}

int rowdot_device(const void *x_hi, const void *x_lo, const float *w, float bias, int64_t M, int n_valid, int ld, float *out, cudaStream_t st) {
  if (M == 0) return DCB_OK;
  const unsigned blocks = (unsigned)((M * 32 + 255) / 256);
  rowdot_kernel<<<blocks, 256, 0, st>>>((const __half *)x_hi, (const __half *)x_lo, w, bias, M, n_valid, ld, out);
  return dcb_check_launch();
}

}  // namespace dcb
