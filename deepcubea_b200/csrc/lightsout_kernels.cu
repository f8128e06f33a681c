// lightsout_kernels.cu -- Lights Out 7x7 environment step (environments/lights_out.py:26-166,
// cpp/environments.cpp:133-208).  A state is 49 cells of 0/1 = one 64-bit word; the 49 children of a parent are
// `parent_bits ^ press_mask(m)`.  A CTA owns a tile of 16 parents (16 x 2401 bytes of children = 2401 x 16 bytes, so every
// tile starts 16-byte aligned):
//   phase 0  one thread per parent: 49 bytes -> bits (aligned word loads + multiply trick)
//   phase 1  one thread per child : bits, hash (same NH hash as the other environments, over the byte form) and solved flag
//   phase 2  one thread per 16 output bytes: bits of the one or two children that cover them -> bytes, coalesced 16-byte stores
#include <cuda_runtime.h>
#include "dcb_internal.h"
#include "expand_core.cuh"
#include "ptx.cuh"

namespace dcb {
namespace {
constexpr int LO = 5, S = 49, A = 49, W = 14, DIM = 7;
constexpr int TILE_P = 16, TILE_C = TILE_P * A, TILE_BYTES = TILE_C * S;     // 784 children, 38416 bytes

__device__ __forceinline__ uint64_t load_state_bits(const uint8_t *base, uint64_t off) {
  uint32_t raw[LoadShape<S>::NRAW], w[W];
  const uint32_t *a = reinterpret_cast<const uint32_t *>(base + (off & ~uint64_t(3)));
#pragma unroll
  for (int k = 0; k < LoadShape<S>::NRAW; k++) raw[k] = ldg_nc_u32(a + k);
  align_state<S, W>(raw, (uint32_t)(off & 3), w);
  return lo_bits_from_words<W>(w);
}

// cells [first, first+count) of the children stream starting at tile child c0: 4 cells -> one output word
__device__ __forceinline__ uint32_t word_at(const uint64_t *cbits, int byte_in_tile) {
  const int c = byte_in_tile / S, j = byte_in_tile - c * S;
  uint64_t x = cbits[c] >> j;
  if (j + 4 > S) x |= cbits[c + 1] << (S - j);            // straddles into the next child (cbits has one pad entry)
  return lo_spread4((uint32_t)x);
}

// PLANNED (dcb_search_expand): a 32-parent tile of the iteration's tile list = two 16-parent CTA tiles; parents ids[src + k],
// children to arena slots dst_slot + k (dst_slot is a multiple of 16), hashes in tile numbering, depth / parent link written here.
template <bool INDEXED, bool PLANNED>
__global__ void __launch_bounds__(256)
lightsout_expand_kernel(const uint8_t *__restrict__ src, const uint32_t *__restrict__ ids, int64_t n, uint8_t *__restrict__ children,
                        uint8_t *__restrict__ solved, uint64_t *__restrict__ hash, const uint4 *__restrict__ tiles,
                        const dcb_step_plan *__restrict__ plan, uint32_t *__restrict__ node_g, uint32_t *__restrict__ slot_parent) {
  __shared__ uint64_t pbits[TILE_P];
  __shared__ uint64_t cbits[TILE_C + 1];
  const int64_t n_tiles = PLANNED ? 2 * (int64_t)plan->n_tiles : (n + TILE_P - 1) / TILE_P;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    int64_t p0 = tile * TILE_P;        // first parent in the output numbering (children / solved)
    int64_t h0 = p0;                   // ... in the hash numbering
    int64_t s0 = p0;                   // ... in ids[]
    int np;
    if constexpr (PLANNED) {
      const uint4 d = tiles[tile >> 1];
      const int half = (int)(tile & 1) * TILE_P;
      np = (int)d.z - half;
      np = np < 0 ? 0 : (np > TILE_P ? TILE_P : np);
      p0 = (int64_t)d.y + half;
      h0 = (tile >> 1) * 32 + half;
      s0 = (int64_t)d.x + half;
    } else {
      np = (int)((n - p0) < TILE_P ? (n - p0) : TILE_P);
    }
    __syncthreads();                                       // previous tile fully written out
    if (threadIdx.x < np) {
      const uint64_t node = INDEXED ? (uint64_t)ids[s0 + threadIdx.x] : (uint64_t)(p0 + threadIdx.x);
      pbits[threadIdx.x] = load_state_bits(src, node * S);
      if constexpr (PLANNED) slot_parent[p0 + threadIdx.x] = (uint32_t)node;
    }
    if (threadIdx.x == 0) cbits[np * A] = 0;               // pad entry read by the last straddling word
    __syncthreads();
    for (int c = threadIdx.x; c < np * A; c += blockDim.x) {
      const int p = c / A, m = c - p * A;
      const uint64_t b = pbits[p] ^ lo_press_mask<DIM>(m);
      cbits[c] = b;
      uint32_t w[W];
      lo_words_from_bits<S, W>(b, w);
      if (hash) hash[(h0 + p) * A + m] = state_hash<W>(w);
      if (solved) solved[(p0 + p) * A + m] = (b == 0) ? 1 : 0;
      if constexpr (PLANNED) node_g[(p0 + p) * A + m] = node_g[ids[s0 + p]] + 1;
    }
    __syncthreads();
    uint8_t *out = children + p0 * (int64_t)(A * S);        // 16-byte aligned: 16 parents x 2401 bytes
    const int bytes = np * A * S;
    const int n_vec = bytes >> 4;
    for (int v = threadIdx.x; v < n_vec; v += blockDim.x) {
      uint4 o;
      o.x = word_at(cbits, 16 * v); o.y = word_at(cbits, 16 * v + 4); o.z = word_at(cbits, 16 * v + 8); o.w = word_at(cbits, 16 * v + 12);
      reinterpret_cast<uint4 *>(out)[v] = o;
    }
    for (int b = (n_vec << 4) + threadIdx.x; b < bytes; b += blockDim.x) {   // partial last tile: up to 15 tail bytes
      const int c = b / S, j = b - c * S;
      out[b] = (uint8_t)((cbits[c] >> j) & 1);
    }
  }
}
}  // namespace

int lightsout_expand_device(const uint8_t *src, const uint32_t *ids, int64_t n, uint8_t *children, uint8_t *solved, uint64_t *hash,
                            cudaStream_t st) {
  if (n == 0) return DCB_OK;
  int64_t blocks = (n + TILE_P - 1) / TILE_P;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (ids) lightsout_expand_kernel<true, false><<<(unsigned)blocks, 256, 0, st>>>(src, ids, n, children, solved, hash, nullptr, nullptr, nullptr, nullptr);
  else lightsout_expand_kernel<false, false><<<(unsigned)blocks, 256, 0, st>>>(src, nullptr, n, children, solved, hash, nullptr, nullptr, nullptr, nullptr);
  return dcb_check_launch();
}

int lightsout_expand_planned_device(uint8_t *arena, const uint32_t *ids, int64_t max_tiles, const uint32_t *tiles, const dcb_step_plan *plan,
                                    uint8_t *node_solved, uint64_t *hash, uint32_t *node_g, uint32_t *slot_parent, cudaStream_t st) {
  int64_t blocks = 2 * max_tiles;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  lightsout_expand_kernel<true, true><<<(unsigned)blocks, 256, 0, st>>>(arena, ids, 0, arena, node_solved, hash, reinterpret_cast<const uint4 *>(tiles),
                                                                        plan, node_g, slot_parent);
  return dcb_check_launch();
}
}  // namespace dcb
