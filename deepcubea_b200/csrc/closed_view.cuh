// closed_view.cuh -- how the CLOSED kernels see a batch of candidate nodes.
//   ContigView : candidates i = 0..m-1 are the nodes first_id + i with explicit hash / g / valid arrays (dcb_closed_insert).
//   TileView   : candidates of one search iteration, addressed through its tile list (include/dcb.h, dcb_search_*): candidate
//                c = child (c % A) of parent lane (c / A) % 32 of tile c / (32*A); the count lives in device memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/dcb.h"

namespace dcb {

struct Cand {
  uint32_t id, g;
  uint64_t key;        // table key: state hash mixed with the instance number, never 0
};

struct ClosedScratch {
  uint32_t *slot;                 // [m] slot index found by the probe (0xffffffff: table full)
  unsigned long long *prev;       // [m] the slot's value before this batch
  uint4 *amb;                     // [m] {slot, id, g, candidate index} of the candidates left to the in-batch fix-up
  uint32_t *counters;             // [0] kept, [1] ambiguous (stand-alone form)
};

struct ContigView {
  const uint64_t *hash;
  const uint32_t *g;
  const uint8_t *valid;
  uint32_t first_id;
  int64_t m;
  __device__ __forceinline__ int64_t count() const { return m; }
  __device__ __forceinline__ bool get(int64_t i, Cand &c) const {
    if (valid && !valid[i]) return false;
    c.id = first_id + (uint32_t)i;
    c.g = g[i];
    c.key = hash[i] ? hash[i] : 1ull;
    return true;
  }
  __device__ __forceinline__ bool same_inst(uint32_t, uint32_t) const { return true; }
};

// instance number -> key mixer (odd multiplier: a bijection of the 64-bit hash space per instance)
__host__ __device__ __forceinline__ uint64_t inst_mix(uint32_t inst) { return (uint64_t)inst * 0xD6E8FEB86659FD93ull; }

struct TileView {
  const uint4 *tiles;             // {src, dst_slot, count, inst}
  const dcb_step_plan *plan;
  const uint64_t *hash;           // [candidate]
  const uint32_t *node_g;         // [node id]
  uint32_t A;                     // moves
  uint32_t nodes_per_inst;        // slots_per_inst * A
  __device__ __forceinline__ int64_t count() const { return (int64_t)plan->n_tiles * 32 * A; }
  __device__ __forceinline__ bool get(int64_t i, Cand &c) const {
    const uint32_t per_tile = 32u * A;
    const uint32_t t = (uint32_t)(i / per_tile), r = (uint32_t)(i - (int64_t)t * per_tile);
    const uint32_t lane = r / A, a = r - lane * A;
    const uint4 d = tiles[t];
    if (lane >= d.z) return false;
    c.id = (d.y + lane) * A + a;
    c.g = node_g[c.id];
    const uint64_t k = hash[i] ^ inst_mix(d.w);
    c.key = k ? k : 1ull;
    return true;
  }
  __device__ __forceinline__ bool same_inst(uint32_t a, uint32_t b) const { return a / nodes_per_inst == b / nodes_per_inst; }
};

}  // namespace dcb
