// closed_table.cu -- CLOSED set as an open-addressing hash table in HBM.
//
// Replaces std::unordered_set<Node*,Hash,NodePointerEq> closed and the SERIAL find / insert /
// "(*found)->depth > node->depth -> overwrite and re-open" loop of cpp/parallel_weighted_astar.cpp:142,
// 243-265 (the reference's `check:` phase, ~12 ms per 24k children on 8 cores), and Instance.closed_dict /
// remove_in_closed of search_methods/astar.py:55, 78-90.
//
// Slot = 16 bytes {u64 key = state hash (0 = empty); u64 val = (g << 32) | node_id (all-ones = unset)}.
// Insert-or-improve for a whole batch runs as two launches:
//   1. insert : linear probing; atomicCAS claims/locates the key, atomicMin folds (g,id) into val --
//               "strictly smaller g wins, ties keep the older node" falls out of the (g,id) ordering
//               because node ids only grow.
//   2. resolve: a candidate is kept iff its own (g,id) is what the slot now holds.  A candidate that lost
//               is compared byte-for-byte with the winner's state in the arena: equal -> true duplicate,
//               dropped; different -> 64-bit hash collision, kept (a non-duplicate is never dropped).
// HBM traffic per candidate: 8 B hash + 4 B g in, one 32-B sector probe (~1.3 probes at load <= 0.5),
// 16 B slot update, 4 B slot index + 1 B keep out.
#include <cuda_runtime.h>
#include "dcb_internal.h"
#include "state_ops.cuh"
#include "ptx.cuh"

namespace dcb {

__global__ void __launch_bounds__(256) closed_clear_kernel(ulonglong2 *tbl, int64_t cap) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < cap; i += (int64_t)gridDim.x * blockDim.x)
    tbl[i] = make_ulonglong2(0ull, ~0ull);
}

__global__ void __launch_bounds__(256)
closed_insert_kernel(unsigned long long *__restrict__ tbl, uint64_t mask, const uint64_t *__restrict__ hash,
                     const uint32_t *__restrict__ g, const uint8_t *__restrict__ valid, uint32_t first_id, int64_t m,
                     uint32_t *__restrict__ slot_out, uint32_t *__restrict__ num_entries) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  bool claimed = false;
  if (i < m) slot_out[i] = 0xFFFFFFFFu;   // "no slot": table full (resolve keeps such candidates)
  if (i < m && (!valid || valid[i])) {
    const uint64_t h = hash[i];
    const unsigned long long myval = ((unsigned long long)g[i] << 32) | (unsigned long long)(first_id + (uint32_t)i);
    uint64_t s = h & mask;
    for (uint64_t probes = 0; probes <= mask; probes++) {
      unsigned long long k = tbl[2 * s];              // cheap read first: most probes hit an occupied slot
      if (k == 0ull) k = atomicCAS(&tbl[2 * s], 0ull, (unsigned long long)h);
      if (k == 0ull || k == h) {
        claimed = (k == 0ull);
        atomicMin(&tbl[2 * s + 1], myval);
        slot_out[i] = (uint32_t)s;
        break;
      }
      s = (s + 1) & mask;
    }
  }
  // one atomic per warp for the entry counter
  const unsigned ballot = __ballot_sync(0xffffffffu, claimed);
  if (num_entries && ballot && (threadIdx.x & 31) == 0) atomicAdd(num_entries, __popc(ballot));
}

template <int ENV>
__global__ void __launch_bounds__(256)
closed_resolve_kernel(const unsigned long long *__restrict__ tbl, const uint8_t *__restrict__ arena,
                      const uint32_t *__restrict__ g, const uint8_t *__restrict__ valid, uint32_t first_id, int64_t m,
                      const uint32_t *__restrict__ slot_in, uint8_t *__restrict__ keep) {
  constexpr int S = EnvTraits<ENV>::S, W = hash_words(S);
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= m) return;
  if (valid && !valid[i]) { keep[i] = 0; return; }
  const uint32_t my_id = first_id + (uint32_t)i;
  const unsigned long long myval = ((unsigned long long)g[i] << 32) | my_id;
  if (slot_in[i] == 0xFFFFFFFFu) { keep[i] = 1; return; }
  const unsigned long long win = tbl[2 * (uint64_t)slot_in[i] + 1];
  uint8_t k = 1;
  if (win != myval) {
    // lost to an older / cheaper node with the same hash: verify it really is the same state
    const uint32_t wid = (uint32_t)win;
    uint32_t ra[LoadShape<S>::NRAW], rb[LoadShape<S>::NRAW], a[W], b[W];
    const uint64_t oa = (uint64_t)my_id * S, ob = (uint64_t)wid * S;
    const uint32_t *pa = reinterpret_cast<const uint32_t *>(arena + (oa & ~uint64_t(3)));
    const uint32_t *pb = reinterpret_cast<const uint32_t *>(arena + (ob & ~uint64_t(3)));
#pragma unroll
    for (int q = 0; q < LoadShape<S>::NRAW; q++) { ra[q] = pa[q]; rb[q] = pb[q]; }
    align_state<S, W>(ra, (uint32_t)(oa & 3), a);
    align_state<S, W>(rb, (uint32_t)(ob & 3), b);
    uint32_t diff = 0;
#pragma unroll
    for (int q = 0; q < W; q++) diff |= a[q] ^ b[q];
    k = diff ? 1 : 0;
  }
  keep[i] = k;
}

__global__ void __launch_bounds__(256)
closed_rehash_kernel(const ulonglong2 *__restrict__ old_tbl, int64_t old_cap, unsigned long long *__restrict__ new_tbl, uint64_t new_mask) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < old_cap; i += (int64_t)gridDim.x * blockDim.x) {
    const ulonglong2 e = old_tbl[i];
    if (e.x == 0ull) continue;
    uint64_t s = e.x & new_mask;
    for (uint64_t probes = 0; probes <= new_mask; probes++) {
      // hashes are unique in the old table, so an empty slot is the only possible landing spot
      if (atomicCAS(&new_tbl[2 * s], 0ull, e.x) == 0ull) { new_tbl[2 * s + 1] = e.y; break; }
      s = (s + 1) & new_mask;
    }
  }
}

int closed_rehash_device(const void *old_tbl, int64_t old_cap, void *new_tbl, int64_t new_cap, cudaStream_t st) {
  int64_t blocks = (old_cap + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  closed_rehash_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const ulonglong2 *>(old_tbl), old_cap,
                                                        reinterpret_cast<unsigned long long *>(new_tbl), (uint64_t)new_cap - 1);
  return dcb_check_launch();
}

int closed_clear_device(void *tbl, int64_t cap, cudaStream_t st) {
  int64_t blocks = (cap + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  closed_clear_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<ulonglong2 *>(tbl), cap);
  return dcb_check_launch();
}

int closed_insert_device(int env, void *tbl, int64_t cap, const uint8_t *arena, const uint64_t *hash, const uint32_t *g,
                         const uint8_t *valid, uint32_t first_id, int64_t m, uint32_t *slot, uint8_t *keep,
                         uint32_t *num_entries, cudaStream_t st) {
  if (m == 0) return DCB_OK;
  const unsigned blocks = (unsigned)((m + 255) / 256);
  closed_insert_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<unsigned long long *>(tbl), (uint64_t)cap - 1, hash, g, valid,
                                               first_id, m, slot, num_entries);
  int rc = dcb_check_launch();
  if (rc) return rc;
  const unsigned long long *t = reinterpret_cast<const unsigned long long *>(tbl);
  switch (env) {
    case 0: closed_resolve_kernel<0><<<blocks, 256, 0, st>>>(t, arena, g, valid, first_id, m, slot, keep); break;
    case 1: closed_resolve_kernel<1><<<blocks, 256, 0, st>>>(t, arena, g, valid, first_id, m, slot, keep); break;
    case 2: closed_resolve_kernel<2><<<blocks, 256, 0, st>>>(t, arena, g, valid, first_id, m, slot, keep); break;
    case 3: closed_resolve_kernel<3><<<blocks, 256, 0, st>>>(t, arena, g, valid, first_id, m, slot, keep); break;
    case 4: closed_resolve_kernel<4><<<blocks, 256, 0, st>>>(t, arena, g, valid, first_id, m, slot, keep); break;
    case 5: closed_resolve_kernel<5><<<blocks, 256, 0, st>>>(t, arena, g, valid, first_id, m, slot, keep); break;
    case 6: closed_resolve_kernel<6><<<blocks, 256, 0, st>>>(t, arena, g, valid, first_id, m, slot, keep); break;
    default: return DCB_ERR_BAD_ENV;
  }
  return dcb_check_launch();
}

}  // namespace dcb
