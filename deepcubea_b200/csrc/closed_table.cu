// closed_table.cu -- CLOSED set as an open-addressing hash table in HBM.
//
// Replaces std::unordered_set<Node*,Hash,NodePointerEq> closed and the SERIAL find / insert /
// "(*found)->depth > node->depth -> overwrite and re-open" loop of cpp/parallel_weighted_astar.cpp:142,
// 243-265 (the reference's `check:` phase, ~12 ms per 24k children on 8 cores), and Instance.closed_dict /
// remove_in_closed of search_methods/astar.py:55, 78-90.
//
// Slot = 16 bytes {u64 key = state hash mixed with the instance number (0 = empty); u64 val = (g << 32) | node_id (all-ones =
// unset)}.  The reference's loop is sequential in child order: candidate i is kept iff its state is unseen or its g is
// STRICTLY smaller than what the table holds when the loop reaches it -- the value stored before the batch, lowered by every
// earlier candidate of the same state.  That rule is reproduced exactly by four launches over the whole batch:
//   1. probe  : linear probing; atomicCAS claims an empty key; the slot's value BEFORE the batch is recorded per candidate
//               (nobody writes values in this launch).
//   2. min    : candidates that beat the recorded value fold (g,id) into the slot with atomicMin -> the slot ends up with the
//               smallest (g,id) of the state, which is also what the sequential loop leaves behind.
//   3. resolve: not better than the old value -> dropped; own (g,id) won -> kept; lost to a winner with a SMALLER id -> dropped
//               (the winner came first and is at least as cheap); lost to a winner with a LARGER id -> the candidate came first
//               with a larger g: the sequential loop keeps it iff no even earlier candidate of the state is at least as cheap.
//               Those rare candidates go to a list and
//   4. fix-up : are settled against each other (id order, running minimum).
// A candidate is only ever dropped after comparing its state byte-for-byte (and its instance) with the node it lost to: a 64-bit
// hash collision keeps the candidate (a non-duplicate is never dropped).
// HBM traffic per candidate: 8 B hash + 4 B g in, one 32-B sector probe (~1.3 probes at load <= 0.5),
// 16 B slot update, 12 B scratch; + 2 x S bytes for the verify of a duplicate.
#include <cuda_runtime.h>
#include "closed_view.cuh"
#include "dcb_internal.h"
#include "ptx.cuh"
#include "state_ops.cuh"

namespace dcb {

__global__ void __launch_bounds__(256) closed_clear_kernel(ulonglong2 *tbl, int64_t cap) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < cap; i += (int64_t)gridDim.x * blockDim.x)
    tbl[i] = make_ulonglong2(0ull, ~0ull);
}

namespace {
constexpr uint32_t kNoSlot = 0xFFFFFFFFu;

__device__ __forceinline__ void warp_count(uint32_t *ctr, bool pred) {
  const unsigned ballot = __ballot_sync(0xffffffffu, pred);
  if (ctr && ballot && (threadIdx.x & 31) == 0) atomicAdd(ctr, __popc(ballot));
}
__device__ __forceinline__ uint32_t warp_append(uint32_t *ctr, bool pred) {
  const unsigned m = __ballot_sync(0xffffffffu, pred);
  if (!m) return 0;
  const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(ctr, (uint32_t)__popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  return base + __popc(m & ((1u << lane) - 1));
}

template <class View>
__global__ void __launch_bounds__(256)
closed_probe_kernel(View v, unsigned long long *__restrict__ tbl, uint64_t mask, ClosedScratch sc, uint32_t *num_entries, uint32_t *error) {
  const int64_t m = v.count();
  for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x; i0 < m; i0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = i0 + threadIdx.x;
    bool claimed = false;
    Cand c;
    if (i < m && v.get(i, c)) {
      uint32_t slot = kNoSlot;
      unsigned long long prev = ~0ull;
      uint64_t s = c.key & mask;
      for (uint64_t probes = 0; probes <= mask; probes++) {
        unsigned long long k = __ldcg(&tbl[2 * s]);          // cheap read first: most probes hit an occupied slot
        if (k == 0ull) k = atomicCAS(&tbl[2 * s], 0ull, (unsigned long long)c.key);
        if (k == 0ull || k == c.key) {
          claimed = (k == 0ull);
          prev = claimed ? ~0ull : __ldcg(&tbl[2 * s + 1]);  // values are not written in this launch: this is the pre-batch value
          slot = (uint32_t)s;
          break;
        }
        s = (s + 1) & mask;
      }
      if (slot == kNoSlot && error) atomicOr(error, 4u);     // table full: the candidate is kept, unrecorded
      sc.slot[i] = slot;
      sc.prev[i] = prev;
    }
    warp_count(num_entries, claimed);
  }
}

template <class View>
__global__ void __launch_bounds__(256) closed_min_kernel(View v, unsigned long long *__restrict__ tbl, ClosedScratch sc) {
  const int64_t m = v.count();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
    Cand c;
    if (!v.get(i, c)) continue;
    const uint32_t slot = sc.slot[i];
    if (slot == kNoSlot) continue;
    const uint32_t prev_g = (uint32_t)(sc.prev[i] >> 32);
    if (c.g < prev_g) atomicMin(&tbl[2 * (uint64_t)slot + 1], ((unsigned long long)c.g << 32) | c.id);
  }
}

// byte-for-byte compare of the states of two nodes of the arena
template <int ENV> __device__ __forceinline__ bool same_state(const uint8_t *__restrict__ arena, uint32_t ida, uint32_t idb) {
  constexpr int S = EnvTraits<ENV>::S, W = hash_words(S);
  uint32_t ra[LoadShape<S>::NRAW], rb[LoadShape<S>::NRAW], a[W], b[W];
  const uint64_t oa = (uint64_t)ida * S, ob = (uint64_t)idb * S;
  const uint32_t *pa = reinterpret_cast<const uint32_t *>(arena + (oa & ~uint64_t(3)));
  const uint32_t *pb = reinterpret_cast<const uint32_t *>(arena + (ob & ~uint64_t(3)));
#pragma unroll
  for (int q = 0; q < LoadShape<S>::NRAW; q++) { ra[q] = pa[q]; rb[q] = pb[q]; }
  align_state<S, W>(ra, (uint32_t)(oa & 3), a);
  align_state<S, W>(rb, (uint32_t)(ob & 3), b);
  uint32_t diff = 0;
#pragma unroll
  for (int q = 0; q < W; q++) diff |= a[q] ^ b[q];
  return diff == 0;
}

template <int ENV, class View>
__global__ void __launch_bounds__(256)
closed_resolve_kernel(View v, const unsigned long long *__restrict__ tbl, const uint8_t *__restrict__ arena, ClosedScratch sc,
                      uint8_t *__restrict__ keep, uint32_t *__restrict__ kept_ids, uint32_t *n_kept, uint32_t *n_amb) {
  const int64_t m = v.count();
  for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x; i0 < m; i0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = i0 + threadIdx.x;
    Cand c;
    const bool valid = i < m && v.get(i, c);
    uint8_t k = 0;          // 0 dropped, 1 kept, 2 undecided (fix-up)
    if (valid) {
      const uint32_t slot = sc.slot[i];
      if (slot == kNoSlot) k = 1;
      else {
        const unsigned long long prev = sc.prev[i];
        const uint32_t prev_g = (uint32_t)(prev >> 32);
        const unsigned long long myval = ((unsigned long long)c.g << 32) | c.id;
        if (c.g >= prev_g) {
          // not cheaper than the node CLOSED already holds: a duplicate, unless that node is another state (collision)
          const uint32_t oid = (uint32_t)prev;
          k = (v.same_inst(c.id, oid) && same_state<ENV>(arena, c.id, oid)) ? 0 : 1;
        } else {
          const unsigned long long win = __ldcg(&tbl[2 * (uint64_t)slot + 1]);
          if (win == myval) k = 1;
          else {
            const uint32_t wid = (uint32_t)win;
            if (!(v.same_inst(c.id, wid) && same_state<ENV>(arena, c.id, wid))) k = 1;      // collision: keep
            else k = (wid < c.id) ? 0 : 2;
          }
        }
      }
      if (keep) keep[i] = (k == 1) ? 1 : 0;
    }
    const uint32_t pos = warp_append(n_kept, k == 1);
    if (k == 1 && kept_ids) kept_ids[pos] = c.id;
    const uint32_t ap = warp_append(n_amb, k == 2);
    if (k == 2) sc.amb[ap] = make_uint4(sc.slot[i], c.id, c.g, (uint32_t)i);
  }
}

// Candidates that came BEFORE their state's in-batch winner with a larger g (but cheaper than the pre-batch value): the
// sequential loop keeps such a candidate iff no earlier candidate of the same state is at least as cheap.  The list is tiny
// (same state reached at two depths inside one batch, deeper one first), so all pairs are compared.
__global__ void __launch_bounds__(256)
closed_fixup_kernel(ClosedScratch sc, const uint32_t *n_amb_ptr, uint8_t *__restrict__ keep, uint32_t *__restrict__ kept_ids, uint32_t *n_kept) {
  const uint32_t n = *n_amb_ptr;
  for (uint32_t i0 = blockIdx.x * blockDim.x; i0 < n; i0 += gridDim.x * blockDim.x) {
    const uint32_t i = i0 + threadIdx.x;
    bool kept = false;
    uint4 me = make_uint4(0, 0, 0, 0);
    if (i < n) {
      me = sc.amb[i];
      kept = true;
      for (uint32_t j = 0; j < n; j++) {
        const uint4 o = sc.amb[j];
        if (o.x == me.x && o.y < me.y && o.z <= me.z) { kept = false; break; }
      }
      if (kept && keep) keep[me.w] = 1;
    }
    const uint32_t pos = warp_append(n_kept, kept);
    if (kept && kept_ids) kept_ids[pos] = me.y;
  }
}

template <int ENV, class View>
int launch_closed(View v, int64_t max_m, void *tbl, int64_t cap, const uint8_t *arena, ClosedScratch sc, uint8_t *keep,
                  uint32_t *kept_ids, uint32_t *n_kept, uint32_t *n_amb, uint32_t *num_entries, uint32_t *error, cudaStream_t st) {
  int64_t blocks = (max_m + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  unsigned long long *t = reinterpret_cast<unsigned long long *>(tbl);
  closed_probe_kernel<View><<<(unsigned)blocks, 256, 0, st>>>(v, t, (uint64_t)cap - 1, sc, num_entries, error);
  closed_min_kernel<View><<<(unsigned)blocks, 256, 0, st>>>(v, t, sc);
  closed_resolve_kernel<ENV, View><<<(unsigned)blocks, 256, 0, st>>>(v, t, arena, sc, keep, kept_ids, n_kept, n_amb);
  closed_fixup_kernel<<<64, 256, 0, st>>>(sc, n_amb, keep, kept_ids, n_kept);
  return dcb_check_launch();
}

template <class View>
int dispatch_closed(int env, View v, int64_t max_m, void *tbl, int64_t cap, const uint8_t *arena, ClosedScratch sc, uint8_t *keep,
                    uint32_t *kept_ids, uint32_t *n_kept, uint32_t *n_amb, uint32_t *num_entries, uint32_t *error, cudaStream_t st) {
  switch (env) {
    case 0: return launch_closed<0>(v, max_m, tbl, cap, arena, sc, keep, kept_ids, n_kept, n_amb, num_entries, error, st);
    case 1: return launch_closed<1>(v, max_m, tbl, cap, arena, sc, keep, kept_ids, n_kept, n_amb, num_entries, error, st);
    case 2: return launch_closed<2>(v, max_m, tbl, cap, arena, sc, keep, kept_ids, n_kept, n_amb, num_entries, error, st);
    case 3: return launch_closed<3>(v, max_m, tbl, cap, arena, sc, keep, kept_ids, n_kept, n_amb, num_entries, error, st);
    case 4: return launch_closed<4>(v, max_m, tbl, cap, arena, sc, keep, kept_ids, n_kept, n_amb, num_entries, error, st);
    case 5: return launch_closed<5>(v, max_m, tbl, cap, arena, sc, keep, kept_ids, n_kept, n_amb, num_entries, error, st);
    case 6: return launch_closed<6>(v, max_m, tbl, cap, arena, sc, keep, kept_ids, n_kept, n_amb, num_entries, error, st);
  }
  return DCB_ERR_BAD_ENV;
}

__global__ void __launch_bounds__(256)
closed_rehash_kernel(const ulonglong2 *__restrict__ old_tbl, int64_t old_cap, unsigned long long *__restrict__ new_tbl, uint64_t new_mask) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < old_cap; i += (int64_t)gridDim.x * blockDim.x) {
    const ulonglong2 e = old_tbl[i];
    if (e.x == 0ull) continue;
    uint64_t s = e.x & new_mask;
    for (uint64_t probes = 0; probes <= new_mask; probes++) {
      // keys are unique in the old table, so an empty slot is the only possible landing spot
      if (atomicCAS(&new_tbl[2 * s], 0ull, e.x) == 0ull) { new_tbl[2 * s + 1] = e.y; break; }
      s = (s + 1) & new_mask;
    }
  }
}
}  // namespace

// scratch carve-up for m candidates: slot u32[m] | prev u64[m] | amb uint4[m] | counters
int64_t closed_scratch_bytes(int64_t m) {
  if (m < 1) m = 1;
  return ((4 * m + 15) / 16 * 16) + 8 * m + 16 * m + 64;
}
ClosedScratch closed_carve(void *scratch, int64_t m) {
  if (m < 1) m = 1;
  uint8_t *p = reinterpret_cast<uint8_t *>(scratch);
  ClosedScratch sc;
  sc.slot = reinterpret_cast<uint32_t *>(p); p += (4 * m + 15) / 16 * 16;
  sc.prev = reinterpret_cast<unsigned long long *>(p); p += 8 * m;
  sc.amb = reinterpret_cast<uint4 *>(p); p += 16 * m;
  sc.counters = reinterpret_cast<uint32_t *>(p);
  return sc;
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

// stand-alone form: candidates i = 0..m-1 are nodes first_id + i (dcb_closed_insert)
int closed_insert_device(int env, void *tbl, int64_t cap, const uint8_t *arena, const uint64_t *hash, const uint32_t *g,
                         const uint8_t *valid, uint32_t first_id, int64_t m, void *scratch, uint8_t *keep,
                         uint32_t *num_entries, cudaStream_t st) {
  if (m == 0) return DCB_OK;
  ClosedScratch sc = closed_carve(scratch, m);
  if (cudaMemsetAsync(sc.counters, 0, 16, st) != cudaSuccess) return dcb_cuda_fail();
  if (cudaMemsetAsync(keep, 0, (size_t)m, st) != cudaSuccess) return dcb_cuda_fail();
  ContigView v{hash, g, valid, first_id, m};
  return dispatch_closed(env, v, m, tbl, cap, arena, sc, keep, nullptr, sc.counters, sc.counters + 1, num_entries, nullptr, st);
}

// search form: candidates come from the iteration's tile list (dcb_search_closed)
int closed_insert_tiles_device(int env, const TileView &v, int64_t max_m, void *tbl, int64_t cap, const uint8_t *arena, void *scratch,
                               uint32_t *kept_ids, dcb_step_plan *plan, cudaStream_t st) {
  ClosedScratch sc = closed_carve(scratch, max_m);
  return dispatch_closed(env, v, max_m, tbl, cap, arena, sc, nullptr, kept_ids, &plan->n_kept, &plan->n_ambiguous, &plan->closed_entries,
                         &plan->error, st);
}

}  // namespace dcb
