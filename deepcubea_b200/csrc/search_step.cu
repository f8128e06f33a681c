// search_step.cu -- the device-driven search iteration (include/dcb.h, dcb_search_*): roots, slot assignment + tile list after
// the pop, cost + push per instance, path reconstruction in a shared arena.  One BWAS iteration
// (cpp/parallel_weighted_astar.cpp:169-330) or one AStar.step over all instances (search_methods/astar.py:256-317) is
//     dcb_search_pop -> dcb_search_expand -> dcb_search_closed -> cost-to-go network -> dcb_search_push
// and every size in between (parents popped, tiles, children kept) lives in device memory: the host enqueues the whole
// iteration without reading anything back.
#include <cuda_runtime.h>
#include "closed_view.cuh"
#include "dcb_internal.h"
#include "state_ops.cuh"

namespace dcb {
int open_pop_device(void *state, uint32_t *key, uint32_t *key_lo, uint32_t *id, int64_t seg_cap, int n_inst, int32_t batch, int mode, int stop_at_goal,
                    int include_solved, int num_moves, const uint8_t *node_solved, const uint32_t *node_g, uint32_t *popped_ids,
                    int64_t popped_stride, void *scratch, const dcb_step_plan *plan, cudaStream_t st);
int64_t open_scratch_bytes(int64_t capacity, int64_t batch, int64_t n_inst);
int closed_insert_tiles_device(int env, const TileView &v, int64_t max_m, void *tbl, int64_t cap, const uint8_t *arena, void *scratch,
                               uint32_t *kept_ids, dcb_step_plan *plan, cudaStream_t st);

namespace {
constexpr uint32_t kNone = 0xFFFFFFFFu;

// ---- roots ------------------------------------------------------------------------------------------
// One thread per instance (parallel_weighted_astar.cpp:160-166; astar.py:244-249, 50-62).
template <int ENV>
__global__ void __launch_bounds__(128)
search_reset_kernel(const uint8_t *__restrict__ roots, int n_inst, int semantics, uint32_t slots_per_inst, uint32_t open_per_inst,
                    uint8_t *__restrict__ arena, uint32_t *__restrict__ node_g, uint8_t *__restrict__ node_solved,
                    uint32_t *__restrict__ slot_parent, unsigned long long *__restrict__ tbl, uint64_t mask, uint32_t *__restrict__ open_key,
                    uint32_t *__restrict__ open_id, dcb_search_inst *__restrict__ inst, dcb_step_plan *__restrict__ plan,
                    uint32_t *__restrict__ kept_ids) {
  constexpr int S = EnvTraits<ENV>::S, A = EnvTraits<ENV>::A, W = hash_words(S);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) {
    dcb_step_plan z = {};
    z.budget = plan->budget;                              // the host's iteration budget outlives a reset
    z.n_running = (uint32_t)n_inst;
    z.n_kept = semantics == 1 ? (uint32_t)n_inst : 0u;
    z.closed_entries = semantics == 0 ? (uint32_t)n_inst : 0u;
    *plan = z;
  }
  if (i >= n_inst) return;
  const uint32_t root_slot = (uint32_t)i * slots_per_inst, root_id = root_slot * A;
  uint32_t w[W];
#pragma unroll
  for (int k = 0; k < W; k++) w[k] = 0;
  const uint8_t *src = roots + (size_t)i * S;
  uint8_t *dst = arena + (size_t)root_id * S;
#pragma unroll
  for (int j = 0; j < S; j++) {
    const uint8_t b = src[j];
    dst[j] = b;
    w[j / 4] |= (uint32_t)b << (8 * (j % 4));
  }
  const bool solved = is_goal<ENV, W>(w);
  node_g[root_id] = 0;
  node_solved[root_id] = solved ? 1 : 0;
  slot_parent[root_slot] = kNone;
  dcb_search_inst z = {};
  z.goal_id = kNone;
  z.goal_key = kNone;
  z.next_slot = 1;
  if (semantics == 0) {
    // root in CLOSED with depth 0 and in OPEN with cost 0 / heuristic 0; it counts as generated
    const uint64_t k0 = state_hash<W>(w) ^ inst_mix((uint32_t)i);
    const unsigned long long key = k0 ? k0 : 1ull;
    uint64_t s = key & mask;
    for (uint64_t probes = 0; probes <= mask; probes++) {
      const unsigned long long k = atomicCAS(&tbl[2 * s], 0ull, key);
      if (k == 0ull || k == key) { atomicMin(&tbl[2 * s + 1], (unsigned long long)root_id); break; }
      s = (s + 1) & mask;
    }
    open_key[(size_t)i * open_per_inst] = 0u;            // float bits of cost 0.0
    open_id[(size_t)i * open_per_inst] = root_id;
    z.open_size = 1;
    z.nodes_generated = 1;
  } else {
    kept_ids[i] = root_id;                                // the caller evaluates the roots, then dcb_search_push
  }
  inst[i] = z;
}

// ---- after the pop: slots, tiles, plan ----------------------------------------------------------------
// Single block.  Every instance's expanded parents get consecutive arena slots starting at a multiple of `align` (so that the
// child block is 16-byte aligned for the TMA store) inside the instance's slot range; the iteration's tile list is laid out
// instance after instance.
__global__ void __launch_bounds__(1024)
search_plan_kernel(dcb_search_inst *__restrict__ inst, int n_inst, uint32_t slots_per_inst, uint32_t align, uint32_t popped_stride, int num_moves,
                   uint32_t batch, uint4 *__restrict__ tiles, dcb_step_plan *__restrict__ plan) {
  __shared__ uint32_t part[1024];
  __shared__ uint32_t carry, tot_parents, running, err;
  const int t = threadIdx.x;
  if (t == 0) { carry = 0; tot_parents = 0; running = 0; err = 0; }
  __syncthreads();
  for (int base = 0; base < n_inst; base += 1024) {
    const int i = base + t;
    uint32_t nt = 0, ne = 0;
    if (i < n_inst) {
      dcb_search_inst *s = inst + i;
      ne = s->n_expand;
      if (ne) {
        const uint32_t b0 = (s->next_slot + align - 1) / align * align;
        if ((uint64_t)b0 + ne > slots_per_inst) {          // node arena of this instance is full: stop it
          s->done = 3; s->n_expand = 0; ne = 0;
          atomicOr(&err, 1u);
        } else {
          s->base_slot = b0;
          s->next_slot = b0 + ne;
          s->nodes_expanded += (uint64_t)ne * (uint64_t)num_moves;
        }
      }
      if (s->overflow && s->done == 0) { s->done = 4; s->n_expand = 0; ne = 0; atomicOr(&err, 2u); }
      nt = (ne + 31) / 32;
      if (s->done == 0) atomicAdd(&running, 1u);
      atomicAdd(&tot_parents, ne);
    }
    part[t] = nt;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
      const uint32_t v = (t >= off) ? part[t - off] : 0;
      __syncthreads();
      part[t] += v;
      __syncthreads();
    }
    const uint32_t off0 = carry + part[t] - nt;
    if (i < n_inst) inst[i].tile_off = off0;
    __syncthreads();
    if (t == 1023) carry += part[1023];
    __syncthreads();
  }
  // tile descriptors: few instances -> threads stride over an instance's tiles; many -> one thread per instance
  if (n_inst <= 64) {
    for (int i = 0; i < n_inst; i++) {
      const dcb_search_inst *s = inst + i;
      const uint32_t ne = s->n_expand, nt = (ne + 31) / 32, off0 = s->tile_off;
      const uint32_t gslot = (uint32_t)i * slots_per_inst + s->base_slot;
      for (uint32_t j = t; j < nt; j += 1024)
        tiles[off0 + j] = make_uint4((uint32_t)i * popped_stride + 32 * j, gslot + 32 * j, min(32u, ne - 32 * j), (uint32_t)i);
    }
  } else {
    for (int i = t; i < n_inst; i += 1024) {
      const dcb_search_inst *s = inst + i;
      const uint32_t ne = s->n_expand, nt = (ne + 31) / 32, off0 = s->tile_off;
      const uint32_t gslot = (uint32_t)i * slots_per_inst + s->base_slot;
      for (uint32_t j = 0; j < nt; j++)
        tiles[off0 + j] = make_uint4((uint32_t)i * popped_stride + 32 * j, gslot + 32 * j, min(32u, ne - 32 * j), (uint32_t)i);
    }
  }
  __syncthreads();
  if (t == 0) {
    plan->n_tiles = carry;
    plan->n_parents = tot_parents;
    plan->n_kept = 0;
    plan->n_ambiguous = 0;
    plan->n_running = running;
    plan->error |= err;
    plan->total_expanded += (uint64_t)tot_parents * (uint64_t)num_moves;
    const uint32_t bud = plan->budget;
    if (bud != 0xFFFFFFFFu && bud > 0 && tot_parents == (uint32_t)n_inst * batch) plan->budget = bud - 1;     // a full-batch iteration
  }
}

// ---- cost + push -----------------------------------------------------------------------------------------
// C++ semantics: cost = h * (!solved) + weight * depth in float32, no FMA contraction (parallel_weighted_astar.cpp:298; g++ -O3
// without -march emits separate mulss/addss).  Python semantics: cost = weight * path_cost + h * (!solved) in float64 with h the
// float32 network output widened (astar.py:196; nnet_utils.py:172-194), kept as a 64-bit order-preserving key in two words.  h is
// clipped at 0 first (nnet_utils.py:193-194).  The node goes to the OPEN segment of the instance that owns it (astar.py:206-209).
__global__ void __launch_bounds__(256)
search_push_kernel(const uint32_t *__restrict__ kept_ids, dcb_step_plan *__restrict__ plan, const float *__restrict__ h,
                   const float *__restrict__ dot_partial, int n_parts, float dot_bias, const uint32_t *__restrict__ node_g,
                   const uint8_t *__restrict__ node_solved, const double *__restrict__ weights, uint32_t nodes_per_inst, int n_inst,
                   uint32_t open_per_inst, dcb_search_inst *__restrict__ inst, uint32_t *__restrict__ open_key,
                   uint32_t *__restrict__ open_key_lo, uint32_t *__restrict__ open_id) {
  const uint32_t n = plan->n_kept;
  if (blockIdx.x == 0 && threadIdx.x == 0) plan->total_kept += n;
  for (uint32_t j0 = blockIdx.x * blockDim.x; j0 < n; j0 += gridDim.x * blockDim.x) {
    const uint32_t j = j0 + threadIdx.x;
    if (j >= n) continue;                                  // (no warp-wide primitive below needs the whole warp)
    const uint32_t id = kept_ids[j];
    const uint32_t ii = n_inst > 1 ? id / nodes_per_inst : 0u;
    float hv;
    if (h) hv = h[j];
    else {
      hv = 0.0f;
      for (int q = 0; q < n_parts; q++) hv += dot_partial[(size_t)j * n_parts + q];     // fixed order: reproducible
      hv += dot_bias;
    }
    hv = fmaxf(hv, 0.0f);
    const float ns = node_solved[id] ? 0.0f : 1.0f;
    uint32_t k_hi, k_lo = 0;
    if (open_key_lo) {
      const double c64 = __dadd_rn(__dmul_rn(weights[ii], (double)node_g[id]), __dmul_rn((double)hv, (double)ns));
      const unsigned long long bits = (unsigned long long)__double_as_longlong(c64);      // costs are >= 0: bit order == value order
      k_hi = (uint32_t)(bits >> 32); k_lo = (uint32_t)bits;
    } else {
      k_hi = __float_as_uint(__fadd_rn(__fmul_rn(hv, ns), __fmul_rn((float)weights[ii], (float)node_g[id])));
    }
    // one atomic per (warp, instance)
    const unsigned peers = __match_any_sync(__activemask(), ii);
    const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(&inst[ii].open_size, (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    const uint32_t pos = base + __popc(peers & ((1u << lane) - 1));
    if (pos < open_per_inst) {
      open_key[(size_t)ii * open_per_inst + pos] = k_hi;
      if (open_key_lo) open_key_lo[(size_t)ii * open_per_inst + pos] = k_lo;
      open_id[(size_t)ii * open_per_inst + pos] = id;
    } else {
      inst[ii].overflow = 1;
    }
  }
}

// parallel_weighted_astar.cpp:336-341 / astar.py:213-229; roots are the nodes whose id is a multiple of nodes_per_inst
__global__ void search_path_kernel(const uint32_t *__restrict__ slot_parent, uint32_t goal_id, int A, uint32_t nodes_per_inst, int32_t max_len,
                                   uint8_t *moves, int32_t *len_out) {
  if (threadIdx.x || blockIdx.x) return;
  int32_t len = 0;
  uint32_t id = goal_id;
  while (id % nodes_per_inst != 0) {
    if (len >= max_len) { *len_out = -1; return; }
    moves[len++] = (uint8_t)(id % A);
    id = slot_parent[id / A];
  }
  for (int32_t i = 0; i < len / 2; i++) { const uint8_t t = moves[i]; moves[i] = moves[len - 1 - i]; moves[len - 1 - i] = t; }
  *len_out = len;
}

inline int env_moves(int env) { return dcb_env_num_moves(env); }
inline uint32_t ceil32(int32_t b) { return (uint32_t)((b + 31) / 32 * 32); }
}  // namespace

int search_reset_device(const dcb_search_ctx &c, const uint8_t *roots, cudaStream_t st) {
  const unsigned blocks = (unsigned)((c.n_inst + 127) / 128);
  unsigned long long *tbl = reinterpret_cast<unsigned long long *>(c.d_closed);
  const uint64_t mask = (uint64_t)c.closed_capacity - 1;
#define DCB_RESET(E)                                                                                                                   \
  search_reset_kernel<E><<<blocks, 128, 0, st>>>(roots, c.n_inst, c.semantics, c.slots_per_inst, c.open_per_inst, c.d_arena, c.d_node_g, \
                                                 c.d_node_solved, c.d_slot_parent, tbl, mask, c.d_open_key, c.d_open_id, c.d_inst, c.d_plan, \
                                                 c.d_kept_ids)
  switch (c.env) {
    case 0: DCB_RESET(0); break;
    case 1: DCB_RESET(1); break;
    case 2: DCB_RESET(2); break;
    case 3: DCB_RESET(3); break;
    case 4: DCB_RESET(4); break;
    case 5: DCB_RESET(5); break;
    case 6: DCB_RESET(6); break;
    default: return DCB_ERR_BAD_ENV;
  }
#undef DCB_RESET
  return dcb_check_launch();
}

int search_pop_device(const dcb_search_ctx &c, int include_solved, cudaStream_t st) {
  const int A = env_moves(c.env);
  const int rc = open_pop_device(c.d_inst, c.d_open_key, c.d_open_key_lo, c.d_open_id, c.open_per_inst, c.n_inst, c.batch, c.semantics, c.semantics == 0 ? 1 : 0,
                                 include_solved, A, c.d_node_solved, c.d_node_g, c.d_popped_ids, ceil32(c.batch), c.d_pop_scratch, c.d_plan, st);
  if (rc) return rc;
  search_plan_kernel<<<1, 1024, 0, st>>>(c.d_inst, c.n_inst, c.slots_per_inst, (uint32_t)dcb_env_slot_align(c.env), ceil32(c.batch), A,
                                         (uint32_t)c.batch, reinterpret_cast<uint4 *>(c.d_tiles), c.d_plan);
  return dcb_check_launch();
}

int search_expand_device(const dcb_search_ctx &c, cudaStream_t st) {
  const int64_t max_tiles = (int64_t)c.n_inst * (ceil32(c.batch) / 32);
  return expand_planned_device(c.env, c.d_arena, c.d_popped_ids, max_tiles, c.d_tiles, c.d_plan, c.d_node_solved, c.d_hash, c.d_node_g,
                               c.d_slot_parent, st);
}

int search_closed_device(const dcb_search_ctx &c, cudaStream_t st) {
  const int A = env_moves(c.env);
  const int64_t max_m = (int64_t)c.n_inst * ceil32(c.batch) * A;
  TileView v{reinterpret_cast<const uint4 *>(c.d_tiles), c.d_plan, c.d_hash, c.d_node_g, (uint32_t)A, c.slots_per_inst * (uint32_t)A};
  return closed_insert_tiles_device(c.env, v, max_m, c.d_closed, c.closed_capacity, c.d_arena, c.d_closed_scratch, c.d_kept_ids, c.d_plan, st);
}

int search_push_device(const dcb_search_ctx &c, const float *h, const float *dot_partial, int n_parts, float dot_bias, cudaStream_t st) {
  const int A = env_moves(c.env);
  int64_t max_m = (int64_t)c.n_inst * ceil32(c.batch) * A;
  int64_t blocks = (max_m + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  search_push_kernel<<<(unsigned)blocks, 256, 0, st>>>(c.d_kept_ids, c.d_plan, h, dot_partial, n_parts, dot_bias, c.d_node_g, c.d_node_solved,
                                                      c.d_weights, c.slots_per_inst * (uint32_t)A, c.n_inst, c.open_per_inst, c.d_inst,
                                                      c.d_open_key, c.d_open_key_lo, c.d_open_id);
  return dcb_check_launch();
}

int search_path_device(const dcb_search_ctx &c, uint32_t node_id, int32_t max_len, uint8_t *moves, int32_t *len, cudaStream_t st) {
  const int A = env_moves(c.env);
  search_path_kernel<<<1, 32, 0, st>>>(c.d_slot_parent, node_id, A, c.slots_per_inst * (uint32_t)A, max_len, moves, len);
  return dcb_check_launch();
}

int64_t search_pop_scratch_bytes(int32_t n_inst, int64_t open_per_inst, int32_t batch) { return open_scratch_bytes(open_per_inst, batch, n_inst); }

}  // namespace dcb
