// node_ops.cu -- per-node bookkeeping of the BWAS loop: child metadata, kept-list compaction, NN input
// gather, cost, path reconstruction.  Nodes are identified by id = slot * A + move; the state of node id
// sits at arena + id * S, so a node's move and parent slot are implicit in its id.
#include <cuda_runtime.h>
#include "dcb_internal.h"
#include "state_ops.cuh"

namespace dcb {

// Node{depth, parent} of parallel_weighted_astar.cpp:80-86, 221-226 as SoA arrays.
__global__ void __launch_bounds__(256)
child_meta_kernel(const uint32_t *__restrict__ parent_ids, int64_t n_parents, int A, uint32_t first_id,
                  uint32_t *__restrict__ node_g, uint32_t *__restrict__ slot_parent) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n_parents * A) return;
  const uint32_t pid = parent_ids[i / A];
  node_g[first_id + i] = node_g[pid] + 1;                      // depth = parent depth + 1 (:219)
  if (i % A == 0) slot_parent[(first_id + i) / A] = pid;
}

__global__ void __launch_bounds__(256)
compact_kept_kernel(const uint8_t *__restrict__ keep, uint32_t first_id, int64_t m, uint32_t *__restrict__ out_ids, uint32_t *counter) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const bool want = i < m && keep[i];
  const unsigned mask = __ballot_sync(0xffffffffu, want);
  if (!mask) return;
  const int lane = threadIdx.x & 31, leader = __ffs(mask) - 1;
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(mask));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (want) out_ids[base + __popc(mask & ((1u << lane) - 1))] = first_id + (uint32_t)i;
}

// state_to_nnet_input on gathered nodes (cube3.py:77-85 colour = sticker/9; cube4 sticker/16; n_puzzle.py:84-89 identity)
template <int DIV>
__global__ void __launch_bounds__(256)
gather_nnet_kernel(const uint8_t *__restrict__ arena, const uint32_t *__restrict__ ids, int64_t m, int S, uint8_t *__restrict__ out) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= m * S) return;
  const int64_t j = t / S;
  const int b = (int)(t - j * S);
  const uint8_t v = arena[(uint64_t)ids[j] * S + b];
  out[t] = DIV == 9 ? (uint8_t)((v * 57u) >> 9) : (DIV == 16 ? (uint8_t)(v >> 4) : v);
}

// cost = h * (!solved) + weight * depth in float32, no FMA contraction (parallel_weighted_astar.cpp:298;
// g++ -O3 without -march emits separate mulss/addss).  h is clipped at 0 first (nnet_utils.py:193-194).
__global__ void __launch_bounds__(256)
cost_kernel(const float *__restrict__ h, const uint32_t *__restrict__ ids, const uint32_t *__restrict__ node_g,
            const uint8_t *__restrict__ node_solved, float weight, int64_t m, float *__restrict__ cost) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= m) return;
  const uint32_t id = ids[j];
  const float hv = fmaxf(h[j], 0.0f);
  const float ns = node_solved[id] ? 0.0f : 1.0f;
  cost[j] = __fadd_rn(__fmul_rn(hv, ns), __fmul_rn(weight, (float)node_g[id]));
}

// parallel_weighted_astar.cpp:336-341 / astar.py:213-229
__global__ void path_kernel(const uint32_t *__restrict__ slot_parent, uint32_t goal_id, int A, int32_t max_len,
                            uint8_t *moves, int32_t *len_out) {
  if (threadIdx.x || blockIdx.x) return;
  int32_t len = 0;
  uint32_t id = goal_id;
  while (id != 0) {
    if (len >= max_len) { *len_out = -1; return; }
    moves[len++] = (uint8_t)(id % A);
    id = slot_parent[id / A];
  }
  for (int32_t i = 0; i < len / 2; i++) { const uint8_t t = moves[i]; moves[i] = moves[len - 1 - i]; moves[len - 1 - i] = t; }
  *len_out = len;
}

int child_meta_device(const uint32_t *parent_ids, int64_t n_parents, int A, uint32_t first_id, uint32_t *node_g,
                      uint32_t *slot_parent, cudaStream_t st) {
  const int64_t m = n_parents * A;
  if (m == 0) return DCB_OK;
  child_meta_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(parent_ids, n_parents, A, first_id, node_g, slot_parent);
  return dcb_check_launch();
}
int compact_kept_device(const uint8_t *keep, uint32_t first_id, int64_t m, uint32_t *out_ids, uint32_t *counter, cudaStream_t st) {
  if (m == 0) return DCB_OK;
  compact_kept_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(keep, first_id, m, out_ids, counter);
  return dcb_check_launch();
}
int gather_nnet_device(int env, const uint8_t *arena, const uint32_t *ids, int64_t m, uint8_t *out, cudaStream_t st) {
  if (env < 0 || env >= DCB_NUM_ENVS) return DCB_ERR_BAD_ENV;
  if (m == 0) return DCB_OK;
  const int S = dcb_env_state_bytes(env);
  const unsigned blocks = (unsigned)((m * S + 255) / 256);
  if (env == 0) gather_nnet_kernel<9><<<blocks, 256, 0, st>>>(arena, ids, m, S, out);
  else if (env == DCB_ENV_CUBE4) gather_nnet_kernel<16><<<blocks, 256, 0, st>>>(arena, ids, m, S, out);
  else gather_nnet_kernel<0><<<blocks, 256, 0, st>>>(arena, ids, m, S, out);
  return dcb_check_launch();
}
int cost_device(const float *h, const uint32_t *ids, const uint32_t *node_g, const uint8_t *node_solved, float weight,
                int64_t m, float *cost, cudaStream_t st) {
  if (m == 0) return DCB_OK;
  cost_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(h, ids, node_g, node_solved, weight, m, cost);
  return dcb_check_launch();
}
int path_device(const uint32_t *slot_parent, uint32_t goal_id, int A, int32_t max_len, uint8_t *moves, int32_t *len, cudaStream_t st) {
  path_kernel<<<1, 32, 0, st>>>(slot_parent, goal_id, A, max_len, moves, len);
  return dcb_check_launch();
}

}  // namespace dcb
