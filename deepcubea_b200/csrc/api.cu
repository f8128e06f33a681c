// api.cu -- the extern "C" surface declared in include/dcb.h: argument validation, error reporting,
// host-buffer convenience entry points.  No kernel lives here.
#include <cuda_runtime.h>
#include <string.h>
#include "cube3_moves.cuh"
#include "cube4_moves.cuh"
#include "dcb_internal.h"

namespace dcb {
static thread_local cudaError_t t_last_cuda = cudaSuccess;
int dcb_record_cuda(cudaError_t e) {
  if (e == cudaSuccess) return DCB_OK;
  t_last_cuda = e;
  return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? DCB_ERR_NO_DEVICE : DCB_ERR_CUDA;
}
int dcb_cuda_fail() { return dcb_record_cuda(cudaGetLastError()); }
int dcb_check_launch() { return dcb_record_cuda(cudaGetLastError()); }

int closed_clear_device(void *tbl, int64_t cap, cudaStream_t st);
int closed_insert_device(int env, void *tbl, int64_t cap, const uint8_t *arena, const uint64_t *hash, const uint32_t *g,
                         const uint8_t *valid, uint32_t first_id, int64_t m, void *scratch, uint8_t *keep,
                         uint32_t *num_entries, cudaStream_t st);
int64_t closed_scratch_bytes(int64_t m);
int closed_rehash_device(const void *old_tbl, int64_t old_cap, void *new_tbl, int64_t new_cap, cudaStream_t st);
int open_clear_device(void *state, int n_inst, cudaStream_t st);
int open_push_device(void *state, uint32_t *key, uint32_t *key_lo, uint32_t *id, int64_t capacity, const float *cost, const uint32_t *ids,
                     uint32_t first_id, const uint8_t *keep, int64_t m, cudaStream_t st);
int64_t open_scratch_bytes(int64_t capacity, int64_t batch, int64_t n_inst);
int open_pop_device(void *state, uint32_t *key, uint32_t *key_lo, uint32_t *id, int64_t seg_cap, int n_inst, int32_t batch, int mode, int stop_at_goal,
                    int include_solved, int num_moves, const uint8_t *node_solved, const uint32_t *node_g, uint32_t *popped_ids,
                    int64_t popped_stride, void *scratch, const dcb_step_plan *plan, cudaStream_t st);
int search_reset_device(const dcb_search_ctx &c, const uint8_t *roots, cudaStream_t st);
int search_pop_device(const dcb_search_ctx &c, int include_solved, cudaStream_t st);
int search_expand_device(const dcb_search_ctx &c, cudaStream_t st);
int search_closed_device(const dcb_search_ctx &c, cudaStream_t st);
int search_push_device(const dcb_search_ctx &c, const float *h, const float *dot_partial, int n_parts, float dot_bias, cudaStream_t st);
int search_path_device(const dcb_search_ctx &c, uint32_t node_id, int32_t max_len, uint8_t *moves, int32_t *len, cudaStream_t st);
int64_t search_pop_scratch_bytes(int32_t n_inst, int64_t open_per_inst, int32_t batch);
int child_meta_device(const uint32_t *parent_ids, int64_t n_parents, int A, uint32_t first_id, uint32_t *node_g,
                      uint32_t *slot_parent, cudaStream_t st);
int compact_kept_device(const uint8_t *keep, uint32_t first_id, int64_t m, uint32_t *out_ids, uint32_t *counter, cudaStream_t st);
int gather_nnet_device(int env, const uint8_t *arena, const uint32_t *ids, int64_t m, uint8_t *out, cudaStream_t st);
int cost_device(const float *h, const uint32_t *ids, const uint32_t *node_g, const uint8_t *node_solved, float weight,
                int64_t m, float *cost, cudaStream_t st);
int resnet_gemm_device(const void *a_hi, const void *a_lo, int64_t lda, const void *w_hi, const void *w_lo, int64_t ldw, const float *bias,
                       float scale, const void *skip_hi, const void *skip_lo, int relu, void *out_hi, void *out_lo, float *out_f32,
                       const float *partial_in, float *partial_out, const float *dot_w, float *dot_partial, int64_t M, int Np, int Kp,
                       const int32_t *m_dev, int32_t m_off, int chunk_k, void *scratch, cudaStream_t st);
int64_t resnet_gemm_scratch_bytes();
int onehot_device(const uint8_t *x, int64_t M, int S, int depth, int Kp, void *out, cudaStream_t st);
int onehot_gather_device(int env, const uint8_t *arena, const uint32_t *ids, int64_t M, int S, int depth, int Kp, void *out, const int32_t *m_dev,
                         int32_t m_off, cudaStream_t st);
int rowdot_device(const void *x_hi, const void *x_lo, const float *w, float bias, int64_t M, int n_valid, int ld, float *out, cudaStream_t st);
int path_device(const uint32_t *slot_parent, uint32_t goal_id, int A, int32_t max_len, uint8_t *moves, int32_t *len, cudaStream_t st);
}  // namespace dcb

using namespace dcb;

namespace {
const int kStateBytes[DCB_NUM_ENVS] = {54, 16, 25, 36, 49, 49, 96};
const int kNumMoves[DCB_NUM_ENVS] = {12, 4, 4, 4, 4, 49, 24};
const int kDim[DCB_NUM_ENVS] = {3, 4, 5, 6, 7, 7, 4};
inline bool env_ok(int env) { return env >= 0 && env < DCB_NUM_ENVS; }
inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline cudaStream_t S(void *s) { return reinterpret_cast<cudaStream_t>(s); }
inline bool pow2(int64_t v) { return v > 0 && (v & (v - 1)) == 0; }
}  // namespace

extern "C" {

int dcb_abi_version(void) { return DCB_ABI_VERSION; }

const char *dcb_error_string(int code) {
  switch (code) {
    case DCB_OK: return "ok";
    case DCB_ERR_BAD_ENV: return "unknown environment id";
    case DCB_ERR_BAD_ARG: return "bad argument";
    case DCB_ERR_ALIGN: return "pointer not 16-byte aligned";
    case DCB_ERR_CUDA: return "CUDA runtime error";
    case DCB_ERR_NO_DEVICE: return "no usable CUDA device";
    case DCB_ERR_CAPACITY: return "capacity exceeded";
  }
  return "unknown error code";
}
const char *dcb_last_cuda_error(void) { return t_last_cuda == cudaSuccess ? "" : cudaGetErrorString(t_last_cuda); }

int dcb_env_num_moves(int env) { return env_ok(env) ? kNumMoves[env] : DCB_ERR_BAD_ENV; }
int dcb_env_state_bytes(int env) { return env_ok(env) ? kStateBytes[env] : DCB_ERR_BAD_ENV; }
int dcb_env_slot_align(int env) {
  if (!env_ok(env)) return DCB_ERR_BAD_ENV;
  const int rec = kStateBytes[env] * kNumMoves[env];
  int a = 1;
  while ((rec * a) % 16) a *= 2;
  return a;
}
int dcb_env_goal_state(int env, uint8_t *h_out) {
  if (!env_ok(env)) return DCB_ERR_BAD_ENV;
  if (!h_out) return DCB_ERR_BAD_ARG;
  const int s = kStateBytes[env];
  for (int j = 0; j < s; j++) h_out[j] = (env == 0 || env == DCB_ENV_CUBE4) ? (uint8_t)j : (env == DCB_ENV_LIGHTSOUT7 ? (uint8_t)0 : (uint8_t)((j + 1) % s));
  return DCB_OK;
}
int dcb_env_move_table(int env, int32_t *h_out, int64_t capacity_elems) {
  if (!env_ok(env)) return DCB_ERR_BAD_ENV;
  if (!h_out) return DCB_ERR_BAD_ARG;
  if (env == 0) {
    if (capacity_elems < 12 * 54) return DCB_ERR_BAD_ARG;
    for (int a = 0; a < 12; a++)
      for (int j = 0; j < 54; j++) h_out[a * 54 + j] = kCube3PermHost[a][j];
    return DCB_OK;
  }
  if (env == DCB_ENV_CUBE4) {                   // perm[24][96], child[j] = parent[perm[a][j]] (cpp/environments.cpp:262-341 in gather form)
    if (capacity_elems < 24 * 96) return DCB_ERR_BAD_ARG;
    for (int a = 0; a < 24; a++)
      for (int j = 0; j < 96; j++) h_out[a * 96 + j] = kCube4PermHost[a][j];
    return DCB_OK;
  }
  const int d = kDim[env];
  if (env == DCB_ENV_LIGHTSOUT7) {              // move_matrix[49][5] of lights_out.py:31-42 / getMoveMat (environments.cpp:133-155)
    if (capacity_elems < (int64_t)d * d * 5) return DCB_ERR_BAD_ARG;
    for (int m = 0; m < d * d; m++) {
      const int x = m / d, y = m % d;
      h_out[m * 5 + 0] = m;
      h_out[m * 5 + 1] = x < d - 1 ? m + d : m;
      h_out[m * 5 + 2] = x > 0 ? m - d : m;
      h_out[m * 5 + 3] = y < d - 1 ? m + 1 : m;
      h_out[m * 5 + 4] = y > 0 ? m - 1 : m;
    }
    return DCB_OK;
  }
  if (capacity_elems < (int64_t)d * d * 4) return DCB_ERR_BAD_ARG;
  for (int i = 0; i < d; i++)
    for (int j = 0; j < d; j++) {
      const int z = i * d + j;
      h_out[z * 4 + 0] = i < d - 1 ? z + d : z;   // U: blank swaps with the tile below
      h_out[z * 4 + 1] = i > 0 ? z - d : z;       // D
      h_out[z * 4 + 2] = j < d - 1 ? z + 1 : z;   // L
      h_out[z * 4 + 3] = j > 0 ? z - 1 : z;       // R
    }
  return DCB_OK;
}

int dcb_expand(int env, const uint8_t *d_parents, int64_t n, uint8_t *d_children, uint8_t *d_solved, uint64_t *d_hash,
               void *stream) {
  if (!env_ok(env)) return DCB_ERR_BAD_ENV;
  if (n < 0 || (n > 0 && (!d_parents || !d_children))) return DCB_ERR_BAD_ARG;
  if (!aligned16(d_children) || !aligned16(d_hash) || (reinterpret_cast<uintptr_t>(d_parents) & 3u) ||
      (reinterpret_cast<uintptr_t>(d_solved) & 3u))
    return DCB_ERR_ALIGN;
  return expand_device(env, d_parents, nullptr, n, d_children, d_solved, d_hash, S(stream));
}
int dcb_expand_indexed(int env, const uint8_t *d_arena, const uint32_t *d_parent_ids, int64_t n, uint8_t *d_children,
                       uint8_t *d_solved, uint64_t *d_hash, void *stream) {
  if (!env_ok(env)) return DCB_ERR_BAD_ENV;
  if (n < 0 || (n > 0 && (!d_arena || !d_children || !d_parent_ids))) return DCB_ERR_BAD_ARG;
  if (!aligned16(d_children) || !aligned16(d_hash) || (reinterpret_cast<uintptr_t>(d_arena) & 3u) ||
      (reinterpret_cast<uintptr_t>(d_solved) & 3u))
    return DCB_ERR_ALIGN;
  return expand_device(env, d_arena, d_parent_ids, n, d_children, d_solved, d_hash, S(stream));
}
int dcb_next_state(int env, const uint8_t *d_states, int64_t n, int action, uint8_t *d_next, void *stream) {
  if (!env_ok(env)) return DCB_ERR_BAD_ENV;
  if (n < 0 || action < 0 || action >= kNumMoves[env] || (n > 0 && (!d_states || !d_next))) return DCB_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(d_states) & 3u) || (reinterpret_cast<uintptr_t>(d_next) & 3u)) return DCB_ERR_ALIGN;
  return next_state_device(env, d_states, n, action, d_next, S(stream));
}
int dcb_is_solved(int env, const uint8_t *d_states, int64_t n, uint8_t *d_solved, void *stream) {
  if (!env_ok(env)) return DCB_ERR_BAD_ENV;
  if (n < 0 || (n > 0 && (!d_states || !d_solved))) return DCB_ERR_BAD_ARG;
  if (reinterpret_cast<uintptr_t>(d_states) & 3u) return DCB_ERR_ALIGN;
  return is_solved_device(env, d_states, n, d_solved, S(stream));
}
int dcb_hash_states(int env, const uint8_t *d_states, int64_t n, uint64_t *d_hash, void *stream) {
  if (!env_ok(env)) return DCB_ERR_BAD_ENV;
  if (n < 0 || (n > 0 && (!d_states || !d_hash))) return DCB_ERR_BAD_ARG;
  if (reinterpret_cast<uintptr_t>(d_states) & 3u) return DCB_ERR_ALIGN;
  return hash_states_device(env, d_states, n, d_hash, S(stream));
}
int dcb_nnet_input(int env, const uint8_t *d_states, int64_t n, uint8_t *d_out, void *stream) {
  if (!env_ok(env)) return DCB_ERR_BAD_ENV;
  if (n < 0 || (n > 0 && (!d_states || !d_out))) return DCB_ERR_BAD_ARG;
  if (!aligned16(d_states) || !aligned16(d_out)) return DCB_ERR_ALIGN;
  return nnet_input_device(env, d_states, n, d_out, S(stream));
}

// ---- host-buffer entry points ------------------------------------------------------------------------
#define DCB_TRY(expr)                                       \
  do {                                                      \
    const int rc__ = dcb_record_cuda(expr);                 \
    if (rc__) { rc = rc__; goto done; }                     \
  } while (0)

int dcb_expand_host(int env, const uint8_t *h_parents, int64_t n, uint8_t *h_children, uint8_t *h_solved, uint64_t *h_hash,
                    int device) {
  if (!env_ok(env)) return DCB_ERR_BAD_ENV;
  if (n < 0 || (n > 0 && (!h_parents || !h_children))) return DCB_ERR_BAD_ARG;
  if (n == 0) return DCB_OK;
  const int64_t s = kStateBytes[env], a = kNumMoves[env];
  uint8_t *d_par = nullptr, *d_ch = nullptr, *d_sv = nullptr;
  uint64_t *d_h = nullptr;
  cudaStream_t st = nullptr;
  int rc = DCB_OK;
  DCB_TRY(cudaSetDevice(device));
  DCB_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  DCB_TRY(cudaMalloc(&d_par, n * s + 16));
  DCB_TRY(cudaMalloc(&d_ch, n * a * s + 16));
  if (h_solved) DCB_TRY(cudaMalloc(&d_sv, n * a + 16));
  if (h_hash) DCB_TRY(cudaMalloc(&d_h, n * a * 8 + 16));
  DCB_TRY(cudaMemcpyAsync(d_par, h_parents, n * s, cudaMemcpyHostToDevice, st));
  rc = expand_device(env, d_par, nullptr, n, d_ch, d_sv, d_h, st);
  if (rc) goto done;
  DCB_TRY(cudaMemcpyAsync(h_children, d_ch, n * a * s, cudaMemcpyDeviceToHost, st));
  if (h_solved) DCB_TRY(cudaMemcpyAsync(h_solved, d_sv, n * a, cudaMemcpyDeviceToHost, st));
  if (h_hash) DCB_TRY(cudaMemcpyAsync(h_hash, d_h, n * a * 8, cudaMemcpyDeviceToHost, st));
  DCB_TRY(cudaStreamSynchronize(st));
done:
  cudaFree(d_par); cudaFree(d_ch); cudaFree(d_sv); cudaFree(d_h);
  if (st) cudaStreamDestroy(st);
  return rc;
}

int dcb_next_state_host(int env, const uint8_t *h_states, int64_t n, int action, uint8_t *h_next, int device) {
  if (!env_ok(env)) return DCB_ERR_BAD_ENV;
  if (n < 0 || action < 0 || action >= kNumMoves[env] || (n > 0 && (!h_states || !h_next))) return DCB_ERR_BAD_ARG;
  if (n == 0) return DCB_OK;
  const int64_t s = kStateBytes[env];
  uint8_t *d_in = nullptr, *d_out = nullptr;
  int rc = DCB_OK;
  DCB_TRY(cudaSetDevice(device));
  DCB_TRY(cudaMalloc(&d_in, n * s + 16));
  DCB_TRY(cudaMalloc(&d_out, n * s + 16));
  DCB_TRY(cudaMemcpy(d_in, h_states, n * s, cudaMemcpyHostToDevice));
  rc = next_state_device(env, d_in, n, action, d_out, nullptr);
  if (rc) goto done;
  DCB_TRY(cudaMemcpy(h_next, d_out, n * s, cudaMemcpyDeviceToHost));
done:
  cudaFree(d_in); cudaFree(d_out);
  return rc;
}

int dcb_is_solved_host(int env, const uint8_t *h_states, int64_t n, uint8_t *h_solved, int device) {
  if (!env_ok(env)) return DCB_ERR_BAD_ENV;
  if (n < 0 || (n > 0 && (!h_states || !h_solved))) return DCB_ERR_BAD_ARG;
  if (n == 0) return DCB_OK;
  const int64_t s = kStateBytes[env];
  uint8_t *d_in = nullptr, *d_out = nullptr;
  int rc = DCB_OK;
  DCB_TRY(cudaSetDevice(device));
  DCB_TRY(cudaMalloc(&d_in, n * s + 16));
  DCB_TRY(cudaMalloc(&d_out, n + 16));
  DCB_TRY(cudaMemcpy(d_in, h_states, n * s, cudaMemcpyHostToDevice));
  rc = is_solved_device(env, d_in, n, d_out, nullptr);
  if (rc) goto done;
  DCB_TRY(cudaMemcpy(h_solved, d_out, n, cudaMemcpyDeviceToHost));
done:
  cudaFree(d_in); cudaFree(d_out);
  return rc;
}

// ---- CLOSED ---------------------------------------------------------------------------------------------
int64_t dcb_closed_bytes(int64_t capacity) { return pow2(capacity) ? capacity * 16 : DCB_ERR_BAD_ARG; }
int dcb_closed_clear(void *d_table, int64_t capacity, void *stream) {
  if (!d_table || !pow2(capacity)) return DCB_ERR_BAD_ARG;
  if (!aligned16(d_table)) return DCB_ERR_ALIGN;
  return closed_clear_device(d_table, capacity, S(stream));
}
int64_t dcb_closed_scratch_bytes(int64_t m) { return m >= 0 ? closed_scratch_bytes(m) : DCB_ERR_BAD_ARG; }
int dcb_closed_insert(int env, void *d_table, int64_t capacity, const uint8_t *d_arena, const uint64_t *d_hash,
                      const uint32_t *d_g, const uint8_t *d_valid, uint32_t first_id, int64_t m, void *d_scratch,
                      uint8_t *d_keep, uint32_t *d_num_entries, void *stream) {
  if (!env_ok(env)) return DCB_ERR_BAD_ENV;
  // slot indices are 32-bit with 0xffffffff = "no slot": the table stops at 2^31 slots (32 GB)
  if (m < 0 || !pow2(capacity) || capacity > (int64_t(1) << 31) || (m > 0 && (!d_table || !d_arena || !d_hash || !d_g || !d_scratch || !d_keep)))
    return DCB_ERR_BAD_ARG;
  if (!aligned16(d_table) || !aligned16(d_scratch) || (reinterpret_cast<uintptr_t>(d_arena) & 3u)) return DCB_ERR_ALIGN;
  return closed_insert_device(env, d_table, capacity, d_arena, d_hash, d_g, d_valid, first_id, m, d_scratch, d_keep, d_num_entries,
                              S(stream));
}
int dcb_closed_rehash(const void *d_old, int64_t old_capacity, void *d_new, int64_t new_capacity, void *stream) {
  if (!d_old || !d_new || !pow2(old_capacity) || !pow2(new_capacity)) return DCB_ERR_BAD_ARG;
  return closed_rehash_device(d_old, old_capacity, d_new, new_capacity, S(stream));
}

// ---- OPEN ------------------------------------------------------------------------------------------------
int dcb_open_clear(dcb_open_state *d_state, void *stream) {
  if (!d_state) return DCB_ERR_BAD_ARG;
  return open_clear_device(d_state, 1, S(stream));
}
int dcb_open_push(dcb_open_state *d_state, uint32_t *d_key, uint32_t *d_id, int64_t capacity, const float *d_cost,
                  const uint32_t *d_ids, uint32_t first_id, const uint8_t *d_keep, int64_t m, void *stream) {
  if (m < 0 || capacity <= 0 || capacity > 0xFFFFFFFFll || (m > 0 && (!d_state || !d_key || !d_id || !d_cost))) return DCB_ERR_BAD_ARG;
  return open_push_device(d_state, d_key, nullptr, d_id, capacity, d_cost, d_ids, first_id, d_keep, m, S(stream));
}
int64_t dcb_open_scratch_bytes(int64_t capacity, int64_t batch) {
  return (capacity > 0 && batch > 0) ? open_scratch_bytes(capacity, batch, 1) : DCB_ERR_BAD_ARG;
}
int dcb_open_pop(dcb_open_state *d_state, uint32_t *d_key, uint32_t *d_id, int64_t capacity, int32_t batch, int stop_at_goal,
                 const uint8_t *d_node_solved, uint32_t *d_popped_ids, void *d_scratch, void *stream) {
  if (!d_state || !d_key || !d_id || !d_popped_ids || !d_scratch || batch <= 0 || capacity <= 0) return DCB_ERR_BAD_ARG;
  if (stop_at_goal && !d_node_solved) return DCB_ERR_BAD_ARG;
  if (!aligned16(d_scratch)) return DCB_ERR_ALIGN;
  return open_pop_device(d_state, d_key, nullptr, d_id, capacity, 1, batch, -1, stop_at_goal, 0, 0, d_node_solved, nullptr, d_popped_ids, batch, d_scratch,
                         nullptr, S(stream));
}

// ---- node bookkeeping --------------------------------------------------------------------------------------
int dcb_child_meta(int env, const uint32_t *d_parent_ids, int64_t n_parents, uint32_t first_id, uint32_t *d_node_g,
                   uint32_t *d_slot_parent, void *stream) {
  if (!env_ok(env)) return DCB_ERR_BAD_ENV;
  if (n_parents < 0 || first_id % kNumMoves[env] || (n_parents > 0 && (!d_parent_ids || !d_node_g || !d_slot_parent))) return DCB_ERR_BAD_ARG;
  return child_meta_device(d_parent_ids, n_parents, kNumMoves[env], first_id, d_node_g, d_slot_parent, S(stream));
}
int dcb_compact_kept(const uint8_t *d_keep, uint32_t first_id, int64_t m, uint32_t *d_out_ids, uint32_t *d_counter, void *stream) {
  if (m < 0 || (m > 0 && (!d_keep || !d_out_ids || !d_counter))) return DCB_ERR_BAD_ARG;
  return compact_kept_device(d_keep, first_id, m, d_out_ids, d_counter, S(stream));
}
int dcb_gather_nnet_input(int env, const uint8_t *d_arena, const uint32_t *d_ids, int64_t m, uint8_t *d_out, void *stream) {
  if (!env_ok(env)) return DCB_ERR_BAD_ENV;
  if (m < 0 || (m > 0 && (!d_arena || !d_ids || !d_out))) return DCB_ERR_BAD_ARG;
  return gather_nnet_device(env, d_arena, d_ids, m, d_out, S(stream));
}
int dcb_compute_cost(const float *d_h, const uint32_t *d_ids, const uint32_t *d_node_g, const uint8_t *d_node_solved,
                     float weight, int64_t m, float *d_cost, void *stream) {
  if (m < 0 || (m > 0 && (!d_h || !d_ids || !d_node_g || !d_node_solved || !d_cost))) return DCB_ERR_BAD_ARG;
  return cost_device(d_h, d_ids, d_node_g, d_node_solved, weight, m, d_cost, S(stream));
}
int dcb_reconstruct_path(int env, const uint32_t *d_slot_parent, uint32_t goal_id, int32_t max_len, uint8_t *d_moves,
                         int32_t *d_len, void *stream) {
  if (!env_ok(env)) return DCB_ERR_BAD_ENV;
  if (!d_slot_parent || !d_moves || !d_len || max_len <= 0) return DCB_ERR_BAD_ARG;
  return path_device(d_slot_parent, goal_id, kNumMoves[env], max_len, d_moves, d_len, S(stream));
}

// ---- device-driven search iteration ---------------------------------------------------------------------------
static int ctx_ok(const dcb_search_ctx *c) {
  if (!c) return DCB_ERR_BAD_ARG;
  if (!env_ok(c->env)) return DCB_ERR_BAD_ENV;
  const int64_t a = kNumMoves[c->env];
  if (c->n_inst <= 0 || c->n_inst > 65535 || c->batch <= 0 || (c->semantics != 0 && c->semantics != 1)) return DCB_ERR_BAD_ARG;
  if (c->slots_per_inst < 64 || c->slots_per_inst % 32 || (int64_t)c->n_inst * c->slots_per_inst * a >= (int64_t(1) << 32)) return DCB_ERR_BAD_ARG;
  if ((c->semantics == 1) != (c->d_open_key_lo != nullptr)) return DCB_ERR_BAD_ARG;    // 64-bit keys iff float64 costs
  if (c->open_per_inst == 0 || !pow2(c->closed_capacity) || c->closed_capacity > (int64_t(1) << 31)) return DCB_ERR_BAD_ARG;
  if (!c->d_arena || !c->d_node_g || !c->d_node_solved || !c->d_slot_parent || !c->d_closed || !c->d_open_key || !c->d_open_id || !c->d_inst ||
      !c->d_plan || !c->d_weights || !c->d_popped_ids || !c->d_tiles || !c->d_hash || !c->d_kept_ids || !c->d_pop_scratch || !c->d_closed_scratch)
    return DCB_ERR_BAD_ARG;
  if (!aligned16(c->d_arena) || !aligned16(c->d_closed) || !aligned16(c->d_hash) || !aligned16(c->d_tiles) || !aligned16(c->d_pop_scratch) ||
      !aligned16(c->d_closed_scratch) || !aligned16(c->d_inst) || !aligned16(c->d_plan) || (reinterpret_cast<uintptr_t>(c->d_node_solved) & 3u))
    return DCB_ERR_ALIGN;
  return DCB_OK;
}
int64_t dcb_search_pop_scratch_bytes(int32_t n_inst, int64_t open_per_inst, int32_t batch) {
  return (n_inst > 0 && open_per_inst > 0 && batch > 0) ? search_pop_scratch_bytes(n_inst, open_per_inst, batch) : DCB_ERR_BAD_ARG;
}
int dcb_search_reset(const dcb_search_ctx *ctx, const uint8_t *d_roots, void *stream) {
  const int rc = ctx_ok(ctx);
  if (rc) return rc;
  if (!d_roots) return DCB_ERR_BAD_ARG;
  return search_reset_device(*ctx, d_roots, S(stream));
}
int dcb_search_pop(const dcb_search_ctx *ctx, int include_solved, void *stream) {
  const int rc = ctx_ok(ctx);
  return rc ? rc : search_pop_device(*ctx, include_solved, S(stream));
}
int dcb_search_expand(const dcb_search_ctx *ctx, void *stream) {
  const int rc = ctx_ok(ctx);
  return rc ? rc : search_expand_device(*ctx, S(stream));
}
int dcb_search_closed(const dcb_search_ctx *ctx, void *stream) {
  const int rc = ctx_ok(ctx);
  return rc ? rc : search_closed_device(*ctx, S(stream));
}
int dcb_search_push(const dcb_search_ctx *ctx, const float *d_h, const float *d_dot_partial, int32_t n_parts, float dot_bias, void *stream) {
  const int rc = ctx_ok(ctx);
  if (rc) return rc;
  if ((!d_h && !d_dot_partial) || (d_dot_partial && n_parts <= 0)) return DCB_ERR_BAD_ARG;
  return search_push_device(*ctx, d_h, d_dot_partial, n_parts, dot_bias, S(stream));
}
int dcb_search_path(const dcb_search_ctx *ctx, uint32_t node_id, int32_t max_len, uint8_t *d_moves, int32_t *d_len, void *stream) {
  const int rc = ctx_ok(ctx);
  if (rc) return rc;
  if (!d_moves || !d_len || max_len <= 0) return DCB_ERR_BAD_ARG;
  return search_path_device(*ctx, node_id, max_len, d_moves, d_len, S(stream));
}

// ---- cost-to-go network: tcgen05 dense layers ----------------------------------------------------------------
int dcb_resnet_gemm_ex(const void *d_a_hi, const void *d_a_lo, int64_t lda, const void *d_w_hi, const void *d_w_lo, int64_t ldw,
                       const float *d_bias, float scale, const void *d_skip_hi, const void *d_skip_lo, int relu, void *d_out_hi,
                       void *d_out_lo, float *d_out_f32, const float *d_partial_in, float *d_partial_out, const float *d_dot_w,
                       float *d_dot_partial, int64_t m, int32_t n_padded, int32_t k_padded, const int32_t *d_m_count, int32_t m_offset,
                       int32_t k_chunk, void *d_scratch, void *stream) {
  if (m < 0 || n_padded <= 0 || k_padded <= 0 || n_padded % 256 || k_padded % 64 || lda < k_padded || ldw < k_padded || (lda % 8) || (ldw % 8))
    return DCB_ERR_BAD_ARG;
  if (m > 0 && (!d_a_hi || !d_w_hi || (!d_partial_out && (!d_bias || (!d_out_hi && !d_dot_w && !d_out_f32))) || (d_dot_w && !d_dot_partial))) return DCB_ERR_BAD_ARG;
  if (k_chunk < 0 || (k_chunk % 64) || (k_chunk > 0 && k_chunk < k_padded && !d_scratch)) return DCB_ERR_BAD_ARG;
  if (d_out_lo && !d_out_hi) return DCB_ERR_BAD_ARG;
  if (!aligned16(d_a_hi) || !aligned16(d_a_lo) || !aligned16(d_w_hi) || !aligned16(d_w_lo) || !aligned16(d_skip_hi) || !aligned16(d_skip_lo) ||
      !aligned16(d_out_hi) || !aligned16(d_out_lo) || !aligned16(d_out_f32) || !aligned16(d_partial_in) || !aligned16(d_partial_out) ||
      !aligned16(d_scratch))
    return DCB_ERR_ALIGN;
  return resnet_gemm_device(d_a_hi, d_a_lo, lda, d_w_hi, d_w_lo, ldw, d_bias, scale, d_skip_hi, d_skip_lo, relu, d_out_hi, d_out_lo, d_out_f32,
                            d_partial_in, d_partial_out, d_dot_w, d_dot_partial, m, n_padded, k_padded, d_m_count, m_offset, k_chunk, d_scratch,
                            S(stream));
}
int dcb_resnet_gemm(const void *d_a_hi, const void *d_a_lo, int64_t lda, const void *d_w_hi, const void *d_w_lo, int64_t ldw,
                    const float *d_bias, float scale, const void *d_skip_hi, const void *d_skip_lo, int relu, void *d_out_hi,
                    void *d_out_lo, float *d_out_f32, const float *d_partial_in, float *d_partial_out, const float *d_dot_w,
                    float *d_dot_partial, int64_t m, int32_t n_padded, int32_t k_padded, void *stream) {
  return dcb_resnet_gemm_ex(d_a_hi, d_a_lo, lda, d_w_hi, d_w_lo, ldw, d_bias, scale, d_skip_hi, d_skip_lo, relu, d_out_hi, d_out_lo, d_out_f32,
                            d_partial_in, d_partial_out, d_dot_w, d_dot_partial, m, n_padded, k_padded, nullptr, 0, 0, nullptr, stream);
}
int64_t dcb_resnet_gemm_scratch_bytes(void) { return resnet_gemm_scratch_bytes(); }
int dcb_onehot_fp16(const uint8_t *d_nnet_in, int64_t m, int32_t state_dim, int32_t depth, int32_t k_padded, void *d_out, void *stream) {
  if (m < 0 || state_dim <= 0 || depth <= 0 || k_padded < state_dim * depth || k_padded % 64 || (m > 0 && (!d_nnet_in || !d_out))) return DCB_ERR_BAD_ARG;
  if (!aligned16(d_out)) return DCB_ERR_ALIGN;
  return onehot_device(d_nnet_in, m, state_dim, depth, k_padded, d_out, S(stream));
}
int dcb_onehot_fp16_nodes(int env, const uint8_t *d_arena, const uint32_t *d_ids, int64_t m, int32_t depth, int32_t k_padded, void *d_out,
                          void *stream) {
  if (!env_ok(env)) return DCB_ERR_BAD_ENV;
  const int s = kStateBytes[env];
  if (m < 0 || depth <= 0 || k_padded < s * depth || k_padded % 64 || (m > 0 && (!d_arena || !d_ids || !d_out))) return DCB_ERR_BAD_ARG;
  if (!aligned16(d_out)) return DCB_ERR_ALIGN;
  return onehot_gather_device(env, d_arena, d_ids, m, s, depth, k_padded, d_out, nullptr, 0, S(stream));
}
int dcb_onehot_fp16_nodes_ex(int env, const uint8_t *d_arena, const uint32_t *d_ids, int64_t m, int32_t depth, int32_t k_padded, void *d_out,
                             const int32_t *d_m_count, int32_t m_offset, void *stream) {
  if (!env_ok(env)) return DCB_ERR_BAD_ENV;
  const int s = kStateBytes[env];
  if (m < 0 || depth <= 0 || k_padded < s * depth || k_padded % 64 || (m > 0 && (!d_arena || !d_ids || !d_out))) return DCB_ERR_BAD_ARG;
  if (!aligned16(d_out)) return DCB_ERR_ALIGN;
  return onehot_gather_device(env, d_arena, d_ids, m, s, depth, k_padded, d_out, d_m_count, m_offset, S(stream));
}
int dcb_rowdot(const void *d_x_hi, const void *d_x_lo, const float *d_w, float bias, int64_t m, int32_t n_valid, int32_t ld, float *d_out,
               void *stream) {
  if (m < 0 || n_valid <= 0 || ld < n_valid || (ld & 7) || (m > 0 && (!d_x_hi || !d_w || !d_out))) return DCB_ERR_BAD_ARG;
  if (!aligned16(d_x_hi) || !aligned16(d_x_lo)) return DCB_ERR_ALIGN;
  return rowdot_device(d_x_hi, d_x_lo, d_w, bias, m, n_valid, ld, d_out, S(stream));
}

}  // extern "C"
