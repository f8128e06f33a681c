/* dcb.h -- C ABI of libdcb_b200.so: the B200 (sm_100a) implementation of DeepCubeA's batched
 * environment step and batched weighted A* (BWAS) node expansion.
 *
 * This header is the drop-in boundary.  The reference has no in-process native library for this path:
 * its native side is a child PROCESS (cpp/parallel_weighted_astar.cpp) driven over argv
 * (parallel_weighted_astar.cpp:352-356), a positional stdout protocol (search_methods/astar.py:529-532)
 * and an AF_UNIX request/response socket (parallel_weighted_astar.cpp:121-136, 275-279;
 * astar.py:571-616).  Each entry point below names the reference code it replaces.  INTEGRATION.md
 * shows the ctypes binding a maintainer would add to search_methods/astar.py / environments/*.py.
 *
 * Conventions
 *   - plain C types only; every `d_` pointer is a DEVICE pointer owned by the caller (PyTorch in this
 *     repo), every `h_` pointer is a HOST pointer; the library never frees caller memory.
 *   - `stream` is a cudaStream_t passed as void*; all device entry points are asynchronous on it and
 *     touch no global mutable state (move tables are compile-time constants), so they are thread-safe
 *     per (device, stream).
 *   - return value: DCB_OK (0) or a negative DCB_ERR_* code.  Nothing throws, exits or prints.
 *   - layouts: states are packed uint8, `state_bytes` per state, no padding (cube3 54, puzzle15 16,
 *     puzzle24 25, puzzle35 36, puzzle48 49, lightsout7 49, cube4 96).  Children of parent p are contiguous, move-minor:
 *     children[(p*num_moves + a)*state_bytes ...] -- the order of Environment.expand
 *     (environments/cube3.py:129-161) and of getNextStates (cpp/environments.cpp:236-243).
 */
#ifndef DCB_H_
#define DCB_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* the library is built with -fvisibility=hidden */
#endif

#define DCB_ABI_VERSION 1

/* ---- status codes ------------------------------------------------------------------------------ */
#define DCB_OK 0
#define DCB_ERR_BAD_ENV (-1)      /* unknown env id                                                  */
#define DCB_ERR_BAD_ARG (-2)      /* null pointer, negative count, bad action, bad capacity          */
#define DCB_ERR_ALIGN (-3)        /* a pointer that must be 16-byte aligned is not                   */
#define DCB_ERR_CUDA (-4)         /* a CUDA runtime call failed; see dcb_last_cuda_error()           */
#define DCB_ERR_NO_DEVICE (-5)    /* no CUDA device / wrong architecture (needs sm_100)              */
#define DCB_ERR_CAPACITY (-6)     /* arena / closed table / open set would overflow                  */

/* ---- environments (utils/env_utils.py:6-28; cpp main :375-390) --------------------------------- */
#define DCB_ENV_CUBE3 0
#define DCB_ENV_PUZZLE15 1
#define DCB_ENV_PUZZLE24 2
#define DCB_ENV_PUZZLE35 3
#define DCB_ENV_PUZZLE48 4
#define DCB_ENV_LIGHTSOUT7 5      /* environments/lights_out.py; LightsOut in cpp/environments.cpp:133-208 */
#define DCB_ENV_CUBE4 6           /* Cube4 in cpp/environments.cpp:262-370, cpp/parallel_weighted_astar.cpp:386 (C++ only in the reference;
                                    96 sticker ids, 24 quarter turns; solved = one colour (id / 16) per face) */
#define DCB_NUM_ENVS 7

int dcb_abi_version(void);
const char *dcb_error_string(int code);
/* cudaGetErrorString of the last CUDA failure seen on the calling thread ("" if none). */
const char *dcb_last_cuda_error(void);

/* Environment.get_num_moves (cube3.py:87-88, n_puzzle.py:91-92); getNumActions (environments.cpp). */
int dcb_env_num_moves(int env);
/* bwas_cpp's state_dim table (astar.py:473-486). */
int dcb_env_state_bytes(int env);
/* Goal state (cube3.py:37 arange(54); n_puzzle.py:41 [1..n*n-1,0]; cube4: arange(96), one of its solved states).
 * h_out: state_bytes bytes. */
int dcb_env_goal_state(int env, uint8_t *h_out);
/* The move tables the kernels were compiled with, for auditing against the reference's:
 * cube3: perm[12][54] with child[j] = parent[perm[a][j]] (cube3.py:163-171, environments.h:75-105); cube4: perm[24][96]
 * (environments.cpp:262-341);
 * puzzles: swap_zero_idxs[n*n][4] (n_puzzle.py:174-214, environments.cpp:4-46); lightsout7: move_matrix[49][5]
 * (lights_out.py:31-42).  h_out: int32. */
int dcb_env_move_table(int env, int32_t *h_out, int64_t capacity_elems);

/* ---- batched environment step, device buffers --------------------------------------------------- */
/* Environment.expand + is_solved on every child + state hash, one launch.
 * Replaces Cube3.expand/_move_np (cube3.py:129-171), NPuzzle.expand/_move_np (n_puzzle.py:136-231),
 * Cube3::getNextStates/isSolved (environments.cpp:222-256), PuzzleN::getNextStates/isSolved (:92-126)
 * and the OpenMP child loop of parallel_weighted_astar.cpp:217-230.
 *   d_parents  [n][S]        d_children [n][A][S] (16-byte aligned)
 *   d_solved   [n][A] u8 (0/1), may be NULL      d_hash [n][A] u64, may be NULL (16-byte aligned)
 * Hash: project-defined 64-bit state hash (never 0); the reference's hashes (boost::hash_range,
 * parallel_weighted_astar.cpp:104-111; CPython bytes hash, cube3.py:17-21) are unobservable. */
int dcb_expand(int env, const uint8_t *d_parents, int64_t n, uint8_t *d_children, uint8_t *d_solved,
               uint64_t *d_hash, void *stream);

/* Same, but parents are gathered by node id from a node arena (state of node i at d_arena + i*S) and the
 * children are appended at d_children (normally d_arena + first_child_id*S).  This is the form the A*
 * loop uses (parallel_weighted_astar.cpp:217-230 reads popped[i]->env).  d_parent_ids [n] u32. */
int dcb_expand_indexed(int env, const uint8_t *d_arena, const uint32_t *d_parent_ids, int64_t n,
                       uint8_t *d_children, uint8_t *d_solved, uint64_t *d_hash, void *stream);

/* Environment.next_state for one action (cube3.py:48-54, n_puzzle.py:46-61; getNextState). */
int dcb_next_state(int env, const uint8_t *d_states, int64_t n, int action, uint8_t *d_next, void *stream);
/* Environment.is_solved (cube3.py:71-75, n_puzzle.py:78-82; isSolved). d_solved [n] u8.  cube4 follows Cube4::isSolved
 * (environments.cpp:356-366): every face one colour (id / 16), not sticker identity. */
int dcb_is_solved(int env, const uint8_t *d_states, int64_t n, uint8_t *d_solved, void *stream);
/* Project-defined state hash of arbitrary states. d_hash [n] u64. */
int dcb_hash_states(int env, const uint8_t *d_states, int64_t n, uint64_t *d_hash, void *stream);
/* Environment.state_to_nnet_input (cube3.py:77-85: colors/9; n_puzzle.py:84-89: tiles) and the same
 * conversion in cpp_listener (astar.py:598-602).  d_out [n][S] u8. */
int dcb_nnet_input(int env, const uint8_t *d_states, int64_t n, uint8_t *d_out, void *stream);

/* ---- batched environment step, HOST buffers (the call a ctypes/cgo-style binding makes) ----------
 * Copies h_parents to the device, runs dcb_expand, copies results back, synchronises.  `device` is the
 * CUDA ordinal.  Scratch device memory is allocated and freed inside the call. */
int dcb_expand_host(int env, const uint8_t *h_parents, int64_t n, uint8_t *h_children, uint8_t *h_solved,
                    uint64_t *h_hash, int device);
int dcb_next_state_host(int env, const uint8_t *h_states, int64_t n, int action, uint8_t *h_next, int device);
int dcb_is_solved_host(int env, const uint8_t *h_states, int64_t n, uint8_t *h_solved, int device);

/* ---- CLOSED: open-addressing hash table in HBM ---------------------------------------------------
 * Replaces std::unordered_set<Node*,Hash,NodePointerEq> closed and the serial find / insert /
 * "depth improved -> re-open" loop of parallel_weighted_astar.cpp:142, 243-265, and Instance.closed_dict /
 * remove_in_closed of astar.py:55, 78-90.
 * Table = `capacity` (power of two) 16-byte slots {u64 key, u64 val}; key = state hash (0 = empty),
 * val = (g << 32) | node_id.  Caller allocates dcb_closed_bytes(capacity) bytes and clears the table
 * with dcb_closed_clear. */
int64_t dcb_closed_bytes(int64_t capacity);
int dcb_closed_clear(void *d_table, int64_t capacity, void *stream);
/* Insert-or-improve for a batch of m candidate nodes whose states already sit in the arena.
 *   d_hash [m] u64, d_g [m] u32 (path cost = depth), node id of candidate i = first_id + i.
 *   d_valid [m] u8 or NULL: candidates with valid==0 are ignored (padding).
 * Result d_keep[i] = 1 iff the reference's child-order loop (:246-261; remove_in_closed, astar.py:78-90) keeps candidate i:
 * its state was never seen, or it is reached with a STRICTLY smaller g than the table holds AT THAT POINT OF THE LOOP --
 * i.e. smaller than the value stored before this batch and than every earlier (lower id) candidate of the same state in
 * this batch.  The table ends up holding the smallest (g, id) per state.  Equal hashes are verified by comparing the S
 * state bytes in the arena; on a true 64-bit collision the candidate is kept (never wrongly dropped).
 * d_scratch: dcb_closed_scratch_bytes(m) bytes, 16-byte aligned.  Four launches (probe, min, resolve, in-batch fix-up). */
int64_t dcb_closed_scratch_bytes(int64_t m);
int dcb_closed_insert(int env, void *d_table, int64_t capacity, const uint8_t *d_arena,
                      const uint64_t *d_hash, const uint32_t *d_g, const uint8_t *d_valid, uint32_t first_id,
                      int64_t m, void *d_scratch, uint8_t *d_keep, uint32_t *d_num_entries, void *stream);
/* Re-insert every entry of an old table into a (larger, cleared) new one. */
int dcb_closed_rehash(const void *d_old, int64_t old_capacity, void *d_new, int64_t new_capacity, void *stream);

/* ---- OPEN: bucket priority queue in HBM ----------------------------------------------------------
 * Replaces std::priority_queue<Node*,...,compareNodeCost> open (parallel_weighted_astar.cpp:141), the pop
 * loop :177-208 and the push loop :309-319; heapq open_set of astar.py:53, 64-76.
 * Storage: flat unsorted arrays d_key [capacity] u32 (cost as non-negative float bits: order-preserving)
 * and d_id [capacity] u32.  A pop is an exact radix select of the `batch` smallest (key,id) composites:
 * two 4096-bucket histogram passes over key bits 31..20 / 19..8, a single-block finish on the boundary
 * bucket, one partition pass.  Ties on cost break towards the smaller node id (the reference's C++ heap
 * order is unspecified; Python's heapq is FIFO, which this matches). */
typedef struct dcb_search_inst {     /* one per problem instance; lives in device memory; the host may copy it back (128 bytes) */
  uint32_t open_size;                /* live entries of this instance's OPEN segment                               */
  uint32_t n_popped;                 /* nodes returned by the last pop (C++ semantics: after the break at the first solved one) */
  uint32_t n_expand;                 /* parents the current iteration expands (0 once `done` fired / the arena is full) */
  uint32_t min_key;                  /* last pop: smallest key popped == popped[0]->cost                          */
  uint32_t goal_id;                  /* C++: cheapest solved node popped so far; Python: solved popped node of smallest path
                                        cost (astar.py:327-333); 0xffffffff = none                                */
  uint32_t goal_key;                 /* C++: its cost bits; Python: its path cost                                  */
  uint32_t done;                     /* 0 running | 1 finished (:205-208 / :191-193; Python: a solved node was popped, astar.py:73)
                                        | 2 OPEN exhausted | 3 node arena full | 4 OPEN segment full               */
  uint32_t n_goals;                  /* Python: solved nodes popped so far (Instance.goal_nodes)                   */
  uint32_t next_slot;                /* next free arena slot, counted inside the instance's own slot range         */
  uint32_t base_slot;                /* this iteration's first child record (instance-local slot)                  */
  uint32_t iterations;               /* pops so far                                                                */
  uint32_t tile_off;                 /* first 32-parent tile of this instance in the current iteration's tile list */
  uint64_t nodes_generated;          /* the reference's counter: C++ 1 + sum popped*A incl. the terminating iteration (:166, :266);
                                        Python sum popped*A (astar.py:168)                                         */
  uint64_t nodes_expanded;           /* children actually materialised in the arena                                */
  uint32_t thr_key, thr_id;          /* last pop: select threshold (everything <= it was removed)                  */
  uint32_t need, prefix, cand_count, n_holes, n_surv, take_all, n_at_pop, n_take, resting;   /* pop-internal     */
  uint32_t overflow;                 /* a push ran past the segment's capacity                                     */
  uint32_t thr_lo;                   /* select threshold, low key word (wide keys)                                 */
  uint32_t reserved[3];
} dcb_search_inst;
typedef dcb_search_inst dcb_open_state;   /* a stand-alone OPEN queue is a search instance that uses only the OPEN fields */
int dcb_open_clear(dcb_open_state *d_state, void *stream);
/* Append m entries with cost d_cost[i] (float32 >= 0) and node id (d_ids ? d_ids[i] : first_id + i);
 * entries with d_keep[i]==0 are skipped (d_keep may be NULL). */
int dcb_open_push(dcb_open_state *d_state, uint32_t *d_key, uint32_t *d_id, int64_t capacity,
                  const float *d_cost, const uint32_t *d_ids, uint32_t first_id, const uint8_t *d_keep,
                  int64_t m, void *stream);
/* (The search loop also keeps 64-bit keys -- float64 costs of the Python AStar, astar.py:196 -- as a second array of low words,
 * dcb_search_ctx.d_open_key_lo; composite = (key, key_lo, id).  The stand-alone queue uses float32 keys.) */
/* Remove the (up to) `batch` cheapest entries and write their node ids, in cost order, to d_popped_ids
 * [batch]; d_state->n_popped says how many.  With `stop_at_goal` the pop is truncated right after the
 * first SOLVED entry in cost order (the `break` at :190-203; the entries behind it go back to OPEN) and the
 * goal bookkeeping / termination rule of :186-208 is applied to d_state (goal_id, goal_key, done).
 * d_node_solved [node id] u8.  d_scratch: dcb_open_scratch_bytes(capacity, batch) bytes, 16-byte aligned. */
int64_t dcb_open_scratch_bytes(int64_t capacity, int64_t batch);   /* one instance; see dcb_search_pop_scratch_bytes */
int dcb_open_pop(dcb_open_state *d_state, uint32_t *d_key, uint32_t *d_id, int64_t capacity, int32_t batch,
                 int stop_at_goal, const uint8_t *d_node_solved, uint32_t *d_popped_ids, void *d_scratch,
                 void *stream);

/* ---- node bookkeeping ----------------------------------------------------------------------------
 * Node ids: id = slot * A + move, state at d_arena + id * S.  Slot 0 holds the root (id 0); each batch of
 * popped parents gets consecutive slots starting at a multiple of dcb_env_slot_align(env), so that its
 * child block starts 16-byte aligned for the TMA store of dcb_expand_indexed. */
int dcb_env_slot_align(int env);
/* Node{depth,parent} of parallel_weighted_astar.cpp:80-86, 219-226: for child i (0 <= i < n_parents*A) of
 * popped parent d_parent_ids[i/A]:  d_node_g[first_id+i] = d_node_g[parent] + 1, and per slot
 * d_slot_parent[(first_id+i)/A] = parent.  first_id must be a multiple of A. */
int dcb_child_meta(int env, const uint32_t *d_parent_ids, int64_t n_parents, uint32_t first_id,
                   uint32_t *d_node_g, uint32_t *d_slot_parent, void *stream);
/* nodesToAdd of :244-262: ids (first_id + i) of the candidates with d_keep[i] != 0, appended (unordered) to
 * d_out_ids; *d_counter (must be zeroed by the caller) receives the count. */
int dcb_compact_kept(const uint8_t *d_keep, uint32_t first_id, int64_t m, uint32_t *d_out_ids,
                     uint32_t *d_counter, void *stream);
/* state_to_nnet_input of the listed nodes (cube3.py:77-85 / astar.py:598-602): d_out [m][S] u8. */
int dcb_gather_nnet_input(int env, const uint8_t *d_arena, const uint32_t *d_ids, int64_t m,
                          uint8_t *d_out, void *stream);
/* cost = max(h,0) * (!solved) + weight * g in float32 without FMA contraction -- the expression of
 * parallel_weighted_astar.cpp:298 (astar.py:196 in float64); clip at 0 is nnet_utils.py:193-194. */
int dcb_compute_cost(const float *d_h, const uint32_t *d_ids, const uint32_t *d_node_g,
                     const uint8_t *d_node_solved, float weight, int64_t m, float *d_cost, void *stream);
/* Path reconstruction (parallel_weighted_astar.cpp:336-341, astar.py:213-229) on the device: writes the
 * moves root->goal into d_moves [max_len] and the length into d_len (-1 if max_len was too small). */
int dcb_reconstruct_path(int env, const uint32_t *d_slot_parent, uint32_t goal_id, int32_t max_len,
                         uint8_t *d_moves, int32_t *d_len, void *stream);

/* ---- device-driven search iteration ------------------------------------------------------------------
 * One BWAS iteration (parallel_weighted_astar.cpp:169-330) / one AStar.step over ALL instances (astar.py:256-317) as a fixed
 * sequence of launches whose sizes live in device memory: nothing between the pop and the push needs the host, so an
 * iteration can be enqueued before the previous one has finished (or be captured in a CUDA graph).
 *
 * n_inst problem instances share one node arena, one CLOSED table (keys mixed with the instance number) and segmented OPEN
 * arrays.  Instance i owns arena slots [i*slots_per_inst, (i+1)*slots_per_inst); node id = global slot * A + move; its root is
 * node i*slots_per_inst*A.  The pop stage leaves, per iteration, a list of 32-parent TILES: tile t = {src, dst_slot, count,
 * inst} -- parents d_popped_ids[src .. src+count), children written to arena slots dst_slot .. dst_slot+count-1 (dst_slot is
 * a multiple of dcb_env_slot_align).  Candidate c of the iteration = child (c % A) of parent lane (c / A) % 32 of tile c / (32*A). */
typedef struct dcb_step_plan {       /* one per engine; device memory; written by the pop stage, read by the later ones (64 bytes) */
  uint32_t n_tiles;                  /* tiles of this iteration                                                    */
  uint32_t n_parents;                /* parents expanded (sum of n_expand)                                         */
  uint32_t n_kept;                   /* children that survived CLOSED == rows for the cost-to-go network           */
  uint32_t n_ambiguous;              /* CLOSED-internal: candidates decided by the in-batch fix-up                 */
  uint32_t closed_entries;           /* occupied CLOSED slots                                                      */
  uint32_t n_running;                /* instances with done == 0 after this pop                                    */
  uint32_t error;                    /* sticky bits: 1 arena full, 2 OPEN segment full, 4 CLOSED table full        */
  uint32_t budget;                   /* full-batch iterations (every instance pops `batch` nodes) the pop stage may still start; the pop
                                        rests at 0; 0xffffffff = unlimited.  Set by the host, counted down on the device, kept across
                                        dcb_search_reset: lets a host that runs one iteration ahead stop after EXACTLY k full iterations */
  uint64_t total_kept;               /* since the last reset                                                       */
  uint64_t total_expanded;           /* children materialised since the last reset                                 */
  uint64_t reserved1[2];
} dcb_step_plan;

typedef struct dcb_search_ctx {      /* HOST struct: geometry + the device buffers of one engine (all caller-owned)  */
  int32_t env, n_inst, batch, semantics;   /* semantics 0: C++ program (parallel_weighted_astar.cpp), 1: Python AStar class   */
  uint32_t slots_per_inst;           /* multiple of 32; n_inst * slots_per_inst * A < 2^32                         */
  uint32_t open_per_inst;            /* OPEN entries per instance                                                  */
  int64_t closed_capacity;           /* power of two, <= 2^31                                                      */
  uint8_t *d_arena;                  /* [n_inst*slots_per_inst*A][S]                                               */
  uint32_t *d_node_g;                /* [nodes] path cost                                                          */
  uint8_t *d_node_solved;            /* [nodes]                                                                    */
  uint32_t *d_slot_parent;           /* [slots] node id of the slot's parent                                       */
  void *d_closed;                    /* dcb_closed_bytes(closed_capacity)                                          */
  uint32_t *d_open_key, *d_open_id;  /* [n_inst][open_per_inst]                                                    */
  uint32_t *d_open_key_lo;           /* [n_inst][open_per_inst] low words of 64-bit keys (semantics 1: float64 costs, astar.py:196);
                                        NULL for semantics 0 (float32 costs, parallel_weighted_astar.cpp:298)      */
  dcb_search_inst *d_inst;           /* [n_inst]                                                                   */
  dcb_step_plan *d_plan;
  const double *d_weights;           /* [n_inst] path-cost weight of each instance (AStar(states, env, fn, weights)); semantics 0
                                        rounds it to float32 like the C++ program's argv parsing                   */
  uint32_t *d_popped_ids;            /* [n_inst][ceil32(batch)]                                                    */
  uint32_t *d_tiles;                 /* [n_inst*ceil(batch/32)] x 4 u32                                            */
  uint64_t *d_hash;                  /* [max candidates = n_inst*ceil32(batch)*A]                                  */
  uint32_t *d_kept_ids;              /* [max candidates]                                                           */
  void *d_pop_scratch;               /* dcb_search_pop_scratch_bytes                                               */
  void *d_closed_scratch;            /* dcb_closed_scratch_bytes(max candidates)                                   */
} dcb_search_ctx;
int64_t dcb_search_pop_scratch_bytes(int32_t n_inst, int64_t open_per_inst, int32_t batch);
/* Roots: d_roots [n_inst][S].  Clears the instance records and the plan, writes the roots into the arena; C++ semantics: root
 * into CLOSED with g = 0 and into OPEN with cost 0 (:160-162), nodes_generated = 1 (:166); Python semantics: root NOT in CLOSED
 * (astar.py:50-62), d_kept_ids[0..n_inst) = the root ids and plan.n_kept = n_inst so that the caller evaluates them and calls
 * dcb_search_push (root cost = heuristic, astar.py:244-249).  The CLOSED table must have been cleared (dcb_closed_clear). */
int dcb_search_reset(const dcb_search_ctx *ctx, const uint8_t *d_roots, void *stream);
/* Pop stage: segmented exact top-`batch` pop of every running instance (Python: every instance without a goal node, or all with
 * include_solved, astar.py:263-265), goal / termination bookkeeping on the device, slot assignment, tile list, plan. */
int dcb_search_pop(const dcb_search_ctx *ctx, int include_solved, void *stream);
/* Expand every tile: children + is_solved + hash (dcb_expand_indexed's kernel), depth and parent link of every child. */
int dcb_search_expand(const dcb_search_ctx *ctx, void *stream);
/* CLOSED insert-or-improve of every candidate (dcb_closed_insert's rule, keyed per instance); survivors are appended
 * (unordered) to d_kept_ids, plan.n_kept counts them. */
int dcb_search_closed(const dcb_search_ctx *ctx, void *stream);
/* cost for d_kept_ids[0..plan.n_kept) and push onto each node's instance's OPEN.  Semantics 0: float32 max(h,0)*(!solved) + weight*g
 * without FMA contraction (dcb_compute_cost; parallel_weighted_astar.cpp:298).  Semantics 1: float64 weight*g + max(h,0)*(!solved) with
 * h the float32 network output widened (astar.py:196, nnet_utils.py:172-194), kept as a 64-bit key.  The heuristic comes either as d_h [rows] or as the fused fc_out partials of dcb_resnet_gemm (d_dot_partial
 * [rows][n_parts], summed in index order, + dot_bias). */
int dcb_search_push(const dcb_search_ctx *ctx, const float *d_h, const float *d_dot_partial, int32_t n_parts, float dot_bias,
                    void *stream);
/* Moves root -> node of the instance that owns `node_id` (dcb_reconstruct_path for a shared arena). */
int dcb_search_path(const dcb_search_ctx *ctx, uint32_t node_id, int32_t max_len, uint8_t *d_moves, int32_t *d_len, void *stream);

/* ---- cost-to-go network (utils/pytorch_models.py:45-86), dense layers on tcgen05 tensor cores ------
 * Eval-mode BatchNorm is folded into the preceding Linear by the caller (deepcubea_b200/nnet/tc_resnet.py).
 * Activations and weights are fp16 "hi" arrays plus optional fp16 "lo" arrays with x = hi + lo (22 significand
 * bits).  One call = one layer (or one K chunk of it):
 *     OUT[m][n] = act( scale * ( PARTIAL_IN[m][n] + sum_k A[m][k] * W[n][k] ) + bias[n] + SKIP[m][n] )
 * with the products A_hi*W_lo (if d_w_lo), A_lo*W_hi (if d_a_lo and d_w_lo) and A_hi*W_hi swept over K in that
 * order into one fp32 TMEM accumulator (small terms first: the tensor core's accumulation truncates).
 * Row-major, K-major operands: A [m][k_padded] with leading dimension lda, W [n_padded][k_padded] with leading
 * dimension ldw (elements, multiples of 8), SKIP / OUT / PARTIAL [m][n_padded]; n_padded % 256 == 0,
 * k_padded % 64 == 0; padding rows / cols of W and bias must be zero.
 * If d_partial_out is set the call only writes the raw fp32 sum (PARTIAL_IN + A*W) there -- used to split a long K
 * into chunks of <= 1024 chained through fp32.  Otherwise OUT is written as fp16 hi (+ lo if d_out_lo) (+ fp32 if
 * d_out_f32); relu != 0 applies max(.,0).
 * Fused fc_out (pytorch_models.py:85): with d_dot_w [n_padded] set, d_dot_partial[m][n_padded/256] receives, per 256-column
 * tile, sum_n OUT[m][n] * d_dot_w[n] in fp32 (fixed order); the caller adds the tile partials and the bias.  d_out_hi may then
 * be NULL (nothing but the dot product leaves the kernel). */
int dcb_resnet_gemm(const void *d_a_hi, const void *d_a_lo, int64_t lda, const void *d_w_hi, const void *d_w_lo, int64_t ldw,
                    const float *d_bias, float scale, const void *d_skip_hi, const void *d_skip_lo, int relu, void *d_out_hi,
                    void *d_out_lo, float *d_out_f32, const float *d_partial_in, float *d_partial_out, const float *d_dot_w,
                    float *d_dot_partial, int64_t m, int32_t n_padded, int32_t k_padded, void *stream);
/* Extended form used by the search loop:
 *   d_m_count / m_offset : optional DEVICE-side row count -- the rows processed are clamp(*d_m_count - m_offset, 0, m), so that an
 *                          iteration can be enqueued (or captured in a CUDA graph) before the host knows how many children
 *                          survived CLOSED; `m` is then the capacity of the row buffers.
 *   k_chunk / d_scratch  : with 0 < k_chunk < k_padded (multiple of 64) the K sweep is folded ON CHIP in chunks of k_chunk per
 *                          TMEM accumulator (the tensor core's fp32 accumulation truncates, so one accumulator should not see
 *                          more than ~2048 K), partial sums passing through d_scratch (dcb_resnet_gemm_scratch_bytes() bytes,
 *                          16-byte aligned; per-CTA, stays in L2) instead of an [m][n] fp32 matrix in HBM. */
int dcb_resnet_gemm_ex(const void *d_a_hi, const void *d_a_lo, int64_t lda, const void *d_w_hi, const void *d_w_lo, int64_t ldw,
                       const float *d_bias, float scale, const void *d_skip_hi, const void *d_skip_lo, int relu, void *d_out_hi,
                       void *d_out_lo, float *d_out_f32, const float *d_partial_in, float *d_partial_out, const float *d_dot_w,
                       float *d_dot_partial, int64_t m, int32_t n_padded, int32_t k_padded, const int32_t *d_m_count, int32_t m_offset,
                       int32_t k_chunk, void *d_scratch, void *stream);
int64_t dcb_resnet_gemm_scratch_bytes(void);
/* F.one_hot of the nnet input (pytorch_models.py:49-52) as fp16 [m][k_padded], column = position*depth + value. */
int dcb_onehot_fp16(const uint8_t *d_nnet_in, int64_t m, int32_t state_dim, int32_t depth, int32_t k_padded, void *d_out,
                    void *stream);
/* Same for the listed nodes of a search's node arena (state of node i at d_arena + i*S): state_to_nnet_input
 * (cube3.py:77-85) and the one-hot encoding in one pass -- the form the A* loop uses for the children that survived CLOSED. */
int dcb_onehot_fp16_nodes(int env, const uint8_t *d_arena, const uint32_t *d_ids, int64_t m, int32_t depth, int32_t k_padded,
                          void *d_out, void *stream);
/* Same with a DEVICE-side row count: rows = clamp(*d_m_count - m_offset, 0, m) (see dcb_resnet_gemm_ex). */
int dcb_onehot_fp16_nodes_ex(int env, const uint8_t *d_arena, const uint32_t *d_ids, int64_t m, int32_t depth, int32_t k_padded,
                             void *d_out, const int32_t *d_m_count, int32_t m_offset, void *stream);
/* fc_out (pytorch_models.py:85): d_out[m] = sum_{n<n_valid} (x_hi+x_lo)[m][n] * d_w[n] + bias, fp32. */
int dcb_rowdot(const void *d_x_hi, const void *d_x_lo, const float *d_w, float bias, int64_t m, int32_t n_valid, int32_t ld,
               float *d_out, void *stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* DCB_H_ */
