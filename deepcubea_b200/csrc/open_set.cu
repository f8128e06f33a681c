// open_set.cu -- OPEN as a bucket priority queue in HBM: exact "pop the B cheapest" without a heap.
//
// Replaces std::priority_queue<Node*,std::vector<Node*>,compareNodeCost> open, the one-at-a-time pop
// loop (with its `break` at the first solved node) and the push loop of
// cpp/parallel_weighted_astar.cpp:141, 177-208, 309-319; heapq open_set / pop_from_open / push_to_open of
// search_methods/astar.py:53, 64-76, and the per-instance loop of astar.py:93-96 (every unsolved instance pops per step).
//
// Entries are (key = cost's float bits, id) in flat unsorted arrays.  Costs are >= 0 so unsigned key order
// is cost order; ties break towards the smaller node id, which makes every (key,id) distinct and the popped
// set unique.  A pop of the `batch` cheapest entries is a radix select over the 64-bit composite key<<32|id:
//   1. two multi-block histogram passes bucket all entries by key bits 31..20 and 19..8 (4096 buckets
//      each, shared-memory privatised) and locate the bucket holding the batch-th entry;
//   2. entries of that boundary bucket (typically a handful) are collected and a single block finishes
//      the select on their remaining 40 bits -> exact threshold T;
//   3. one partition pass removes everything <= T, back-filling the holes from the array tail;
//   4. the popped list is sorted (bucketed rank sort) into cost order (the reference's pop order, so child node ids are
//      deterministic) and the goal / termination rule of :190-208 (or astar.py:73) is applied on the device; in the C++
//      semantics entries behind the first solved pop go back to OPEN, exactly like the reference's `break`.
// HBM traffic per pop: 4 passes x 4 B per open entry; per push: 8 B per entry.
//
// WIDE KEYS: the Python AStar adds its costs in float64 (search_methods/astar.py:196).  A 64-bit cost key is kept as two arrays,
// key (the high word: what every streaming pass reads) and key_lo (read only for entries that tie on the high word); the composite
// becomes (key, key_lo, id) = 96 bits.  With key_lo == NULL (the C++ program's float32 costs) the low word is 0 everywhere.
//
// SEGMENTED: every kernel runs with gridDim.y = number of problem instances.  Instance i owns the OPEN segment
// key/id[i * seg_cap ...], the record states[i] and its own slice of the scratch buffer, so many A* instances are popped by
// the same launches as one.  Nothing here needs the host: counts stay in the per-instance records.
#include <cuda_runtime.h>
#include "dcb_internal.h"

namespace dcb {

namespace {
constexpr int kBins = 4096;
constexpr uint32_t kNone = 0xFFFFFFFFu;

typedef dcb_search_inst OpenState;      // per-instance record (include/dcb.h); dcb_open_state is the same type

__device__ __forceinline__ uint32_t warp_agg_inc(uint32_t *ctr, bool pred) {
  // one atomic per warp; returns this lane's slot (undefined if !pred)
  const unsigned m = __ballot_sync(0xffffffffu, pred);
  if (!m) return 0;
  const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(ctr, (uint32_t)__popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  return base + __popc(m & ((1u << lane) - 1));
}

__global__ void open_clear_kernel(OpenState *s) {
  if (threadIdx.x == 0) {
    OpenState z = {};
    z.goal_id = kNone;
    z.goal_key = kNone;
    s[blockIdx.x] = z;
  }
}

__global__ void __launch_bounds__(256)
open_push_kernel(OpenState *s, uint32_t *key, uint32_t *key_lo, uint32_t *id, uint32_t capacity, const float *cost, const uint32_t *ids,
                 uint32_t first_id, const uint8_t *keep, int64_t m) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const bool want = i < m && (!keep || keep[i]);
  const uint32_t pos = warp_agg_inc(&s->open_size, want);
  if (want) {
    if (pos < capacity) {
      key[pos] = __float_as_uint(cost[i]);
      if (key_lo) key_lo[pos] = 0u;
      id[pos] = ids ? ids[i] : first_id + (uint32_t)i;
    } else {
      s->overflow = 1;   // size keeps counting; the host raises DCB_ERR_CAPACITY
    }
  }
}

// ---- pop ------------------------------------------------------------------------------------------
constexpr int kSortBins = 4096;
struct SortScratch {                     // lives in the pop scratch buffer
  unsigned long long min64;            // smallest / largest popped composite (set by the partition pass)
  unsigned long long max64;
  uint32_t counts[kSortBins + 1];      // bucket sizes, then (after the scan) bucket start offsets
  uint32_t cursor[kSortBins];
};

// where instance blockIdx.y keeps its things
struct Pop3 { unsigned long long hl; uint32_t id, pad; };      // one popped entry: (key << 32 | key_lo, id)
__device__ __forceinline__ bool pop3_less(const Pop3 &a, const Pop3 &b) { return a.hl < b.hl || (a.hl == b.hl && a.id < b.id); }

struct Seg {
  OpenState *states;
  uint32_t *key, *key_lo, *id;   // key_lo may be NULL (narrow keys)
  uint64_t seg_cap;              // OPEN entries per instance
  uint8_t *scratch;
  uint64_t scratch_stride;       // bytes of pop scratch per instance
  uint32_t *popped_ids;          // [n_inst][popped_stride]
  uint32_t popped_stride;
  int32_t batch;
};
struct Carve {                   // one instance's view
  OpenState *s;
  uint32_t *key, *key_lo, *id;
  uint32_t *hist;
  SortScratch *ss;
  unsigned long long *cand;      // boundary-bucket candidates: (key << 32 | key_lo) ...
  uint32_t *cand_id;             // ... and their ids
  Pop3 *popped, *sorted, *tmp;
  uint32_t *holes, *surv, *popped_ids;
};
__host__ __device__ inline uint64_t sort_scratch_bytes() { return (sizeof(SortScratch) + 15) / 16 * 16; }
// per-instance scratch layout (bytes): hist 2*4096*4 | sort scratch | cand 8*cap | cand_id 4*cap | popped 16*batch | sorted 16*batch | tmp 16*batch
// | holes 4*batch | surv 4*batch
__host__ __device__ inline uint64_t pop_scratch_stride(uint64_t seg_cap, uint64_t batch) {
  return (2 * kBins * 4 + sort_scratch_bytes() + 12 * ((seg_cap + 3) / 4 * 4) + 56 * batch + 255) / 256 * 256;
}
__device__ __forceinline__ Carve carve(const Seg &g) {
  const uint32_t inst = blockIdx.y;
  Carve c;
  c.s = g.states + inst;
  c.key = g.key + inst * g.seg_cap;
  c.key_lo = g.key_lo ? g.key_lo + inst * g.seg_cap : nullptr;
  c.id = g.id + inst * g.seg_cap;
  const uint64_t cap4 = (g.seg_cap + 3) / 4 * 4;
  uint8_t *p = g.scratch + inst * g.scratch_stride;
  c.hist = reinterpret_cast<uint32_t *>(p); p += 2 * kBins * 4;
  c.ss = reinterpret_cast<SortScratch *>(p); p += sort_scratch_bytes();
  c.cand = reinterpret_cast<unsigned long long *>(p); p += 8 * cap4;
  c.cand_id = reinterpret_cast<uint32_t *>(p); p += 4 * cap4;
  c.popped = reinterpret_cast<Pop3 *>(p); p += 16 * (uint64_t)g.batch;
  c.sorted = reinterpret_cast<Pop3 *>(p); p += 16 * (uint64_t)g.batch;
  c.tmp = reinterpret_cast<Pop3 *>(p); p += 16 * (uint64_t)g.batch;
  c.holes = reinterpret_cast<uint32_t *>(p); p += 4 * (uint64_t)g.batch;
  c.surv = reinterpret_cast<uint32_t *>(p);
  c.popped_ids = g.popped_ids + (uint64_t)inst * g.popped_stride;
  return c;
}

// mode: 0 = C++ semantics, 1 = Python semantics (instances that found their goal rest unless include_solved), -1 = stand-alone
// queue (dcb_open_pop: C++ goal rule when asked, never rests)
__global__ void open_pop_begin_kernel(Seg g, int mode, int include_solved, const dcb_step_plan *plan) {
  const Carve c = carve(g);
  OpenState *s = c.s;
  for (int i = threadIdx.x; i < 2 * kBins; i += blockDim.x) c.hist[i] = 0;
  for (int i = threadIdx.x; i <= kSortBins; i += blockDim.x) c.ss->counts[i] = 0;
  for (int i = threadIdx.x; i < kSortBins; i += blockDim.x) c.ss->cursor[i] = 0;
  if (threadIdx.x == 0) { c.ss->min64 = ~0ull; c.ss->max64 = 0ull; }
  if (threadIdx.x == 0) {
    // an instance rests once it is done (C++: the loop test at :169; Python: astar.py:263-265 skips instances with a goal node)
    const bool rest = (mode >= 0 && s->done != 0 && !(mode == 1 && include_solved && s->done == 1)) || (plan && plan->budget == 0);
    const uint32_t n = s->open_size;
    const uint32_t b = rest ? 0u : (n < (uint32_t)g.batch ? n : (uint32_t)g.batch);
    s->resting = rest ? 1u : 0u;
    s->n_at_pop = n;
    s->n_take = b;
    s->need = b;                 // select the b smallest
    s->take_all = (b == n || b == 0);
    s->prefix = 0;
    s->cand_count = 0;
    s->n_holes = 0;
    s->n_surv = 0;
    s->n_popped = 0;
    s->n_expand = 0;
    s->thr_key = kNone;
    s->thr_lo = kNone;
    s->thr_id = kNone;
  }
}

// LEVEL 0: bucket = key >> 20 over all entries.  LEVEL 1: bucket = (key >> 8) & 0xFFF over entries whose
// top 12 bits equal the level-0 boundary bucket.
template <int LEVEL>
__global__ void __launch_bounds__(512) open_hist_kernel(Seg g) {
  const Carve c = carve(g);
  const OpenState *s = c.s;
  if (s->take_all) return;
  __shared__ uint32_t sh[kBins];
  for (int i = threadIdx.x; i < kBins; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const uint32_t n = s->n_at_pop, prefix = s->prefix;
  const uint32_t *__restrict__ key = c.key;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t k = key[i];
    if (LEVEL == 0) atomicAdd(&sh[k >> 20], 1u);
    else if ((k >> 20) == prefix) atomicAdd(&sh[(k >> 8) & 0xFFF], 1u);
  }
  __syncthreads();
  uint32_t *h = c.hist + LEVEL * kBins;
  for (int i = threadIdx.x; i < kBins; i += blockDim.x)
    if (sh[i]) atomicAdd(&h[i], sh[i]);
}

// Single block: find the bucket in which the cumulative count reaches `need`.
template <int LEVEL> __global__ void __launch_bounds__(1024) open_scan_kernel(Seg g) {
  const Carve c = carve(g);
  OpenState *s = c.s;
  if (s->take_all) return;
  __shared__ uint32_t part[1024];
  const uint32_t *h = c.hist + LEVEL * kBins;
  const int t = threadIdx.x;
  const uint32_t need = s->need, prefix_in = s->prefix;   // read before anyone writes them back
  uint32_t cnt[4], sum = 0;
#pragma unroll
  for (int q = 0; q < 4; q++) { cnt[q] = h[4 * t + q]; sum += cnt[q]; }
  part[t] = sum;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {   // Hillis-Steele inclusive scan
    const uint32_t v = (t >= off) ? part[t - off] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  const uint32_t incl = part[t], excl = incl - sum;
  if (excl < need && need <= incl) {           // exactly one thread
    uint32_t below = excl;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      if (below + cnt[q] >= need) {
        s->prefix = (LEVEL == 0) ? (uint32_t)(4 * t + q) : ((prefix_in << 12) | (uint32_t)(4 * t + q));
        s->need = need - below;                // still to take from inside this bucket
        break;
      }
      below += cnt[q];
    }
  }
}

// Collect the composite keys of the boundary bucket (top 24 key bits == prefix).
__global__ void __launch_bounds__(512) open_collect_kernel(Seg g) {
  const Carve c = carve(g);
  OpenState *s = c.s;
  if (s->take_all) return;
  const uint32_t n = s->n_at_pop, prefix = s->prefix;
  const uint32_t *__restrict__ key = c.key, *__restrict__ id = c.id;
  for (uint32_t i0 = blockIdx.x * blockDim.x; i0 < n; i0 += gridDim.x * blockDim.x) {
    const uint32_t i = i0 + threadIdx.x;
    const bool hit = i < n && (key[i] >> 8) == prefix;
    const uint32_t pos = warp_agg_inc(&s->cand_count, hit);
    if (hit) {
      c.cand[pos] = ((unsigned long long)key[i] << 32) | (c.key_lo ? c.key_lo[i] : 0u);
      c.cand_id[pos] = id[i];
    }
  }
}

// Single block: the `need`-th smallest of the candidates, by 8-bit radix passes over the 72 bits still open: the low 8 bits of the
// key, key_lo, id (the top 24 key bits are the bucket prefix).  A candidate is the 96-bit value (key, key_lo, id).
__global__ void __launch_bounds__(1024) open_select_finish_kernel(Seg g) {
  const Carve c = carve(g);
  OpenState *s = c.s;
  if (s->take_all) return;
  __shared__ uint32_t hist[256];
  __shared__ unsigned __int128 sel_prefix;
  __shared__ uint32_t sel_need;
  const uint32_t nc = s->cand_count;
  const unsigned long long *cand = c.cand;
  const uint32_t *cand_id = c.cand_id;
  const int n_pass = c.key_lo ? 9 : 5;               // narrow keys: key_lo is 0 everywhere, its four digits need no pass
  if (threadIdx.x == 0) { sel_prefix = (unsigned __int128)s->prefix; sel_need = s->need; }   // 24 bits known
  __syncthreads();
  for (int pass = 0; pass < n_pass; pass++) {
    // digit positions inside the 96-bit value: pass 0 = bits 71..64; then (wide) 63..32 key_lo; then 31..0 id
    const int shift = c.key_lo ? 64 - 8 * pass : (pass == 0 ? 64 : 32 - 8 * pass);
    if (threadIdx.x < 256) hist[threadIdx.x] = 0;
    __syncthreads();
    const unsigned __int128 pre = sel_prefix;
    const int pre_shift = (!c.key_lo && pass == 1) ? 64 : shift + 8;      // narrow: skip over the (all-zero) key_lo digits
    for (uint32_t i = threadIdx.x; i < nc; i += blockDim.x) {
      const unsigned __int128 v = ((unsigned __int128)cand[i] << 32) | cand_id[i];
      if ((v >> pre_shift) == pre) atomicAdd(&hist[(uint32_t)(v >> shift) & 0xFF], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t need = sel_need, below = 0;
      int b = 0;
      for (; b < 256; b++) {
        if (below + hist[b] >= need) break;
        below += hist[b];
      }
      unsigned __int128 p2 = pre;
      if (!c.key_lo && pass == 1) p2 <<= 32;                                // the key_lo digits of a narrow key: zeros
      sel_prefix = (p2 << 8) | (unsigned __int128)b;
      sel_need = need - below;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const unsigned __int128 t = sel_prefix;          // the full 96-bit threshold
    s->thr_key = (uint32_t)(t >> 64);
    s->thr_lo = (uint32_t)(t >> 32);
    s->thr_id = (uint32_t)t;
  }
}

// Partition: pop everything <= threshold; remember holes in the kept prefix and survivors in the tail.
__global__ void __launch_bounds__(512) open_partition_kernel(Seg g) {
  const Carve c = carve(g);
  OpenState *s = c.s;
  const uint32_t b = s->n_take;
  if (b == 0) return;
  const uint32_t n = s->n_at_pop;
  const uint32_t new_size = n - b;
  const uint32_t tk = s->thr_key, tl = s->thr_lo, ti = s->thr_id;
  const uint32_t *__restrict__ key = c.key, *__restrict__ id = c.id;
  for (uint32_t i0 = blockIdx.x * blockDim.x; i0 < n; i0 += gridDim.x * blockDim.x) {
    const uint32_t i = i0 + threadIdx.x;
    bool pop = false, hole = false, sv = false;
    uint32_t k = 0, l = 0, d = 0;
    if (i < n) {
      k = key[i];
      if (k <= tk) {
        d = id[i];
        l = c.key_lo ? c.key_lo[i] : 0u;
        pop = (k < tk) || (l < tl) || (l == tl && d <= ti);
      }
      hole = pop && i < new_size;
      sv = !pop && i >= new_size;
    }
    const uint32_t pp = warp_agg_inc(&s->n_popped, pop);
    if (pop) {
      Pop3 v;
      v.hl = ((unsigned long long)k << 32) | l; v.id = d; v.pad = 0;
      c.popped[pp] = v;
      atomicMin(&c.ss->min64, v.hl);
      atomicMax(&c.ss->max64, v.hl);
    }
    const uint32_t hp = warp_agg_inc(&s->n_holes, hole);
    if (hole) c.holes[hp] = i;
    const uint32_t sp = warp_agg_inc(&s->n_surv, sv);
    if (sv) c.surv[sp] = i;
  }
}

__global__ void __launch_bounds__(256) open_fill_holes_kernel(Seg g) {
  const Carve c = carve(g);
  const uint32_t cnt = c.s->n_holes;   // == n_surv
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < cnt; j += gridDim.x * blockDim.x) {
    c.key[c.holes[j]] = c.key[c.surv[j]];
    if (c.key_lo) c.key_lo[c.holes[j]] = c.key_lo[c.surv[j]];
    c.id[c.holes[j]] = c.id[c.surv[j]];
  }
}

// Sorting the popped entries (all distinct) into cost order = the reference's pop order.  Bucket by a monotone map of the 64-bit
// key onto kSortBins equal slices of [min, max], counting-sort into bucket order, then rank only within a bucket (a handful of
// elements, compared as (key, key_lo, id)): O(b) instead of the O(b^2) plain rank sort (175 us at b = 20000).
__device__ __forceinline__ uint32_t sort_bucket(unsigned long long v, unsigned long long lo, double scale) {
  const double x = (double)(v - lo) * scale;                  // monotone in v
  const uint32_t b = (uint32_t)x;
  return b < (uint32_t)kSortBins ? b : (uint32_t)(kSortBins - 1);
}
__device__ __forceinline__ double sort_scale(const SortScratch *ss) {
  const unsigned long long hi = ss->max64, lo = ss->min64;
  return (hi >= lo) ? (double)kSortBins / ((double)(hi - lo) + 1.0) : 0.0;
}
__global__ void __launch_bounds__(256) open_sort_count_kernel(Seg g) {
  const Carve c = carve(g);
  const uint32_t b = c.s->n_popped;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < b; i += gridDim.x * blockDim.x)
    atomicAdd(&c.ss->counts[sort_bucket(c.popped[i].hl, c.ss->min64, sort_scale(c.ss))], 1u);
}
__global__ void __launch_bounds__(1024) open_sort_scan_kernel(Seg g) {      // exclusive scan of kSortBins counts
  const Carve c = carve(g);
  if (c.s->n_popped == 0) return;
  SortScratch *ss = c.ss;
  __shared__ uint32_t part[1024];
  const int t = threadIdx.x;
  uint32_t cnt[4], sum = 0;
#pragma unroll
  for (int q = 0; q < 4; q++) { cnt[q] = ss->counts[4 * t + q]; sum += cnt[q]; }
  part[t] = sum;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    const uint32_t v = (t >= off) ? part[t - off] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  uint32_t run = part[t] - sum;
#pragma unroll
  for (int q = 0; q < 4; q++) { ss->counts[4 * t + q] = run; run += cnt[q]; }
  if (t == 1023) ss->counts[kSortBins] = run;
}
__global__ void __launch_bounds__(256) open_sort_scatter_kernel(Seg g) {
  const Carve c = carve(g);
  const uint32_t b = c.s->n_popped;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < b; i += gridDim.x * blockDim.x) {
    const Pop3 v = c.popped[i];
    const uint32_t k = sort_bucket(v.hl, c.ss->min64, sort_scale(c.ss));
    c.tmp[c.ss->counts[k] + atomicAdd(&c.ss->cursor[k], 1u)] = v;
  }
}
__global__ void __launch_bounds__(256) open_sort_rank_kernel(Seg g) {
  const Carve c = carve(g);
  const uint32_t b = c.s->n_popped;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < b; i += gridDim.x * blockDim.x) {
    const Pop3 v = c.tmp[i];
    const uint32_t k = sort_bucket(v.hl, c.ss->min64, sort_scale(c.ss));
    const uint32_t lo = c.ss->counts[k], hi = c.ss->counts[k + 1];
    uint32_t rank = lo;
    for (uint32_t j = lo; j < hi; j++) rank += pop3_less(c.tmp[j], v);
    c.sorted[rank] = v;
  }
}

// Single block per instance: goal bookkeeping + termination and, in the C++ semantics, un-popping the entries behind the
// first solved pop.
//   mode 0, stop_at_goal: parallel_weighted_astar.cpp:186-208.   mode 0, !stop_at_goal: plain pop (no goal logic).
//   mode 1: search_methods/astar.py:69-76 -- every popped node stands; solved pops are goal nodes (:73); the caller's loop ends
//           once an instance has one (:421); the answer is the goal node of smallest path cost, first one on ties (:327-333).
__global__ void __launch_bounds__(1024)
open_finalize_kernel(Seg g, int mode, int stop_at_goal, int num_moves, const uint8_t *__restrict__ node_solved,
                     const uint32_t *__restrict__ node_g) {
  const Carve c = carve(g);
  OpenState *s = c.s;
  __shared__ uint32_t first_solved;
  __shared__ unsigned long long best_goal;       // python: (g << 32) | position in pop order
  __shared__ uint32_t goals_here;
  if (s->resting) return;
  const uint32_t b = s->n_popped;
  const uint32_t n = s->n_at_pop;
  const uint32_t new_size = n - b;
  const Pop3 *__restrict__ sorted = c.sorted;
  if (threadIdx.x == 0) { first_solved = kNone; best_goal = ~0ull; goals_here = 0; }
  __syncthreads();
  if (mode <= 0 && stop_at_goal && node_solved) {
    uint32_t best = kNone;
    for (uint32_t j = threadIdx.x; j < b; j += blockDim.x)
      if (node_solved[sorted[j].id]) { best = j; break; }   // j ascending per thread
    if (best != kNone) atomicMin(&first_solved, best);
  } else if (mode == 1) {
    for (uint32_t j = threadIdx.x; j < b; j += blockDim.x) {
      const uint32_t nid = sorted[j].id;
      if (node_solved[nid]) {
        atomicAdd(&goals_here, 1u);
        atomicMin(&best_goal, ((unsigned long long)node_g[nid] << 32) | j);
      }
    }
  }
  __syncthreads();
  const uint32_t fs = first_solved;
  const uint32_t m = (fs != kNone) ? fs + 1 : b;                    // pops that stand
  for (uint32_t j = threadIdx.x; j < b; j += blockDim.x) {
    const Pop3 v = sorted[j];
    if (j < m) c.popped_ids[j] = v.id;
    else {                                                         // back to OPEN
      c.key[new_size + (j - m)] = (uint32_t)(v.hl >> 32);
      if (c.key_lo) c.key_lo[new_size + (j - m)] = (uint32_t)v.hl;
      c.id[new_size + (j - m)] = v.id;
    }
  }
  if (threadIdx.x == 0) {
    uint32_t done = s->done;
    const uint32_t min_key = b ? (uint32_t)(sorted[0].hl >> 32) : kNone;
    if (mode <= 0) {
      const bool goal_prev = s->goal_id != kNone;
      if (fs != kNone) {
        const uint32_t gk = (uint32_t)(sorted[fs].hl >> 32), gi = sorted[fs].id;
        if (g.batch == 1) { s->goal_id = gi; s->goal_key = gk; done = 1; }            // :191-193
        else if (!goal_prev || s->goal_key > gk) { s->goal_id = gi; s->goal_key = gk; }  // :195-199
      }
      if (stop_at_goal && goal_prev && b && min_key >= s->goal_key) done = 1;         // :205-208
    } else if (goals_here) {
      const uint32_t gg = (uint32_t)(best_goal >> 32), pos = (uint32_t)best_goal;
      if (s->goal_id == kNone || gg < s->goal_key) { s->goal_id = sorted[pos].id; s->goal_key = gg; }
      s->n_goals += goals_here;
      done = 1;
    }
    if (b == 0) done = 2;                                                            // OPEN exhausted
    s->min_key = min_key;
    s->done = done;
    s->n_popped = m;
    s->open_size = new_size + (b - m);
    s->iterations += 1;
    s->nodes_generated += (uint64_t)m * (uint64_t)num_moves;     // :266 -- counted on the terminating iteration too; astar.py:168
    // C++: the terminating iteration's children are never needed (the loop test :169 fails first); Python expands every pop
    s->n_expand = (mode <= 0 && done) ? 0u : m;
  }
}
}  // namespace

// ---- host launchers --------------------------------------------------------------------------------
int open_clear_device(void *state, int n_inst, cudaStream_t st) {
  open_clear_kernel<<<(unsigned)n_inst, 32, 0, st>>>(reinterpret_cast<OpenState *>(state));
  return dcb_check_launch();
}

int open_push_device(void *state, uint32_t *key, uint32_t *key_lo, uint32_t *id, int64_t capacity, const float *cost, const uint32_t *ids,
                     uint32_t first_id, const uint8_t *keep, int64_t m, cudaStream_t st) {
  if (m == 0) return DCB_OK;
  open_push_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(reinterpret_cast<OpenState *>(state), key, key_lo, id, (uint32_t)capacity, cost,
                                                               ids, first_id, keep, m);
  return dcb_check_launch();
}

int64_t open_scratch_bytes(int64_t capacity, int64_t batch, int64_t n_inst) {
  return (int64_t)pop_scratch_stride((uint64_t)capacity, (uint64_t)batch) * n_inst + 256;
}

// Pop for n_inst instances at once.  seg_cap = OPEN entries per instance, popped ids of instance i at popped_ids + i*popped_stride.
int open_pop_device(void *state, uint32_t *key, uint32_t *key_lo, uint32_t *id, int64_t seg_cap, int n_inst, int32_t batch, int mode, int stop_at_goal,
                    int include_solved, int num_moves, const uint8_t *node_solved, const uint32_t *node_g, uint32_t *popped_ids,
                    int64_t popped_stride, void *scratch, const dcb_step_plan *plan, cudaStream_t st) {
  Seg g;
  g.states = reinterpret_cast<OpenState *>(state);
  g.key = key; g.key_lo = key_lo; g.id = id;
  g.seg_cap = (uint64_t)seg_cap;
  g.scratch = reinterpret_cast<uint8_t *>(scratch);
  g.scratch_stride = pop_scratch_stride((uint64_t)seg_cap, (uint64_t)batch);
  g.popped_ids = popped_ids;
  g.popped_stride = (uint32_t)popped_stride;
  g.batch = batch;
  // grid-stride passes over a whole segment: all SMs for one instance, fewer blocks each as the instance count grows
  int full_blocks = (148 * 4 + n_inst - 1) / n_inst;
  const int64_t seg_blocks = (seg_cap + 511) / 512;
  if (full_blocks > seg_blocks) full_blocks = (int)seg_blocks;
  if (full_blocks < 1) full_blocks = 1;
  int batch_blocks = (batch + 255) / 256;
  if (batch_blocks > full_blocks * 2) batch_blocks = full_blocks * 2;
  const dim3 one(1, n_inst), full(full_blocks, n_inst), bat(batch_blocks, n_inst);
  open_pop_begin_kernel<<<one, 1024, 0, st>>>(g, mode, include_solved, plan);
  open_hist_kernel<0><<<full, 512, 0, st>>>(g);
  open_scan_kernel<0><<<one, 1024, 0, st>>>(g);
  open_hist_kernel<1><<<full, 512, 0, st>>>(g);
  open_scan_kernel<1><<<one, 1024, 0, st>>>(g);
  open_collect_kernel<<<full, 512, 0, st>>>(g);
  open_select_finish_kernel<<<one, 1024, 0, st>>>(g);
  open_partition_kernel<<<full, 512, 0, st>>>(g);
  open_fill_holes_kernel<<<bat, 256, 0, st>>>(g);
  open_sort_count_kernel<<<bat, 256, 0, st>>>(g);
  open_sort_scan_kernel<<<one, 1024, 0, st>>>(g);
  open_sort_scatter_kernel<<<bat, 256, 0, st>>>(g);
  open_sort_rank_kernel<<<bat, 256, 0, st>>>(g);
  open_finalize_kernel<<<one, 1024, 0, st>>>(g, mode, stop_at_goal, num_moves, node_solved, node_g);
  return dcb_check_launch();
}

}  // namespace dcb
