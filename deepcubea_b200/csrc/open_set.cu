// open_set.cu -- OPEN as a bucket priority queue in HBM: exact "pop the B cheapest" without a heap.
//
// Replaces std::priority_queue<Node*,std::vector<Node*>,compareNodeCost> open, the one-at-a-time pop
// loop (with its `break` at the first solved node) and the push loop of
// cpp/parallel_weighted_astar.cpp:141, 177-208, 309-319; heapq open_set / pop_from_open / push_to_open of
// search_methods/astar.py:53, 64-76.
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
//      deterministic) and the goal / termination rule of :190-208 is applied on the device; entries behind
//      the first solved pop go back to OPEN, exactly like the reference's `break`.
// HBM traffic per pop: 4 passes x 4 B per open entry; per push: 8 B per entry.
#include <cuda_runtime.h>
#include "dcb_internal.h"

namespace dcb {

namespace {
constexpr int kBins = 4096;
constexpr uint32_t kNone = 0xFFFFFFFFu;

struct OpenState {           // mirrors dcb_open_state (include/dcb.h)
  uint32_t size, n_popped, thr_key, thr_id, min_key, goal_id, goal_key, done;
  uint32_t overflow, need, prefix, cand_count, n_holes, n_surv, take_all, n_at_pop;
};
static_assert(sizeof(OpenState) == sizeof(dcb_open_state), "state layout");

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
    *s = z;
  }
}

__global__ void __launch_bounds__(256)
open_push_kernel(OpenState *s, uint32_t *key, uint32_t *id, uint32_t capacity, const float *cost, const uint32_t *ids,
                 uint32_t first_id, const uint8_t *keep, int64_t m) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const bool want = i < m && (!keep || keep[i]);
  const uint32_t pos = warp_agg_inc(&s->size, want);
  if (want) {
    if (pos < capacity) {
      key[pos] = __float_as_uint(cost[i]);
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

__global__ void open_pop_begin_kernel(OpenState *s, uint32_t *hist, int32_t batch, SortScratch *ss) {
  for (int i = threadIdx.x; i < 2 * kBins; i += blockDim.x) hist[i] = 0;
  for (int i = threadIdx.x; i <= kSortBins; i += blockDim.x) ss->counts[i] = 0;
  for (int i = threadIdx.x; i < kSortBins; i += blockDim.x) ss->cursor[i] = 0;
  if (threadIdx.x == 0) { ss->min64 = ~0ull; ss->max64 = 0ull; }
  if (threadIdx.x == 0) {
    const uint32_t n = s->size;
    const uint32_t b = n < (uint32_t)batch ? n : (uint32_t)batch;
    s->n_at_pop = n;
    s->need = b;                 // select the b smallest
    s->take_all = (b == n);
    s->prefix = 0;
    s->cand_count = 0;
    s->n_holes = 0;
    s->n_surv = 0;
    s->n_popped = 0;
    s->thr_key = kNone;
    s->thr_id = kNone;
  }
}

// LEVEL 0: bucket = key >> 20 over all entries.  LEVEL 1: bucket = (key >> 8) & 0xFFF over entries whose
// top 12 bits equal the level-0 boundary bucket.
template <int LEVEL>
__global__ void __launch_bounds__(512) open_hist_kernel(const OpenState *s, const uint32_t *__restrict__ key, uint32_t *hist) {
  if (s->take_all) return;
  __shared__ uint32_t sh[kBins];
  for (int i = threadIdx.x; i < kBins; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const uint32_t n = s->n_at_pop, prefix = s->prefix;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t k = key[i];
    if (LEVEL == 0) atomicAdd(&sh[k >> 20], 1u);
    else if ((k >> 20) == prefix) atomicAdd(&sh[(k >> 8) & 0xFFF], 1u);
  }
  __syncthreads();
  uint32_t *h = hist + LEVEL * kBins;
  for (int i = threadIdx.x; i < kBins; i += blockDim.x)
    if (sh[i]) atomicAdd(&h[i], sh[i]);
}

// Single block: find the bucket in which the cumulative count reaches `need`.
template <int LEVEL> __global__ void __launch_bounds__(1024) open_scan_kernel(OpenState *s, const uint32_t *hist) {
  if (s->take_all) return;
  __shared__ uint32_t part[1024];
  const uint32_t *h = hist + LEVEL * kBins;
  const int t = threadIdx.x;
  const uint32_t need = s->need, prefix_in = s->prefix;   // read before anyone writes them back
  uint32_t c[4], sum = 0;
#pragma unroll
  for (int q = 0; q < 4; q++) { c[q] = h[4 * t + q]; sum += c[q]; }
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
      if (below + c[q] >= need) {
        s->prefix = (LEVEL == 0) ? (uint32_t)(4 * t + q) : ((prefix_in << 12) | (uint32_t)(4 * t + q));
        s->need = need - below;                // still to take from inside this bucket
        break;
      }
      below += c[q];
    }
  }
}

// Collect the composite keys of the boundary bucket (top 24 key bits == prefix).
__global__ void __launch_bounds__(512)
open_collect_kernel(OpenState *s, const uint32_t *__restrict__ key, const uint32_t *__restrict__ id, unsigned long long *cand) {
  if (s->take_all) return;
  const uint32_t n = s->n_at_pop, prefix = s->prefix;
  for (uint32_t i0 = blockIdx.x * blockDim.x; i0 < n; i0 += gridDim.x * blockDim.x) {
    const uint32_t i = i0 + threadIdx.x;
    const bool hit = i < n && (key[i] >> 8) == prefix;
    const uint32_t pos = warp_agg_inc(&s->cand_count, hit);
    if (hit) cand[pos] = ((unsigned long long)key[i] << 32) | id[i];
  }
}

// Single block: the `need`-th smallest of the candidates, by 8-bit radix passes over their low 40 bits.
__global__ void __launch_bounds__(1024) open_select_finish_kernel(OpenState *s, const unsigned long long *cand) {
  if (s->take_all) return;
  __shared__ uint32_t hist[256];
  __shared__ unsigned long long sel_prefix;
  __shared__ uint32_t sel_need;
  const uint32_t c = s->cand_count;
  if (threadIdx.x == 0) { sel_prefix = (unsigned long long)s->prefix; sel_need = s->need; }   // 24 bits known
  __syncthreads();
  for (int pass = 0; pass < 5; pass++) {
    const int shift = 32 - 8 * pass;            // digit = bits [shift+7 .. shift] of the 64-bit composite
    if (threadIdx.x < 256) hist[threadIdx.x] = 0;
    __syncthreads();
    const unsigned long long pre = sel_prefix;
    for (uint32_t i = threadIdx.x; i < c; i += blockDim.x) {
      const unsigned long long v = cand[i];
      if ((v >> (shift + 8)) == pre) atomicAdd(&hist[(uint32_t)(v >> shift) & 0xFF], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t need = sel_need, below = 0;
      int b = 0;
      for (; b < 256; b++) {
        if (below + hist[b] >= need) break;
        below += hist[b];
      }
      sel_prefix = (pre << 8) | (unsigned long long)b;
      sel_need = need - below;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    s->thr_key = (uint32_t)(sel_prefix >> 32);
    s->thr_id = (uint32_t)sel_prefix;
  }
}

// Partition: pop everything <= threshold; remember holes in the kept prefix and survivors in the tail.
__global__ void __launch_bounds__(512)
open_partition_kernel(OpenState *s, const uint32_t *__restrict__ key, const uint32_t *__restrict__ id, int32_t batch,
                      unsigned long long *popped, uint32_t *holes, uint32_t *surv, SortScratch *ss) {
  const uint32_t n = s->n_at_pop;
  const uint32_t b = n < (uint32_t)batch ? n : (uint32_t)batch;
  const uint32_t new_size = n - b;
  const uint32_t tk = s->thr_key, ti = s->thr_id;
  for (uint32_t i0 = blockIdx.x * blockDim.x; i0 < n; i0 += gridDim.x * blockDim.x) {
    const uint32_t i = i0 + threadIdx.x;
    bool pop = false, hole = false, sv = false;
    uint32_t k = 0, d = 0;
    if (i < n) {
      k = key[i];
      if (k <= tk) {
        d = id[i];
        pop = (k < tk) || (d <= ti);
      }
      hole = pop && i < new_size;
      sv = !pop && i >= new_size;
    }
    const uint32_t pp = warp_agg_inc(&s->n_popped, pop);
    if (pop) {
      const unsigned long long v = ((unsigned long long)k << 32) | d;
      popped[pp] = v;
      atomicMin(&ss->min64, v);
      atomicMax(&ss->max64, v);
    }
    const uint32_t hp = warp_agg_inc(&s->n_holes, hole);
    if (hole) holes[hp] = i;
    const uint32_t sp = warp_agg_inc(&s->n_surv, sv);
    if (sv) surv[sp] = i;
  }
}

__global__ void __launch_bounds__(256) open_fill_holes_kernel(const OpenState *s, uint32_t *key, uint32_t *id, const uint32_t *holes, const uint32_t *surv) {
  const uint32_t cnt = s->n_holes;   // == n_surv
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < cnt; j += gridDim.x * blockDim.x) {
    key[holes[j]] = key[surv[j]];
    id[holes[j]] = id[surv[j]];
  }
}

// Sorting the popped composites (all distinct) into cost order = the reference's pop order.  Bucket by a monotone map of
// the value onto kSortBins equal slices of [min, threshold], counting-sort into bucket order, then rank only within a
// bucket (a handful of elements): O(b) instead of the O(b^2) plain rank sort (175 us at b = 20000).
__device__ __forceinline__ uint32_t sort_bucket(unsigned long long v, unsigned long long lo, double scale) {
  const double x = (double)(v - lo) * scale;                  // monotone in v
  const uint32_t b = (uint32_t)x;
  return b < (uint32_t)kSortBins ? b : (uint32_t)(kSortBins - 1);
}
__device__ __forceinline__ double sort_scale(const OpenState *, const SortScratch *ss) {
  const unsigned long long hi = ss->max64, lo = ss->min64;
  return (hi >= lo) ? (double)kSortBins / ((double)(hi - lo) + 1.0) : 0.0;
}
__global__ void __launch_bounds__(256) open_sort_count_kernel(const OpenState *s, const unsigned long long *__restrict__ popped, SortScratch *ss) {
  const uint32_t b = s->n_popped, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b) return;
  atomicAdd(&ss->counts[sort_bucket(popped[i], ss->min64, sort_scale(s, ss))], 1u);
}
__global__ void __launch_bounds__(1024) open_sort_scan_kernel(SortScratch *ss) {      // exclusive scan of kSortBins counts
  __shared__ uint32_t part[1024];
  const int t = threadIdx.x;
  uint32_t c[4], sum = 0;
#pragma unroll
  for (int q = 0; q < 4; q++) { c[q] = ss->counts[4 * t + q]; sum += c[q]; }
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
  for (int q = 0; q < 4; q++) { ss->counts[4 * t + q] = run; run += c[q]; }
  if (t == 1023) ss->counts[kSortBins] = run;
}
__global__ void __launch_bounds__(256)
open_sort_scatter_kernel(const OpenState *s, const unsigned long long *__restrict__ popped, SortScratch *ss, unsigned long long *tmp) {
  const uint32_t b = s->n_popped, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b) return;
  const unsigned long long v = popped[i];
  const uint32_t k = sort_bucket(v, ss->min64, sort_scale(s, ss));
  tmp[ss->counts[k] + atomicAdd(&ss->cursor[k], 1u)] = v;
}
__global__ void __launch_bounds__(256)
open_sort_rank_kernel(const OpenState *s, const SortScratch *ss, const unsigned long long *__restrict__ tmp, unsigned long long *sorted) {
  const uint32_t b = s->n_popped, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b) return;
  const unsigned long long v = tmp[i];
  const uint32_t k = sort_bucket(v, ss->min64, sort_scale(s, ss));
  const uint32_t lo = ss->counts[k], hi = ss->counts[k + 1];
  uint32_t rank = lo;
  for (uint32_t j = lo; j < hi; j++) rank += tmp[j] < v;
  sorted[rank] = v;
}

// Single block: goal bookkeeping + termination (parallel_weighted_astar.cpp:186-208) and un-popping the
// entries behind the first solved pop.
__global__ void __launch_bounds__(1024)
open_finalize_kernel(OpenState *s, uint32_t *key, uint32_t *id, int32_t batch, int stop_at_goal,
                     const uint8_t *__restrict__ node_solved, const unsigned long long *__restrict__ sorted, uint32_t *popped_ids) {
  __shared__ uint32_t first_solved;
  const uint32_t b = s->n_popped;
  const uint32_t n = s->n_at_pop;
  const uint32_t new_size = n - b;
  if (threadIdx.x == 0) first_solved = kNone;
  __syncthreads();
  if (stop_at_goal && node_solved) {
    uint32_t best = kNone;
    for (uint32_t j = threadIdx.x; j < b; j += blockDim.x)
      if (node_solved[(uint32_t)sorted[j]]) { best = j; break; }   // j ascending per thread
    if (best != kNone) atomicMin(&first_solved, best);
  }
  __syncthreads();
  const uint32_t fs = first_solved;
  const uint32_t m = (fs != kNone) ? fs + 1 : b;                    // pops that stand
  for (uint32_t j = threadIdx.x; j < b; j += blockDim.x) {
    const unsigned long long v = sorted[j];
    if (j < m) popped_ids[j] = (uint32_t)v;
    else {                                                         // back to OPEN
      key[new_size + (j - m)] = (uint32_t)(v >> 32);
      id[new_size + (j - m)] = (uint32_t)v;
    }
  }
  if (threadIdx.x == 0) {
    const bool goal_prev = s->goal_id != kNone;
    uint32_t done = s->done;
    if (fs != kNone) {
      const uint32_t gk = (uint32_t)(sorted[fs] >> 32), gi = (uint32_t)sorted[fs];
      if (batch == 1) { s->goal_id = gi; s->goal_key = gk; done = 1; }            // :191-193
      else if (!goal_prev || s->goal_key > gk) { s->goal_id = gi; s->goal_key = gk; }  // :195-199
    }
    const uint32_t min_key = b ? (uint32_t)(sorted[0] >> 32) : kNone;
    if (goal_prev && b && min_key >= s->goal_key) done = 1;                         // :205-208
    if (b == 0) done = 2;                                                            // OPEN exhausted
    s->min_key = min_key;
    s->done = done;
    s->n_popped = m;
    s->size = new_size + (b - m);
  }
}
}  // namespace

// ---- host launchers --------------------------------------------------------------------------------
int open_clear_device(void *state, cudaStream_t st) {
  open_clear_kernel<<<1, 32, 0, st>>>(reinterpret_cast<OpenState *>(state));
  return dcb_check_launch();
}

int open_push_device(void *state, uint32_t *key, uint32_t *id, int64_t capacity, const float *cost, const uint32_t *ids,
                     uint32_t first_id, const uint8_t *keep, int64_t m, cudaStream_t st) {
  if (m == 0) return DCB_OK;
  open_push_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(reinterpret_cast<OpenState *>(state), key, id, (uint32_t)capacity, cost,
                                                               ids, first_id, keep, m);
  return dcb_check_launch();
}

// scratch layout (bytes): hist 2*4096*4 | sort scratch | cand 8*cap | popped 8*batch | sorted 8*batch | tmp 8*batch | holes 4*batch | surv 4*batch
int64_t open_scratch_bytes(int64_t capacity, int64_t batch) {
  return 2 * kBins * 4 + (int64_t)((sizeof(SortScratch) + 15) / 16 * 16) + 8 * capacity + 32 * batch + 256;
}

int open_pop_device(void *state, uint32_t *key, uint32_t *id, int64_t capacity, int32_t batch, int stop_at_goal,
                    const uint8_t *node_solved, uint32_t *popped_ids, void *scratch, cudaStream_t st) {
  OpenState *s = reinterpret_cast<OpenState *>(state);
  uint8_t *p = reinterpret_cast<uint8_t *>(scratch);
  uint32_t *hist = reinterpret_cast<uint32_t *>(p); p += 2 * kBins * 4;
  SortScratch *ss = reinterpret_cast<SortScratch *>(p); p += (sizeof(SortScratch) + 15) / 16 * 16;
  unsigned long long *cand = reinterpret_cast<unsigned long long *>(p); p += 8 * capacity;
  unsigned long long *popped = reinterpret_cast<unsigned long long *>(p); p += 8 * (int64_t)batch;
  unsigned long long *sorted = reinterpret_cast<unsigned long long *>(p); p += 8 * (int64_t)batch;
  unsigned long long *tmp = reinterpret_cast<unsigned long long *>(p); p += 8 * (int64_t)batch;
  uint32_t *holes = reinterpret_cast<uint32_t *>(p); p += 4 * (int64_t)batch;
  uint32_t *surv = reinterpret_cast<uint32_t *>(p);
  const int full_blocks = 148 * 4;   // grid-stride passes over the whole array
  open_pop_begin_kernel<<<1, 1024, 0, st>>>(s, hist, batch, ss);
  open_hist_kernel<0><<<full_blocks, 512, 0, st>>>(s, key, hist);
  open_scan_kernel<0><<<1, 1024, 0, st>>>(s, hist);
  open_hist_kernel<1><<<full_blocks, 512, 0, st>>>(s, key, hist);
  open_scan_kernel<1><<<1, 1024, 0, st>>>(s, hist);
  open_collect_kernel<<<full_blocks, 512, 0, st>>>(s, key, id, cand);
  open_select_finish_kernel<<<1, 1024, 0, st>>>(s, cand);
  open_partition_kernel<<<full_blocks, 512, 0, st>>>(s, key, id, batch, popped, holes, surv, ss);
  open_fill_holes_kernel<<<(batch + 255) / 256, 256, 0, st>>>(s, key, id, holes, surv);
  const unsigned sort_blocks = (unsigned)((batch + 255) / 256);
  open_sort_count_kernel<<<sort_blocks, 256, 0, st>>>(s, popped, ss);
  open_sort_scan_kernel<<<1, 1024, 0, st>>>(ss);
  open_sort_scatter_kernel<<<sort_blocks, 256, 0, st>>>(s, popped, ss, tmp);
  open_sort_rank_kernel<<<sort_blocks, 256, 0, st>>>(s, ss, tmp, sorted);
  open_finalize_kernel<<<1, 1024, 0, st>>>(s, key, id, batch, stop_at_goal, node_solved, sorted, popped_ids);
  return dcb_check_launch();
}

}  // namespace dcb
