// expand_kernels.cu -- the batched environment step on sm_100a.
//
// expand_kernel<ENV, INDEXED>: one lane owns one parent.  The parent's S bytes are pulled into registers
// with aligned 32-bit loads (+ funnel shift for the env's natural misalignment), all A children are
// produced in registers (cube3: PRMT networks; n-puzzle: SIMD-within-register blank swap), hashed
// (IMAD.WIDE pair products) and solved-checked, and the lane's A*S-byte child record is written to a
// per-warp shared-memory staging tile.  A 32-parent tile of records is contiguous in the output, so one
// elected lane ships it with a single TMA bulk store (cp.async.bulk shared->global); the warp only
// waits for the TMA engine to finish READING the tile before refilling it.  HBM traffic is therefore
// the algorithmic minimum: S bytes read per parent, A*S + A + 8A bytes written.
//
// Replaces Cube3::getNextStates/isSolved, PuzzleN::getNextStates/isSolved (cpp/environments.cpp:92-126,
// 222-256), the OpenMP loop of parallel_weighted_astar.cpp:217-230, and Environment.expand/_move_np/
// is_solved of environments/cube3.py:71-75,129-171 and environments/n_puzzle.py:78-82,136-231.
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>
#include "dcb_internal.h"
#include "expand_core.cuh"
#include "ptx.cuh"

namespace dcb {


// ---- vector stores of a run of record words into the lane's staging slot --------------------------
template <int K0, int I, int N, int ALIGNW> struct StoreWords {
  static __device__ __forceinline__ void run(uint32_t *base, const uint32_t (&r)[N]) {
    if constexpr (I < N) {
      constexpr int k = K0 + I;
      if constexpr (ALIGNW >= 4 && k % 4 == 0 && I + 4 <= N) {
        sts128(base + k, r[I], r[I + 1], r[I + 2], r[I + 3]);
        StoreWords<K0, I + 4, N, ALIGNW>::run(base, r);
      } else if constexpr (ALIGNW >= 2 && k % 2 == 0 && I + 2 <= N) {
        sts64(base + k, r[I], r[I + 1]);
        StoreWords<K0, I + 2, N, ALIGNW>::run(base, r);
      } else {
        sts32(base + k, r[I]);
        StoreWords<K0, I + 1, N, ALIGNW>::run(base, r);
      }
    }
  }
};

template <int ENV> struct SmemSink {
  using Sh = ExpandShape<ENV>;
  // widest store the lane stride (REC_WORDS) keeps aligned; cube3: 162 words -> 8-byte stores, which are
  // bank-conflict free at that stride (162 = 2 mod 32: a half-warp's 64-bit stores tile all 32 banks)
  static constexpr int ALIGNW = (Sh::REC_WORDS % 4 == 0) ? 4 : ((Sh::REC_WORDS % 2 == 0) ? 2 : 1);
  uint32_t *rec;
  uint64_t *hash_out;   // this parent's A hashes (16-byte aligned) or nullptr
  uint16_t *solved_out; // this parent's A flags as move pairs or nullptr
  uint64_t h_even;
  uint32_t s_even;

  template <int K0, int N> __device__ __forceinline__ void store_record_words(const uint32_t (&r)[N]) {
    StoreWords<K0, 0, N, ALIGNW>::run(rec, r);
  }
  template <int MOVE> __device__ __forceinline__ void store_hash(uint64_t h) {
    if constexpr (MOVE % 2 == 0) h_even = h;
    else if (hash_out) stg_cs_v2u64(hash_out + MOVE - 1, h_even, h);
  }
  template <int MOVE> __device__ __forceinline__ void store_solved(bool s) {
    if constexpr (MOVE % 2 == 0) s_even = s ? 1u : 0u;
    else if (solved_out) solved_out[MOVE / 2] = (uint16_t)(s_even | ((s ? 1u : 0u) << 8));
  }
};

template <int S> __device__ __forceinline__ void load_raw(const uint8_t *base, uint64_t byte_off, uint32_t (&raw)[LoadShape<S>::NRAW]) {
  const uint32_t *a = reinterpret_cast<const uint32_t *>(base + (byte_off & ~uint64_t(3)));
#pragma unroll
  for (int k = 0; k < LoadShape<S>::NRAW; k++) raw[k] = ldg_nc_u32(a + k);
}

__device__ __forceinline__ void team_sync(int split, int team) {
  if (split == 1) __syncwarp();
  else asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "r"(32 * split) : "memory");
}

// TEAMS staging tiles per CTA, each filled by SPLIT warps: lane l of every warp of a team owns parent l of the tile and
// warp `member` produces children [member*A/SPLIT, (member+1)*A/SPLIT).  Splitting the 12 moves over two warps halves
// the time a tile spends being computed, so a larger share of the SM's staging memory is in flight to HBM at any time.
//
// PAD > 0 (cube4): a record of 2304 bytes is a multiple of 128, so with records back to back every lane's staging stores
// would land in the same banks (32-way conflict).  The lanes' slots are then PAD bytes apart in shared memory (2320-byte
// stride: eight lanes' 16-byte stores tile all 32 banks) and every lane ships its own record with its own bulk store -- the
// tile is no longer one contiguous block, but 2304-byte bulk copies are still large enough for the TMA engine.
//
// PLANNED (the search loop, dcb_search_expand): the work is the iteration's TILE LIST in device memory (include/dcb.h) instead
// of a contiguous parent range -- tile t = {src, dst_slot, count, inst}: parents ids[src + lane], children to arena slots
// dst_slot + lane, solved flags to node_solved[(dst_slot + lane) * A ...], hashes to hash[(t * 32 + lane) * A ...]; the tile
// count is read from the plan, so the launch needs no host-side size.  The lane that owns a parent also writes its children's
// depth (parent depth + 1, parallel_weighted_astar.cpp:219) and the slot's parent link (Node::parent, :221).
struct ExpandPlan {
  const uint4 *tiles;
  const dcb_step_plan *plan;
  uint32_t *node_g;
  uint32_t *slot_parent;
};

template <int ENV, bool INDEXED, int TEAMS, int SPLIT, int PAD = 0, bool PLANNED = false>
__global__ void __launch_bounds__(TEAMS * SPLIT * 32)
expand_kernel(const uint8_t *__restrict__ src, const uint32_t *__restrict__ ids, int64_t n,
              uint8_t *__restrict__ children, uint8_t *__restrict__ solved, uint64_t *__restrict__ hash, ExpandPlan ep) {
  using Sh = ExpandShape<ENV>;
  static_assert(Sh::A % (SPLIT * Sh::GROUP) == 0, "moves must split into whole packing groups");
  static_assert(PAD == 0 || (SPLIT == 1 && PAD % 16 == 0 && Sh::REC_BYTES % 16 == 0), "per-lane bulk stores need 16-byte records");
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int team = warp / SPLIT, member = warp % SPLIT;
  constexpr int kTileBytes = 32 * Sh::REC_BYTES;
  uint8_t *tile_smem = smem_raw + team * (32 * (Sh::REC_BYTES + PAD));
  uint32_t *lane_rec = reinterpret_cast<uint32_t *>(tile_smem + lane * (Sh::REC_BYTES + PAD));
  const bool issuer = member == 0 && lane == 0;

  const int64_t n_tiles = PLANNED ? (int64_t)ep.plan->n_tiles : ((n + 31) >> 5);
  const int64_t tile_stride = (int64_t)gridDim.x * TEAMS;
  // software pipeline: the parent words of the next tile are requested before this tile is computed
  uint32_t raw_next[LoadShape<Sh::S>::NRAW];
  uint64_t off_next = 0;
  uint4 desc_next = make_uint4(0, 0, 0, 0);
  int64_t tile = (int64_t)blockIdx.x * TEAMS + team;
  auto issue_loads = [&](int64_t t) {
    if constexpr (PLANNED) {
      if (t < n_tiles) {
        desc_next = ep.tiles[t];
        if ((uint32_t)lane < desc_next.z) {
          off_next = (uint64_t)ids[desc_next.x + lane] * Sh::S;
          load_raw<Sh::S>(src, off_next, raw_next);
        }
      }
    } else {
      const int64_t p = t * 32 + lane;
      if (t < n_tiles && p < n) {
        const uint64_t node = INDEXED ? (uint64_t)ids[p] : (uint64_t)p;
        off_next = node * Sh::S;
        load_raw<Sh::S>(src, off_next, raw_next);
      }
    }
  };
  issue_loads(tile);
  for (; tile < n_tiles; tile += tile_stride) {
    const uint4 desc = desc_next;
    // p: index of this lane's parent in the output numbering (children / solved at p * A ...); hp: in the hash numbering
    const int64_t p = PLANNED ? (int64_t)desc.y + lane : tile * 32 + lane;
    const int64_t hp = tile * 32 + lane;
    const bool valid = PLANNED ? ((uint32_t)lane < desc.z) : (p < n);
    uint32_t raw[LoadShape<Sh::S>::NRAW];
#pragma unroll
    for (int k = 0; k < LoadShape<Sh::S>::NRAW; k++) raw[k] = raw_next[k];
    const uint64_t off = off_next;
    issue_loads(tile + tile_stride);
    if constexpr (PAD > 0) {
      bulk_wait_read_all();                // this lane's previous record has left its staging slot
    } else {
      if (issuer) bulk_wait_read_all();    // the previous bulk store has drained this team's staging tile
      team_sync(SPLIT, team);
    }
    if (valid) {
      uint32_t w[Sh::W];
      align_state<Sh::S, Sh::W>(raw, (uint32_t)(off & 3), w);
      SmemSink<ENV> sink;
      sink.rec = lane_rec;
      sink.hash_out = hash ? hash + hp * Sh::A : nullptr;
      sink.solved_out = solved ? reinterpret_cast<uint16_t *>(solved + p * Sh::A) : nullptr;
      if constexpr (PLANNED) {
        if (member == 0) {                                   // Node{depth, parent} of the children (:219-226)
          const uint32_t pid = (uint32_t)(off / Sh::S);
          const uint32_t gc = ep.node_g[pid] + 1;
          uint32_t *gdst = ep.node_g + p * Sh::A;
#pragma unroll
          for (int a = 0; a < Sh::A; a++) gdst[a] = gc;
          ep.slot_parent[p] = pid;
        }
      }
      constexpr int PER = Sh::A / SPLIT;
      if constexpr (SPLIT == 1) expand_parent<ENV, SmemSink<ENV>, 0, Sh::A>(w, sink);
      else if (member == 0) expand_parent<ENV, SmemSink<ENV>, 0, PER>(w, sink);
      else expand_parent<ENV, SmemSink<ENV>, PER, 2 * PER>(w, sink);
    }
    fence_proxy_async_smem();
    if constexpr (PAD > 0) {
      if (valid) {
        bulk_store_s2g(children + p * (int64_t)Sh::REC_BYTES, reinterpret_cast<uint8_t *>(lane_rec), (uint32_t)Sh::REC_BYTES);
        bulk_commit();
      }
      continue;
    }
    team_sync(SPLIT, team);
    const int64_t rem = PLANNED ? (int64_t)desc.z : n - tile * 32;
    const uint32_t bytes = (uint32_t)((rem < 32 ? rem : 32) * Sh::REC_BYTES);
    const uint32_t bulk = bytes & ~15u;
    uint8_t *gdst = PLANNED ? children + (int64_t)desc.y * Sh::REC_BYTES : children + tile * (int64_t)kTileBytes;
    if (issuer && bulk) {
      bulk_store_s2g(gdst, tile_smem, bulk);
      bulk_commit();
    }
    // a partial last tile can leave 4..12 trailing bytes that are not a 16-byte multiple
    if (member == 0 && lane < ((bytes - bulk) >> 2))
      reinterpret_cast<uint32_t *>(gdst + bulk)[lane] = reinterpret_cast<const uint32_t *>(tile_smem + bulk)[lane];
  }
  if (PAD > 0 || issuer) bulk_wait_read_all();
}

// ---- secondary single-state kernels (Environment.next_state / is_solved / hash / nnet input) -------
template <int ENV, int MODE>   // MODE 0: next_state, 1: is_solved, 2: hash
__global__ void __launch_bounds__(256) state_kernel(const uint8_t *__restrict__ states, int64_t n, int action,
                                                    uint8_t *__restrict__ out_states, uint8_t *__restrict__ out_flag,
                                                    uint64_t *__restrict__ out_hash) {
  using Sh = ExpandShape<ENV>;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    uint32_t raw[LoadShape<Sh::S>::NRAW], w[Sh::W];
    const uint64_t off = (uint64_t)p * Sh::S;
    load_raw<Sh::S>(states, off, raw);
    align_state<Sh::S, Sh::W>(raw, (uint32_t)(off & 3), w);
    if (MODE == 1) {
      out_flag[p] = is_goal<ENV, Sh::W>(w) ? 1 : 0;
    } else if (MODE == 2) {
      out_hash[p] = state_hash<Sh::W>(w);
    } else {
      uint32_t zm[Sh::W], c[Sh::W];
      if constexpr (ENV == 5) {                     // Lights Out: XOR with the press mask (lights_out.py:156-166)
        uint32_t mk[Sh::W];
        lo_words_from_bits<Sh::S, Sh::W>(lo_press_mask<EnvTraits<ENV>::DIM>(action), mk);
#pragma unroll
        for (int i = 0; i < Sh::W; i++) { zm[i] = 0; c[i] = w[i] ^ mk[i]; }
      } else {
        if constexpr (EnvTraits<ENV>::kPuzzle) puzzle_blank_mask<EnvTraits<ENV>::DIM, Sh::W>(w, zm);
        else {
#pragma unroll
          for (int i = 0; i < Sh::W; i++) zm[i] = 0;
        }
#pragma unroll
        for (int i = 0; i < Sh::W; i++) c[i] = w[i];
        ApplyAction<ENV, 0>::run(action, w, zm, c);
      }
      uint8_t *o = out_states + off;
      if constexpr (Sh::S % 4 == 0) {
#pragma unroll
        for (int k = 0; k < Sh::S / 4; k++) reinterpret_cast<uint32_t *>(o)[k] = c[k];
      } else if constexpr (Sh::S % 2 == 0) {
#pragma unroll
        for (int k = 0; k < Sh::S / 2; k++) reinterpret_cast<uint16_t *>(o)[k] = (uint16_t)(c[k / 2] >> (16 * (k % 2)));
      } else {
#pragma unroll
        for (int k = 0; k < Sh::S; k++) o[k] = (uint8_t)(c[k / 4] >> (8 * (k % 4)));
      }
    }
  }
}

// nnet input: cube3 sticker id / 9 -> colour id (cube3.py:77-85); cube4 sticker id / 16 (same rule, 16 stickers a face);
// puzzles / Lights Out: identity (n_puzzle.py:84-89)
template <int DIV> __global__ void __launch_bounds__(256) nnet_input_kernel(const uint8_t *__restrict__ in, int64_t nbytes, uint8_t *__restrict__ out) {
  const int64_t nvec = nbytes >> 4;
  const int64_t t0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, ts = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = t0; i < nvec; i += ts) {
    uint4 v = reinterpret_cast<const uint4 *>(in)[i];
    if (DIV == 9) {
      uint32_t *x = reinterpret_cast<uint32_t *>(&v);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        uint32_t r = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) r |= ((((x[k] >> (8 * b)) & 0xFF) * 57u) >> 9) << (8 * b);   // floor(v/9), v < 64
        x[k] = r;
      }
    } else if (DIV == 16) {
      uint32_t *x = reinterpret_cast<uint32_t *>(&v);
#pragma unroll
      for (int k = 0; k < 4; k++) x[k] = (x[k] >> 4) & 0x0F0F0F0Fu;
    }
    reinterpret_cast<uint4 *>(out)[i] = v;
  }
  for (int64_t i = (nvec << 4) + t0; i < nbytes; i += ts)
    out[i] = DIV == 9 ? (uint8_t)((in[i] * 57u) >> 9) : (DIV == 16 ? (uint8_t)(in[i] >> 4) : in[i]);
}

// ---- launchers -------------------------------------------------------------------------------------
static int g_num_sms = 0;
static int num_sms() {
  if (!g_num_sms) {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    g_num_sms = v > 0 ? v : 148;
  }
  return g_num_sms;
}

template <int ENV, bool INDEXED, int TEAMS, int SPLIT, int PAD = 0>
static int launch_expand_cfg(const uint8_t *src, const uint32_t *ids, int64_t n, uint8_t *children, uint8_t *solved, uint64_t *hash,
                             cudaStream_t st) {
  using Sh = ExpandShape<ENV>;
  constexpr int smem = TEAMS * 32 * (Sh::REC_BYTES + PAD);
  constexpr int threads = TEAMS * SPLIT * 32;
  static bool configured = false;
  static int blocks_per_sm = 1;
  auto kern = expand_kernel<ENV, INDEXED, TEAMS, SPLIT, PAD>;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return dcb_cuda_fail();
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, threads, smem) != cudaSuccess) return dcb_cuda_fail();
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    configured = true;
  }
  const int64_t n_tiles = (n + 31) / 32;
  int64_t blocks = (n_tiles + TEAMS - 1) / TEAMS;
  const int64_t cap = (int64_t)num_sms() * blocks_per_sm;
  if (blocks > cap) blocks = cap;
  kern<<<(unsigned)blocks, threads, smem, st>>>(src, ids, n, children, solved, hash, ExpandPlan{});
  return dcb_check_launch();
}

// the search loop's form: work = the iteration's tile list, sizes in device memory
template <int ENV, int TEAMS, int SPLIT, int PAD = 0>
static int launch_expand_planned_cfg(const uint8_t *arena, const uint32_t *ids, int64_t max_tiles, const ExpandPlan &ep, uint8_t *node_solved,
                                     uint64_t *hash, cudaStream_t st) {
  using Sh = ExpandShape<ENV>;
  constexpr int smem = TEAMS * 32 * (Sh::REC_BYTES + PAD);
  constexpr int threads = TEAMS * SPLIT * 32;
  static bool configured = false;
  static int blocks_per_sm = 1;
  auto kern = expand_kernel<ENV, true, TEAMS, SPLIT, PAD, true>;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return dcb_cuda_fail();
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, threads, smem) != cudaSuccess) return dcb_cuda_fail();
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    configured = true;
  }
  int64_t blocks = (max_tiles + TEAMS - 1) / TEAMS;
  const int64_t cap = (int64_t)num_sms() * blocks_per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  kern<<<(unsigned)blocks, threads, smem, st>>>(arena, ids, 0, const_cast<uint8_t *>(arena), node_solved, hash, ep);
  return dcb_check_launch();
}

// cube3 configuration: DCB_EXPAND_CFG="<teams>x<split>" picks an experimental shape (tools/sweep_expand.py)
static int cube3_cfg() {
  static int cfg = -1;
  if (cfg < 0) {
    const char *e = getenv("DCB_EXPAND_CFG");
    cfg = 0;
    if (e) {
      if (!strcmp(e, "4x1")) cfg = 1;
      else if (!strcmp(e, "5x1")) cfg = 2;
      else if (!strcmp(e, "2x2")) cfg = 3;
      else if (!strcmp(e, "5x2")) cfg = 4;
      else if (!strcmp(e, "1x2")) cfg = 5;
    }
  }
  return cfg;
}

template <int ENV, bool INDEXED>
static int launch_expand(const uint8_t *src, const uint32_t *ids, int64_t n, uint8_t *children, uint8_t *solved,
                         uint64_t *hash, cudaStream_t st) {
  if (n == 0) return DCB_OK;
  if constexpr (ENV == 0) {
    switch (cube3_cfg()) {
      case 1: return launch_expand_cfg<0, INDEXED, 4, 1>(src, ids, n, children, solved, hash, st);
      case 2: return launch_expand_cfg<0, INDEXED, 5, 1>(src, ids, n, children, solved, hash, st);
      case 3: return launch_expand_cfg<0, INDEXED, 2, 2>(src, ids, n, children, solved, hash, st);
      case 4: return launch_expand_cfg<0, INDEXED, 5, 2>(src, ids, n, children, solved, hash, st);
      case 5: return launch_expand_cfg<0, INDEXED, 1, 2>(src, ids, n, children, solved, hash, st);
      default: return launch_expand_cfg<0, INDEXED, 4, 1>(src, ids, n, children, solved, hash, st);
    }
  } else if constexpr (ENV == 6) {
    // cube4: a 32-parent tile of records is 72 KB, so three staging tiles fill the SM's shared memory.  Default: the
    // bank-conflict-free staging layout with one bulk store per record (measured 4092 vs 3522 GB/s algorithmic, r01);
    // DCB_CUBE4_PAD=0 selects the contiguous layout with one bulk store per tile.
    static int pad = -1;
    if (pad < 0) { const char *e = getenv("DCB_CUBE4_PAD"); pad = (e && e[0] == '0') ? 0 : 1; }
    if (pad) return launch_expand_cfg<6, INDEXED, 3, 1, 16>(src, ids, n, children, solved, hash, st);
    return launch_expand_cfg<6, INDEXED, 3, 1>(src, ids, n, children, solved, hash, st);
  } else {
    return launch_expand_cfg<ENV, INDEXED, 4, 1>(src, ids, n, children, solved, hash, st);
  }
}

template <bool INDEXED>
static int dispatch_expand(int env, const uint8_t *src, const uint32_t *ids, int64_t n, uint8_t *children, uint8_t *solved,
                           uint64_t *hash, cudaStream_t st) {
  switch (env) {
    case 0: return launch_expand<0, INDEXED>(src, ids, n, children, solved, hash, st);
    case 1: return launch_expand<1, INDEXED>(src, ids, n, children, solved, hash, st);
    case 2: return launch_expand<2, INDEXED>(src, ids, n, children, solved, hash, st);
    case 3: return launch_expand<3, INDEXED>(src, ids, n, children, solved, hash, st);
    case 4: return launch_expand<4, INDEXED>(src, ids, n, children, solved, hash, st);
    case 5: return lightsout_expand_device(src, ids, n, children, solved, hash, st);
    case 6: return launch_expand<6, INDEXED>(src, ids, n, children, solved, hash, st);
  }
  return DCB_ERR_BAD_ENV;
}

int expand_planned_device(int env, uint8_t *arena, const uint32_t *ids, int64_t max_tiles, const uint32_t *tiles, const dcb_step_plan *plan,
                          uint8_t *node_solved, uint64_t *hash, uint32_t *node_g, uint32_t *slot_parent, cudaStream_t st) {
  const ExpandPlan ep{reinterpret_cast<const uint4 *>(tiles), plan, node_g, slot_parent};
  // A* launches are small (625 tiles at B = 20000): one tile per team, and for cube3 two warps share a tile's 12 moves so the
  // tile is computed in half the time (DCB_EXPAND_CFG=4x1 selects the streaming shape)
  switch (env) {
    case 0:
      if (cube3_cfg() == 1) return launch_expand_planned_cfg<0, 4, 1>(arena, ids, max_tiles, ep, node_solved, hash, st);
      return launch_expand_planned_cfg<0, 2, 2>(arena, ids, max_tiles, ep, node_solved, hash, st);
    case 1: return launch_expand_planned_cfg<1, 4, 1>(arena, ids, max_tiles, ep, node_solved, hash, st);
    case 2: return launch_expand_planned_cfg<2, 4, 1>(arena, ids, max_tiles, ep, node_solved, hash, st);
    case 3: return launch_expand_planned_cfg<3, 4, 1>(arena, ids, max_tiles, ep, node_solved, hash, st);
    case 4: return launch_expand_planned_cfg<4, 4, 1>(arena, ids, max_tiles, ep, node_solved, hash, st);
    case 5: return lightsout_expand_planned_device(arena, ids, max_tiles, tiles, plan, node_solved, hash, node_g, slot_parent, st);
    case 6: return launch_expand_planned_cfg<6, 3, 1, 16>(arena, ids, max_tiles, ep, node_solved, hash, st);
  }
  return DCB_ERR_BAD_ENV;
}

int expand_device(int env, const uint8_t *parents, const uint32_t *ids, int64_t n, uint8_t *children, uint8_t *solved,
                  uint64_t *hash, cudaStream_t st) {
  return ids ? dispatch_expand<true>(env, parents, ids, n, children, solved, hash, st)
             : dispatch_expand<false>(env, parents, nullptr, n, children, solved, hash, st);
}

template <int MODE>
static int dispatch_state(int env, const uint8_t *states, int64_t n, int action, uint8_t *out_states, uint8_t *out_flag,
                          uint64_t *out_hash, cudaStream_t st) {
  if (n == 0) return DCB_OK;
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  switch (env) {
    case 0: state_kernel<0, MODE><<<(unsigned)blocks, 256, 0, st>>>(states, n, action, out_states, out_flag, out_hash); break;
    case 1: state_kernel<1, MODE><<<(unsigned)blocks, 256, 0, st>>>(states, n, action, out_states, out_flag, out_hash); break;
    case 2: state_kernel<2, MODE><<<(unsigned)blocks, 256, 0, st>>>(states, n, action, out_states, out_flag, out_hash); break;
    case 3: state_kernel<3, MODE><<<(unsigned)blocks, 256, 0, st>>>(states, n, action, out_states, out_flag, out_hash); break;
    case 4: state_kernel<4, MODE><<<(unsigned)blocks, 256, 0, st>>>(states, n, action, out_states, out_flag, out_hash); break;
    case 5: state_kernel<5, MODE><<<(unsigned)blocks, 256, 0, st>>>(states, n, action, out_states, out_flag, out_hash); break;
    case 6: state_kernel<6, MODE><<<(unsigned)blocks, 256, 0, st>>>(states, n, action, out_states, out_flag, out_hash); break;
    default: return DCB_ERR_BAD_ENV;
  }
  return dcb_check_launch();
}

int next_state_device(int env, const uint8_t *states, int64_t n, int action, uint8_t *out, cudaStream_t st) {
  return dispatch_state<0>(env, states, n, action, out, nullptr, nullptr, st);
}
int is_solved_device(int env, const uint8_t *states, int64_t n, uint8_t *out, cudaStream_t st) {
  return dispatch_state<1>(env, states, n, 0, nullptr, out, nullptr, st);
}
int hash_states_device(int env, const uint8_t *states, int64_t n, uint64_t *out, cudaStream_t st) {
  return dispatch_state<2>(env, states, n, 0, nullptr, nullptr, out, st);
}
int nnet_input_device(int env, const uint8_t *states, int64_t n, uint8_t *out, cudaStream_t st) {
  if (env < 0 || env >= DCB_NUM_ENVS) return DCB_ERR_BAD_ENV;
  if (n == 0) return DCB_OK;
  const int64_t nbytes = n * dcb_env_state_bytes(env);
  int64_t blocks = ((nbytes >> 4) + 255) / 256 + 1;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (env == 0) nnet_input_kernel<9><<<(unsigned)blocks, 256, 0, st>>>(states, nbytes, out);
  else if (env == DCB_ENV_CUBE4) nnet_input_kernel<16><<<(unsigned)blocks, 256, 0, st>>>(states, nbytes, out);
  else nnet_input_kernel<0><<<(unsigned)blocks, 256, 0, st>>>(states, nbytes, out);
  return dcb_check_launch();
}

}  // namespace dcb
