// state_ops.cuh -- register-level state math shared by every kernel (and by tests/host_check.cpp).
//
// A state of S bytes lives in W = 2*ceil(S/8) little-endian u32 words with the bytes past S zeroed, so the
// words feed the hash directly.  Everything here is branch-free straight-line code after unrolling: move
// application never indexes registers dynamically.
#pragma once
#include "cube3_moves.cuh"
#include "cube4_moves.cuh"
#include "intrinsics.cuh"

namespace dcb {

constexpr int kEnvCube3 = 0;

template <int ENV> struct EnvTraits;
template <> struct EnvTraits<0> { static constexpr int S = 54, A = 12, DIM = 3; static constexpr bool kPuzzle = false; };
template <> struct EnvTraits<1> { static constexpr int S = 16, A = 4, DIM = 4; static constexpr bool kPuzzle = true; };
template <> struct EnvTraits<2> { static constexpr int S = 25, A = 4, DIM = 5; static constexpr bool kPuzzle = true; };
template <> struct EnvTraits<3> { static constexpr int S = 36, A = 4, DIM = 6; static constexpr bool kPuzzle = true; };
template <> struct EnvTraits<4> { static constexpr int S = 49, A = 4, DIM = 7; static constexpr bool kPuzzle = true; };
template <> struct EnvTraits<5> { static constexpr int S = 49, A = 49, DIM = 7; static constexpr bool kPuzzle = false; };   // Lights Out 7x7
template <> struct EnvTraits<6> { static constexpr int S = 96, A = 24, DIM = 4; static constexpr bool kPuzzle = false; };   // 4x4x4 cube (cpp/environments.cpp:262-370)

DCB_HOSTDEV constexpr int hash_words(int s) { return 2 * ((s + 7) / 8); }
DCB_HOSTDEV constexpr int gcd4(int s) { return (s % 4 == 0) ? 4 : ((s % 2 == 0) ? 2 : 1); }

// ---------------------------------------------------------------------------------------------------
// Hash: NH pair-product universal hash (keys from splitmix64, see oracle/oracle_env.py) + murmur3 fmix64.
// On sm_100a each term is one IMAD.WIDE.U32 with 64-bit accumulate; the adds are IADD3.
// ---------------------------------------------------------------------------------------------------
constexpr uint64_t kHashSeed = 0x9E3779B97F4A7C15ull;
DCB_HOSTDEV constexpr uint32_t hash_key(int i) {
  constexpr uint32_t k[24] = {0xa82e9745u, 0x275e0ca1u, 0xa22c9073u, 0x922741ddu, 0x98b201b5u, 0xebbdc7b7u,
                              0xed62b2c5u, 0xdedc17fbu, 0x963a6a25u, 0xa4b0b7d7u, 0xd613fdbfu, 0xd5658cfdu,
                              0x6b33c2a1u, 0x7a9e4ccbu, 0x3ae98f97u, 0x6d554fd7u, 0x984ad69du, 0x7f654559u,
                              0xcae5c8b1u, 0x589234d5u, 0x6dd827ffu, 0x8e496109u, 0xc5e648bbu, 0x99730cdbu};
  return k[i];
}

DCB_HOSTDEV uint64_t fmix64(uint64_t h) {
  h ^= h >> 33; h *= 0xFF51AFD7ED558CCDull;
  h ^= h >> 33; h *= 0xC4CEB9FE1A85EC53ull;
  h ^= h >> 33;
  return h ? h : 1ull;   // 0 is the closed table's EMPTY key
}

template <int W> DCB_DEV uint64_t state_hash(const uint32_t (&w)[W]) {
  static_assert(W % 2 == 0 && W <= 24, "hash words");
  uint64_t acc = kHashSeed;
#pragma unroll
  for (int i = 0; i < W; i += 2)
    acc += (uint64_t)(uint32_t)(w[i] + hash_key(i)) * (uint64_t)(uint32_t)(w[i + 1] + hash_key(i + 1));
  return fmix64(acc);
}

// ---------------------------------------------------------------------------------------------------
// Goal states
// ---------------------------------------------------------------------------------------------------
template <int ENV> DCB_HOSTDEV constexpr uint8_t goal_byte(int j) {
  // cube3: sticker identity (cube3.py:37, :71-75).  n-puzzle: [1..n*n-1, 0] (n_puzzle.py:41).  Lights Out: all off.
  // cube4: sticker identity (one of its many solved states, see is_goal<6>).
  return (ENV == 0 || ENV == 6) ? (uint8_t)j : (ENV == 5 ? (uint8_t)0 : (uint8_t)((j + 1) % EnvTraits<ENV>::S));
}
template <int ENV> DCB_HOSTDEV constexpr uint32_t goal_word(int w) {
  constexpr int S = EnvTraits<ENV>::S;
  uint32_t v = 0;
  for (int b = 0; b < 4; b++) {
    const int j = 4 * w + b;
    if (j < S) v |= (uint32_t)goal_byte<ENV>(j) << (8 * b);
  }
  return v;
}
// kGoalIsUnique: one solved state, so "hash equals the goal's hash" pre-filters the exact compare.  The 4x4x4 cube is
// solved when every face shows one colour = sticker id / 16 (Cube4::isSolved, cpp/environments.cpp:356-366): whole-cube
// rotations and stickers exchanged inside a face count too, so it is tested directly on the high nibbles.
template <int ENV> struct GoalTraits { static constexpr bool kGoalIsUnique = ENV != 6; };
template <int ENV, int W> DCB_DEV bool is_goal(const uint32_t (&w)[W]) {
  uint32_t diff = 0;
  if constexpr (ENV == 6) {
#pragma unroll
    for (int f = 0; f < 6; f++) {
      const uint32_t colour = (w[4 * f] & 0xF0u) * 0x01010101u;        // byte 0's colour nibble in all four bytes
#pragma unroll
      for (int k = 0; k < 4; k++) diff |= (w[4 * f + k] & 0xF0F0F0F0u) ^ colour;
    }
  } else {
#pragma unroll
    for (int i = 0; i < W; i++) diff |= w[i] ^ goal_word<ENV>(i);
  }
  return diff == 0;
}

// ---------------------------------------------------------------------------------------------------
// Static byte shifts of a W-word little-endian byte array (bytes shifted out are lost, zeros shifted in)
// ---------------------------------------------------------------------------------------------------
template <int W, int K> DCB_DEV void shift_up_bytes(const uint32_t (&in)[W], uint32_t (&out)[W]) {
  constexpr int q = K / 4, r = K % 4;   // out byte j = in byte j-K
#pragma unroll
  for (int w = 0; w < W; w++) {
    const uint32_t hi = (w - q >= 0) ? in[(w - q >= 0) ? w - q : 0] : 0u;
    const uint32_t lo = (w - q - 1 >= 0) ? in[(w - q - 1 >= 0) ? w - q - 1 : 0] : 0u;
    out[w] = (r == 0) ? hi : __funnelshift_l(lo, hi, 8 * r);
  }
}
template <int W, int K> DCB_DEV void shift_down_bytes(const uint32_t (&in)[W], uint32_t (&out)[W]) {
  constexpr int q = K / 4, r = K % 4;   // out byte j = in byte j+K
#pragma unroll
  for (int w = 0; w < W; w++) {
    const uint32_t lo = (w + q < W) ? in[(w + q < W) ? w + q : 0] : 0u;
    const uint32_t hi = (w + q + 1 < W) ? in[(w + q + 1 < W) ? w + q + 1 : 0] : 0u;
    out[w] = (r == 0) ? lo : __funnelshift_r(lo, hi, 8 * r);
  }
}

// ---------------------------------------------------------------------------------------------------
// n-puzzle: all four children of one parent, SIMD-within-register.
//   moves U,D,L,R swap the blank with the tile at (i+1,j),(i-1,j),(i,j+1),(i,j-1); an illegal move leaves
//   the state unchanged (n_puzzle.py:174-231; cpp/environments.cpp:4-46, 92-104).
//   zmask marks the blank byte; shifting it by +-DIM / +-1 bytes marks the tile to swap; XOR-ing the tile
//   value into both positions performs the swap.  No table, no dynamic byte indexing.
// ---------------------------------------------------------------------------------------------------
template <int DIM> DCB_HOSTDEV constexpr uint32_t col_mask(int w, int excluded_col) {
  // 0xFF for every byte j < DIM*DIM in word w whose column (j % DIM) != excluded_col
  uint32_t v = 0;
  for (int b = 0; b < 4; b++) {
    const int j = 4 * w + b;
    if (j < DIM * DIM && (j % DIM) != excluded_col) v |= 0xFFu << (8 * b);
  }
  return v;
}
template <int S> DCB_HOSTDEV constexpr uint32_t valid_mask(int w) {
  uint32_t v = 0;
  for (int b = 0; b < 4; b++)
    if (4 * w + b < S) v |= 0xFFu << (8 * b);
  return v;
}

template <int DIM, int MOVE, int W> DCB_DEV void puzzle_child(const uint32_t (&p)[W], const uint32_t (&zm)[W], uint32_t (&c)[W]) {
  constexpr int S = DIM * DIM;
  uint32_t z[W], sm[W], t[W], tz[W];
  if (MOVE == 0) {          // U: swap with byte z + DIM
#pragma unroll
    for (int w = 0; w < W; w++) z[w] = zm[w];
    shift_up_bytes<W, DIM>(z, sm);
  } else if (MOVE == 1) {   // D: swap with byte z - DIM
#pragma unroll
    for (int w = 0; w < W; w++) z[w] = zm[w];
    shift_down_bytes<W, DIM>(z, sm);
  } else if (MOVE == 2) {   // L: swap with byte z + 1, only if the blank is not in the last column
#pragma unroll
    for (int w = 0; w < W; w++) z[w] = zm[w] & col_mask<DIM>(w, DIM - 1);
    shift_up_bytes<W, 1>(z, sm);
  } else {                  // R: swap with byte z - 1, only if the blank is not in the first column
#pragma unroll
    for (int w = 0; w < W; w++) z[w] = zm[w] & col_mask<DIM>(w, 0);
    shift_down_bytes<W, 1>(z, sm);
  }
#pragma unroll
  for (int w = 0; w < W; w++) t[w] = p[w] & sm[w] & valid_mask<S>(w);   // tile value at its own position
  if (MOVE == 0) shift_down_bytes<W, DIM>(t, tz);
  else if (MOVE == 1) shift_up_bytes<W, DIM>(t, tz);
  else if (MOVE == 2) shift_down_bytes<W, 1>(t, tz);
  else shift_up_bytes<W, 1>(t, tz);
#pragma unroll
  for (int w = 0; w < W; w++) c[w] = p[w] ^ t[w] ^ tz[w];
}

template <int DIM, int W> DCB_DEV void puzzle_blank_mask(const uint32_t (&p)[W], uint32_t (&zm)[W]) {
#pragma unroll
  for (int w = 0; w < W; w++) zm[w] = __vcmpeq4(p[w], 0u) & valid_mask<DIM * DIM>(w);
}

// ---------------------------------------------------------------------------------------------------
// Lights Out (environments/lights_out.py:26-166, cpp/environments.cpp:133-208): a state is DIM*DIM cells of 0/1, i.e. a
// 49-bit word for the 7x7 board; pressing cell m toggles m and its in-board neighbours m+-DIM (x axis) and m+-1 (y axis).
// All arithmetic happens on the bit form; bytes <-> bits conversions are multiply tricks on 4-byte words.
// ---------------------------------------------------------------------------------------------------
template <int DIM> DCB_HOSTDEV uint64_t lo_press_mask(int m) {
  const int x = m / DIM, y = m % DIM;                       // lights_out.py:34-42
  uint64_t k = 1ull << m;
  if (x < DIM - 1) k |= 1ull << (m + DIM);
  if (x > 0) k |= 1ull << (m - DIM);
  if (y < DIM - 1) k |= 1ull << (m + 1);
  if (y > 0) k |= 1ull << (m - 1);
  return k;
}
// 4 cells (one per byte, values 0/1) -> 4 bits
DCB_HOSTDEV uint32_t lo_pack4(uint32_t w) { return (((w & 0x01010101u) * 0x01020408u) >> 24) & 0xFu; }
// 4 bits -> 4 cells (one per byte)
DCB_HOSTDEV uint32_t lo_spread4(uint32_t nib) { return ((nib & 0xFu) * 0x00204081u) & 0x01010101u; }
template <int W> DCB_DEV uint64_t lo_bits_from_words(const uint32_t (&w)[W]) {
  uint64_t b = 0;
#pragma unroll
  for (int k = 0; k < W; k++) b |= (uint64_t)lo_pack4(w[k]) << (4 * k);
  return b;
}
template <int S, int W> DCB_DEV void lo_words_from_bits(uint64_t bits, uint32_t (&w)[W]) {
#pragma unroll
  for (int k = 0; k < W; k++) w[k] = lo_spread4((uint32_t)(bits >> (4 * k))) & valid_mask<S>(k);
}

// ---------------------------------------------------------------------------------------------------
// Generic single child: ENV, MOVE compile-time
// ---------------------------------------------------------------------------------------------------
template <int ENV, int MOVE, int W> struct ChildOf;
template <int MOVE> struct ChildOf<0, MOVE, 14> {
  static DCB_DEV void apply(const uint32_t (&p)[14], const uint32_t (&)[14], uint32_t (&c)[14]) { cube3_move<MOVE>(p, c); }
};
template <int MOVE> struct ChildOf<6, MOVE, 24> {
  static DCB_DEV void apply(const uint32_t (&p)[24], const uint32_t (&)[24], uint32_t (&c)[24]) { cube4_move<MOVE>(p, c); }
};
template <int ENV, int MOVE, int W> struct ChildOf {
  static DCB_DEV void apply(const uint32_t (&p)[W], const uint32_t (&zm)[W], uint32_t (&c)[W]) {
    puzzle_child<EnvTraits<ENV>::DIM, MOVE, W>(p, zm, c);
  }
};

// ---------------------------------------------------------------------------------------------------
// Record packing: NC consecutive children of S bytes each, given as aligned W-word arrays (bytes >= S
// zero), concatenated without padding into NC*S/4 record words.
// ---------------------------------------------------------------------------------------------------
template <int S, int W> DCB_DEV uint32_t bytes_at(const uint32_t (&x)[W], int o) {
  // 4 bytes of x starting at (static) byte offset o; bytes past the array read as 0
  const int q = o / 4, r = o % 4;
  const uint32_t lo = (q < W) ? x[(q < W) ? q : 0] : 0u;
  const uint32_t hi = (q + 1 < W) ? x[(q + 1 < W) ? q + 1 : 0] : 0u;
  return (r == 0) ? lo : __funnelshift_r(lo, hi, 8 * r);
}

template <int S, int NC, int W> DCB_DEV void pack_record(const uint32_t (&ch)[NC][W], uint32_t (&rec)[NC * S / 4]) {
  static_assert((NC * S) % 4 == 0, "record must be a whole number of words");
#pragma unroll
  for (int k = 0; k < NC * S / 4; k++) {
    const int c = (4 * k) / S, o = 4 * k - S * c;
    uint32_t v = bytes_at<S, W>(ch[c], o);
    if (o + 4 > S) {                       // straddles into child c+1 (bytes past S of child c are zero)
      const int n0 = S - o;
      v |= ch[(c + 1 < NC) ? c + 1 : c][0] << (8 * n0);
    }
    rec[k] = v;
  }
}

// ---------------------------------------------------------------------------------------------------
// Loading a state from a byte address with arbitrary (env-dependent) alignment into aligned words.
// raw[] must hold NRAW = (S + 4 - gcd4(S) + 3) / 4 consecutive aligned words starting at (addr & ~3).
// ---------------------------------------------------------------------------------------------------
template <int S> struct LoadShape { static constexpr int NRAW = (S + (4 - gcd4(S)) + 3) / 4; };

template <int S, int W> DCB_DEV void align_state(const uint32_t (&raw)[LoadShape<S>::NRAW], uint32_t byte_in_word, uint32_t (&w)[W]) {
  constexpr int NRAW = LoadShape<S>::NRAW;
  const uint32_t sh = 8 * byte_in_word;
#pragma unroll
  for (int k = 0; k < W; k++) {
    const uint32_t lo = (k < NRAW) ? raw[(k < NRAW) ? k : 0] : 0u;
    const uint32_t hi = (k + 1 < NRAW) ? raw[(k + 1 < NRAW) ? k + 1 : 0] : 0u;
    const uint32_t v = (gcd4(S) == 4) ? lo : __funnelshift_r(lo, hi, sh);
    w[k] = v & valid_mask<S>(k);
  }
}

}  // namespace dcb
