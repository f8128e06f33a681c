// host_check.cpp -- compiles the product's device math (deepcubea_b200/csrc/*.cuh) with plain g++ so the
// CPU-only test tier can compare it with the oracle: PRMT move networks, n-puzzle mask logic, record
// packing, hash, is_solved.  Built by tests/conftest.py into tests/_host/libhostcheck.so.  Not product code.
#include <cstring>
#include "expand_core.cuh"

using namespace dcb;

namespace {
template <int ENV> struct HostSink {
  using Sh = ExpandShape<ENV>;
  uint8_t *rec; uint8_t *solved; uint64_t *hash;
  template <int K0, int N> void store_record_words(const uint32_t (&r)[N]) { std::memcpy(rec + 4 * K0, r, 4 * N); }
  template <int MOVE> void store_hash(uint64_t h) { hash[MOVE] = h; }
  template <int MOVE> void store_solved(bool s) { solved[MOVE] = s ? 1 : 0; }
};

template <int ENV> void load_words(const uint8_t *base, int64_t off, uint32_t (&w)[ExpandShape<ENV>::W]) {
  constexpr int S = EnvTraits<ENV>::S;
  constexpr int NRAW = LoadShape<S>::NRAW;
  // emulate the kernel's aligned-word loads + funnel shift (base is 4-byte aligned by construction)
  const int64_t a0 = off & ~int64_t(3);
  uint32_t raw[NRAW];
  for (int k = 0; k < NRAW; k++) std::memcpy(&raw[k], base + a0 + 4 * k, 4);
  align_state<S, ExpandShape<ENV>::W>(raw, (uint32_t)(off & 3), w);
}

template <int ENV> void expand_env(const uint8_t *parents, int64_t n, uint8_t *children, uint8_t *solved, uint64_t *hash) {
  using Sh = ExpandShape<ENV>;
  for (int64_t p = 0; p < n; p++) {
    uint32_t w[Sh::W];
    load_words<ENV>(parents, p * Sh::S, w);
    HostSink<ENV> sink{children + p * Sh::REC_BYTES, solved + p * Sh::A, hash + p * Sh::A};
    expand_parent<ENV>(w, sink);
  }
}

template <int ENV> void next_env(const uint8_t *states, int64_t n, int action, uint8_t *out) {
  using Sh = ExpandShape<ENV>;
  for (int64_t p = 0; p < n; p++) {
    uint32_t w[Sh::W], zm[Sh::W], c[Sh::W];
    load_words<ENV>(states, p * Sh::S, w);
    if (EnvTraits<ENV>::kPuzzle) puzzle_blank_mask<EnvTraits<ENV>::DIM, Sh::W>(w, zm);
    else for (int i = 0; i < Sh::W; i++) zm[i] = 0;
    ApplyAction<ENV, 0>::run(action, w, zm, c);
    std::memcpy(out + p * Sh::S, c, Sh::S);
  }
}
}  // namespace

// Lights Out 7x7: the bit-form math of csrc/lightsout_kernels.cu (bytes -> bits, press masks, bits -> bytes, hash, solved)
extern "C" int hc_lightsout_expand(const uint8_t *parents, int64_t n, uint8_t *children, uint8_t *solved, uint64_t *hash) {
  constexpr int S = 49, A = 49, W = 14;
  for (int64_t p = 0; p < n; p++) {
    uint32_t w[W];
    load_words<5>(parents, p * S, w);
    const uint64_t pb = lo_bits_from_words<W>(w);
    for (int m = 0; m < A; m++) {
      const uint64_t b = pb ^ lo_press_mask<7>(m);
      uint32_t c[W];
      lo_words_from_bits<S, W>(b, c);
      std::memcpy(children + (p * A + m) * S, c, S);
      hash[p * A + m] = state_hash<W>(c);
      solved[p * A + m] = b == 0;
    }
  }
  return 0;
}

// `parents` must be readable for 4 bytes past the end (callers pad).
extern "C" int hc_expand(int env, const uint8_t *parents, int64_t n, uint8_t *children, uint8_t *solved, uint64_t *hash) {
  switch (env) {
    case 0: expand_env<0>(parents, n, children, solved, hash); return 0;
    case 1: expand_env<1>(parents, n, children, solved, hash); return 0;
    case 2: expand_env<2>(parents, n, children, solved, hash); return 0;
    case 3: expand_env<3>(parents, n, children, solved, hash); return 0;
    case 4: expand_env<4>(parents, n, children, solved, hash); return 0;
    case 6: expand_env<6>(parents, n, children, solved, hash); return 0;
  }
  return -1;
}
extern "C" int hc_is_goal(int env, const uint8_t *states, int64_t n, uint8_t *out) {
  if (env != 6) return -1;
  for (int64_t p = 0; p < n; p++) {
    uint32_t w[ExpandShape<6>::W];
    load_words<6>(states, p * 96, w);
    out[p] = is_goal<6, ExpandShape<6>::W>(w) ? 1 : 0;
  }
  return 0;
}
extern "C" int hc_next_state(int env, const uint8_t *states, int64_t n, int action, uint8_t *out) {
  switch (env) {
    case 0: next_env<0>(states, n, action, out); return 0;
    case 1: next_env<1>(states, n, action, out); return 0;
    case 2: next_env<2>(states, n, action, out); return 0;
    case 3: next_env<3>(states, n, action, out); return 0;
    case 4: next_env<4>(states, n, action, out); return 0;
    case 6: next_env<6>(states, n, action, out); return 0;
  }
  return -1;
}
