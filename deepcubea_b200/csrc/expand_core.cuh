// expand_core.cuh -- "all children of one parent" as straight-line register code, parameterised on a
// sink so the CUDA kernel (shared-memory staging + TMA bulk store) and the CPU-side test harness
// (tests/host_check.cpp) run the very same math.
//
// Sink concept:
//   template <int K0, int N> void store_record_words(const uint32_t (&rec)[N]);  // record words K0..K0+N-1
//   template <int MOVE> void store_hash(uint64_t h);
//   template <int MOVE> void store_solved(bool s);
#pragma once
#include "state_ops.cuh"

namespace dcb {

template <int ENV> struct ExpandShape {
  static constexpr int S = EnvTraits<ENV>::S;
  static constexpr int A = EnvTraits<ENV>::A;
  static constexpr int W = hash_words(S);
  static constexpr int REC_BYTES = S * A;
  static constexpr int REC_WORDS = S * A / 4;
  static constexpr int GROUP = (ENV == 0) ? 2 : (ENV == 6 ? 1 : 4);   // children packed per group (cube3: pairs of 108 B; cube4: 96 B = whole words)
  static constexpr int GROUP_WORDS = GROUP * S / 4;
  static_assert(ENV == 5 || ((S * A) % 4 == 0 && (GROUP * S) % 4 == 0), "record shape");   // Lights Out (2401-byte records) has its own kernel
};

template <int ENV, int M0, int G, class Sink> struct ExpandGroup {
  static DCB_DEV void run(const uint32_t (&p)[ExpandShape<ENV>::W], const uint32_t (&zm)[ExpandShape<ENV>::W],
                          uint64_t goal_hash, Sink &sink) {
    using Sh = ExpandShape<ENV>;
    uint32_t ch[G][Sh::W];
    expand_children<0>(p, zm, goal_hash, ch, sink);
    uint32_t rec[G * Sh::S / 4];
    pack_record<Sh::S, G, Sh::W>(ch, rec);
    sink.template store_record_words<M0 * Sh::S / 4, G * Sh::S / 4>(rec);
  }
  template <int I>
  static DCB_DEV void expand_children(const uint32_t (&p)[ExpandShape<ENV>::W], const uint32_t (&zm)[ExpandShape<ENV>::W],
                                      uint64_t goal_hash, uint32_t (&ch)[G][ExpandShape<ENV>::W], Sink &sink) {
    using Sh = ExpandShape<ENV>;
    if constexpr (I < G) {
      ChildOf<ENV, M0 + I, Sh::W>::apply(p, zm, ch[I]);
      const uint64_t h = state_hash<Sh::W>(ch[I]);
      sink.template store_hash<M0 + I>(h);
      // is_solved: a solved child must hash to the goal's hash; verify the (rare) candidates exactly.  Environments with
      // many solved states (cube4) are tested directly.
      bool solved = false;
      if constexpr (!GoalTraits<ENV>::kGoalIsUnique) solved = is_goal<ENV, Sh::W>(ch[I]);
      else if (h == goal_hash) solved = is_goal<ENV, Sh::W>(ch[I]);
      sink.template store_solved<M0 + I>(solved);
      expand_children<I + 1>(p, zm, goal_hash, ch, sink);
    }
  }
};

template <int ENV, int M0, int M1, class Sink> struct ExpandRange {     // children M0 .. M1-1
  static DCB_DEV void run(const uint32_t (&p)[ExpandShape<ENV>::W], const uint32_t (&zm)[ExpandShape<ENV>::W],
                          uint64_t goal_hash, Sink &sink) {
    using Sh = ExpandShape<ENV>;
    static_assert(M0 % Sh::GROUP == 0 && M1 % Sh::GROUP == 0, "move ranges are whole packing groups");
    if constexpr (M0 < M1) {
      ExpandGroup<ENV, M0, Sh::GROUP, Sink>::run(p, zm, goal_hash, sink);
      ExpandRange<ENV, M0 + Sh::GROUP, M1, Sink>::run(p, zm, goal_hash, sink);
    }
  }
};

template <int ENV> DCB_DEV uint64_t goal_hash_value() {
  using Sh = ExpandShape<ENV>;
  uint32_t g[Sh::W];
#pragma unroll
  for (int i = 0; i < Sh::W; i++) g[i] = goal_word<ENV>(i);
  return state_hash<Sh::W>(g);
}

// p: aligned parent words (bytes >= S zero).  Children [M0, M1) only (default: all) -- lets several warps share a parent.
template <int ENV, class Sink, int M0 = 0, int M1 = ExpandShape<ENV>::A>
DCB_DEV void expand_parent(const uint32_t (&p)[ExpandShape<ENV>::W], Sink &sink) {
  using Sh = ExpandShape<ENV>;
  uint32_t zm[Sh::W];
  if constexpr (EnvTraits<ENV>::kPuzzle) {
    puzzle_blank_mask<EnvTraits<ENV>::DIM, Sh::W>(p, zm);
  } else {
#pragma unroll
    for (int i = 0; i < Sh::W; i++) zm[i] = 0;
  }
  const uint64_t gh = goal_hash_value<ENV>();
  ExpandRange<ENV, M0, M1, Sink>::run(p, zm, gh, sink);
}

// One child for a runtime action (Environment.next_state).
template <int ENV, int M> struct ApplyAction {
  static DCB_DEV void run(int action, const uint32_t (&p)[ExpandShape<ENV>::W], const uint32_t (&zm)[ExpandShape<ENV>::W],
                          uint32_t (&c)[ExpandShape<ENV>::W]) {
    if constexpr (M < EnvTraits<ENV>::A) {
      if (action == M) ChildOf<ENV, M, ExpandShape<ENV>::W>::apply(p, zm, c);
      else ApplyAction<ENV, M + 1>::run(action, p, zm, c);
    }
  }
};

}  // namespace dcb
