/* C restatement of the reference's native environment step.  TEST INFRASTRUCTURE ONLY -- never linked
 * into or called by the product library; used by tests/ as the fast checker at BASELINE sizes and by
 * bench.py's cpu_baseline leg.  Tables are passed in by the caller (tests/golden/*.json), so nothing here
 * is copied from the reference's literals.
 *
 *   oracle_cube3_expand   cpp/environments.cpp:222-243 (getNextState: 24 scalar copies per move;
 *                         getNextStates: all 12 moves) + :249-256 (isSolved: state[i] == i)
 *   oracle_puzzle_expand  cpp/environments.cpp:92-113 (swap blank with swapZeroIdxs[zIdx][action])
 *                         + :119-126 (isSolved: state[i] == (i+1) % numTiles)
 *   oracle_cube4_expand   cpp/environments.cpp:327-350 (Cube4::getNextState(s), gather form) + :356-366 (isSolved: one colour
 *                         = id / 16 per face)
 *   oracle_hash64         project-defined hash (see oracle/oracle_env.py:state_hash64)
 */
#include <stdint.h>
#include <string.h>

void oracle_cube3_expand(const uint8_t *parents, int64_t n, const int32_t *idx_new /*[12][24]*/,
                         const int32_t *idx_old /*[12][24]*/, uint8_t *children /*[n][12][54]*/,
                         uint8_t *solved /*[n][12]*/) {
#pragma omp parallel for schedule(static)
  for (int64_t p = 0; p < n; p++) {
    const uint8_t *cur = parents + p * 54;
    for (int a = 0; a < 12; a++) {
      uint8_t *nxt = children + (p * 12 + a) * 54;
      memcpy(nxt, cur, 54);
      for (int i = 0; i < 24; i++) nxt[idx_new[a * 24 + i]] = cur[idx_old[a * 24 + i]];
      uint8_t ok = 1;
      for (int i = 0; i < 54; i++) ok &= (uint8_t)(nxt[i] == i);
      solved[p * 12 + a] = ok;
    }
  }
}

void oracle_puzzle_expand(const uint8_t *parents, int64_t n, int dim, const int32_t *swap /*[dim*dim][4]*/,
                          uint8_t *children /*[n][4][dim*dim]*/, uint8_t *solved /*[n][4]*/) {
  const int s = dim * dim;
#pragma omp parallel for schedule(static)
  for (int64_t p = 0; p < n; p++) {
    const uint8_t *cur = parents + p * s;
    int z = 0;
    for (int i = 0; i < s; i++)
      if (cur[i] == 0) { z = i; break; }
    for (int a = 0; a < 4; a++) {
      uint8_t *nxt = children + (p * 4 + a) * s;
      memcpy(nxt, cur, (size_t)s);
      const int sw = swap[z * 4 + a];
      const uint8_t val = nxt[sw];
      nxt[z] = val;
      nxt[sw] = 0;
      uint8_t ok = 1;
      for (int i = 0; i < s; i++) ok &= (uint8_t)(nxt[i] == (uint8_t)((i + 1) % s));
      solved[p * 4 + a] = ok;
    }
  }
}

/* cpp/environments.cpp:171-196 (LightsOut::getNextState(s): newState[moveMat[a][i]] = (state[...] + 1) % 2 from the ORIGINAL
 * state) + :198-206 (isSolved: all zero). */
void oracle_lightsout_expand(const uint8_t *parents, int64_t n, int dim, const int32_t *move_mat /*[dim*dim][5]*/,
                             uint8_t *children /*[n][dim*dim][dim*dim]*/, uint8_t *solved /*[n][dim*dim]*/) {
  const int s = dim * dim;
#pragma omp parallel for schedule(static)
  for (int64_t p = 0; p < n; p++) {
    const uint8_t *cur = parents + p * s;
    for (int a = 0; a < s; a++) {
      uint8_t *nxt = children + (p * s + a) * s;
      memcpy(nxt, cur, (size_t)s);
      for (int i = 0; i < 5; i++) nxt[move_mat[a * 5 + i]] = (uint8_t)((cur[move_mat[a * 5 + i]] + 1) % 2);
      uint8_t ok = 1;
      for (int i = 0; i < s; i++) ok &= (uint8_t)(nxt[i] == 0);
      solved[p * s + a] = ok;
    }
  }
}

void oracle_cube4_expand(const uint8_t *parents, int64_t n, const int32_t *perm /*[24][96]*/, uint8_t *children /*[n][24][96]*/,
                         uint8_t *solved /*[n][24]*/) {
#pragma omp parallel for schedule(static)
  for (int64_t p = 0; p < n; p++) {
    const uint8_t *cur = parents + p * 96;
    for (int a = 0; a < 24; a++) {
      uint8_t *nxt = children + (p * 24 + a) * 96;
      for (int j = 0; j < 96; j++) nxt[j] = cur[perm[a * 96 + j]];
      uint8_t ok = 1;
      for (int side = 0; side < 6; side++)
        for (int i = 1; i < 16; i++) ok &= (uint8_t)(nxt[side * 16 + i] / 16 == nxt[side * 16] / 16);
      solved[p * 24 + a] = ok;
    }
  }
}

void oracle_hash64(const uint8_t *states, int64_t n, int state_dim, const uint32_t *keys /*[24]*/,
                   uint64_t seed, uint64_t *out) {
  const int w = 2 * ((state_dim + 7) / 8);
#pragma omp parallel for schedule(static)
  for (int64_t k = 0; k < n; k++) {
    uint8_t buf[96];
    memset(buf, 0, sizeof buf);
    memcpy(buf, states + k * state_dim, (size_t)state_dim);
    uint64_t acc = seed;
    for (int i = 0; i < w; i += 2) {
      uint32_t a, b;
      memcpy(&a, buf + 4 * i, 4);
      memcpy(&b, buf + 4 * i + 4, 4);
      acc += (uint64_t)(uint32_t)(a + keys[i]) * (uint64_t)(uint32_t)(b + keys[i + 1]);
    }
    acc ^= acc >> 33; acc *= 0xFF51AFD7ED558CCDull;
    acc ^= acc >> 33; acc *= 0xC4CEB9FE1A85EC53ull;
    acc ^= acc >> 33;
    out[k] = acc ? acc : 1;
  }
}
