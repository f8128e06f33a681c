// Test infrastructure (oracle/): drives the UNMODIFIED reference Cube4 (cpp/environments.cpp:262-370, compiled where it lies
// under /root/reference by `make -C oracle ref`) to produce golden vectors for the cube4 environment.
//   stdin : n, then n states of 96 bytes (text integers)
//   stdout: per state 24 children (96 integers each) then isSolved of the state and of every child (25 flags)
#include <cstdint>
#include <cstdio>
#include <vector>
#include "environments.h"

int main() {
  int n = 0;
  if (scanf("%d", &n) != 1) return 1;
  for (int i = 0; i < n; i++) {
    std::vector<uint8_t> s(96);
    for (int j = 0; j < 96; j++) { int v; if (scanf("%d", &v) != 1) return 1; s[j] = (uint8_t)v; }
    Cube4 root(s);
    std::vector<Environment *> ch = root.getNextStates();
    for (size_t a = 0; a < ch.size(); a++) {
      const std::vector<uint8_t> c = ch[a]->getState();
      for (int j = 0; j < 96; j++) printf("%d ", (int)c[j]);
      printf("\n");
    }
    printf("%d", root.isSolved() ? 1 : 0);
    for (size_t a = 0; a < ch.size(); a++) { printf(" %d", ch[a]->isSolved() ? 1 : 0); delete ch[a]; }
    printf("\n");
  }
  return 0;
}
