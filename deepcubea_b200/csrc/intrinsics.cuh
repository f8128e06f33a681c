// intrinsics.cuh -- qualifiers and, for plain g++ builds of tests/host_check.cpp, bit-exact host
// emulations of the handful of CUDA integer intrinsics the state math uses.  Compiling the SAME device
// math with g++ lets the CPU-only test tier check the PRMT networks, the hash and the puzzle mask logic
// against the oracle before any GPU time is spent.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define DCB_DEV __device__ __forceinline__
#define DCB_HOSTDEV __host__ __device__ __forceinline__
#else
#define DCB_DEV inline
#define DCB_HOSTDEV inline
// PRMT, default mode: result byte i = byte (sel nibble i & 7) of the 8-byte pool {y:x}; nibble bit 3
// replicates the sign bit of that byte (never used by this code base, asserted in the generator).
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s) {
  const uint64_t pool = ((uint64_t)y << 32) | x;
  uint32_t r = 0;
  for (int i = 0; i < 4; i++) {
    const uint32_t nib = (s >> (4 * i)) & 0xF;
    r |= (uint32_t)((pool >> (8 * (nib & 7))) & 0xFF) << (8 * i);
  }
  return r;
}
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) {
  const uint64_t v = ((uint64_t)hi << 32) | lo;
  return (uint32_t)(v >> (sh & 31));
}
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t sh) {
  const uint64_t v = ((uint64_t)hi << 32) | lo;
  return (uint32_t)((v << (sh & 31)) >> 32);
}
static inline uint32_t __vcmpeq4(uint32_t a, uint32_t b) {
  uint32_t r = 0;
  for (int i = 0; i < 4; i++)
    if (((a >> (8 * i)) & 0xFF) == ((b >> (8 * i)) & 0xFF)) r |= 0xFFu << (8 * i);
  return r;
}
#endif
