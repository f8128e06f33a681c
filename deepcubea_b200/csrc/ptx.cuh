// ptx.cuh -- thin inline-PTX wrappers (sm_100a): TMA 1-D bulk copies, vector shared/global stores.
#pragma once
#include <stdint.h>

namespace dcb {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Make generic-proxy shared-memory writes visible to the async (TMA) proxy.  Every writer executes it.
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// TMA bulk store shared -> global (SASS: UBLKCP).  16-byte aligned addresses, size % 16 == 0.
__device__ __forceinline__ void bulk_store_s2g(void *gdst, const void *ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// Wait until all committed bulk groups have finished READING their shared-memory source.
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void sts32(uint32_t *p, uint32_t a) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(smem_u32(p)), "r"(a) : "memory"); }
__device__ __forceinline__ void sts64(uint32_t *p, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(smem_u32(p)), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t *p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(p)), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// streaming (evict-first) global stores for write-once outputs
__device__ __forceinline__ void stg_cs_v2u64(uint64_t *p, uint64_t a, uint64_t b) {
  asm volatile("st.global.cs.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void stg_cs_u32(uint32_t *p, uint32_t a) { asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(a) : "memory"); }
__device__ __forceinline__ uint32_t ldg_nc_u32(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

}  // namespace dcb
