// dcb_internal.h -- declarations shared between the translation units of libdcb_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/dcb.h"

namespace dcb {
int dcb_cuda_fail();          // records cudaGetLastError() for dcb_last_cuda_error(); returns DCB_ERR_CUDA
int dcb_check_launch();       // DCB_OK or dcb_cuda_fail() after a kernel launch
int dcb_record_cuda(cudaError_t e);

int expand_device(int env, const uint8_t *parents, const uint32_t *ids, int64_t n, uint8_t *children, uint8_t *solved,
                  uint64_t *hash, cudaStream_t st);
int lightsout_expand_device(const uint8_t *src, const uint32_t *ids, int64_t n, uint8_t *children, uint8_t *solved, uint64_t *hash,
                            cudaStream_t st);
int expand_planned_device(int env, uint8_t *arena, const uint32_t *ids, int64_t max_tiles, const uint32_t *tiles, const dcb_step_plan *plan,
                          uint8_t *node_solved, uint64_t *hash, uint32_t *node_g, uint32_t *slot_parent, cudaStream_t st);
int lightsout_expand_planned_device(uint8_t *arena, const uint32_t *ids, int64_t max_tiles, const uint32_t *tiles, const dcb_step_plan *plan,
                                    uint8_t *node_solved, uint64_t *hash, uint32_t *node_g, uint32_t *slot_parent, cudaStream_t st);
int next_state_device(int env, const uint8_t *states, int64_t n, int action, uint8_t *out, cudaStream_t st);
int is_solved_device(int env, const uint8_t *states, int64_t n, uint8_t *out, cudaStream_t st);
int hash_states_device(int env, const uint8_t *states, int64_t n, uint64_t *out, cudaStream_t st);
int nnet_input_device(int env, const uint8_t *states, int64_t n, uint8_t *out, cudaStream_t st);
}  // namespace dcb
