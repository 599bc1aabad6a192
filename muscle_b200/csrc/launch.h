// launch.h -- host-visible launch surface of the kernel translation units.
#pragma once
#include "common.cuh"

#define MB_MAX_C 16
// k_posterior_sm (state in shared memory, runtime columns-per-lane)
// op 0: launch, 1: occupancy (blocks/SM) -> *out, 2: static smem struct bytes -> *out
bool mb_post_sm_dispatch(bool mega, int op, dim3 grid, size_t smem, cudaStream_t st, const PostParams *P, int *out);
