// launch.h -- host-visible launch surface of the kernel translation units.
#pragma once
#include "common.cuh"

#define MB_MAX_C 16
// op 0: launch, 1: occupancy (blocks/SM) -> *out, 2: static smem struct bytes -> *out
bool mb_post_dispatch_g0(int C, int op, dim3 grid, size_t smem, cudaStream_t st, const PostParams *P, int *out);
bool mb_post_dispatch_g1(int C, int op, dim3 grid, size_t smem, cudaStream_t st, const PostParams *P, int *out);
bool mb_post_dispatch_g2(int C, int op, dim3 grid, size_t smem, cudaStream_t st, const PostParams *P, int *out);
bool mb_post_dispatch_g3(int C, int op, dim3 grid, size_t smem, cudaStream_t st, const PostParams *P, int *out);

static inline bool mb_post_dispatch(int C, int op, dim3 grid, size_t smem, cudaStream_t st, const PostParams *P, int *out)
	{
	return mb_post_dispatch_g0(C, op, grid, smem, st, P, out) || mb_post_dispatch_g1(C, op, grid, smem, st, P, out)
	  || mb_post_dispatch_g2(C, op, grid, smem, st, P, out) || mb_post_dispatch_g3(C, op, grid, smem, st, P, out);
	}

// k_posterior_sm (state in shared memory, runtime columns-per-lane): op as above
bool mb_post_sm_dispatch(int op, dim3 grid, size_t smem, cudaStream_t st, const PostParams *P, int *out);
