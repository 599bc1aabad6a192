// post_sm.cu -- launch surface of k_posterior_sm (post_kernel_sm.cuh)
#include "post_kernel_sm.cuh"
#include "launch.h"

template <bool MEGA>
static void dispatch(int op, dim3 grid, size_t smem, cudaStream_t st, const PostParams *P, int *out)
	{
	if (op == 0)
		{
		cudaFuncSetAttribute(k_posterior_sm<MEGA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
		k_posterior_sm<MEGA><<<grid, 32*MB_WARPS_PER_BLOCK, smem, st>>>(*P);
		}
	else if (op == 1)
		{
		cudaFuncSetAttribute(k_posterior_sm<MEGA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, k_posterior_sm<MEGA>, 32*MB_WARPS_PER_BLOCK, smem);
		}
	else
		*out = (int) sizeof(PostSmemHdr);
	}

bool mb_post_sm_dispatch(bool mega, int op, dim3 grid, size_t smem, cudaStream_t st, const PostParams *P, int *out)
	{
	if (mega)
		dispatch<true>(op, grid, smem, st, P, out);
	else
		dispatch<false>(op, grid, smem, st, P, out);
	return true;
	}
