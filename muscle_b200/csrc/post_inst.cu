// post_inst.cu -- explicit instantiations of k_posterior<C> for C in [MB_INST_LO, MB_INST_HI].
// Compiled several times with different ranges so that the 16 variants build in parallel.
#include "post_kernel.cuh"
#include "launch.h"

#ifndef MB_INST_LO
#error "define MB_INST_LO / MB_INST_HI / MB_INST_GROUP"
#endif

template <int C>
static bool dispatch(int want, int op, dim3 grid, size_t smem, cudaStream_t st, const PostParams *P, int *out)
	{
	if (want == C)
		{
		if (op == 0)
			{
			cudaFuncSetAttribute(k_posterior<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
			k_posterior<C><<<grid, 32*MB_WARPS_PER_BLOCK, smem, st>>>(*P);
			}
		else if (op == 1)
			{
			cudaFuncSetAttribute(k_posterior<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
			cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, k_posterior<C>, 32*MB_WARPS_PER_BLOCK, smem);
			}
		else
			*out = (int) sizeof(PostSmem<C>);
		return true;
		}
	if constexpr (C < MB_INST_HI)
		return dispatch<C + 1>(want, op, grid, smem, st, P, out);
	return false;
	}

#define MB_CAT2(a, b) a##b
#define MB_CAT(a, b) MB_CAT2(a, b)

bool MB_CAT(mb_post_dispatch_g, MB_INST_GROUP)(int C, int op, dim3 grid, size_t smem, cudaStream_t st,
  const PostParams *P, int *out)
	{
	if (C < MB_INST_LO || C > MB_INST_HI)
		return false;
	return dispatch<MB_INST_LO>(C, op, grid, smem, st, P, out);
	}
