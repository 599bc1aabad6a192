// common.cuh -- shared device helpers and parameter blocks for libmuscle_b200 (sm_100a).
//
// Arithmetic contract (SURVEY.md section 7 hard part 1, Appendix A): every fp32 operation on the
// pair-HMM path is an explicit round-to-nearest add/mul (__fadd_rn/__fmul_rn are never fused into
// FMA) in the reference's association order, so Forward/Backward/total are bit-identical to the
// reference built with -ffp-contract=off.  The file is also compiled with -fmad=false.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/muscle_b200.h"

#define MB_FULL 0xffffffffu
#define MB_LOG_ZERO (-2e20f)          // scoretype.h:89
#define MB_MAX_K 64                    // residue classes the smem tables can hold
#define MB_CAP MB200_MAX_ROW_NNZ       // candidate entries per posterior row
#define MB_WARPS_PER_BLOCK 4            // posterior kernel: warps (pairs in flight) per CTA

struct MbHmm
	{
	float tSM, tSI, tSJ, tMM, tMI, tMJ, tII, tIM, tJJ, tJM;   // hmmscores.h:1-13
	float minScore;                                            // logf(0.01f) from the host
	int   K, KS;                                               // classes, row stride of matchT
	int   pad;                                                 // class used for out-of-range residues
	};

// LOGEXP1 pieces (scoretype.h:95-105), {c3,c2,c1,c0} for x<=1, <=2.5, <=4.5, else.
__constant__ float4 c_logexp1[4] =
	{
	{ -0.009350833524763f, 0.130659527668286f, 0.498799810682272f, 0.693203116424741f },
	{ -0.014532321752540f, 0.139942324101744f, 0.495635523139337f, 0.692140569840976f },
	{ -0.004605031767994f, 0.063427417320019f, 0.695956496475118f, 0.514272634594009f },
	{ -0.000458661602210f, 0.009695946122598f, 0.930734667215156f, 0.168037164329057f },
	};

#define ADD(a, b) __fadd_rn((a), (b))
#define MUL(a, b) __fmul_rn((a), (b))

// LOG_ADD (scoretype.h:107-124), branch-free and bit-identical:
//  * |x-y| == (larger - smaller) exactly (round-to-nearest is sign symmetric);
//  * the reference's `smaller == LOG_ZERO` test is subsumed: if only the smaller is LOG_ZERO the gap
//    is ~2e20 >= 7.5 and the larger is returned; if both are LOG_ZERO the gap is 0 and
//    LOGEXP1(0)+LOG_ZERO rounds back to LOG_ZERO, the value the reference returns;
//  * the piece is chosen by index into a 4-entry float4 table in shared memory (one LDS.128,
//    conflict free: equal indexes broadcast, different indexes hit different banks).
struct LogAdd
	{
	const float4 *coef;      // shared memory
	__device__ __forceinline__ float operator()(float x, float y) const
		{
		const float d = fabsf(__fsub_rn(x, y));
		const float lo = fminf(x, y);
		const float hi = fmaxf(x, y);
		// piece index by a 2-step binary search (FSETP, FSEL, FSETP, SEL, predicated add)
		const bool p2 = d > 2.5f;
		const float thr = p2 ? 4.5f : 1.0f;
		const float4 *a = p2 ? coef + 2 : coef;
		if (d > thr)
			a += 1;
		const float4 c = *a;
		float p = ADD(MUL(c.x, d), c.y);
		p = ADD(MUL(p, d), c.z);
		p = ADD(MUL(p, d), c.w);
		const float r = ADD(p, lo);
		return d >= 7.5f ? hi : r;
		}
	};

struct PostParams
	{
	MbHmm h;
	const float   *matchT;          // K*KS
	const float   *insT;            // K
	const uint8_t *codes;           // residue classes, all sequences back to back
	const uint64_t *seqoff;
	const uint32_t *seqlen;
	const uint32_t *px, *py;        // store pair -> sequence ids
	const uint32_t *order;          // work list (store pair ids), longest first
	uint32_t nwork;
	uint32_t *counter;              // work-stealing cursor for this launch
	// per-warp scratch
	float       *fm;      size_t fm_stride;      // floats
	float4      *edge;    size_t edge_stride;    // float4, two ping-pong halves
	mb200_entry *rows;    size_t rows_stride;    // entries
	uint8_t     *rowcnt;  size_t rowcnt_stride;  // bytes
	uint32_t     lxmax;                          // rows the scratch was sized for
	uint32_t     fm_rows;                        // (lxmax+33): steps per strip slot in fm
	uint32_t     cmax;                           // k_posterior_sm: columns per lane the smem state is sized for
	// outputs
	uint32_t       *rowoff;         // concatenated, pair k at rowbase[k], LX+1 entries
	const uint64_t *rowbase;
	mb200_entry    *entries;        // bump-allocated
	uint64_t        ent_cap;
	unsigned long long *ent_cursor;
	uint64_t       *entbase;        // per store pair
	uint32_t       *nnz;            // per store pair
	float          *ea;             // per store pair
	int            *err;            // sticky error code
	// optional dense dumps for one pair (parity surface, mb200_calc_post_dense)
	float *dbg_fwd, *dbg_bwd, *dbg_post, *dbg_total;
	};
