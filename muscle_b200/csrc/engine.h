// engine.h -- context layout shared by the host translation units of libmuscle_b200.so.
#pragma once
#include <vector>
#include "launch.h"

struct DevBuf
	{
	void  *p = nullptr;
	size_t cap = 0;
	int  ensure(size_t bytes);     // grow-only; 0 on success
	void release();
	};

struct PostDebug { float *fwd, *bwd, *post, *total; };

struct mb200_ctx
	{
	int device = 0;
	cudaDeviceProp prop;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
	char err[512];
	mb200_stats stats = {};

	// HMM
	bool have_hmm = false;
	MbHmm hmm = {};
	uint8_t byte2class[256] = {};
	DevBuf d_matchT, d_insT;

	// sequences
	uint32_t nseq = 0;
	std::vector<uint8_t>  h_bytes;
	std::vector<uint64_t> h_off;
	std::vector<uint32_t> h_len;
	DevBuf d_codes, d_seqoff, d_seqlen;
	// Mega feature-profile emissions (mb200_set_seqs_mega): d_codes then holds 8 letters per position
	bool mega = false;
	uint32_t mega_nf = 0, mega_tsize = 0, mega_base[8] = {}, mega_alpha[8] = {};
	DevBuf d_megaT, d_insP;

	// the store: sparse posteriors of the listed pairs (replaces MPCFlat::m_SparsePosts1/2)
	bool store_valid = false;
	bool store_allpairs = false;      // holds all N(N-1)/2 pairs in reference order
	bool ea_allpairs = false;         // d_ea holds the EA of all pairs (computed on this device)
	uint32_t store_p_lo = 0, store_p_hi = 0;
	bool store_packed = false;        // entries are packed in store order (entbase ascending, no holes)
	bool store_tr_valid = false;      // transposed orientation + permutation built
	bool tr_values_stale = false;     // forward values changed since the transposed copy was refreshed
	bool store_masks_valid = false;   // column bit masks of both orientations built (relax.cu)
	std::vector<uint64_t> h_tr_rowbase;
	uint64_t store_nnz = 0;
	uint32_t nnz_per_row_cap = 12;    // entry pool sizing guess (retry with the exact count on overflow)
	std::vector<uint32_t> h_px, h_py;
	std::vector<uint64_t> h_rowbase;
	std::vector<uint32_t> h_nnz;
	std::vector<uint64_t> h_entbase;
	bool h_index_valid = false;
	DevBuf d_px, d_py, d_order;
	DevBuf d_rowoff, d_rowbase, d_entries, d_cursor, d_entbase, d_nnz, d_ea;     // d_cursor: control block, see prepare_plan
	DevBuf d_entries2;                // second value buffer for the Jacobi update
	DevBuf d_pack_off, d_pack_ent;    // packed image for exchange/export
	uint64_t xchg_n_offsets = 0, xchg_n_entries = 0;      // sizes announced by mb200_store_exchange_begin
	// transposed orientation of every pair (rows = positions of Y) for the relax kernel
	DevBuf d_tr_rowoff, d_tr_rowbase, d_tr_entries, d_tr_entbase, d_tr_perm;
	// column bit masks of every sparse row, both orientations (relax.cu): hdr = {first word slot,
	// w0 | nw << 16} per row, words = {mask of columns 32w..32w+31, index of the first such entry}
	DevBuf d_mk_hdr, d_mk_words, d_tr_mk_hdr, d_tr_mk_words;
	// relax work order (pairs of [p_lo,p_hi) in 2-D tile order) and the range it was built for
	DevBuf d_relax_order;
	uint32_t relax_order_n = 0, relax_order_lo = 0, relax_order_hi = 0;
	DevBuf d_tmp, d_tmp2;
	// device-resident MSAs (align.cu): position -> column of every residue
	// (same layout as d_codes) and, on the host, the column count of the MSA each sequence is in
	DevBuf d_p2c, d_join;             // d_join: per-join scratch (maps, Post, traceback, path)
	DevBuf d_stage;                   // BuildPost staging slots
	std::vector<uint32_t> h_msa_cols;
	bool msa_valid = false;
	uint32_t *h_pinned = nullptr;     // 64 words of pinned host memory for small read-backs

	// cached launch plan of the posterior stage (depends only on the pair list)
	bool plan_valid = false, plan_is_allpairs = false;
	int plan_force_c = 0;
	uint64_t plan_cells = 0, plan_est_entries = 0, last_used_entries = 0;
	std::vector<std::vector<uint32_t>> plan_bins;
	std::vector<uint32_t> plan_bin_start, plan_lxmax, plan_lymax;

	// per-warp scratch of the posterior kernel
	DevBuf d_fm[MB_MAX_C + 1], d_edge[MB_MAX_C + 1], d_rows[MB_MAX_C + 1], d_rowcnt[MB_MAX_C + 1];   // per bin
	static constexpr int kStreams = 4;          // column-width bins run concurrently to overlap their tails
	cudaStream_t aux[kStreams] = {};
	cudaEvent_t aux_done[kStreams] = {};
	DevBuf d_dbg;
	int debug_force_c = 0;
	int occ_cache[MB_MAX_C + 1] = {}, occ_cache_k = -1;     // resident CTAs/SM of k_posterior_sm per size class
	};

int mb_fail(mb200_ctx *ctx, int code, const char *fmt, ...);
int mb_store_pack_inplace(mb200_ctx *ctx);
int mb_store_build_transposed(mb200_ctx *ctx);
int mb_store_refresh_transposed(mb200_ctx *ctx);
int mb_store_build_masks(mb200_ctx *ctx);
void mb_allpairs_list(uint32_t n, uint32_t p_lo, uint32_t p_hi, std::vector<uint32_t> &px, std::vector<uint32_t> &py);
