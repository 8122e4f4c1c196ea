// libvoroffset_b200.so - host side of the C ABI declared in include/voroffset_b200.h.
// CUDA runtime only (no torch, no CPU fallback). Kernels are in kernels.cuh / scan.cuh.
#include "voroffset_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <functional>
#include <limits>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

#include "kernels.cuh"
#include "pass1_tile.cuh"
#include "scan.cuh"
#include "dexelize.cuh"

using namespace vo;

// ---------------------------------------------------------------------------------------------------
// objects
// ---------------------------------------------------------------------------------------------------
struct vo_ctx {
	int device = 0;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
	std::string err;
	uint64_t launches = 0;
	// small persistent device scratch: [0] mid-pool cursor, [1] stage-pool cursor, [2] redo count, [3] big-tile count
	unsigned long long *d_ctr = nullptr;
	// caller-visible timing marks (vo_mark / vo_elapsed_ms) and per-kernel profile of the last dilation
	cudaEvent_t mark[8] = {};
	cudaEvent_t kev[4] = {};          // [0,1] around k_pass1<CAP_FAST>, [2,3] around k_pass2<CAP_FAST>
	bool kev_valid[2] = {false, false};
	void *table_cache = nullptr;      // TableCache*: cap tables of the last radius, kept on the device
	// pool sizes the recent calls needed (+25 %), forgotten slowly: opening / closing alternate between a dilation
	// that needs next to nothing and an erosion that needs millions of entries - a hint that followed the last call
	// only would make every second pass run twice
	unsigned long long pool_hint = 0;  // mid-pool entries (pass 1)
	unsigned long long stage_hint = 0; // staging-pool entries (staged gathers)
	unsigned long long *dbg_tiles = nullptr;   // vo_set_option("tile_debug", "<device pointer>"): per-tile statistics
	uint32_t *ovf = nullptr;          // spill area of the tile kernel's survivor lists (pass1_tile.cuh), allocated on first use
	uint64_t out_hint = 0;            // intervals of the last pipelined result (+12 %): sizes the pinned span buffer
	bool no_pipeline = false;         // vo_set_option("pipeline", "off")
	// last resort for lists beyond CAP_BIG (with_huge_lists): the redo launches run with CAP_HUGE and lists in `huge_scratch`
	bool huge = false, list_overflow = false;
	double2 *huge_scratch = nullptr;
	bool tile_order = true;           // vo_set_option("tile_order", "off"): pass-1 tiles in row-major order instead of expensive first
	int multi_warps = 0;              // vo_set_option("multi_warps", "N"): warps per CTA of the tile kernel's multi-interval launches (0: chosen by
	                                  // TilePlan::init - fewer than 16 for deep columns, whose sorted-list unions keep their lists in local memory:
	                                  // 16 warps x 32 lanes x ~26 intervals outgrow the L1 and every insertion waits for the L2)
	int p2_mode = -1;                 // vo_set_option("pass2_union", "auto" | "registers" | "lists"): running union of pass 2 on shallow input, see pass2()
	int p2_list_calls = 0;            // > 0: pass 2 keeps its list-capable union for this many more calls
	int gen_mode = -1;                // vo_set_option("tile_general", "auto" | "redo" | "inline"): -1 as described below, 0 never inline, 1 always
	bool last_lean1 = false;          // the last tile launch set started with the 20-warp variant (note_pass1_redo)
	int gen_inline_calls = 0;         // > 0: the first tile launch keeps its inline sorted-list union for this many more calls (k_pass1_tile<..., GEN>:
	                                  // a pass 1 that handed slots to the redo launch probably met "complex" classes; 0: the 20-warp variant without it)
	double pooled_per_column = -1;    // mid-pool entries per column the last tile-kernel pass 1 with multi-interval tiles needed (-1: none yet)
	int tile_ctas = 1;                // vo_set_option("tile_ctas", "N"): CTAs per SM of the pass-1 tile kernel
	int tile_dbuf = -1;               // vo_set_option("tile_dbuf", "auto" | "on" | "off"): double-buffered candidate staging of the tile kernel
	int cand_order = -1;              // vo_set_option("cand_order", "auto" | "column" | "layer"): order in which the lean multi-interval launches walk a tile's candidates (Pass1TileArgs::layer_major)
	int tile_lean = -1;               // vo_set_option("tile_lean", "auto" | "on" | "off"): candidates of the list launches left in global memory
	int band_split = 2;               // vo_set_option("band_split", "N"): a band's pass-1 launch set takes 1/N of the SMs (host-buffer pipeline)
	int pipe_warps = 64;              // vo_set_option("pipe_warps", "N"): warps per tile-kernel CTA in the host-buffer pipeline (default: as many as fit)
	int band_free = 0;                // vo_set_option("band_free", "N"): SMs no pass-1 launch set of the pipeline takes (room for its small kernels and pass 2)
	int pipe_bands = 8;               // vo_set_option("bands", "N"): row bands of the pipelined host-buffer path
	std::vector<int> pipe_wts;        // vo_set_option("band_weights", "1,2,3,4,..."): relative heights of the bands (empty: equal bands)
	bool pipe_first_full = true;      // vo_set_option("pipe_first_full", "on"): the first band's tile launches take every SM
	bool pipe_interleave = false;     // vo_set_option("pipe_interleave", "on"): enqueue order pass 1 (b + 1), second half (b), ...
	bool pipe_lean = true;            // vo_set_option("pipe_lean", "on"): the bands of the host-buffer call leave out the launches that are normally idle
	                                  // (the tile kernel's list launches, the redo launches of both passes); a call that did need one is
	                                  // repeated on the plain path and the context remembers which (pipe_lists, pipe_redo)
	bool pipe_lists = false, pipe_redo = false;
	unsigned long long pipe_dry = 0;             // (-DVO_KTRACE builds) vo_set_option("pipe_dry", "<intervals>"): the host-buffer call without its kernels - copies, events
	                                  // and host round trips only, downloads sized from the previous result: the floor its transfer pattern sets
	int copy_align = 256;             // vo_set_option("copy_align", "N"): the band copies of the host-buffer call start and end on N-byte boundaries
	                                  // (0: exactly the band's bytes). The copy engines run a copy whose ends are only 16-byte aligned a fifth
	                                  // slower when both directions are busy (profiles/r2ah_pcie_probe6.txt); the few bytes of the
	                                  // neighbouring band that travel along are the same bytes, or are overwritten by that band's own copy
	bool copy_batch = true;           // vo_set_option("copy_batch", "on"): the two copies of a band and direction (offsets, spans) as ONE
	                                  // cudaMemcpyBatchAsync: the copy engine takes them without the gap between two stream operations
	                                  // (bare pattern, 8 / 16 bands: 1.46 -> 1.39 / 1.64 -> 1.46 ms, scripts/probes/pcie_probe7.cu)
	bool pipe_ahead = false;          // vo_set_option("pipe_ahead", "on"): a band's offsets download is enqueued before the host knows the band's total
	                                  // ("off", the default since copy_batch: offsets and spans of a band leave together as one batch once the
	                                  // total is known - 1.98 against 2.01 ms)
	bool pipe_order_one = true;       // vo_set_option("pipe_order_one", "on"): a band's tile order by one CTA in one launch
	int pipe_warps0 = 0;              // vo_set_option("pipe_warps0", "N"): warps per tile-kernel CTA for the FIRST band only (0: as the others) - few
	                                  // warps per SM run their tiles faster, which is what the time to the first result needs
	int pipe_quota = 0;               // vo_set_option("pipe_quota", "N"): tiles per warp of the pipeline's first tile launch (0: persistent CTAs)
	int pipe_ctas = 0;                // vo_set_option("pipe_ctas", "N"): CTAs per SM of the pipeline's tile launches (0: as tile_ctas)
	bool pipe_mid = false;            // vo_set_option("pipe_mid", "on"): thresholds and tile order of every band on a stream of their own, one
	                                  // priority level above the tile launches
	cudaStream_t s_mid = nullptr;
	size_t ovf_areas = 0;             // per-warp spill areas per bank of `ovf`
	bool slab_overlap = true;         // vo_set_option("slab", "overlap" | "serial"): pass 1 of the halo-independent rows while the halos travel
	int slab_reserve = 8;             // vo_set_option("slab_reserve", "N"): SMs the interior launch of a slab step leaves to the NCCL kernels
	cudaStream_t s_in = nullptr, s_out = nullptr;   // copy streams of the pipelined host-buffer path
	cudaStream_t s_p[2] = {nullptr, nullptr};       // its two pass-1 streams (consecutive bands overlap)
	cudaStream_t s_hi[3] = {nullptr, nullptr, nullptr};   // its second-half streams (highest priority): two for alternating bands, one for the pass-1 redo launches
	cudaStream_t s_ctl = nullptr;                   // its control stream (band totals -> host)
	cudaStream_t s_out2 = nullptr;                  // the stream of its SM-driven span downloads (k_copy_out)
	int copy_out_ctas = 0;                          // vo_set_option("copy_out", "N"): CTAs of k_copy_out; 0 (default): the spans leave through the copy
	                                                // engine, which needs a host round trip per band for their size
	std::vector<cudaEvent_t> pipe_ev;               // its (reused) events
	// single-pass scan + compaction (k_scan_compact): tile state words, ticket counter (device) and their host mirrors
	unsigned long long *scan_state = nullptr;   // [scan_cap] state words, then the ticket counter
	size_t scan_cap = 0;
	unsigned long long scan_ticket = 0;         // tickets handed out so far
	uint32_t scan_epoch = 0;                    // epoch of the last call (1 .. 2^22 - 1)
	unsigned long long out_cap_hint = 0;        // intervals the recent staged results needed (+12 %): sizes the spans buffer up front
	bool fused_scan = true;                     // vo_set_option("scan", "fused" | "classic")
	unsigned long long last_ctr[18] = {};
	bool staged_ctr_clean = false;    // pass 1 has just zeroed every counter: the staged gather that follows need not zero its own again   // the counters as the last read_counters saw them (a deferred pass 1 is judged after pass 2)
	bool force_tile_pass1 = false;
	bool force_simple_pass1 = false;  // vo_set_option("pass1", "simple"): always use the one-thread-per-(x,y,j) kernel
	// Large scratch blocks (>= 1 MiB) released by dfree are kept WHOLE and handed to the next request they fit
	// (dalloc). The driver's stream-ordered pool splits a released block for smaller requests; an erosion / closing
	// alternates between requests of very different sizes (a 2.4 GB mid volume, pools sized by decaying hints) and
	// every few calls the pool had to map fresh memory: 5 ms operations took 15-35 ms now and then.
	std::unordered_map<void *, size_t> big_live;            // blocks handed out -> capacity
	std::vector<std::pair<void *, size_t>> big_free;        // released blocks, oldest first
	size_t big_free_bytes = 0, big_free_limit = (size_t)48 << 30;   // vo_create sets the limit to a third of the device memory
	bool block_cache = true;          // vo_set_option("block_cache", "off")
	int redo_recent = 0;              // > 0: a recent call needed a redo launch (lists beyond the fast capacities) - the next calls
	                                  // enqueue the redo launches up front again instead of finding out at the end
	uint32_t max_epoch = 0;           // tags the "largest list" word ([15] of d_ctr) of a staged gather
	int erosion_mode = 0;             // vo_set_option("erosion", "auto" | "dual" | "general"): 0 = dual form where the input qualifies
	                                  // (erode_dual), 1 = dual form or VO_ERR_ARG (tests), 2 = always complement - dilate - complement
	uint64_t dual_erosions = 0;       // erosions that ran in dual form
};

struct vo_dvol {
	int nx = 0, ny = 0;
	uint64_t nspans = 0;
	uint32_t *off = nullptr;
	double2 *spans = nullptr;
	long long max_cnt = -1;         // upper bound on the intervals of any column where one is known (-1: unknown): a volume known to
	                                // hold at most one per column skips the multi-interval launches of pass 1
	mutable int dual_state = 0;     // erode_dual: 0 = not tried, 1 = qualified, 2 = did not (several intervals in a column, data at the bounds)
};

struct vo_dmid {
	int nx = 0, ny = 0, J = 0;
	double R = 0;
	double2 *slots = nullptr;
	double2 *pool = nullptr;
	uint16_t *flags = nullptr;      // [2][ny*nx]: class window (lo | hi << 8) needed by the consumer rows above / below each mid column
	unsigned long long *tilemask = nullptr;   // [2][ny * ceil(nx / P1_TX)]: OR of the windows per pass-1 tile
	uint64_t pool_cap = 0, pool_used = 0;
	bool shallow = false;           // pass 1 ran on input known to hold at most one interval per column: pass 2 with more CTAs per SM (P2_SHALLOW)
	bool redo_skipped = false;      // ... and without its redo launch: any entry in the redo list means "repeat"
	bool deferred = false;          // pass 1 returned without reading its counters: the caller checks them after pass 2
	unsigned int redo_cap = 0;      // ... against these
};
static_assert(P1_TX == P2_TX, "pass 2 reads the tile masks of pass 1: same tile width");

struct vo_slab;   // a y-slab dilation in flight (defined with its helpers below)

namespace {

int fail(vo_ctx *ctx, int code, const std::string &msg)
{
	if (ctx) ctx->err = msg;
	return code;
}

#define VO_CUDA(call)                                                                                   \
	do {                                                                                                \
		cudaError_t e_ = (call);                                                                        \
		if (e_ != cudaSuccess) {                                                                        \
			cudaGetLastError();                                                                         \
			return fail(ctx, e_ == cudaErrorMemoryAllocation ? VO_ERR_NOMEM : VO_ERR_CUDA,              \
			            std::string(#call) + ": " + cudaGetErrorString(e_));                            \
		}                                                                                               \
	} while (0)

#define VO_TRY(expr)                                                                                    \
	do {                                                                                                \
		int rc_ = (expr);                                                                               \
		if (rc_ != VO_OK) return rc_;                                                                   \
	} while (0)

inline unsigned long long next_hint(unsigned long long hint, unsigned long long used)
{
	return std::max(used + used / 4, hint - hint / 8);
}

inline unsigned int blocks_for(unsigned long long n, int threads)
{
	unsigned long long b = (n + threads - 1) / threads;
	return (unsigned int)(b ? b : 1);
}

void release_big_free(vo_ctx *ctx, size_t keep_bytes)
{
	size_t i = 0;
	while (ctx->big_free_bytes > keep_bytes && i < ctx->big_free.size()) {
		cudaFreeAsync(ctx->big_free[i].first, ctx->stream);
		ctx->big_free_bytes -= ctx->big_free[i].second;
		++i;
	}
	ctx->big_free.erase(ctx->big_free.begin(), ctx->big_free.begin() + i);
}

template <typename T> int dalloc(vo_ctx *ctx, T **p, unsigned long long count)
{
	*p = nullptr;
	size_t bytes = (size_t)std::max<unsigned long long>(count, 1ull) * sizeof(T);
	const bool big = bytes >= (1u << 20) && ctx->block_cache;
	// Large requests are rounded up to eight size classes per octave: a block released by the previous call then fits
	// even when the grid differs by a border column (erosion) or the result by a few intervals.
	if (bytes >= (1u << 20)) {
		size_t q = (size_t)1 << 17;
		while ((q << 4) <= bytes) q <<= 1;
		bytes = (bytes + q - 1) / q * q;
	}
	if (big) {
		// smallest released block that holds the request without wasting more than half of itself. Reuse is in stream
		// order like cudaMallocAsync after cudaFreeAsync: every block is allocated, first used and released on
		// ctx->stream (other streams join it through events before a release).
		int best = -1;
		for (int i = 0; i < (int)ctx->big_free.size(); ++i) {
			const size_t cap = ctx->big_free[i].second;
			if (cap >= bytes && cap / 2 <= bytes && (best < 0 || cap < ctx->big_free[best].second)) best = i;
		}
		if (best >= 0) {
			*p = (T *)ctx->big_free[best].first;
			ctx->big_live[*p] = ctx->big_free[best].second;
			ctx->big_free_bytes -= ctx->big_free[best].second;
			ctx->big_free.erase(ctx->big_free.begin() + best);
			return VO_OK;
		}
	}
	cudaError_t e = cudaMallocAsync((void **)p, bytes, ctx->stream);
	if (e == cudaErrorMemoryAllocation && !ctx->big_free.empty()) {
		cudaGetLastError();
		release_big_free(ctx, 0);
		cudaStreamSynchronize(ctx->stream);
		e = cudaMallocAsync((void **)p, bytes, ctx->stream);
	}
	VO_CUDA(e);
	if (big) ctx->big_live[*p] = bytes;
	return VO_OK;
}

template <typename T> void dfree(vo_ctx *ctx, T *p)
{
	if (!p) return;
	auto it = ctx->big_live.find((void *)p);
	if (it == ctx->big_live.end()) { cudaFreeAsync((void *)p, ctx->stream); return; }
	ctx->big_free.emplace_back(it->first, it->second);
	ctx->big_free_bytes += it->second;
	ctx->big_live.erase(it);
	if (ctx->big_free_bytes > ctx->big_free_limit) release_big_free(ctx, ctx->big_free_limit);
}

// RAII for temporaries allocated on the context stream
template <typename T> struct Tmp {
	vo_ctx *ctx;
	T *p = nullptr;
	explicit Tmp(vo_ctx *c) : ctx(c) {}
	~Tmp() { dfree(ctx, p); }
	Tmp(const Tmp &) = delete;
	Tmp &operator=(const Tmp &) = delete;
};

// ---- pinned host blocks handed to the caller, recycled through vo_free -------------------------------
std::mutex g_host_mu;
std::unordered_map<void *, size_t> g_host_live;            // ptr -> capacity
std::vector<std::pair<void *, size_t>> g_host_cache;       // released blocks kept for reuse
int g_live_contexts = 0;                                   // the cache is emptied when the last context goes (guarded by g_host_mu)

void *host_block(size_t bytes)
{
	bytes = std::max<size_t>(bytes, 64);
	{
		std::lock_guard<std::mutex> lk(g_host_mu);
		int best = -1;
		for (int i = 0; i < (int)g_host_cache.size(); ++i)
			if (g_host_cache[i].second >= bytes && g_host_cache[i].second <= 2 * bytes + (1u << 20) &&
			    (best < 0 || g_host_cache[i].second < g_host_cache[best].second))
				best = i;
		if (best >= 0) {
			auto blk = g_host_cache[best];
			g_host_cache.erase(g_host_cache.begin() + best);
			g_host_live[blk.first] = blk.second;
			return blk.first;
		}
	}
	void *p = nullptr;
	if (cudaMallocHost(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
	std::lock_guard<std::mutex> lk(g_host_mu);
	g_host_live[p] = bytes;
	return p;
}

// ---- cap tables (the reference's exact fp64 operation order, computed once on the host) --------------
struct Tables {
	int J = 0;
	std::vector<double> H;     // 'ours'  [j = |dy|][|dx|]
	std::vector<int> reach;    // 'ours'  floor(r1(j))
	std::vector<double> HB;    // 'brute_force' [|dy|][|dx|]
};

// 'ours': r1 = R (dy == 0) or sqrt(R*R - dy*dy) (Voronoi2D.cpp:704-716); a piece lives while
// |dx| <= floor(r1) (SeparatePower2D.cpp:241,266); h = sqrt(r1*r1 - dx*dx) (SeparatePower2D.cpp:314).
// 'brute_force': dx^2 + dy^2 <= R*R, dz = sqrt(R*R - dx^2 - dy^2) (VoronoiBruteForce.cpp:48-50).
Tables make_tables(double R)
{
	Tables t;
	t.J = (int)std::floor(R);
	const int n = t.J + 1;
	t.H.assign((size_t)n * n, -1.0);
	t.HB.assign((size_t)n * n, -1.0);
	t.reach.assign(n, 0);
	for (int dy = 0; dy < n; ++dy) {
		const double dyd = (double)dy;
		volatile double r1 = (dy == 0) ? R : std::sqrt(R * R - dyd * dyd);
		const int reach = (int)std::floor(r1);
		t.reach[dy] = std::min(reach, t.J);
		for (int dx = 0; dx < n; ++dx) {
			const double dxd = (double)dx;
			if (dx <= reach) {
				volatile double a = r1 * r1;
				volatile double b = dxd * dxd;
				volatile double d = a - b;
				t.H[(size_t)dy * n + dx] = std::sqrt(d);
			}
			volatile double p2 = std::pow(dxd, 2.0) + std::pow(dyd, 2.0);
			volatile double rr = R * R;
			if (p2 <= rr) {
				volatile double q = rr - std::pow(dxd, 2.0);
				volatile double q2 = q - std::pow(dyd, 2.0);
				t.HB[(size_t)dy * n + dx] = std::sqrt(q2);
			}
		}
	}
	return t;
}

int check_dims(vo_ctx *ctx, int nx, int ny)
{
	if (nx < 0 || ny < 0) return fail(ctx, VO_ERR_ARG, "negative grid size");
	if ((unsigned long long)nx * (unsigned long long)ny >= (1ull << 32) - 8) return fail(ctx, VO_ERR_OVERFLOW, "grid too large for uint32 CSR offsets");
	return VO_OK;
}

// counts -> offsets; returns the grand total (checked against the uint32 range)
int scan_counts(vo_ctx *ctx, const uint32_t *cnt, unsigned long long n, uint32_t *off, unsigned long long *total)
{
	if (n == 0) {
		VO_CUDA(cudaMemsetAsync(off, 0, sizeof(uint32_t), ctx->stream));
		*total = 0;
		return VO_OK;
	}
	const unsigned int ntiles = blocks_for(n, SCAN_TILE);
	Tmp<unsigned long long> sums(ctx);
	VO_TRY(dalloc(ctx, &sums.p, (unsigned long long)ntiles + 1));
	k_scan_reduce<<<ntiles, SCAN_THREADS, 0, ctx->stream>>>(cnt, n, sums.p);
	k_scan_tiles<<<1, 1024, 0, ctx->stream>>>(sums.p, ntiles);
	k_scan_apply<<<ntiles, SCAN_THREADS, 0, ctx->stream>>>(cnt, n, sums.p, off);
	ctx->launches += 3;
	VO_CUDA(cudaGetLastError());
	unsigned long long tot = 0;
	VO_CUDA(cudaMemcpyAsync(&tot, sums.p + ntiles, sizeof(tot), cudaMemcpyDeviceToHost, ctx->stream));
	VO_CUDA(cudaStreamSynchronize(ctx->stream));
	if (tot >= (1ull << 32)) return fail(ctx, VO_ERR_OVERFLOW, "result has more than 2^32-1 intervals");
	*total = tot;
	return VO_OK;
}

int new_dvol(vo_ctx *ctx, int nx, int ny, vo_dvol **out)
{
	vo_dvol *v = new (std::nothrow) vo_dvol();
	if (!v) return fail(ctx, VO_ERR_NOMEM, "out of host memory");
	v->nx = nx;
	v->ny = ny;
	int rc = dalloc(ctx, &v->off, (unsigned long long)nx * ny + 1);
	if (rc) { delete v; return rc; }
	*out = v;
	return VO_OK;
}

void free_dvol(vo_ctx *ctx, vo_dvol *v)
{
	if (!v) return;
	dfree(ctx, v->off);
	dfree(ctx, v->spans);
	delete v;
}

// Staged lists of `nlists` lists -> new device CSR volume of shape (nx, ny).
struct StageBuf {
	vo_ctx *ctx;
	Stage st{};
	explicit StageBuf(vo_ctx *c) : ctx(c) {}
	~StageBuf() { dfree(ctx, st.cnt); dfree(ctx, st.inl); dfree(ctx, st.pool); }
	int alloc(unsigned long long nlists, unsigned long long pool_cap)
	{
		VO_TRY(dalloc(ctx, &st.cnt, nlists));
		VO_TRY(dalloc(ctx, &st.inl, nlists * STAGE_INLINE));
		VO_TRY(dalloc(ctx, &st.pool, pool_cap));
		st.pool_cap = pool_cap;
		st.cursor = ctx->d_ctr + 1;
		return VO_OK;
	}
	int regrow(unsigned long long pool_cap)
	{
		dfree(ctx, st.pool);
		st.pool = nullptr;
		VO_TRY(dalloc(ctx, &st.pool, pool_cap));
		st.pool_cap = pool_cap;
		return VO_OK;
	}
};

struct RedoBuf {
	vo_ctx *ctx;
	Redo rd{};
	explicit RedoBuf(vo_ctx *c) : ctx(c) {}
	~RedoBuf() { dfree(ctx, rd.list); }
	int alloc(unsigned int cap, int counter = 2)
	{
		VO_TRY(dalloc(ctx, &rd.list, cap));
		rd.cap = cap;
		rd.count = reinterpret_cast<unsigned int *>(ctx->d_ctr + counter);
		return VO_OK;
	}
};

constexpr int NCTR = 16;  // pass 1: [0] mid-pool cursor [2] redo count [3] big-tile count [4] redo failures [5] multi-tile count
                          //         [6] [7] list cursors [10] tile cursor of the first launch; staged gathers (pass 2, ...): [1] stage-pool cursor [8] redo count [9] failures
                          //         [11] invalid input offsets seen by k_thresh (banded host-buffer call)

constexpr int NREAD = NCTR + 2;   // ... what read_counters brings back: the counters, [NCTR] erosion's "data outside the z range" flag,
                                  // [NCTR + 1] the grand total of the last fused scan + compaction (written, not accumulated)
int read_counters(vo_ctx *ctx, unsigned long long h[NREAD])
{
	VO_CUDA(cudaMemcpyAsync(h, ctx->d_ctr, NREAD * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
	VO_CUDA(cudaStreamSynchronize(ctx->stream));
	static_assert(NREAD == 18, "vo_ctx::last_ctr");
	std::memcpy(ctx->last_ctr, h, NREAD * sizeof(unsigned long long));
	return VO_OK;
}

// Generic driver for the staged gather kernels: first launch with CAP_FAST over all lists, then the
// redo launch with CAP_BIG over the lists whose running union outgrew CAP_FAST; the staging pool is
// regrown and the launch repeated if it was too small (the cursor keeps counting past the capacity).
// The whole chain - gather, redo, prefix sum - is enqueued without a host round trip; the single
// synchronisation at the end returns the counters and the grand total together.
constexpr unsigned int REDO_GRID = 148 * 4;
constexpr unsigned int HUGE_GRID = 74;      // last-resort redo launch: 74 x 128 threads x CAP_HUGE x 16 bytes = 4.6 GiB of scratch

// "a list outgrew the redo capacity": remembered, so that the primitive can be repeated with the last-resort launch
int fail_list_overflow(vo_ctx *ctx)
{
	ctx->list_overflow = true;
	return fail(ctx, VO_ERR_OVERFLOW, ctx->huge ? "a dexel list needs more than 32768 disjoint intervals in its running union"
	                                           : "a dexel list needs more than 512 disjoint intervals in its running union");
}

// Runs a primitive (a dilation: both passes, brute force, the 2D rows); if it fails because a running union outgrew
// CAP_BIG, runs it once more with the redo launches in their last-resort form (CAP_HUGE, lists in global memory: slow,
// but the reference has no such limit at all). The scratch only lives for that second run.
template <typename F>
int with_huge_lists(vo_ctx *ctx, F f)
{
	if (ctx->huge) return f();
	ctx->list_overflow = false;
	int rc = f();
	if (rc != VO_ERR_OVERFLOW || !ctx->list_overflow) return rc;
	const size_t bytes = (size_t)HUGE_GRID * 128 * CAP_HUGE * sizeof(double2);
	if (cudaMalloc((void **)&ctx->huge_scratch, bytes) != cudaSuccess) {
		cudaGetLastError();
		ctx->huge_scratch = nullptr;
		return fail(ctx, VO_ERR_NOMEM, "scratch of the last-resort launch (a dexel list needs more than 512 disjoint intervals in its running union)");
	}
	ctx->huge = true;
	ctx->err.clear();
	rc = f();
	ctx->huge = false;
	cudaStreamSynchronize(ctx->stream);
	cudaFree(ctx->huge_scratch);
	ctx->huge_scratch = nullptr;
	return rc;
}

// tile state of k_scan_compact for `ntiles` tiles: grown when needed, cleared when grown and when the epoch wraps
// (growing and the epoch wrap clear the array and wait for that: rare, and other streams may use it next)
int scan_reserve(vo_ctx *ctx, unsigned int ntiles)
{
	if (ctx->scan_cap >= ntiles) return VO_OK;
	if (ctx->scan_state) { cudaDeviceSynchronize(); cudaFree(ctx->scan_state); ctx->scan_state = nullptr; ctx->scan_cap = 0; }
	const size_t cap = std::max<size_t>(2 * (size_t)ntiles, 4096);
	VO_CUDA(cudaMalloc((void **)&ctx->scan_state, (cap + 1) * sizeof(unsigned long long)));
	VO_CUDA(cudaMemsetAsync(ctx->scan_state, 0, (cap + 1) * sizeof(unsigned long long), ctx->stream));
	VO_CUDA(cudaStreamSynchronize(ctx->stream));
	ctx->scan_cap = cap; ctx->scan_ticket = 0; ctx->scan_epoch = 0;
	return VO_OK;
}

int scan_prepare(vo_ctx *ctx, unsigned int ntiles, uint32_t *epoch, unsigned long long *ticket_base)
{
	if (ntiles == 0) { *epoch = 0; *ticket_base = ctx->scan_ticket; return VO_OK; }
	VO_TRY(scan_reserve(ctx, ntiles));
	if (++ctx->scan_epoch >= (1u << SCAN_EPOCH_BITS)) {
		VO_CUDA(cudaMemsetAsync(ctx->scan_state, 0, ctx->scan_cap * sizeof(unsigned long long), ctx->stream));
		VO_CUDA(cudaStreamSynchronize(ctx->stream));
		ctx->scan_epoch = 1;
	}
	*epoch = ctx->scan_epoch;
	*ticket_base = ctx->scan_ticket;
	ctx->scan_ticket += ntiles;
	return VO_OK;
}

// done_ev (optional): recorded behind the last kernel, before the counters travel back
template <typename Args, typename LaunchFast, typename LaunchBig>
int run_staged(vo_ctx *ctx, Args &args, unsigned long long nlists, unsigned long long pool_guess,
               LaunchFast launch_fast, LaunchBig launch_big, int nx, int ny, vo_dvol **out, cudaEvent_t done_ev = nullptr)
{
	StageBuf sb(ctx);
	RedoBuf rb(ctx);
	// (the last call's need is the best guess for repeated calls on similar data: no second gather)
	// (bounded by what the lists can need at most - CAP_BIG intervals each - not by a guess: dense volumes, ten intervals
	// per column, used to run every gather three times because the hint was cut to four per list)
	VO_TRY(sb.alloc(nlists, std::max(pool_guess, std::min(ctx->stage_hint, (unsigned long long)CAP_BIG * nlists + 65536ull))));
	const unsigned int redo_cap = (unsigned int)std::min<unsigned long long>(std::max<unsigned long long>(nlists, 1ull), 1ull << 22);
	VO_TRY(rb.alloc(redo_cap, 8));
	vo_dvol *v = nullptr;
	VO_TRY(new_dvol(ctx, nx, ny, &v));
	struct Guard { vo_ctx *c; vo_dvol *&p; ~Guard() { if (p) free_dvol(c, p); } } guard{ctx, v};
	const bool fused = ctx->fused_scan && nlists > 0;
	// (small grids: tiles of 512 lists instead of 2048, so that every SM gets some)
	const bool small_tiles = fused && blocks_for(nlists, SCAN_TILE) < 4 * 148;
	const unsigned int ntiles = blocks_for(nlists, small_tiles ? SCAN_THREADS * 2 : SCAN_TILE);
	Tmp<unsigned long long> sums(ctx);
	if (!fused) VO_TRY(dalloc(ctx, &sums.p, (unsigned long long)ntiles + 1));
	// fused: the spans buffer exists before the gather (sized from the recent results, at least one interval per list)
	// and ONE kernel writes offsets and spans; a result that outgrows the buffer is compacted again into an exact one
	unsigned long long out_cap = 0;
	if (fused) {
		out_cap = std::min<unsigned long long>(std::max(ctx->out_cap_hint, nlists + 65536ull), (1ull << 32) - 1);
		VO_TRY(dalloc(ctx, &v->spans, out_cap));
	}
	// The redo launch is normally idle (no list outgrows the fast capacity): it is left out, and the call repeats with it
	// when the counters say that a list did (and keeps launching it for the calls that follow).
	bool skip_redo = ctx->redo_recent == 0 && !ctx->huge;
	if (ctx->redo_recent > 0) --ctx->redo_recent;
	for (int attempt = 0; attempt < 4; ++attempt) {
		// only the counters of the staged gather: a pass 1 may be in flight on the same stream (pipelined path)
		if (!ctx->staged_ctr_clean) {
			VO_CUDA(cudaMemsetAsync(ctx->d_ctr + 1, 0, sizeof(unsigned long long), ctx->stream));
			VO_CUDA(cudaMemsetAsync(ctx->d_ctr + 8, 0, 2 * sizeof(unsigned long long), ctx->stream));
		}
		ctx->staged_ctr_clean = false;
		unsigned long long h[NREAD] = {0}, total = 0;
		if (nlists) {
			args.st = sb.st;
			args.redo = rb.rd;
			args.wk = Work{nullptr, nlists, nullptr, 0u, nullptr};
			launch_fast(args);
			args.wk = Work{rb.rd.list, 0ull, rb.rd.count, rb.rd.cap, reinterpret_cast<unsigned int *>(ctx->d_ctr + 9)};
			args.wk.huge = ctx->huge_scratch;                    // (ctx->huge: the launch_big of every caller launches its CAP_HUGE form)
			if (!skip_redo) launch_big(args, REDO_GRID);
			if (fused) {
				uint32_t epoch = 0;
				unsigned long long tbase = 0;
				VO_TRY(scan_prepare(ctx, ntiles, &epoch, &tbase));
				if (++ctx->max_epoch == 0) {                     // (the tag wrapped: start over from a cleared word)
					VO_CUDA(cudaMemsetAsync(ctx->d_ctr + 15, 0, sizeof(unsigned long long), ctx->stream));
					ctx->max_epoch = 1;
				}
				if (small_tiles)
					k_scan_compact<2><<<ntiles, SCAN_THREADS, 0, ctx->stream>>>(sb.st, nlists, v->off, v->spans, out_cap, ctx->scan_state,
					                                                            ctx->scan_state + ctx->scan_cap, tbase, epoch, ctx->d_ctr + NCTR + 1,
					                                                            nullptr, ctx->d_ctr + 15, ctx->max_epoch);
				else
					k_scan_compact<SCAN_ITEMS><<<ntiles, SCAN_THREADS, 0, ctx->stream>>>(sb.st, nlists, v->off, v->spans, out_cap, ctx->scan_state,
					                                                                     ctx->scan_state + ctx->scan_cap, tbase, epoch, ctx->d_ctr + NCTR + 1,
					                                                                     nullptr, ctx->d_ctr + 15, ctx->max_epoch);
				ctx->launches += skip_redo ? 2 : 3;
			} else {
				k_scan_reduce<<<ntiles, SCAN_THREADS, 0, ctx->stream>>>(sb.st.cnt, nlists, sums.p);
				k_scan_tiles<<<1, 1024, 0, ctx->stream>>>(sums.p, ntiles);
				k_scan_apply<<<ntiles, SCAN_THREADS, 0, ctx->stream>>>(sb.st.cnt, nlists, sums.p, v->off);
				ctx->launches += skip_redo ? 4 : 5;
			}
			VO_CUDA(cudaGetLastError());
			if (!fused) VO_CUDA(cudaMemcpyAsync(&total, sums.p + ntiles, sizeof(total), cudaMemcpyDeviceToHost, ctx->stream));
		} else {
			VO_CUDA(cudaMemsetAsync(v->off, 0, sizeof(uint32_t), ctx->stream));
		}
		if (done_ev) cudaEventRecord(done_ev, ctx->stream);
		VO_TRY(read_counters(ctx, h));
		if (fused) total = h[NCTR + 1];
		if (skip_redo && h[8] > 0) { skip_redo = false; ctx->redo_recent = 16; continue; }
		if (h[8] > redo_cap) return fail(ctx, VO_ERR_OVERFLOW, "too many lists outgrew the fast running-union capacity");
		if (h[9]) return fail_list_overflow(ctx);
		ctx->stage_hint = next_hint(ctx->stage_hint, h[1]);
		if (h[1] > sb.st.pool_cap) { VO_TRY(sb.regrow(h[1] + h[1] / 8 + 1024)); continue; }
		if (total >= (1ull << 32)) return fail(ctx, VO_ERR_OVERFLOW, "result has more than 2^32-1 intervals");
		if (fused) ctx->out_cap_hint = next_hint(ctx->out_cap_hint, total + total / 8 + 1024);
		if (!fused || total > out_cap) {
			if (v->spans) { dfree(ctx, v->spans); v->spans = nullptr; }
			VO_TRY(dalloc(ctx, &v->spans, total));
			if (nlists) {
				k_compact<<<blocks_for(nlists, 256), 256, 0, ctx->stream>>>(sb.st, nlists, v->off, v->spans);
				ctx->launches++;
				VO_CUDA(cudaGetLastError());
			}
		}
		v->nspans = total;
		if (nlists == 0) v->max_cnt = 0;
		else if (fused) v->max_cnt = (uint32_t)(h[15] >> 32) == ctx->max_epoch ? (long long)(uint32_t)h[15] : 0;
		*out = v;
		v = nullptr;          // released from the guard
		return VO_OK;
	}
	return fail(ctx, VO_ERR_OVERFLOW, "staging pool did not converge");
}

struct DevTables {
	vo_ctx *ctx;
	double *H = nullptr, *HB = nullptr;
	int *reach = nullptr;
	int J = 0;
	explicit DevTables(vo_ctx *c) : ctx(c) {}
	~DevTables() { dfree(ctx, H); dfree(ctx, HB); dfree(ctx, reach); }
	int upload(const Tables &t)
	{
		J = t.J;
		const size_t n = (size_t)(J + 1) * (J + 1);
		VO_TRY(dalloc(ctx, &H, n));
		VO_TRY(dalloc(ctx, &HB, n));
		VO_TRY(dalloc(ctx, &reach, (unsigned long long)J + 1));
		VO_CUDA(cudaMemcpyAsync(H, t.H.data(), n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
		VO_CUDA(cudaMemcpyAsync(HB, t.HB.data(), n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
		VO_CUDA(cudaMemcpyAsync(reach, t.reach.data(), (size_t)(J + 1) * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
		VO_CUDA(cudaStreamSynchronize(ctx->stream)); // host vectors may go out of scope
		return VO_OK;
	}
};

int check_radius(vo_ctx *ctx, double R)
{
	if (!(R >= 0.0) || !(R < 4096.0)) return fail(ctx, VO_ERR_ARG, "radius must be in [0, 4096) dexels");
	return VO_OK;
}

// Dominance bounds of the tile kernel (pass1_tile.cuh), from the same cap table the candidates use:
//   Dmono[d]  = min over d' >= d and classes j with reach[j] >= d' of H[j][d'-1] - H[j][d']   (near, x)
//   Emono[j]  = min over j' >= j and d <= reach[j'] of H[j'-1][d] - H[j'][d]                    (near, y)
//   G[jt][d]  = max over 1 <= d' <= d and j <= min(jt, jmax[d'+1]) of H[j][d'] - H[j][d'+1]    (far, x)
//   Ef[d][c]  = max over 1 <= c' <= c of H[c'-1][d] - H[c'][d], +inf for c > jmax[d]           (far, y)
// the near bounds non-decreasing with entry J+1 = +inf; the far bounds non-decreasing along d / c, rounded
// UP to float (a larger bound only prunes less), +inf where no farther neighbour is in reach.
struct TileTables {
	vo_ctx *ctx;
	double *Dmono = nullptr, *Emono = nullptr, *Ht = nullptr;   // Ht[d*JPP + j] = H[j][d], rows padded
	float *G = nullptr, *Ef = nullptr;
	uint8_t *jmax = nullptr;
	explicit TileTables(vo_ctx *c) : ctx(c) {}
	~TileTables() { dfree(ctx, Dmono); dfree(ctx, Emono); dfree(ctx, Ht); dfree(ctx, G); dfree(ctx, Ef); dfree(ctx, jmax); }
	static float round_up(double x)
	{
		float f = (float)x;
		if ((double)f < x) f = std::nextafterf(f, std::numeric_limits<float>::infinity());
		return f;
	}
	int upload(const Tables &t)
	{
		const int J = t.J, n = J + 1;
		if (J > 255) return VO_OK;                       // (the tile kernel is only used up to J = 63)
		std::vector<double> dm((size_t)J + 2, 0.0), em((size_t)J + 2, 0.0);
		const double inf = std::numeric_limits<double>::infinity();
		const float finf = std::numeric_limits<float>::infinity();
		dm[J + 1] = em[J + 1] = inf;
		for (int d = J; d >= 1; --d) {
			double m = inf;
			for (int j = 0; j < n; ++j)
				if (t.reach[j] >= d) m = std::min(m, t.H[(size_t)j * n + d - 1] - t.H[(size_t)j * n + d]);
			dm[d] = std::min(m, dm[d + 1]);
		}
		for (int j = J; j >= 1; --j) {
			double m = inf;
			for (int d = 0; d <= t.reach[j]; ++d) m = std::min(m, t.H[(size_t)(j - 1) * n + d] - t.H[(size_t)j * n + d]);
			em[j] = std::min(m, em[j + 1]);
		}
		std::vector<uint8_t> jm((size_t)J + 2, 0);
		for (int d = 0; d <= J; ++d)
			for (int j = 0; j < n; ++j) if (t.reach[j] >= d) jm[d] = (uint8_t)j;
		std::vector<float> g((size_t)n * n, 0.0f), ef((size_t)n * (n + 1), 0.0f);
		for (int jt = 0; jt < n; ++jt) {
			double run = 0.0;
			for (int d = 1; d < J; ++d) {
				for (int j = 0; j <= std::min(jt, (int)jm[d + 1]); ++j)
					run = std::max(run, t.H[(size_t)j * n + d] - t.H[(size_t)j * n + d + 1]);
				g[(size_t)jt * n + d] = round_up(run);
			}
			g[(size_t)jt * n + J] = finf;
		}
		for (int d = 0; d <= J; ++d) {
			double run = 0.0;
			for (int c = 1; c <= J + 1; ++c) {
				if (c <= (int)jm[d]) {
					run = std::max(run, t.H[(size_t)(c - 1) * n + d] - t.H[(size_t)c * n + d]);
					ef[(size_t)d * (n + 1) + c] = round_up(run);
				} else ef[(size_t)d * (n + 1) + c] = finf;
			}
		}
		const int jpp = pass1_jpp(J);
		std::vector<double> ht((size_t)n * jpp, -1.0);
		for (int j = 0; j < n; ++j)
			for (int d = 0; d < n; ++d) ht[(size_t)d * jpp + j] = t.H[(size_t)j * n + d];
		VO_TRY(dalloc(ctx, &Ht, (unsigned long long)ht.size()));
		VO_CUDA(cudaMemcpyAsync(Ht, ht.data(), ht.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
		VO_TRY(dalloc(ctx, &Dmono, (unsigned long long)J + 2));
		VO_TRY(dalloc(ctx, &Emono, (unsigned long long)J + 2));
		VO_TRY(dalloc(ctx, &G, (unsigned long long)g.size()));
		VO_TRY(dalloc(ctx, &Ef, (unsigned long long)ef.size()));
		VO_TRY(dalloc(ctx, &jmax, (unsigned long long)jm.size()));
		VO_CUDA(cudaMemcpyAsync(Dmono, dm.data(), dm.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
		VO_CUDA(cudaMemcpyAsync(Emono, em.data(), em.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
		VO_CUDA(cudaMemcpyAsync(G, g.data(), g.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
		VO_CUDA(cudaMemcpyAsync(Ef, ef.data(), ef.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
		VO_CUDA(cudaMemcpyAsync(jmax, jm.data(), jm.size(), cudaMemcpyHostToDevice, ctx->stream));
		VO_CUDA(cudaStreamSynchronize(ctx->stream));
		return VO_OK;
	}
};

// The tables only depend on the radius: keep the last set on the device so that repeated calls with the
// same radius (opening / closing, benchmark steps, slabs) do not upload or synchronise again.
struct TableCache {
	double R;
	Tables t;
	DevTables dt;
	TileTables tt;
	bool has_tile = false;
	TableCache(vo_ctx *ctx, double r) : R(r), t(make_tables(r)), dt(ctx), tt(ctx) {}
};

void free_table_cache(vo_ctx *ctx)
{
	delete static_cast<TableCache *>(ctx->table_cache);
	ctx->table_cache = nullptr;
}

int get_tables(vo_ctx *ctx, double R, bool need_tile, TableCache **out)
{
	TableCache *tc = static_cast<TableCache *>(ctx->table_cache);
	if (!tc || std::memcmp(&tc->R, &R, sizeof(double)) != 0) {
		free_table_cache(ctx);
		tc = new (std::nothrow) TableCache(ctx, R);
		if (!tc) return fail(ctx, VO_ERR_NOMEM, "out of host memory");
		ctx->table_cache = tc;
		int rc = tc->dt.upload(tc->t);
		if (rc) { free_table_cache(ctx); return rc; }
	}
	if (need_tile && !tc->has_tile) {
		int rc = tc->tt.upload(tc->t);
		if (rc) { free_table_cache(ctx); return rc; }
		tc->has_tile = true;
	}
	*out = tc;
	return VO_OK;
}

// Launch plan of the tile kernel (pass1_tile.cuh): candidate-buffer sizes, warps per CTA and shared memory of the
// three launches, from the grid and the mean fill. One CTA per SM; warps work on their own.
// thresholds of the columns [ta.c_begin, ta.c_end): one thread per column, or four (one per neighbour) when the
// columns hold several intervals each
inline void launch_thresh(const ThreshArgs &ta, double k_in, cudaStream_t s)
{
	const unsigned long long n = ta.c_end - ta.c_begin;
	const size_t smem = 2 * (size_t)(ta.J + 2) * sizeof(double);
	if (k_in >= 4.0) k_thresh_quad<<<blocks_for(TH_Q * n, 256), 256, smem, s>>>(ta);
	else k_thresh<<<blocks_for(n, 256), 256, smem, s>>>(ta);
}

struct TilePlan {
	int J = 0, tiles_xw = 0, tiles_x = 0, sms = 148, cps = 1, quota = 0;
	int cmax_small = 0, cmax_big = 0, cmax_multi = 0;
	int nw_small = 1, nw_big = 1, nw_multi = 1, nw_bigmulti = 1, nw_small_dual = 1;
	bool db_small = true, db_big = true, db_multi = true, db_bigmulti = false;   // candidates double-buffered (pass1_warp_smem)
	bool multi_bounded = false;
	bool lean1 = true;                // first launch without the inline sorted-list union (20 warps; vo_ctx::gen_inline_calls)
	bool lean_big = false, lean_multi = false, lean_bigmulti = false;             // ... or left in global memory
	size_t smem_bigmulti = 0;
	size_t smem_small = 0, smem_big = 0, smem_multi = 0, smem_small_dual = 0;
	bool db_small_dual = true;
	static constexpr int CMAX = 2048;       // largest candidate buffer (11-bit candidate ids in the survivor lists)
	static bool fits(int J, double k_in) { return J <= 63 && k_in * (P1_W + 2 * J) <= 0.75 * CMAX; }
	// warps_cap: fewer warps per CTA than the registers allow (the host-buffer pipeline leaves room on every SM for the
	// small kernels of the other bands, which otherwise wait for a persistent tile CTA to retire)
	// spill areas of the survivor lists: one per warp of the largest grid, two banks (a second launch set may run beside
	// the first - slab boundary rows, alternating pipeline bands)
	static int reserve_ovf(vo_ctx *ctx, size_t areas)
	{
		if (ctx->ovf && ctx->ovf_areas >= areas) return VO_OK;
		if (ctx->ovf) { cudaDeviceSynchronize(); cudaFree(ctx->ovf); ctx->ovf = nullptr; ctx->ovf_areas = 0; }
		cudaError_t ea = cudaMalloc((void **)&ctx->ovf, 2 * areas * P1_OVF * P1_W * sizeof(uint32_t));
		if (ea != cudaSuccess) { cudaGetLastError(); ctx->ovf = nullptr; return fail(ctx, VO_ERR_NOMEM, "spill area of the tile kernel"); }
		ctx->ovf_areas = areas;
		return VO_OK;
	}
	// first-launch grid over `ntiles` tiles
	unsigned int grid_small(unsigned int ntiles, int sms_avail, int nw = 0) const
	{
		if (nw <= 0) nw = nw_small;
		if (quota > 0) return std::max(1u, (ntiles + (unsigned int)(nw * quota) - 1) / (unsigned int)(nw * quota));
		return (unsigned int)std::max(1u, std::min<unsigned int>((unsigned int)(sms_avail * cps), (ntiles + nw - 1) / nw));
	}
	int init(vo_ctx *ctx, int nx, int J_, double k_in, int warps_cap = 64, int cps_override = 0, int quota_ = 0, unsigned int ntiles_max = 0, long long max_cnt = -1, int ny_hint = 0)
	{
		J = J_;
		quota = quota_;
		tiles_xw = (nx + P1_W - 1) / P1_W;
		tiles_x = (nx + P1_TX - 1) / P1_TX;
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
		const int SEG = P1_W + 2 * J;
		auto pick = [](double est) { int c = 64; while (c < CMAX && c < est) c <<= 1; return c; };
		// launch 1 only sees single-interval columns: at most SEG candidates
		cmax_small = std::min(pick(k_in * SEG * 1.5), (SEG + 15) & ~15);
		cmax_big = CMAX;
		// launch 3 (multi-interval columns): a quarter above the mean fill, in steps of 32 (shared memory per warp is what
		// bounds the warps per SM of this launch; the few tiles beyond it go to the redo list)
		cmax_multi = std::min(CMAX, std::max(128, ((int)std::ceil(k_in * SEG * 1.25) + 31) & ~31));
		// `cps` CTAs per SM share its shared memory (228 KB, 1 KB of it reserved per CTA) and its 16 warps' worth of
		// registers: several small CTAs give an SM back piecewise when a launch runs out of tiles, one large CTA only
		// when its last warp is done
		cps = std::max(1, std::min(cps_override > 0 ? cps_override : ctx->tile_ctas, 8));
		const size_t budget = std::min<size_t>(220 * 1024, 228 * 1024 / cps - 1024 - 256);
		auto warps = [&](int cmax, int lcap, int maxw, bool dbuf, bool lean) {
			const size_t per = pass1_warp_smem(J, cmax, lcap, dbuf, lean), tab = pass1_table_smem(J) + 32;
			if (budget < tab + per) return 1;
			return (int)std::max<size_t>(1, std::min<size_t>(std::min(maxw, warps_cap) / cps, (budget - tab) / per));
		};
		// The staging of a launch, in order of preference: double-buffered candidates in shared memory where that still
		// leaves room for every warp the registers allow, one buffer otherwise, and for the launches that pull tiles from a
		// list (may_lean) nothing but the column map when even that costs warps.
		auto plan = [&](int cmax, int lcap, int maxw, bool may_lean, int &nw, bool &db, bool &lean, size_t &smem) {
			const int want = std::max(1, std::min(maxw, warps_cap) / cps);
			lean = false;
			db = ctx->tile_dbuf > 0 || (ctx->tile_dbuf < 0 && warps(cmax, lcap, maxw, true, false) >= want);
			if (!db && may_lean && ctx->tile_lean != 0 && (ctx->tile_lean > 0 || warps(cmax, lcap, maxw, false, false) < want)) lean = true;
			nw = warps(cmax, lcap, maxw, db, lean);
			smem = pass1_tile_smem(J, cmax, lcap, nw, db, lean);
		};
		bool dummy = false;
		lean1 = ctx->gen_mode == 0 || (ctx->gen_mode < 0 && ctx->gen_inline_calls == 0);
		if (!lean1 && ctx->gen_inline_calls > 0) --ctx->gen_inline_calls;
		plan(cmax_small, P1_LCAP_S, lean1 ? P1_MAXWARPS_LEAN : P1_MAXWARPS, false, nw_small, db_small, dummy, smem_small);
		{   // (the dual form's first launch never meets a complex class: always the 20-warp variant)
			bool db = true, ln = false;
			plan(cmax_small, P1_LCAP_S, P1_MAXWARPS_LEAN, false, nw_small_dual, db, ln, smem_small_dual);
			db_small_dual = db;
		}
		plan(cmax_big, P1_LCAP_M, P1_MAXWARPS, true, nw_big, db_big, lean_big, smem_big);
		// (vo_ctx::multi_warps) deep columns - known from the volume, or seen in the last mid pool - go through the sorted-list
		// union, whose local-memory lists want the L1 to themselves: 16 warps thrash it (profiles/r2ce_layer_major_ab.txt: lattice 512,
		// R = 5 / 8 / 12: 1.30 / 2.05 / 2.55 ms with 16 warps, 1.23 / 1.87 / 2.12 with 14, 1.29 / 1.83 / 1.96 with 12)
		// - unless the grid has about one tile per warp anyway (lattice 256, R = 12: 0.76 ms with 16 warps, 0.99 with 12)
		const bool few_tiles = ny_hint > 0 && (long long)tiles_xw * ny_hint < 2ll * sms * P1_MAXWARPS_M;
		const bool deep = (max_cnt >= 3 || ctx->pooled_per_column >= 8.0) && !few_tiles;
		const int mw = std::min(P1_MAXWARPS_M, ctx->multi_warps > 0 ? ctx->multi_warps : deep ? (J <= 6 ? 14 : 12) : P1_MAXWARPS_M);
		plan(cmax_multi, P1_LCAP_M, mw, true, nw_multi, db_multi, lean_multi, smem_multi);
		if (lean_multi) {                // lean is cheap per candidate (4 bytes): one launch for every multi-interval tile, sized by what
			// a tile can hold at most where the deepest column is known (vo_dvol::max_cnt) - shared memory that is not asked
			// for stays L1, which the local-memory lists of the sorted-list union live in
			cmax_multi = cmax_big;
			if (max_cnt > 0 && max_cnt * SEG < (long long)cmax_big) { cmax_multi = std::max(128, ((int)(max_cnt * SEG) + 15) & ~15); multi_bounded = true; }
			plan(cmax_multi, P1_LCAP_M, mw, true, nw_multi, db_multi, lean_multi, smem_multi);
		}
		plan(cmax_big, P1_LCAP_M, mw, true, nw_bigmulti, db_bigmulti, lean_bigmulti, smem_bigmulti);   // launch 4: the multi-interval tiles beyond cmax_multi
		cudaError_t e = cudaFuncSetAttribute(k_pass1_tile<CAP_FAST, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_small);
		if (e == cudaSuccess) e = cudaFuncSetAttribute(k_pass1_tile<CAP_FAST, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_big);
		if (e == cudaSuccess) e = cudaFuncSetAttribute(k_pass1_tile<CAP_FAST, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(smem_multi, smem_bigmulti));
		if (e == cudaSuccess) e = cudaFuncSetAttribute(k_pass1_tile<CAP_FAST, false, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_small);
		if (e == cudaSuccess) e = cudaFuncSetAttribute(k_pass1_tile<CAP_FAST, false, false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_small_dual);
		if (e == cudaSuccess) e = cudaFuncSetAttribute(k_pass1_tile<CAP_FAST, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_big);
		if (e != cudaSuccess) return fail(ctx, VO_ERR_CUDA, std::string("k_pass1_tile smem: ") + cudaGetErrorString(e));
		size_t areas = (size_t)sms * std::max(P1_MAXWARPS, P1_MAXWARPS_LEAN);
		if (quota > 0) areas = std::max(areas, (size_t)grid_small(ntiles_max, sms) * std::max(nw_small, nw_small_dual));
		return reserve_ovf(ctx, areas);
	}
	// The four launches over the tiles [tile0, tile0 + ntiles). `g` holds the data pointers; the lists and cursors
	// ([3] big count, [5] multi count, [12] big multi count, [6] [7] [10] [13] cursors of d_ctr) must be zero.
	// A second range [tile0b, tile0b + ntilesb) may follow the first; reserve_sms CTAs fewer are launched (room for the
	// NCCL kernels of a halo exchange in flight).
	void launch(vo_ctx *ctx, Pass1TileArgs g, unsigned int tile0, unsigned int ntiles0, unsigned int *big_tiles, unsigned int *multi_tiles,
	            cudaStream_t s, unsigned int tile0b = 0, unsigned int ntilesb = 0, int reserve_sms = 0,
	            unsigned long long *bank = nullptr, const unsigned int *order = nullptr, bool dual = false, int ovf_bank = -1,
	            bool single = false, unsigned int *sticky = nullptr) const
	{
		// sticky: with `single`, counts the tiles that would have needed a launch that is not made (the caller repeats with them)
		// single: every column is KNOWN to hold at most one interval (vo_dvol::max_cnt) - no tile can be a multi-interval
		// one, and none can hold more candidates than its segment has columns: the list launches that could only find
		// empty lists are not made
		if (!bank) bank = ctx->d_ctr;                       // ([3] [5] [6] [7] [10] [12] [13] of `bank`: the lists and cursors of this launch set)
		const unsigned int ntiles = ntiles0 + ntilesb;
		const int sms = std::max(1, this->sms - reserve_sms);
		g.J = J; g.tiles_xw = tiles_xw; g.tiles_x = tiles_x; g.tile0 = tile0; g.ntiles = ntiles; g.tile0b = tile0b; g.ntiles0 = ntiles0;
		if (ovf_bank < 0) ovf_bank = bank == ctx->d_ctr ? 0 : 1;
		g.ovf = ctx->ovf + (size_t)ovf_bank * ctx->ovf_areas * P1_OVF * P1_W;
		g.dbg = ctx->dbg_tiles;
		g.layer_major = ctx->cand_order;
		unsigned int *big_count = reinterpret_cast<unsigned int *>(bank + 3);
		unsigned int *multi_count = reinterpret_cast<unsigned int *>(bank + 5);
		g.big_count = big_count; g.multi_tiles = multi_tiles; g.multi_count = multi_count;
		// (a large buffer may leave room for fewer CTAs per SM than planned: the grid is then more than one wave, which
		// the tile counter does not mind)
		auto grid = [&](int nw) { return (unsigned int)std::max(1u, std::min<unsigned int>((unsigned int)(sms * cps), (ntiles + nw - 1) / nw)); };
		// launch 1: single-interval tiles, small candidate buffer
		g.cmax = cmax_small; g.dbuf = db_small; g.lean = 0; g.tiles = nullptr; g.tiles_count = nullptr; g.order = order;
		g.tiles_next = reinterpret_cast<unsigned int *>(bank + 10);
		g.big_tiles = cmax_small < cmax_big ? big_tiles : nullptr;
		g.quota = quota;
		const bool launch2 = cmax_small < cmax_big && !(single && cmax_small >= P1_W + 2 * J);
		g.sticky_multi = single ? sticky : nullptr; g.sticky_big = (single && !launch2) ? sticky : nullptr;
		if (dual) {
			g.dbuf = db_small_dual;
			k_pass1_tile<CAP_FAST, false, false, true, false><<<grid_small(ntiles, sms, nw_small_dual), 32 * nw_small_dual, smem_small_dual, s>>>(g);
		} else if (lean1) k_pass1_tile<CAP_FAST, false, false, false, false><<<grid_small(ntiles, sms), 32 * nw_small, smem_small, s>>>(g);
		else k_pass1_tile<CAP_FAST, false, false><<<grid_small(ntiles, sms), 32 * nw_small, smem_small, s>>>(g);
		ctx->last_lean1 = !dual && lean1;
		g.quota = 0;
		ctx->launches++;
		g.sticky_multi = g.sticky_big = nullptr;
		if (launch2) {                  // launch 2: the single-interval tiles that need the large buffer
			g.cmax = cmax_big; g.dbuf = db_big; g.lean = lean_big; g.tiles = big_tiles; g.tiles_count = big_count; g.big_tiles = nullptr;
			g.tiles_next = reinterpret_cast<unsigned int *>(bank + 6);
			if (dual) k_pass1_tile<CAP_FAST, false, true, true><<<grid(nw_big), 32 * nw_big, smem_big, s>>>(g);
			else k_pass1_tile<CAP_FAST, false, true><<<grid(nw_big), 32 * nw_big, smem_big, s>>>(g);
			ctx->launches++;
		}
		// (dual form: every column holds one interval; a tile that says otherwise is only counted, [5], and the caller
		// falls back to the general erosion)
		if (dual || single) return;
		// launch 3: tiles with multi-interval columns (two hulls per class), candidate buffer a quarter above the mean fill;
		// the tiles beyond it are collected again (in big_tiles, which launch 2 is done with) for launch 4
		unsigned int *bigmulti_count = reinterpret_cast<unsigned int *>(bank + 12);
		const bool four = cmax_multi < cmax_big && !multi_bounded;   // (bounded: no tile can hold more than cmax_multi candidates)
		g.cmax = cmax_multi; g.dbuf = db_multi; g.lean = lean_multi; g.tiles = multi_tiles; g.tiles_count = multi_count;
		g.big_tiles = four ? big_tiles : nullptr; g.big_count = bigmulti_count;
		g.tiles_next = reinterpret_cast<unsigned int *>(bank + 7);
		k_pass1_tile<CAP_FAST, true, true><<<grid(nw_multi), 32 * nw_multi, smem_multi, s>>>(g);
		ctx->launches++;
		if (four) {                     // launch 4: the largest buffer; whatever exceeds that goes to the redo list
			g.cmax = cmax_big; g.dbuf = db_bigmulti; g.lean = lean_bigmulti; g.tiles = big_tiles; g.tiles_count = bigmulti_count; g.big_tiles = nullptr;
			g.tiles_next = reinterpret_cast<unsigned int *>(bank + 13);
			k_pass1_tile<CAP_FAST, true, true><<<grid(nw_bigmulti), 32 * nw_bigmulti, smem_bigmulti, s>>>(g);
			ctx->launches++;
		}
	}
	// Expensive tiles first: `est` (k_thresh, per tile) -> `order`, a permutation of the positions of the same two
	// ranges launch() takes; scratch = [2 * P1_NBUCKET] zeroed counters.
	static void order_tiles(vo_ctx *ctx, const unsigned int *est, unsigned int *scratch, unsigned int *order, unsigned int tile0,
	                        unsigned int ntiles0, unsigned int tile0b, unsigned int ntilesb, cudaStream_t s, bool one_cta = false)
	{
		OrderArgs oa;
		oa.est = est; oa.tile0 = tile0; oa.ntiles0 = ntiles0; oa.tile0b = tile0b; oa.ntiles = ntiles0 + ntilesb;
		oa.hist = scratch; oa.order = order;
		if (!oa.ntiles) return;
		if (one_cta && oa.ntiles <= P1_ORDER_ONE_MAX) {
			k_order_one<<<1, 1024, 0, s>>>(oa);
			ctx->launches++;
			return;
		}
		k_order_count<<<blocks_for(oa.ntiles, 256), 256, 0, s>>>(oa);
		k_order_place<<<blocks_for(oa.ntiles, 256), 256, 0, s>>>(oa);
		ctx->launches += 2;
	}
};

// ---- 'ours' pass 1 ------------------------------------------------------------------------------------
// clip_lo / clip_hi: the caller only reads the dilation inside (clip_lo, clip_hi) (erosion); the pruning of the tile
// kernel may then ignore what happens outside. -inf / +inf: the exact dilation everywhere.
// defer: enqueue only - no host round trip; the counters that say whether the pools were large enough are read by the
// caller together with those of pass 2 (which never follows a reference beyond a pool), and a pass 1 that fell short is
// repeated without `defer`. Saves one synchronisation per dilation.
// A pass 1 whose first tile launch ran without the inline sorted-list union (TilePlan::lean1) and that handed slots to the
// redo launch has probably met "complex" classes: the next calls of this context take the inline variant again.
inline void note_pass1_redo(vo_ctx *ctx, unsigned long long redo_count)
{
	if (ctx->last_lean1 && redo_count != 0) ctx->gen_inline_calls = 64;
	ctx->last_lean1 = false;
}

// Would pass 1 of this volume take the tile kernel?
bool pass1_uses_tile(const vo_ctx *ctx, const vo_dvol *in, double R)
{
	const int J0 = (int)std::floor(R);
	const unsigned long long ncols = (unsigned long long)in->nx * in->ny;
	const double k_in = ncols ? (double)in->nspans / (double)ncols : 0.0;
	// (small problems do not fill the machine with one lane per column: the simple kernel has J+1 times more threads;
	// dense columns tip the balance earlier - the tile kernel's class windows also spare pass 2 most of its reads)
	const bool big = (double)ncols * (J0 + 1) * std::max(1.0, k_in) >= (double)(2ull << 20) || ctx->force_tile_pass1;
	return ncols > 0 && TilePlan::fits(J0, k_in) && big && !ctx->force_simple_pass1;
}

// dual: the mid volume of the MIRRORED intervals (erode_dual; tile kernel only, always deferred). [dual_lo, dual_hi] =
// the bounds of the erosion's complement (zmin - 1, zmax + 1): every interval must lie strictly inside.
struct DualSpec { double lo, hi; };

int pass1(vo_ctx *ctx, const vo_dvol *in, double R, vo_dmid **out,
          double clip_lo = -std::numeric_limits<double>::infinity(), double clip_hi = std::numeric_limits<double>::infinity(),
          bool defer = false, const DualSpec *dual = nullptr)
{
	VO_TRY(check_radius(ctx, R));
	const int J0 = (int)std::floor(R);
	const unsigned long long ncols = (unsigned long long)in->nx * in->ny;
	// tile kernel (pass1_tile.cuh) whenever its tables and tile fit in shared memory and a row segment's
	// candidates are expected to fit; otherwise the one-thread-per-(x,y,j) kernel does everything
	const int TX = P1_TX;
	const double k_in = ncols ? (double)in->nspans / (double)ncols : 0.0;
	const bool use_tile = pass1_uses_tile(ctx, in, R);
	if (dual && (!use_tile || !defer)) return fail(ctx, VO_ERR_ARG, "dual pass 1 needs the tile kernel");
	TableCache *tc = nullptr;
	VO_TRY(get_tables(ctx, R, use_tile, &tc));
	const Tables &t = tc->t;
	DevTables &dt = tc->dt;
	TileTables &tt = tc->tt;

	vo_dmid *m = new (std::nothrow) vo_dmid();
	if (!m) return fail(ctx, VO_ERR_NOMEM, "out of host memory");
	m->nx = in->nx; m->ny = in->ny; m->J = t.J; m->R = R;
	m->shallow = in->max_cnt >= 0 && in->max_cnt <= 1;
	const unsigned long long nslots = ncols * (t.J + 1);
	int rc = dalloc(ctx, &m->slots, nslots);
	unsigned long long pool_cap = std::max(65536ull + (unsigned long long)(t.J + 1) * (in->nspans / 4), ctx->pool_hint);
	if (rc == VO_OK) rc = dalloc(ctx, &m->pool, pool_cap);
	if (rc == VO_OK) rc = dalloc(ctx, &m->flags, 2 * ncols);
	const unsigned long long nmask = 2ull * in->ny * ((in->nx + TX - 1) / TX);
	// (the tile order's counters and cost estimates sit behind the tile masks: one block, one memset)
	const unsigned long long ntiles_pre = (unsigned long long)((in->nx + P1_W - 1) / P1_W) * in->ny;
	const bool ordered_pre = use_tile && ctx->tile_order && ntiles_pre >= 4096;
	const unsigned long long nest = ordered_pre ? (ntiles_pre + 2 * P1_NBUCKET + 1) / 2 : 0;
	if (rc == VO_OK) rc = dalloc(ctx, &m->tilemask, nmask + nest);
	Tmp<uint4> thr(ctx);
	if (rc == VO_OK && use_tile) rc = dalloc(ctx, &thr.p, in->nspans);
	TilePlan plan;
	if (rc == VO_OK && use_tile) rc = plan.init(ctx, in->nx, t.J, k_in, 64, 0, 0, 0, in->max_cnt, in->ny);
	const unsigned long long ntiles = (unsigned long long)plan.tiles_xw * in->ny;
	Tmp<unsigned int> big_tiles(ctx), multi_tiles(ctx);
	if (rc == VO_OK && use_tile) rc = dalloc(ctx, &big_tiles.p, ntiles);
	if (rc == VO_OK && use_tile) rc = dalloc(ctx, &multi_tiles.p, ntiles);
	// tile order (expensive first): [2 * P1_NBUCKET counters | cost estimate per tile], and the permutation
	Tmp<unsigned int> order(ctx);
	const bool ordered = ordered_pre;
	struct { unsigned int *p; } est{rc == VO_OK && ordered ? reinterpret_cast<unsigned int *>(m->tilemask + nmask) : nullptr};
	if (rc == VO_OK && ordered) rc = dalloc(ctx, &order.p, ntiles);
	RedoBuf rb(ctx);
	const unsigned int redo_cap = (unsigned int)std::min<unsigned long long>(std::max<unsigned long long>(nslots, 1ull), 1ull << 22);
	if (rc == VO_OK) rc = rb.alloc(redo_cap);
	if (rc != VO_OK) { vo_dmid_free(ctx, m); return rc; }
	m->pool_cap = pool_cap;
	auto bail = [&](int code) { vo_dmid_free(ctx, m); return code; };
	bool tile_now = use_tile;
	for (int attempt = 0; attempt < 4; ++attempt) {
		cudaError_t e = cudaMemsetAsync(ctx->d_ctr, 0, NCTR * sizeof(unsigned long long), ctx->stream);
		if (e != cudaSuccess) return bail(fail(ctx, VO_ERR_CUDA, cudaGetErrorString(e)));
		ctx->staged_ctr_clean = defer;                       // (pass 2 follows at once and need not zero its counters again)
		Pass1Args a;
		a.nx = in->nx; a.ny = in->ny; a.J = t.J;
		a.off = in->off; a.spans = in->spans; a.H = dt.H; a.reach = dt.reach;
		a.mid = m->slots; a.pool = m->pool; a.cursor = ctx->d_ctr; a.pool_cap = m->pool_cap;
		a.redo = rb.rd;
		a.wk = Work{nullptr, nslots, nullptr, 0u, nullptr};
		if (nslots && tile_now) {
			cudaMemsetAsync(m->tilemask, 0, (nmask + nest) * sizeof(unsigned long long), ctx->stream);
			ThreshArgs ta;
			ta.nx = in->nx; ta.ny = in->ny; ta.J = t.J; ta.off = in->off; ta.spans = in->spans;
			ta.Dmono = tt.Dmono; ta.Emono = tt.Emono; ta.G = tt.G; ta.reach = dt.reach; ta.thr = thr.p;
			ta.c_begin = 0; ta.c_end = ncols; ta.clip_lo = clip_lo; ta.clip_hi = clip_hi;
			if (dual) { ta.dual = 1; ta.dual_lo = dual->lo; ta.dual_hi = dual->hi; ta.dual_bad = reinterpret_cast<unsigned int *>(ctx->d_ctr + 14); }
			if (ordered) { ta.est = est.p + 2 * P1_NBUCKET; ta.tiles_xw = plan.tiles_xw; }
			launch_thresh(ta, k_in, ctx->stream);
			ctx->launches++;
			if (ordered) TilePlan::order_tiles(ctx, ta.est, est.p, order.p, 0u, (unsigned int)ntiles, 0u, 0u, ctx->stream);
			Pass1TileArgs g;
			g.nx = in->nx; g.ny = in->ny;
			g.off = in->off; g.spans = in->spans; g.thr = thr.p; g.Ht = tt.Ht; g.Ef = tt.Ef; g.jmax = tt.jmax;
			g.mid = m->slots; g.flags = m->flags; g.tilemask = m->tilemask; g.pool = m->pool; g.cursor = ctx->d_ctr; g.pool_cap = m->pool_cap; g.redo = rb.rd;
			cudaEventRecord(ctx->kev[0], ctx->stream);
			plan.launch(ctx, g, 0u, (unsigned int)ntiles, big_tiles.p, multi_tiles.p, ctx->stream, 0u, 0u, 0, nullptr, ordered ? order.p : nullptr, dual != nullptr,
			            -1, in->max_cnt >= 0 && in->max_cnt <= 1);
			cudaEventRecord(ctx->kev[1], ctx->stream);
			ctx->kev_valid[0] = true;
		} else if (nslots) {
			// every class of every column is computed: all class windows = [0, J + 1)
			// (an 8-bit window bound only reaches class 254: larger radii get the "every class" sentinel)
			k_fill16<<<blocks_for(2 * ncols, 256), 256, 0, ctx->stream>>>(m->flags, 2 * ncols, t.J + 1 <= 255 ? (uint16_t)((t.J + 1) << 8) : FLAG_ALL);
			cudaMemsetAsync(m->tilemask, 0xFF, nmask * sizeof(unsigned long long), ctx->stream);
			ctx->launches++;
			cudaEventRecord(ctx->kev[0], ctx->stream);
			k_pass1<CAP_FAST><<<blocks_for(nslots, 128), 128, 0, ctx->stream>>>(a);
			cudaEventRecord(ctx->kev[1], ctx->stream);
			ctx->kev_valid[0] = true;
			ctx->launches++;
		}
		// (deferred: the redo launch - normally idle - is left out like run_staged's; the caller, who reads the counters
		// after pass 2, repeats the dilation without `defer` when a list did outgrow the fast capacity)
		const bool skip_redo = defer && ctx->redo_recent == 0 && attempt == 0 && !ctx->huge;
		m->redo_skipped = skip_redo;
		if (nslots && !dual && !skip_redo) {
			// redo launch over the device-side list (fixed grid, reads the count itself)
			a.wk = Work{rb.rd.list, 0ull, rb.rd.count, rb.rd.cap, reinterpret_cast<unsigned int *>(ctx->d_ctr + 4)};
			a.wk.huge = ctx->huge_scratch;
			if (ctx->huge) k_pass1<CAP_HUGE><<<HUGE_GRID, 128, 0, ctx->stream>>>(a);
			else k_pass1<CAP_BIG><<<REDO_GRID, 128, 0, ctx->stream>>>(a);
			ctx->launches++;
		}
		e = cudaGetLastError();
		if (e != cudaSuccess) return bail(fail(ctx, VO_ERR_CUDA, std::string("k_pass1: ") + cudaGetErrorString(e)));
		if (defer) { m->deferred = true; m->redo_cap = redo_cap; *out = m; return VO_OK; }
		unsigned long long h[NREAD];
		rc = read_counters(ctx, h);
		if (rc) return bail(rc);
		note_pass1_redo(ctx, h[2]);
		if (h[2] > redo_cap) {
			if (tile_now) { tile_now = false; continue; }       // too many oversized tiles: simple kernel for everything
			return bail(fail(ctx, VO_ERR_OVERFLOW, "too many lists outgrew the fast running-union capacity"));
		}
		if (h[4]) return bail(fail_list_overflow(ctx));
		if (h[0] <= m->pool_cap) {
			m->pool_used = h[0]; ctx->pool_hint = next_hint(ctx->pool_hint, h[0]);
			if (tile_now && h[5]) ctx->pooled_per_column = (double)h[0] / (double)std::max<unsigned long long>(1, (unsigned long long)in->nx * in->ny);
			*out = m;
			return VO_OK;
		}
		dfree(ctx, m->pool);
		m->pool = nullptr;
		m->pool_cap = h[0] + h[0] / 8 + 1024;
		rc = dalloc(ctx, &m->pool, m->pool_cap);
		if (rc) return bail(rc);
	}
	return bail(fail(ctx, VO_ERR_OVERFLOW, "mid pool did not converge"));
}

int pass2(vo_ctx *ctx, const vo_dmid *m, int y0, int y1, vo_dvol **out, cudaEvent_t done_ev = nullptr)
{
	if (y0 < 0 || y1 > m->ny || y0 > y1) return fail(ctx, VO_ERR_ARG, "pass 2 row range outside the mid volume");
	Pass2Args a;
	a.nx = m->nx; a.ny = m->ny; a.J = m->J; a.y0 = y0; a.y1 = y1;
	a.mid = m->slots; a.flags = m->flags; a.tilemask = m->tilemask; a.pool = m->pool; a.pool_cap = m->pool_cap;
	const unsigned long long nlists = (unsigned long long)m->nx * (y1 - y0);
	cudaStream_t s = ctx->stream;
	// Shallow input (one interval per column): nearly every output column is one or two intervals, which the running union
	// keeps in registers. The kernel instantiated with capacity 2 has no list, no call into the list code and no local
	// memory (C5 k_pass2_rows 0.175 -> 0.155 ms); a third interval sends the column to the redo launch. A call that sent
	// more than one column in 128 there goes back to the list-capable kernel for the next 64 calls.
	const bool regonly = a.J <= 32 && m->shallow && (ctx->p2_mode == 0 || (ctx->p2_mode < 0 && ctx->p2_list_calls == 0));
	if (!regonly && ctx->p2_list_calls > 0) --ctx->p2_list_calls;
	const int rc2 = run_staged(ctx, a, nlists, 65536ull + nlists / 8,
		[&](Pass2Args &g) {
			cudaEventRecord(ctx->kev[2], s);
			if (regonly)
				k_pass2_rows<2, false, P2_REGONLY><<<(unsigned int)((g.nx + P2_TX - 1) / P2_TX) * (unsigned int)(g.y1 - g.y0), P2_TX, 0, s>>>(g);
			else if (g.J <= 32 && m->shallow)
				k_pass2_rows<CAP_FAST, false, P2_SHALLOW><<<(unsigned int)((g.nx + P2_TX - 1) / P2_TX) * (unsigned int)(g.y1 - g.y0), P2_TX, 0, s>>>(g);
			else if (g.J <= 32)
				k_pass2_rows<CAP_FAST, false><<<(unsigned int)((g.nx + P2_TX - 1) / P2_TX) * (unsigned int)(g.y1 - g.y0), P2_TX, 0, s>>>(g);
			else if (g.J <= 63)
				k_pass2_rows<CAP_FAST><<<(unsigned int)((g.nx + P2_TX - 1) / P2_TX) * (unsigned int)(g.y1 - g.y0), P2_TX, 0, s>>>(g);
			else
				k_pass2<CAP_FAST><<<blocks_for(g.wk.n, 128), 128, 0, s>>>(g);
			cudaEventRecord(ctx->kev[3], s);
			ctx->kev_valid[1] = true;
		},
		[&](Pass2Args &g, unsigned int grid) {
			if (ctx->huge) k_pass2<CAP_HUGE><<<HUGE_GRID, 128, 0, s>>>(g);
			else k_pass2<CAP_BIG><<<grid, 128, 0, s>>>(g);
		},
		m->nx, y1 - y0, out, done_ev);
	if (regonly && ctx->last_ctr[8] > std::max<unsigned long long>(64, nlists / 128)) ctx->p2_list_calls = 64;
	return rc2;
}

// pass 2 of the dual form (k_pass2_rows_dual): hull of the mirrored slots, empty columns in reach, negateInv's clamping
int pass2_dual(vo_ctx *ctx, const vo_dmid *m, const uint8_t *dist, const int *reach, double lo, double hi, vo_dvol **out, cudaEvent_t done_ev = nullptr,
               int y0 = 0, int y1 = -1)
{
	if (y1 < 0) y1 = m->ny;
	if (y0 < 0 || y1 > m->ny || y0 > y1) return fail(ctx, VO_ERR_ARG, "pass 2 row range outside the mid volume");
	Pass2Args a;
	a.nx = m->nx; a.ny = m->ny; a.J = m->J; a.y0 = y0; a.y1 = y1;
	a.mid = m->slots; a.flags = m->flags; a.tilemask = m->tilemask; a.pool = m->pool; a.pool_cap = m->pool_cap;
	a.dist = dist; a.reach = reach; a.lo = lo; a.hi = hi;
	a.dist_tmin = dist + (size_t)m->nx * m->ny;                 // (k_empty_dist leaves the tile minima behind the distances)
	const unsigned long long nlists = (unsigned long long)m->nx * (y1 - y0);
	cudaStream_t s = ctx->stream;
	return run_staged(ctx, a, nlists, 65536ull,
		[&](Pass2Args &g) {
			cudaEventRecord(ctx->kev[2], s);
			if (g.J <= 32) k_pass2_rows_dual<false><<<(unsigned int)((g.nx + P2_TX - 1) / P2_TX) * (unsigned int)(g.y1 - g.y0), P2_TX, 0, s>>>(g);
			else k_pass2_rows_dual<true><<<(unsigned int)((g.nx + P2_TX - 1) / P2_TX) * (unsigned int)(g.y1 - g.y0), P2_TX, 0, s>>>(g);
			cudaEventRecord(ctx->kev[3], s);
			ctx->kev_valid[1] = true;
		},
		[&](Pass2Args &, unsigned int) {},                       // (never needed: at most one interval per column)
		m->nx, y1 - y0, out, done_ev);
}

int brute(vo_ctx *ctx, const vo_dvol *in, double R, vo_dvol **out)
{
	VO_TRY(check_radius(ctx, R));
	TableCache *tc = nullptr;
	VO_TRY(get_tables(ctx, R, false, &tc));
	const Tables &t = tc->t;
	DevTables &dt = tc->dt;
	BruteArgs a;
	a.nx = in->nx; a.ny = in->ny; a.J = t.J;
	a.off = in->off; a.spans = in->spans; a.HB = dt.HB;
	const unsigned long long nlists = (unsigned long long)in->nx * in->ny;
	cudaStream_t s = ctx->stream;
	return run_staged(ctx, a, nlists, 65536ull + nlists / 8,
		[&](BruteArgs &g) { k_brute<CAP_FAST><<<blocks_for(g.wk.n, 128), 128, 0, s>>>(g); },
		[&](BruteArgs &g, unsigned int grid) {
			if (ctx->huge) k_brute<CAP_HUGE><<<HUGE_GRID, 128, 0, s>>>(g);
			else k_brute<CAP_BIG><<<grid, 128, 0, s>>>(g);
		},
		in->nx, in->ny, out);
}

struct PassTimes { double ms1 = 0, ms2 = 0; };

// dilation of a resident volume; fills the per-pass device times
int dilate_once(vo_ctx *ctx, int method, const vo_dvol *in, double R, vo_dvol **out, PassTimes *pt, double clip_lo, double clip_hi)
{
	float t1 = 0, t2 = 0;
	VO_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
	if (method == VO_METHOD_OURS) {
		for (int attempt = 0;; ++attempt) {
			vo_dmid *mid = nullptr;
			VO_TRY(pass1(ctx, in, R, &mid, clip_lo, clip_hi, attempt == 0 && !ctx->huge));
			cudaEventRecord(ctx->ev[1], ctx->stream);
			const bool deferred = mid->deferred, redo_skipped = mid->redo_skipped;
			const uint64_t pool_cap = mid->pool_cap;
			const unsigned int redo_cap = mid->redo_cap;
			int rc = pass2(ctx, mid, 0, mid->ny, out, ctx->ev[2]);   // (ev[2] behind the last kernel: run_staged has synchronised past it)
			vo_dmid_free(ctx, mid);
			if (deferred) {
				// pass 2's synchronisation has read every counter: was the deferred pass 1 complete?
				const unsigned long long *h = ctx->last_ctr;
				const bool short1 = h[0] > pool_cap || h[2] > redo_cap || h[4] != 0 || (redo_skipped && h[2] != 0);
				if (redo_skipped && h[2] != 0) ctx->redo_recent = 16;
				note_pass1_redo(ctx, h[2]);
				if (rc == VO_OK && !short1) {
					ctx->pool_hint = next_hint(ctx->pool_hint, h[0]);
					if (h[5]) ctx->pooled_per_column = (double)h[0] / (double)std::max<unsigned long long>(1, (unsigned long long)in->nx * in->ny);
					if (h[5] && std::getenv("VO_TRACE")) std::fprintf(stderr, "[vo trace] mid-pool entries per column %.2f\n", ctx->pooled_per_column);
					break;
				}
				if (rc == VO_OK) { free_dvol(ctx, *out); *out = nullptr; }
				else if (!short1) return rc;
				// repeat, synchronously this time (it regrows its pool / reports what cannot be done)
				if (h[0] > pool_cap) ctx->pool_hint = std::max<unsigned long long>(ctx->pool_hint, h[0] + h[0] / 4);
				ctx->err.clear();
				continue;
			}
			VO_TRY(rc);
			break;
		}
		cudaEventElapsedTime(&t1, ctx->ev[0], ctx->ev[1]);
		cudaEventElapsedTime(&t2, ctx->ev[1], ctx->ev[2]);
	} else if (method == VO_METHOD_BRUTE_FORCE) {
		VO_TRY(brute(ctx, in, R, out));
		VO_CUDA(cudaEventRecord(ctx->ev[2], ctx->stream));
		VO_CUDA(cudaEventSynchronize(ctx->ev[2]));
		cudaEventElapsedTime(&t1, ctx->ev[0], ctx->ev[2]);
	} else {
		return fail(ctx, VO_ERR_ARG, "Invalid method");
	}
	if (pt) { pt->ms1 = t1; pt->ms2 = t2; }
	return VO_OK;
}

int dilate(vo_ctx *ctx, int method, const vo_dvol *in, double R, vo_dvol **out, PassTimes *pt,
           double clip_lo = -std::numeric_limits<double>::infinity(), double clip_hi = std::numeric_limits<double>::infinity())
{
	return with_huge_lists(ctx, [&] { return dilate_once(ctx, method, in, R, out, pt, clip_lo, clip_hi); });
}

// negate (Voronoi.cpp:18-55) / negateInv (Voronoi.cpp:57-89) / vor2d negate. Fused (default): ONE launch (count ->
// look-back prefix sum -> fill, k_complement_fused) into a buffer sized by the bound "intervals + 1 per column", one
// synchronisation for the exact total. Classic: count -> scan -> fill.
template <typename Op>
int complement_fused(vo_ctx *ctx, typename Op::Args &a, unsigned long long nlists, unsigned long long cap, vo_dvol *v,
                     unsigned int *d_flag, unsigned int *h_flag)
{
	VO_TRY(dalloc(ctx, &v->spans, cap));
	const unsigned int ntiles = blocks_for(nlists, CF_TILE);
	Tmp<unsigned long long> tot(ctx);
	VO_TRY(dalloc(ctx, &tot.p, 1));
	uint32_t epoch = 0;
	unsigned long long tbase = 0, total = 0;
	VO_TRY(scan_prepare(ctx, ntiles, &epoch, &tbase));
	k_complement_fused<Op><<<ntiles, SCAN_THREADS, 0, ctx->stream>>>(a, nlists, v->off, v->spans, cap, ctx->scan_state,
	                                                                 ctx->scan_state + ctx->scan_cap, tbase, epoch, tot.p);
	ctx->launches++;
	VO_CUDA(cudaGetLastError());
	VO_CUDA(cudaMemcpyAsync(&total, tot.p, sizeof(total), cudaMemcpyDeviceToHost, ctx->stream));
	if (d_flag && h_flag) VO_CUDA(cudaMemcpyAsync(h_flag, d_flag, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
	VO_CUDA(cudaStreamSynchronize(ctx->stream));
	if (total > cap) return fail(ctx, VO_ERR_OVERFLOW, "complement larger than its bound");
	v->nspans = total;
	return VO_OK;
}

// by0 / by1: border rows before / after (default: `border` on every side; a y-slab of a sharded grid passes the rows of
// the global border it owns). h_outside (optional): receives the "data outside [lo, hi]" flag of `outside`.
int negate(vo_ctx *ctx, const vo_dvol *in, int border, double lo, double hi, vo_dvol **out, unsigned int *outside = nullptr,
           int by0 = -1, int by1 = -1, unsigned int *h_outside = nullptr)
{
	if (by0 < 0) by0 = border;
	if (by1 < 0) by1 = border;
	const int mx = in->nx + 2 * border, my = in->ny + by0 + by1;
	VO_TRY(check_dims(ctx, mx, my));
	const unsigned long long nlists = (unsigned long long)mx * my;
	vo_dvol *v = nullptr;
	VO_TRY(new_dvol(ctx, mx, my, &v));
	NegArgs a;
	a.nx = in->nx; a.ny = in->ny; a.border = border; a.by0 = by0; a.by1 = by1; a.lo = lo; a.hi = hi;
	a.off = in->off; a.spans = in->spans; a.cnt = nullptr; a.out_off = nullptr; a.out_spans = nullptr; a.outside = outside;
	if (outside) cudaMemsetAsync(outside, 0, sizeof(unsigned int), ctx->stream);
	if (h_outside) *h_outside = 0;
	if (ctx->fused_scan && nlists && in->nspans + nlists < (1ull << 32)) {
		const int rc = complement_fused<NegOp>(ctx, a, nlists, in->nspans + nlists, v, outside, h_outside);
		if (rc) { free_dvol(ctx, v); return rc; }
		v->max_cnt = in->max_cnt < 0 ? -1 : in->max_cnt + 1;
		*out = v;
		return VO_OK;
	}
	Tmp<uint32_t> cnt(ctx);
	int rc = dalloc(ctx, &cnt.p, nlists);
	if (rc) { free_dvol(ctx, v); return rc; }
	a.cnt = cnt.p;
	if (nlists) { k_negate<false><<<blocks_for(nlists, 256), 256, 0, ctx->stream>>>(a, nlists); ctx->launches++; }
	unsigned long long total = 0;
	rc = scan_counts(ctx, cnt.p, nlists, v->off, &total);
	if (rc == VO_OK) rc = dalloc(ctx, &v->spans, total);
	if (rc) { free_dvol(ctx, v); return rc; }
	v->nspans = total;
	a.out_off = v->off; a.out_spans = v->spans;
	if (nlists) { k_negate<true><<<blocks_for(nlists, 256), 256, 0, ctx->stream>>>(a, nlists); ctx->launches++; }
	cudaError_t e = cudaGetLastError();
	if (e == cudaSuccess && outside && h_outside) {
		e = cudaMemcpyAsync(h_outside, outside, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream);
		if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
	}
	if (e != cudaSuccess) { free_dvol(ctx, v); return fail(ctx, VO_ERR_CUDA, std::string("k_negate: ") + cudaGetErrorString(e)); }
	*out = v;
	return VO_OK;
}

int negate_inv(vo_ctx *ctx, const vo_dvol *in, int border, double lo, double hi, vo_dvol **out, int by0 = -1, int by1 = -1)
{
	if (by0 < 0) by0 = border;
	if (by1 < 0) by1 = border;
	const int nx = in->nx - 2 * border, ny = in->ny - by0 - by1;
	if (nx < 0 || ny < 0) return fail(ctx, VO_ERR_ARG, "negateInv on a grid smaller than its border");
	const unsigned long long nlists = (unsigned long long)nx * ny;
	vo_dvol *v = nullptr;
	VO_TRY(new_dvol(ctx, nx, ny, &v));
	NegInvArgs a;
	a.mx = in->nx; a.my = in->ny; a.border = border; a.by0 = by0; a.lo = lo; a.hi = hi;
	a.off = in->off; a.spans = in->spans; a.cnt = nullptr; a.out_off = nullptr; a.out_spans = nullptr;
	if (ctx->fused_scan && nlists && in->nspans + nlists < (1ull << 32)) {
		const int rc = complement_fused<NegInvOp>(ctx, a, nlists, in->nspans + nlists, v, nullptr, nullptr);
		if (rc) { free_dvol(ctx, v); return rc; }
		v->max_cnt = in->max_cnt < 0 ? -1 : in->max_cnt + 1;
		*out = v;
		return VO_OK;
	}
	Tmp<uint32_t> cnt(ctx);
	int rc = dalloc(ctx, &cnt.p, nlists);
	if (rc) { free_dvol(ctx, v); return rc; }
	a.cnt = cnt.p;
	if (nlists) { k_negate_inv<false><<<blocks_for(nlists, 256), 256, 0, ctx->stream>>>(a, nlists); ctx->launches++; }
	unsigned long long total = 0;
	rc = scan_counts(ctx, cnt.p, nlists, v->off, &total);
	if (rc == VO_OK) rc = dalloc(ctx, &v->spans, total);
	if (rc) { free_dvol(ctx, v); return rc; }
	v->nspans = total;
	a.out_off = v->off; a.out_spans = v->spans;
	if (nlists) { k_negate_inv<true><<<blocks_for(nlists, 256), 256, 0, ctx->stream>>>(a, nlists); ctx->launches++; }
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) { free_dvol(ctx, v); return fail(ctx, VO_ERR_CUDA, std::string("k_negate_inv: ") + cudaGetErrorString(e)); }
	*out = v;
	return VO_OK;
}

// erosion = negate, dilation, negateInv (Voronoi.cpp:8-17). `dil(neg, clip_lo, clip_hi, unpruned, &out)` dilates the
// complement; by0 / by1 = rows of the one-line border this volume owns (1 / 1 for a whole grid; a y-slab of a sharded grid
// only has the border rows at the ends of the global grid).
template <typename DilateFn>
int erode_with(vo_ctx *ctx, const vo_dvol *in, double zmin, double zmax, int by0, int by1, DilateFn dil, vo_dvol **out)
{
	const double z_min = zmin - 1, z_max = zmax + 1;
	vo_dvol *neg = nullptr, *dl = nullptr;
	unsigned int *outside = reinterpret_cast<unsigned int *>(ctx->d_ctr + NCTR);
	unsigned int h_outside = 0;
	VO_TRY(negate(ctx, in, 1, z_min, z_max, &neg, outside, by0, by1, &h_outside));
	// negateInv only reads the dilated complement inside (z_min + 1, z_max - 1): everything at or beyond those bounds
	// is dropped (MorphologyOperators.cpp:292-312), so 'ours' may prune with that clip range (pass1_tile.cuh: nn_of)
	// Data outside [z_min, z_max] (no head-room: offset3d's -p) turns the complement into something that is not a
	// set of intervals (negate_ray prepends / appends the bounds without looking); the reference still computes with
	// it. k_negate raises a flag then, and that dilation takes the unpruned one-thread-per-slot kernel, which folds
	// whatever it is given exactly like the reference's unions do.
	int rc = dil(neg, z_min + 1, z_max - 1, h_outside != 0, &dl);
	free_dvol(ctx, neg);
	VO_TRY(rc);
	rc = negate_inv(ctx, dl, 1, z_min + 1, z_max - 1, out, by0, by1);
	free_dvol(ctx, dl);
	return rc;
}

// Erosion in DUAL form ('ours'; kernels.cuh: k_pass2_rows_dual has the derivation): for a volume whose columns hold at
// most one interval, strictly inside (zmin - 1, zmax + 1), the reference's complement - dilate - complement
// (Voronoi.cpp:8-17) is, column by column, the intersection of the eroded intervals [a_q + h_q, b_q - h_q] over the
// pairs a dilation visits, and empty where an empty column (or the one-line border) lies in reach. The tile kernel
// computes that intersection as the hull of the mirrored intervals with the same tables and the same pruning, so the
// erosion costs one dilation-sized pass instead of a dilation of the (denser, two-layer) complement between two
// complement kernels. Whether the input qualifies is checked by k_thresh while it runs (no separate pass, no
// synchronisation): DUAL_NA = it did not (or cannot be known to), the caller takes the general path.
constexpr int DUAL_NA = -2;
// (y0, y1: only these rows of the result - a y-slab that was handed its halo rows; default: all)
int erode_dual(vo_ctx *ctx, const vo_dvol *in, double zmin, double zmax, double R, vo_dvol **out, PassTimes *pt, int y0 = 0, int y1 = -1)
{
	const unsigned long long ncols = (unsigned long long)in->nx * in->ny;
	if (in->dual_state == 2 || ncols == 0 || in->nspans > ncols || ctx->force_simple_pass1 || (in->max_cnt > 1)) return DUAL_NA;
	if (check_radius(ctx, R) != VO_OK || !pass1_uses_tile(ctx, in, R)) { ctx->err.clear(); return DUAL_NA; }
	float t1 = 0, t2 = 0;
	VO_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
	Tmp<uint8_t> dist(ctx);
	const unsigned long long ntmin = (unsigned long long)in->ny * ((in->nx + P2_TX - 1) / P2_TX);     // (tile minima behind the distances)
	VO_TRY(dalloc(ctx, &dist.p, ncols + ntmin));
	k_empty_dist<<<(unsigned int)in->ny, ED_THREADS, (size_t)((in->nx + 31) / 32 + (in->nx + P2_TX - 1) / P2_TX) * sizeof(uint32_t), ctx->stream>>>(in->off, in->nx, dist.p, dist.p + ncols);
	ctx->launches++;
	vo_dmid *mid = nullptr;
	const DualSpec ds{zmin - 1, zmax + 1};
	const double inf = std::numeric_limits<double>::infinity();
	VO_TRY(pass1(ctx, in, R, &mid, -inf, inf, true, &ds));
	cudaEventRecord(ctx->ev[1], ctx->stream);
	const unsigned int redo_cap = mid->redo_cap;
	TableCache *tc = static_cast<TableCache *>(ctx->table_cache);    // (pass 1 has just made these the current tables)
	int rc = pass2_dual(ctx, mid, dist.p, tc->dt.reach, zmin, zmax, out, ctx->ev[2], y0, y1);
	vo_dmid_free(ctx, mid);
	VO_TRY(rc);
	// pass 2's synchronisation has read every counter: [14] a column did not qualify, [5] a tile saw a multi-interval
	// column, [2] anything was left to the (non-dual) redo kernel
	const unsigned long long *h = ctx->last_ctr;
	(void)redo_cap;
	if (h[14] != 0 || h[5] != 0 || h[2] != 0 || h[4] != 0) {
		free_dvol(ctx, *out);
		*out = nullptr;
		in->dual_state = 2;
		return DUAL_NA;
	}
	in->dual_state = 1;
	ctx->dual_erosions++;
	cudaEventElapsedTime(&t1, ctx->ev[0], ctx->ev[1]);
	cudaEventElapsedTime(&t2, ctx->ev[1], ctx->ev[2]);
	if (pt) { pt->ms1 = t1; pt->ms2 = t2; }
	return VO_OK;
}

int erode(vo_ctx *ctx, int method, const vo_dvol *in, double zmin, double zmax, double R, vo_dvol **out, PassTimes *pt)
{
	if (method == VO_METHOD_OURS && ctx->erosion_mode != 2) {
		const int rc = erode_dual(ctx, in, zmin, zmax, R, out, pt);
		if (rc != DUAL_NA) return rc;
		if (ctx->erosion_mode == 1) return fail(ctx, VO_ERR_ARG, "erosion = dual: the input does not qualify for the dual form");
	}
	return erode_with(ctx, in, zmin, zmax, 1, 1, [&](const vo_dvol *neg, double clo, double chi, bool unpruned, vo_dvol **o) {
		const bool saved_simple = ctx->force_simple_pass1;
		if (unpruned) ctx->force_simple_pass1 = true;
		const int rc = dilate(ctx, method, neg, R, o, pt, clo, chi);
		ctx->force_simple_pass1 = saved_simple;
		return rc;
	}, out);
}

int morph3d_dev(vo_ctx *ctx, int op, int method, const vo_dvol *in, double zmin, double zmax, double R,
                vo_dvol **out, PassTimes *pt)
{
	if (method != VO_METHOD_OURS && method != VO_METHOD_BRUTE_FORCE) return fail(ctx, VO_ERR_ARG, "Invalid method");
	switch (op) {
	case VO_OP_DILATION: return dilate(ctx, method, in, R, out, pt);
	case VO_OP_EROSION: return erode(ctx, method, in, zmin, zmax, R, out, pt);
	case VO_OP_OPENING: {   // offset3d.cpp:129-133
		vo_dvol *tmp = nullptr;
		VO_TRY(erode(ctx, method, in, zmin, zmax, R, &tmp, pt));
		int rc = dilate(ctx, method, tmp, R, out, pt);
		free_dvol(ctx, tmp);
		return rc;
	}
	case VO_OP_CLOSING: {   // offset3d.cpp:124-128
		vo_dvol *tmp = nullptr;
		VO_TRY(dilate(ctx, method, in, R, &tmp, pt));
		int rc = erode(ctx, method, tmp, zmin, zmax, R, out, pt);
		free_dvol(ctx, tmp);
		return rc;
	}
	default: return fail(ctx, VO_ERR_ARG, "Operation");
	}
}

// vor2d: rows live in a vo_dvol with nx = rows, ny = 1
int dilate2d(vo_ctx *ctx, const vo_dvol *in, int width, double R, int complement, vo_dvol **out)
{
	// (R = r * rows for a dilation, DoubleCompressedImage.cpp:685-686: easily thousands of rows. Only rows of the image -
	// and, for the erosion sweep, the two sentinel rows next to it - contribute: the table ends there, the radius itself
	// is not limited like the 3D one)
	if (!(R >= 0.0) || !(R < 1.0e9)) return fail(ctx, VO_ERR_ARG, "radius must be in [0, 1e9) rows");
	const int J = (int)std::min<double>(std::floor(R), (double)in->nx + 1.0);
	std::vector<double> h2((size_t)J + 1);
	for (int di = 0; di <= J; ++di) {
		const double d = (double)di;
		volatile double a = R * R;
		volatile double b = d * d;
		volatile double c = a - b;
		h2[di] = std::sqrt(c);     // DoubleVoronoi.cpp:718
	}
	Tmp<double> dh(ctx);
	VO_TRY(dalloc(ctx, &dh.p, (unsigned long long)J + 1));
	VO_CUDA(cudaMemcpyAsync(dh.p, h2.data(), h2.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
	VO_CUDA(cudaStreamSynchronize(ctx->stream));
	Dil2dArgs a;
	a.rows = in->nx; a.J = J; a.complement = complement; a.W = (double)width;
	a.off = in->off; a.spans = in->spans; a.h2 = dh.p;
	const unsigned long long nlists = (unsigned long long)in->nx;
	cudaStream_t s = ctx->stream;
	return with_huge_lists(ctx, [&] { return run_staged(ctx, a, nlists, 65536ull + 4 * in->nspans,
		[&](Dil2dArgs &g) { k_dilate2d_block<<<(unsigned int)g.wk.n, D2B_THREADS, 0, s>>>(g); },   // CTA per row; overflow -> redo
		[&](Dil2dArgs &g, unsigned int grid) {
			if (ctx->huge) k_dilate2d<CAP_HUGE><<<2 * HUGE_GRID, 64, 0, s>>>(g);
			else k_dilate2d<CAP_BIG><<<grid, 64, 0, s>>>(g);
		},
		in->nx, 1, out); });
}

int morph2d_dev(vo_ctx *ctx, int op, const vo_dvol *in, int width, double r, vo_dvol **out)
{
	if (in->ny != 1 && in->nx != 0) return fail(ctx, VO_ERR_ARG, "2D rows must be passed as an nx = rows, ny = 1 volume");
	switch (op) {
	case VO_OP2D_DILATE:   // DoubleCompressedImage.cpp:680-689: R = r * rows
		return dilate2d(ctx, in, width, r * (double)(size_t)in->nx, 0, out);
	case VO_OP2D_ERODE: {  // DoubleCompressedImage.cpp:693-703: R = r, then negate()
		vo_dvol *tmp = nullptr;
		VO_TRY(dilate2d(ctx, in, width, r, 1, &tmp));
		int rc = negate(ctx, tmp, 0, 0.0, (double)width, out);
		free_dvol(ctx, tmp);
		return rc;
	}
	case VO_OP2D_NEGATE: return negate(ctx, in, 0, 0.0, (double)width, out);
	case VO_OP2D_OPEN: {   // erode then dilate (DoubleCompressedImage.cpp:715-719)
		vo_dvol *tmp = nullptr;
		VO_TRY(morph2d_dev(ctx, VO_OP2D_ERODE, in, width, r, &tmp));
		int rc = morph2d_dev(ctx, VO_OP2D_DILATE, tmp, width, r, out);
		free_dvol(ctx, tmp);
		return rc;
	}
	case VO_OP2D_CLOSE: {  // dilate then erode (DoubleCompressedImage.cpp:707-711)
		vo_dvol *tmp = nullptr;
		VO_TRY(morph2d_dev(ctx, VO_OP2D_DILATE, in, width, r, &tmp));
		int rc = morph2d_dev(ctx, VO_OP2D_ERODE, tmp, width, r, out);
		free_dvol(ctx, tmp);
		return rc;
	}
	default: return fail(ctx, VO_ERR_ARG, "Operation");
	}
}

int xor_dev(vo_ctx *ctx, const vo_dvol *A, const vo_dvol *B, double zmin, double zmax, double spacing,
            vo_dvol **out, double *volume)
{
	if (A->nx != B->nx || A->ny != B->ny) return fail(ctx, VO_ERR_ARG, "xor needs two volumes on the same grid");
	const unsigned long long nlists = (unsigned long long)A->nx * A->ny;
	vo_dvol *v = nullptr;
	VO_TRY(new_dvol(ctx, A->nx, A->ny, &v));
	Tmp<uint32_t> cnt(ctx);
	Tmp<double> len(ctx), part(ctx);
	int rc = dalloc(ctx, &cnt.p, nlists);
	if (rc == VO_OK) rc = dalloc(ctx, &len.p, nlists);
	const unsigned int nb = 256;
	if (rc == VO_OK) rc = dalloc(ctx, &part.p, nb);
	if (rc) { free_dvol(ctx, v); return rc; }
	XorArgs a;
	a.off_a = A->off; a.sp_a = A->spans; a.off_b = B->off; a.sp_b = B->spans; a.lo = zmin; a.hi = zmax;
	a.cnt = cnt.p; a.out_off = nullptr; a.out_spans = nullptr; a.col_len = len.p;
	if (nlists) { k_xor<false><<<blocks_for(nlists, 256), 256, 0, ctx->stream>>>(a, nlists); ctx->launches++; }
	unsigned long long total = 0;
	rc = scan_counts(ctx, cnt.p, nlists, v->off, &total);
	if (rc == VO_OK) rc = dalloc(ctx, &v->spans, total);
	if (rc) { free_dvol(ctx, v); return rc; }
	v->nspans = total;
	a.out_off = v->off; a.out_spans = v->spans;
	double vol = 0;
	if (nlists) {
		k_xor<true><<<blocks_for(nlists, 256), 256, 0, ctx->stream>>>(a, nlists);
		k_sum<<<nb, 256, 0, ctx->stream>>>(len.p, nlists, part.p);
		ctx->launches += 2;
		std::vector<double> hp(nb);
		cudaError_t e = cudaMemcpyAsync(hp.data(), part.p, nb * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
		if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
		if (e != cudaSuccess) { free_dvol(ctx, v); return fail(ctx, VO_ERR_CUDA, std::string("k_xor: ") + cudaGetErrorString(e)); }
		for (double p : hp) vol += p;
	}
	if (volume) *volume = spacing * spacing * spacing * vol;
	*out = v;
	return VO_OK;
}

// CSR offsets must be non-decreasing (every kernel indexes spans[] with them). Branch-free so that it vectorises:
// ~1.5 ms for the 4.2 M columns of a 2048^2 grid, hidden behind the upload it guards.
bool offsets_sorted(const uint32_t *off, unsigned long long n, uint32_t *max_count = nullptr)
{
	uint32_t bad = 0, mx = 0;
	for (unsigned long long i = 0; i < n; ++i) {
		bad |= (uint32_t)(off[i] > off[i + 1]);
		const uint32_t c = off[i + 1] - off[i];
		mx = c > mx ? c : mx;
	}
	if (max_count) *max_count = mx;
	return bad == 0;
}

int upload(vo_ctx *ctx, int nx, int ny, const uint32_t *off, const double *spans, vo_dvol **out)
{
	VO_TRY(check_dims(ctx, nx, ny));
	const unsigned long long n = (unsigned long long)nx * ny;
	if (!off) return fail(ctx, VO_ERR_ARG, "off is NULL");
	if (off[0] != 0) return fail(ctx, VO_ERR_ARG, "off[0] must be 0");
	const uint64_t m = off[n];
	if (m && !spans) return fail(ctx, VO_ERR_ARG, "spans is NULL");
	vo_dvol *v = nullptr;
	VO_TRY(new_dvol(ctx, nx, ny, &v));
	int rc = dalloc(ctx, &v->spans, m);
	if (rc) { free_dvol(ctx, v); return rc; }
	v->nspans = m;
	// (the copies only need off[n]: the check below runs on the host while they are in flight, and nothing is
	// launched on the volume before it has passed)
	cudaError_t e = cudaMemcpyAsync(v->off, off, (n + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream);
	if (e == cudaSuccess && m) e = cudaMemcpyAsync(v->spans, spans, m * sizeof(double2), cudaMemcpyHostToDevice, ctx->stream);
	if (e != cudaSuccess) { free_dvol(ctx, v); cudaGetLastError(); return fail(ctx, VO_ERR_CUDA, std::string("upload: ") + cudaGetErrorString(e)); }
	uint32_t max_count = 0;
	if (!offsets_sorted(off, n, &max_count)) {
		cudaStreamSynchronize(ctx->stream);
		free_dvol(ctx, v);
		return fail(ctx, VO_ERR_ARG, "offsets must be non-decreasing");
	}
	v->max_cnt = max_count;
	*out = v;
	return VO_OK;
}

int download_new(vo_ctx *ctx, const vo_dvol *v, uint32_t **out_off, double **out_spans, uint64_t *out_nspans)
{
	const unsigned long long n = (unsigned long long)v->nx * v->ny;
	uint32_t *ho = (uint32_t *)host_block((n + 1) * sizeof(uint32_t));
	double *hs = (double *)host_block(std::max<uint64_t>(v->nspans, 1) * sizeof(double2));
	if (!ho || !hs) { vo_free(ho); vo_free(hs); return fail(ctx, VO_ERR_NOMEM, "pinned host allocation failed"); }
	cudaError_t e = cudaMemcpyAsync(ho, v->off, (n + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream);
	if (e == cudaSuccess && v->nspans) e = cudaMemcpyAsync(hs, v->spans, v->nspans * sizeof(double2), cudaMemcpyDeviceToHost, ctx->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
	if (e != cudaSuccess) { vo_free(ho); vo_free(hs); cudaGetLastError(); return fail(ctx, VO_ERR_CUDA, std::string("download: ") + cudaGetErrorString(e)); }
	*out_off = ho; *out_spans = hs;
	if (out_nspans) *out_nspans = v->nspans;
	return VO_OK;
}

// ---------------------------------------------------------------------------------------------------
// Pipelined host-buffer dilation ('ours'): the grid is cut into bands of rows and the three stages
//   H2D of band b+1  |  pass 1 of band b, pass 2 + compaction of band b-1  |  D2H of band b-2
// run concurrently on three streams, so the PCIe transfers hide behind the kernels (and vice versa).
// Pass 1 of a row needs only that row (its y-thresholds also the two neighbouring rows), pass 2 of a row
// needs the mid rows within floor(R): band b-1 can be finished as soon as pass 1 of band b is enqueued.
// Results are identical to the plain path (same kernels, same tables). Returns PIPE_NA when the case is
// not worth / not able to be pipelined (small grids, simple-kernel cases, any pool overflow): the caller
// then takes the plain path.
// ---------------------------------------------------------------------------------------------------
constexpr int PIPE_NA = -1;

// copy streams and events live in the context (created on first use, destroyed with it)
struct PipeRes {
	vo_ctx *ctx;
	cudaStream_t s_in = nullptr, s_out = nullptr;
	size_t used = 0;
	explicit PipeRes(vo_ctx *c) : ctx(c) {}
	~PipeRes()
	{
		if (s_in) cudaStreamSynchronize(s_in);
		if (s_out) cudaStreamSynchronize(s_out);
	}
	bool init()
	{
		if (!ctx->s_in && cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking) != cudaSuccess) return false;
		if (!ctx->s_out && cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking) != cudaSuccess) return false;
		s_in = ctx->s_in; s_out = ctx->s_out;
		return true;
	}
	cudaEvent_t event()
	{
		if (used == ctx->pipe_ev.size()) {
			cudaEvent_t e = nullptr;
			cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
			ctx->pipe_ev.push_back(e);
		}
		return ctx->pipe_ev[used++];
	}
};

// Row window (keep0, keep1; default: everything): only the rows [keep0, keep1) of the result are produced and returned
// (offsets starting at 0) - a y-slab of a larger grid passed WITH its floor(R) ghost rows on either side, cut from host
// memory (vo_morph3d_rows: the multi-GPU host-buffer call, where the host holds every slab's neighbours anyway, so the
// halo needs no device-to-device exchange). `off` may then point into the middle of a larger CSR: the offsets are taken
// as they are (off[0] != 0) and `spans` is the base of the whole span array.
int dilate_ours_pipelined(vo_ctx *ctx, int nx, int ny, const uint32_t *off, const double *spans, double R,
                          uint32_t **out_off, double **out_spans, uint64_t *out_nspans, PassTimes *pt, int keep0 = 0, int keep1 = -1)
{
	if (ctx->force_simple_pass1 || ctx->no_pipeline) return PIPE_NA;
	if (check_radius(ctx, R) != VO_OK) return PIPE_NA;             // let the plain path report it
	const int J = (int)std::floor(R);
	const unsigned long long ncols = (unsigned long long)nx * ny;
	if (keep1 < 0) keep1 = ny;
	if (nx <= 0 || ny <= 0 || !off || ncols >= (1ull << 32) - 8 || keep0 < 0 || keep1 > ny || keep0 >= keep1) return PIPE_NA;
	const bool whole = keep0 == 0 && keep1 == ny;
	if (whole && off[0] != 0) return PIPE_NA;
	const uint32_t obase = off[0];                                 // the device arrays are indexed by the offsets as they are: pointers shifted by this
	if (off[ncols] < obase) return fail(ctx, VO_ERR_ARG, "offsets must be non-decreasing");
	const uint64_t nspans = off[ncols] - obase;
	const double k_in = (double)nspans / (double)ncols;
	// Bands of rows: eight by default (vo_set_option("bands", N)) - enough to overlap the copies with the passes, few
	// enough that a launch set of pass 1 still has many tiles per warp. (Small bands at both ends and large ones in
	// between, or a first band half as tall - early first download, short last stage - were tried and give nothing: a
	// band's chain of stream-ordered launches costs the same whatever its height)
	// Every band is at least 2 (floor(R) + 1) rows high (pass 2 of a band then only reaches into its two neighbours).
	std::vector<int> ys;
	{
		const int hmin = 2 * (J + 1);
		std::vector<int> wts;
		wts.assign(std::max(3, ctx->pipe_bands), 1);
		if (ctx->pipe_wts.size() >= 3) wts = ctx->pipe_wts;
		int wsum = 0;
		for (int w : wts) wsum += w;
		if (ny / wsum < hmin) {                              // too few rows for that shape: equal bands of the minimum height or more
			const int n = std::max(1, std::min((int)wts.size(), ny / hmin));
			wts.assign(n, 1); wsum = n;
		}
		ys.push_back(0);
		int acc = 0;
		for (size_t i = 0; i + 1 < wts.size(); ++i) {
			acc += wts[i];
			const int y = std::min(ny, (int)(((long long)ny * acc / wsum + 7) & ~7ll));
			if (y - ys.back() >= hmin && ny - y >= hmin) ys.push_back(y);
		}
		ys.push_back(ny);
	}
	const int nb = (int)ys.size() - 1;
	int BH = 0;                                              // the tallest band
	for (int b = 0; b < nb; ++b) BH = std::max(BH, ys[b + 1] - ys[b]);
	// The second half (pass 2 .. download) works on bands shifted UP by floor(R) rows: its rows [ys2[b], ys2[b+1]) only
	// read mid rows below ys[b+1], i.e. pass 1 of the bands 0 .. b - a band's result no longer waits for pass 1 of the
	// band behind it. (The last one takes the floor(R) rows that are left over.)
	std::vector<int> ys2(ys);
	for (int b = 1; b < nb; ++b) ys2[b] = ys[b] - J;
	int BH2 = 0;
	for (int b = 0; b < nb; ++b) BH2 = std::max(BH2, ys2[b + 1] - ys2[b]);
	if (nb < 3 || ncols * (unsigned long long)(J + 1) < (48ull << 20) || !TilePlan::fits(J, k_in)) return PIPE_NA;
	if (nspans && !spans) return PIPE_NA;
	// The offsets have not been validated (upload() does that for the plain path): the band boundaries are checked here, so
	// that every copy below stays inside the buffers; everything in between is checked on the device by k_thresh, which
	// reads every offset anyway (ThreshArgs::bad) - a host loop over 4 M offsets would cost more than a band's upload.
	for (int b = 0; b <= nb; ++b) {
		const uint32_t o = off[(unsigned long long)ys[b] * nx];
		if (o < obase || o - obase > nspans || (b > 0 && o < off[(unsigned long long)ys[b - 1] * nx])) return fail(ctx, VO_ERR_ARG, "offsets must be non-decreasing");
	}
	for (int b = 1; b < nb; ++b) {                           // (a band's upload ends one row below it)
		const unsigned long long c = (unsigned long long)std::min(ny, ys[b] + 1) * nx;
		if (off[c] - obase > nspans || off[c] < off[(unsigned long long)ys[b] * nx] || off[c] > off[(unsigned long long)ys[b + 1] * nx])
			return fail(ctx, VO_ERR_ARG, "offsets must be non-decreasing");
	}

	TableCache *tc = nullptr;
	VO_TRY(get_tables(ctx, R, true, &tc));
	DevTables &dt = tc->dt;
	TileTables &tt = tc->tt;

	PipeRes pr(ctx);
	if (!pr.init()) { cudaGetLastError(); return PIPE_NA; }

	// device input, mid volume, scratch
	vo_dvol *in = nullptr;
	VO_TRY(new_dvol(ctx, nx, ny, &in));
	struct InGuard { vo_ctx *c; vo_dvol *p; ~InGuard() { free_dvol(c, p); } } in_guard{ctx, in};
	// (a row window starts at some interval obase of the caller's array: the device copy starts obase % 16 entries into its
	// block, so that host and device addresses of an interval are aligned alike - vo_ctx::copy_align)
	const uint32_t in_pad = obase % 16u;
	VO_TRY(dalloc(ctx, &in->spans, nspans + 16));
	in->nspans = nspans;
	vo_dmid *m = new (std::nothrow) vo_dmid();
	if (!m) return fail(ctx, VO_ERR_NOMEM, "out of host memory");
	struct MidGuard { vo_ctx *c; vo_dmid *p; ~MidGuard() { vo_dmid_free(c, p); } } mid_guard{ctx, m};
	m->nx = nx; m->ny = ny; m->J = J; m->R = R;
	const unsigned long long nslots = ncols * (J + 1);
	VO_TRY(dalloc(ctx, &m->slots, nslots));
	m->pool_cap = std::max(65536ull + (unsigned long long)(J + 1) * (nspans / 4), ctx->pool_hint);
	VO_TRY(dalloc(ctx, &m->pool, m->pool_cap));
	VO_TRY(dalloc(ctx, &m->flags, 2 * ncols));
	Tmp<uint4> thr(ctx);
	VO_TRY(dalloc(ctx, &thr.p, nspans));
	double2 *const d_sp = in->spans + in_pad - obase;       // indexed by the host's offsets as they are
	uint4 *const d_thr = thr.p - obase;
	TilePlan plan, plan0;
	// (the first band's plan first: init() leaves the kernels' shared-memory limit at what the LAST call asked for, the larger one)
	if (ctx->pipe_warps0 > 0) VO_TRY(plan0.init(ctx, nx, J, k_in, ctx->pipe_warps0, ctx->pipe_ctas, 0, 0));
	VO_TRY(plan.init(ctx, nx, J, k_in, ctx->pipe_warps, ctx->pipe_ctas, ctx->pipe_quota,
	                 (unsigned int)(((nx + P1_W - 1) / P1_W) * (unsigned int)BH)));
	if (ctx->pipe_warps0 <= 0) plan0 = plan;
	const int tiles_x = plan.tiles_x;
	VO_TRY(dalloc(ctx, &m->tilemask, 2ull * ny * tiles_x));
	const unsigned long long ntiles = (unsigned long long)plan.tiles_xw * ny;
	Tmp<unsigned int> big_tiles(ctx), multi_tiles(ctx), big_tiles1(ctx), multi_tiles1(ctx);
	VO_TRY(dalloc(ctx, &big_tiles.p, ntiles));
	VO_TRY(dalloc(ctx, &multi_tiles.p, ntiles));
	VO_TRY(dalloc(ctx, &big_tiles1.p, ntiles));
	VO_TRY(dalloc(ctx, &multi_tiles1.p, ntiles));
	// tile order per band (expensive first): [per band 2 * P1_NBUCKET counters | cost estimate per tile], permutations per stream
	Tmp<unsigned int> est(ctx), order0(ctx), order1(ctx);
	const bool ordered = ctx->tile_order;
	if (ordered) {
		VO_TRY(dalloc(ctx, &est.p, ntiles + 2ull * P1_NBUCKET * nb));      // [nb][2 * P1_NBUCKET] bucket counters | cost estimate per tile
		VO_TRY(dalloc(ctx, &order0.p, ntiles));
		VO_TRY(dalloc(ctx, &order1.p, ntiles));
	}
	for (auto &st : ctx->s_p)
		if (!st && cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return PIPE_NA; }
	for (auto &st : ctx->s_hi) {       // the second half (pass 2, scan, compaction) of a band outranks pass 1 of the bands behind it
		if (st) continue;
		int least = 0, greatest = 0;
		cudaDeviceGetStreamPriorityRange(&least, &greatest);
		if (cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, greatest) != cudaSuccess) { cudaGetLastError(); return PIPE_NA; }
	}
	if (!ctx->s_mid) {                 // thresholds and tile order: between the tile launches and the second halves
		int least = 0, greatest = 0;
		cudaDeviceGetStreamPriorityRange(&least, &greatest);
		if (cudaStreamCreateWithPriority(&ctx->s_mid, cudaStreamNonBlocking, (least + greatest) / 2) != cudaSuccess) { cudaGetLastError(); return PIPE_NA; }
	}
	// the tile lists and cursors of every band's launch set (with pipe_mid: the thresholds of band b + 2 may run beside
	// the tile launches of band b, so k_thresh cannot be the one to zero a shared bank)
	Tmp<unsigned long long> banks(ctx);
	const bool mid = ctx->pipe_mid;
	if (mid) VO_TRY(dalloc(ctx, &banks.p, 16ull * nb));
	// redo lists of pass 1, one per band (slot ids that outgrew the fast paths; their counts sit behind the band totals in `gb`)
	const unsigned int rcap1 = (unsigned int)std::min<unsigned long long>((unsigned long long)nx * BH * (J + 1), 1ull << 20);
	Tmp<unsigned long long> redo1(ctx);
	VO_TRY(dalloc(ctx, &redo1.p, (unsigned long long)rcap1 * nb));

	// second half: whole-grid staging, per-band redo lists and scan scratch, the result volume on the device
	const uint64_t hs_cap = std::min<uint64_t>(std::max<uint64_t>(ctx->out_hint, 2 * nspans + (1u << 16)), (1ull << 32) - 1);
	const unsigned long long dcap = hs_cap;
	StageBuf sb(ctx);
	VO_TRY(sb.alloc(ncols, 65536ull + ncols / 8));
	const unsigned int rcap2 = (unsigned int)std::min<unsigned long long>((unsigned long long)nx * BH2, 1ull << 22);
	Tmp<unsigned long long> redo2(ctx), sums(ctx), gb(ctx);
	VO_TRY(dalloc(ctx, &redo2.p, (unsigned long long)rcap2 * nb));
	const unsigned int nt_max = blocks_for((unsigned long long)nx * BH2, SCAN_TILE);
	VO_TRY(dalloc(ctx, &sums.p, ((unsigned long long)nt_max + 1) * nb));
	VO_TRY(dalloc(ctx, &gb.p, 3ull * nb + 1));                // [nb + 1] running totals, [nb] redo counts of pass 2, [nb] of pass 1
	vo_dvol *dout = nullptr;
	VO_TRY(new_dvol(ctx, nx, ny, &dout));
	struct OutGuard { vo_ctx *c; vo_dvol *p; ~OutGuard() { free_dvol(c, p); } } out_guard{ctx, dout};
	VO_TRY(dalloc(ctx, &dout->spans, dcap));
	if (!ctx->s_ctl && cudaStreamCreateWithFlags(&ctx->s_ctl, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return PIPE_NA; }
	if (!ctx->s_out2) {
		int least = 0, greatest = 0;
		cudaDeviceGetStreamPriorityRange(&least, &greatest);
		if (cudaStreamCreateWithPriority(&ctx->s_out2, cudaStreamNonBlocking, greatest) != cudaSuccess) { cudaGetLastError(); return PIPE_NA; }
	}
	// Downloads: the offsets of a band have a size the host knows (copy engine, enqueued up front behind the band's
	// event); the spans do not - the SMs copy them (k_copy_out reads the range on the device), or the host waits for
	// every band's total and enqueues a copy of that size (copy_out = 0; a round trip per band, ~0.2 ms per call)
	const bool sm_out = ctx->copy_out_ctas > 0;

	// result buffers (pinned): offsets are exact, the span buffer is sized from the last result
	const unsigned long long ckeep0 = (unsigned long long)keep0 * nx, nkeep = (unsigned long long)(keep1 - keep0) * nx;
	uint32_t *ho = (uint32_t *)host_block((nkeep + 1) * sizeof(uint32_t));
	double *hs = (double *)host_block(hs_cap * sizeof(double2));
	unsigned long long *h_tot = (unsigned long long *)host_block((size_t)nb * sizeof(unsigned long long));
	auto drop_host = [&]() { vo_free(ho); vo_free(hs); };
	if (!ho || !hs || !h_tot) { drop_host(); vo_free(h_tot); return fail(ctx, VO_ERR_NOMEM, "pinned host allocation failed"); }

	// copy boundaries in elements (vo_ctx::copy_align): only where source and destination are aligned alike
	const unsigned long long al_bytes = ctx->copy_align >= 16 && (ctx->copy_align & (ctx->copy_align - 1)) == 0 ? (unsigned long long)ctx->copy_align : 0ull;
	auto aligned_ptr = [&](const void *p) { return al_bytes && (reinterpret_cast<uintptr_t>(p) & (al_bytes - 1)) == 0; };
	const unsigned long long al_off = aligned_ptr(off) && aligned_ptr(in->off) ? al_bytes / sizeof(uint32_t) : 0ull;
	const unsigned long long al_sp = aligned_ptr(spans) && aligned_ptr(in->spans) && al_bytes <= 16 * sizeof(double2) ? al_bytes / sizeof(double2) : 0ull;
	const unsigned long long al_off_out = aligned_ptr(ho) && aligned_ptr(dout->off) && ckeep0 % (al_bytes ? al_bytes / sizeof(uint32_t) : 1) == 0 ? al_bytes / sizeof(uint32_t) : 0ull;
	const unsigned long long al_sp_out = aligned_ptr(hs) && aligned_ptr(dout->spans) ? al_bytes / sizeof(double2) : 0ull;
	// two copies of one direction: one batch where the runtime has it (vo_ctx::copy_batch), else one after the other
	auto copy_pair = [&](void *d0, const void *s0, size_t n0, void *d1, const void *s1, size_t n1, cudaMemcpyKind kind, cudaStream_t st) {
		if (ctx->copy_batch && n0 && n1) {
			void *dsts[2] = {d0, d1}, *srcs[2] = {const_cast<void *>(s0), const_cast<void *>(s1)};
			size_t sizes[2] = {n0, n1}, idx[1] = {0}, fail_idx = 0;
			cudaMemcpyAttributes attr{};
			attr.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
			if (cudaMemcpyBatchAsync(dsts, srcs, sizes, 2, &attr, idx, 1, &fail_idx, st) == cudaSuccess) return;
			cudaGetLastError();
			ctx->copy_batch = false;                            // (a driver without it: plain copies from now on)
		}
		if (n0) cudaMemcpyAsync(d0, s0, n0, kind, st);
		if (n1) cudaMemcpyAsync(d1, s1, n1, kind, st);
	};
	// offsets of the lists [c0, c0 + n) of the result (+ the closing one for the last band) to the host
	auto download_offsets = [&](unsigned long long c0, unsigned long long n, bool last) {
		unsigned long long a0 = c0, a1 = c0 + n + (last ? 1 : 0);
		if (al_off_out) { a0 &= ~(al_off_out - 1); if (!last) a1 = (a1 + al_off_out - 1) & ~(al_off_out - 1); }
		cudaMemcpyAsync(ho + (a0 - ckeep0), dout->off + a0, (a1 - a0) * sizeof(uint32_t), cudaMemcpyDeviceToHost, pr.s_out);
	};
	// stage 0: all uploads are enqueued up front, one event per band
	const auto host_t0 = std::chrono::steady_clock::now();
	cudaEvent_t ev_t0 = nullptr;
	if (std::getenv("VO_TRACE")) { cudaEventCreate(&ev_t0); cudaEventRecord(ev_t0, pr.s_in); }
	std::vector<cudaEvent_t> ev_in(nb), ev_done(nb);
	for (int b = 0; b < nb; ++b) {
		// rows [ys[b] + 1, ys[b + 1] + 1): one row ahead of the band, so that pass 1 of band b (whose thresholds read the
		// row below) can start as soon as ITS upload is done instead of waiting for the next one
		const int y0 = b == 0 ? 0 : ys[b] + 1, y1 = std::min(ny, ys[b + 1] + 1);
		const unsigned long long c0 = (unsigned long long)y0 * nx, c1 = (unsigned long long)y1 * nx;
		{
			unsigned long long a0 = c0, a1 = c1 + 1;            // offsets [a0, a1)
			if (al_off) { a0 = c0 & ~(al_off - 1); a1 = std::min<unsigned long long>(ncols + 1, (c1 + 1 + al_off - 1) & ~(al_off - 1)); }
			unsigned long long i0 = off[c0], i1 = off[c1];      // spans [i0, i1)
			if (i1 > i0 && al_sp) { i0 &= ~(al_sp - 1); i1 = std::min<unsigned long long>((unsigned long long)obase + nspans, (i1 + al_sp - 1) & ~(al_sp - 1)); }
			copy_pair(in->off + a0, off + a0, (a1 - a0) * sizeof(uint32_t), d_sp + i0, spans + 2 * (size_t)i0, (size_t)(i1 > i0 ? i1 - i0 : 0) * sizeof(double2),
			          cudaMemcpyHostToDevice, pr.s_in);
		}
		ev_in[b] = pr.event();
		ev_done[b] = pr.event();
		cudaEventRecord(ev_in[b], pr.s_in);
		KT_MARK_STREAM(1000 + b, pr.s_in);
	}
	if (cudaGetLastError() != cudaSuccess) { drop_host(); return PIPE_NA; }
	cudaEvent_t ev_up_done = nullptr;
	if (std::getenv("VO_TRACE")) { cudaEventCreate(&ev_up_done); cudaEventRecord(ev_up_done, pr.s_in); }

	cudaStream_t sh[2] = {ctx->s_hi[0], ctx->s_hi[1]};
	// VO_TRACE=1: device-side timeline of the call on stderr (development aid, scripts/e2e_bands.py)
	static const bool trace = std::getenv("VO_TRACE") != nullptr;
	std::vector<std::pair<std::string, cudaEvent_t>> marks;
	static const bool trace_few = trace && std::getenv("VO_TRACE")[0] == '2';   // VO_TRACE=2: band 0 and the downloads only (perturbs less)
	auto mark = [&](const char *name, int b, cudaStream_t st) {
		if (!trace || (trace_few && b != 0 && std::strncmp(name, "download", 8) != 0)) return;
		cudaEvent_t ev;
		cudaEventCreate(&ev);
		cudaEventRecord(ev, st);
		const double host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count();
		char buf[96];
		std::snprintf(buf, sizeof buf, "%s %d (enqueued at host %.3f ms)", name, b, host_ms);
		marks.emplace_back(buf, ev);
	};
	if (trace && ev_t0) marks.emplace_back("start 0", ev_t0);
	cudaMemsetAsync(ctx->d_ctr, 0, NCTR * sizeof(unsigned long long), ctx->stream);
	cudaMemsetAsync(m->tilemask, 0, 2ull * ny * tiles_x * sizeof(unsigned long long), ctx->stream);
	if (ordered) cudaMemsetAsync(est.p, 0, (ntiles + 2ull * P1_NBUCKET * nb) * sizeof(unsigned int), ctx->stream);
	cudaMemsetAsync(gb.p, 0, (3ull * nb + 1) * sizeof(unsigned long long), ctx->stream);
	if (mid) cudaMemsetAsync(banks.p, 0, 16ull * nb * sizeof(unsigned long long), ctx->stream);
	cudaEventRecord(ctx->ev[0], ctx->stream);
	{
		cudaEvent_t ev_init = pr.event();
		cudaEventRecord(ev_init, ctx->stream);
		for (auto st : ctx->s_p) cudaStreamWaitEvent(st, ev_init, 0);
		for (auto st : ctx->s_hi) cudaStreamWaitEvent(st, ev_init, 0);
		cudaStreamWaitEvent(ctx->s_mid, ev_init, 0);
	}
	std::vector<cudaEvent_t> ev_p1(nb), ev_r1(nb), ev_scan(nb);

	// launch parameters shared by all bands
	ThreshArgs ta;
	ta.nx = nx; ta.ny = ny; ta.J = J; ta.off = in->off; ta.spans = d_sp;
	ta.Dmono = tt.Dmono; ta.Emono = tt.Emono; ta.G = tt.G; ta.reach = dt.reach; ta.thr = d_thr;
	ta.clip_lo = -std::numeric_limits<double>::infinity(); ta.clip_hi = std::numeric_limits<double>::infinity();
	ta.bad = reinterpret_cast<unsigned int *>(ctx->d_ctr + 11); ta.nspans = (uint32_t)(obase + nspans); ta.nspans_lo = obase;
	Pass1TileArgs g;
	g.nx = nx; g.ny = ny;
	g.off = in->off; g.spans = d_sp; g.thr = d_thr; g.Ht = tt.Ht; g.Ef = tt.Ef; g.jmax = tt.jmax;
	g.mid = m->slots; g.flags = m->flags; g.tilemask = m->tilemask; g.pool = m->pool; g.cursor = ctx->d_ctr; g.pool_cap = m->pool_cap;
	g.bad = ta.bad;
	Pass1Args a1;
	a1.nx = nx; a1.ny = ny; a1.J = J; a1.off = in->off; a1.spans = d_sp; a1.H = dt.H; a1.reach = dt.reach;
	a1.mid = m->slots; a1.pool = m->pool; a1.cursor = ctx->d_ctr; a1.pool_cap = m->pool_cap;
	auto redo_of = [&](int b) { return Redo{redo1.p + (size_t)b * rcap1, reinterpret_cast<unsigned int *>(gb.p + 2 * nb + 1 + b), rcap1}; };

	// Pass 1 of band b runs on one of two streams, with that stream's own tile lists and cursors: the launch set of
	// band b+1 starts while the last heavy tiles of band b are still being worked on (every launch of the tile kernel
	// ends with such a tail), its CTAs taking over the SMs as those of band b retire.
	// (normally idle launches left out: see vo_ctx::pipe_lean)
	const bool no_lists = ctx->pipe_lean && !ctx->pipe_lists, no_redo = ctx->pipe_lean && !ctx->pipe_redo;
	const bool p2_regonly = no_lists && (ctx->p2_mode == 0 || (ctx->p2_mode < 0 && ctx->p2_list_calls == 0));   // (see pass2())
	if (!p2_regonly && ctx->p2_list_calls > 0) --ctx->p2_list_calls;
	auto pass1_band = [&](int b) {
		const int y0 = ys[b], y1 = ys[b + 1], w = b & 1;
		cudaStream_t sp = ctx->s_p[w];
		cudaStream_t st = mid ? ctx->s_mid : sp;             // thresholds and order
		unsigned long long *bank = mid ? banks.p + 16ull * b : w ? ctx->d_ctr + NCTR + 8 : ctx->d_ctr;
		cudaStreamWaitEvent(st, ev_in[b], 0);                // (the upload of a band includes the first row of the next: thresholds look one row down)
		mark("pass1 begin", b, st);
		ta.zero_bank = mid ? nullptr : bank;                 // the lists and cursors of this launch set, zeroed by k_thresh itself
		ta.c_begin = (unsigned long long)y0 * nx; ta.c_end = (unsigned long long)y1 * nx;
		const unsigned int t0 = (unsigned int)plan.tiles_xw * (unsigned int)y0, nt = (unsigned int)plan.tiles_xw * (unsigned int)(y1 - y0);
		unsigned int *ord = nullptr;
		if (ordered) {
			ta.est = est.p + 2ull * P1_NBUCKET * nb; ta.tiles_xw = plan.tiles_xw;
			ord = (w ? order1.p : order0.p) + t0;
		}
		if (ctx->pipe_dry) { ev_p1[b] = pr.event(); cudaEventRecord(ev_p1[b], sp); return; }
		launch_thresh(ta, k_in, st);
		ctx->launches++;
		mark("  thresh end", b, st);
		if (ordered) TilePlan::order_tiles(ctx, ta.est, est.p + 2ull * P1_NBUCKET * b, ord, t0, nt, 0u, 0u, st, ctx->pipe_order_one);
		mark("  order end", b, st);
		if (mid) { cudaEvent_t e = pr.event(); cudaEventRecord(e, st); cudaStreamWaitEvent(sp, e, 0); }
		g.redo = redo_of(b);
		// (the two streams' launch sets may each take a share of the SMs, so that they run side by side instead of the
		// second waiting for CTAs of the first to retire)
		const int usable = std::max(ctx->band_split, plan.sms - std::max(0, ctx->band_free));
		const int reserve = (b == 0 && ctx->pipe_first_full) ? 0 : plan.sms - usable / std::max(1, ctx->band_split);
		(b == 0 ? plan0 : plan).launch(ctx, g, t0, nt, w ? big_tiles1.p : big_tiles.p, w ? multi_tiles1.p : multi_tiles.p, sp, 0u, 0u, reserve, bank, ord, false, w,
		                               no_lists, reinterpret_cast<unsigned int *>(ctx->d_ctr + 14));
		ev_p1[b] = pr.event();
		cudaEventRecord(ev_p1[b], sp);
		mark("pass1 end", b, sp);
	};
	// Whatever outgrew the fast paths of pass 1 in band b is redone at the head of that band's second half (normally an
	// idle launch: no stream of its own, no extra hop between the streams). ev_r1[b] = pass 1 of the bands 0 .. b is
	// complete, redone slots included.
	auto pass1_redo = [&](int b, cudaStream_t st) {
		cudaStreamWaitEvent(st, ev_p1[b], 0);
		if (b > 0) cudaStreamWaitEvent(st, ev_r1[b - 1], 0);
		a1.redo = redo_of(b);
		a1.wk = Work{a1.redo.list, 0ull, a1.redo.count, a1.redo.cap, reinterpret_cast<unsigned int *>(ctx->d_ctr + 4)};
		if (!no_redo) {
			k_pass1<CAP_BIG><<<REDO_GRID, 128, 0, st>>>(a1);
			ctx->launches++;
		}
		ev_r1[b] = pr.event();
		cudaEventRecord(ev_r1[b], st);
	};

	// Second half of a band: pass 2 -> prefix sum -> compaction, enqueued WITHOUT a host round trip. The bands share
	// whole-grid staging buffers (each uses its slice) and write one result volume on the device; the running interval
	// total lives on the device (gb[b] = intervals of the bands before b), so the offsets come out global. The host only
	// learns the totals to size the downloads, on a control stream of its own: the compute streams never wait for it.
	int rc = VO_OK;
	// declared last, hence destroyed first: on every way out the side streams are drained BEFORE the temporaries above
	// go back to the pool of the context stream (stream-ordered frees only order against that stream)
	struct Join { vo_ctx *c; ~Join() { for (auto st : c->s_hi) cudaStreamSynchronize(st); cudaStreamSynchronize(c->s_ctl); for (auto st : c->s_p) cudaStreamSynchronize(st); if (c->s_mid) cudaStreamSynchronize(c->s_mid); if (c->s_out2) cudaStreamSynchronize(c->s_out2); } } join{ctx};

	std::vector<cudaEvent_t> ev_tot(nb);
	// single-pass scan + compaction per band: tile state reserved up front (no allocation between the launches)
	const bool fused = ctx->fused_scan;
	std::vector<uint32_t> scan_epoch(nb, 0u);
	std::vector<unsigned long long> scan_tbase(nb, 0ull);
	// (row window: a band's second half covers its rows inside [keep0, keep1); bands with none only take part in pass 1.
	// prev_act = the last band before b with rows to produce: the running total and the scan chain skip the others.)
	std::vector<int> prev_act(nb, -1);
	std::vector<char> act(nb, 0);
	int last_act = -1;
	for (int b = 0, pa = -1; b < nb; ++b) {
		prev_act[b] = pa;
		act[b] = std::max(ys2[b], keep0) < std::min(ys2[b + 1], keep1);
		if (act[b]) { pa = b; last_act = b; }
	}
	auto rows2 = [&](int b) { return std::max(0, std::min(ys2[b + 1], keep1) - std::max(ys2[b], keep0)); };
	if (fused) {
		unsigned int most = 1;
		for (int b = 0; b < nb; ++b) most = std::max(most, blocks_for((unsigned long long)nx * rows2(b), SCAN_TILE));
		VO_TRY(scan_reserve(ctx, most));
		for (int b = 0; b < nb; ++b) {
			if (!act[b]) continue;
			const unsigned int nt = blocks_for((unsigned long long)nx * rows2(b), SCAN_TILE);
			VO_TRY(scan_prepare(ctx, nt, &scan_epoch[b], &scan_tbase[b]));
		}
	}
	auto second_half = [&](int b) {
		const int y0 = std::max(ys2[b], keep0), y1 = std::min(ys2[b + 1], keep1);
		if (!act[b]) {                                      // pass 1's redo launch still has to run (and keep the ev_r1 chain whole)
			pass1_redo(b, sh[b & 1]);
			return;
		}
		const int pb = prev_act[b];
		const unsigned long long c0 = (unsigned long long)y0 * nx, nlists = (unsigned long long)nx * (y1 - y0);
		const unsigned int nt = blocks_for(nlists, SCAN_TILE);
		unsigned long long *sums_b = sums.p + (size_t)b * (nt_max + 1);
		Stage st = sb.st;
		st.cnt += c0; st.inl += c0 * STAGE_INLINE;
		const Redo rd{redo2.p + (size_t)b * rcap2, reinterpret_cast<unsigned int *>(gb.p + nb + 1 + b), rcap2};
		cudaStream_t sm = sh[b & 1];                        // consecutive bands overlap; only the running total is a chain
		pass1_redo(b, sm);                                   // rows [ys2[b], ys2[b+1]) read mid rows below ys[b+1]: pass 1 of the bands 0 .. b
		mark("pass2 begin", b, sm);
		Pass2Args a2;
		a2.nx = nx; a2.ny = ny; a2.J = J; a2.y0 = y0; a2.y1 = y1;
		a2.mid = m->slots; a2.flags = m->flags; a2.tilemask = m->tilemask; a2.pool = m->pool; a2.pool_cap = m->pool_cap; a2.st = st; a2.redo = rd;
		a2.wk = Work{nullptr, nlists, nullptr, 0u, nullptr};
		if (ctx->pipe_dry) {
			cudaEventRecord(ev_done[b], sm);
			cudaStreamWaitEvent(ctx->s_ctl, ev_done[b], 0);
			cudaMemcpyAsync(h_tot + b, gb.p + b + 1, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->s_ctl);
			ev_tot[b] = pr.event();
			cudaEventRecord(ev_tot[b], ctx->s_ctl);
			return;
		}
		if (J <= 32 && no_lists && p2_regonly) k_pass2_rows<2, false, P2_REGONLY><<<(unsigned int)((nx + P2_TX - 1) / P2_TX) * (unsigned int)(y1 - y0), P2_TX, 0, sm>>>(a2);
		else if (J <= 32 && no_lists) k_pass2_rows<CAP_FAST, false, P2_SHALLOW><<<(unsigned int)((nx + P2_TX - 1) / P2_TX) * (unsigned int)(y1 - y0), P2_TX, 0, sm>>>(a2);
		else if (J <= 32) k_pass2_rows<CAP_FAST, false><<<(unsigned int)((nx + P2_TX - 1) / P2_TX) * (unsigned int)(y1 - y0), P2_TX, 0, sm>>>(a2);
		else k_pass2_rows<CAP_FAST><<<(unsigned int)((nx + P2_TX - 1) / P2_TX) * (unsigned int)(y1 - y0), P2_TX, 0, sm>>>(a2);
		a2.wk = Work{rd.list, 0ull, rd.count, rd.cap, reinterpret_cast<unsigned int *>(ctx->d_ctr + 9)};
		if (!no_redo) k_pass2<CAP_BIG><<<REDO_GRID, 128, 0, sm>>>(a2);
		if (fused) {
			// ONE kernel: offsets (from the running total gb[b] of the bands before) and spans of the band; the bands'
			// kernels follow one another (the running total is a chain anyway, and each takes a few microseconds)
			if (pb >= 0) cudaStreamWaitEvent(sm, ev_scan[pb], 0);
			k_scan_compact<SCAN_ITEMS><<<nt, SCAN_THREADS, 0, sm>>>(st, nlists, dout->off + c0, dout->spans, dcap, ctx->scan_state,
			                                            ctx->scan_state + ctx->scan_cap, scan_tbase[b], scan_epoch[b], gb.p + b + 1, gb.p + pb + 1);
			ev_scan[b] = pr.event();
			cudaEventRecord(ev_scan[b], sm);
			ctx->launches += 3;
		} else {
			k_scan_reduce<<<nt, SCAN_THREADS, 0, sm>>>(st.cnt, nlists, sums_b);
			if (pb >= 0) cudaStreamWaitEvent(sm, ev_scan[pb], 0);
			k_scan_tiles<<<1, 1024, 0, sm>>>(sums_b, nt, gb.p + pb + 1, gb.p + b + 1);
			ev_scan[b] = pr.event();
			cudaEventRecord(ev_scan[b], sm);
			k_scan_apply<<<nt, SCAN_THREADS, 0, sm>>>(st.cnt, nlists, sums_b, dout->off + c0);
			k_compact<<<blocks_for(nlists, 256), 256, 0, sm>>>(st, nlists, dout->off + c0, dout->spans, dcap);
			ctx->launches += 6;
		}
		cudaEventRecord(ev_done[b], sm);
		mark("pass2 end", b, sm);
		if (sm_out) {
			cudaStreamWaitEvent(pr.s_out, ev_done[b], 0);
			download_offsets(c0, nlists, b == last_act);
			cudaStreamWaitEvent(ctx->s_out2, ev_done[b], 0);
			mark("download begin", b, ctx->s_out2);
			k_copy_out<<<ctx->copy_out_ctas, COPY_OUT_THREADS, 0, ctx->s_out2>>>(dout->spans, reinterpret_cast<double2 *>(hs), gb.p + pb + 1, gb.p + b + 1, dcap);
			ctx->launches++;
			mark("download end", b, ctx->s_out2);
			return;
		}
		cudaStreamWaitEvent(ctx->s_ctl, ev_done[b], 0);
		cudaMemcpyAsync(h_tot + b, gb.p + b + 1, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->s_ctl);
		ev_tot[b] = pr.event();
		cudaEventRecord(ev_tot[b], ctx->s_ctl);
	};

	// Everything is enqueued up front: pass 1 of every band (it only waits for its upload), then the second halves
	// (pass 2 of band b-1 needs pass 1 of band b). After that the host follows the bands and starts their downloads.
	if (ctx->pipe_interleave) {
		// (second half of band b right behind pass 1 of band b + 1: the host reaches it ~0.2 ms earlier)
		pass1_band(0);
		for (int b = 1; b < nb; ++b) { pass1_band(b); second_half(b - 1); }
		second_half(nb - 1);
	} else {
		for (int b = 0; b < nb; ++b) pass1_band(b);
		for (int b = 0; b < nb; ++b) second_half(b);
	}
	uint64_t base = 0;          // intervals of the bands downloaded so far
	// The offsets of a band have a size the host knows: their copy is enqueued AHEAD, behind the band's event, and runs
	// while the host is still waiting for the band's total - the round trip (total -> host -> enqueue) hides behind it.
	// (Not all of them up front: a stream runs in enqueue order, and the spans of band b must not queue behind the
	// offsets of the bands after it.)
	auto offsets_ahead = [&](int b, bool now = false) {
		if (!ctx->pipe_ahead && !now) return;
		while (!now && b < nb && !act[b]) ++b;
		if (b >= nb) return;
		const int y0 = std::max(ys2[b], keep0), y1 = std::min(ys2[b + 1], keep1);
		const unsigned long long c0 = (unsigned long long)y0 * nx, nlists = (unsigned long long)nx * (y1 - y0);
		cudaStreamWaitEvent(pr.s_out, ev_done[b], 0);
		mark("download begin", b, pr.s_out);
		KT_MARK_STREAM(2000 + b, pr.s_out);
		download_offsets(c0, nlists, b == last_act);
	};
	const bool host_driven = !(sm_out && !ctx->pipe_dry);
	if (host_driven) offsets_ahead(0);
	for (int b = 0; b < nb && rc == VO_OK && host_driven; ++b) {
		if (!act[b]) continue;
		if (cudaEventSynchronize(ev_tot[b]) != cudaSuccess) { rc = PIPE_NA; break; }
		uint64_t tot = h_tot[b];
		if (ctx->pipe_dry) tot = (uint64_t)ctx->pipe_dry * (uint64_t)(b + 1) / (uint64_t)nb;
		if (tot > dcap) {                                   // result buffers too small (first call, or a much larger result): plain path
			ctx->out_hint = std::max<uint64_t>(ctx->out_hint, 2 * tot);
			rc = PIPE_NA;
			break;
		}
		{
			unsigned long long j0 = base, j1 = tot;
			if (j1 > j0 && al_sp_out) { j0 &= ~(al_sp_out - 1); j1 = std::min<unsigned long long>(dcap, (j1 + al_sp_out - 1) & ~(al_sp_out - 1)); }
			const size_t nsp = (size_t)(j1 > j0 ? j1 - j0 : 0) * sizeof(double2);
			if (!ctx->pipe_ahead) {                             // offsets and spans of the band in one go
				const int y0 = std::max(ys2[b], keep0), y1 = std::min(ys2[b + 1], keep1);
				const unsigned long long c0 = (unsigned long long)y0 * nx, n = (unsigned long long)nx * (y1 - y0);
				unsigned long long a0 = c0, a1 = c0 + n + (b == last_act ? 1 : 0);
				if (al_off_out) { a0 &= ~(al_off_out - 1); if (b != last_act) a1 = (a1 + al_off_out - 1) & ~(al_off_out - 1); }
				cudaStreamWaitEvent(pr.s_out, ev_done[b], 0);
				mark("download begin", b, pr.s_out);
				KT_MARK_STREAM(2000 + b, pr.s_out);
				copy_pair(ho + (a0 - ckeep0), dout->off + a0, (a1 - a0) * sizeof(uint32_t), hs + 2 * j0, dout->spans + j0, nsp, cudaMemcpyDeviceToHost, pr.s_out);
			} else if (nsp) cudaMemcpyAsync(hs + 2 * j0, dout->spans + j0, nsp, cudaMemcpyDeviceToHost, pr.s_out);
		}
		mark("download end", b, pr.s_out);
		KT_MARK_STREAM(3000 + b, pr.s_out);
		base = tot;
		offsets_ahead(b + 1);
	}
	{
		for (auto st : ctx->s_p) { cudaEvent_t e = pr.event(); cudaEventRecord(e, st); cudaStreamWaitEvent(ctx->stream, e, 0); }
		for (auto st : ctx->s_hi) { cudaEvent_t e = pr.event(); cudaEventRecord(e, st); cudaStreamWaitEvent(ctx->stream, e, 0); }
	}
	cudaEventRecord(ctx->ev[2], ctx->stream);
	unsigned long long h[NREAD];
	std::vector<unsigned long long> hgb(3 * (size_t)nb + 1, 0ull);
	if (rc == VO_OK && cudaMemcpyAsync(hgb.data(), gb.p, hgb.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) rc = PIPE_NA;
	if (rc == VO_OK) rc = read_counters(ctx, h);
	cudaStreamSynchronize(pr.s_out);
	cudaStreamSynchronize(ctx->s_out2);
	cudaStreamSynchronize(pr.s_in);
	if (rc == VO_OK && sm_out && !ctx->pipe_dry) {              // the total the device arrived at (k_copy_out stopped at the buffer's end)
		base = last_act >= 0 ? hgb[(size_t)last_act + 1] : 0;
		if (base > dcap) { ctx->out_hint = std::max<uint64_t>(ctx->out_hint, 2 * base); rc = PIPE_NA; }
	}
	vo_free(h_tot);
	if (rc == VO_OK) {
		bool redo = h[9] != 0 || h[1] > sb.st.pool_cap;         // rare: the plain path regrows / reports
		for (int b = 0; b < nb; ++b) redo = redo || (unsigned int)hgb[nb + 1 + b] > rcap2 || (unsigned int)hgb[2 * nb + 1 + b] > rcap1;
		// launches that were left out and turned out to be needed: the plain path does this call, the next ones make them
		if (p2_regonly) {                                        // (as pass2(): columns with a third interval)
			unsigned long long redo2_count = 0;
			for (int b = 0; b < nb; ++b) redo2_count += (unsigned int)hgb[nb + 1 + b];
			if (redo2_count) ctx->p2_list_calls = 64;
		}
		unsigned long long redo1_count = 0;
		for (int b = 0; b < nb; ++b) redo1_count += (unsigned int)hgb[2 * nb + 1 + b];
		note_pass1_redo(ctx, redo1_count);                           // (slots handed over by the 20-warp first launch: see there)
		if (no_redo)
			for (int b = 0; b < nb; ++b)
				if ((unsigned int)hgb[nb + 1 + b] || (unsigned int)hgb[2 * nb + 1 + b]) { ctx->pipe_redo = true; redo = true; }
		if (no_lists && h[14]) { ctx->pipe_lists = true; redo = true; }
		if (redo) rc = PIPE_NA;
	}
	if (trace && !marks.empty()) {
		float t = 0;
		if (ev_up_done) { cudaEventElapsedTime(&t, marks[0].second, ev_up_done); std::fprintf(stderr, "[vo trace] %-18s %8.3f ms\n", "uploads end", t); cudaEventDestroy(ev_up_done); }
		for (auto &mk : marks) {
			if (cudaEventElapsedTime(&t, marks[0].second, mk.second) == cudaSuccess) std::fprintf(stderr, "[vo trace] %8.3f ms  %s\n", t, mk.first.c_str());
		}
		for (auto &mk : marks) cudaEventDestroy(mk.second);
		cudaGetLastError();
	}
	if (rc == VO_OK && cudaGetLastError() != cudaSuccess) rc = PIPE_NA;
	if (rc == VO_OK && h[11]) { drop_host(); return fail(ctx, VO_ERR_ARG, "offsets must be non-decreasing"); }
	if (rc == VO_OK && (h[0] > m->pool_cap || h[4])) {
		ctx->pool_hint = std::max<unsigned long long>(ctx->pool_hint, h[0] + h[0] / 4);
		rc = PIPE_NA;                                           // let the plain path deal with it (it regrows / reports)
	}
	if (rc != VO_OK) { drop_host(); return rc; }
	ctx->pool_hint = next_hint(ctx->pool_hint, h[0]);
	ctx->out_hint = base + base / 8 + (1u << 16);
	float tms = 0;
	cudaEventElapsedTime(&tms, ctx->ev[0], ctx->ev[2]);
	if (pt) { pt->ms1 = tms; pt->ms2 = 0; }                     // the passes interleave: one figure for both
	*out_off = ho; *out_spans = hs;
	if (out_nspans) *out_nspans = base;
	return VO_OK;
}


// ---------------------------------------------------------------------------------------------------
// Overlapped y-slab dilation (multi-GPU, voroffset_b200/slab.py). A rank owns ny rows and needs floor(R) input
// rows of each neighbour. vo_slab_begin lays out the EXTENDED volume [prev halo | own rows | next halo] before the
// halos exist - their spans get reserved regions of agreed capacity, the previous halo right-aligned in its region
// so that the CSR stays contiguous - and enqueues pass 1 of the own rows that do not depend on a halo. The caller
// runs the NCCL exchange meanwhile; vo_slab_finish drops the halos in, runs pass 1 on the 2 + 2 floor(R) remaining
// rows and pass 2 on the own rows. Same kernels and tables as the plain path: the rows are bit-identical.
// ---------------------------------------------------------------------------------------------------
} // namespace (reopened below)

struct vo_slab {
	vo_ctx *ctx = nullptr;
	vo_dvol *ext = nullptr;
	vo_dmid *mid = nullptr;
	int nx = 0, ny = 0, J = 0, jp = 0, jn = 0;      // own rows; halo rows before / after (0 or J)
	uint64_t cap_prev = 0, cap_next = 0, n_own = 0;
	double R = 0;
	double clip_lo = -std::numeric_limits<double>::infinity(), clip_hi = std::numeric_limits<double>::infinity();   // erosion: the result is only read inside
	TilePlan plan;
	uint4 *thr = nullptr;
	unsigned int *big_tiles = nullptr, *multi_tiles = nullptr;
	unsigned int *big_tiles_b = nullptr, *multi_tiles_b = nullptr;   // ... of the boundary rows' launch set
	unsigned int *est = nullptr, *order = nullptr;  // tile order (expensive first): [2 sets x 2 P1_NBUCKET counters | cost per tile]; permutations of the two sets
	unsigned long long ntiles = 0;
	double k_in = 0;
	cudaEvent_t ev_setup = nullptr, ev_side = nullptr;
	unsigned long long *redo_list = nullptr;
	unsigned int redo_cap = 0;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
	bool overlapped = false;
	bool dual = false;                              // erosion in dual form (erode_dual): mirrored intervals through pass 1, k_pass2_rows_dual
	double dual_zmin = 0, dual_zmax = 0;            // ... the z range of the erosion
};

namespace {

void free_slab(vo_slab *S)
{
	if (!S) return;
	vo_ctx *ctx = S->ctx;
	free_dvol(ctx, S->ext);
	if (S->mid) vo_dmid_free(ctx, S->mid);
	dfree(ctx, S->thr); dfree(ctx, S->big_tiles); dfree(ctx, S->multi_tiles); dfree(ctx, S->redo_list);
	dfree(ctx, S->big_tiles_b); dfree(ctx, S->multi_tiles_b); dfree(ctx, S->est); dfree(ctx, S->order);
	for (cudaEvent_t e : {S->ev0, S->ev1, S->ev2, S->ev_setup, S->ev_side}) if (e) cudaEventDestroy(e);
	delete S;
}

// pass 1 (thresholds + the three tile launches + redo) of rows [y0, y1) and [y0b, y1b) of the extended volume; no
// host sync
void slab_redo(vo_slab *S);

// `side` = true: the launch set of the boundary rows, on the side stream with its own tile lists and cursors, so
// that it can run next to the set of the interior rows (their redo launch follows once, after both).
void slab_pass1_rows(vo_slab *S, int y0, int y1, int y0b = 0, int y1b = 0, int reserve_sms = 0, bool side = false)
{
	if (y1 <= y0) { y0 = y0b; y1 = y1b; y0b = y1b = 0; }
	if (y1 <= y0) return;
	vo_ctx *ctx = S->ctx;
	cudaStream_t sm = side ? ctx->s_in : ctx->stream;
	unsigned long long *bank = side ? ctx->d_ctr + NCTR + 8 : ctx->d_ctr;
	TableCache *tc = static_cast<TableCache *>(ctx->table_cache);
	const int nx = S->nx, J = S->J;
	ThreshArgs ta;
	ta.zero_bank = bank;                                    // the lists and cursors of this launch set, zeroed by k_thresh (no memsets of their own)
	ta.nx = nx; ta.ny = S->ext->ny; ta.J = J; ta.off = S->ext->off; ta.spans = S->ext->spans;
	ta.Dmono = tc->tt.Dmono; ta.Emono = tc->tt.Emono; ta.G = tc->tt.G; ta.reach = tc->dt.reach; ta.thr = S->thr;
	ta.clip_lo = S->clip_lo; ta.clip_hi = S->clip_hi;
	if (S->dual) { ta.dual = 1; ta.dual_lo = S->dual_zmin - 1; ta.dual_hi = S->dual_zmax + 1; ta.dual_bad = reinterpret_cast<unsigned int *>(ctx->d_ctr + 14); }
	ta.c_begin = (unsigned long long)y0 * nx; ta.c_end = (unsigned long long)y1 * nx;
	if (S->est) { ta.est = S->est + 4 * P1_NBUCKET; ta.tiles_xw = S->plan.tiles_xw; }
	launch_thresh(ta, S->k_in, sm);
	ctx->launches++;
	if (y1b > y0b) {
		ta.zero_bank = nullptr;
		ta.c_begin = (unsigned long long)y0b * nx; ta.c_end = (unsigned long long)y1b * nx;
		launch_thresh(ta, S->k_in, sm);
		ctx->launches++;
	}
	const unsigned int *order = nullptr;
	if (S->est) {      // (each of the two launch sets of a step orders its tiles once: the counters were zeroed by vo_slab_begin)
		unsigned int *ord = S->order + (side ? S->ntiles : 0ull);
		TilePlan::order_tiles(ctx, ta.est, S->est + (side ? 2 * P1_NBUCKET : 0), ord, (unsigned int)S->plan.tiles_xw * (unsigned int)y0,
		                      (unsigned int)S->plan.tiles_xw * (unsigned int)(y1 - y0), (unsigned int)S->plan.tiles_xw * (unsigned int)y0b,
		                      (unsigned int)S->plan.tiles_xw * (unsigned int)std::max(0, y1b - y0b), sm, true);   // (one CTA where the set is small: the boundary rows)
		order = ord;
	}
	vo_dmid *m = S->mid;
	Redo rd{S->redo_list, reinterpret_cast<unsigned int *>(ctx->d_ctr + 2), S->redo_cap};
	Pass1TileArgs g;
	g.nx = nx; g.ny = S->ext->ny;
	g.off = S->ext->off; g.spans = S->ext->spans; g.thr = S->thr; g.Ht = tc->tt.Ht; g.Ef = tc->tt.Ef; g.jmax = tc->tt.jmax;
	g.mid = m->slots; g.flags = m->flags; g.tilemask = m->tilemask; g.pool = m->pool; g.cursor = ctx->d_ctr; g.pool_cap = m->pool_cap; g.redo = rd;
	const bool interior = reserve_sms != 0;              // the large launch: the one vo_last_profile reports
	reserve_sms = std::max(reserve_sms, 0);
	if (interior) cudaEventRecord(ctx->kev[0], sm);
	S->plan.launch(ctx, g, (unsigned int)S->plan.tiles_xw * (unsigned int)y0, (unsigned int)S->plan.tiles_xw * (unsigned int)(y1 - y0),
	               side ? S->big_tiles_b : S->big_tiles, side ? S->multi_tiles_b : S->multi_tiles, sm,
	               (unsigned int)S->plan.tiles_xw * (unsigned int)y0b,
	               (unsigned int)S->plan.tiles_xw * (unsigned int)std::max(0, y1b - y0b), reserve_sms, bank, order, S->dual);
	if (interior) { cudaEventRecord(ctx->kev[1], sm); ctx->kev_valid[0] = true; }
	if (side) return;
	slab_redo(S);
}

// lists that outgrew the fast capacity, and oversized tiles: one strided launch over the redo list (idempotent)
void slab_redo(vo_slab *S)
{
	if (S->dual) return;                                    // (nothing can land on the redo list; slab_finish checks that nothing did)
	vo_ctx *ctx = S->ctx;
	cudaStream_t sm = ctx->stream;
	TableCache *tc = static_cast<TableCache *>(ctx->table_cache);
	const int nx = S->nx, J = S->J;
	vo_dmid *m = S->mid;
	Redo rd{S->redo_list, reinterpret_cast<unsigned int *>(ctx->d_ctr + 2), S->redo_cap};
	Pass1Args a1;
	a1.nx = nx; a1.ny = S->ext->ny; a1.J = J; a1.off = S->ext->off; a1.spans = S->ext->spans; a1.H = tc->dt.H; a1.reach = tc->dt.reach;
	a1.mid = m->slots; a1.pool = m->pool; a1.cursor = ctx->d_ctr; a1.pool_cap = m->pool_cap; a1.redo = rd;
	a1.wk = Work{rd.list, 0ull, rd.count, rd.cap, reinterpret_cast<unsigned int *>(ctx->d_ctr + 4)};
	k_pass1<CAP_BIG><<<REDO_GRID, 128, 0, sm>>>(a1);
	ctx->launches++;
}

struct HaloOut { void *d_off; void *d_spans; uint64_t cap; };   // send buffer of one neighbour (d_off == NULL: none)

int slab_begin(vo_ctx *ctx, const vo_dvol *own, double R, int has_prev, int has_next, uint64_t cap_prev, uint64_t cap_next,
               HaloOut to_prev, HaloOut to_next, void *wait_stream, vo_slab **out,
               double clip_lo = -std::numeric_limits<double>::infinity(), double clip_hi = std::numeric_limits<double>::infinity(),
               bool dual = false, double dual_zmin = 0, double dual_zmax = 0)
{
	VO_TRY(check_radius(ctx, R));
	const int J = (int)std::floor(R), nx = own->nx, ny = own->ny;
	const int jp = has_prev ? J : 0, jn = has_next ? J : 0;
	if (!has_prev) cap_prev = 0;
	if (!has_next) cap_next = 0;
	const int ey = jp + ny + jn;
	VO_TRY(check_dims(ctx, nx, ey));
	const unsigned long long ncols = (unsigned long long)nx * ey, nown = (unsigned long long)nx * ny;
	const uint64_t total = cap_prev + own->nspans + cap_next;
	const double k_in = nown ? (double)own->nspans / (double)nown : 0.0;
	if (J < 1 || ny < 2 || total >= (1ull << 32) || !TilePlan::fits(J, k_in) || ctx->force_simple_pass1 ||
	    (double)ncols * (J + 1) * std::max(1.0, k_in) < (double)(2ull << 20))
		return VO_ERR_ARG;                                   // not a case for the overlapped path (the caller takes the plain one)
	TableCache *tc = nullptr;
	VO_TRY(get_tables(ctx, R, true, &tc));
	vo_slab *S = new (std::nothrow) vo_slab();
	if (!S) return fail(ctx, VO_ERR_NOMEM, "out of host memory");
	S->ctx = ctx; S->nx = nx; S->ny = ny; S->J = J; S->jp = jp; S->jn = jn; S->R = R;
	S->cap_prev = cap_prev; S->cap_next = cap_next; S->n_own = own->nspans; S->clip_lo = clip_lo; S->clip_hi = clip_hi;
	S->dual = dual; S->dual_zmin = dual_zmin; S->dual_zmax = dual_zmax;
	auto bail = [&](int rc) { free_slab(S); return rc; };
	int rc = new_dvol(ctx, nx, ey, &S->ext);
	if (rc == VO_OK) rc = dalloc(ctx, &S->ext->spans, total);
	if (rc) return bail(rc);
	S->ext->nspans = total;
	vo_dmid *m = new (std::nothrow) vo_dmid();
	if (!m) return bail(fail(ctx, VO_ERR_NOMEM, "out of host memory"));
	S->mid = m;
	m->nx = nx; m->ny = ey; m->J = J; m->R = R;
	m->shallow = own->max_cnt >= 0 && own->max_cnt <= 1;     // (a hint: the halo rows usually look like the rows next to them)
	const unsigned long long nslots = ncols * (J + 1);
	m->pool_cap = std::max(65536ull + (unsigned long long)(J + 1) * (total / 4), ctx->pool_hint);
	rc = dalloc(ctx, &m->slots, nslots);
	if (rc == VO_OK) rc = dalloc(ctx, &m->pool, m->pool_cap);
	if (rc == VO_OK) rc = dalloc(ctx, &m->flags, 2 * ncols);
	S->k_in = k_in;
	if (rc == VO_OK) rc = S->plan.init(ctx, nx, J, k_in);
	const unsigned long long nmask = 2ull * ey * S->plan.tiles_x, ntiles = (unsigned long long)S->plan.tiles_xw * ey;
	if (rc == VO_OK) rc = dalloc(ctx, &m->tilemask, nmask);
	if (rc == VO_OK) rc = dalloc(ctx, &S->thr, total);
	if (rc == VO_OK) rc = dalloc(ctx, &S->big_tiles, ntiles);
	if (rc == VO_OK) rc = dalloc(ctx, &S->multi_tiles, ntiles);
	const unsigned long long ntiles_b = (unsigned long long)S->plan.tiles_xw * (jp + jn + 2);
	if (rc == VO_OK) rc = dalloc(ctx, &S->big_tiles_b, ntiles_b);
	if (rc == VO_OK) rc = dalloc(ctx, &S->multi_tiles_b, ntiles_b);
	S->ntiles = ntiles;
	if (rc == VO_OK && ctx->tile_order) rc = dalloc(ctx, &S->est, ntiles + 4 * P1_NBUCKET);
	if (rc == VO_OK && ctx->tile_order) rc = dalloc(ctx, &S->order, 2 * ntiles);
	if (rc == VO_OK && !ctx->s_in && cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); rc = fail(ctx, VO_ERR_CUDA, "cudaStreamCreate"); }
	S->redo_cap = (unsigned int)std::min<unsigned long long>(nslots, 1ull << 22);
	if (rc == VO_OK) rc = dalloc(ctx, &S->redo_list, S->redo_cap);
	if (rc) return bail(rc);
	bool ok = cudaEventCreate(&S->ev0) == cudaSuccess && cudaEventCreate(&S->ev1) == cudaSuccess && cudaEventCreate(&S->ev2) == cudaSuccess &&
	          cudaEventCreateWithFlags(&S->ev_setup, cudaEventDisableTiming) == cudaSuccess &&
	          cudaEventCreateWithFlags(&S->ev_side, cudaEventDisableTiming) == cudaSuccess;
	if (!ok) { cudaGetLastError(); return bail(fail(ctx, VO_ERR_CUDA, "cudaEventCreate")); }
	cudaStream_t sm = ctx->stream;
	cudaEventRecord(S->ev0, sm);
	// the halos this rank SENDS are packed first (device side, no host round trip), so that the exchange can start
	// while pass 1 runs; `wait_stream` (the caller's communication stream) is made to wait for them
	if (has_prev && to_prev.d_off) {
		k_halo_pack<<<64, 256, 0, sm>>>(own->off, own->spans, 0ull, (unsigned long long)J * nx, static_cast<uint32_t *>(to_prev.d_off),
		                                static_cast<double2 *>(to_prev.d_spans), to_prev.cap);
		ctx->launches++;
	}
	if (has_next && to_next.d_off) {
		k_halo_pack<<<64, 256, 0, sm>>>(own->off, own->spans, (unsigned long long)(ny - J) * nx, (unsigned long long)ny * nx,
		                                static_cast<uint32_t *>(to_next.d_off), static_cast<double2 *>(to_next.d_spans), to_next.cap);
		ctx->launches++;
	}
	if ((to_prev.d_off || to_next.d_off)) {
		cudaEventRecord(S->ev1, sm);
		cudaStreamWaitEvent(static_cast<cudaStream_t>(wait_stream), S->ev1, 0);
	}
	cudaMemsetAsync(ctx->d_ctr, 0, NCTR * sizeof(unsigned long long), sm);
	cudaMemsetAsync(m->tilemask, 0, nmask * sizeof(unsigned long long), sm);
	if (S->est) cudaMemsetAsync(S->est, 0, (ntiles + 4 * P1_NBUCKET) * sizeof(unsigned int), sm);
	// own rows into the extended volume: offsets shifted by the reserved region of the previous halo
	cudaMemcpyAsync(S->ext->off + (size_t)jp * nx, own->off, (nown + 1) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, sm);
	if (cap_prev) { k_rebase<<<blocks_for(nown + 1, 256), 256, 0, sm>>>(S->ext->off + (size_t)jp * nx, nown + 1, 0u, (uint32_t)cap_prev); ctx->launches++; }
	if (own->nspans) cudaMemcpyAsync(S->ext->spans + cap_prev, own->spans, own->nspans * sizeof(double2), cudaMemcpyDeviceToDevice, sm);
	// "overlap" (default): the rows whose thresholds only read own rows start now (a few SMs stay free for the
	// exchange's kernels and for what follows) and the boundary rows join them from the side stream in vo_slab_finish.
	// "serial": everything in vo_slab_finish, as one launch set after the exchange (C5 slabs on 2 GPUs: 1.72 ms per
	// step against 1.39 ms overlapped; a single GPU takes 1.26 ms for the same rows).
	S->overlapped = ctx->slab_overlap;
	cudaEventRecord(S->ev_setup, sm);                       // (the side stream of vo_slab_finish starts from here)
	if (S->overlapped) slab_pass1_rows(S, jp + (jp ? 1 : 0), jp + ny - (jn ? 1 : 0), 0, 0, std::max(1, ctx->slab_reserve));
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) return bail(fail(ctx, VO_ERR_CUDA, std::string("slab_begin: ") + cudaGetErrorString(e)));
	*out = S;
	return VO_OK;
}

int slab_finish(vo_slab *S, const void *d_off_prev, const void *d_spans_prev, uint64_t n_prev,
                const void *d_off_next, const void *d_spans_next, uint64_t n_next, vo_dvol **out, double *ms1, double *ms2)
{
	vo_ctx *ctx = S->ctx;
	cudaStream_t sm = ctx->stream;
	const int nx = S->nx, ny = S->ny, jp = S->jp, jn = S->jn;
	if ((jp && (n_prev > S->cap_prev || !d_off_prev)) || (jn && (n_next > S->cap_next || !d_off_next)))
		return fail(ctx, VO_ERR_OVERFLOW, "halo larger than its reserved region");
	// Overlapped: the halos and pass 1 of the boundary rows go to the side stream, so that they run NEXT TO the launch
	// set of the interior rows (which left a few SMs free; its CTAs retire one by one and the boundary CTAs take their
	// place) instead of adding their own tail after it. Serial: everything on the main stream.
	cudaStream_t sh = S->overlapped ? ctx->s_in : sm;
	if (S->overlapped) cudaStreamWaitEvent(sh, S->ev_setup, 0);
	if (jp) {
		const unsigned long long n = (unsigned long long)jp * nx;          // (the entry after the last halo column is the first own offset)
		cudaMemcpyAsync(S->ext->off, d_off_prev, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, sh);
		k_rebase<<<blocks_for(n, 256), 256, 0, sh>>>(S->ext->off, n, 0u, (uint32_t)(S->cap_prev - n_prev));
		ctx->launches++;
		if (n_prev) cudaMemcpyAsync(S->ext->spans + (S->cap_prev - n_prev), d_spans_prev, n_prev * sizeof(double2), cudaMemcpyDeviceToDevice, sh);
	}
	if (jn) {
		const unsigned long long n = (unsigned long long)jn * nx + 1;
		uint32_t *dst = S->ext->off + (size_t)(jp + ny) * nx;
		cudaMemcpyAsync(dst, d_off_next, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, sh);
		k_rebase<<<blocks_for(n, 256), 256, 0, sh>>>(dst, n, 0u, (uint32_t)(S->cap_prev + S->n_own));
		ctx->launches++;
		if (n_next) cudaMemcpyAsync(S->ext->spans + S->cap_prev + S->n_own, d_spans_next, n_next * sizeof(double2), cudaMemcpyDeviceToDevice, sh);
	}
	if (S->overlapped) {
		slab_pass1_rows(S, 0, jp ? jp + 1 : 0, jp + ny - 1, jn ? jp + ny + jn : 0, 0, true);
		cudaEventRecord(S->ev_side, sh);
		cudaStreamWaitEvent(sm, S->ev_side, 0);
		slab_redo(S);                                       // (whatever either launch set left on the redo list)
	} else slab_pass1_rows(S, 0, jp + ny + jn, 0, 0, -1);
	cudaEventRecord(S->ev1, sm);
	unsigned long long h[NREAD];
	VO_TRY(read_counters(ctx, h));
	VO_CUDA(cudaGetLastError());
	if (h[0] > S->mid->pool_cap) { ctx->pool_hint = h[0] + h[0] / 4; return fail(ctx, VO_ERR_OVERFLOW, "mid pool too small"); }
	note_pass1_redo(ctx, h[2]);
	if (h[2] > S->redo_cap || h[4]) return fail(ctx, VO_ERR_OVERFLOW, "too many lists outgrew the fast running-union capacity");
	S->mid->pool_used = h[0];
	ctx->pool_hint = next_hint(ctx->pool_hint, h[0]);
	if (S->dual) {
		// (the ranks agreed beforehand that every slab qualifies: k_dual_check in vo_mg.cuh)
		if (h[14] || h[5] || h[2]) return fail(ctx, VO_ERR_OVERFLOW, "dual erosion: a column did not qualify after all");
		Tmp<uint8_t> dist(ctx);
		const unsigned long long ncols_ext = (unsigned long long)nx * S->ext->ny, ntmin = (unsigned long long)S->ext->ny * ((nx + P2_TX - 1) / P2_TX);
		VO_TRY(dalloc(ctx, &dist.p, ncols_ext + ntmin));
		k_empty_dist<<<(unsigned int)S->ext->ny, ED_THREADS, (size_t)((nx + 31) / 32 + (nx + P2_TX - 1) / P2_TX) * sizeof(uint32_t), sm>>>(S->ext->off, nx, dist.p, dist.p + ncols_ext);
		ctx->launches++;
		TableCache *tc = static_cast<TableCache *>(ctx->table_cache);
		VO_TRY(pass2_dual(ctx, S->mid, dist.p, tc->dt.reach, S->dual_zmin, S->dual_zmax, out, nullptr, jp, jp + ny));
	} else
	VO_TRY(pass2(ctx, S->mid, jp, jp + ny, out));
	VO_CUDA(cudaEventRecord(S->ev2, sm));
	VO_CUDA(cudaEventSynchronize(S->ev2));
	float t1 = 0, t2 = 0;
	cudaEventElapsedTime(&t1, S->ev0, S->ev1);
	cudaEventElapsedTime(&t2, S->ev1, S->ev2);
	if (ms1) *ms1 = t1;
	if (ms2) *ms2 = t2;
	return VO_OK;
}

// ---- dexeliser (the step before the path): compute_sign of src/vor3d/Dexelize.cpp:166-225, facets as work items ----
int dexelize_dev(vo_ctx *ctx, uint64_t nv, const double *verts, uint64_t nf, const int32_t *tris,
                 double ox, double oy, double spacing, int nx, int ny, vo_dvol **out)
{
	VO_TRY(check_dims(ctx, nx, ny));
	if (!(spacing > 0)) return fail(ctx, VO_ERR_ARG, "dexelize: spacing must be positive");
	if ((unsigned long long)nx * (unsigned long long)ny >= (1ull << 32) - 2 * DEX_CELLS) return fail(ctx, VO_ERR_OVERFLOW, "dexelize: grid too large for 32-bit cell numbers");
	if (nv >= (1ull << 31) || nf >= (1ull << 31) / 3) return fail(ctx, VO_ERR_OVERFLOW, "dexelize: mesh too large for 32-bit indices");
	if ((nv && !verts) || (nf && !tris)) return fail(ctx, VO_ERR_ARG, "dexelize: null mesh arrays");
	const unsigned long long ncols = (unsigned long long)nx * ny;
	vo_dvol *v = nullptr;
	VO_TRY(new_dvol(ctx, nx, ny, &v));
	std::unique_ptr<vo_dvol, std::function<void(vo_dvol *)>> guard(v, [ctx](vo_dvol *p) { free_dvol(ctx, p); });
	if (ncols == 0 || nf == 0 || nv == 0) {
		VO_CUDA(cudaMemsetAsync(v->off, 0, (ncols + 1) * sizeof(uint32_t), ctx->stream));
		VO_TRY(dalloc(ctx, &v->spans, 0));
		*out = guard.release();
		return VO_OK;
	}
	Tmp<double> dV(ctx), raw(ctx);
	Tmp<int> dF(ctx);
	Tmp<int4> box(ctx);
	Tmp<uint32_t> nchunk(ctx), chunk_off(ctx), cnt(ctx), raw_off(ctx), pairs(ctx);
	Tmp<unsigned int> bad(ctx);
	VO_TRY(dalloc(ctx, &dV.p, 3 * nv));
	VO_TRY(dalloc(ctx, &dF.p, 3 * nf));
	VO_TRY(dalloc(ctx, &box.p, nf));
	VO_TRY(dalloc(ctx, &nchunk.p, nf));
	VO_TRY(dalloc(ctx, &chunk_off.p, nf + 1));
	VO_TRY(dalloc(ctx, &cnt.p, ncols));
	VO_TRY(dalloc(ctx, &raw_off.p, ncols + 1));
	VO_TRY(dalloc(ctx, &pairs.p, ncols));
	VO_TRY(dalloc(ctx, &bad.p, 1));
	VO_CUDA(cudaMemcpyAsync(dV.p, verts, 3 * nv * sizeof(double), cudaMemcpyDefault, ctx->stream));
	VO_CUDA(cudaMemcpyAsync(dF.p, tris, 3 * nf * sizeof(int), cudaMemcpyDefault, ctx->stream));
	VO_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(unsigned int), ctx->stream));
	VO_CUDA(cudaMemsetAsync(cnt.p, 0, ncols * sizeof(uint32_t), ctx->stream));
	DexArgs a{};
	a.V = dV.p; a.F = dF.p; a.nv = (unsigned int)nv; a.nf = (unsigned int)nf;
	a.ox = ox; a.oy = oy; a.spacing = spacing; a.nx = nx; a.ny = ny;
	a.box = box.p; a.nchunk = nchunk.p; a.chunk_off = chunk_off.p; a.cnt = cnt.p; a.bad = bad.p;
	k_dex_plan<<<blocks_for(nf, 256), 256, 0, ctx->stream>>>(a);
	ctx->launches++;
	unsigned long long nchunks = 0;
	{
		// scan_counts checks its total against the uint32 range of the CSR offsets, which also bounds the chunk ids
		int rc = scan_counts(ctx, nchunk.p, nf, chunk_off.p, &nchunks);
		if (rc == VO_ERR_OVERFLOW) return fail(ctx, rc, "dexelize: too many (facet, column-chunk) pairs");
		VO_TRY(rc);
	}
	unsigned int h_bad = 0;
	VO_CUDA(cudaMemcpyAsync(&h_bad, bad.p, sizeof(h_bad), cudaMemcpyDeviceToHost, ctx->stream));
	VO_CUDA(cudaStreamSynchronize(ctx->stream));
	if (h_bad) return fail(ctx, VO_ERR_ARG, "dexelize: a facet names a vertex outside [0, nv)");
	a.nchunks = nchunks;
	const unsigned int hit_blocks = blocks_for(nchunks * 32ull, 256);
	if (nchunks) { k_dex_hits<false><<<hit_blocks, 256, 0, ctx->stream>>>(a); ctx->launches++; }
	unsigned long long total_raw = 0, total = 0;
	{
		int rc = scan_counts(ctx, cnt.p, ncols, raw_off.p, &total_raw);
		if (rc == VO_ERR_OVERFLOW) return fail(ctx, rc, "dexelize: more than 2^32-1 crossings");
		VO_TRY(rc);
	}
	k_dex_pairs<<<blocks_for(ncols, 256), 256, 0, ctx->stream>>>(cnt.p, ncols, pairs.p);
	ctx->launches++;
	VO_TRY(scan_counts(ctx, pairs.p, ncols, v->off, &total));
	VO_TRY(dalloc(ctx, &raw.p, total_raw));
	VO_TRY(dalloc(ctx, &v->spans, total));
	v->nspans = total;
	a.raw_off = raw_off.p; a.raw = raw.p; a.out_off = v->off; a.out = v->spans;
	if (total_raw) {
		VO_CUDA(cudaMemsetAsync(cnt.p, 0, ncols * sizeof(uint32_t), ctx->stream));
		k_dex_hits<true><<<hit_blocks, 256, 0, ctx->stream>>>(a);
		k_dex_sort<<<blocks_for(ncols, 256), 256, 0, ctx->stream>>>(a, ncols);
		ctx->launches += 2;
	}
	VO_CUDA(cudaGetLastError());
	*out = guard.release();
	return VO_OK;
}

int concat_rows(vo_ctx *ctx, const vo_dvol *a, const vo_dvol *b, const vo_dvol *c, vo_dvol **out);

struct DeviceGuard {
	int prev = -1;
	explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
	~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

} // namespace

// ---------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------
extern "C" {

const char *vo_version(void) { return "voroffset_b200 0.1 (sm_100a)"; }
int vo_span_bytes(void) { return (int)sizeof(double2); }

int vo_create(int device, vo_ctx **out)
{
	if (!out) return VO_ERR_ARG;
	*out = nullptr;
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); return VO_ERR_CUDA; }
	if (device < 0 || device >= ndev) return VO_ERR_ARG;
	vo_ctx *ctx = new (std::nothrow) vo_ctx();
	if (!ctx) return VO_ERR_NOMEM;
	ctx->device = device;
	bool ok = cudaSetDevice(device) == cudaSuccess;
	ok = ok && cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess;
	for (int i = 0; i < 3 && ok; ++i) ok = cudaEventCreate(&ctx->ev[i]) == cudaSuccess;
	for (int i = 0; i < 8 && ok; ++i) ok = cudaEventCreate(&ctx->mark[i]) == cudaSuccess;
	for (int i = 0; i < 4 && ok; ++i) ok = cudaEventCreate(&ctx->kev[i]) == cudaSuccess;
	ok = ok && cudaMalloc((void **)&ctx->d_ctr, (NCTR + 8 + NCTR) * sizeof(unsigned long long)) == cudaSuccess;   // [NCTR]: erosion's "data outside the z range" flag; [NCTR + 8 ...]: tile cursors of a second, concurrent launch set (slab boundary rows)
	if (ok) {
		// keep freed blocks in the stream-ordered pool: steady-state calls then allocate without the driver
		cudaMemPool_t pool;
		if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
			uint64_t thr = UINT64_MAX;
			cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
		}
	}
	if (ok) {
		size_t free_b = 0, total_b = 0;
		if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && total_b) ctx->big_free_limit = total_b / 3;
		else cudaGetLastError();
	}
	{ std::lock_guard<std::mutex> lk(g_host_mu); ++g_live_contexts; }
	if (!ok) { cudaGetLastError(); vo_destroy(ctx); return VO_ERR_CUDA; }
	*out = ctx;
	return VO_OK;
}

void vo_destroy(vo_ctx *ctx)
{
	if (!ctx) return;
	DeviceGuard g(ctx->device);
	if (ctx->stream) cudaStreamSynchronize(ctx->stream);
	free_table_cache(ctx);
	if (ctx->stream) { release_big_free(ctx, 0); cudaStreamSynchronize(ctx->stream); }
	if (ctx->d_ctr) cudaFree(ctx->d_ctr);
	if (ctx->ovf) cudaFree(ctx->ovf);
	if (ctx->scan_state) cudaFree(ctx->scan_state);
	for (auto &e : ctx->ev) if (e) cudaEventDestroy(e);
	for (auto &e : ctx->mark) if (e) cudaEventDestroy(e);
	for (auto &e : ctx->kev) if (e) cudaEventDestroy(e);
	for (auto &e : ctx->pipe_ev) if (e) cudaEventDestroy(e);
	for (auto &st : ctx->s_p) if (st) cudaStreamDestroy(st);
	for (auto &st : ctx->s_hi) if (st) cudaStreamDestroy(st);
	if (ctx->s_ctl) cudaStreamDestroy(ctx->s_ctl);
	if (ctx->s_out2) cudaStreamDestroy(ctx->s_out2);
	if (ctx->s_mid) cudaStreamDestroy(ctx->s_mid);
	if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
	if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
	if (ctx->stream) cudaStreamDestroy(ctx->stream);
	delete ctx;
	// pinned result blocks waiting for reuse (vo_free) are process-wide; the last context releases them
	std::lock_guard<std::mutex> lk(g_host_mu);
	if (--g_live_contexts <= 0) {
		g_live_contexts = 0;
		for (auto &blk : g_host_cache) cudaFreeHost(blk.first);
		g_host_cache.clear();
	}
}

const char *vo_last_error(const vo_ctx *ctx) { return ctx ? ctx->err.c_str() : "no context"; }
void *vo_stream(const vo_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
uint64_t vo_launch_count(const vo_ctx *ctx) { return ctx ? ctx->launches : 0; }

int vo_set_option(vo_ctx *ctx, const char *key, const char *value)
{
	if (!ctx || !key || !value) return VO_ERR_ARG;
	if (std::strcmp(key, "pipeline") == 0) {
		if (std::strcmp(value, "off") == 0) { ctx->no_pipeline = true; return VO_OK; }
		if (std::strcmp(value, "on") == 0 || std::strcmp(value, "auto") == 0) { ctx->no_pipeline = false; return VO_OK; }
	}
	if (std::strcmp(key, "erosion") == 0) {
		if (std::strcmp(value, "auto") == 0) { ctx->erosion_mode = 0; return VO_OK; }
		if (std::strcmp(value, "dual") == 0) { ctx->erosion_mode = 1; return VO_OK; }
		if (std::strcmp(value, "general") == 0) { ctx->erosion_mode = 2; return VO_OK; }
	}
	if (std::strcmp(key, "block_cache") == 0) {
		if (std::strcmp(value, "on") == 0) {
			DeviceGuard g(ctx->device);
			size_t free_b = 0, total_b = 0;
			ctx->block_cache = true;
			ctx->big_free_limit = cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && total_b ? total_b / 3 : (size_t)48 << 30;
			return VO_OK;
		}
		if (std::strcmp(value, "off") == 0) {
			DeviceGuard g(ctx->device);
			ctx->block_cache = false;
			ctx->big_free_limit = 0;       // blocks still out are released to the driver's pool when they come back
			release_big_free(ctx, 0);
			return VO_OK;
		}
	}
	if (std::strcmp(key, "slab") == 0) {
		if (std::strcmp(value, "overlap") == 0) { ctx->slab_overlap = true; return VO_OK; }
		if (std::strcmp(value, "serial") == 0) { ctx->slab_overlap = false; return VO_OK; }
	}
	if (std::strcmp(key, "slab_reserve") == 0) {
		const int n = std::atoi(value);
		if (n >= 1 && n <= 96) { ctx->slab_reserve = n; return VO_OK; }
	}
	if (std::strcmp(key, "bands") == 0) {
		const int n = std::atoi(value);
		if (n >= 3 && n <= 64) { ctx->pipe_bands = n; return VO_OK; }
	}
	if (std::strcmp(key, "band_weights") == 0) {
		std::vector<int> w;
		for (const char *c = value; *c;) {
			char *end = nullptr;
			const long v = std::strtol(c, &end, 10);
			if (end == c || v < 1 || v > 1000) return fail(ctx, VO_ERR_ARG, "band_weights: positive integers separated by commas or colons");
			w.push_back((int)v);
			c = (*end == ',' || *end == ':') ? end + 1 : end;
			if (end == c && *c) return fail(ctx, VO_ERR_ARG, "band_weights: positive integers separated by commas or colons");
		}
		if (w.size() == 1 && w[0] == 1) w.clear();           // "1": equal bands again
		if (!w.empty() && (w.size() < 3 || w.size() > 64)) return fail(ctx, VO_ERR_ARG, "band_weights: 3 to 64 bands");
		ctx->pipe_wts = w;
		return VO_OK;
	}
	if (std::strcmp(key, "pipe_quota") == 0) {
		const int n = std::atoi(value);
		if (n >= 0 && n <= 4096) { ctx->pipe_quota = n; return VO_OK; }
	}
	if (std::strcmp(key, "pipe_ctas") == 0) {
		const int n = std::atoi(value);
		if (n >= 0 && n <= 8) { ctx->pipe_ctas = n; return VO_OK; }
	}
	if (std::strcmp(key, "pipe_first_full") == 0 || std::strcmp(key, "pipe_interleave") == 0) {
		bool &flag = key[5] == 'f' ? ctx->pipe_first_full : ctx->pipe_interleave;
		if (std::strcmp(value, "on") == 0) { flag = true; return VO_OK; }
		if (std::strcmp(value, "off") == 0) { flag = false; return VO_OK; }
	}
	if (std::strcmp(key, "pipe_mid") == 0) {
		if (std::strcmp(value, "on") == 0) { ctx->pipe_mid = true; return VO_OK; }
		if (std::strcmp(value, "off") == 0) { ctx->pipe_mid = false; return VO_OK; }
	}
	if (std::strcmp(key, "copy_out") == 0) {
		const int n = std::atoi(value);
		if (n >= 0 && n <= 1024) { ctx->copy_out_ctas = n; return VO_OK; }
	}
	if (std::strcmp(key, "pipe_lean") == 0) {
		if (std::strcmp(value, "on") == 0) { ctx->pipe_lean = true; ctx->pipe_lists = ctx->pipe_redo = false; return VO_OK; }
		if (std::strcmp(value, "off") == 0) { ctx->pipe_lean = false; return VO_OK; }
	}
	if (std::strcmp(key, "multi_warps") == 0) {
		const int n = std::atoi(value);
		if (n >= 0 && n <= 64) { ctx->multi_warps = n; return VO_OK; }
	}
	if (std::strcmp(key, "copy_align") == 0) {
		const int n = std::atoi(value);
		if (n == 0 || (n >= 16 && n <= 65536 && (n & (n - 1)) == 0)) { ctx->copy_align = n; return VO_OK; }
	}
	if (std::strcmp(key, "copy_batch") == 0) {
		if (std::strcmp(value, "on") == 0) { ctx->copy_batch = true; return VO_OK; }
		if (std::strcmp(value, "off") == 0) { ctx->copy_batch = false; return VO_OK; }
	}
	if (std::strcmp(key, "pipe_ahead") == 0) {
		if (std::strcmp(value, "on") == 0) { ctx->pipe_ahead = true; return VO_OK; }
		if (std::strcmp(value, "off") == 0) { ctx->pipe_ahead = false; return VO_OK; }
	}
	if (std::strcmp(key, "pipe_order_one") == 0) {
		if (std::strcmp(value, "on") == 0) { ctx->pipe_order_one = true; return VO_OK; }
		if (std::strcmp(value, "off") == 0) { ctx->pipe_order_one = false; return VO_OK; }
	}
	if (std::strcmp(key, "pipe_warps0") == 0) {
		const int n = std::atoi(value);
		if (n >= 0 && n <= 64) { ctx->pipe_warps0 = n; return VO_OK; }
	}
	if (std::strcmp(key, "pipe_warps") == 0) {
		const int n = std::atoi(value);
		if (n >= 1 && n <= 64) { ctx->pipe_warps = n; return VO_OK; }
	}
	if (std::strcmp(key, "band_free") == 0) {
		const int n = std::atoi(value);
		if (n >= 0 && n <= 96) { ctx->band_free = n; return VO_OK; }
	}
	if (std::strcmp(key, "band_split") == 0) {
		const int n = std::atoi(value);
		if (n >= 1 && n <= 4) { ctx->band_split = n; return VO_OK; }
	}
	if (std::strcmp(key, "tile_ctas") == 0) {
		const int n = std::atoi(value);
		if (n >= 1 && n <= 8) { ctx->tile_ctas = n; return VO_OK; }
	}
	if (std::strcmp(key, "scan") == 0) {
		if (std::strcmp(value, "fused") == 0) { ctx->fused_scan = true; return VO_OK; }
		if (std::strcmp(value, "classic") == 0) { ctx->fused_scan = false; return VO_OK; }
	}
	if (std::strcmp(key, "pass2_union") == 0) {
		if (std::strcmp(value, "auto") == 0) { ctx->p2_mode = -1; ctx->p2_list_calls = 0; return VO_OK; }
		if (std::strcmp(value, "registers") == 0) { ctx->p2_mode = 0; return VO_OK; }
		if (std::strcmp(value, "lists") == 0) { ctx->p2_mode = 1; return VO_OK; }
	}
	if (std::strcmp(key, "tile_general") == 0) {
		if (std::strcmp(value, "auto") == 0) { ctx->gen_mode = -1; ctx->gen_inline_calls = 0; return VO_OK; }
		if (std::strcmp(value, "redo") == 0) { ctx->gen_mode = 0; return VO_OK; }
		if (std::strcmp(value, "inline") == 0) { ctx->gen_mode = 1; return VO_OK; }
	}
	if (std::strcmp(key, "cand_order") == 0) {
		if (std::strcmp(value, "auto") == 0) { ctx->cand_order = -1; return VO_OK; }
		if (std::strcmp(value, "column") == 0) { ctx->cand_order = 0; return VO_OK; }
		if (std::strcmp(value, "layer") == 0) { ctx->cand_order = 1; return VO_OK; }
	}
	if (std::strcmp(key, "tile_lean") == 0) {
		if (std::strcmp(value, "auto") == 0) { ctx->tile_lean = -1; return VO_OK; }
		if (std::strcmp(value, "on") == 0) { ctx->tile_lean = 1; return VO_OK; }
		if (std::strcmp(value, "off") == 0) { ctx->tile_lean = 0; return VO_OK; }
	}
	if (std::strcmp(key, "tile_dbuf") == 0) {
		if (std::strcmp(value, "auto") == 0) { ctx->tile_dbuf = -1; return VO_OK; }
		if (std::strcmp(value, "on") == 0) { ctx->tile_dbuf = 1; return VO_OK; }
		if (std::strcmp(value, "off") == 0) { ctx->tile_dbuf = 0; return VO_OK; }
	}
	if (std::strcmp(key, "tile_order") == 0) {
		if (std::strcmp(value, "off") == 0) { ctx->tile_order = false; return VO_OK; }
		if (std::strcmp(value, "on") == 0) { ctx->tile_order = true; return VO_OK; }
	}
#ifdef VO_KTRACE                                         // development builds only (ktrace.cuh, scripts/ktrace_view.py)
	if (std::strcmp(key, "pipe_dry") == 0) { ctx->pipe_dry = std::strtoull(value, nullptr, 0); return VO_OK; }
	if (std::strcmp(key, "ktrace") == 0) {
		DeviceGuard g(ctx->device);
		const unsigned int cap = (unsigned int)std::strtoul(value, nullptr, 0);
		KTraceBuf kb{nullptr, nullptr, cap};
		if (cap) {
			VO_CUDA(cudaMalloc(&kb.rec, 32ull * cap));
			VO_CUDA(cudaMalloc(&kb.count, sizeof(unsigned int)));
			VO_CUDA(cudaMemset(kb.count, 0, sizeof(unsigned int)));
		}
		VO_CUDA(cudaMemcpyToSymbol(c_kt, &kb, sizeof kb));
		return VO_OK;
	}
	if (std::strcmp(key, "ktrace_dump") == 0) {
		DeviceGuard g(ctx->device);
		VO_CUDA(cudaDeviceSynchronize());
		KTraceBuf kb;
		VO_CUDA(cudaMemcpyFromSymbol(&kb, c_kt, sizeof kb));
		if (!kb.rec) return fail(ctx, VO_ERR_ARG, "ktrace is off");
		unsigned int n = 0;
		VO_CUDA(cudaMemcpy(&n, kb.count, sizeof n, cudaMemcpyDeviceToHost));
		n = std::min(n, kb.cap);
		std::vector<unsigned long long> rec(4ull * n);
		if (n) VO_CUDA(cudaMemcpy(rec.data(), kb.rec, 32ull * n, cudaMemcpyDeviceToHost));
		VO_CUDA(cudaMemset(kb.count, 0, sizeof(unsigned int)));
		FILE *f = std::fopen(value, "w");
		if (!f) return fail(ctx, VO_ERR_ARG, "ktrace_dump: cannot open the file");
		std::fprintf(f, "id,sm,aux,block,t0,t1\n");
		for (unsigned int i = 0; i < n; ++i)
			std::fprintf(f, "%llu,%llu,%llu,%llu,%llu,%llu\n", rec[4ull * i] & 0xffffffffull, rec[4ull * i] >> 32, rec[4ull * i + 1] & 0xffffffffull,
			             rec[4ull * i + 1] >> 32, rec[4ull * i + 2], rec[4ull * i + 3]);
		std::fclose(f);
		return VO_OK;
	}
#endif
#ifdef VO_TILE_DEBUG                                     // development builds only (scripts/tile_costs.py): a raw device pointer
	if (std::strcmp(key, "tile_debug") == 0) {
		ctx->dbg_tiles = reinterpret_cast<unsigned long long *>(std::strtoull(value, nullptr, 0));
		return VO_OK;
	}
#endif
	if (std::strcmp(key, "pass1") == 0) {
		if (std::strcmp(value, "simple") == 0) { ctx->force_simple_pass1 = true; ctx->force_tile_pass1 = false; return VO_OK; }
		if (std::strcmp(value, "tile") == 0) { ctx->force_simple_pass1 = false; ctx->force_tile_pass1 = true; return VO_OK; }
		if (std::strcmp(value, "auto") == 0) { ctx->force_simple_pass1 = false; ctx->force_tile_pass1 = false; return VO_OK; }
	}
	return fail(ctx, VO_ERR_ARG, "unknown option");
}

int vo_mark(vo_ctx *ctx, int slot)
{
	if (!ctx || slot < 0 || slot >= 8) return VO_ERR_ARG;
	DeviceGuard g(ctx->device);
	VO_CUDA(cudaEventRecord(ctx->mark[slot], ctx->stream));
	return VO_OK;
}

int vo_elapsed_ms(vo_ctx *ctx, int slot_a, int slot_b, double *ms)
{
	if (!ctx || !ms || slot_a < 0 || slot_a >= 8 || slot_b < 0 || slot_b >= 8) return VO_ERR_ARG;
	DeviceGuard g(ctx->device);
	VO_CUDA(cudaEventSynchronize(ctx->mark[slot_b]));
	float t = 0;
	VO_CUDA(cudaEventElapsedTime(&t, ctx->mark[slot_a], ctx->mark[slot_b]));
	*ms = t;
	return VO_OK;
}

int vo_last_profile(vo_ctx *ctx, double *k_pass1_ms, double *k_pass2_ms)
{
	if (!ctx) return VO_ERR_ARG;
	DeviceGuard g(ctx->device);
	VO_CUDA(cudaStreamSynchronize(ctx->stream));
	float t = 0;
	if (k_pass1_ms) { *k_pass1_ms = 0; if (ctx->kev_valid[0]) { VO_CUDA(cudaEventElapsedTime(&t, ctx->kev[0], ctx->kev[1])); *k_pass1_ms = t; } }
	if (k_pass2_ms) { *k_pass2_ms = 0; if (ctx->kev_valid[1]) { VO_CUDA(cudaEventElapsedTime(&t, ctx->kev[2], ctx->kev[3])); *k_pass2_ms = t; } }
	return VO_OK;
}

int vo_dvol_from_device(vo_ctx *ctx, int nx, int ny, const void *d_off, const void *d_spans, uint64_t nspans, vo_dvol **out)
{
	if (!ctx || !out || !d_off) return VO_ERR_ARG;
	ctx->err.clear();
	DeviceGuard g(ctx->device);
	VO_TRY(check_dims(ctx, nx, ny));
	if (nspans >= (1ull << 32)) return fail(ctx, VO_ERR_OVERFLOW, "more than 2^32-1 intervals");
	const unsigned long long n = (unsigned long long)nx * ny;
	vo_dvol *v = nullptr;
	VO_TRY(new_dvol(ctx, nx, ny, &v));
	int rc = dalloc(ctx, &v->spans, nspans);
	if (rc) { free_dvol(ctx, v); return rc; }
	v->nspans = nspans;
	cudaError_t e = cudaMemcpyAsync(v->off, d_off, (n + 1) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream);
	if (e == cudaSuccess && nspans) e = cudaMemcpyAsync(v->spans, d_spans, nspans * sizeof(double2), cudaMemcpyDeviceToDevice, ctx->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
	if (e != cudaSuccess) { free_dvol(ctx, v); cudaGetLastError(); return fail(ctx, VO_ERR_CUDA, std::string("from_device: ") + cudaGetErrorString(e)); }
	*out = v;
	return VO_OK;
}

void vo_free(void *p)
{
	if (!p) return;
	std::lock_guard<std::mutex> lk(g_host_mu);
	auto it = g_host_live.find(p);
	if (it == g_host_live.end()) return;
	std::pair<void *, size_t> blk(it->first, it->second);
	g_host_live.erase(it);
	if (g_host_cache.size() < 8) g_host_cache.push_back(blk);
	else cudaFreeHost(blk.first);
}

int vo_dvol_upload(vo_ctx *ctx, int nx, int ny, const uint32_t *off, const double *spans, vo_dvol **out)
{
	if (!ctx || !out) return VO_ERR_ARG;
	ctx->err.clear();
	DeviceGuard g(ctx->device);
	VO_TRY(upload(ctx, nx, ny, off, spans, out));
	VO_CUDA(cudaStreamSynchronize(ctx->stream));
	return VO_OK;
}

int vo_dvol_download(vo_ctx *ctx, const vo_dvol *v, uint32_t *off, double *spans)
{
	if (!ctx || !v || !off) return VO_ERR_ARG;
	ctx->err.clear();
	DeviceGuard g(ctx->device);
	const unsigned long long n = (unsigned long long)v->nx * v->ny;
	VO_CUDA(cudaMemcpyAsync(off, v->off, (n + 1) * sizeof(uint32_t), cudaMemcpyDefault, ctx->stream));
	if (v->nspans) {
		if (!spans) return fail(ctx, VO_ERR_ARG, "spans is NULL");
		VO_CUDA(cudaMemcpyAsync(spans, v->spans, v->nspans * sizeof(double2), cudaMemcpyDefault, ctx->stream));
	}
	VO_CUDA(cudaStreamSynchronize(ctx->stream));
	return VO_OK;
}

int vo_dvol_info(const vo_dvol *v, int *nx, int *ny, uint64_t *nspans, const void **d_off, const void **d_spans)
{
	if (!v) return VO_ERR_ARG;
	if (nx) *nx = v->nx;
	if (ny) *ny = v->ny;
	if (nspans) *nspans = v->nspans;
	if (d_off) *d_off = v->off;
	if (d_spans) *d_spans = v->spans;
	return VO_OK;
}

void vo_dvol_free(vo_ctx *ctx, vo_dvol *v)
{
	if (!ctx || !v) return;
	DeviceGuard g(ctx->device);
	free_dvol(ctx, v);
}

int vo_dvol_rows(vo_ctx *ctx, const vo_dvol *v, int y0, int y1, vo_dvol **out)
{
	if (!ctx || !v || !out) return VO_ERR_ARG;
	ctx->err.clear();
	DeviceGuard g(ctx->device);
	if (y0 < 0 || y1 > v->ny || y0 > y1) return fail(ctx, VO_ERR_ARG, "row range outside the volume");
	const unsigned long long c0 = (unsigned long long)y0 * v->nx, c1 = (unsigned long long)y1 * v->nx;
	uint32_t ends[2] = {0, 0};
	VO_CUDA(cudaMemcpyAsync(&ends[0], v->off + c0, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
	VO_CUDA(cudaMemcpyAsync(&ends[1], v->off + c1, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
	VO_CUDA(cudaStreamSynchronize(ctx->stream));
	vo_dvol *r = nullptr;
	VO_TRY(new_dvol(ctx, v->nx, y1 - y0, &r));
	r->nspans = ends[1] - ends[0];
	r->max_cnt = v->max_cnt;
	int rc = dalloc(ctx, &r->spans, r->nspans);
	if (rc) { free_dvol(ctx, r); return rc; }
	cudaMemcpyAsync(r->off, v->off + c0, (c1 - c0 + 1) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream);
	if (r->nspans) cudaMemcpyAsync(r->spans, v->spans + ends[0], r->nspans * sizeof(double2), cudaMemcpyDeviceToDevice, ctx->stream);
	k_rebase<<<blocks_for(c1 - c0 + 1, 256), 256, 0, ctx->stream>>>(r->off, c1 - c0 + 1, ends[0], 0u);
	ctx->launches++;
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) { free_dvol(ctx, r); return fail(ctx, VO_ERR_CUDA, cudaGetErrorString(e)); }
	*out = r;
	return VO_OK;
}

int vo_dvol_rows_to(vo_ctx *ctx, const vo_dvol *v, int y0, int y1, void *d_off, void *d_spans, uint64_t cap_spans, uint64_t *nspans)
{
	if (!ctx || !v || !d_off || !nspans) return VO_ERR_ARG;
	ctx->err.clear();
	DeviceGuard g(ctx->device);
	if (y0 < 0 || y1 > v->ny || y0 > y1) return fail(ctx, VO_ERR_ARG, "row range outside the volume");
	const unsigned long long c0 = (unsigned long long)y0 * v->nx, c1 = (unsigned long long)y1 * v->nx;
	uint32_t ends[2] = {0, 0};
	VO_CUDA(cudaMemcpyAsync(&ends[0], v->off + c0, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
	VO_CUDA(cudaMemcpyAsync(&ends[1], v->off + c1, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
	VO_CUDA(cudaStreamSynchronize(ctx->stream));
	const uint64_t n = ends[1] - ends[0];
	*nspans = n;
	VO_CUDA(cudaMemcpyAsync(d_off, v->off + c0, (c1 - c0 + 1) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
	k_rebase<<<blocks_for(c1 - c0 + 1, 256), 256, 0, ctx->stream>>>(static_cast<uint32_t *>(d_off), c1 - c0 + 1, ends[0], 0u);
	ctx->launches++;
	if (n && n <= cap_spans) {
		if (!d_spans) return fail(ctx, VO_ERR_ARG, "d_spans is NULL");
		VO_CUDA(cudaMemcpyAsync(d_spans, v->spans + ends[0], n * sizeof(double2), cudaMemcpyDeviceToDevice, ctx->stream));
	}
	VO_CUDA(cudaStreamSynchronize(ctx->stream));
	return VO_OK;
}

int vo_dvol_concat_rows(vo_ctx *ctx, const vo_dvol *a, const vo_dvol *b, const vo_dvol *c, vo_dvol **out)
{
	if (!ctx || !out) return VO_ERR_ARG;
	ctx->err.clear();
	DeviceGuard g(ctx->device);
	return concat_rows(ctx, a, b, c, out);
}

} // extern "C"

namespace {
int concat_rows(vo_ctx *ctx, const vo_dvol *a, const vo_dvol *b, const vo_dvol *c, vo_dvol **out)
{
	const vo_dvol *parts[3] = {a, b, c};
	int nx = -1, ny = 0;
	uint64_t total = 0;
	for (auto p : parts) {
		if (!p) continue;
		if (nx < 0) nx = p->nx;
		if (p->nx != nx) return fail(ctx, VO_ERR_ARG, "slabs must share nx");
		ny += p->ny;
		total += p->nspans;
	}
	if (nx < 0) return fail(ctx, VO_ERR_ARG, "nothing to concatenate");
	if (total >= (1ull << 32)) return fail(ctx, VO_ERR_OVERFLOW, "result has more than 2^32-1 intervals");
	vo_dvol *r = nullptr;
	VO_TRY(new_dvol(ctx, nx, ny, &r));
	r->nspans = total;
	r->max_cnt = 0;
	for (auto p : parts) if (p) r->max_cnt = (r->max_cnt < 0 || p->max_cnt < 0) ? -1 : std::max(r->max_cnt, p->max_cnt);
	int rc = dalloc(ctx, &r->spans, total);
	if (rc) { free_dvol(ctx, r); return rc; }
	unsigned long long col = 0;
	uint64_t sp = 0;
	for (auto p : parts) {
		if (!p) continue;
		const unsigned long long n = (unsigned long long)p->nx * p->ny;
		cudaMemcpyAsync(r->off + col, p->off, (n + 1) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream);
		if (sp) { k_rebase<<<blocks_for(n + 1, 256), 256, 0, ctx->stream>>>(r->off + col, n + 1, 0u, (uint32_t)sp); ctx->launches++; }
		if (p->nspans) cudaMemcpyAsync(r->spans + sp, p->spans, p->nspans * sizeof(double2), cudaMemcpyDeviceToDevice, ctx->stream);
		col += n;
		sp += p->nspans;
	}
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) { free_dvol(ctx, r); return fail(ctx, VO_ERR_CUDA, cudaGetErrorString(e)); }
	*out = r;
	return VO_OK;
}
} // namespace

extern "C" {

int vo_morph3d_dev(vo_ctx *ctx, int op, int method, const vo_dvol *in, double zmin, double zmax, double radius,
                   vo_dvol **out, double *ms_pass1, double *ms_pass2)
{
	if (!ctx || !in || !out) return VO_ERR_ARG;
	ctx->err.clear();
	DeviceGuard g(ctx->device);
	PassTimes pt;
	VO_TRY(morph3d_dev(ctx, op, method, in, zmin, zmax, radius, out, &pt));
	VO_CUDA(cudaStreamSynchronize(ctx->stream));
	if (ms_pass1) *ms_pass1 = pt.ms1;
	if (ms_pass2) *ms_pass2 = pt.ms2;
	return VO_OK;
}

int vo_xor3d_dev(vo_ctx *ctx, const vo_dvol *a, const vo_dvol *b, double zmin, double zmax, double spacing,
                 vo_dvol **out, double *volume)
{
	if (!ctx || !a || !b || !out) return VO_ERR_ARG;
	ctx->err.clear();
	DeviceGuard g(ctx->device);
	VO_TRY(xor_dev(ctx, a, b, zmin, zmax, spacing, out, volume));
	VO_CUDA(cudaStreamSynchronize(ctx->stream));
	return VO_OK;
}

int vo_pass1_dev(vo_ctx *ctx, const vo_dvol *in, double radius, vo_dmid **mid, double *ms)
{
	if (!ctx || !in || !mid) return VO_ERR_ARG;
	ctx->err.clear();
	DeviceGuard g(ctx->device);
	VO_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
	VO_TRY(pass1(ctx, in, radius, mid));
	VO_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
	VO_CUDA(cudaEventSynchronize(ctx->ev[1]));
	float t = 0;
	cudaEventElapsedTime(&t, ctx->ev[0], ctx->ev[1]);
	if (ms) *ms = t;
	return VO_OK;
}

int vo_pass2_dev(vo_ctx *ctx, const vo_dmid *mid, int y0, int y1, vo_dvol **out, double *ms)
{
	if (!ctx || !mid || !out) return VO_ERR_ARG;
	ctx->err.clear();
	DeviceGuard g(ctx->device);
	VO_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
	VO_TRY(pass2(ctx, mid, y0, y1, out));
	VO_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
	VO_CUDA(cudaEventSynchronize(ctx->ev[1]));
	float t = 0;
	cudaEventElapsedTime(&t, ctx->ev[0], ctx->ev[1]);
	if (ms) *ms = t;
	return VO_OK;
}

int vo_dexelize_dev(vo_ctx *ctx, uint64_t nv, const double *verts, uint64_t nf, const int32_t *tris,
                    double origin_x, double origin_y, double spacing, int nx, int ny, vo_dvol **out, double *ms)
{
	if (!ctx || !out) return VO_ERR_ARG;
	ctx->err.clear();
	DeviceGuard g(ctx->device);
	VO_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
	VO_TRY(dexelize_dev(ctx, nv, verts, nf, tris, origin_x, origin_y, spacing, nx, ny, out));
	VO_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
	VO_CUDA(cudaEventSynchronize(ctx->ev[1]));
	float t = 0;
	cudaEventElapsedTime(&t, ctx->ev[0], ctx->ev[1]);
	if (ms) *ms = t;
	return VO_OK;
}

void vo_dmid_free(vo_ctx *ctx, vo_dmid *m)
{
	if (!ctx || !m) return;
	DeviceGuard g(ctx->device);
	dfree(ctx, m->slots);
	dfree(ctx, m->pool);
	dfree(ctx, m->flags);
	dfree(ctx, m->tilemask);
	delete m;
}

int vo_dmid_info(const vo_dmid *m, int *nx, int *ny, int *classes, uint64_t *bytes)
{
	if (!m) return VO_ERR_ARG;
	if (nx) *nx = m->nx;
	if (ny) *ny = m->ny;
	if (classes) *classes = m->J + 1;
	if (bytes) *bytes = ((uint64_t)m->nx * m->ny * (m->J + 1) + m->pool_used) * sizeof(double2);
	return VO_OK;
}

int vo_slab_begin(vo_ctx *ctx, const vo_dvol *own, double radius, int has_prev, int has_next, uint64_t cap_prev, uint64_t cap_next,
                  void *d_off_to_prev, void *d_spans_to_prev, uint64_t cap_to_prev,
                  void *d_off_to_next, void *d_spans_to_next, uint64_t cap_to_next, void *comm_stream, vo_slab **out)
{
	if (!ctx || !own || !out) return VO_ERR_ARG;
	ctx->err.clear();
	DeviceGuard g(ctx->device);
	*out = nullptr;
	return slab_begin(ctx, own, radius, has_prev, has_next, cap_prev, cap_next, HaloOut{d_off_to_prev, d_spans_to_prev, cap_to_prev},
	                  HaloOut{d_off_to_next, d_spans_to_next, cap_to_next}, comm_stream, out);
}

int vo_slab_finish(vo_ctx *ctx, vo_slab *slab, const void *d_off_prev, const void *d_spans_prev, uint64_t n_prev,
                   const void *d_off_next, const void *d_spans_next, uint64_t n_next, vo_dvol **out, double *ms_pass1, double *ms_pass2)
{
	if (!ctx || !slab || !out || slab->ctx != ctx) return VO_ERR_ARG;
	ctx->err.clear();
	DeviceGuard g(ctx->device);
	const int rc = slab_finish(slab, d_off_prev, d_spans_prev, n_prev, d_off_next, d_spans_next, n_next, out, ms_pass1, ms_pass2);
	free_slab(slab);
	return rc;
}

void vo_slab_abort(vo_ctx *ctx, vo_slab *slab)
{
	if (!ctx || !slab) return;
	DeviceGuard g(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	free_slab(slab);
}

int vo_morph2d_dev(vo_ctx *ctx, int op, const vo_dvol *rows, int width, double r, vo_dvol **out, double *ms)
{
	if (!ctx || !rows || !out) return VO_ERR_ARG;
	ctx->err.clear();
	DeviceGuard g(ctx->device);
	VO_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
	VO_TRY(morph2d_dev(ctx, op, rows, width, r, out));
	VO_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
	VO_CUDA(cudaEventSynchronize(ctx->ev[1]));
	float t = 0;
	cudaEventElapsedTime(&t, ctx->ev[0], ctx->ev[1]);
	if (ms) *ms = t;
	return VO_OK;
}

int vo_morph3d(vo_ctx *ctx, int op, int method, int nx, int ny, double zmin, double zmax,
               const uint32_t *off, const double *spans, double radius,
               uint32_t **out_off, double **out_spans, uint64_t *out_nspans, double *ms_pass1, double *ms_pass2)
{
	if (!ctx || !out_off || !out_spans) return VO_ERR_ARG;
	ctx->err.clear();
	DeviceGuard g(ctx->device);
	PassTimes pt;
	if (op == VO_OP_DILATION && method == VO_METHOD_OURS) {
		// large grids: upload, the two passes and the download overlap band by band
		const int prc = dilate_ours_pipelined(ctx, nx, ny, off, spans, radius, out_off, out_spans, out_nspans, &pt);
		if (prc == VO_OK) {
			if (ms_pass1) *ms_pass1 = pt.ms1;
			if (ms_pass2) *ms_pass2 = pt.ms2;
			return VO_OK;
		}
		if (prc != PIPE_NA) return prc;
		ctx->err.clear();
	}
	vo_dvol *in = nullptr, *res = nullptr;
	VO_TRY(upload(ctx, nx, ny, off, spans, &in));
	int rc = morph3d_dev(ctx, op, method, in, zmin, zmax, radius, &res, &pt);
	free_dvol(ctx, in);
	VO_TRY(rc);
	rc = download_new(ctx, res, out_off, out_spans, out_nspans);
	free_dvol(ctx, res);
	VO_TRY(rc);
	if (ms_pass1) *ms_pass1 = pt.ms1;
	if (ms_pass2) *ms_pass2 = pt.ms2;
	return VO_OK;
}

int vo_morph3d_rows(vo_ctx *ctx, int op, int method, int nx, int ny, double zmin, double zmax,
                    const uint32_t *off, const double *spans, double radius, int row0, int row1,
                    uint32_t **out_off, double **out_spans, uint64_t *out_nspans, double *ms_pass1, double *ms_pass2)
{
	if (!ctx || !out_off || !out_spans) return VO_ERR_ARG;
	ctx->err.clear();
	DeviceGuard g(ctx->device);
	if (nx < 0 || ny < 0 || !off) return fail(ctx, VO_ERR_ARG, "bad grid / offsets");
	if (row0 < 0 || row1 > ny || row0 > row1) return fail(ctx, VO_ERR_ARG, "row window outside the grid");
	PassTimes pt;
	if (op == VO_OP_DILATION && method == VO_METHOD_OURS && row0 < row1) {
		const int prc = dilate_ours_pipelined(ctx, nx, ny, off, spans, radius, out_off, out_spans, out_nspans, &pt, row0, row1);
		if (prc == VO_OK) {
			if (ms_pass1) *ms_pass1 = pt.ms1;
			if (ms_pass2) *ms_pass2 = pt.ms2;
			return VO_OK;
		}
		if (prc != PIPE_NA) return prc;
		ctx->err.clear();
	}
	// plain path: the window as a volume of its own (offsets rebased on the host), the operator, the rows asked for
	VO_TRY(check_dims(ctx, nx, ny));
	const unsigned long long n = (unsigned long long)nx * ny;
	const uint32_t obase = off[0];
	std::vector<uint32_t> rel;
	const uint32_t *o = off;
	if (obase) {
		rel.resize(n + 1);
		for (unsigned long long i = 0; i <= n; ++i) {
			if (off[i] < obase) return fail(ctx, VO_ERR_ARG, "offsets must be non-decreasing");
			rel[i] = off[i] - obase;
		}
		o = rel.data();
	}
	vo_dvol *in = nullptr, *res = nullptr, *rows = nullptr;
	VO_TRY(upload(ctx, nx, ny, o, spans ? spans + 2 * (size_t)obase : nullptr, &in));
	int rc = morph3d_dev(ctx, op, method, in, zmin, zmax, radius, &res, &pt);
	free_dvol(ctx, in);
	VO_TRY(rc);
	rc = vo_dvol_rows(ctx, res, row0, row1, &rows);
	free_dvol(ctx, res);
	VO_TRY(rc);
	rc = download_new(ctx, rows, out_off, out_spans, out_nspans);
	free_dvol(ctx, rows);
	VO_TRY(rc);
	if (ms_pass1) *ms_pass1 = pt.ms1;
	if (ms_pass2) *ms_pass2 = pt.ms2;
	return VO_OK;
}

int vo_morph2d(vo_ctx *ctx, int op, int rows, int width, const uint32_t *off, const double *spans, double r,
               uint32_t **out_off, double **out_spans, uint64_t *out_nspans, double *ms)
{
	if (!ctx || !out_off || !out_spans) return VO_ERR_ARG;
	ctx->err.clear();
	DeviceGuard g(ctx->device);
	if (rows < 0 || width < 0) return fail(ctx, VO_ERR_ARG, "negative image size");
	vo_dvol *in = nullptr, *res = nullptr;
	VO_TRY(upload(ctx, rows, 1, off, spans, &in));
	VO_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
	int rc = morph2d_dev(ctx, op, in, width, r, &res);
	free_dvol(ctx, in);
	VO_TRY(rc);
	VO_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
	rc = download_new(ctx, res, out_off, out_spans, out_nspans);
	free_dvol(ctx, res);
	VO_TRY(rc);
	float t = 0;
	cudaEventElapsedTime(&t, ctx->ev[0], ctx->ev[1]);
	if (ms) *ms = t;
	return VO_OK;
}

int vo_xor3d(vo_ctx *ctx, int nx, int ny, double zmin, double zmax, double spacing,
             const uint32_t *off_a, const double *spans_a, const uint32_t *off_b, const double *spans_b,
             uint32_t **out_off, double **out_spans, uint64_t *out_nspans, double *volume)
{
	if (!ctx || !out_off || !out_spans) return VO_ERR_ARG;
	ctx->err.clear();
	DeviceGuard g(ctx->device);
	vo_dvol *a = nullptr, *b = nullptr, *res = nullptr;
	VO_TRY(upload(ctx, nx, ny, off_a, spans_a, &a));
	int rc = upload(ctx, nx, ny, off_b, spans_b, &b);
	if (rc == VO_OK) rc = xor_dev(ctx, a, b, zmin, zmax, spacing, &res, volume);
	free_dvol(ctx, a);
	free_dvol(ctx, b);
	VO_TRY(rc);
	rc = download_new(ctx, res, out_off, out_spans, out_nspans);
	free_dvol(ctx, res);
	return rc;
}

} // extern "C"

#include "vo_mg.cuh"
