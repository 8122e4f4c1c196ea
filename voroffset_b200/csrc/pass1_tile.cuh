// 'ours', pass 1, tile kernel with exact 2D dominance pruning.
//
// What it computes is what k_pass1 (kernels.cuh) computes - for an output column x of row y and a radius
// class j the union over |dx| <= reach[j] of the neighbours' intervals grown by H[j][|dx|] - but only
// for the (candidate, class) pairs that can still matter, and only for the classes some consumer of
// pass 2 will read:
//
//  * the row segment [x0-J, x0+TX+J) is staged once in shared memory as a flat candidate array
//    (interval, segment column, thresholds) straight from the CSR;
//  * x-dominance (per candidate, per side): if the neighbouring column one step closer to the output has
//    an interval q with  max(a_q - a_p, b_p - b_q) <= Dmono[d],  Dmono[d] = min over d' >= d and live
//    classes j of H[j][d'-1] - H[j][d'],  then for EVERY class q's capped interval contains p's for all
//    outputs at distance >= d on that side, and p is skipped there;
//  * y-dominance (k_ythresh, per interval, per side): the same test against the column one row closer to
//    the consumer row, with  Emono[j] = min over j' >= j and d <= reach[j'] of H[j'-1][d] - H[j'][d]:
//    the consumer at row distance >= Ty on that side is served by the neighbouring row's class j-1 slot,
//    so p takes part only in the classes j < max(Ty_up, Ty_dn);
//  * a thread owns one output column: it collects its surviving candidates (a short list in shared
//    memory), then evaluates the classes 0 .. Tmax-1 only, where Tmax is the largest class any survivor
//    still needs; the two per-side maxima are published as flags so that pass 2 skips every other slot
//    without reading it (the mid volume stays unwritten there);
//  * classes are evaluated EIGHT AT A TIME in registers: a survivor is loaded once, its eight caps come
//    from one row of the transposed table with 128-bit shared loads, and the eight (lo, hi) running
//    unions are independent instruction streams (no dependent load chain per candidate). A class whose
//    union is not a single interval is flagged and redone by the general list-based path;
//  * lanes run along x: slot writes are full 128-byte lines.
//
// Why this is exact: containment is decided on the same table values the candidates are built from and
// fp64 subtraction is monotone, so "q's capped interval contains p's" carries over to the rounded
// endpoints; every dropped pair is contained in a pair one lattice step (in x or in y) closer to the
// consumer, |dx| + |dy| strictly decreases along such a chain, and the pair at (0,0) is never dropped -
// so the union every consumer sees is unchanged, bit for bit. This is the role the reference's
// Voronoi-vertex / power-diagram events play (Voronoi2D.cpp:329-586, SeparatePower2D.cpp:118-293:
// seeds are retired once their cell no longer reaches the sweep line) in a conservative, data-parallel
// form. On the C5 torus it keeps ~7 of ~38 candidates per column and ~13 of 33 classes.
#pragma once
#include <cooperative_groups.h>
#include <cooperative_groups/scan.h>

#include "kernels.cuh"

namespace vo {

constexpr int P1_TX = 128;      // output columns (= threads) per CTA
constexpr int P1_LCAP = 32;     // survivors listed per output column (more: the candidate range is re-scanned)
constexpr int P1_CB = 8;        // classes evaluated together in registers

// ---- y-direction dominance thresholds, one thread per column ----------------------------------------
struct YThreshArgs {
	int nx, ny, J;
	const uint32_t *off;
	const double2 *spans;
	const double *Emono;    // J+2: Emono[j], j = 1..J; Emono[J+1] = +inf
	uint16_t *ty;           // per interval: Ty_up | Ty_dn << 8, each in [1, J+1]
	unsigned long long c_begin, c_end;   // columns processed by this launch (a band of rows, or everything)
};

__device__ __forceinline__ int first_ge(const double *tab, int J, double need)
{
	int lo = 1, hi = J + 1;                       // tab is non-decreasing, tab[J+1] = +inf
	while (lo < hi) { const int mid = (lo + hi) >> 1; if (tab[mid] >= need) hi = mid; else lo = mid + 1; }
	return lo;
}

__global__ void __launch_bounds__(256) k_ythresh(YThreshArgs a)
{
	extern __shared__ double s_E[];
	for (int i = threadIdx.x; i < a.J + 2; i += blockDim.x) s_E[i] = a.Emono[i];
	__syncthreads();
	const unsigned long long c = a.c_begin + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= a.c_end) return;
	const uint32_t o0 = a.off[c], o1 = a.off[c + 1];
	if (o0 == o1) return;
	const int y = (int)(c / (unsigned)a.nx);
	const double inf = __longlong_as_double(0x7FF0000000000000LL);
	uint32_t u0 = 0, u1 = 0, d0 = 0, d1 = 0;
	if (y > 0) { u0 = a.off[c - a.nx]; u1 = a.off[c - a.nx + 1]; }
	if (y < a.ny - 1) { d0 = a.off[c + a.nx]; d1 = a.off[c + a.nx + 1]; }
	for (uint32_t k = o0; k < o1; ++k) {
		const double2 p = a.spans[k];
		const double m = 1e-9 + 1e-13 * (fabs(p.x) + fabs(p.y));
		double need_up = inf, need_dn = inf;
		for (uint32_t q = u0; q < u1; ++q) { const double2 v = a.spans[q]; need_up = fmin(need_up, fmax(v.x - p.x, p.y - v.y)); }
		for (uint32_t q = d0; q < d1; ++q) { const double2 v = a.spans[q]; need_dn = fmin(need_dn, fmax(v.x - p.x, p.y - v.y)); }
		const int tu = first_ge(s_E, a.J, need_up + m), td = first_ge(s_E, a.J, need_dn + m);
		a.ty[k] = (uint16_t)(tu | (td << 8));
	}
}

// ---- pass 1 ------------------------------------------------------------------------------------------
struct Pass1TileArgs {
	int nx, ny, J, cmax, tiles_x;
	unsigned int tile0;     // first tile of a launch over all tiles of a band of rows
	const uint32_t *off;
	const double2 *spans;
	const uint16_t *ty;     // k_ythresh output
	const double *Ht;       // (J+1) rows of JPP = roundup(J+1, 8) doubles: Ht[d*JPP + j] = H[j][d]
	const int *reach;       // J+1
	const double *Dmono;    // J+2
	double2 *mid;
	uint8_t *flags;         // [2][ny*nx]: classes needed by the consumer rows above ([0]) / below ([1])
	double2 *pool;
	unsigned long long *cursor;
	unsigned long long pool_cap;
	Redo redo;              // slot ids to be (re)done by k_pass1: list overflow and oversized tiles
	// Device-side dispatch (no host round trip between the launches):
	//   launch 1  <MULTI=false>, small candidate buffer, all tiles. Tiles with a multi-interval column go to
	//             multi_tiles, tiles with more than cmax candidates to big_tiles.
	//   launch 2  <MULTI=false>, large buffer, tiles = big_tiles.
	//   launch 3  <MULTI=true> (two hulls per class), tiles = multi_tiles.
	// A launch over a list uses the full grid; CTAs beyond *tiles_count exit at once.
	const unsigned int *tiles;        // NULL: all tiles
	const unsigned int *tiles_count;
	unsigned int *tiles_next;         // list launches: next list position to hand out
	unsigned int *big_tiles;          // NULL: oversized tiles go to the redo list (k_pass1)
	unsigned int *big_count;
	unsigned int *multi_tiles;
	unsigned int *multi_count;
};

__host__ __device__ inline int pass1_jpp(int J) { return (J + 1 + P1_CB - 1) / P1_CB * P1_CB; }

__host__ __device__ inline size_t pass1_tile_smem(int J, int cmax)
{
	const size_t JP = (size_t)J + 1, SEG = (size_t)P1_TX + 2 * J;
	size_t b = 0;
	b += (size_t)cmax * sizeof(double2);                    // candidates
	b += JP * pass1_jpp(J) * sizeof(double);                // Ht
	b += (JP + 1) * sizeof(double);                         // Dmono
	b += ((SEG + 2) & ~(size_t)1) * sizeof(uint32_t);       // segment offsets
	b += (size_t)cmax * sizeof(uint32_t);                   // first surviving output | width << 8 | column << 16 | T << 24
	b += ((JP + 1) & ~(size_t)1) * sizeof(int);             // reach
	b += (size_t)P1_LCAP * P1_TX * sizeof(uint16_t);        // survivor lists [s][thread]
	b += (size_t)cmax * sizeof(uint16_t);                   // y thresholds
	b += JP;                                                // largest class within reach of a distance
	return b + 32;
}

// Pool space for `n` entries, one atomic per converged group of threads instead of one per thread.
__device__ __forceinline__ unsigned long long pool_alloc(unsigned long long *cursor, unsigned int n)
{
	namespace cg = cooperative_groups;
	auto g = cg::coalesced_threads();
	const unsigned int pre = cg::exclusive_scan(g, n);
	const unsigned int total = g.shfl(pre + n, g.size() - 1);
	unsigned long long base = 0;
	if (g.thread_rank() == 0) base = atomicAdd(cursor, (unsigned long long)total);
	return g.shfl(base, 0) + pre;
}

// Per-thread view of the staged tile (phase 2).
struct TileThread {
	const double2 *cand;
	const double *Ht;
	const uint32_t *sv;        // first | width << 8 | column << 16 | (T-1) << 24 | layer << 30
	const uint16_t *list;
	const uint8_t *jmax;
	int JP, JPP, xi, ix, kb, niter, Tmax, y, x0;
	bool direct;

	__device__ __forceinline__ bool survivor(int s, int &k, uint32_t &w) const
	{
		k = direct ? kb + s : (int)list[s * P1_TX + xi];
		w = sv[k];
		return !direct || (uint32_t)(ix - (int)(w & 0xffu)) <= ((w >> 8) & 0xffu);
	}
	__device__ __forceinline__ int dist(uint32_t w) const { return abs((int)((w >> 16) & 0xffu) - ix); }
	// classes [0, classes(w, d)) take this survivor
	__device__ __forceinline__ int classes(uint32_t w, int d) const { return min((int)((w >> 24) & 0x3fu) + 1, (int)jmax[d] + 1); }
};

// General path for one class: sorted list of disjoint intervals (any number of components).
template <int CAP>
__device__ __noinline__ double2 class_general(const Pass1TileArgs &a, const TileThread &t, int j, unsigned long long slot)
{
	double2 ulist[CAP];
	RunUnion<CAP> u(ulist);
	for (int s = 0; s < t.niter; ++s) {
		int k;
		uint32_t w;
		if (!t.survivor(s, k, w)) continue;
		const int d = t.dist(w);
		if (j < t.classes(w, d)) {
			const double2 ab = t.cand[k];
			const double hh = t.Ht[(size_t)d * t.JPP + j];
			u.insert(ab.x - hh, ab.y + hh);
		}
	}
	if (u.overflow) { redo_push(a.redo, slot); return slot_empty(); }
	if (u.n == 0) return slot_empty();
	if (u.n == 1) return make_double2(u.s0, u.e0);
	const unsigned long long pb = atomicAdd(a.cursor, (unsigned long long)u.n);
	if (pb + u.n <= a.pool_cap)
		for (int q = 0; q < u.n; ++q) a.pool[pb + q] = u.L[q];
	return slot_pool(pb, (unsigned int)u.n);
}

// Classes CB at a time in registers, NL hulls per class. Hull l collects the survivors that are the l-th
// interval of their column (the last hull also takes any further ones): for shells, slabs and complements
// (erosion) each such layer unions to ONE interval, so a class costs NL (lo, hi) pairs and no list. A
// survivor that misses the running hull of its layer makes the class "complex" (general path).
template <int CB, int NL, int CAP>
__device__ __forceinline__ void eval_classes(const Pass1TileArgs &a, const TileThread &t)
{
	const double inf = __longlong_as_double(0x7FF0000000000000LL);
	const unsigned int full = (1u << CB) - 1u;
	for (int cb = 0; cb < t.Tmax; cb += CB) {
		double lo[NL][CB], hi[NL][CB];
		unsigned int seen[NL];
#pragma unroll
		for (int l = 0; l < NL; ++l) {
			seen[l] = 0;
#pragma unroll
			for (int q = 0; q < CB; ++q) { lo[l][q] = inf; hi[l][q] = -inf; }
		}
		unsigned int complex_mask = 0;
		for (int s = 0; s < t.niter; ++s) {
			int k;
			uint32_t w;
			if (!t.survivor(s, k, w)) continue;
			const int d = t.dist(w);
			const int te = t.classes(w, d) - cb;           // classes [cb, cb + te) take this survivor
			if (te <= 0) continue;
			const double2 ab = t.cand[k];
			const double2 *hp = reinterpret_cast<const double2 *>(t.Ht + (size_t)d * t.JPP + cb);
			double h[CB];
#pragma unroll
			for (int q = 0; q < CB / 2; ++q) { const double2 v = hp[q]; h[2 * q] = v.x; h[2 * q + 1] = v.y; }
			// branch-free: a class that does not take this survivor gets the cap -inf, i.e. the candidate
			// (+inf, -inf), which leaves its hull untouched
			const unsigned int valid = te >= CB ? full : ((1u << te) - 1u);
			const int lay = NL == 1 ? 0 : min((int)(w >> 30), NL - 1);
#pragma unroll
			for (int l = 0; l < NL; ++l) {
				if (NL == 1 || lay == l) {
					unsigned int miss = 0;
#pragma unroll
					for (int q = 0; q < CB; ++q) {
						const double hq = (q < te) ? h[q] : -inf;
						const double cs = ab.x - hq, ce = ab.y + hq;
						miss |= (cs <= hi[l][q] && ce >= lo[l][q]) ? 0u : (1u << q);
						lo[l][q] = cs < lo[l][q] ? cs : lo[l][q];
						hi[l][q] = ce > hi[l][q] ? ce : hi[l][q];
					}
					complex_mask |= miss & valid & seen[l];
					seen[l] |= valid;
				}
			}
		}
		// classes whose two hulls stay apart need two pool entries: one allocation for the whole block
		unsigned int two_mask = 0;
		if (NL == 2) {
#pragma unroll
			for (int q = 0; q < CB; ++q) {
				const bool both = lo[0][q] <= hi[0][q] && lo[NL - 1][q] <= hi[NL - 1][q];
				const bool apart = !(lo[NL - 1][q] <= hi[0][q] && hi[NL - 1][q] >= lo[0][q]);
				if (cb + q < t.Tmax && both && apart && !((complex_mask >> q) & 1u)) two_mask |= 1u << q;
			}
		}
		unsigned long long pb = 0;
		if (two_mask) pb = pool_alloc(a.cursor, 2u * __popc(two_mask));
#pragma unroll
		for (int q = 0; q < CB; ++q) {
			const int j = cb + q;
			if (j >= t.Tmax) break;
			const unsigned long long slot = ((unsigned long long)t.y * t.JP + j) * a.nx + t.x0 + t.xi;
			double2 out;
			if ((complex_mask >> q) & 1u) out = class_general<CAP>(a, t, j, slot);
			else if (NL == 1) out = make_double2(lo[0][q], hi[0][q]);          // (+inf, -inf) is the empty slot
			else if ((two_mask >> q) & 1u) {
				const bool first0 = lo[0][q] < lo[NL - 1][q];
				const double2 h0 = make_double2(lo[0][q], hi[0][q]), h1 = make_double2(lo[NL - 1][q], hi[NL - 1][q]);
				if (pb + 2 <= a.pool_cap) { a.pool[pb] = first0 ? h0 : h1; a.pool[pb + 1] = first0 ? h1 : h0; }
				out = slot_pool(pb, 2u);
				pb += 2;
			} else {
				// at most one interval: the hulls overlap, or one (or both) is empty ((+inf, -inf) drops out of min / max)
				const double l0 = lo[0][q] < lo[NL - 1][q] ? lo[0][q] : lo[NL - 1][q];
				const double h0 = hi[0][q] > hi[NL - 1][q] ? hi[0][q] : hi[NL - 1][q];
				out = make_double2(l0, h0);
			}
			a.mid[slot] = out;
		}
	}
}

template <int CAP, bool MULTI>
__device__ __forceinline__ void pass1_tile_body(const Pass1TileArgs &a, const unsigned int tile)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const int J = a.J, JP = J + 1, JPP = pass1_jpp(J), TX = P1_TX, SEG = TX + 2 * J;
	double2 *s_cand = reinterpret_cast<double2 *>(smem_raw);
	double *s_Ht = reinterpret_cast<double *>(s_cand + a.cmax);
	double *s_D = s_Ht + (size_t)JP * JPP;
	uint32_t *s_off = reinterpret_cast<uint32_t *>(s_D + JP + 1);
	uint32_t *s_sv = s_off + ((SEG + 2) & ~1);
	int *s_reach = reinterpret_cast<int *>(s_sv + a.cmax);
	uint16_t *s_list = reinterpret_cast<uint16_t *>(s_reach + ((JP + 1) & ~1));
	uint16_t *s_ty = s_list + (size_t)P1_LCAP * TX;
	uint8_t *s_jmax = reinterpret_cast<uint8_t *>(s_ty + a.cmax);

	const int tid = threadIdx.x, nthr = blockDim.x;
	const int y = (int)(tile / (unsigned)a.tiles_x);
	const int x0 = (int)(tile % (unsigned)a.tiles_x) * TX;
	const int txe = min(TX, a.nx - x0);
	const size_t rowbase = (size_t)y * a.nx, ncols_all = (size_t)a.nx * a.ny;

	// ---- phase 0: stage the row segment -------------------------------------------------------
	bool multi = false;                                     // does a column of the segment hold several intervals?
	for (int i = tid; i <= SEG; i += nthr) {
		const int gx = min(max(x0 - J + i, 0), a.nx);      // columns outside the grid collapse to empty ranges
		const uint32_t o = __ldg(a.off + rowbase + gx);
		s_off[i] = o;
		if (!MULTI && i < SEG) multi |= __ldg(a.off + rowbase + min(max(x0 - J + i + 1, 0), a.nx)) - o > 1u;
	}
	if (__syncthreads_or(multi)) {                          // (never taken by the MULTI variant)
		if (tid == 0) a.multi_tiles[atomicAdd(a.multi_count, 1u)] = tile;   // -> two-hull variant, launch 3
		return;
	}
	const uint32_t base = s_off[0];
	const int ncand = (int)(s_off[SEG] - base);
	if (ncand > a.cmax) {
		if (a.big_tiles) {                                  // run again with a larger candidate buffer
			if (tid == 0) a.big_tiles[atomicAdd(a.big_count, 1u)] = tile;
			return;
		}
		// no larger buffer: leave every slot of the tile to k_pass1
		if (tid < txe) { a.flags[rowbase + x0 + tid] = (uint8_t)JP; a.flags[ncols_all + rowbase + x0 + tid] = (uint8_t)JP; }
		for (int idx = tid; idx < JP * txe; idx += nthr) {
			const int j = idx / txe, xi = idx % txe;
			redo_push(a.redo, ((unsigned long long)y * JP + j) * a.nx + x0 + xi);
		}
		return;
	}
	if (ncand == 0) {                                       // nothing in reach: no slot of the tile is needed
		if (tid < txe) { a.flags[rowbase + x0 + tid] = 0; a.flags[ncols_all + rowbase + x0 + tid] = 0; }
		return;
	}
	for (int i = tid; i < JP * JPP; i += nthr) s_Ht[i] = __ldg(a.Ht + i);
	for (int i = tid; i < JP + 1; i += nthr) s_D[i] = __ldg(a.Dmono + i);
	for (int i = tid; i < JP; i += nthr) s_reach[i] = __ldg(a.reach + i);
	for (int k = tid; k < ncand; k += nthr) { s_cand[k] = __ldg(a.spans + base + k); s_ty[k] = __ldg(a.ty + base + k); }
	for (int i = tid; i < SEG; i += nthr)
		for (uint32_t k = s_off[i] - base; k < s_off[i + 1] - base; ++k) s_sv[k] = (uint32_t)i << 16;
	__syncthreads();
	// largest class whose reach covers distance d (reach is non-increasing in j)
	for (int d = tid; d < JP; d += nthr) {
		int jm = 0;
		for (int j = 0; j < JP; ++j) if (s_reach[j] >= d) jm = j;
		s_jmax[d] = (uint8_t)jm;
	}

	// ---- phase 1: x-dominance thresholds ----------------------------------------------------------
	for (int k = tid; k < ncand; k += nthr) {
		const int i = (int)(s_sv[k] >> 16);
		const double2 p = s_cand[k];
		const double m = 1e-9 + 1e-13 * (fabs(p.x) + fabs(p.y));
		double need_hi = __longlong_as_double(0x7FF0000000000000LL), need_lo = need_hi;
		if (i > 0)
			for (uint32_t q = s_off[i - 1] - base; q < s_off[i] - base; ++q)
				need_hi = fmin(need_hi, fmax(s_cand[q].x - p.x, p.y - s_cand[q].y));
		if (i < SEG - 1)
			for (uint32_t q = s_off[i + 1] - base; q < s_off[i + 2] - base; ++q)
				need_lo = fmin(need_lo, fmax(s_cand[q].x - p.x, p.y - s_cand[q].y));
		const int t_hi = first_ge(s_D, J, need_hi + m);   // dominated for outputs at distance >= t_hi on the left
		const int t_lo = first_ge(s_D, J, need_lo + m);   // ... on the right
		// survives for the outputs ix in [i - (t_hi-1), i + (t_lo-1)] (segment coordinates)
		const int first = max(i - (t_hi - 1), 0);
		const int last = min(i + (t_lo - 1), SEG - 1);
		const uint32_t ty = s_ty[k];
		const uint32_t layer = min((uint32_t)k - (s_off[i] - base), 3u);      // position of the interval inside its column
		s_sv[k] = (uint32_t)first | ((uint32_t)(last - first) << 8) | ((uint32_t)i << 16) | ((max(ty & 0xffu, ty >> 8) - 1u) << 24) | (layer << 30);
	}
	__syncthreads();
	if (tid >= txe) return;

	// ---- phase 2: one thread per output column ------------------------------------------------------
	const int xi = tid, ix = xi + J;
	const int kb = (int)(s_off[ix - J] - base), ke = (int)(s_off[ix + J + 1] - base);
	int S = 0, Fup = 0, Fdn = 0;
	uint32_t maxlayer = 0;
	for (int k = kb; k < ke; ++k) {
		const uint32_t w = s_sv[k];
		if ((uint32_t)(ix - (int)(w & 0xffu)) <= ((w >> 8) & 0xffu)) {
			maxlayer = max(maxlayer, w >> 30);
			const int d = abs((int)((w >> 16) & 0xffu) - ix);
			const int jm = (int)s_jmax[d] + 1;            // classes that can reach this distance
			const uint32_t ty = s_ty[k];
			Fup = max(Fup, min((int)(ty & 0xffu), jm));
			Fdn = max(Fdn, min((int)(ty >> 8), jm));
			if (S < P1_LCAP) s_list[S * TX + xi] = (uint16_t)k;
			++S;
		}
	}
	a.flags[rowbase + x0 + xi] = (uint8_t)Fup;
	a.flags[ncols_all + rowbase + x0 + xi] = (uint8_t)Fdn;
	const int Tmax = max(Fup, Fdn);
	// more survivors than the list holds (steep walls, many layers): re-scan the candidate range instead
	TileThread t;
	t.cand = s_cand; t.Ht = s_Ht; t.sv = s_sv; t.list = s_list; t.jmax = s_jmax;
	t.JP = JP; t.JPP = JPP; t.xi = xi; t.ix = ix; t.kb = kb; t.Tmax = Tmax; t.y = y; t.x0 = x0;
	t.direct = S > P1_LCAP;
	t.niter = t.direct ? ke - kb : S;
	if (!MULTI || maxlayer == 0) eval_classes<P1_CB, 1, CAP>(a, t);   // every survivor is the first interval of its column
	else eval_classes<P1_CB / 2, 2, CAP>(a, t);                       // two hulls per class, four classes at a time
}

// LIST = false: launch over all tiles, one tile per CTA. LIST = true: launch over a collected list; a fixed
// grid strides over it (the list length is only known on the device).
template <int CAP, bool MULTI, bool LIST>
__global__ void __launch_bounds__(P1_TX, MULTI ? 5 : 6) k_pass1_tile(Pass1TileArgs a)
{
	if (!LIST) pass1_tile_body<CAP, MULTI>(a, a.tile0 + blockIdx.x);
	else {
		// one resident wave of CTAs pulls tiles from the list (tile costs vary a lot: dynamic beats strided)
		__shared__ unsigned int s_next;
		const unsigned int n = *a.tiles_count;
		for (;;) {
			if (threadIdx.x == 0) s_next = atomicAdd(a.tiles_next, 1u);
			__syncthreads();
			const unsigned int i = s_next;
			if (i >= n) break;
			pass1_tile_body<CAP, MULTI>(a, a.tiles[i]);
			__syncthreads();                                // the next tile reuses the shared buffers (and s_next)
		}
	}
}

} // namespace vo
