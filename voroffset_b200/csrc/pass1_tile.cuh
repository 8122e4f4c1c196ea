// 'ours', pass 1, tile kernel with exact two-sided 2D dominance pruning.
//
// What it computes is what k_pass1 (kernels.cuh) computes - for an output column x of row y and a radius
// class j the union over |dx| <= reach[j] of the neighbours' intervals grown by H[j][|dx|] - but only for
// the (candidate, consumer) pairs that can still matter, and only for the classes some consumer of pass 2
// will read.
//
// A PAIR is (input interval p of column (cx, y), consumer column (x, yc)): distance d = |cx - x| in x,
// class j = |y - yc| in y, contribution [a_p - H[j][d], b_p + H[j][d]]. A pair may be dropped when another
// legitimate pair of the same consumer contains its contribution; we test the four lattice neighbours of p
// and call nn(p, N) = min over intervals q of the neighbouring column N of max(a_q - a_p, b_p - b_q)
// ("how far the best q is from containing p"; negative: q contains p with room to spare):
//
//   near, x : N one column CLOSER to the consumer. Dropped for every class at distances >= t where
//             nn + m <= Dmono[t], Dmono[d] = min over d' >= d and live classes of H[j][d'-1] - H[j][d'].
//   far, x  : N one column FARTHER. Dropped at distance d for the classes N still reaches (j <= jmax[d+1])
//             when -nn - m >= G[T-1][d], G[jt][d] = max over d' <= d, j <= min(jt, jmax[d'+1]) of
//             H[j][d'] - H[j][d'+1]; the classes in (jmax[d+1], jmax[d]] keep the pair.
//   near, y : N one row closer (class j-1, same d). Dropped for the consumers at row distance >= Ty where
//             nn + m <= Emono[Ty], Emono[j] = min over j' >= j, d <= reach[j'] of H[j'-1][d] - H[j'][d].
//   far, y  : N one row farther (class j+1, same d). Dropped for the classes j < Tf(d), Tf(d) = number of
//             leading classes with -nn - m >= Ef[d][j+1], Ef[d][c] = max over c' <= c of H[c'-1][d] - H[c'][d]
//             (+inf beyond the reach of distance d).
//
// So a surviving pair takes part in two class windows, [lo_u, hi_u) for the consumers above and
// [lo_d, hi_d) for those below; class 0 (the consumer in the same row) needs both. Why this is exact:
// every test is decided on the table values the contributions are built from and fp64 subtraction is
// monotone, so containment carries over to the rounded endpoints; the margin m makes every dominator's
// lower endpoint STRICTLY smaller, so chains of dominators cannot cycle and end at a kept pair; the far
// bounds are rounded up (float), the near bounds are exact doubles with the margin - conservative pruning
// never changes the union. This is the role the reference's Voronoi-vertex / power-diagram events play
// (Voronoi2D.cpp:329-586, SeparatePower2D.cpp:118-293: seeds are retired once their cell no longer reaches
// the sweep line) in a data-parallel form. On the C5 torus ~2.6 of ~38 candidates per column survive, in
// ~3.4 classes each (near tests alone: 7.6 candidates x 15 classes).
//
// Kernel structure (CTA = one row segment of P1_TX output columns):
//   phase 0  the segment [x0-J, x0+TX+J) of row y is staged in shared memory as a flat candidate array;
//   phase 1  thread per candidate: the four nn values (columns cx-1 / cx+1 from shared memory, rows y-1 / y+1
//            from global memory) -> near / far thresholds, packed per candidate;
//   phase 2  thread per output column: scan of its candidate range -> short survivor list with the class
//            windows -> per-column window hulls (published as flags: pass 2 reads only those slots, pass 1
//            writes only those) -> classes evaluated CB at a time in registers ((lo, hi) hulls per layer;
//            a class whose union is not one interval per layer is redone by the general list path).
// Lanes run along x: slot writes are full lines.
#pragma once
#include <cooperative_groups.h>
#include <cooperative_groups/scan.h>

#include "kernels.cuh"

namespace vo {

constexpr int P1_TX = 128;      // output columns (= threads) per CTA
constexpr int P1_CB = 4;        // classes evaluated together in registers
constexpr int P1_LCAP_S = 12;   // survivors listed per output column, single-interval launch
constexpr int P1_LCAP_M = 32;   // ... multi-interval / large-buffer launches (more: the range is re-scanned)

struct Pass1TileArgs {
	int nx, ny, J, cmax, tiles_x;
	unsigned int tile0;     // first tile of a launch over all tiles of a band of rows
	const uint32_t *off;
	const double2 *spans;
	const double *Ht;       // (J+1) rows of JPP doubles: Ht[d*JPP + j] = H[j][d] (-1 beyond the reach / the table)
	const double *Dmono;    // J+2
	const double *Emono;    // J+2
	const float *G;         // (J+1)*(J+1): G[jt*(J+1) + d], rounded up, G[.][J] = +inf
	const float *Ef;        // (J+1)*(J+2): Ef[d*(J+2) + c], rounded up, +inf for c > jmax[d]
	const uint8_t *jmax;    // J+2: largest class whose reach covers distance d (jmax[J+1] = 0, never used for a live pair)
	double2 *mid;
	uint16_t *flags;        // [2][ny*nx]: lo | hi << 8 of the class window needed by the consumers above ([0]) / below ([1])
	double2 *pool;
	unsigned long long *cursor;
	unsigned long long pool_cap;
	Redo redo;              // slot ids to be (re)done by k_pass1: list overflow and oversized tiles
	// Device-side dispatch (no host round trip between the launches):
	//   launch 1  <MULTI=false>, small candidate buffer, all tiles. Tiles with a multi-interval column go to
	//             multi_tiles, tiles with more than cmax candidates to big_tiles.
	//   launch 2  <MULTI=false>, large buffer, tiles = big_tiles.
	//   launch 3  <MULTI=true> (two hulls per class), tiles = multi_tiles.
	// A launch over a list uses one resident wave of CTAs pulling tiles with an atomic counter.
	const unsigned int *tiles;        // NULL: all tiles
	const unsigned int *tiles_count;
	unsigned int *tiles_next;         // list launches: next list position to hand out
	unsigned int *big_tiles;          // NULL: oversized tiles go to the redo list (k_pass1)
	unsigned int *big_count;
	unsigned int *multi_tiles;
	unsigned int *multi_count;
};

__host__ __device__ inline int pass1_jpp(int J) { return (J + 1 + P1_CB + 1) & ~1; }

__host__ __device__ inline size_t pass1_tile_smem(int J, int cmax, int lcap)
{
	const size_t JP = (size_t)J + 1, SEG = (size_t)P1_TX + 2 * J;
	size_t b = 0;
	b += (size_t)cmax * sizeof(double2);                    // candidates
	b += JP * pass1_jpp(J) * sizeof(double);                // Ht
	b += 2 * (JP + 1) * sizeof(double);                     // Dmono, Emono
	b += ((JP * JP + 1) & ~(size_t)1) * sizeof(float);      // G
	b += ((JP * (JP + 1) + 1) & ~(size_t)1) * sizeof(float);// Ef
	b += 2 * (size_t)cmax * sizeof(float);                  // far values (up / down consumers)
	b += 2 * (size_t)cmax * sizeof(uint32_t);               // packed thresholds
	b += ((SEG + 2) & ~(size_t)1) * sizeof(uint32_t);       // segment offsets
	b += 2 * (size_t)lcap * P1_TX * sizeof(uint32_t);       // survivor lists [s][thread]: candidate word, window word
	b += (JP + 1 + 3) & ~(size_t)3;                         // jmax
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

__device__ __forceinline__ int first_ge(const double *tab, int J, double need)
{
	int lo = 1, hi = J + 1;                       // tab is non-decreasing, tab[J+1] = +inf
	while (lo < hi) { const int mid = (lo + hi) >> 1; if (tab[mid] >= need) hi = mid; else lo = mid + 1; }
	return lo;
}

// first index in [1, last] whose entry exceeds v; tab is non-decreasing and tab[last] = +inf
__device__ __forceinline__ int first_gt(const float *tab, int last, float v)
{
	if (!(tab[1] <= v)) return 1;                 // the common case: no far dominance at all
	int lo = 2, hi = last;
	while (lo < hi) { const int mid = (lo + hi) >> 1; if (tab[mid] > v) hi = mid; else lo = mid + 1; }
	return lo;
}

// nn(p, column): min over the intervals q of the column of max(a_q - a_p, b_p - b_q)
__device__ __forceinline__ double nn_of(const double2 p, const double2 *q, uint32_t q0, uint32_t q1)
{
	double r = __longlong_as_double(0x7FF0000000000000LL);
	for (uint32_t k = q0; k < q1; ++k) { const double2 v = q[k]; r = fmin(r, fmax(v.x - p.x, p.y - v.y)); }
	return r;
}

// Per-thread view of the staged tile (phase 2).
template <int LCAP>
struct TileThread {
	const double2 *cand;
	const double *Ht;
	const float *Ef;
	const float *vu, *vd;      // far values per candidate (consumers above / below)
	const uint32_t *wa;        // first | width << 8 | column << 16 | layer << 24
	const uint32_t *wb;        // tfL | tfR << 8 | (Ty_up - 1) << 16 | (Ty_dn - 1) << 24
	uint32_t *listA, *listB;   // [s * P1_TX + xi]
	const uint8_t *jmax;
	int J, JPP, xi, ix, kb, niter, y, x0;
	bool direct;

	// Is candidate k a surviving pair for this output, before the (d-dependent) far test in y? On success
	// `e` = k | d << 16 | base << 24 (base: first class the x-far test leaves to this pair).
	__device__ __forceinline__ bool near_alive(int k, uint32_t &e) const
	{
		const uint32_t w = wa[k];
		if ((uint32_t)(ix - (int)(w & 0xffu)) > ((w >> 8) & 0xffu)) return false;
		const int i = (int)((w >> 16) & 0xffu), d = abs(i - ix);
		const uint32_t b = wb[k];
		const int tf = ix < i ? (int)(b & 0xffu) : (int)((b >> 8) & 0xffu);
		const int T = (int)max((b >> 16) & 0xffu, b >> 24) + 1;
		const int base = (d >= 1 && d < tf) ? (int)jmax[d + 1] + 1 : 0;
		if (base >= min(T, (int)jmax[d] + 1)) return false;
		e = (uint32_t)k | ((uint32_t)d << 16) | ((uint32_t)base << 24);
		return true;
	}
	// Class windows of a near-alive pair: lo_u | hi_u << 8 | lo_d << 16 | hi_d << 24 (0 = both empty).
	__device__ __forceinline__ uint32_t windows(uint32_t e) const
	{
		const int k = (int)(e & 0xffffu), d = (int)((e >> 16) & 0xffu), base = (int)(e >> 24);
		const uint32_t b = wb[k];
		const int jm1 = (int)jmax[d] + 1;
		int hu = min((int)((b >> 16) & 0xffu) + 1, jm1), hd = min((int)(b >> 24) + 1, jm1);
		const float *row = Ef + (size_t)d * (J + 2);
		int lu = max(base, first_gt(row, J + 1, vu[k]) - 1);
		int ld = max(base, first_gt(row, J + 1, vd[k]) - 1);
		if (lu > 0 || ld > 0) { lu = max(lu, 1); ld = max(ld, 1); }      // class 0 needs both windows
		if (lu >= hu) lu = hu = 0;
		if (ld >= hd) ld = hd = 0;
		return (uint32_t)lu | ((uint32_t)hu << 8) | ((uint32_t)ld << 16) | ((uint32_t)hd << 24);
	}
	// s-th entry of the survivor loop: candidate word and window word (0: not a survivor).
	__device__ __forceinline__ bool survivor(int s, uint32_t &e, uint32_t &w) const
	{
		if (!direct) { e = listA[s * P1_TX + xi]; w = listB[s * P1_TX + xi]; return w != 0; }
		if (!near_alive(kb + s, e)) return false;
		w = windows(e);
		return w != 0;
	}
};

// bits q in [0, CB) with l <= cb + q < h
template <int CB>
__device__ __forceinline__ unsigned int range_mask(int l, int h, int cb)
{
	const int a = min(max(l - cb, 0), CB), b = min(max(h - cb, 0), CB);
	return ((1u << b) - 1u) & ~((1u << a) - 1u);
}
template <int CB>
__device__ __forceinline__ unsigned int window_mask(uint32_t w, int cb)
{
	return range_mask<CB>((int)(w & 0xffu), (int)((w >> 8) & 0xffu), cb) | range_mask<CB>((int)((w >> 16) & 0xffu), (int)(w >> 24), cb);
}

// General path for one class: sorted list of disjoint intervals (any number of components).
template <int CAP, int LCAP>
__device__ __noinline__ double2 class_general(const Pass1TileArgs &a, const TileThread<LCAP> &t, int j, unsigned long long slot)
{
	double2 ulist[CAP];
	RunUnion<CAP> u(ulist);
	for (int s = 0; s < t.niter; ++s) {
		uint32_t e, w;
		if (!t.survivor(s, e, w)) continue;
		if (window_mask<1>(w, j)) {
			const double2 ab = t.cand[e & 0xffffu];
			const double hh = t.Ht[(size_t)((e >> 16) & 0xffu) * t.JPP + j];
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

// Classes [cb, cb + CB) in registers, NL hulls per class. Hull l collects the survivors that are the l-th
// interval of their column (the last hull also takes any further ones): for shells, slabs and complements
// (erosion) each such layer unions to ONE interval, so a class costs NL (lo, hi) pairs and no list. A
// survivor that misses the running hull of its layer makes the class "complex" (general path).
// `need` = the classes of the block some consumer reads (bit q = class cb + q).
template <int CB, int NL, int CAP, int LCAP>
__device__ __forceinline__ void eval_block(const Pass1TileArgs &a, const TileThread<LCAP> &t, int cb, unsigned int need)
{
	const double inf = __longlong_as_double(0x7FF0000000000000LL);
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
		uint32_t e, w;
		if (!t.survivor(s, e, w)) continue;
		const unsigned int valid = window_mask<CB>(w, cb);
		if (!valid) continue;
		const int k = (int)(e & 0xffffu), d = (int)((e >> 16) & 0xffu);
		const double2 ab = t.cand[k];
		const double *hp = t.Ht + (size_t)d * t.JPP + cb;
		double h[CB];
#pragma unroll
		for (int q = 0; q < CB; ++q) h[q] = hp[q];
		// branch-free: a class that does not take this survivor gets the cap -inf, i.e. the candidate
		// (+inf, -inf), which leaves its hull untouched
		const int lay = NL == 1 ? 0 : min((int)((t.wa[k] >> 24) & 3u), NL - 1);
#pragma unroll
		for (int l = 0; l < NL; ++l) {
			if (NL == 1 || lay == l) {
				unsigned int miss = 0;
#pragma unroll
				for (int q = 0; q < CB; ++q) {
					const double hq = ((valid >> q) & 1u) ? h[q] : -inf;
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
			if (((need >> q) & 1u) && both && apart && !((complex_mask >> q) & 1u)) two_mask |= 1u << q;
		}
	}
	unsigned long long pb = 0;
	if (two_mask) pb = pool_alloc(a.cursor, 2u * __popc(two_mask));
#pragma unroll
	for (int q = 0; q < CB; ++q) {
		if (!((need >> q) & 1u)) continue;
		const int j = cb + q;
		const unsigned long long slot = ((unsigned long long)t.y * (t.J + 1) + j) * a.nx + t.x0 + t.xi;
		double2 out;
		if ((complex_mask >> q) & 1u) out = class_general<CAP, LCAP>(a, t, j, slot);
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

// every class of the two windows [UL, UH) u [DL, DH), CB at a time, blocks starting at a needed class
template <int CB, int NL, int CAP, int LCAP>
__device__ __forceinline__ void eval_classes(const Pass1TileArgs &a, const TileThread<LCAP> &t, int UL, int UH, int DL, int DH)
{
	const int jend = max(UH, DH);
	int cb = UH > UL ? (DH > DL ? min(UL, DL) : UL) : DL;
	while (cb < jend) {
		const unsigned int need = range_mask<CB>(UL, UH, cb) | range_mask<CB>(DL, DH, cb);
		eval_block<CB, NL, CAP, LCAP>(a, t, cb, need);
		cb += CB;
		// next needed class at or after cb
		const bool in_u = cb >= UL && cb < UH, in_d = cb >= DL && cb < DH;
		if (!in_u && !in_d) {
			int nxt = jend;
			if (UH > UL && UL >= cb) nxt = min(nxt, UL);
			if (DH > DL && DL >= cb) nxt = min(nxt, DL);
			cb = nxt;
		}
	}
}

template <int CAP, bool MULTI, int LCAP>
__device__ __forceinline__ void pass1_tile_body(const Pass1TileArgs &a, const unsigned int tile)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const int J = a.J, JP = J + 1, JPP = pass1_jpp(J), TX = P1_TX, SEG = TX + 2 * J;
	double2 *s_cand = reinterpret_cast<double2 *>(smem_raw);
	double *s_Ht = reinterpret_cast<double *>(s_cand + a.cmax);
	double *s_D = s_Ht + (size_t)JP * JPP;
	double *s_E = s_D + JP + 1;
	float *s_G = reinterpret_cast<float *>(s_E + JP + 1);
	float *s_Ef = s_G + ((JP * JP + 1) & ~1);
	float *s_vu = s_Ef + ((JP * (JP + 1) + 1) & ~1);
	float *s_vd = s_vu + a.cmax;
	uint32_t *s_wa = reinterpret_cast<uint32_t *>(s_vd + a.cmax);
	uint32_t *s_wb = s_wa + a.cmax;
	uint32_t *s_off = s_wb + a.cmax;
	uint32_t *s_listA = s_off + ((SEG + 2) & ~1);
	uint32_t *s_listB = s_listA + (size_t)LCAP * TX;
	uint8_t *s_jmax = reinterpret_cast<uint8_t *>(s_listB + (size_t)LCAP * TX);

	const int tid = threadIdx.x, nthr = blockDim.x;
	const int y = (int)(tile / (unsigned)a.tiles_x);
	const int x0 = (int)(tile % (unsigned)a.tiles_x) * TX;
	const int txe = min(TX, a.nx - x0);
	const size_t rowbase = (size_t)y * a.nx, ncols_all = (size_t)a.nx * a.ny;
	const uint16_t full = (uint16_t)(JP << 8);             // window [0, J+1)

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
		if (tid < txe) { a.flags[rowbase + x0 + tid] = full; a.flags[ncols_all + rowbase + x0 + tid] = full; }
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
	for (int i = tid; i < JP + 1; i += nthr) { s_D[i] = __ldg(a.Dmono + i); s_E[i] = __ldg(a.Emono + i); }
	for (int i = tid; i < JP * JP; i += nthr) s_G[i] = __ldg(a.G + i);
	for (int i = tid; i < JP * (JP + 1); i += nthr) s_Ef[i] = __ldg(a.Ef + i);
	for (int i = tid; i < JP + 1; i += nthr) s_jmax[i] = __ldg(a.jmax + i);
	for (int k = tid; k < ncand; k += nthr) s_cand[k] = __ldg(a.spans + base + k);
	for (int i = tid; i < SEG; i += nthr)
		for (uint32_t k = s_off[i] - base; k < s_off[i + 1] - base; ++k) s_wa[k] = (uint32_t)i << 16;
	__syncthreads();

	// ---- phase 1: dominance thresholds, thread per candidate -------------------------------------------
	for (int k = tid; k < ncand; k += nthr) {
		const int i = (int)(s_wa[k] >> 16);
		const double2 p = s_cand[k];
		const double inf = __longlong_as_double(0x7FF0000000000000LL);
		const double m = 1e-9 + 1e-13 * (fabs(p.x) + fabs(p.y));
		double nnL = inf, nnR = inf, nnU = inf, nnD = inf;
		if (i > 0) nnL = nn_of(p, s_cand, s_off[i - 1] - base, s_off[i] - base);
		if (i < SEG - 1) nnR = nn_of(p, s_cand, s_off[i + 1] - base, s_off[i + 2] - base);
		const size_t c = rowbase + (size_t)(x0 - J + i);                // the candidate's own column (inside the grid)
		if (y > 0) nnU = nn_of(p, a.spans, __ldg(a.off + c - a.nx), __ldg(a.off + c - a.nx + 1));
		if (y < a.ny - 1) nnD = nn_of(p, a.spans, __ldg(a.off + c + a.nx), __ldg(a.off + c + a.nx + 1));
		// near: dominated for outputs at distance >= t (x) / consumers at row distance >= Ty (y)
		const int tnL = first_ge(s_D, J, nnL + m), tnR = first_ge(s_D, J, nnR + m);
		const int tyu = first_ge(s_E, J, nnU + m), tyd = first_ge(s_E, J, nnD + m);
		// far, x: dominated for the outputs on the left at distances [1, tfL) by the right neighbour, ...
		const float *g = s_G + (size_t)(max(tyu, tyd) - 1) * JP;
		const int tfL = first_gt(g, J, __double2float_rd(-nnR - m));
		const int tfR = first_gt(g, J, __double2float_rd(-nnL - m));
		// survives the near tests for the outputs ix in [i - (tnL-1), i + (tnR-1)] (segment coordinates)
		const int first = max(i - (tnL - 1), 0);
		const int last = min(i + (tnR - 1), SEG - 1);
		const uint32_t layer = min((uint32_t)k - (s_off[i] - base), 3u);     // position of the interval inside its column
		s_wa[k] = (uint32_t)first | ((uint32_t)(last - first) << 8) | ((uint32_t)i << 16) | (layer << 24);
		s_wb[k] = (uint32_t)tfL | ((uint32_t)tfR << 8) | ((uint32_t)(tyu - 1) << 16) | ((uint32_t)(tyd - 1) << 24);
		// far, y: the consumers above are served by row y+1 instead, those below by row y-1
		s_vu[k] = __double2float_rd(-nnD - m);
		s_vd[k] = __double2float_rd(-nnU - m);
	}
	__syncthreads();
	if (tid >= txe) return;

	// ---- phase 2: one thread per output column ------------------------------------------------------
	const int xi = tid, ix = xi + J;
	TileThread<LCAP> t;
	t.cand = s_cand; t.Ht = s_Ht; t.Ef = s_Ef; t.vu = s_vu; t.vd = s_vd; t.wa = s_wa; t.wb = s_wb;
	t.listA = s_listA; t.listB = s_listB; t.jmax = s_jmax;
	t.J = J; t.JPP = JPP; t.xi = xi; t.ix = ix; t.y = y; t.x0 = x0;
	const int kb = (int)(s_off[ix - J] - base), ke = (int)(s_off[ix + J + 1] - base);
	t.kb = kb;
	int S = 0;
	for (int k = kb; k < ke; ++k) {
		uint32_t e;
		if (t.near_alive(k, e)) {
			if (S < LCAP) s_listA[S * TX + xi] = e;
			++S;
		}
	}
	// more survivors than the list holds (steep walls, many layers): re-scan the candidate range instead
	t.direct = S > LCAP;
	t.niter = t.direct ? ke - kb : S;
	int UL = 255, UH = 0, DL = 255, DH = 0;
	uint32_t maxlayer = 0;
	for (int s = 0; s < t.niter; ++s) {
		uint32_t e, w;
		if (t.direct) { if (!t.near_alive(kb + s, e)) continue; }
		else e = s_listA[s * TX + xi];
		w = t.windows(e);
		if (!t.direct) s_listB[s * TX + xi] = w;
		if (w == 0) continue;
		if (MULTI) maxlayer = max(maxlayer, (s_wa[e & 0xffffu] >> 24) & 3u);
		const int lu = (int)(w & 0xffu), hu = (int)((w >> 8) & 0xffu), ld = (int)((w >> 16) & 0xffu), hd = (int)(w >> 24);
		if (hu > lu) { UL = min(UL, lu); UH = max(UH, hu); }
		if (hd > ld) { DL = min(DL, ld); DH = max(DH, hd); }
	}
	if (UH == 0) UL = 0;
	if (DH == 0) DL = 0;
	a.flags[rowbase + x0 + xi] = (uint16_t)(UL | (UH << 8));
	a.flags[ncols_all + rowbase + x0 + xi] = (uint16_t)(DL | (DH << 8));
	if (UH == 0 && DH == 0) return;
	if (!MULTI || maxlayer == 0) eval_classes<P1_CB, 1, CAP, LCAP>(a, t, UL, UH, DL, DH);   // every survivor is the first interval of its column
	else eval_classes<P1_CB, 2, CAP, LCAP>(a, t, UL, UH, DL, DH);                           // two hulls per class
}

// LIST = false: launch over all tiles, one tile per CTA. LIST = true: launch over a collected list; a fixed
// grid strides over it (the list length is only known on the device).
template <int CAP, bool MULTI, bool LIST>
__global__ void __launch_bounds__(P1_TX, MULTI ? 4 : 5) k_pass1_tile(Pass1TileArgs a)
{
	constexpr int LCAP = (MULTI || LIST) ? P1_LCAP_M : P1_LCAP_S;
	if (!LIST) pass1_tile_body<CAP, MULTI, LCAP>(a, a.tile0 + blockIdx.x);
	else {
		// one resident wave of CTAs pulls tiles from the list (tile costs vary a lot: dynamic beats strided)
		__shared__ unsigned int s_next;
		const unsigned int n = *a.tiles_count;
		for (;;) {
			if (threadIdx.x == 0) s_next = atomicAdd(a.tiles_next, 1u);
			__syncthreads();
			const unsigned int i = s_next;
			if (i >= n) break;
			pass1_tile_body<CAP, MULTI, LCAP>(a, a.tiles[i]);
			__syncthreads();                                // the next tile reuses the shared buffers (and s_next)
		}
	}
}

// flags of the one-thread-per-slot kernel: every class of every column is computed and needed
__global__ void __launch_bounds__(256) k_fill16(uint16_t *p, unsigned long long n, uint16_t v)
{
	const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) p[i] = v;
}

} // namespace vo
