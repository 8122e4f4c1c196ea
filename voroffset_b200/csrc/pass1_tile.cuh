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
// Kernel structure:
//   k_thresh      thread per input column: the four nn values of each of its intervals -> near / far thresholds,
//                 16 bytes per interval (tile independent, so halo candidates are not recomputed per tile);
//   k_pass1_tile  one CTA per SM, cap tables staged in shared memory once; every warp pulls tiles (row segments of
//                 P1_W = 32 output columns) on its own, without CTA barriers. Per tile:
//     phase 0  the segment [x0-J, x0+TX+J) of row y is staged as a flat candidate array (interval + thresholds);
//     phase 1  thread per candidate: walks the output columns it survives for, computes the class windows of
//              each surviving pair (the far test in y depends on the distance) and appends the pair to that
//              output's list in shared memory;
//     phase 2  thread per output column: window hulls of its list (published as flags: pass 2 reads only
//              those slots, pass 1 writes only those) -> classes evaluated CB at a time in registers ((lo, hi)
//              hulls per layer; a class whose union is not one interval per layer is redone by the general
//              list path). Lists that overflow are replaced by a re-scan of the candidate range.
// Lanes run along x: slot writes are full lines.
#pragma once
#include <cooperative_groups.h>
#include <cooperative_groups/scan.h>

#include "kernels.cuh"

namespace vo {

constexpr int P1_W = 32;        // output columns of a tile = lanes of the warp that owns it
constexpr int P1_TX = 128;      // columns per tile mask (the unit pass 2 skips producer rows by)
constexpr int P1_CB = 4;        // classes evaluated together in registers
constexpr int P1_LCAP_S = 24;   // survivors listed per output column, single-interval launch
constexpr int P1_LCAP_M = 32;   // ... multi-interval / large-buffer launches (more: the range is re-scanned)
#ifndef P1_MAXWARPS_V
#define P1_MAXWARPS_V 16
#endif
constexpr int P1_MAXWARPS = P1_MAXWARPS_V; // warps per CTA (one CTA per SM; fewer when the per-warp buffers are large). (20 warps at 96
                                // registers were tried: the spills cost more than the extra warps give, 0.71 vs 0.65 ms on C5.)
#ifndef P1_MAXWARPS_M_V
#define P1_MAXWARPS_M_V 16
#endif
constexpr int P1_MAXWARPS_M = P1_MAXWARPS_M_V; // ... of the two-hull variant
#ifndef P1_MAXWARPS_LEAN_V
#define P1_MAXWARPS_LEAN_V 20
#endif
constexpr int P1_MAXWARPS_LEAN = P1_MAXWARPS_LEAN_V; // ... of the first launch without the inline sorted-list union (k_pass1_tile<..., GEN = false>: 88
                                // registers; 22 / 23 warps at 80 registers were no faster than 20, profiles/r2ci_defer_general_ab.txt)
constexpr int P1_OVF = 224;     // further survivors per output column kept in a global-memory spill area of the warp

__device__ __forceinline__ int first_ge(const double *tab, int J, double need)
{
	if (tab[1] >= need) return 1;                 // the two common cases first: contained in the neighbour outright,
	if (!(tab[J] >= need)) return J + 1;          // ... or no neighbour / nowhere near it
	int lo = 2, hi = J;                           // tab is non-decreasing, tab[J+1] = +inf
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

// nn(p, column): min over the intervals q of the column of max(a_q - a_p, b_p - b_q).
// Clipped form (erosion: the dilated complement is only read inside (clo, chi), Voronoi.cpp:57-89 drops the rest):
// an endpoint at or beyond the clip bound is SATURATED - whatever cap it gets, it stays outside the range - so
// on that side a saturated q contains everything and nothing unsaturated contains a saturated p. Without this
// the complement's intervals, which all share the endpoint zmin-1 or zmax+1, could never dominate each other
// from farther away (the closer one always reaches lower), and every candidate up to the tangent point survived.
// The intervals of a column are ascending and disjoint, so along q the first term (a_q - a_p) grows and the second
// (b_p - b_q) falls: the minimum sits where they cross. `qs` (kept by the caller across the ascending p of one column,
// initially q0) skips the q that lie entirely below p except the nearest, and the scan stops once the first term alone
// reaches the best value - a column of k intervals costs O(k) per neighbour instead of O(k^2) (lattices: 10 x 10).
// Any SUBSET of q gives a valid, merely more conservative value, so unsorted input only prunes less.
// dual (ThreshArgs::dual): p and the q are mirrored intervals (-z1, -z2) - the formula is the same, only the
// "is an interval" test and the scan start, which rely on z1 <= z2 and on the order inside a column, are dropped
// (dual columns hold one interval).
__device__ __forceinline__ double nn_of(const double2 p, const double2 *q, uint32_t &qs, uint32_t q1, double clo, double chi, bool dual = false)
{
	const double inf = __longlong_as_double(0x7FF0000000000000LL);
	const bool plo = p.x <= clo, phi = p.y >= chi;
	double r = inf;
	while (!dual && qs + 1 < q1 && __ldg(q + qs + 1).y < p.x) ++qs;
	for (uint32_t k = qs; k < q1; ++k) {
		double2 v = __ldg(q + k);
		if (dual) v = make_double2(-v.x, -v.y);
		else if (!(v.x <= v.y)) continue;                  // (not an interval, see k_thresh)
		const double ta = v.x <= clo ? -inf : (plo ? inf : v.x - p.x);
		const double tb = v.y >= chi ? -inf : (phi ? inf : p.y - v.y);
		r = fmin(r, fmax(ta, tb));
		if (ta >= r) break;
	}
	return r;
}

// margin by which the farther neighbour contains p (rounded down; finite, so that the +inf sentinels of the far
// tables always stop a search; -inf: never dominated from farther away)
__device__ __forceinline__ float far_value(double nn, double m, bool never)
{
	if (never) return __int_as_float(0xff800000);
	return fminf(__double2float_rd(-nn - m), 3.0e38f);
}

// ---- dominance thresholds, one thread per input column ------------------------------------------------
// thr[k] = { tnL | tnR << 8 | tfL << 16 | tfR << 24,
//            (Ty_up - 1) | (Ty_dn - 1) << 8 | layer << 16 | reach[T-1] << 24,  float bits of v_up,  of v_dn }
//   tnL / tnR : near, x: dropped for the outputs on the left / right at distances >= tn       (1 .. J+1)
//   tfL / tfR : far, x: dropped for the outputs on the left / right at distances in [1, tf)   (1 .. J), except
//               for the classes beyond the farther neighbour's reach (only at distances >= reach[T-1])
//   Ty_up/dn  : near, y: takes part in the classes j < Ty for the consumers above / below, T = max of both
//   v_up/dn   : far, y: -nn - m of the row FARTHER from the consumers above / below (rounded down)
//   layer     : position of the interval inside its column (3 = third or later)
struct ThreshArgs {
	int nx, ny, J;
	const uint32_t *off;
	const double2 *spans;
	const double *Dmono;    // J+2
	const double *Emono;    // J+2
	const float *G;         // (J+1)*(J+1)
	const int *reach;       // J+1
	uint4 *thr;             // per interval
	unsigned long long c_begin, c_end;   // columns processed by this launch (a band of rows, or everything)
	double clip_lo, clip_hi;             // the result is only read inside (clip_lo, clip_hi); -inf / +inf: everywhere
	unsigned int *est = nullptr;         // optional, [ny * tiles_xw], zeroed by the host: estimated cost of every pass-1 tile
	int tiles_xw = 0;                    //   (surviving pairs weighted by their classes; only the ORDER of the tiles uses it)
	unsigned long long *zero_bank = nullptr;   // optional: the tile lists and cursors ([3] [5] [6] [7] [10]) of the launch set that
	                                           //   follows on this stream are zeroed here instead of by three memsets of their own
	// optional (the banded host-buffer call, whose input nobody has validated yet): offsets that are not
	// non-decreasing or point beyond `nspans` raise *bad and are read as empty columns; the tile kernel of the band
	// then does nothing and the host reports VO_ERR_ARG
	unsigned int *bad = nullptr;
	uint32_t nspans = 0, nspans_lo = 0;                  // valid offsets: [nspans_lo, nspans] (a row window of a larger CSR starts above 0)
	// Dual form (erosion of a volume with at most one interval per column, vo_lib.cu: erode_dual): the thresholds of the
	// MIRRORED intervals (-z1, -z2), whose "dilation" hull is the erosion's intersection. Columns that do not qualify -
	// several intervals, an interval that is not strictly inside (dual_lo, dual_hi) or has z1 > z2 - raise *dual_bad.
	int dual = 0;
	double dual_lo = 0, dual_hi = 0;
	unsigned int *dual_bad = nullptr;
};

// offsets of one column as read from untrusted input (ThreshArgs::bad): an invalid pair becomes an empty column
__device__ __forceinline__ void thresh_guard(const ThreshArgs &a, uint32_t &q0, uint32_t &q1)
{
	if (a.bad && (q1 < q0 || q1 > a.nspans || q0 < a.nspans_lo)) { *a.bad = 1u; q0 = q1 = a.nspans_lo; }
}

__device__ __forceinline__ void thresh_zero_bank(const ThreshArgs &a)
{
	if (a.zero_bank && blockIdx.x == 0 && threadIdx.x == 0) {
		a.zero_bank[3] = 0; a.zero_bank[5] = 0; a.zero_bank[6] = 0; a.zero_bank[7] = 0; a.zero_bank[10] = 0;
		a.zero_bank[12] = 0; a.zero_bank[13] = 0;
	}
}

#ifndef TH_MINB
#define TH_MINB 8
#endif
__global__ void __launch_bounds__(256, TH_MINB) k_thresh(ThreshArgs a)
{
	KT_SCOPE(KT_THRESH, a.c_begin / (unsigned)a.nx, threadIdx.x == 0);
	extern __shared__ double s_DE[];
	double *s_D = s_DE, *s_E = s_DE + a.J + 2;
	for (int i = threadIdx.x; i < a.J + 2; i += blockDim.x) { s_D[i] = a.Dmono[i]; s_E[i] = a.Emono[i]; }
	__syncthreads();
	thresh_zero_bank(a);
	const unsigned long long c = a.c_begin + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	const bool in_range = c < a.c_end;
	if (!a.est && !in_range) return;
	uint32_t o0 = 0, o1 = 0;
	if (in_range) { o0 = __ldg(a.off + c); o1 = __ldg(a.off + c + 1); thresh_guard(a, o0, o1); }
	if (!a.est && o0 == o1) return;
	const int J = a.J, JP = J + 1;
	const int x = in_range ? (int)(c % (unsigned)a.nx) : 0, y = in_range ? (int)(c / (unsigned)a.nx) : 0;
	uint32_t l0 = 0, l1 = 0, r0 = 0, r1 = 0, u0 = 0, u1 = 0, d0 = 0, d1 = 0;
	if (o0 != o1) {
		if (x > 0) { l0 = __ldg(a.off + c - 1); l1 = o0; thresh_guard(a, l0, l1); }
		if (x < a.nx - 1) { r0 = o1; r1 = __ldg(a.off + c + 2); thresh_guard(a, r0, r1); }
		if (y > 0) { u0 = __ldg(a.off + c - a.nx); u1 = __ldg(a.off + c - a.nx + 1); thresh_guard(a, u0, u1); }
		if (y < a.ny - 1) { d0 = __ldg(a.off + c + a.nx); d1 = __ldg(a.off + c + a.nx + 1); thresh_guard(a, d0, d1); }
	}
	unsigned int cost_l = 0, cost_s = 0, cost_r = 0;     // pairs of this column with outputs in the tile on the left / its own / on the right
	const int xi = x & (P1_W - 1);
	const bool dual = a.dual != 0;
	if (dual && o1 - o0 > 1u) *a.dual_bad = 1u;
	for (uint32_t k = o0; k < o1; ++k) {
		double2 p = __ldg(a.spans + k);
		if (dual) {
			if (!(p.x > a.dual_lo && p.x <= p.y && p.y < a.dual_hi)) *a.dual_bad = 1u;
			p = make_double2(-p.x, -p.y);
		}
		const double m = 1e-9 + 1e-13 * (fabs(p.x) + fabs(p.y));
		// An entry with z1 > z2 is not an interval; it is never pruned and never prunes. (The one known source - the
		// erosion of data outside [zmin, zmax], where negate_ray prepends the bound without looking,
		// MorphologyOperators.cpp:241-248 - is sent to the unpruned kernel by erode() anyway.)
		const bool proper = dual || p.x <= p.y;
		const double pinf = __longlong_as_double(0x7FF0000000000000LL);
		// (l0, r0, u0, d0 advance with p: nn_of's scan start)
		const double nnL = proper ? nn_of(p, a.spans, l0, l1, a.clip_lo, a.clip_hi, dual) : pinf, nnR = proper ? nn_of(p, a.spans, r0, r1, a.clip_lo, a.clip_hi, dual) : pinf;
		const double nnU = proper ? nn_of(p, a.spans, u0, u1, a.clip_lo, a.clip_hi, dual) : pinf, nnD = proper ? nn_of(p, a.spans, d0, d1, a.clip_lo, a.clip_hi, dual) : pinf;
		// an interval saturated on both sides covers the whole range: it yields to a CLOSER one of its kind (near
		// tests) but never to a farther one - otherwise two of them could drop each other
		const bool whole = p.x <= a.clip_lo && p.y >= a.clip_hi;
		const int tnL = first_ge(s_D, J, nnL + m), tnR = first_ge(s_D, J, nnR + m);
		const int tyu = first_ge(s_E, J, nnU + m), tyd = first_ge(s_E, J, nnD + m);
		const int T = max(tyu, tyd);
		int tfL = 1, tfR = 1;
		if (J >= 1) {
			const float *g = a.G + (size_t)(T - 1) * JP;
			tfL = first_gt(g, J, far_value(nnR, m, whole));         // outputs on the left: the right neighbour is farther
			tfR = first_gt(g, J, far_value(nnL, m, whole));
		}
		const uint32_t layer = min(k - o0, 3u);
		uint4 t;
		t.x = (uint32_t)tnL | ((uint32_t)tnR << 8) | ((uint32_t)tfL << 16) | ((uint32_t)tfR << 24);
		t.y = (uint32_t)(tyu - 1) | ((uint32_t)(tyd - 1) << 8) | (layer << 16) | ((uint32_t)__ldg(a.reach + T - 1) << 24);
		t.z = __float_as_uint(far_value(nnD, m, whole));            // consumers above: row y+1 is farther
		t.w = __float_as_uint(far_value(nnU, m, whole));
		a.thr[k] = t;
		if (a.est) {
			// the distances Tile::scatter will list this interval at: {0} and [s, hi] on the left, [s, hi] on the right
			const int rx = __ldg(a.reach + T - 1);
			const unsigned int w = 4u + (unsigned int)T;
			const int hiL = min(tnL - 1, J), sL = max(1, min(tfL, rx)), hiR = min(tnR - 1, J), sR = max(1, min(tfR, rx));
			const int nL = max(0, hiL - sL + 1), nLs = max(0, min(hiL, xi) - sL + 1);
			const int nR = max(0, hiR - sR + 1), nRs = max(0, min(hiR, P1_W - 1 - xi) - sR + 1);
			cost_s += w * (unsigned int)(1 + nLs + nRs);
			cost_l += w * (unsigned int)(nL - nLs);
			cost_r += w * (unsigned int)(nR - nRs);
		}
	}
	if (a.est) {
		// one atomic per warp and tile (lanes are consecutive columns: usually a single tile)
		const unsigned int tile = in_range ? (unsigned int)y * (unsigned int)a.tiles_xw + (unsigned int)(x / P1_W) : 0xffffffffu;
		// (a warp usually lies inside one tile - always when the row length is a multiple of the tile width: one plain
		// reduction then, instead of grouping the lanes by tile first)
		const bool one_tile = __all_sync(0xffffffffu, tile == __shfl_sync(0xffffffffu, tile, 0));
		const unsigned int grp = one_tile ? 0xffffffffu : __match_any_sync(0xffffffffu, tile);
		const unsigned int sl = __reduce_add_sync(grp, cost_l), ss = __reduce_add_sync(grp, cost_s), sr = __reduce_add_sync(grp, cost_r);
		if (in_range && (threadIdx.x & 31) == __ffs(grp) - 1) {
			const int tx = x / P1_W;
			if (ss) atomicAdd(a.est + tile, ss);
			if (sl && tx > 0) atomicAdd(a.est + tile - 1, sl);
			if (sr && tx + 1 < a.tiles_xw) atomicAdd(a.est + tile + 1, sr);
		}
	}
}

// Four lanes per column, one per lattice neighbour (0: left, 1: right, 2: row above, 3: row below): each computes
// the nn values of the column's intervals against ITS neighbour and the near threshold that follows from them, the
// quad exchanges them by shuffles, the x lanes then search the far thresholds (they need T = max of the y lanes'
// results) and lane 0 writes the 16 bytes. Columns with many intervals (lattices: 10 x 10 pairs per neighbour)
// get four times the parallelism. Used for volumes with several intervals per column (launch_thresh in vo_lib.cu);
// for sparse height-field-like volumes the four-fold thread count costs more than it gives (C5: 0.27 vs 0.12 ms).
constexpr int TH_Q = 4;
__global__ void __launch_bounds__(256) k_thresh_quad(ThreshArgs a)
{
	KT_SCOPE(KT_THRESH, a.c_begin / (unsigned)a.nx, threadIdx.x == 0);
	extern __shared__ double s_DE[];
	double *s_D = s_DE, *s_E = s_DE + a.J + 2;
	for (int i = threadIdx.x; i < a.J + 2; i += blockDim.x) { s_D[i] = a.Dmono[i]; s_E[i] = a.Emono[i]; }
	__syncthreads();
	thresh_zero_bank(a);
	const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	const unsigned long long c = a.c_begin + t / TH_Q;
	const int dir = (int)(threadIdx.x & (TH_Q - 1)), lane = (int)(threadIdx.x & 31);
	const unsigned int qmask = 0xFu << (lane & ~(TH_Q - 1));
	const bool in_range = c < a.c_end;
	if (!a.est && !in_range) return;
	uint32_t o0 = 0, o1 = 0;
	if (in_range) { o0 = __ldg(a.off + c); o1 = __ldg(a.off + c + 1); thresh_guard(a, o0, o1); }
	if (!a.est && o0 == o1) return;
	const int J = a.J, JP = J + 1;
	const int x = in_range ? (int)(c % (unsigned)a.nx) : 0, y = in_range ? (int)(c / (unsigned)a.nx) : 0;
	uint32_t q0 = 0, q1 = 0;                             // this lane's neighbour column (empty range: none)
	if (o0 != o1) {
		if (dir == 0) { if (x > 0) { q0 = __ldg(a.off + c - 1); q1 = o0; } }
		else if (dir == 1) { if (x < a.nx - 1) { q0 = o1; q1 = __ldg(a.off + c + 2); } }
		else if (dir == 2) { if (y > 0) { q0 = __ldg(a.off + c - a.nx); q1 = __ldg(a.off + c - a.nx + 1); } }
		else { if (y < a.ny - 1) { q0 = __ldg(a.off + c + a.nx); q1 = __ldg(a.off + c + a.nx + 1); } }
		thresh_guard(a, q0, q1);
	}
	unsigned int cost_l = 0, cost_s = 0, cost_r = 0;     // pairs of this column with outputs in the tile on the left / its own / on the right
	const int xi = x & (P1_W - 1);
	for (uint32_t k = o0; k < o1; ++k) {                 // (the same trip count in all four lanes of a quad)
		const double2 p = __ldg(a.spans + k);
		const double m = 1e-9 + 1e-13 * (fabs(p.x) + fabs(p.y));
		// An entry with z1 > z2 is not an interval; it is never pruned and never prunes. (The one known source - the
		// erosion of data outside [zmin, zmax], where negate_ray prepends the bound without looking,
		// MorphologyOperators.cpp:241-248 - is sent to the unpruned kernel by erode() anyway.)
		const bool proper = p.x <= p.y;
		const double pinf = __longlong_as_double(0x7FF0000000000000LL);
		const double nn = proper ? nn_of(p, a.spans, q0, q1, a.clip_lo, a.clip_hi) : pinf;
		// an interval saturated on both sides covers the whole range: it yields to a CLOSER one of its kind (near
		// tests) but never to a farther one - otherwise two of them could drop each other
		const bool whole = p.x <= a.clip_lo && p.y >= a.clip_hi;
		const int tnear = first_ge(dir < 2 ? s_D : s_E, J, nn + m);      // tnL, tnR, Ty_up, Ty_dn by lane
		const float fv = far_value(nn, m, whole);
		const int tnL = __shfl_sync(qmask, tnear, 0, TH_Q), tnR = __shfl_sync(qmask, tnear, 1, TH_Q);
		const int tyu = __shfl_sync(qmask, tnear, 2, TH_Q), tyd = __shfl_sync(qmask, tnear, 3, TH_Q);
		const int T = max(tyu, tyd);
		// far, x: the outputs on the LEFT have the right neighbour as the farther one, so lane 1 finds tfL and lane 0 tfR
		int tfar = 1;
		if (J >= 1 && dir < 2) tfar = first_gt(a.G + (size_t)(T - 1) * JP, J, fv);
		const int tfL = __shfl_sync(qmask, tfar, 1, TH_Q), tfR = __shfl_sync(qmask, tfar, 0, TH_Q);
		const float vD = __shfl_sync(qmask, fv, 3, TH_Q), vU = __shfl_sync(qmask, fv, 2, TH_Q);
		if (dir == 0) {
			const int rx = __ldg(a.reach + T - 1);
			const uint32_t layer = min(k - o0, 3u);
			uint4 th;
			th.x = (uint32_t)tnL | ((uint32_t)tnR << 8) | ((uint32_t)tfL << 16) | ((uint32_t)tfR << 24);
			th.y = (uint32_t)(tyu - 1) | ((uint32_t)(tyd - 1) << 8) | (layer << 16) | ((uint32_t)rx << 24);
			th.z = __float_as_uint(vD);                                 // consumers above: row y+1 is farther
			th.w = __float_as_uint(vU);
			a.thr[k] = th;
			if (a.est) {
				// the distances Tile::scatter will list this interval at: {0} and [s, hi] on the left, [s, hi] on the right
				const unsigned int w = 4u + (unsigned int)T;
				const int hiL = min(tnL - 1, J), sL = max(1, min(tfL, rx)), hiR = min(tnR - 1, J), sR = max(1, min(tfR, rx));
				const int nL = max(0, hiL - sL + 1), nLs = max(0, min(hiL, xi) - sL + 1);
				const int nR = max(0, hiR - sR + 1), nRs = max(0, min(hiR, P1_W - 1 - xi) - sR + 1);
				cost_s += w * (unsigned int)(1 + nLs + nRs);
				cost_l += w * (unsigned int)(nL - nLs);
				cost_r += w * (unsigned int)(nR - nRs);
			}
		}
	}
	if (a.est) {
		// one atomic per warp and tile (a warp holds eight consecutive columns: usually a single tile)
		const unsigned int tile = in_range ? (unsigned int)y * (unsigned int)a.tiles_xw + (unsigned int)(x / P1_W) : 0xffffffffu;
		const bool one_tile = __all_sync(0xffffffffu, tile == __shfl_sync(0xffffffffu, tile, 0));
		const unsigned int grp = one_tile ? 0xffffffffu : __match_any_sync(0xffffffffu, tile);
		const unsigned int sl = __reduce_add_sync(grp, cost_l), ss = __reduce_add_sync(grp, cost_s), sr = __reduce_add_sync(grp, cost_r);
		if (in_range && lane == __ffs(grp) - 1) {
			const int tx = x / P1_W;
			if (ss) atomicAdd(a.est + tile, ss);
			if (sl && tx > 0) atomicAdd(a.est + tile - 1, sl);
			if (sr && tx + 1 < a.tiles_xw) atomicAdd(a.est + tile + 1, sr);
		}
	}
}

// ---- tile order: expensive tiles first (costs differ by two orders of magnitude; a launch that ends with the
// expensive ones ends with a long tail of a few busy warps). Buckets of log2(cost), descending; the order inside a
// bucket is whatever the atomics give.
constexpr int P1_NBUCKET = 32;
struct OrderArgs {
	const unsigned int *est;
	unsigned int tile0, ntiles0, tile0b, ntiles;       // positions [0, ntiles) <-> tiles, as in Pass1TileArgs
	unsigned int *hist;                                // [2 * P1_NBUCKET], zeroed by the host: counts, then fill cursors
	unsigned int *order;                               // [ntiles]
};
__device__ __forceinline__ int order_bucket(unsigned int est) { return P1_NBUCKET - 1 - (est ? 32 - __clz(est) : 0) * (P1_NBUCKET - 1) / 32; }

__global__ void __launch_bounds__(256) k_order_count(OrderArgs a)
{
	KT_SCOPE(KT_ORDER_COUNT, a.tile0, threadIdx.x == 0);
	__shared__ unsigned int h[P1_NBUCKET];
	if (threadIdx.x < P1_NBUCKET) h[threadIdx.x] = 0;
	__syncthreads();
	const unsigned int pos = blockIdx.x * blockDim.x + threadIdx.x;
	if (pos < a.ntiles) {
		const unsigned int tile = pos < a.ntiles0 ? a.tile0 + pos : a.tile0b + (pos - a.ntiles0);
		atomicAdd(&h[order_bucket(__ldg(a.est + tile))], 1u);
	}
	__syncthreads();
	if (threadIdx.x < P1_NBUCKET && h[threadIdx.x]) atomicAdd(a.hist + threadIdx.x, h[threadIdx.x]);
}

__global__ void __launch_bounds__(256) k_order_place(OrderArgs a)
{
	KT_SCOPE(KT_ORDER_PLACE, a.tile0, threadIdx.x == 0);
	__shared__ unsigned int base[P1_NBUCKET], h[P1_NBUCKET], start[P1_NBUCKET];
	if (threadIdx.x < P1_NBUCKET) h[threadIdx.x] = 0;
	if (threadIdx.x == 0) {
		unsigned int acc = 0;
		for (int b = 0; b < P1_NBUCKET; ++b) { base[b] = acc; acc += a.hist[b]; }
	}
	__syncthreads();
	const unsigned int pos = blockIdx.x * blockDim.x + threadIdx.x;
	unsigned int tile = 0, rank = 0;
	int b = -1;
	if (pos < a.ntiles) {
		tile = pos < a.ntiles0 ? a.tile0 + pos : a.tile0b + (pos - a.ntiles0);
		b = order_bucket(__ldg(a.est + tile));
		rank = atomicAdd(&h[b], 1u);
	}
	__syncthreads();
	if (threadIdx.x < P1_NBUCKET && h[threadIdx.x]) start[threadIdx.x] = atomicAdd(a.hist + P1_NBUCKET + threadIdx.x, h[threadIdx.x]);
	__syncthreads();
	if (b >= 0) a.order[base[b] + start[b] + rank] = tile;
}

// The same permutation by ONE CTA (count, bucket starts and placement between CTA barriers): for the launch sets of a
// band of the host-buffer call, where two dependent launches cost more than the few microseconds this takes.
constexpr unsigned int P1_ORDER_ONE_MAX = 1u << 15;
__global__ void __launch_bounds__(1024) k_order_one(OrderArgs a)
{
	KT_SCOPE(KT_ORDER_COUNT, a.tile0, threadIdx.x == 0);
	__shared__ unsigned int h[P1_NBUCKET], base[P1_NBUCKET];
	if (threadIdx.x < P1_NBUCKET) h[threadIdx.x] = 0;
	__syncthreads();
	for (unsigned int pos = threadIdx.x; pos < a.ntiles; pos += blockDim.x) {
		const unsigned int tile = pos < a.ntiles0 ? a.tile0 + pos : a.tile0b + (pos - a.ntiles0);
		atomicAdd(&h[order_bucket(__ldg(a.est + tile))], 1u);
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned int acc = 0;
		for (int b = 0; b < P1_NBUCKET; ++b) { base[b] = acc; acc += h[b]; }
	}
	__syncthreads();
	for (unsigned int pos = threadIdx.x; pos < a.ntiles; pos += blockDim.x) {
		const unsigned int tile = pos < a.ntiles0 ? a.tile0 + pos : a.tile0b + (pos - a.ntiles0);
		a.order[atomicAdd(&base[order_bucket(__ldg(a.est + tile))], 1u)] = tile;
	}
}

// ---- pass 1 ------------------------------------------------------------------------------------------
struct Pass1TileArgs {
	int nx, ny, J, cmax;
	int dbuf = 1;           // candidates double-buffered (pass1_warp_smem)
	int lean = 0;           // candidates and thresholds stay in global memory (pass1_warp_smem)
	int tiles_xw;           // tiles (P1_W columns) per row
	int tiles_x;            // tile masks (P1_TX columns) per row
	unsigned int tile0, ntiles;         // tiles [tile0, tile0 + ntiles0) and [tile0b, tile0b + ntiles - ntiles0) when `tiles` is NULL
	unsigned int tile0b, ntiles0;
	const uint32_t *off;
	const double2 *spans;
	const uint4 *thr;       // k_thresh output
	const double *Ht;       // (J+1) rows of JPP doubles: Ht[d*JPP + j] = H[j][d] (-1 beyond the reach / the table)
	const float *Ef;        // (J+1)*(J+2): Ef[d*(J+2) + c], rounded up, +inf for c > jmax[d]
	const uint8_t *jmax;    // J+2: largest class whose reach covers distance d
	double2 *mid;
	uint16_t *flags;        // [2][ny*nx]: lo | hi << 8 of the class window needed by the consumers above ([0]) / below ([1])
	unsigned long long *tilemask;   // [2][ny*tiles_x]: OR of those windows over the columns of a tile (bit j = class j), zeroed by the host
	double2 *pool;
	unsigned long long *cursor;
	unsigned long long pool_cap;
	uint32_t *ovf;          // [CTAs * warps][P1_OVF][P1_W]: survivor entries beyond the shared-memory lists
	const unsigned int *bad = nullptr;   // optional: raised by k_thresh when the (untrusted) offsets of the launch set's rows are
	                                     //   invalid - the launch then does nothing (ThreshArgs::bad)
	unsigned int *sticky_multi = nullptr, *sticky_big = nullptr;   // optional: count the tiles handed to a list (multi_tiles / big_tiles)
	                                     //   whose launch the caller has left out (never zeroed by a launch set: the banded host-buffer call
	                                     //   reads it once at the end and repeats with the list launches when it is not 0)
	unsigned long long *dbg;   // development aid (NULL normally): per tile {cycles, candidates, survivor entries, cycles of phase 1}
	Redo redo;              // slot ids to be (re)done by k_pass1: list overflow and oversized tiles
	// Device-side dispatch (no host round trip between the launches). Every launch is one resident wave of CTAs
	// pulling tiles with an atomic counter (tile costs vary a lot):
	//   launch 1  <MULTI=false>, small candidate buffer, all tiles. Tiles with a multi-interval column go to
	//             multi_tiles, tiles with more than cmax candidates to big_tiles.
	//   launch 2  <MULTI=false>, large buffer, tiles = big_tiles.
	//   launch 3  <MULTI=true> (two hulls per class), tiles = multi_tiles.
	int layer_major = -1;             // lean multi-interval launches, order in which phase 1 walks the candidates of a tile: -1 layer-major for tiles
	                                  //      with three layers and more (see `stage` in k_pass1_tile), 0 column-major, 1 layer-major (vo_set_option("cand_order", ...))
	int quota = 0;                    // > 0: a warp takes at most this many tiles and leaves (launch 1 of the host-buffer pipeline: a
	                                  //      grid of short-lived CTAs, so that the small kernels of the other bands get SM slots all
	                                  //      the time instead of waiting for a resident wave to retire); 0: warps stay until the tiles run out
	const unsigned int *tiles;        // NULL: all tiles of [tile0, tile0 + ntiles)
	const unsigned int *order;        // with tiles == NULL, optional: the same tiles as a permutation (k_order_place), expensive ones first
	const unsigned int *tiles_count;  // length of `tiles` (device side)
	unsigned int *tiles_next;         // next position to hand out
	unsigned int *big_tiles;          // NULL: oversized tiles go to the redo list (k_pass1)
	unsigned int *big_count;
	unsigned int *multi_tiles;
	unsigned int *multi_count;
};

__host__ __device__ inline int pass1_jpp(int J) { return (J + 1 + P1_CB + 1) & ~1; }

// dynamic shared memory: the cap tables once per CTA, then one staging area per warp
__host__ __device__ inline size_t pass1_table_smem(int J)
{
	const size_t JP = (size_t)J + 1;
	size_t b = JP * pass1_jpp(J) * sizeof(double);          // Ht
	b += ((JP * (JP + 1) + 3) & ~(size_t)3) * sizeof(float);// Ef
	b += (JP + 1 + 15) & ~(size_t)15;                       // jmax
	return b;
}
// dbuf: the candidates are double-buffered (the next tile is staged while the current one is in phase 2). Large buffers
// (dense columns: erosion's complement) take one: 34 instead of 52 bytes per candidate, i.e. half again as many
// warps per SM, against ~1.5 us of exposed copy latency per tile - tiles that large take tens of microseconds.
// lean: nothing but the segment column and the layer of a candidate (2 bytes) lives in shared memory; phase 1 reads the
// thresholds from global memory (coalesced, one batch ahead), phase 2 reads the few surviving candidates through L1.
// Used where the buffers would otherwise leave room for a few warps only (lattices with ten intervals per column: 430
// candidates per tile on average, far more in some): 2048 candidates cost 4 KB per warp instead of 70-106 KB.
__host__ __device__ inline size_t pass1_warp_smem(int J, int cmax, int lcap, bool dbuf = true, bool lean = false)
{
	const size_t SEG = (size_t)P1_W + 2 * J, nb = (dbuf && !lean) ? 2 : 1;
	size_t b = 0;
	if (!lean) {
		b += nb * (size_t)cmax * sizeof(double2);           // candidates
		b += (size_t)cmax * sizeof(uint4);                  // their thresholds: only read by phase 1, ONE buffer, refilled after it
	}
	b += 2 * ((SEG + 4) & ~(size_t)3) * sizeof(uint32_t);   // segment offsets (always double-buffered)
	b += (size_t)P1_W * sizeof(uint32_t);                   // list lengths
	b += (size_t)lcap * P1_W * sizeof(uint32_t);            // survivor lists [s][lane]
	b += 2 * nb * (((size_t)cmax + 15) & ~(size_t)15);      // segment column and layer of each candidate
	if (lean) b += 2 * (((size_t)cmax + 15) & ~(size_t)15); // layer-major order of the candidates (uint16, Tile::perm)
	b += 2 * sizeof(unsigned long long);                    // mbarriers of the two staging buffers
	return b;
}
__host__ __device__ inline size_t pass1_tile_smem(int J, int cmax, int lcap, int nwarps, bool dbuf = true, bool lean = false)
{
	return pass1_table_smem(J) + (size_t)nwarps * pass1_warp_smem(J, cmax, lcap, dbuf, lean) + 32;
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

// Class windows of a pair from its thresholds: lo_u | hi_u << 8 | lo_d << 16 | hi_d << 24 (0 = both empty).
// base: first class the far test in x leaves to the pair; fu / fd: classes the far test in y drops (leading).
__device__ __forceinline__ uint32_t pair_windows(int tyu, int tyd, int jm1, int base, int fu, int fd)
{
	int hu = min(tyu, jm1), hd = min(tyd, jm1);
	int lu = max(base, fu), ld = max(base, fd);
	if (lu > 0 || ld > 0) { lu = max(lu, 1); ld = max(ld, 1); }          // class 0 needs both windows
	if (lu >= hu) lu = hu = 0;
	if (ld >= hd) ld = hd = 0;
	return (uint32_t)lu | ((uint32_t)hu << 8) | ((uint32_t)ld << 16) | ((uint32_t)hd << 24);
}

// Shared-memory layout of the tile kernel (see pass1_tile_smem).
struct TableSmem {
	double *Ht;
	float *Ef;
	uint8_t *jmax;
	__device__ __forceinline__ TableSmem(unsigned char *raw, int J)
	{
		const int JP = J + 1;
		Ht = reinterpret_cast<double *>(raw);
		Ef = reinterpret_cast<float *>(Ht + (size_t)JP * pass1_jpp(J));
		jmax = reinterpret_cast<uint8_t *>(Ef + ((JP * (JP + 1) + 3) & ~3));
	}
};
struct WarpSmem {
	double2 *cand[2];
	uint4 *thr;
	uint32_t *off[2], *cnt, *list;
	uint8_t *ci[2], *ly[2];
	uint16_t *perm;                // lean only: the candidates in layer-major order (see `stage` in k_pass1_tile)
	unsigned long long *mbar;      // [2]
	__device__ __forceinline__ WarpSmem(unsigned char *raw, int J, int cmax, int lcap, bool dbuf, bool lean)
	{
		const int SEG = P1_W + 2 * J, c16 = (cmax + 15) & ~15;
		cand[0] = reinterpret_cast<double2 *>(raw);
		cand[1] = dbuf ? cand[0] + cmax : cand[0];          // (single buffer: both names, one array)
		thr = reinterpret_cast<uint4 *>(cand[1] + cmax);
		off[0] = lean ? reinterpret_cast<uint32_t *>(raw) : reinterpret_cast<uint32_t *>(thr + cmax);   // (lean: no candidate / threshold arrays)
		off[1] = off[0] + ((SEG + 4) & ~3);
		cnt = off[1] + ((SEG + 4) & ~3);
		list = cnt + P1_W;
		ci[0] = reinterpret_cast<uint8_t *>(list + (size_t)lcap * P1_W);
		ci[1] = dbuf ? ci[0] + c16 : ci[0];
		ly[0] = ci[1] + c16;
		ly[1] = dbuf ? ly[0] + c16 : ly[0];
		perm = reinterpret_cast<uint16_t *>(ly[1] + c16);
		mbar = reinterpret_cast<unsigned long long *>(ly[1] + c16 + (lean ? 2 * c16 : 0));
	}
};

// ---- bulk asynchronous copies (TMA, 1D) global -> shared, completion on an mbarrier -----------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned int count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned int bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned int parity)
{
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"WAIT_%=:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		"@p bra DONE_%=;\n"
		"bra WAIT_%=;\n"
		"DONE_%=:\n"
		"}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// bytes: multiple of 16; src / dst 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned int bytes, unsigned long long *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Survivor list entry: k | lo_u << 11 | n_u << 17 | lo_d << 21 | n_d << 27 (candidate index, the two class windows
// as start + length; lengths up to 15, longer windows are split over several entries).
__device__ __forceinline__ uint32_t entry_windows(uint32_t e)        // -> lo_u | hi_u << 8 | lo_d << 16 | hi_d << 24
{
	const uint32_t lu = (e >> 11) & 63u, nu = (e >> 17) & 15u, ld = (e >> 21) & 63u, nd = (e >> 27) & 15u;
	return lu | ((lu + nu) << 8) | (ld << 16) | ((ld + nd) << 24);
}

// The staged tile.
template <int LCAP>
struct Tile {
	const double2 *cand;
	const uint4 *thr;          // thresholds of the candidates: valid during phase 1 only (the buffer is refilled for the next tile)
	const uint4 *gthr;         // ... the same in global memory (phase 2, direct mode)
	const uint8_t *ci;         // segment column of each candidate
	const uint8_t *ly;         // its layer (position inside its column, 3 = third or later)
	const double *Ht;
	const float *Ef;
	const uint8_t *jmax;
	uint32_t *cnt;             // [P1_W] list lengths
	uint32_t *list;            // [s * P1_W + xi]
	uint32_t *ovf;             // [(s - LCAP) * P1_W + xi], global memory
	int J, JPP;

	__device__ __forceinline__ void push(int xo, int k, uint32_t w) const
	{
		int lu = (int)(w & 0xffu), nu = (int)((w >> 8) & 0xffu) - lu, ld = (int)((w >> 16) & 0xffu), nd = (int)(w >> 24) - ld;
		do {
			const int cu = min(nu, 15), cd = min(nd, 15);
			const uint32_t pos = atomicAdd(cnt + xo, 1u);
			const uint32_t e = (uint32_t)k | ((uint32_t)lu << 11) | ((uint32_t)cu << 17) | ((uint32_t)ld << 21) | ((uint32_t)cd << 27);
			if (pos < (uint32_t)LCAP) list[pos * P1_W + xo] = e;
			else if (pos < (uint32_t)(LCAP + P1_OVF)) ovf[(pos - LCAP) * P1_W + xo] = e;
			lu += cu; nu -= cu; ld += cd; nd -= cd;
		} while (nu > 0 || nd > 0);
	}

	// Phase 1: candidate k appends itself to the lists of the output columns it survives for.
	__device__ __forceinline__ void scatter(const uint4 th, int k, int txe) const
	{
		const int i = (int)ci[k];
		const int tyu = (int)(th.y & 0xffu) + 1, tyd = (int)((th.y >> 8) & 0xffu) + 1, T = max(tyu, tyd), rx = (int)(th.y >> 24);
		const float vu = __uint_as_float(th.z), vd = __uint_as_float(th.w);
#pragma unroll 1
		for (int side = 0; side < 2; ++side) {             // 0: outputs on the left (ix = i - d, incl. d = 0), 1: on the right
			const int tn = side ? (int)((th.x >> 8) & 0xffu) : (int)(th.x & 0xffu);
			const int tf = side ? (int)(th.x >> 24) : (int)((th.x >> 16) & 0xffu);
			// distances that hit an output column [J, J + txe) and survive the near test
			int d = side ? max(J - i, 1) : max(i - (J + txe - 1), 0);
			const int dhi = min(side ? J + txe - 1 - i : i - J, min(tn - 1, J));
			int fu = -1, fd = -1;                           // far, y: dropped leading classes (non-increasing in d)
#pragma unroll 1
			for (; d <= dhi; ++d) {
				int base = 0;
				if (d >= 1 && d < tf) {
					// far-dominated in x: only the classes beyond the farther neighbour's reach remain (d >= reach[T-1])
					if (d < rx) { d = min(tf, rx) - 1; continue; }
					base = (int)jmax[d + 1] + 1;
				}
				const int jm1 = (int)jmax[d] + 1;
				if (base >= min(T, jm1)) continue;
				const float *row = Ef + (size_t)d * (J + 2);
				if (fu < 0) { fu = first_gt(row, J + 1, vu) - 1; fd = first_gt(row, J + 1, vd) - 1; }
				else {
					while (fu > 0 && row[fu] > vu) --fu;
					while (fd > 0 && row[fd] > vd) --fd;
				}
				const uint32_t w = pair_windows(tyu, tyd, jm1, base, fu, fd);
				if (w == 0) continue;
				push((side ? i + d : i - d) - J, k, w);
			}
		}
	}
};

// Per-thread view of the staged tile (phase 2).
template <int LCAP>
struct TileThread {
	Tile<LCAP> t;
	int xi, ix, kb, niter, y, x0;
	bool direct;

	// direct mode (the list overflowed): candidate k of the range re-tested for this output
	__device__ __forceinline__ bool pair_direct(int k, int &d, uint32_t &w) const
	{
		const uint4 th = __ldg(t.gthr + k);
		const int i = (int)t.ci[k];
		d = abs(i - ix);
		const bool left = ix < i;
		const int tn = left ? (int)(th.x & 0xffu) : (int)((th.x >> 8) & 0xffu);
		if (d >= 1 && d >= tn) return false;
		const int tf = left ? (int)((th.x >> 16) & 0xffu) : (int)(th.x >> 24);
		const int tyu = (int)(th.y & 0xffu) + 1, tyd = (int)((th.y >> 8) & 0xffu) + 1;
		const int base = (d >= 1 && d < tf) ? (int)t.jmax[d + 1] + 1 : 0;
		const int jm1 = (int)t.jmax[d] + 1;
		if (base >= min(max(tyu, tyd), jm1)) return false;
		const float *row = t.Ef + (size_t)d * (t.J + 2);
		const int fu = first_gt(row, t.J + 1, __uint_as_float(th.z)) - 1, fd = first_gt(row, t.J + 1, __uint_as_float(th.w)) - 1;
		w = pair_windows(tyu, tyd, jm1, base, fu, fd);
		return w != 0;
	}
	// s-th entry of the survivor loop: candidate k, its distance d and its window word
	__device__ __forceinline__ bool survivor(int s, int &k, int &d, uint32_t &w) const
	{
		if (!direct) {
			const uint32_t e = s < LCAP ? t.list[s * P1_W + xi] : t.ovf[(s - LCAP) * P1_W + xi];
			k = (int)(e & 2047u);
			d = abs((int)t.ci[k] - ix);
			w = entry_windows(e);
			return true;
		}
		k = kb + s;
		return pair_direct(k, d, w);
	}
	__device__ __forceinline__ int layer(int k) const { return (int)t.ly[k]; }
};

// bits q in [0, CB) with l <= cb + q < h
template <int CB>
__device__ __forceinline__ unsigned int range_mask(int l, int h, int cb)
{
	const int a = min(max(l - cb, 0), CB), b = min(max(h - cb, 0), CB);
	return ((1u << b) - 1u) & ~((1u << a) - 1u);
}
// bits [lo, hi) of a 64-bit class mask (hi <= 64)
__device__ __forceinline__ unsigned long long class_mask(int lo, int hi)
{
	if (hi <= lo) return 0ull;
	const unsigned long long upto = hi >= 64 ? ~0ull : ((1ull << hi) - 1ull);
	return upto & ~((1ull << lo) - 1ull);
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
	// (list launches - LCAP == P1_LCAP_M - one survivor ahead: the candidate and its cap are loaded before the previous
	// survivor goes into the list; an insertion from the third interval on is a call, which no load is moved across. Lattices
	// 2 - 3 % of the launch. The first launch keeps the plain loop: there the function is all but idle, and the longer live
	// ranges cost the kernel around it 5 % - C5 k_pass1_tile 0.764 -> 0.803 ms, profiles/r2ch_ab.txt)
	if (LCAP != P1_LCAP_M) {
		for (int s = 0; s < t.niter; ++s) {
			int k, d;
			uint32_t w;
			if (!t.survivor(s, k, d, w)) continue;
			if (window_mask<1>(w, j)) {
				const double2 ab = t.t.cand[k];
				const double hh = t.t.Ht[(size_t)d * t.t.JPP + j];
				u.insert(ab.x - hh, ab.y + hh);
			}
		}
	}
	bool have = false;
	double2 ab_p = make_double2(0.0, 0.0);
	double hh_p = 0.0;
	for (int s = 0; LCAP == P1_LCAP_M && s < t.niter; ++s) {
		int k, d;
		uint32_t w;
		if (!t.survivor(s, k, d, w)) continue;
		if (window_mask<1>(w, j)) {
			const double2 ab = t.t.cand[k];
			const double hh = t.t.Ht[(size_t)d * t.t.JPP + j];
			if (have) u.insert(ab_p.x - hh_p, ab_p.y + hh_p);
			ab_p = ab; hh_p = hh; have = true;
		}
	}
	if (have) u.insert(ab_p.x - hh_p, ab_p.y + hh_p);
	if (u.overflow) { redo_push(a.redo, slot); return slot_empty(); }
	if (u.n == 0) return slot_empty();
	if (u.n == 1) return make_double2(u.s0, u.e0);
	const unsigned long long pb = atomicAdd(a.cursor, (unsigned long long)u.n);
	if (pb + u.n <= a.pool_cap)
		for (int q = 0; q < u.n; ++q) a.pool[pb + q] = u.get(q);
	return slot_pool(pb, (unsigned int)u.n);
}

// Classes [cb, cb + CB) in registers, NL hulls per class. Hull l collects the survivors that are the l-th
// interval of their column (the last hull also takes any further ones): for shells, slabs and complements
// (erosion) each such layer unions to ONE interval, so a class costs NL (lo, hi) pairs and no list. A
// survivor that misses the running hull of its layer makes the class "complex" (general path).
// `need` = the classes of the block some consumer reads (bit q = class cb + q).
// DUAL (erosion in dual form, vo_lib.cu: erode_dual): the candidates are read as the mirrored intervals (-z1, -z2);
// the hull (min, max) IS the result then - the intersection of the eroded intervals, mirrored - so no class is ever
// "complex".
template <int CB, int NL, int CAP, int LCAP, bool DUAL = false, bool GEN = true>
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
		int k, d;
		uint32_t w;
		if (!t.survivor(s, k, d, w)) continue;
		const unsigned int valid = window_mask<CB>(w, cb);
		if (!valid) continue;
		double2 ab = t.t.cand[k];
		if (DUAL) ab = make_double2(-ab.x, -ab.y);
		const double *hp = t.t.Ht + (size_t)d * t.t.JPP + cb;
		double h[CB];
#pragma unroll
		for (int q = 0; q < CB; ++q) h[q] = hp[q];
		// branch-free: a class that does not take this survivor gets the cap -inf, i.e. the candidate
		// (+inf, -inf), which leaves its hull untouched
		const int lay = NL == 1 ? 0 : min(t.layer(k), NL - 1);
#pragma unroll
		for (int l = 0; l < NL; ++l) {
			if (NL == 1 || lay == l) {
				unsigned int miss = 0;
#pragma unroll
				for (int q = 0; q < CB; ++q) {
					const double hq = ((valid >> q) & 1u) ? h[q] : -inf;
					const double cs = ab.x - hq, ce = ab.y + hq;
					if (!DUAL) miss |= (cs <= hi[l][q] && ce >= lo[l][q]) ? 0u : (1u << q);
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
		const unsigned long long slot = ((unsigned long long)t.y * (t.t.J + 1) + j) * a.nx + t.x0 + t.xi;
		double2 out;
		if ((complex_mask >> q) & 1u) {
			if (GEN) out = class_general<CAP, LCAP>(a, t, j, slot);
			else { redo_push(a.redo, slot); out = slot_empty(); }     // (the redo launch computes the slot from scratch)
		}
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
template <int CB, int NL, int CAP, int LCAP, bool DUAL = false, bool GEN = true>
__device__ __forceinline__ void eval_classes(const Pass1TileArgs &a, const TileThread<LCAP> &t, int UL, int UH, int DL, int DH)
{
	const int jend = max(UH, DH);
	int cb = UH > UL ? (DH > DL ? min(UL, DL) : UL) : DL;
	while (cb < jend) {
		const unsigned int need = range_mask<CB>(UL, UH, cb) | range_mask<CB>(DL, DH, cb);
		eval_block<CB, NL, CAP, LCAP, DUAL, GEN>(a, t, cb, need);
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

// What a tile turns out to be once its segment offsets are known (uniform over the CTA).
enum TileKind { TK_NORMAL = 0, TK_EMPTY, TK_MULTI, TK_BIG, TK_REDO };

struct TileHead {
	unsigned int tile;
	int y, x0, txe, ncand, kind;
	uint32_t base;
};

// Phase 2 of a staged tile: lane per output column (the whole warp enters; `active` = owns a column).
template <int CAP, bool MULTI, int LCAP, bool DUAL = false, bool GEN = true>
__device__ __forceinline__ void tile_phase2(const Pass1TileArgs &a, const TileHead &h, const Tile<LCAP> &tl, const uint32_t *s_off)
{
	const int J = a.J, lane = threadIdx.x & 31;
	const size_t rowbase = (size_t)h.y * a.nx, ncols_all = (size_t)a.nx * a.ny;
	const bool active = lane < h.txe;
	const int xi = lane, ix = xi + J;
	TileThread<LCAP> t;
	t.t = tl; t.xi = xi; t.ix = ix; t.y = h.y; t.x0 = h.x0;
	int UL = 255, UH = 0, DL = 255, DH = 0;
	int maxlayer = 0;
	if (active) {
		t.kb = (int)(s_off[ix - J] - h.base);
		const int S = (int)tl.cnt[xi];
		// more survivors than the lists hold (very many layers): re-scan the candidate range instead
		t.direct = S > LCAP + P1_OVF;
		t.niter = t.direct ? (int)(s_off[ix + J + 1] - h.base) - t.kb : S;
		for (int s = 0; s < t.niter; ++s) {
			int k, d;
			uint32_t w;
			if (!t.survivor(s, k, d, w)) continue;
			if (MULTI) maxlayer = max(maxlayer, t.layer(k));
			const int lu = (int)(w & 0xffu), hu = (int)((w >> 8) & 0xffu), ld = (int)((w >> 16) & 0xffu), hd = (int)(w >> 24);
			if (hu > lu) { UL = min(UL, lu); UH = max(UH, hu); }
			if (hd > ld) { DL = min(DL, ld); DH = max(DH, hd); }
		}
		if (UH == 0) UL = 0;
		if (DH == 0) DL = 0;
		a.flags[rowbase + h.x0 + xi] = (uint16_t)(UL | (UH << 8));
		a.flags[ncols_all + rowbase + h.x0 + xi] = (uint16_t)(DL | (DH << 8));
		tl.cnt[xi] = 0;                                    // ready for the next tile
	} else { UL = UH = DL = DH = 0; }
	// OR of the windows over the tile: pass 2 skips a producer row whose tile mask lacks the class
	{
		const unsigned long long mu = class_mask(UL, UH), md = class_mask(DL, DH);
		const unsigned int mu0 = __reduce_or_sync(0xffffffffu, (unsigned int)mu), mu1 = __reduce_or_sync(0xffffffffu, (unsigned int)(mu >> 32));
		const unsigned int md0 = __reduce_or_sync(0xffffffffu, (unsigned int)md), md1 = __reduce_or_sync(0xffffffffu, (unsigned int)(md >> 32));
		if (lane == 0) {
			const size_t ntl = (size_t)a.tiles_x * a.ny, mt = (size_t)h.y * a.tiles_x + h.x0 / P1_TX;
			if (mu0 | mu1) atomicOr(a.tilemask + mt, (unsigned long long)mu0 | ((unsigned long long)mu1 << 32));
			if (md0 | md1) atomicOr(a.tilemask + ntl + mt, (unsigned long long)md0 | ((unsigned long long)md1 << 32));
		}
	}
	if (!active || (UH == 0 && DH == 0)) return;
	if (!MULTI || maxlayer == 0) eval_classes<P1_CB, 1, CAP, LCAP, DUAL, GEN>(a, t, UL, UH, DL, DH);   // every survivor is the first interval of its column
	else if (maxlayer == 1) eval_classes<P1_CB, 2, CAP, LCAP>(a, t, UL, UH, DL, DH);        // two hulls per class
	else {
		// three layers and more (lattices, stacks of plates): two hulls per class would come out "complex" for nearly
		// every class - straight to the sorted-list union
		for (int j = min(UH > UL ? UL : DL, DH > DL ? DL : UL); j < max(UH, DH); ++j) {
			if (!((j >= UL && j < UH) || (j >= DL && j < DH))) continue;
			const unsigned long long slot = ((unsigned long long)t.y * (J + 1) + j) * a.nx + t.x0 + t.xi;
			a.mid[slot] = class_general<CAP, LCAP>(a, t, j, slot);
		}
	}
}

// Tiles that are not processed here: empty ones, ones for another launch, ones left to k_pass1.
template <bool MULTI>
__device__ __forceinline__ void tile_other(const Pass1TileArgs &a, const TileHead &h)
{
	const int lane = threadIdx.x & 31, JP = a.J + 1;
	const size_t rowbase = (size_t)h.y * a.nx, ncols_all = (size_t)a.nx * a.ny;
	if (h.kind == TK_EMPTY) {                               // nothing in reach: no slot of the tile is needed
		if (lane < h.txe) { a.flags[rowbase + h.x0 + lane] = 0; a.flags[ncols_all + rowbase + h.x0 + lane] = 0; }
	} else if (h.kind == TK_MULTI) {                        // -> two-hull variant, launch 3
		if (lane == 0) { a.multi_tiles[atomicAdd(a.multi_count, 1u)] = h.tile; if (a.sticky_multi) atomicAdd(a.sticky_multi, 1u); }
	} else if (h.kind == TK_BIG) {                          // -> launch 2, larger candidate buffer
		if (lane == 0) { a.big_tiles[atomicAdd(a.big_count, 1u)] = h.tile; if (a.sticky_big) atomicAdd(a.sticky_big, 1u); }
	} else {                                                // TK_REDO: leave every slot of the tile to k_pass1
		const uint16_t full = (uint16_t)(JP << 8);         // window [0, J+1)
		if (lane < h.txe) { a.flags[rowbase + h.x0 + lane] = full; a.flags[ncols_all + rowbase + h.x0 + lane] = full; }
		if (lane == 0) {
			const size_t ntl = (size_t)a.tiles_x * a.ny, mt = (size_t)h.y * a.tiles_x + h.x0 / P1_TX;
			atomicOr(a.tilemask + mt, ~0ull);
			atomicOr(a.tilemask + ntl + mt, ~0ull);
		}
		for (int idx = lane; idx < JP * h.txe; idx += 32) {
			const int j = idx / h.txe, xi = idx % h.txe;
			redo_push(a.redo, ((unsigned long long)h.y * JP + j) * a.nx + h.x0 + xi);
		}
	}
}

// Layer-major order of a tile's candidates (every column's first interval, then every column's second, ...): phase 1 of
// the lean multi-interval launches walks them in this order, so the survivors of an output column reach its list - and
// later the sorted-list union of class_general - roughly in ascending z (neighbouring columns cut the same features),
// where an insertion is an append or a merge with the tail. In column-major order every column after the first inserted
// its intervals into the middle of the list: shifts, bisections and merge scans on four lanes of 32 were a third of that
// launch on a lattice with ten intervals per column. Only the order changes, never the result. mode: -1 = for tiles
// with three layers and more (below that every class is two hulls, which no order changes), 0 never, 1 always.
// Called by the whole warp; returns whether `perm` now holds the order. (Out of line: the tile kernel sits at its
// register limit.)
__device__ __noinline__ bool layer_major_order(const uint32_t *s_off, uint32_t base, int SEG, uint16_t *perm, int mode)
{
	constexpr int NR = 5;                                   // SEG <= 32 + 2 * 63
	const int lane = threadIdx.x & 31;
	uint32_t cn[NR], kf[NR];
	uint32_t maxc = 0;
#pragma unroll
	for (int r = 0; r < NR; ++r) {
		const int i = lane + r * 32;
		cn[r] = i < SEG ? s_off[i + 1] - s_off[i] : 0u;
		kf[r] = i < SEG ? s_off[i] - base : 0u;
		maxc = max(maxc, cn[r]);
	}
	maxc = __reduce_max_sync(0xffffffffu, maxc);
	if (!(mode > 0 || (mode < 0 && maxc >= 3u))) return false;
	const unsigned int lt = (1u << lane) - 1u;
	uint32_t at = 0;
	for (uint32_t l = 0; l < maxc; ++l) {
#pragma unroll
		for (int r = 0; r < NR; ++r) {
			if (r * 32 < SEG) {
				const bool has = cn[r] > l;
				const unsigned int bal = __ballot_sync(0xffffffffu, has);
				if (has) perm[at + __popc(bal & lt)] = (uint16_t)(kf[r] + l);
				at += __popc(bal);
			}
		}
	}
	return true;
}

// One CTA per SM; the cap tables are staged once per CTA, then every WARP works on its own: it pulls tiles
// (P1_W output columns of one row) with an atomic counter and never meets a CTA barrier again - tile costs vary
// a lot (steep walls), and a warp that lags only delays itself.
// LIST = false: the tiles [tile0, tile0 + ntiles). LIST = true: the tiles of a list collected by an earlier
// launch (its length is only known on the device).
// Software pipeline over the tiles of a warp (global latency is what bounds a tile otherwise):
//   - the position of the tile after next is fetched (atomicAdd) while the current tile is processed;
//   - the segment offsets of the NEXT tile are loaded into registers before phase 1 of the current tile and
//     published to shared memory when phase 1 ends;
//   - right after that one lane starts the bulk copies (TMA, cp.async.bulk) of the next tile's candidates and
//     thresholds into the other staging buffer; they land during phase 2 of the current tile and are awaited
//     (mbarrier) at the top of the next iteration.
// GEN = false (first launch only): a class whose hull turns out "complex" is handed to the redo launch instead of being
// folded here by class_general. The call - and the registers the ABI keeps free around it - is what held this kernel at 128
// registers with spills; without it 88 registers and no spill, i.e. 20 warps per SM instead of 16 (C5 k_pass1_tile
// 0.645 -> 0.565 ms). Complex classes are rare in height-field-like input; a context that meets them falls back to the
// inline variant for its next calls (vo_ctx::gen_inline_calls).
template <int CAP, bool MULTI, bool LIST, bool DUAL = false, bool GEN = true>
__global__ void __launch_bounds__(32 * (MULTI ? P1_MAXWARPS_M : GEN ? P1_MAXWARPS : P1_MAXWARPS_LEAN), 1) k_pass1_tile(Pass1TileArgs a)
{
	constexpr int LCAP = (MULTI || LIST) ? P1_LCAP_M : P1_LCAP_S;
	constexpr int NR = 5;                                   // segment offsets per lane: P1_W + 2 * 63 + 1 <= 32 * NR
	extern __shared__ __align__(16) unsigned char smem_raw[];
	KT_SCOPE(LIST ? KT_TILE_LIST : KT_TILE, a.tile0 / (unsigned)a.tiles_xw, (threadIdx.x & 31) == 0);     // (per warp: warps leave on their own)
	const unsigned int n = LIST ? *a.tiles_count : a.ntiles;
	if (n == 0) return;
	if (a.bad) {                                            // invalid input offsets (banded host-buffer call): nothing is safe to read
		__shared__ unsigned int s_bad;                      // (one read per CTA: another band's k_thresh may raise the flag meanwhile)
		if (threadIdx.x == 0) s_bad = *a.bad;
		__syncthreads();
		if (s_bad) return;
	}
	const int J = a.J, JP = J + 1, SEG = P1_W + 2 * J, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const unsigned int FULL = 0xffffffffu;
	const TableSmem tb(smem_raw, J);
	for (int i = threadIdx.x; i < JP * pass1_jpp(J); i += blockDim.x) tb.Ht[i] = __ldg(a.Ht + i);
	for (int i = threadIdx.x; i < JP * (JP + 1); i += blockDim.x) tb.Ef[i] = __ldg(a.Ef + i);
	for (int i = threadIdx.x; i < JP + 1; i += blockDim.x) tb.jmax[i] = __ldg(a.jmax + i);
	const bool lean = LIST && a.lean != 0, dbuf = a.dbuf != 0 && !lean;
	const WarpSmem sm(smem_raw + pass1_table_smem(J) + (size_t)warp * pass1_warp_smem(J, a.cmax, LCAP, dbuf, lean), J, a.cmax, LCAP, dbuf, lean);
	sm.cnt[lane] = 0;
	if (lane == 0) {
		mbar_init(sm.mbar + 0, 1); mbar_init(sm.mbar + 1, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();                                        // the only CTA barrier: tables (and mbarriers) are in place

	Tile<LCAP> tl;
	tl.Ht = tb.Ht; tl.Ef = tb.Ef; tl.jmax = tb.jmax; tl.cnt = sm.cnt; tl.list = sm.list; tl.J = J;
	tl.ovf = a.ovf + ((size_t)blockIdx.x * (blockDim.x >> 5) + warp) * (size_t)P1_OVF * P1_W; tl.JPP = pass1_jpp(J);

	// segment offsets of the tile at list position `pos` -> registers (o[r][1] = the next column's offset, for the
	// multi-interval test)
	auto load_offsets = [&](unsigned int pos, uint32_t (&o)[NR][2], TileHead &h) {
		h.tile = LIST ? a.tiles[pos] : a.order ? __ldg(a.order + pos) : (pos < a.ntiles0 ? a.tile0 + pos : a.tile0b + (pos - a.ntiles0));
		h.y = (int)(h.tile / (unsigned)a.tiles_xw);
		h.x0 = (int)(h.tile % (unsigned)a.tiles_xw) * P1_W;
		h.txe = min(P1_W, a.nx - h.x0);
		const size_t rowbase = (size_t)h.y * a.nx;
#pragma unroll
		for (int r = 0; r < NR; ++r) {
			const int i = lane + r * 32;
			o[r][0] = o[r][1] = 0;
			if (i <= SEG) {
				o[r][0] = __ldg(a.off + rowbase + min(max(h.x0 - J + i, 0), a.nx));   // columns outside the grid collapse to empty ranges
				if (!MULTI && i < SEG) o[r][1] = __ldg(a.off + rowbase + min(max(h.x0 - J + i + 1, 0), a.nx));
			}
		}
	};
	// registers -> shared memory; returns "a column of the segment holds several intervals" (warp-uniform)
	auto store_offsets = [&](const uint32_t (&o)[NR][2], uint32_t *s_off) -> bool {
		bool multi = false;
#pragma unroll
		for (int r = 0; r < NR; ++r) {
			const int i = lane + r * 32;
			if (i <= SEG) { s_off[i] = o[r][0]; if (!MULTI && i < SEG) multi |= o[r][1] - o[r][0] > 1u; }
		}
		__syncwarp();                                       // the stores are visible to the whole warp
		return __any_sync(FULL, multi);
	};
	auto classify = [&](TileHead &h, const uint32_t *s_off, bool multi) {
		h.base = s_off[0];
		h.ncand = (int)(s_off[SEG] - h.base);
		h.kind = multi ? TK_MULTI : h.ncand > a.cmax ? (a.big_tiles ? TK_BIG : TK_REDO) : h.ncand == 0 ? TK_EMPTY : TK_NORMAL;
	};
	// start the staging of a NORMAL tile into buffer b: bulk copies by one lane, the column map by all
	bool perm_on = false;                                   // lean multi-interval launches: the tile staged last is walked in layer-major order
	auto stage = [&](const TileHead &h, int b) {
		if (lane == 0 && !lean) {
			const unsigned int bytes = (unsigned int)h.ncand * 16u;
			mbar_expect_tx(sm.mbar + b, 2u * bytes);
			bulk_g2s(sm.cand[b], a.spans + h.base, bytes, sm.mbar + b);
			bulk_g2s(sm.thr, a.thr + h.base, bytes, sm.mbar + b);
		}
		const uint32_t *s_off = sm.off[b];
		for (int i = lane; i < SEG; i += 32)
			for (uint32_t k0 = s_off[i] - h.base, k = k0; k < s_off[i + 1] - h.base; ++k) { sm.ci[b][k] = (uint8_t)i; sm.ly[b][k] = (uint8_t)min(k - k0, 3u); }
		if (MULTI && lean) perm_on = layer_major_order(s_off, h.base, SEG, sm.perm, a.layer_major);
	};
	int claims = 0;                                         // tiles this warp has asked for (a.quota)
	auto fetch_pos = [&]() -> unsigned int {
		if (a.quota > 0 && claims >= a.quota) return 0xffffffffu;
		++claims;
		return lane == 0 ? atomicAdd(a.tiles_next, 1u) : 0u;
	};

	// prologue: the first tile is loaded synchronously, the positions of the next two are in flight
	unsigned int pos = __shfl_sync(FULL, fetch_pos(), 0);
	if (pos >= n) return;
	unsigned int npos_raw = fetch_pos();                    // position of the next tile (lane 0 holds it)
	TileHead cur, nxt;
	uint32_t o[NR][2];
	int buf = 0;
	unsigned int phase[2] = {0u, 0u};
	load_offsets(pos, o, cur);
	{
		const bool multi = store_offsets(o, sm.off[0]);
		classify(cur, sm.off[0], multi);
		if (cur.kind == TK_NORMAL) stage(cur, 0);
		__syncwarp();
	}

	for (;;) {
		const unsigned int npos = __shfl_sync(FULL, npos_raw, 0);
		const bool have_next = npos < n;
		npos_raw = fetch_pos();                              // the one after it (used at the top of the next iteration)
		if (have_next) load_offsets(npos, o, nxt);

		// ---- current tile, phase 1 ----
#ifdef VO_TILE_DEBUG
		const long long dbg_t0 = a.dbg ? clock64() : 0;
#endif
		if (cur.kind == TK_NORMAL) {
			tl.gthr = a.thr + cur.base; tl.ci = sm.ci[buf]; tl.ly = sm.ly[buf];
			if (!lean) {
				mbar_wait(sm.mbar + buf, phase[buf]);        // candidates and thresholds have landed
				phase[buf] ^= 1u;
				tl.cand = sm.cand[buf]; tl.thr = sm.thr;
				for (int k = lane; k < cur.ncand; k += 32) tl.scatter(sm.thr[k], k, cur.txe);
			} else {
				__syncwarp();                                // (the column map of this tile, written by all lanes)
				tl.cand = a.spans + cur.base; tl.thr = tl.gthr;
				// (MULTI: candidates in layer-major order, see `stage`)
				auto cand_at = [&](int q) -> int { return (MULTI && perm_on) ? (int)sm.perm[q] : q; };
				int kc = lane < cur.ncand ? cand_at(lane) : 0;
				uint4 th = lane < cur.ncand ? __ldg(tl.gthr + kc) : make_uint4(0u, 0u, 0u, 0u);
				for (int k = lane; k < cur.ncand; k += 32) {     // thresholds one batch ahead of their use
					const int kn = k + 32 < cur.ncand ? cand_at(k + 32) : 0;
					const uint4 nx4 = k + 32 < cur.ncand ? __ldg(tl.gthr + kn) : make_uint4(0u, 0u, 0u, 0u);
					tl.scatter(th, kc, cur.txe);
					th = nx4; kc = kn;
				}
			}
		} else tile_other<MULTI>(a, cur);

		// ---- publish the next tile's offsets, start its staging (single candidate buffer: only after phase 2) ----
		if (have_next) {
			const bool nmulti = store_offsets(o, sm.off[buf ^ 1]);
			classify(nxt, sm.off[buf ^ 1], nmulti);
			if (dbuf && nxt.kind == TK_NORMAL) stage(nxt, buf ^ 1);
		}
		__syncwarp();                                       // phase 1 of the current tile is complete (lists visible)

		// ---- current tile, phase 2 ----
#ifdef VO_TILE_DEBUG                                     // (development builds: per-tile statistics for scripts/tile_costs.py)
		unsigned int dbg_entries = 0;
		const long long dbg_t1 = a.dbg ? clock64() : 0;
		if (a.dbg && cur.kind == TK_NORMAL) { __syncwarp(); dbg_entries = __reduce_add_sync(FULL, sm.cnt[lane]); }
#endif
		if (cur.kind == TK_NORMAL) tile_phase2<CAP, MULTI, LCAP, DUAL, GEN>(a, cur, tl, sm.off[buf]);
		__syncwarp();                                       // lists, counters and the staging buffer are free again
#ifdef VO_TILE_DEBUG
		if (a.dbg && lane == 0 && cur.kind == TK_NORMAL) {     // (scripts/tile_costs.py)
			unsigned long long *d = a.dbg + 4ull * cur.tile;
			d[0] = (unsigned long long)(clock64() - dbg_t0); d[1] = (unsigned long long)cur.ncand; d[2] = dbg_entries;
			d[3] = (unsigned long long)(dbg_t1 - dbg_t0);
		}
#endif
		if (!have_next) break;
		if (!dbuf && nxt.kind == TK_NORMAL) { stage(nxt, buf ^ 1); __syncwarp(); }
		cur = nxt;
		buf ^= 1;
	}
}

// flags of the one-thread-per-slot kernel: every class of every column is computed and needed
__global__ void __launch_bounds__(256) k_fill16(uint16_t *p, unsigned long long n, uint16_t v)
{
	const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) p[i] = v;
}

} // namespace vo
