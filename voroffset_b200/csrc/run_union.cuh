// Interval algebra on the device: the running union every gather kernel folds its candidates into.
//
// Replaces the reference's per-line interval helpers (src/vor3d/MorphologyOperators.cpp:15-34
// appendSegment, :117-158 unionSegs, MorphologyOperators.hpp:7-19 addSegmentAtTheEnd): a sorted list of
// disjoint CLOSED intervals; a candidate that touches or overlaps existing intervals coalesces with
// them ("next.start <= prev.end" merges, exactly the reference's rule). Because the union of closed
// intervals is associative and min/max are exact, the result does not depend on the order in which
// candidates arrive - that is what lets a data-parallel gather reproduce the sweep bit for bit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ktrace.cuh"

namespace vo {

// Slot encoding shared by the intermediate volume and the staging buffers:
//   x <= y            one interval [x, y]
//   x = +inf, y = -inf   empty
//   x = NaN-tagged    the list lives in a pool: low 48 bits of x = first index, bits of y = count
__device__ __forceinline__ double2 slot_empty()
{
	return make_double2(__longlong_as_double(0x7FF0000000000000LL), __longlong_as_double((long long)0xFFF0000000000000ULL));
}
__device__ __forceinline__ double2 slot_pool(unsigned long long base, unsigned int n)
{
	return make_double2(__longlong_as_double((long long)(0x7FF8000000000000ULL | base)), __longlong_as_double((long long)n));
}
__device__ __forceinline__ bool slot_is_pool(double2 s)
{
	return (((unsigned int)__double2hiint(s.x)) & 0xFFFF0000u) == 0x7FF80000u;
}
__device__ __forceinline__ unsigned long long slot_pool_base(double2 s)
{
	return ((unsigned long long)__double_as_longlong(s.x)) & 0x0000FFFFFFFFFFFFULL;
}
__device__ __forceinline__ unsigned int slot_pool_count(double2 s)
{
	return (unsigned int)__double_as_longlong(s.y);
}

// Slow path of the running union: insert [s, e] into the sorted disjoint list L[0..n) (n >= 3).
// Returns the new length, or -1 when the list would outgrow CAP. Kept out of line (and free of any
// reference to the caller's scalar state) so that the fast-path state stays in registers.
// proper: every entry that ever went into the list had s <= e, so the upper ends ascend and the first interval that is
// not entirely below the candidate can be found by looking at the last one (candidates tend to arrive in ascending
// runs: the intervals of one input column after the other) and then by bisection, instead of walking up from the
// bottom - the walk was a third of the tile kernel's time on lattices with ~26 intervals per column. (An entry with
// s > e - not an interval; the reference carries those through its unions - can leave the list unsorted: the walk then.)
template <int CAP>
__device__ __noinline__ int run_union_insert_list(double2 *L, int n, double s, double e, bool proper = false)
{
	int i = 0;
	if (proper) {
		if (L[n - 1].y < s) i = n;
		else if (L[n - 2].y < s) i = n - 1;            // (n >= 3) lands on the last interval: the other common case
		else {
			int hi = n - 2;                            // invariant: L[hi].y >= s, everything below i is < s
			while (i < hi) { const int mid = (i + hi) >> 1; if (L[mid].y < s) i = mid + 1; else hi = mid; }
		}
	} else
		while (i < n && L[i].y < s) ++i;          // intervals entirely below the candidate
	if (i == n) {                                  // append
		if (n == CAP) return -1;
		L[n] = make_double2(s, e);
		return n + 1;
	}
	if (L[i].x > e) {                              // falls into a gap: open a new interval at i
		if (n == CAP) return -1;
		for (int k = n; k > i; --k) L[k] = L[k - 1];
		L[i] = make_double2(s, e);
		return n + 1;
	}
	double ns = fmin(s, L[i].x), ne = fmax(e, L[i].y);
	int j = i + 1;
	while (j < n && L[j].x <= e) { ne = fmax(ne, L[j].y); ++j; }
	L[i] = make_double2(ns, ne);
	const int drop = j - i - 1;
	if (drop > 0) {
		for (int k = j; k < n; ++k) L[k - drop] = L[k];
		n -= drop;
	}
	return n;
}

// Sorted disjoint closed intervals. Up to TWO intervals live in registers - one is the overwhelmingly common case
// for smooth solids, two is what every column of an erosion's complement holds (a part below and a part above the
// solid) and what shells give; from three on the list lives in the caller-provided array L (local memory). The
// struct only holds scalars and a pointer, so it is scalar-replaced into registers.
template <int CAP>
struct RunUnion {
	double s0, e0, s1, e1;          // n == 1: (s0, e0); n == 2: (s0, e0) < (s1, e1)
	int n;
	bool overflow;
	bool proper;                    // n >= 3: every entry of L came from candidates with s <= e (run_union_insert_list)
	double2 *L;                     // valid for n >= 3 only: read through get()

	__device__ __forceinline__ explicit RunUnion(double2 *list) : s0(0), e0(0), s1(0), e1(0), n(0), overflow(false), proper(false), L(list) {}

	__device__ __forceinline__ void init() { n = 0; overflow = false; proper = false; }

	__device__ __forceinline__ double2 get(int k) const
	{
		if (n <= 2) return k == 0 ? make_double2(s0, e0) : make_double2(s1, e1);
		return L[k];
	}

	__device__ __forceinline__ void insert(double s, double e)
	{
		if (n == 1) {
			if (s <= e0 && e >= s0) {         // touches or overlaps: coalesce in registers
				s0 = (s < s0) ? s : s0;           // plain compare-select: no NaN handling needed here
				e0 = (e > e0) ? e : e0;
				return;
			}
			if (CAP < 2) { overflow = true; return; }
			if (e < s0) { s1 = s0; e1 = e0; s0 = s; e0 = e; }
			else        { s1 = s; e1 = e; }
			n = 2;
			return;
		}
		if (n == 0) { s0 = s; e0 = e; n = 1; return; }
		if (n == 2 && s <= e) {            // (an entry with s > e is not an interval: the list path below folds it the way it always has)
			const bool t0 = s <= e0 && e >= s0, t1 = s <= e1 && e >= s1;
			if (t0 && t1) { s0 = (s < s0) ? s : s0; e0 = (e > e1) ? e : e1; n = 1; return; }   // bridges the gap
			if (t0) { s0 = (s < s0) ? s : s0; e0 = (e > e0) ? e : e0; return; }
			if (t1) { s1 = (s < s1) ? s : s1; e1 = (e > e1) ? e : e1; return; }
			if (CAP < 3) { overflow = true; return; }
			// a third component: the list moves to L, sorted
			const double2 a = make_double2(s0, e0), b = make_double2(s1, e1), c = make_double2(s, e);
			if (e < s0) { L[0] = c; L[1] = a; L[2] = b; }
			else if (e < s1) { L[0] = a; L[1] = c; L[2] = b; }
			else { L[0] = a; L[1] = b; L[2] = c; }
			n = 3;
			proper = s0 <= e0 && s1 <= e1;
			return;
		}
		if (overflow) return;
		if (n == 2) { L[0] = make_double2(s0, e0); L[1] = make_double2(s1, e1); proper = false; }
		proper = proper && s <= e;
		const int r = run_union_insert_list<CAP>(L, n, s, e, proper);
		if (r < 0) { overflow = true; return; }
		n = r;
		if (n <= 2) { s0 = L[0].x; e0 = L[0].y; if (n == 2) { s1 = L[1].x; e1 = L[1].y; } }
	}
};

} // namespace vo
