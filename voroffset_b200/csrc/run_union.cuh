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

// Slow path of the running union: insert [s, e] into the sorted disjoint list L[0..n) (n >= 2).
// Returns the new length, or -1 when the list would outgrow CAP. Kept out of line (and free of any
// reference to the caller's scalar state) so that the fast-path state stays in registers.
template <int CAP>
__device__ __noinline__ int run_union_insert_list(double2 *L, int n, double s, double e)
{
	int i = 0;
	while (i < n && L[i].y < s) ++i;              // intervals entirely below the candidate
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

// Sorted disjoint closed intervals. One interval lives in registers (the overwhelmingly common case for
// smooth solids); from two on the list lives in the caller-provided array L (local memory). The struct
// only holds scalars and a pointer, so it is scalar-replaced into registers.
template <int CAP>
struct RunUnion {
	double s0, e0;
	int n;
	bool overflow;
	double2 *L;

	__device__ __forceinline__ explicit RunUnion(double2 *list) : s0(0), e0(0), n(0), overflow(false), L(list) {}

	__device__ __forceinline__ void init() { n = 0; overflow = false; }

	__device__ __forceinline__ double2 get(int k) const { return n == 1 ? make_double2(s0, e0) : L[k]; }

	__device__ __forceinline__ void insert(double s, double e)
	{
		if (n == 1) {
			if (s <= e0 && e >= s0) {         // touches or overlaps: coalesce in registers
				s0 = (s < s0) ? s : s0;           // plain compare-select: no NaN handling needed here
				e0 = (e > e0) ? e : e0;
				return;
			}
			if (CAP < 2) { overflow = true; return; }
			if (e < s0) { L[0] = make_double2(s, e); L[1] = make_double2(s0, e0); }
			else        { L[0] = make_double2(s0, e0); L[1] = make_double2(s, e); }
			n = 2;
			return;
		}
		if (n == 0) { s0 = s; e0 = e; n = 1; return; }
		if (overflow) return;
		const int r = run_union_insert_list<CAP>(L, n, s, e);
		if (r < 0) { overflow = true; return; }
		n = r;
		if (n == 1) { s0 = L[0].x; e0 = L[0].y; }
	}
};

} // namespace vo
