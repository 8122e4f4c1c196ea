// Device dexeliser: triangle mesh -> dexel volume (the step before the morphology path).
//
// Restates vor3d::compute_sign (src/vor3d/Dexelize.cpp:166-225) and its helpers orientation /
// point_in_triangle_2d (Dexelize.cpp:56-92, after SDFGen), orient_2d_inexact (Dexelize.cpp:94-113) and
// intersect_ray_z (Dexelize.cpp:136-162). The reference walks the columns one after the other and asks an
// AABB tree for the facets whose bounding box contains the column centre; per (column, facet) pair the test is
// independent, so here the loop is turned inside out: the FACETS are the work items. The columns of a facet's box
// are numbered row by row and cut into chunks of DEX_CELLS; a warp takes a chunk, 32 consecutive cells at a time
// (a facet of a fine mesh covers a handful of columns: one warp, one step; a facet as large as the grid becomes
// thousands of chunks), hits are counted per column, prefix-summed, written on a second identical sweep, and
// every column then sorts its handful of z values (std::sort in the reference, Dexelize.cpp:210). fp64
// throughout, compiled with --fmad=false: every hit carries the reference's operation order.
#pragma once
#include <stdint.h>

namespace vo {

constexpr int DEX_CELLS = 256;   // columns of a facet's box per chunk (one warp: DEX_CELLS / 32 steps)
constexpr int DEX_SORT_REG = 8;  // z values a column sorts in registers; longer lists are sorted in place

struct DexArgs {
	const double *V;              // [3 * nv]
	const int *F;                 // [3 * nf]
	unsigned int nv, nf;
	double ox, oy, spacing;       // column centre = ((x + 0.5) * spacing + ox, (y + 0.5) * spacing + oy)  (CompressedVolumeBase.cpp:5-11)
	int nx, ny;
	int4 *box;                    // [nf] column box (x0, x1, y0, y1) of each facet, x1 < x0: nothing to do
	uint32_t *nchunk;             // [nf]
	const uint32_t *chunk_off;    // [nf + 1] exclusive prefix sum of nchunk
	unsigned long long nchunks;
	uint32_t *cnt;                // [nx * ny] hits per column (count sweep) / write cursor (fill sweep)
	const uint32_t *raw_off;      // [nx * ny + 1] prefix sum of cnt
	double *raw;                  // [raw_off[nx * ny]] unsorted z values
	const uint32_t *out_off;      // [nx * ny + 1] prefix sum of cnt / 2 (intervals)
	double2 *out;                 // the volume's spans
	unsigned int *bad;            // set when a facet names a vertex >= nv
};

// Dexelize.cpp:58-70
__device__ __forceinline__ int dex_orientation(double x1, double y1, double x2, double y2, double &twice_signed_area)
{
	twice_signed_area = y1 * x2 - x1 * y2;
	if (twice_signed_area > 0) return 1;
	if (twice_signed_area < 0) return -1;
	if (y2 > y1) return 1;
	if (y2 < y1) return -1;
	if (x1 > x2) return 1;
	if (x1 < x2) return -1;
	return 0;
}

// Dexelize.cpp:74-92; a, b, c are only normalised when the point is inside
__device__ __forceinline__ bool dex_point_in_triangle(double x0, double y0, double x1, double y1, double x2, double y2,
                                                      double x3, double y3, double &a, double &b, double &c)
{
	x1 -= x0; x2 -= x0; x3 -= x0;
	y1 -= y0; y2 -= y0; y3 -= y0;
	const int signa = dex_orientation(x2, y2, x3, y3, a);
	if (signa == 0) return false;
	const int signb = dex_orientation(x3, y3, x1, y1, b);
	if (signb != signa) return false;
	const int signc = dex_orientation(x1, y1, x2, y2, c);
	if (signc != signa) return false;
	const double sum = a + b + c;
	a /= sum; b /= sum; c /= sum;
	return true;
}

// One thread per facet: the columns whose centre can lie in the facet's xy bounding box, as chunks. Facets whose
// projection has zero signed area never produce a hit (intersect_ray_z returns 0 for them, Dexelize.cpp:150-157).
__global__ void k_dex_plan(DexArgs a)
{
	const unsigned int f = blockIdx.x * blockDim.x + threadIdx.x;
	if (f >= a.nf) return;
	const unsigned int i1 = (unsigned int)a.F[3 * f], i2 = (unsigned int)a.F[3 * f + 1], i3 = (unsigned int)a.F[3 * f + 2];
	int4 box = make_int4(0, -1, 0, -1);
	uint32_t n = 0;
	if (i1 >= a.nv || i2 >= a.nv || i3 >= a.nv) {
		atomicOr(a.bad, 1u);
	} else {
		const double p1x = a.V[3 * i1], p1y = a.V[3 * i1 + 1], p2x = a.V[3 * i2], p2y = a.V[3 * i2 + 1];
		const double p3x = a.V[3 * i3], p3y = a.V[3 * i3 + 1];
		const double det = (p2x - p1x) * (p3y - p1y) - (p2y - p1y) * (p3x - p1x);   // orient_2d_inexact, Dexelize.cpp:103-112
		const double bx0 = fmin(p1x, fmin(p2x, p3x)), bx1 = fmax(p1x, fmax(p2x, p3x));
		const double by0 = fmin(p1y, fmin(p2y, p3y)), by1 = fmax(p1y, fmax(p2y, p3y));
		// centres inside the box: ceil(t0) <= x <= floor(t1) with t = (b - origin) / spacing - 0.5; rounding the other way
		// leaves up to one column of slack on either side, far more than the rounding error of t, and the exact
		// containment test is made per column against the centre itself (k_dex_hits)
		const double fx0 = floor((bx0 - a.ox) / a.spacing - 0.5), fx1 = ceil((bx1 - a.ox) / a.spacing - 0.5);
		const double fy0 = floor((by0 - a.oy) / a.spacing - 0.5), fy1 = ceil((by1 - a.oy) / a.spacing - 0.5);
		if (det != 0 && fx1 >= 0 && fy1 >= 0 && fx0 <= a.nx - 1 && fy0 <= a.ny - 1) {   // (NaN coordinates fail these)
			box.x = (int)fmax(fx0, 0.0); box.y = (int)fmin(fx1, (double)(a.nx - 1));
			box.z = (int)fmax(fy0, 0.0); box.w = (int)fmin(fy1, (double)(a.ny - 1));
			if (box.y >= box.x && box.w >= box.z)
				n = (uint32_t)(((unsigned long long)(box.y - box.x + 1) * (unsigned long long)(box.w - box.z + 1) + DEX_CELLS - 1) / DEX_CELLS);
			else
				box = make_int4(0, -1, 0, -1);
		}
	}
	a.box[f] = box;
	a.nchunk[f] = n;
}

// One warp per chunk. FILL == false counts the hits of every column, FILL == true writes them (same tests, same
// answers); the order of a column's values depends on the atomics and is fixed by k_dex_sort.
template <bool FILL>
__global__ void __launch_bounds__(256) k_dex_hits(DexArgs a)
{
	const unsigned long long w = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (w >= a.nchunks) return;
	const int lane = threadIdx.x & 31;
	// facet of chunk w: the last f with chunk_off[f] <= w. 32-ary search, one probe per lane: four dependent rounds
	// for half a million facets instead of twenty.
	unsigned int lo = 0, hi = a.nf;       // chunk_off[lo] <= w < chunk_off[hi]
	while (hi - lo > 1) {
		const unsigned int step = (hi - lo + 31) / 32;
		const unsigned int probe = lo + (lane + 1) * step;
		const bool le = probe < hi && a.chunk_off[probe] <= w;
		const unsigned int k = __popc(__ballot_sync(0xffffffffu, le));   // chunk_off ascends: the votes are a prefix
		hi = min(hi, lo + (k + 1) * step);
		lo += k * step;
	}
	const unsigned int f = lo;
	const int4 box = a.box[f];
	const unsigned int wd = (unsigned int)(box.y - box.x + 1);
	const unsigned int ncell = wd * (unsigned int)(box.w - box.z + 1);          // <= nx * ny < 2^32 - 2 DEX_CELLS (dexelize_dev)
	const unsigned int base = (unsigned int)(w - a.chunk_off[f]) * DEX_CELLS;
	const int i1 = a.F[3 * f], i2 = a.F[3 * f + 1], i3 = a.F[3 * f + 2];
	const double p1x = a.V[3 * i1], p1y = a.V[3 * i1 + 1], p1z = a.V[3 * i1 + 2];
	const double p2x = a.V[3 * i2], p2y = a.V[3 * i2 + 1], p2z = a.V[3 * i2 + 2];
	const double p3x = a.V[3 * i3], p3y = a.V[3 * i3 + 1], p3z = a.V[3 * i3 + 2];
	// the AABB query of the reference hands compute_sign exactly the facets whose box contains the centre
	// (Dexelize.cpp:190-207: a query box of zero extent in x and y)
	const double bx0 = fmin(p1x, fmin(p2x, p3x)), bx1 = fmax(p1x, fmax(p2x, p3x));
	const double by0 = fmin(p1y, fmin(p2y, p3y)), by1 = fmax(p1y, fmax(p2y, p3y));
	for (int it = 0; it < DEX_CELLS / 32; ++it) {
		if (base + (unsigned int)(it * 32) >= ncell) break;       // (base < ncell < 2^32 - 2 DEX_CELLS: no wrap-around, dexelize_dev checks)
		const unsigned int id = base + (unsigned int)(it * 32 + lane);
		if (id >= ncell) continue;
		const unsigned int row = id / wd;
		const int x = box.x + (int)(id - row * wd), y = box.z + (int)row;
		const double cx = (x + 0.5) * a.spacing + a.ox;
		const double cy = (y + 0.5) * a.spacing + a.oy;
		if (!(cx >= bx0 && cx <= bx1 && cy >= by0 && cy <= by1)) continue;
		double u, v, t;
		if (!dex_point_in_triangle(cx, cy, p1x, p1y, p2x, p2y, p3x, p3y, u, v, t)) continue;
		const unsigned long long c = (unsigned long long)x + (unsigned long long)a.nx * y;
		const uint32_t k = atomicAdd(a.cnt + c, 1u);
		if (FILL) {
			const double z = u * p1z + v * p2z + t * p3z;      // Dexelize.cpp:149
			a.raw[a.raw_off[c] + k] = z / a.spacing;          // Dexelize.cpp:204
		}
	}
}

// cnt -> intervals per column (an odd last crossing - an open surface - is dropped: the CSR holds intervals)
__global__ void k_dex_pairs(const uint32_t *cnt, unsigned long long n, uint32_t *pairs)
{
	const unsigned long long c = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (c < n) pairs[c] = cnt[c] >> 1;
}

// One thread per column: ascending sort of its crossings (Dexelize.cpp:210), pairs out.
__global__ void k_dex_sort(DexArgs a, unsigned long long ncols)
{
	const unsigned long long c = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= ncols) return;
	const uint32_t b = a.raw_off[c], n = a.raw_off[c + 1] - b;
	if (n < 2) return;
	double2 *dst = a.out + a.out_off[c];
	if (n <= DEX_SORT_REG) {
		double z[DEX_SORT_REG];
#pragma unroll
		for (int i = 0; i < DEX_SORT_REG; ++i) z[i] = i < (int)n ? a.raw[b + i] : __longlong_as_double(0x7ff0000000000000ll);
		// odd-even transposition network: fully unrolled, stays in registers
#pragma unroll
		for (int r = 0; r < DEX_SORT_REG; ++r)
#pragma unroll
			for (int i = r & 1; i + 1 < DEX_SORT_REG; i += 2) {
				const double lo = fmin(z[i], z[i + 1]), hi = fmax(z[i], z[i + 1]);
				z[i] = lo; z[i + 1] = hi;
			}
#pragma unroll
		for (int i = 0; i + 1 < DEX_SORT_REG; i += 2)
			if (i + 1 < (int)n) dst[i >> 1] = make_double2(z[i], z[i + 1]);
		return;
	}
	double *r = a.raw + b;
	for (uint32_t i = 1; i < n; ++i) {
		const double v = r[i];
		uint32_t j = i;
		while (j > 0 && r[j - 1] > v) { r[j] = r[j - 1]; --j; }
		r[j] = v;
	}
	for (uint32_t i = 0; i + 1 < n; i += 2) dst[i >> 1] = make_double2(r[i], r[i + 1]);
}

} // namespace vo
