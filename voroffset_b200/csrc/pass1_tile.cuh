// 'ours', pass 1, tile kernel: one CTA per (row y, tile of TX output columns).
//
// What it computes is exactly what k_pass1 (kernels.cuh) computes - for every output column x of the row
// and every radius class j the union over |dx| <= reach[j] of the neighbours' intervals grown by
// H[j][|dx|] - but organised for the machine:
//
//  * the row segment [x0-J, x0+TX+J) is staged ONCE in shared memory as a flat candidate array
//    (interval, segment column), instead of every (x, j) thread walking the CSR on its own;
//  * a warp owns one output column at a time and its LANES ARE THE RADIUS CLASSES: a candidate is
//    broadcast from shared memory and lane j adds its own cap Ht[|dx|][j] (conflict-free row of the
//    transposed table), so the 33 classes of R = 32 cost one instruction stream;
//  * candidates that cannot matter are skipped with an EXACT dominance test (below), found 32 at a
//    time with a ballot;
//  * results are collected in a shared-memory tile [j][x] and written to the mid volume as full
//    128-byte lines.
//
// Exact pruning. Let p be an interval of column c and q an interval of the neighbouring column c' that
// is one step closer to the output column x (|c'-x| = d-1, |c-x| = d). For class j the two candidates
// are [a_p - H_j(d), b_p + H_j(d)] and [a_q - H_j(d-1), b_q + H_j(d-1)], and q is alive whenever p is.
// With D_j(d) = H_j(d-1) - H_j(d) > 0, q's candidate contains p's for EVERY class as soon as
//     max(a_q - a_p, b_p - b_q) <= min_j D_j(d),
// and then p can be dropped without changing the union (containment is decided on the same table
// values the candidates are built from; fp64 subtraction is monotone, so containment of the real
// numbers carries over to the rounded endpoints). The host passes Dmono[d] = min over d' >= d and all
// live classes of D_j(d') (non-decreasing in d), so "dominated at distance d" implies "dominated at
// every larger distance" and one threshold byte per candidate and side suffices. Dominance is
// transitive and distances strictly decrease along a chain, so every dropped candidate is contained in
// a kept one. This plays the role of the reference's Voronoi pruning (Voronoi2D.cpp:329-586: a seed is
// retired once its cell no longer reaches the sweep line) but is conservative, branch-light and
// identical for all classes. On the torus workloads it keeps ~7 of ~38 candidates per column (R = 32).
#pragma once
#include "kernels.cuh"

namespace vo {

struct Pass1TileArgs {
	int nx, ny, J, TX, cmax, tiles_x;
	const uint32_t *off;
	const double2 *spans;
	const double *Ht;       // (J+1)*(J+1), transposed cap table: Ht[d*(J+1) + j] = H[j][d] (-1 = out of reach)
	const int *reach;       // J+1
	const double *Dmono;    // J+2: Dmono[d], d = 1..J; Dmono[J+1] = +inf
	double2 *mid;
	double2 *pool;
	unsigned long long *cursor;
	unsigned long long pool_cap;
	Redo redo;              // slot ids to be (re)done by k_pass1: list overflow and oversized tiles
};

__host__ __device__ inline size_t pass1_tile_smem(int J, int TX, int cmax)
{
	const size_t JP = (size_t)J + 1, SEG = (size_t)TX + 2 * J;
	size_t b = 0;
	b += JP * TX * sizeof(double2);          // out tile
	b += (size_t)cmax * sizeof(double2);     // candidates
	b += JP * JP * sizeof(double);           // Ht
	b += (JP + 1) * sizeof(double);          // Dmono
	b += ((SEG + 1 + 1) & ~(size_t)1) * sizeof(uint32_t); // segment offsets (even count keeps alignment)
	b += (size_t)cmax * sizeof(uint32_t);    // per candidate: first surviving output | width << 8 | column << 16
	return b + 16;
}

template <int CAP>
__global__ void __launch_bounds__(512) k_pass1_tile(Pass1TileArgs a)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const int J = a.J, JP = J + 1, TX = a.TX, SEG = TX + 2 * J;
	double2 *s_out = reinterpret_cast<double2 *>(smem_raw);
	double2 *s_cand = s_out + (size_t)JP * TX;
	double *s_Ht = reinterpret_cast<double *>(s_cand + a.cmax);
	double *s_D = s_Ht + (size_t)JP * JP;
	uint32_t *s_off = reinterpret_cast<uint32_t *>(s_D + JP + 1);
	uint32_t *s_sv = s_off + ((SEG + 2) & ~1);   // per candidate: first surviving output | width << 8 | column << 16

	const int tid = threadIdx.x, nthr = blockDim.x;
	const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
	const int y = blockIdx.x / a.tiles_x;
	const int x0 = (blockIdx.x % a.tiles_x) * TX;
	const int txe = min(TX, a.nx - x0);
	const size_t rowbase = (size_t)y * a.nx;

	// ---- phase 0: stage the row segment -------------------------------------------------------
	for (int i = tid; i <= SEG; i += nthr) {
		const int gx = min(max(x0 - J + i, 0), a.nx);      // columns outside the grid collapse to empty ranges
		s_off[i] = __ldg(a.off + rowbase + gx);
	}
	__syncthreads();
	const uint32_t base = s_off[0];
	const int ncand = (int)(s_off[SEG] - base);
	if (ncand > a.cmax) {                                   // oversized segment: leave the tile to k_pass1
		for (int idx = tid; idx < JP * txe; idx += nthr) {
			const int j = idx / txe, xi = idx % txe;
			redo_push(a.redo, ((unsigned long long)y * JP + j) * a.nx + x0 + xi);
		}
		return;
	}
	if (ncand == 0) {                                       // nothing in reach: the whole tile is empty
		const double2 e = slot_empty();
		for (int idx = tid; idx < JP * txe; idx += nthr) {
			const int j = idx / txe, xi = idx % txe;
			a.mid[((size_t)y * JP + j) * a.nx + x0 + xi] = e;
		}
		return;
	}
	for (int i = tid; i < JP * JP; i += nthr) s_Ht[i] = __ldg(a.Ht + i);
	for (int i = tid; i < JP + 1; i += nthr) s_D[i] = __ldg(a.Dmono + i);
	for (int k = tid; k < ncand; k += nthr) s_cand[k] = __ldg(a.spans + base + k);
	for (int i = tid; i < SEG; i += nthr)
		for (uint32_t k = s_off[i] - base; k < s_off[i + 1] - base; ++k) s_sv[k] = (uint32_t)i << 16;
	__syncthreads();

	// ---- phase 1: dominance thresholds ----------------------------------------------------------
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
		need_hi += m;
		need_lo += m;
		// first d in [1, J] with Dmono[d] >= need (Dmono is non-decreasing, Dmono[J+1] = +inf)
		int lo = 1, hi = JP;
		while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_D[mid] >= need_hi) hi = mid; else lo = mid + 1; }
		const int t_hi = lo;           // dominated for outputs at distance >= t_hi on the left of the candidate
		lo = 1; hi = JP;
		while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_D[mid] >= need_lo) hi = mid; else lo = mid + 1; }
		const int t_lo = lo;           // ... on the right
		// the candidate survives for the outputs ix in [i - (t_hi-1), i + (t_lo-1)] (segment coordinates);
		// packed as first | width << 8 | column << 16
		const int first = max(i - (t_hi - 1), 0);
		const int last = min(i + (t_lo - 1), SEG - 1);
		s_sv[k] = (uint32_t)first | ((uint32_t)(last - first) << 8) | ((uint32_t)i << 16);
	}
	__syncthreads();

	// ---- phase 2: one warp per output column, lanes = radius classes ----------------------------
	double2 ulist[CAP];
	for (int jbase = 0; jbase <= J; jbase += 32) {
		const int j = jbase + lane;
		const bool active = j <= J;
		const int Xj = active ? __ldg(a.reach + j) : -1;
		const int Xmax = __shfl_sync(0xffffffffu, Xj, 0);       // classes are ordered by decreasing reach
		const double *Htj = s_Ht + (active ? j : 0);
		for (int xi = warp; xi < txe; xi += nwarp) {
			const int ix = xi + J;
			// candidates within the largest reach of this round of classes
			const int kb = (int)(s_off[ix - Xmax] - base), ke = (int)(s_off[ix + Xmax + 1] - base);
			RunUnion<CAP> u(ulist);
			for (int kk = kb; kk < ke; kk += 32) {
				const int k = kk + lane;
				bool sv = false;
				if (k < ke) {
					const uint32_t w = s_sv[k];
					sv = (uint32_t)(ix - (int)(w & 0xffu)) <= ((w >> 8) & 0xffu);
				}
				unsigned m = __ballot_sync(0xffffffffu, sv);
				while (m) {
					const int k2 = kk + __ffs(m) - 1;
					m &= m - 1;
					const int d = abs((int)(s_sv[k2] >> 16) - ix);
					if (d <= Xj) {
						const double2 ab = s_cand[k2];
						const double h = Htj[d * JP];
						u.insert(ab.x - h, ab.y + h);
					}
				}
			}
			if (active) {
				double2 out;
				if (u.overflow) {
					redo_push(a.redo, ((unsigned long long)y * JP + j) * a.nx + x0 + xi);
					out = slot_empty();
				} else if (u.n == 0) out = slot_empty();
				else if (u.n == 1) out = make_double2(u.s0, u.e0);
				else {
					const unsigned long long pb = atomicAdd(a.cursor, (unsigned long long)u.n);
					if (pb + u.n <= a.pool_cap)
						for (int q = 0; q < u.n; ++q) a.pool[pb + q] = u.L[q];
					out = slot_pool(pb, (unsigned int)u.n);
				}
				s_out[(size_t)j * TX + xi] = out;
			}
		}
	}
	__syncthreads();

	// ---- phase 3: coalesced store of the tile ---------------------------------------------------
	for (int idx = tid; idx < JP * TX; idx += nthr) {
		const int j = idx / TX, xi = idx % TX;
		if (xi < txe) a.mid[((size_t)y * JP + j) * a.nx + x0 + xi] = s_out[idx];
	}
}

} // namespace vo
