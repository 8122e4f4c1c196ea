// Gather kernels of the dexel-morphology hot path (sm_100a).
//
// Data layout in HBM
//   dexel volume   : CSR, off[nx*ny+1] (uint32) + spans[M] (double2 = (z1,z2)), column (x,y) = x + nx*y.
//   mid volume     : dense slots mid[(y*(J+1) + j)*nx + x] (double2), j = |dy| radius class, J = floor(R);
//                    a slot is one interval, empty, or a reference into the mid pool (run_union.cuh).
//   staged output  : cnt[c] + 2 inline slots per list + a pool for longer lists; k_compact turns it into
//                    canonical CSR after the prefix sum over cnt.
//   cap tables     : host-computed doubles (vo_lib.cu: make_tables) so that every cap height carries the
//                    reference's exact fp64 operation order, independent of GPU sqrt / FMA behaviour.
#pragma once
#include <type_traits>

#include "run_union.cuh"
#include "scan.cuh"

namespace vo {

constexpr int CAP_FAST = 32;    // running-union capacity of the first launch
constexpr int CAP_BIG = 512;    // capacity of the redo launch for lists that outgrew CAP_FAST
constexpr int CAP_HUGE = 32768; // ... of the last-resort redo launch (a call that met a list beyond CAP_BIG is repeated with it): the
                                // lists live in a global-memory scratch, one slice per thread of a small fixed grid (HUGE_GRID)
constexpr int STAGE_INLINE = 2; // inline slots per staged list

struct Stage {
	uint32_t *cnt;                 // [nlists]
	double2 *inl;                  // [nlists * STAGE_INLINE]
	double2 *pool;                 // [pool_cap]
	unsigned long long *cursor;    // pool allocation cursor (keeps counting past pool_cap)
	unsigned long long pool_cap;
};

struct Redo {
	unsigned long long *list;      // ids of lists whose running union overflowed CAP_FAST
	unsigned int *count;
	unsigned int cap;
};

// Which lists a gather launch processes. First launch: ids 0..n-1 (list == NULL). Redo launch: the ids the
// first launch pushed to the redo list; their number is only known on the device, so the launch uses a
// fixed grid, reads *count itself and strides over the list - the host never has to synchronise in between.
struct Work {
	const unsigned long long *list;
	unsigned long long n;
	const unsigned int *count;     // device-side length of `list` (redo launch), capped by cap
	unsigned int cap;
	unsigned int *fail;            // redo launch: incremented when an item outgrows CAP_BIG as well
	double2 *huge = nullptr;       // last-resort redo launch (CAP_HUGE): CAP_HUGE entries per thread of the grid
};

// Where the running union of a gather item keeps its list from the third interval on: local memory, or - last-resort
// launch - this thread's slice of the global scratch.
template <int CAP>
struct ListStore {
	double2 a[CAP];
	__device__ __forceinline__ double2 *ptr(const Work &) { return a; }
};
template <>
struct ListStore<2> {           // (a running union of capacity 2 lives in registers: no list at all)
	__device__ __forceinline__ double2 *ptr(const Work &) { return nullptr; }
};
template <>
struct ListStore<CAP_HUGE> {
	__device__ __forceinline__ double2 *ptr(const Work &wk) { return wk.huge + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * CAP_HUGE; }
};

// CAP == CAP_FAST is the first launch (one id per thread, no loop); CAP == CAP_BIG is the redo launch.
#define VO_FOR_WORK(CAPV, wk, c)                                                                                 \
	const unsigned long long vo_n_ = (CAPV) != CAP_FAST ? (unsigned long long)min(*(wk).count, (wk).cap) : (wk).n; \
	for (unsigned long long vo_t_ = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x, c = 0;           \
	     vo_t_ < vo_n_ && ((c = (CAPV) != CAP_FAST ? (wk).list[vo_t_] : vo_t_), true);                          \
	     vo_t_ = (CAPV) != CAP_FAST ? vo_t_ + (unsigned long long)gridDim.x * blockDim.x : vo_n_)

template <int CAP>
__device__ __forceinline__ void stage_emit(const Stage &st, size_t c, const RunUnion<CAP> &u)
{
	st.cnt[c] = (uint32_t)u.n;
	if (u.n <= STAGE_INLINE) {
		for (int k = 0; k < u.n; ++k) st.inl[c * STAGE_INLINE + k] = u.get(k);
	} else {
		unsigned long long base = atomicAdd(st.cursor, (unsigned long long)u.n);
		if (base + u.n <= st.pool_cap)
			for (int k = 0; k < u.n; ++k) st.pool[base + k] = u.get(k);
		st.inl[c * STAGE_INLINE] = slot_pool(base, (unsigned int)u.n);
	}
}

__device__ __forceinline__ void redo_push(const Redo &rd, unsigned long long id)
{
	unsigned int k = atomicAdd(rd.count, 1u);
	if (k < rd.cap) rd.list[k] = id;
}

// a list outgrew its running-union capacity: first launch -> redo list, redo launch -> failure counter
__device__ __forceinline__ void overflow_item(const Work &wk, const Redo &rd, unsigned long long id)
{
	if (wk.count) atomicAdd(wk.fail, 1u);   // redo launch
	else redo_push(rd, id);
}

// ---------------------------------------------------------------------------------------------------
// 'ours', pass 1 (x-direction). One thread per (x, y, j): the union over |dx| <= reach[j] of the
// intervals of column (x+dx, y) grown by H[j][|dx|] = sqrt(r1(j)^2 - dx^2), r1(j) = sqrt(R^2 - j^2).
// This is the per-slice 2D dilation the reference's VoronoiMorpho2D sweep computes
// (src/vor3d/Voronoi2D.cpp:591-739 driven by halfDilate, HalfDilationOperator.cpp:6-29), evaluated
// for every radius class so that pass 2 needs no arithmetic at all.
// Lanes run along x: offsets and spans of neighbouring columns are read coalesced.
// ---------------------------------------------------------------------------------------------------
struct Pass1Args {
	int nx, ny, J;
	const uint32_t *off;
	const double2 *spans;
	const double *H;        // (J+1)*(J+1), [j][|dx|]
	const int *reach;       // J+1
	double2 *mid;
	double2 *pool;
	unsigned long long *cursor;
	unsigned long long pool_cap;
	Redo redo;
	Work wk;
};

template <int CAP>
__device__ __forceinline__ void pass1_item(const Pass1Args &a, unsigned long long slot)
{
	const int x = (int)(slot % (unsigned)a.nx);
	const unsigned long long rest = slot / (unsigned)a.nx;
	const int j = (int)(rest % (unsigned)(a.J + 1));
	const int y = (int)(rest / (unsigned)(a.J + 1));

	ListStore<CAP> ulist;
	RunUnion<CAP> u(ulist.ptr(a.wk));
	const int X = a.reach[j];
	const int lo = max(-X, -x), hi = min(X, a.nx - 1 - x);
	const double *Hrow = a.H + (size_t)j * (a.J + 1);
	const size_t c0 = (size_t)y * a.nx + x;
	uint32_t o0 = __ldg(a.off + c0 + lo);
	for (int dx = lo; dx <= hi; ++dx) {
		const uint32_t o1 = __ldg(a.off + c0 + dx + 1);
		if (o1 > o0) {
			const double h = __ldg(Hrow + abs(dx));
			for (uint32_t k = o0; k < o1; ++k) {
				const double2 v = __ldg(a.spans + k);
				u.insert(v.x - h, v.y + h);
			}
		}
		o0 = o1;
	}
	double2 out;
	if (u.overflow) { overflow_item(a.wk, a.redo, slot); out = slot_empty(); }
	else if (u.n == 0) out = slot_empty();
	else if (u.n == 1) out = make_double2(u.s0, u.e0);
	else {
		unsigned long long base = atomicAdd(a.cursor, (unsigned long long)u.n);
		if (base + u.n <= a.pool_cap)
			for (int k = 0; k < u.n; ++k) a.pool[base + k] = u.get(k);
		out = slot_pool(base, (unsigned int)u.n);
	}
	a.mid[slot] = out;
}

template <int CAP>
__global__ void __launch_bounds__(128) k_pass1(Pass1Args a)
{
	KT_SCOPE(KT_PASS1, 0, threadIdx.x == 0);
	VO_FOR_WORK(CAP, a.wk, slot) pass1_item<CAP>(a, slot);
}

// ---------------------------------------------------------------------------------------------------
// 'ours', pass 2 (y-direction). One thread per output column (x, y): the plain union of class |dy| of
// column (x, y+dy) for |dy| <= J. Replaces the SeparatePowerMorpho2D sweep + unionMap
// (src/vor3d/SeparatePower2D.cpp:215-341, HalfDilationOperator.hpp:5-16). Lanes run along x, so each
// (dy) step is one coalesced 512-byte row segment of the mid volume.
// ---------------------------------------------------------------------------------------------------
struct Pass2Args {
	int nx, ny, J;          // grid of the mid volume
	int y0, y1;             // rows produced
	const double2 *mid;
	const uint16_t *flags;  // [2][ny*nx] per mid column: lo | hi << 8; the classes lo <= j < hi of flags[0][c] are
	                        // needed by the consumer rows above (y - j), those of flags[1][c] by the rows below
	                        // (y + j); every other slot was never written by pass 1 and must not be read
	                        // (pass1_tile.cuh)
	const unsigned long long *tilemask;   // [2][ny * ceil(nx / 128)]: OR of the class windows (bit j = class j) of the
	                        // columns of a pass-1 tile, consumers above ([0]) / below ([1]); NULL-free only for J <= 63
	const double2 *pool;
	unsigned long long pool_cap = ~0ull;   // entries of `pool`. The banded host-buffer call and the slab step run pass 2 before the
	                        // host has seen pass 1's counters: a pass 1 that outgrew its pool leaves references beyond
	                        // the end (the host then discards the result and repeats with a larger pool), which must
	                        // not be followed
	Stage st;
	Redo redo;
	Work wk;
	// dual form only (k_pass2_rows_dual): distance of every column to the nearest EMPTY column of its row (k_empty_dist),
	// the reach of every class, and the range negateInv keeps (zmin, zmax of the erosion)
	const uint8_t *dist = nullptr;
	const uint8_t *dist_tmin = nullptr;   // ... and the smallest of them per pass-2 tile and row (k_empty_dist)
	const int *reach = nullptr;
	double lo = 0, hi = 0;
};

// class window flag (lo | hi << 8): is class j needed? FLAG_ALL = every class (written by the one-thread-per-slot pass 1
// when floor(R) + 1 does not fit the 8-bit window bound; lo = hi = 255 never occurs as a window, empty ones are 0).
constexpr uint16_t FLAG_ALL = 0xFFFFu;
__device__ __forceinline__ bool flag_has(uint16_t w, int j) { return w == FLAG_ALL || ((int)(w & 0xffu) <= j && j < (int)(w >> 8)); }
// ... where the sentinel cannot occur (the row kernels: floor(R) <= 63): lo <= j < hi as one unsigned compare
__device__ __forceinline__ bool flag_in(uint16_t w, int j) { const unsigned int lo = w & 0xffu; return (unsigned int)j - lo < (unsigned int)(w >> 8) - lo; }

template <int CAP>
__device__ __forceinline__ void pass2_take(RunUnion<CAP> &u, const double2 *slot, const double2 *pool, unsigned long long pool_cap)
{
	const double2 s = __ldg(slot);
	if (s.x <= s.y) u.insert(s.x, s.y);
	else if (slot_is_pool(s)) {
		const unsigned long long base = slot_pool_base(s);
		const unsigned int n = slot_pool_count(s);
		if (base + n > pool_cap) return;
		for (unsigned int k = 0; k < n; ++k) {
			const double2 v = __ldg(pool + base + k);
			u.insert(v.x, v.y);
		}
	}
}

// The slots of output column (x, y) named by two bit masks (bit j-1 of m_up: class j of row y-j, of m_dn: class j
// of row y+j; self: class 0 of the own row), fetched four at a time (independent loads) and folded into the union.
// bit scans of the class masks: 32-bit words where floor(R) <= 32 (half the instructions of the 64-bit ones)
__device__ __forceinline__ int mask_ffs(unsigned int m) { return __ffs((int)m); }
__device__ __forceinline__ int mask_ffs(unsigned long long m) { return __ffsll((long long)m); }

template <int CAP, typename M = unsigned long long>
__device__ __forceinline__ void pass2_gather(const Pass2Args &a, RunUnion<CAP> &u, int x, int y, M m_up, M m_dn, bool self_needed)
{
	const size_t nx = (size_t)a.nx, midrow = (size_t)(a.J + 1) * nx;
	const double2 *self = a.mid + (size_t)y * midrow + x;
	if (self_needed) pass2_take(u, self, a.pool, a.pool_cap);
	const size_t step_up = midrow - nx, step_dn = midrow + nx;     // slot (y-j, class j) = self - j*step_up, ...
	while (m_up | m_dn) {
		const double2 *p[4];
		int n = 0;
#pragma unroll
		for (int i = 0; i < 4; ++i) {
			p[i] = self;
			if (m_up) { const int j = mask_ffs(m_up); m_up &= m_up - 1; p[i] = self - (size_t)j * step_up; n = i + 1; }
			else if (m_dn) { const int j = mask_ffs(m_dn); m_dn &= m_dn - 1; p[i] = self + (size_t)j * step_dn; n = i + 1; }
		}
		double2 v[4];
#pragma unroll
		for (int i = 0; i < 4; ++i) v[i] = __ldg(p[i]);
#pragma unroll
		for (int i = 0; i < 4; ++i) {
			if (i < n) {
				if (v[i].x <= v[i].y) u.insert(v[i].x, v[i].y);
				else if (slot_is_pool(v[i])) {
					const unsigned long long base = slot_pool_base(v[i]);
					const unsigned int cnt = slot_pool_count(v[i]);
					if (base + cnt <= a.pool_cap)
						for (unsigned int k = 0; k < cnt; ++k) { const double2 w = __ldg(a.pool + base + k); u.insert(w.x, w.y); }
				}
			}
		}
	}
}

template <int CAP>
__device__ __forceinline__ void pass2_item(const Pass2Args &a, unsigned long long c)   // c: output list, rows relative to y0
{
	const int x = (int)(c % (unsigned)a.nx);
	const int y = a.y0 + (int)(c / (unsigned)a.nx);

	ListStore<CAP> ulist;
	RunUnion<CAP> u(ulist.ptr(a.wk));
	const int up = min(a.J, y), dn = min(a.J, a.ny - 1 - y);    // rows available above / below
	const size_t nx = (size_t)a.nx, midrow = (size_t)(a.J + 1) * nx;
	const uint16_t *f_up = a.flags, *f_dn = a.flags + (size_t)a.ny * nx;
	const size_t cc = (size_t)y * nx + x;
	const double2 *self = a.mid + (size_t)y * midrow + x;
	if (a.J <= 63) {
		// which rows are needed? All flags are loaded back to back (independent loads) into two bit masks:
		// bit j-1 of m_up = row y-j is needed (this output lies BELOW that row: its "dn" window decides),
		// bit j-1 of m_dn = row y+j is needed (its "up" window decides).
		unsigned long long m_up = 0, m_dn = 0;
		{
			const uint16_t *f = f_dn + cc - nx;
#pragma unroll 8
			for (int j = 1; j <= up; ++j, f -= nx) m_up |= (unsigned long long)flag_has(__ldg(f), j) << (j - 1);
			f = f_up + cc + nx;
#pragma unroll 8
			for (int j = 1; j <= dn; ++j, f += nx) m_dn |= (unsigned long long)flag_has(__ldg(f), j) << (j - 1);
		}
		pass2_gather(a, u, x, y, m_up, m_dn, flag_has(__ldg(f_up + cc), 0) || flag_has(__ldg(f_dn + cc), 0));
	} else {
		// general form (more than 64 classes: every class of every column was computed by k_pass1)
		const uint16_t *f = f_dn + (size_t)(y - up) * nx + x;
		const double2 *row = a.mid + (size_t)(y - up) * midrow + x;
		for (int j = up; j >= 1; --j, f += nx, row += midrow)
			if (flag_has(__ldg(f), j)) pass2_take(u, row + (size_t)j * nx, a.pool, a.pool_cap);
		if (flag_has(__ldg(f_up + cc), 0) || flag_has(__ldg(f_dn + cc), 0)) pass2_take(u, self, a.pool, a.pool_cap);
		f = f_up + (size_t)(y + 1) * nx + x;
		row = a.mid + (size_t)(y + 1) * midrow + x;
		for (int j = 1; j <= dn; ++j, f += nx, row += midrow)
			if (flag_has(__ldg(f), j)) pass2_take(u, row + (size_t)j * nx, a.pool, a.pool_cap);
	}
	if (u.overflow) { overflow_item(a.wk, a.redo, c); a.st.cnt[c] = 0; return; }
	stage_emit(a.st, (size_t)c, u);
}

template <int CAP>
__global__ void __launch_bounds__(128) k_pass2(Pass2Args a)
{
	KT_SCOPE(KT_PASS2, a.y0, threadIdx.x == 0);
	VO_FOR_WORK(CAP, a.wk, c) pass2_item<CAP>(a, c);
}

// Row-segment form of pass 2 (J <= 63): a CTA owns the P2_TX = P1_TX output columns of one tile of pass 1 in one
// row. Pass 1 leaves, per tile and side, the OR of the class windows of its columns (`tilemask`); a producer row
// whose tile mask lacks class j cannot concern any consumer of the segment at row distance j. The lanes of a warp
// test the 2J tile masks in parallel (two rows per lane), the ballots name the ~10 candidate rows, and only for
// those are the per-column windows read - instead of 2J+1 flags per consumer.
constexpr int P2_TX = 128;
// Resident CTAs per SM the row-segment kernels are compiled for: 8 (64 registers) where the running unions may grow
// lists (the spills of a tighter budget cost more than the warps give: C3 pass 2 0.225 -> 0.255 ms), P2_SHALLOW = 12
// (40 registers, ~230 bytes of spills) where the mid slots hold one interval each - height-field-like input, the dual
// form: the gather is bound by memory latency, C5 k_pass2_rows 0.197 -> 0.174 ms, 14 and 16 give no more.
constexpr int P2_DEEP = 8, P2_SHALLOW = 12;
#ifndef P2_REGONLY_V
#define P2_REGONLY_V 12
#endif
constexpr int P2_REGONLY = P2_REGONLY_V;   // CTAs per SM of the register-only variant (running union of capacity 2, vo_lib.cu: pass2())

// WIDE = false: floor(R) <= 32, the class masks are 32-bit words.
// The union of output column (tile * P2_TX + threadIdx.x, y) into `u`; every thread of the CTA calls it (warp ballots),
// false = no such column (beyond the end of the row).
template <int CAP, bool WIDE>
__device__ __forceinline__ bool pass2_rows_union(const Pass2Args &a, int tiles_x, int tile, int y, RunUnion<CAP> &u)
{
	typedef typename std::conditional<WIDE, unsigned long long, unsigned int>::type M;
	const int x = tile * P2_TX + (int)threadIdx.x;
	const int lane = threadIdx.x & 31;
	const size_t nx = (size_t)a.nx;
	const unsigned long long *t_up = a.tilemask, *t_dn = a.tilemask + (size_t)a.ny * tiles_x;
	// coarse: bit j-1 of c_up = the tile of row y-j holds a column whose "dn" window has class j, ...
	M c_up = 0, c_dn = 0;
#pragma unroll
	for (int h = 0; h < (WIDE ? 2 : 1); ++h) {
		const int j = lane + 1 + 32 * h;
		bool pu = false, pd = false;
		if (j <= a.J) {
			if (y - j >= 0) pu = (__ldg(t_dn + (size_t)(y - j) * tiles_x + tile) >> j) & 1ull;
			if (y + j < a.ny) pd = (__ldg(t_up + (size_t)(y + j) * tiles_x + tile) >> j) & 1ull;
		}
		c_up |= (M)__ballot_sync(0xffffffffu, pu) << (32 * h);
		c_dn |= (M)__ballot_sync(0xffffffffu, pd) << (32 * h);
	}
	if (x >= a.nx) return false;
	const uint16_t *f_up = a.flags, *f_dn = a.flags + (size_t)a.ny * nx;
	const size_t cc = (size_t)y * nx + x;
	// windows of the candidate rows, four independent loads at a time (the masks are warp-uniform)
	M m_up = 0, m_dn = 0;
	const uint16_t w_up = __ldg(f_up + cc), w_dn = __ldg(f_dn + cc);
	while (c_up | c_dn) {
		int jj[4];
		const uint16_t *p[4];
#pragma unroll
		for (int i = 0; i < 4; ++i) {
			jj[i] = 0; p[i] = f_up + cc;
			if (c_up) { const int j = mask_ffs(c_up); c_up &= c_up - 1; jj[i] = -j; p[i] = f_dn + cc - (size_t)j * nx; }
			else if (c_dn) { const int j = mask_ffs(c_dn); c_dn &= c_dn - 1; jj[i] = j; p[i] = f_up + cc + (size_t)j * nx; }
		}
		uint16_t w[4];
#pragma unroll
		for (int i = 0; i < 4; ++i) w[i] = __ldg(p[i]);
#pragma unroll
		for (int i = 0; i < 4; ++i) {
			if (jj[i] < 0) m_up |= (M)flag_in(w[i], -jj[i]) << (-jj[i] - 1);
			else if (jj[i] > 0) m_dn |= (M)flag_in(w[i], jj[i]) << (jj[i] - 1);
		}
	}
	pass2_gather<CAP, M>(a, u, x, y, m_up, m_dn, flag_in(w_up, 0) || flag_in(w_dn, 0));
	return true;
}

template <int CAP, bool WIDE = true, int MINB = P2_DEEP>
__global__ void __launch_bounds__(P2_TX, MINB) k_pass2_rows(Pass2Args a)
{
	KT_SCOPE(KT_PASS2_ROWS, a.y0, threadIdx.x == 0);
	const int tiles_x = (a.nx + P2_TX - 1) / P2_TX;
	const int tile = (int)(blockIdx.x % (unsigned)tiles_x);
	const int y = a.y0 + (int)(blockIdx.x / (unsigned)tiles_x);
	ListStore<CAP> ulist;
	RunUnion<CAP> u(ulist.ptr(a.wk));
	if (!pass2_rows_union<CAP, WIDE>(a, tiles_x, tile, y, u)) return;
	const unsigned long long c = (unsigned long long)(y - a.y0) * a.nx + tile * P2_TX + threadIdx.x;
	if (u.overflow) { redo_push(a.redo, c); a.st.cnt[c] = 0; return; }
	stage_emit(a.st, (size_t)c, u);
}

// (Tried in round 2 and dropped: the same pass writing canonical CSR itself - CTAs taking tickets in list order and the
// prefix sum of their interval counts running across them as a decoupled look-back while the unions are still in
// registers. Bit-identical, but 32768 small tiles of very unequal cost wait for each other in ticket order: 0.575 ms
// against 0.199 + 0.040 ms for this kernel plus k_scan_compact at C5.)

// ---------------------------------------------------------------------------------------------------
// Erosion in DUAL form, for volumes with at most one interval [a, b] per column strictly inside the bounds
// (vo_lib.cu: erode_dual). The reference erodes by dilating the complement (Voronoi.cpp:8-17): the column's complement
// is [zmin-1, a] and [b, zmax+1] (negate_ray, MorphologyOperators.cpp:230-259), an empty column's - and the one-line
// border's - is everything. In the dilated complement of an output column all lower parts share the point zmin-1
// and all upper parts zmax+1, so it is [.., L] u [U, ..] with L = max (a_q + h_q), U = min (b_q - h_q) over the same
// (column, cap) pairs a dilation visits, ONE interval when U <= L or when an empty column lies in reach; negateInv
// (negate_ray_range, MorphologyOperators.cpp:282-315) then keeps [L, U] clamped to [lo, hi]. Pass 1 (the tile kernel,
// DUAL) has computed, per class, the hull of the MIRRORED contributions [-a_q - h_q, -b_q + h_q]: the same table
// values and the same IEEE operations up to sign, so L and U carry the reference's bits. Nothing is complemented,
// no border is added, no list is ever built.
//
// k_empty_dist: per row, the distance (0 .. 255) of every column to the nearest empty column, the two virtual
// columns x = -1 and x = nx (the border) counting as empty. One CTA per row: occupancy bit mask of the row in shared
// memory, then a word-wise search to either side.
// ---------------------------------------------------------------------------------------------------
constexpr int ED_THREADS = 256;
// tmin [ny * ceil(nx / P2_TX)]: the smallest distance inside every pass-2 tile of the row - k_pass2_rows_dual skips its
// per-column scan of the 2 floor(R) + 1 rows when no tile above or below comes within reach (the inside of a solid)
// (shared memory: ceil(nx / 32) occupancy words + ceil(nx / P2_TX) tile minima)
__global__ void __launch_bounds__(ED_THREADS) k_empty_dist(const uint32_t *__restrict__ off, int nx, uint8_t *__restrict__ dist, uint8_t *__restrict__ tmin)
{
	extern __shared__ uint32_t s_occ[];                     // bit x = column x is EMPTY
	const int y = blockIdx.x, nwords = (nx + 31) >> 5, ntx = (nx + P2_TX - 1) / P2_TX;
	uint32_t *s_tmin = s_occ + nwords;
	const uint32_t *row = off + (size_t)y * nx;
	for (int i = threadIdx.x; i < ntx; i += ED_THREADS) s_tmin[i] = 255u;
	for (int x0 = 0; x0 < nwords * 32; x0 += ED_THREADS) {
		const int x = x0 + (int)threadIdx.x;
		const bool empty = x < nx && __ldg(row + x + 1) == __ldg(row + x);
		const unsigned int b = __ballot_sync(0xffffffffu, empty);
		if ((threadIdx.x & 31) == 0 && (x >> 5) < nwords) s_occ[x >> 5] = b;
	}
	__syncthreads();
	for (int x0 = 0; x0 < nx; x0 += ED_THREADS) {           // (whole warps stay in the loop: the tile minimum is a warp reduction)
		const int x = x0 + (int)threadIdx.x;
		int best = 255;
		if (x < nx) {
			best = min(min(x + 1, nx - x), 255);
			// nearest empty column at or below x
			int wi = x >> 5;
			uint32_t bits = s_occ[wi] & (0xffffffffu >> (31 - (x & 31)));
			while (true) {
				if (bits) { best = min(best, x - (wi * 32 + 31 - __clz(bits))); break; }
				if (wi == 0 || x - (wi * 32 - 1) >= best) break;
				bits = s_occ[--wi];
			}
			// ... above x
			wi = x >> 5;
			bits = s_occ[wi] & (0xffffffffu << (x & 31));
			while (true) {
				if (bits) { best = min(best, wi * 32 + __ffs(bits) - 1 - x); break; }
				if (wi + 1 >= nwords || (wi + 1) * 32 - x >= best) break;
				bits = s_occ[++wi];
			}
			dist[(size_t)y * nx + x] = (uint8_t)best;
		}
		// (a warp's 32 consecutive columns start at a multiple of 32: one tile)
		const unsigned int wmin = __reduce_min_sync(0xffffffffu, (unsigned int)best);
		if ((threadIdx.x & 31) == 0 && (x0 + (int)(threadIdx.x & ~31u)) < nx) atomicMin(s_tmin + (x0 + (int)(threadIdx.x & ~31u)) / P2_TX, wmin);
	}
	__syncthreads();
	for (int i = threadIdx.x; i < ntx; i += ED_THREADS) tmin[(size_t)y * ntx + i] = (uint8_t)s_tmin[i];
}

// Does a volume qualify for the dual form? *flag is raised when a column holds several intervals or an interval is not
// strictly inside (lo, hi) (or when the host already knows a reason, init_bad). The single-GPU path lets k_thresh find
// out while it runs; the ranks of a multi-GPU group have to AGREE before they exchange anything, so they check first.
__global__ void __launch_bounds__(256) k_dual_check(const uint32_t *__restrict__ off, const double2 *__restrict__ spans, unsigned long long ncols,
                                                    double lo, double hi, unsigned int init_bad, unsigned int *flag)
{
	const unsigned long long c = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (c == 0 && init_bad) *flag = 1u;
	if (c >= ncols) return;
	const uint32_t o0 = __ldg(off + c), o1 = __ldg(off + c + 1);
	if (o1 == o0) return;
	bool bad = o1 - o0 > 1u;
	if (!bad) { const double2 p = __ldg(spans + o0); bad = !(p.x > lo && p.x <= p.y && p.y < hi); }
	if (bad) *flag = 1u;
}

// Pass 2 of the dual form: same segment / tile-mask logic as k_pass2_rows; the fold is a hull (no running union), an
// output column with an empty column (or the border) in reach is empty, the rest goes through negateInv's clamping.
template <bool WIDE = true>                             // WIDE = false: floor(R) <= 32, class masks in 32-bit words (as k_pass2_rows)
__global__ void __launch_bounds__(P2_TX, P2_SHALLOW) k_pass2_rows_dual(Pass2Args a)
{
	typedef typename std::conditional<WIDE, unsigned long long, unsigned int>::type M;
	__shared__ int s_reach[64];
	if (threadIdx.x <= (unsigned)a.J) s_reach[threadIdx.x] = __ldg(a.reach + threadIdx.x);
	__syncthreads();
	const int tiles_x = (a.nx + P2_TX - 1) / P2_TX;
	const int tile = (int)(blockIdx.x % (unsigned)tiles_x);
	const int y = a.y0 + (int)(blockIdx.x / (unsigned)tiles_x);
	const int x = tile * P2_TX + (int)threadIdx.x;
	const int lane = threadIdx.x & 31;
	const size_t nx = (size_t)a.nx;
	const unsigned long long *t_up = a.tilemask, *t_dn = a.tilemask + (size_t)a.ny * tiles_x;
	M c_up = 0, c_dn = 0;
#pragma unroll
	for (int h = 0; h < (WIDE ? 2 : 1); ++h) {
		const int j = lane + 1 + 32 * h;
		bool pu = false, pd = false;
		if (j <= a.J) {
			if (y - j >= 0) pu = (__ldg(t_dn + (size_t)(y - j) * tiles_x + tile) >> j) & 1ull;
			if (y + j < a.ny) pd = (__ldg(t_up + (size_t)(y + j) * tiles_x + tile) >> j) & 1ull;
		}
		c_up |= (M)__ballot_sync(0xffffffffu, pu) << (32 * h);
		c_dn |= (M)__ballot_sync(0xffffffffu, pd) << (32 * h);
	}
	// an empty column within reach[|dy|] of x in row y + dy, |dy| <= J (rows outside the grid are the border: empty)
	bool killed = y < a.J || a.ny - 1 - y < a.J;
	// (tile minima of the distances: can anything in the 2J + 1 rows of this tile come within reach at all? - whole warps)
	bool near_empty = true;
	if (a.dist_tmin && !killed) {
		near_empty = false;
		for (int j = lane; j <= a.J; j += 32) {
			const int r = s_reach[j];
			near_empty |= (int)__ldg(a.dist_tmin + (size_t)(y - j) * tiles_x + tile) <= r || (int)__ldg(a.dist_tmin + (size_t)(y + j) * tiles_x + tile) <= r;
		}
		near_empty = __any_sync(0xffffffffu, near_empty);
	}
	if (x >= a.nx) return;
	const size_t cc = (size_t)y * nx + x;
	const unsigned long long c = (unsigned long long)(y - a.y0) * nx + x;
	if (!killed && near_empty) {
		const uint8_t *d = a.dist + cc;
		killed = (int)__ldg(d) <= s_reach[0];
		for (int j0 = 1; j0 <= a.J && !killed; j0 += 4) {
			uint8_t v[8];
#pragma unroll
			for (int i = 0; i < 4; ++i) {
				const int j = min(j0 + i, a.J);
				v[2 * i] = __ldg(d - (size_t)j * nx); v[2 * i + 1] = __ldg(d + (size_t)j * nx);
			}
#pragma unroll
			for (int i = 0; i < 4; ++i) {
				const int r = s_reach[min(j0 + i, a.J)];
				killed |= (int)v[2 * i] <= r || (int)v[2 * i + 1] <= r;
			}
		}
	}
	if (killed) { a.st.cnt[c] = 0; return; }
	const uint16_t *f_up = a.flags, *f_dn = a.flags + (size_t)a.ny * nx;
	M m_up = 0, m_dn = 0;
	const uint16_t w_up = __ldg(f_up + cc), w_dn = __ldg(f_dn + cc);
	while (c_up | c_dn) {
		int jj[4];
		const uint16_t *p[4];
#pragma unroll
		for (int i = 0; i < 4; ++i) {
			jj[i] = 0; p[i] = f_up + cc;
			if (c_up) { const int j = mask_ffs(c_up); c_up &= c_up - 1; jj[i] = -j; p[i] = f_dn + cc - (size_t)j * nx; }
			else if (c_dn) { const int j = mask_ffs(c_dn); c_dn &= c_dn - 1; jj[i] = j; p[i] = f_up + cc + (size_t)j * nx; }
		}
		uint16_t w[4];
#pragma unroll
		for (int i = 0; i < 4; ++i) w[i] = __ldg(p[i]);
#pragma unroll
		for (int i = 0; i < 4; ++i) {
			if (jj[i] < 0) m_up |= (M)flag_in(w[i], -jj[i]) << (-jj[i] - 1);
			else if (jj[i] > 0) m_dn |= (M)flag_in(w[i], jj[i]) << (jj[i] - 1);
		}
	}
	// hull of the mirrored slots: X = min (-a - h) = -L, Y = max (-b + h) = -U
	const double inf = __longlong_as_double(0x7FF0000000000000LL);
	double X = inf, Y = -inf;
	const size_t midrow = (size_t)(a.J + 1) * nx;
	const double2 *self = a.mid + (size_t)y * midrow + x;
	if (flag_in(w_up, 0) || flag_in(w_dn, 0)) { const double2 v = __ldg(self); X = v.x < X ? v.x : X; Y = v.y > Y ? v.y : Y; }
	const size_t step_up = midrow - nx, step_dn = midrow + nx;
	while (m_up | m_dn) {
		const double2 *p[4];
		int n = 0;
#pragma unroll
		for (int i = 0; i < 4; ++i) {
			p[i] = self;
			if (m_up) { const int j = mask_ffs(m_up); m_up &= m_up - 1; p[i] = self - (size_t)j * step_up; n = i + 1; }
			else if (m_dn) { const int j = mask_ffs(m_dn); m_dn &= m_dn - 1; p[i] = self + (size_t)j * step_dn; n = i + 1; }
		}
		double2 v[4];
#pragma unroll
		for (int i = 0; i < 4; ++i) v[i] = __ldg(p[i]);
#pragma unroll
		for (int i = 0; i < 4; ++i)
			if (i < n) { X = v[i].x < X ? v[i].x : X; Y = v[i].y > Y ? v[i].y : Y; }
	}
	const double L = -X, U = -Y;
	// the dilated complement is [.., L] u [U, ..] (one interval when they touch); negate_ray_range drops the leading
	// events <= lo and the trailing ones >= hi and re-inserts a bound where an even number went
	if (U <= L || U <= a.lo || L >= a.hi) { a.st.cnt[c] = 0; return; }
	a.st.cnt[c] = 1u;
	a.st.inl[c * STAGE_INLINE] = make_double2(L <= a.lo ? a.lo : L, U >= a.hi ? a.hi : U);
}

// ---------------------------------------------------------------------------------------------------
// 'brute_force' (src/vor3d/VoronoiBruteForce.cpp:16-100): one thread per output column gathers the
// whole disc. HB[|dy|][|dx|] = sqrt(R*R - dx^2 - dy^2) where dx^2 + dy^2 <= R*R, else -1.
// ---------------------------------------------------------------------------------------------------
struct BruteArgs {
	int nx, ny, J;
	const uint32_t *off;
	const double2 *spans;
	const double *HB;       // (J+1)*(J+1), [|dy|][|dx|]
	Stage st;
	Redo redo;
	Work wk;
};

template <int CAP>
__device__ __forceinline__ void brute_item(const BruteArgs &a, unsigned long long c)
{
	const int x = (int)(c % (unsigned)a.nx);
	const int y = (int)(c / (unsigned)a.nx);

	ListStore<CAP> ulist;
	RunUnion<CAP> u(ulist.ptr(a.wk));
	const int ylo = max(-a.J, -y), yhi = min(a.J, a.ny - 1 - y);
	const int xlo = max(-a.J, -x), xhi = min(a.J, a.nx - 1 - x);
	for (int dy = ylo; dy <= yhi; ++dy) {
		const double *Hrow = a.HB + (size_t)abs(dy) * (a.J + 1);
		const size_t c0 = (size_t)(y + dy) * a.nx + x;
		uint32_t o0 = __ldg(a.off + c0 + xlo);
		for (int dx = xlo; dx <= xhi; ++dx) {
			const uint32_t o1 = __ldg(a.off + c0 + dx + 1);
			if (o1 > o0) {
				const double h = __ldg(Hrow + abs(dx));
				if (h >= 0.0)
					for (uint32_t k = o0; k < o1; ++k) {
						const double2 v = __ldg(a.spans + k);
						u.insert(v.x - h, v.y + h);
					}
			}
			o0 = o1;
		}
	}
	if (u.overflow) { overflow_item(a.wk, a.redo, c); a.st.cnt[c] = 0; return; }
	stage_emit(a.st, (size_t)c, u);
}

template <int CAP>
__global__ void __launch_bounds__(128) k_brute(BruteArgs a)
{
	VO_FOR_WORK(CAP, a.wk, c) brute_item<CAP>(a, c);
}

// ---------------------------------------------------------------------------------------------------
// vor2d (src/vor2d/DoubleCompressedImage.cpp:680-703, DoubleVoronoi.h:101-148, DoubleVoronoi.cpp:713-725):
// one thread per row i gathers rows i+di, |di| <= J, caps h2[|di|] = sqrt(R*R - di*di), every candidate
// clamped to [0, W] and dropped when it lies outside. complement != 0 is the erosion sweep: the seeds
// are the gaps of each row with extremes -1 and W, and rows -1 and `rows` are full sentinels.
// ---------------------------------------------------------------------------------------------------
struct Dil2dArgs {
	int rows, J, complement;
	double W;
	const uint32_t *off;
	const double2 *spans;
	const double *h2;       // J+1
	Stage st;
	Redo redo;
	Work wk;
};

template <int CAP>
__device__ __forceinline__ void clamp_insert(RunUnion<CAP> &u, double y1, double y2, double h, double W)
{
	const double a = (y1 - h > 0) ? y1 - h : 0;
	const double b = (y2 + h < W) ? y2 + h : W;
	if (a > W || b < 0) return;
	u.insert(a, b);
}

template <int CAP>
__device__ __forceinline__ void dilate2d_item(const Dil2dArgs &a, unsigned long long c)
{
	const int i = (int)c;
	ListStore<CAP> ulist;
	RunUnion<CAP> u(ulist.ptr(a.wk));
	for (int di = -a.J; di <= a.J; ++di) {
		const int r = i + di;
		const double h = __ldg(a.h2 + abs(di));
		if (!a.complement) {
			if (r < 0 || r >= a.rows) continue;
			for (uint32_t k = a.off[r]; k < a.off[r + 1]; ++k) {
				const double2 v = __ldg(a.spans + k);
				clamp_insert(u, v.x, v.y, h, a.W);
			}
		} else {
			if (r < -1 || r > a.rows) continue;
			if (r == -1 || r == a.rows) { clamp_insert(u, -1.0, a.W, h, a.W); continue; }
			double j2 = a.W;
			for (long long k = (long long)a.off[r + 1] - 1; k >= (long long)a.off[r]; --k) {
				const double2 v = __ldg(a.spans + k);
				clamp_insert(u, v.y, j2, h, a.W);
				j2 = v.x;
			}
			clamp_insert(u, -1.0, j2, h, a.W);
		}
	}
	if (u.overflow) { overflow_item(a.wk, a.redo, c); a.st.cnt[c] = 0; return; }
	stage_emit(a.st, (size_t)c, u);
}

template <int CAP>
__global__ void __launch_bounds__(64) k_dilate2d(Dil2dArgs a)
{
	VO_FOR_WORK(CAP, a.wk, c) dilate2d_item<CAP>(a, c);
}

// ---------------------------------------------------------------------------------------------------
// vor2d, block-per-row variant: rows of a 2D image hold tens of intervals and gather from 2J+1 rows, so a
// row's union has up to a few thousand candidates - too long for one thread's running list, too little
// for the tile kernel. One CTA per row: candidates (clamped like DoubleVoronoi.cpp:713-725) go to shared
// memory, are sorted by start (bitonic), a prefix maximum of the ends marks where a new component starts
// (start > every earlier end; touching intervals merge like appendSegment), and the components are written
// straight to the staged output. Rows with more than D2B_MAX candidates go to the redo list and are
// handled by the one-thread-per-row kernel.
// ---------------------------------------------------------------------------------------------------
constexpr int D2B_MAX = 2048;       // candidates per row held in shared memory
constexpr int D2B_THREADS = 256;

__device__ __forceinline__ double2 clamp_candidate(double y1, double y2, double h, double W)
{
	const double a = (y1 - h > 0) ? y1 - h : 0;
	const double b = (y2 + h < W) ? y2 + h : W;
	if (a > W || b < 0) return slot_empty();
	return make_double2(a, b);
}

__global__ void __launch_bounds__(D2B_THREADS) k_dilate2d_block(Dil2dArgs a)
{
	__shared__ double s_key[D2B_MAX];      // starts
	__shared__ double s_val[D2B_MAX];      // ends, later their running maximum
	__shared__ int s_rowbase[130];         // first candidate slot of each source row (2J+1 <= 129)
	__shared__ int s_scan[D2B_THREADS];
	__shared__ int s_total, s_ncomp;
	__shared__ unsigned long long s_pool;
	const int i = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
	if (i >= a.rows) return;
	const int nsrc = 2 * a.J + 1;
	const double inf = __longlong_as_double(0x7FF0000000000000LL);
	if (nsrc > 129) {                                      // more source rows than s_rowbase holds: one-thread-per-row kernel
		if (tid == 0) { redo_push(a.redo, (unsigned long long)i); a.st.cnt[i] = 0; }
		return;
	}
	// candidates per source row (complement rows have one more seed than intervals; sentinels have one)
	if (tid == 0) {
		int tot = 0;
		for (int q = 0; q < nsrc; ++q) {
			const int r = i - a.J + q;
			int c = 0;
			if (!a.complement) { if (r >= 0 && r < a.rows) c = (int)(a.off[r + 1] - a.off[r]); }
			else if (r == -1 || r == a.rows) c = 1;
			else if (r >= 0 && r < a.rows) c = (int)(a.off[r + 1] - a.off[r]) + 1;
			s_rowbase[q] = tot;
			tot += c;
		}
		s_rowbase[nsrc] = tot;
		s_total = tot;
	}
	__syncthreads();
	const int total = s_total;
	if (total > D2B_MAX) {                                 // too long for shared memory: one-thread-per-row kernel
		if (tid == 0) { redo_push(a.redo, (unsigned long long)i); a.st.cnt[i] = 0; }
		return;
	}
	int npow = 1;
	while (npow < total) npow <<= 1;
	for (int k = total + tid; k < npow; k += nthr) { s_key[k] = inf; s_val[k] = -inf; }
	// fill: a warp per source row, lanes over its seeds
	const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
	for (int q = warp; q < nsrc; q += nwarp) {
		const int r = i - a.J + q;
		const int b0 = s_rowbase[q], n = s_rowbase[q + 1] - b0;
		if (n == 0) continue;
		const double h = __ldg(a.h2 + abs(r - i));
		if (!a.complement) {
			const uint32_t o = a.off[r];
			for (int k = lane; k < n; k += 32) {
				const double2 v = __ldg(a.spans + o + k);
				const double2 c = clamp_candidate(v.x, v.y, h, a.W);
				s_key[b0 + k] = c.x; s_val[b0 + k] = c.y;
			}
		} else if (r == -1 || r == a.rows) {
			if (lane == 0) { const double2 c = clamp_candidate(-1.0, a.W, h, a.W); s_key[b0] = c.x; s_val[b0] = c.y; }
		} else {
			// seeds of the complement: [-1, a_0], [b_0, a_1], ..., [b_last, W]   (DoubleVoronoi.h:132-144)
			const uint32_t o = a.off[r];
			for (int k = lane; k < n; k += 32) {
				const double lo = (k == 0) ? -1.0 : __ldg(a.spans + o + k - 1).y;
				const double hi = (k == n - 1) ? a.W : __ldg(a.spans + o + k).x;
				const double2 c = clamp_candidate(lo, hi, h, a.W);
				s_key[b0 + k] = c.x; s_val[b0 + k] = c.y;
			}
		}
	}
	__syncthreads();
	// bitonic sort by start (dropped candidates carry start = +inf and end = -inf: they sink to the end)
	for (int k = 2; k <= npow; k <<= 1)
		for (int j = k >> 1; j > 0; j >>= 1) {
			for (int t = tid; t < npow; t += nthr) {
				const int p = t ^ j;
				if (p > t) {
					const bool up = (t & k) == 0;
					const double kt = s_key[t], kp = s_key[p];
					if ((kt > kp) == up) {
						s_key[t] = kp; s_key[p] = kt;
						const double vt = s_val[t]; s_val[t] = s_val[p]; s_val[p] = vt;
					}
				}
			}
			__syncthreads();
		}
	// running maximum of the ends (each thread owns a contiguous chunk), component starts, their ranks
	const int chunk = (npow + nthr - 1) / nthr;
	const int c0 = tid * chunk, c1 = min(c0 + chunk, npow);
	double m = -inf;
	for (int k = c0; k < c1; ++k) { m = s_val[k] > m ? s_val[k] : m; }
	__shared__ double s_cmax[D2B_THREADS];
	s_cmax[tid] = m;
	__syncthreads();
	double before = -inf;                                   // maximum end of everything before this chunk
	for (int q = 0; q < tid; ++q) before = s_cmax[q] > before ? s_cmax[q] : before;
	int nstart = 0;
	double run = before;
	for (int k = c0; k < c1; ++k) {
		if (s_key[k] != inf && s_key[k] > run) ++nstart;   // opens a new component (the very first has run = -inf)
		run = s_val[k] > run ? s_val[k] : run;
	}
	s_scan[tid] = nstart;
	__syncthreads();
	if (tid == 0) {
		int acc = 0;
		for (int q = 0; q < nthr; ++q) { const int v = s_scan[q]; s_scan[q] = acc; acc += v; }
		s_ncomp = acc;
		a.st.cnt[i] = (uint32_t)acc;
		if (acc > STAGE_INLINE) {
			s_pool = atomicAdd(a.st.cursor, (unsigned long long)acc);
			a.st.inl[(size_t)i * STAGE_INLINE] = slot_pool(s_pool, (unsigned int)acc);
		}
	}
	__syncthreads();
	const int ncomp = s_ncomp;
	if (ncomp == 0) return;
	const bool pooled = ncomp > STAGE_INLINE;
	if (pooled && s_pool + ncomp > a.st.pool_cap) return;   // the host regrows the pool and repeats
	double2 *dst = pooled ? a.st.pool + s_pool : a.st.inl + (size_t)i * STAGE_INLINE;
	// component c = [start of its first element, running maximum just before the next component starts]:
	// the thread that meets the NEXT start closes a component; the last one ends at the overall maximum.
	int comp = s_scan[tid] - 1;                             // component the chunk begins inside (-1: none yet)
	run = before;
	for (int k = c0; k < c1; ++k) {
		if (s_key[k] == inf) break;
		if (s_key[k] > run) {
			if (comp >= 0) dst[comp].y = run;
			++comp;
			dst[comp].x = s_key[k];
		}
		run = s_val[k] > run ? s_val[k] : run;
	}
	if (tid == 0) {
		double all = -inf;
		for (int q = 0; q < nthr; ++q) all = s_cmax[q] > all ? s_cmax[q] : all;
		dst[ncomp - 1].y = all;
	}
}

// Staged lists -> canonical CSR (after the exclusive scan of cnt). Lists that would end beyond `cap` intervals
// are skipped (the caller notices from the total and repeats with a larger buffer).
__global__ void __launch_bounds__(256) k_compact(Stage st, unsigned long long nlists,
                                                 const uint32_t *__restrict__ off, double2 *__restrict__ spans,
                                                 unsigned long long cap = ~0ull)
{
	const unsigned long long c = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= nlists) return;
	const uint32_t n = st.cnt[c];
	if (n == 0) return;
	if ((unsigned long long)off[c] + n > cap) return;
	double2 *dst = spans + off[c];
	if (n <= STAGE_INLINE) {
		for (uint32_t k = 0; k < n; ++k) dst[k] = st.inl[c * STAGE_INLINE + k];
	} else {
		const unsigned long long base = slot_pool_base(st.inl[c * STAGE_INLINE]);
		if (base + n > st.pool_cap) return;       // the staging pool overflowed (banded call: the host only learns afterwards and repeats)
		for (uint32_t k = 0; k < n; ++k) dst[k] = st.pool[base + k];
	}
}

// Staged lists -> canonical CSR in ONE launch (scan.cuh: decoupled look-back): offsets of every list, off[nlists] and
// *total_out = grand total, spans copied where the list ends inside `cap` intervals.
template <int ITEMS>                                    // lists per thread: 8 for millions of lists, 2 when that would leave SMs without a tile
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_compact(Stage st, unsigned long long nlists, uint32_t *__restrict__ off,
                                                               double2 *__restrict__ spans, unsigned long long cap,
                                                               unsigned long long *state, unsigned long long *ticket,
                                                               unsigned long long ticket_base, uint32_t epoch,
                                                               unsigned long long *total_out,
                                                               const unsigned long long *base_in = nullptr,
                                                               unsigned long long *max_word = nullptr, uint32_t max_tag = 0)
{
	KT_SCOPE(KT_SCAN_COMPACT, 0, threadIdx.x == 0);
	// max_word (optional): receives max_tag << 32 | the largest list length seen (atomicMax: a word left by an earlier
	// call carries a smaller tag, so nobody has to clear it)
	// base_in (optional): the offsets start from *base_in instead of 0 (a band of rows of a larger volume: `spans` and
	// `cap` are the whole volume's, *total_out = the running total after this band)
	constexpr int TILE = SCAN_THREADS * ITEMS;
	__shared__ uint32_t s_tile;
	__shared__ unsigned long long s_prefix;
	__shared__ uint32_t s_rel[TILE + 1];              // offsets of the tile's lists relative to the tile's first
	if (threadIdx.x == 0) s_tile = (uint32_t)(atomicAdd(ticket, 1ull) - ticket_base);
	__syncthreads();
	const uint32_t tile = s_tile;
	const unsigned long long tbase = (unsigned long long)tile * TILE, base = tbase + (unsigned long long)threadIdx.x * ITEMS;
	uint32_t c[ITEMS];
	unsigned long long sum = 0;
#pragma unroll
	for (int i = 0; i < ITEMS; ++i) {
		const unsigned long long k = base + i;
		c[i] = k < nlists ? st.cnt[k] : 0u;
		sum += c[i];
	}
	unsigned long long tot;
	const unsigned long long exl = block_excl_scan(sum, &tot);
	volatile unsigned long long *vstate = state;
	if (threadIdx.x == 0) {
		vstate[tile] = scan_word(tot, epoch, tile == 0 ? SCAN_INCL : SCAN_AGG);
		if (tile == 0) s_prefix = 0;
	}
	if (max_word) {
		uint32_t mx = 0;
#pragma unroll
		for (int i = 0; i < ITEMS; ++i) mx = c[i] > mx ? c[i] : mx;
		mx = __reduce_max_sync(0xffffffffu, mx);
		if ((threadIdx.x & 31) == 0 && mx) {
			const unsigned long long w = ((unsigned long long)max_tag << 32) | mx;
			if (*(volatile unsigned long long *)max_word < w) atomicMax(max_word, w);   // (nearly every warp finds its value there already)
		}
	}
	{
		uint32_t r = (uint32_t)exl;
#pragma unroll
		for (int i = 0; i < ITEMS; ++i) { s_rel[threadIdx.x * ITEMS + i] = r; r += c[i]; }
		if (threadIdx.x == blockDim.x - 1) s_rel[TILE] = r;
	}
	if (tile > 0 && threadIdx.x < 32) {
		const unsigned long long excl = scan_look_back(vstate, tile, epoch);
		if (threadIdx.x == 0) {
			vstate[tile] = scan_word(excl + tot, epoch, SCAN_INCL);
			s_prefix = excl;
		}
	}
	__syncthreads();
	const unsigned long long first = s_prefix + (base_in ? *base_in : 0ull);    // global offset of the tile's first list
	if (tile == gridDim.x - 1 && threadIdx.x == 0) {          // the last ticket is the last tile: its end is the grand total
		off[nlists] = (uint32_t)(first + tot);
		*total_out = first + tot;
	}
	// lists of the tile by stride: neighbouring threads take neighbouring lists (coalesced, ITEMS independent copies
	// in flight per thread)
#pragma unroll
	for (int i = 0; i < ITEMS; ++i) {
		const int j = i * SCAN_THREADS + (int)threadIdx.x;
		const unsigned long long k = tbase + j;
		if (k >= nlists) break;
		const unsigned long long ex = first + s_rel[j];
		const uint32_t n = s_rel[j + 1] - s_rel[j];
		off[k] = (uint32_t)ex;
		if (n && ex + n <= cap) {
			double2 *dst = spans + ex;
			if (n <= STAGE_INLINE) {
				for (uint32_t q = 0; q < n; ++q) dst[q] = st.inl[k * STAGE_INLINE + q];
			} else {
				const unsigned long long pb = slot_pool_base(st.inl[k * STAGE_INLINE]);
				if (pb + n <= st.pool_cap)
					for (uint32_t q = 0; q < n; ++q) dst[q] = st.pool[pb + q];
			}
		}
	}
}

// ---------------------------------------------------------------------------------------------------
// Complement kernels of the erosion composite (src/vor3d/Voronoi.cpp:18-89).
// negate: output grid (nx+2*border)^2-ish with `border` empty lines added on every side, every column
// complemented inside [lo, hi] with negate_ray's exact '==' tests (MorphologyOperators.cpp:230-259).
// border = 1 for the 3D erosion, 0 for the vor2d negate() (DoubleCompressedImage.cpp:438-468).
// ---------------------------------------------------------------------------------------------------
struct NegArgs {
	int nx, ny, border;     // source grid, border columns on either side in x
	int by0, by1;           // border rows before / after (a y-slab of a sharded grid only has the rows of the global border it owns)
	double lo, hi;
	const uint32_t *off;
	const double2 *spans;
	uint32_t *cnt;          // count kernel output
	const uint32_t *out_off;
	double2 *out_spans;
	unsigned int *outside;  // (count pass, optional) set when a column has data outside [lo, hi]: the complement is then
	                        // not a set of proper intervals and erosion must not prune with the clip range
};

__device__ __forceinline__ bool neg_src(const NegArgs &a, unsigned long long c, uint32_t &o0, uint32_t &o1)
{
	const int mx = a.nx + 2 * a.border;
	const int x = (int)(c % (unsigned)mx) - a.border;
	const int y = (int)(c / (unsigned)mx) - a.by0;
	if (x < 0 || x >= a.nx || y < 0 || y >= a.ny) { o0 = o1 = 0; return false; }
	const size_t s = (size_t)y * a.nx + x;
	o0 = a.off[s];
	o1 = a.off[s + 1];
	return true;
}

// one column of negate: number of intervals of its complement (and the "data outside [lo, hi]" flag)
__device__ __forceinline__ uint32_t neg_count(const NegArgs &a, unsigned long long c)
{
	uint32_t o0, o1;
	neg_src(a, c, o0, o1);
	const uint32_t k = o1 - o0;
	if (k == 0) return 1u;
	const bool f = (a.spans[o0].x == a.lo);          // first event == lo: erased
	const bool l = (a.spans[o1 - 1].y == a.hi);      // last event == hi: popped
	if (a.outside && (a.spans[o0].x < a.lo || a.spans[o1 - 1].y > a.hi)) *a.outside = 1u;
	return k + 1 - (f ? 1u : 0u) - (l ? 1u : 0u);
}

// ... and the intervals themselves, written to dst[0 .. neg_count)
__device__ __forceinline__ void neg_fill(const NegArgs &a, unsigned long long c, double2 *dst)
{
	uint32_t o0, o1;
	neg_src(a, c, o0, o1);
	if (o1 == o0) { dst[0] = make_double2(a.lo, a.hi); return; }
	const bool f = (a.spans[o0].x == a.lo);
	const bool l = (a.spans[o1 - 1].y == a.hi);
	// event sequence: [lo if !f] z1_0? ... pairs re-formed from consecutive events
	double start = f ? a.spans[o0].y : a.lo;
	uint32_t w = 0;
	if (f) {
		// events after erasing the first: y_0, x_1, y_1, ...
		for (uint32_t i = o0 + 1; i < o1; ++i) { dst[w++] = make_double2(start, a.spans[i].x); start = a.spans[i].y; }
	} else {
		for (uint32_t i = o0; i < o1; ++i) { dst[w++] = make_double2(start, a.spans[i].x); start = a.spans[i].y; }
	}
	if (!l) dst[w++] = make_double2(start, a.hi);
}

template <bool FILL>
__global__ void __launch_bounds__(256) k_negate(NegArgs a, unsigned long long nlists)
{
	const unsigned long long c = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= nlists) return;
	if (FILL) neg_fill(a, c, a.out_spans + a.out_off[c]);
	else a.cnt[c] = neg_count(a, c);
}

// negateInv: strip `border` lines on every side, then negate_ray_range (MorphologyOperators.cpp:282-315):
// leading events <= lo and trailing events >= hi are dropped; a bound is re-inserted when an even
// number of events was dropped on that side; an empty column becomes [lo, hi].
struct NegInvArgs {
	int mx, my, border;     // source (bordered) grid, border columns on either side in x
	int by0;                // border rows stripped before the first row (the rows after the last are simply not visited)
	double lo, hi;
	const uint32_t *off;
	const double2 *spans;
	uint32_t *cnt;
	const uint32_t *out_off;
	double2 *out_spans;
};

__device__ __forceinline__ double ev_at(const double2 *sp, uint32_t base, uint32_t i)
{
	const double2 v = sp[base + (i >> 1)];
	return (i & 1u) ? v.y : v.x;
}

// one column of negateInv: FILL = false returns the number of intervals, FILL = true also writes them to dst
template <bool FILL>
__device__ __forceinline__ uint32_t neg_inv_col(const NegInvArgs &a, unsigned long long c, double2 *dst)
{
	const int nx = a.mx - 2 * a.border;
	const int x = (int)(c % (unsigned)nx) + a.border;
	const int y = (int)(c / (unsigned)nx) + a.by0;
	const size_t s = (size_t)y * a.mx + x;
	const uint32_t o0 = a.off[s], o1 = a.off[s + 1];
	const uint32_t n = 2 * (o1 - o0);               // events
	if (n == 0) {
		if (FILL) dst[0] = make_double2(a.lo, a.hi);
		return 1u;
	}
	uint32_t cf = 0;
	while (cf < n && ev_at(a.spans, o0, cf) <= a.lo) ++cf;
	const uint32_t pre = (cf % 2 == 0) ? 1u : 0u;    // lo re-inserted in front
	uint32_t k = pre + (n - cf);                     // length of the edited list
	uint32_t cl = 0;
	while (k > 0) {
		const double last = (k - 1 >= pre) ? ev_at(a.spans, o0, cf + (k - 1 - pre)) : a.lo;
		if (last >= a.hi) { --k; ++cl; } else break;
	}
	const uint32_t app = (cl % 2 == 0) ? 1u : 0u;    // hi re-appended
	const uint32_t total = k + app;                  // events of the result
	if (!FILL) return total / 2;
	// event t of the result: t < min(pre,k) -> lo ; t < k -> e[cf + t - pre] ; t == k (app) -> hi
	const uint32_t npre = (pre < k) ? pre : k;
	for (uint32_t t = 0; t + 1 < total + 0u; t += 2) {
		double v[2];
		for (uint32_t q = 0; q < 2; ++q) {
			const uint32_t tt = t + q;
			v[q] = (tt < npre) ? a.lo : (tt < k) ? ev_at(a.spans, o0, cf + tt - pre) : a.hi;
		}
		dst[t >> 1] = make_double2(v[0], v[1]);
	}
	return total / 2;
}

template <bool FILL>
__global__ void __launch_bounds__(256) k_negate_inv(NegInvArgs a, unsigned long long nlists)
{
	const unsigned long long c = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= nlists) return;
	if (FILL) neg_inv_col<true>(a, c, a.out_spans + a.out_off[c]);
	else a.cnt[c] = neg_inv_col<false>(a, c, nullptr);
}

// Both complements in ONE launch each (count -> single-pass prefix sum with look-back -> fill, scan.cuh) instead of
// count / reduce / scan / apply / fill with a host round trip in the middle. The output buffer holds `cap` intervals
// (the caller passes the bound k + 1 per column, which both complements respect); *total_out = the exact total.
struct NegOp {
	typedef NegArgs Args;
	static __device__ __forceinline__ uint32_t count(const Args &a, unsigned long long c) { return neg_count(a, c); }
	static __device__ __forceinline__ void fill(const Args &a, unsigned long long c, double2 *dst) { neg_fill(a, c, dst); }
};
struct NegInvOp {
	typedef NegInvArgs Args;
	static __device__ __forceinline__ uint32_t count(const Args &a, unsigned long long c) { return neg_inv_col<false>(a, c, nullptr); }
	static __device__ __forceinline__ void fill(const Args &a, unsigned long long c, double2 *dst) { neg_inv_col<true>(a, c, dst); }
};

constexpr int CF_ITEMS = 4;                             // columns per thread
constexpr int CF_TILE = SCAN_THREADS * CF_ITEMS;

template <typename Op>
__global__ void __launch_bounds__(SCAN_THREADS) k_complement_fused(typename Op::Args a, unsigned long long nlists, uint32_t *__restrict__ off,
                                                                   double2 *__restrict__ spans, unsigned long long cap,
                                                                   unsigned long long *state, unsigned long long *ticket,
                                                                   unsigned long long ticket_base, uint32_t epoch,
                                                                   unsigned long long *total_out)
{
	__shared__ uint32_t s_tile;
	__shared__ unsigned long long s_prefix;
	__shared__ uint32_t s_n[CF_TILE];                       // counts, then offsets relative to the tile's first column
	if (threadIdx.x == 0) s_tile = (uint32_t)(atomicAdd(ticket, 1ull) - ticket_base);
	__syncthreads();
	const uint32_t tile = s_tile;
	const unsigned long long tbase = (unsigned long long)tile * CF_TILE;
	// counts by stride (neighbouring threads read neighbouring columns), scanned in column order
#pragma unroll
	for (int i = 0; i < CF_ITEMS; ++i) {
		const int j = i * SCAN_THREADS + (int)threadIdx.x;
		const unsigned long long c = tbase + j;
		s_n[j] = c < nlists ? Op::count(a, c) : 0u;
	}
	__syncthreads();
	uint32_t n[CF_ITEMS];
	unsigned long long sum = 0;
#pragma unroll
	for (int i = 0; i < CF_ITEMS; ++i) { n[i] = s_n[threadIdx.x * CF_ITEMS + i]; sum += n[i]; }
	unsigned long long tot;
	const unsigned long long exl = block_excl_scan(sum, &tot);
	volatile unsigned long long *vstate = state;
	if (threadIdx.x == 0) {
		vstate[tile] = scan_word(tot, epoch, tile == 0 ? SCAN_INCL : SCAN_AGG);
		if (tile == 0) s_prefix = 0;
	}
	{
		uint32_t r = (uint32_t)exl;
#pragma unroll
		for (int i = 0; i < CF_ITEMS; ++i) { s_n[threadIdx.x * CF_ITEMS + i] = r; r += n[i]; }
	}
	if (tile > 0 && threadIdx.x < 32) {
		const unsigned long long excl = scan_look_back(vstate, tile, epoch);
		if (threadIdx.x == 0) {
			vstate[tile] = scan_word(excl + tot, epoch, SCAN_INCL);
			s_prefix = excl;
		}
	}
	__syncthreads();
	const unsigned long long first = s_prefix;
	if (tile == gridDim.x - 1 && threadIdx.x == 0) {
		off[nlists] = (uint32_t)(first + tot);
		*total_out = first + tot;
	}
#pragma unroll
	for (int i = 0; i < CF_ITEMS; ++i) {
		const int j = i * SCAN_THREADS + (int)threadIdx.x;
		const unsigned long long c = tbase + j;
		if (c >= nlists) break;
		const unsigned long long ex = first + s_n[j];
		const uint32_t cnt = (j + 1 < CF_TILE ? s_n[j + 1] : (uint32_t)tot) - s_n[j];
		off[c] = (uint32_t)ex;
		if (cnt && ex + cnt <= cap) Op::fill(a, c, spans + ex);
	}
}

// ---------------------------------------------------------------------------------------------------
// xor of two same-grid volumes (src/vor3d/Voronoi.cpp:91-111, MorphologyOperators.cpp:334-374):
// symmetric difference inside [zmin, zmax], slivers shorter than 1e-10 dropped. One thread per column
// walks both sorted event lists once. cnt pass and fill pass share the walk.
// ---------------------------------------------------------------------------------------------------
struct XorArgs {
	const uint32_t *off_a; const double2 *sp_a;
	const uint32_t *off_b; const double2 *sp_b;
	double lo, hi;
	uint32_t *cnt;
	const uint32_t *out_off;
	double2 *out_spans;
	double *col_len;        // per-column summed length (fill pass)
};

// membership of z-range pieces: the reference computes (a \ b) u (b \ a) through negate_ray / unionSegs;
// for sorted disjoint inputs inside [lo, hi] that is the set of maximal runs where exactly one of the
// two lists covers, with touching runs coalesced (unionSegs merges start <= end) and both inputs
// clipped the way negate_ray clips (events equal to the bounds vanish).
template <bool FILL>
__global__ void __launch_bounds__(256) k_xor(XorArgs a, unsigned long long nlists)
{
	const unsigned long long c = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= nlists) return;
	const uint32_t a0 = a.off_a[c], na = 2 * (a.off_a[c + 1] - a0);
	const uint32_t b0 = a.off_b[c], nb = 2 * (a.off_b[c + 1] - b0);
	uint32_t ia = 0, ib = 0, w = 0;
	bool in_a = false, in_b = false, open = false;
	double start = 0, len = 0, cur_s = 0, cur_e = 0;
	bool have = false;
	double2 *dst = FILL ? a.out_spans + a.out_off[c] : nullptr;
	while (ia < na || ib < nb) {
		const double za = (ia < na) ? ev_at(a.sp_a, a0, ia) : 0.0;
		const double zb = (ib < nb) ? ev_at(a.sp_b, b0, ib) : 0.0;
		double z;
		if (ib >= nb || (ia < na && za <= zb)) { z = za; in_a = !in_a; ++ia; if (ib < nb && zb == z) { in_b = !in_b; ++ib; } }
		else { z = zb; in_b = !in_b; ++ib; }
		const bool x = (in_a != in_b);
		if (x && !open) { open = true; start = z; }
		else if (!x && open) {
			open = false;
			// piece [start, z]; coalesce with the previous piece when touching
			if (have && start <= cur_e) { if (z > cur_e) cur_e = z; }
			else {
				if (have && !(cur_e - cur_s < 1e-10)) { if (FILL) dst[w] = make_double2(cur_s, cur_e); len += cur_e - cur_s; ++w; }
				cur_s = start; cur_e = z; have = true;
			}
		}
	}
	if (have && !(cur_e - cur_s < 1e-10)) { if (FILL) dst[w] = make_double2(cur_s, cur_e); len += cur_e - cur_s; ++w; }
	if (FILL) a.col_len[c] = len; else a.cnt[c] = w;
}

// Intervals [*begin, min(*end, cap)) of a result volume to the caller's pinned host buffer through the SMs (16-byte
// stores over PCIe, four in flight per thread): the range is only known on the device, so a copy-engine download
// would need a host round trip per band of the host-buffer call first (vo_lib.cu: dilate_ours_pipelined).
constexpr int COPY_OUT_THREADS = 256;
__global__ void __launch_bounds__(COPY_OUT_THREADS) k_copy_out(const double2 *__restrict__ src, double2 *__restrict__ dst_host,
                                                               const unsigned long long *begin, const unsigned long long *end,
                                                               unsigned long long cap)
{
	KT_SCOPE(KT_COPY_OUT, 0, threadIdx.x == 0);
	const unsigned long long i0 = *begin, i1 = min(*end, cap);
	const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
	unsigned long long i = i0 + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	for (; i + 3 * stride < i1; i += 4 * stride) {
		const double2 a = src[i], b = src[i + stride], c = src[i + 2 * stride], d = src[i + 3 * stride];
		dst_host[i] = a; dst_host[i + stride] = b; dst_host[i + 2 * stride] = c; dst_host[i + 3 * stride] = d;
	}
	for (; i < i1; i += stride) dst_host[i] = src[i];
}

// Rebase a copied slice of offsets so that it starts at zero.
__global__ void __launch_bounds__(256) k_rebase(uint32_t *off, unsigned long long n, uint32_t base, uint32_t add)
{
	const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) off[i] = off[i] - base + add;
}

// Rows [c0, c1) (columns) of a volume packed for a halo message: offsets rebased to 0, then the header
// [interval count, overflow flag], spans copied when they fit `cap` intervals (voroffset_b200/slab.py protocol).
__global__ void __launch_bounds__(256) k_halo_pack(const uint32_t *__restrict__ off, const double2 *__restrict__ spans,
                                                   unsigned long long c0, unsigned long long c1, uint32_t *__restrict__ out_off,
                                                   double2 *__restrict__ out_spans, unsigned long long cap)
{
	const uint32_t b = off[c0], e = off[c1];
	const unsigned long long n = c1 - c0, cnt = e - b;
	const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x, t0 = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	for (unsigned long long i = t0; i <= n; i += stride) out_off[i] = off[c0 + i] - b;
	if (t0 == 0) { out_off[n + 1] = (uint32_t)(cnt & 0x7fffffffu); out_off[n + 2] = cnt > cap ? 1u : 0u; }
	if (cnt <= cap)
		for (unsigned long long k = t0; k < cnt; k += stride) out_spans[k] = spans[b + k];
}

// Sum of (z2 - z1) per block, accumulated in double with a fixed tree order; host adds the partials.
__global__ void __launch_bounds__(256) k_sum(const double *__restrict__ v, unsigned long long n, double *__restrict__ partial)
{
	__shared__ double sh[256];
	double s = 0;
	for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x)
		s += v[i];
	sh[threadIdx.x] = s;
	__syncthreads();
	for (int d = 128; d > 0; d >>= 1) {
		if ((int)threadIdx.x < d) sh[threadIdx.x] += sh[threadIdx.x + d];
		__syncthreads();
	}
	if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

} // namespace vo
