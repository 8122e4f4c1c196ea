/*
 * TEST INFRASTRUCTURE ONLY. This file is the CPU *checker* for voroffset_b200; nothing in the product
 * path (voroffset_b200/, include/, the C-ABI library, the C++ adapters, the CLIs) may include, link,
 * import or execute it. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs load liboracle.so.
 *
 * What it is: a plain-C restatement of the NET EFFECT of the reference's dexel-morphology hot path
 * (geometryprocessing/voroffset, citations relative to /root/reference):
 *
 *   dilation, method 'ours'  (src/vor3d/VoronoiVorPower.cpp:24-96):
 *     pass 1 emits, for every seed interval [a,b] of column (x, y+dy) still alive at line y
 *     (|dy| <= floor(R): Voronoi2D.cpp:651,658), the piece (a, b, r1) with r1 = R if dy == 0 else
 *     sqrt(R*R - dy*dy) (Voronoi2D.cpp:704-716). Pass 2 keeps a piece of column (x+dx, y) alive while
 *     |dx| <= floor(r1) (SeparatePower2D.cpp:241,266) and emits [a - h, b + h], h = sqrt(r1*r1 - dx*dx)
 *     (SeparatePower2D.cpp:312-317,328-333); everything is merged with the closed-interval rule
 *     "next.start <= prev.end coalesces" (MorphologyOperators.cpp:15-34, hpp:7-19). The reference's
 *     Voronoi / power-diagram bookkeeping only prunes candidates that cannot change that union, so the
 *     result is the plain union below. This is pinned bit-for-bit against the compiled reference
 *     (oracle/_ref) by tests/test_oracle_vs_ref.py and the committed fixtures in tests/golden/.
 *   dilation, method 'brute_force' (src/vor3d/VoronoiBruteForce.cpp:16-100): restated literally
 *     (clamped neighbour ranges, dx^2+dy^2 <= R*R, dz = sqrt(R*R - dx^2 - dy^2), endpoint events sorted
 *     by coordinate, running counter).
 *   erosion = negate, dilation, negateInv (src/vor3d/Voronoi.cpp:8-89, MorphologyOperators.cpp:230-315).
 *   opening / closing as app/cli3d/offset3d.cpp:124-133 composes them.
 *   xor (src/vor3d/Voronoi.cpp:91-111, MorphologyOperators.cpp:319-374).
 *   2D: DoubleCompressedImage::dilate/erode/open/close/negate (src/vor2d/DoubleCompressedImage.cpp:
 *     438-468,680-719; DoubleVoronoi.h:101-148; DoubleVoronoi.cpp:713-725).
 *   Either side of the path (sections at the end of this file, each with its own header):
 *     dexeliser   compute_sign and helpers (src/vor3d/Dexelize.cpp:56-225)         - PARITY UNPINNED (geogram absent)
 *     2D ingest   DoubleCompressedImage::fromImage / scanLine / unionIntersections
 *                 (src/vor2d/DoubleCompressedImage.cpp:25-111)                       - pinned against oracle/_ref
 *
 * Layout everywhere: column (x,y) of an nx*ny volume is list number c = x + nx*y; list c holds the
 * intervals ev[2*off[c]] .. ev[2*off[c+1]) as (z1,z2) pairs, ascending and disjoint.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double s, e; } iv_t;

typedef struct { iv_t *v; size_t n, cap; } ivbuf_t;

static int ivbuf_push(ivbuf_t *b, double s, double e)
{
	if (b->n == b->cap) {
		size_t nc = b->cap ? 2 * b->cap : 64;
		iv_t *nv = (iv_t *)realloc(b->v, nc * sizeof(iv_t));
		if (!nv) return 1;
		b->v = nv; b->cap = nc;
	}
	b->v[b->n].s = s; b->v[b->n].e = e; b->n++;
	return 0;
}

static int cmp_start(const void *a, const void *b)
{
	const iv_t *p = (const iv_t *)a, *q = (const iv_t *)b;
	return (p->s > q->s) - (p->s < q->s);
}

/* Closed-interval union in place; follows appendSegment / addSegmentAtTheEnd
 * (MorphologyOperators.cpp:15-34, MorphologyOperators.hpp:7-19): a candidate whose start is <= the
 * running end is absorbed. Returns the number of merged intervals. */
static size_t union_sorted(iv_t *v, size_t n)
{
	if (n == 0) return 0;
	qsort(v, n, sizeof(iv_t), cmp_start);
	size_t m = 0;
	for (size_t i = 1; i < n; ++i) {
		if (v[i].s <= v[m].e) { if (v[i].e > v[m].e) v[m].e = v[i].e; }
		else v[++m] = v[i];
	}
	return m + 1;
}

/* Per-list results are produced independently (possibly by several threads) and then packed. */
typedef struct { iv_t *v; size_t n; } list_t;

static int pack_lists(list_t *lists, uint64_t nlists, uint64_t **out_off, double **out_ev)
{
	uint64_t *off = (uint64_t *)malloc((nlists + 1) * sizeof(uint64_t));
	if (!off) return 1;
	off[0] = 0;
	for (uint64_t c = 0; c < nlists; ++c) off[c + 1] = off[c] + lists[c].n;
	double *ev = (double *)malloc((2 * off[nlists] + 2) * sizeof(double));
	if (!ev) { free(off); return 1; }
	for (uint64_t c = 0; c < nlists; ++c) {
		for (size_t k = 0; k < lists[c].n; ++k) {
			ev[2 * (off[c] + k)] = lists[c].v[k].s;
			ev[2 * (off[c] + k) + 1] = lists[c].v[k].e;
		}
		free(lists[c].v);
	}
	free(lists);
	*out_off = off; *out_ev = ev;
	return 0;
}

static int list_from(list_t *dst, const iv_t *v, size_t n)
{
	dst->n = n; dst->v = NULL;
	if (n) {
		dst->v = (iv_t *)malloc(n * sizeof(iv_t));
		if (!dst->v) return 1;
		memcpy(dst->v, v, n * sizeof(iv_t));
	}
	return 0;
}

void oracle_free(void *p) { free(p); }

/* ---- tiny pthread parallel-for (this image has no libgomp) ------------------------------------ */
#include <pthread.h>
#include <stdatomic.h>
static int g_threads = 1;
void oracle_set_threads(int n) { g_threads = n > 0 ? n : 1; }
int oracle_get_threads(void) { return g_threads; }

typedef void (*chunk_fn)(int64_t begin, int64_t end, void *ctx);
typedef struct { atomic_llong next; int64_t n, chunk; chunk_fn fn; void *ctx; } pf_t;
static void *pf_worker(void *p)
{
	pf_t *w = (pf_t *)p;
	for (;;) {
		int64_t b = atomic_fetch_add(&w->next, w->chunk);
		if (b >= w->n) break;
		int64_t e = b + w->chunk < w->n ? b + w->chunk : w->n;
		w->fn(b, e, w->ctx);
	}
	return NULL;
}
static void parallel_for(int64_t n, int64_t chunk, chunk_fn fn, void *ctx)
{
	pf_t w; atomic_init(&w.next, 0); w.n = n; w.chunk = chunk; w.fn = fn; w.ctx = ctx;
	int nt = g_threads;
	if (nt <= 1 || n <= chunk) { pf_worker(&w); return; }
	pthread_t *th = (pthread_t *)malloc((size_t)nt * sizeof(pthread_t));
	int started = 0;
	for (int i = 1; i < nt; ++i) if (pthread_create(&th[started], NULL, pf_worker, &w) == 0) started++;
	pf_worker(&w);
	for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
	free(th);
}

/* ------------------------------------------------------------------------------------------------
 * Cap tables. 'ours' uses the reference's two-step operation order; the table is indexed
 * [|dy|][|dx|] and holds -1 where the pair is out of reach.
 *   r1(dy) = R                       if dy == 0      (Voronoi2D.cpp:708-711)
 *          = sqrt(R*R - dy*dy)       otherwise       (Voronoi2D.cpp:714)
 *   h      = sqrt(r1*r1 - dx*dx)                     (SeparatePower2D.cpp:314)
 *   alive  : |dy| <= floor(R) (Voronoi2D.cpp:651,658), |dx| <= floor(r1) (SeparatePower2D.cpp:266)
 * ---------------------------------------------------------------------------------------------- */
int oracle_cap_table_ours(double R, int *out_J, double **out_tab)
{
	if (!(R >= 0)) return 1;
	int J = (int)floor(R);
	double *t = (double *)malloc((size_t)(J + 1) * (J + 1) * sizeof(double));
	if (!t) return 1;
	for (int dy = 0; dy <= J; ++dy) {
		double dyd = (double)dy;
		double r1 = (dy == 0) ? R : sqrt(R * R - dyd * dyd);
		int reach = (int)floor(r1);
		for (int dx = 0; dx <= J; ++dx) {
			double dxd = (double)dx;
			t[dy * (J + 1) + dx] = (dx <= reach) ? sqrt(r1 * r1 - dxd * dxd) : -1.0;
		}
	}
	*out_J = J; *out_tab = t;
	return 0;
}

/* dilation 'ours': plain union over the reach set with the two-step caps (see file header). */
typedef struct { int nx, ny, J; const uint64_t *off; const double *ev; const double *tab; double R; list_t *lists; atomic_int fail; } dil_ctx_t;

static void ours_chunk(int64_t c0, int64_t c1, void *p)
{
	dil_ctx_t *k = (dil_ctx_t *)p;
	const int nx = k->nx, ny = k->ny, J = k->J;
	ivbuf_t buf = {0, 0, 0};
	int fail = 0;
	for (int64_t c = c0; c < c1 && !fail; ++c) {
		int x = (int)(c % nx), y = (int)(c / nx);
		buf.n = 0;
		for (int dy = -J; dy <= J; ++dy) {
			int yy = y + dy;
			if (yy < 0 || yy >= ny) continue;
			const double *row = k->tab + (size_t)abs(dy) * (J + 1);
			for (int dx = -J; dx <= J; ++dx) {
				int xx = x + dx;
				if (xx < 0 || xx >= nx) continue;
				double h = row[abs(dx)];
				if (h < 0) continue;
				uint64_t cc = (uint64_t)xx + (uint64_t)nx * yy;
				for (uint64_t i = k->off[cc]; i < k->off[cc + 1]; ++i)
					fail |= ivbuf_push(&buf, k->ev[2 * i] - h, k->ev[2 * i + 1] + h);
			}
		}
		size_t m = union_sorted(buf.v, buf.n);
		fail |= list_from(&k->lists[c], buf.v, m);
	}
	free(buf.v);
	if (fail) atomic_store(&k->fail, 1);
}

int oracle_dilate3d_ours(int nx, int ny, const uint64_t *off, const double *ev, double R,
	uint64_t **out_off, double **out_ev)
{
	int J; double *tab;
	if (oracle_cap_table_ours(R, &J, &tab)) return 1;
	const uint64_t N = (uint64_t)nx * ny;
	list_t *lists = (list_t *)calloc(N ? N : 1, sizeof(list_t));
	if (!lists) { free(tab); return 1; }
	dil_ctx_t k = { nx, ny, J, off, ev, tab, R, lists, 0 };
	parallel_for((int64_t)N, 256, ours_chunk, &k);
	free(tab);
	if (atomic_load(&k.fail)) return 1;
	return pack_lists(lists, N, out_off, out_ev);
}

/* dilation 'brute_force', literal restatement of VoronoiBruteForce.cpp:16-100. */
typedef struct { double z; int w; } bev_t;
static int cmp_bev(const void *a, const void *b)
{
	const bev_t *p = (const bev_t *)a, *q = (const bev_t *)b;
	return (p->z > q->z) - (p->z < q->z);
}

static void brute_chunk(int64_t c0, int64_t c1, void *p)
{
	dil_ctx_t *q = (dil_ctx_t *)p;
	const int nx = q->nx, ny = q->ny;
	const double R = q->R;
	const uint64_t *off = q->off; const double *ev = q->ev;
	bev_t *tmp = NULL; size_t tn = 0, tcap = 0;
	ivbuf_t res = {0, 0, 0};
	int fail = 0;
	for (int64_t c = c0; c < c1 && !fail; ++c) {
		int x = (int)(c % nx), y = (int)(c / nx);
		tn = 0;
		/* VoronoiBruteForce.cpp:39-47: ranges ceil(c-R)..floor(c+R), clamped to the grid */
		for (int ry = (int)ceil(y - R); ry <= (int)floor(y + R) && ry <= ny - 1; ry++) {
			for (int rx = (int)ceil(x - R); rx <= (int)floor(x + R) && rx <= nx - 1; rx++) {
				rx = rx > 0 ? rx : 0;
				ry = ry > 0 ? ry : 0;
				if (pow(rx - x, 2.0) + pow(ry - y, 2.0) <= R * R) {
					double dz = sqrt(R * R - pow(rx - x, 2.0) - pow(ry - y, 2.0));
					uint64_t cc = (uint64_t)rx + (uint64_t)nx * ry;
					for (uint64_t k = off[cc]; k < off[cc + 1]; ++k) {
						if (tn + 2 > tcap) {
							size_t nc = tcap ? 2 * tcap : 256;
							bev_t *nt = (bev_t *)realloc(tmp, nc * sizeof(bev_t));
							if (!nt) { fail = 1; break; }
							tmp = nt; tcap = nc;
						}
						tmp[tn].z = ev[2 * k] - dz; tmp[tn].w = -1; tn++;
						tmp[tn].z = ev[2 * k + 1] + dz; tmp[tn].w = 1; tn++;
					}
				}
			}
		}
		/* VoronoiBruteForce.cpp:71-100: sort by coordinate, running counter */
		res.n = 0;
		if (tn) {
			qsort(tmp, tn, sizeof(bev_t), cmp_bev);
			int flag = tmp[0].w;
			double start = tmp[0].z;
			for (size_t k = 1; k < tn;) {
				while (flag != 0) { flag += tmp[k].w; k++; }
				fail |= ivbuf_push(&res, start, tmp[k - 1].z);
				if (k < tn) { start = tmp[k].z; flag = tmp[k].w; k++; }
			}
		}
		fail |= list_from(&q->lists[c], res.v, res.n);
	}
	free(tmp); free(res.v);
	if (fail) atomic_store(&q->fail, 1);
}

int oracle_dilate3d_brute(int nx, int ny, const uint64_t *off, const double *ev, double R,
	uint64_t **out_off, double **out_ev)
{
	const uint64_t N = (uint64_t)nx * ny;
	list_t *lists = (list_t *)calloc(N ? N : 1, sizeof(list_t));
	if (!lists) return 1;
	dil_ctx_t k = { nx, ny, 0, off, ev, NULL, R, lists, 0 };
	parallel_for((int64_t)N, 256, brute_chunk, &k);
	if (atomic_load(&k.fail)) return 1;
	return pack_lists(lists, N, out_off, out_ev);
}

/* ------------------------------------------------------------------------------------------------
 * Erosion pieces (Voronoi.cpp:18-89).
 * negate: (nx+2)*(ny+2) grid, data shifted by (+1,+1), the four border lines empty, then
 * negate_ray(column, zlo, zhi) (MorphologyOperators.cpp:230-259) on every column including the border.
 * ---------------------------------------------------------------------------------------------- */
static int negate_ray(const double *e, size_t n /* events */, double lo, double hi, ivbuf_t *out)
{
	out->n = 0;
	if (n == 0) return ivbuf_push(out, lo, hi);
	/* event list after the edits: optional lo, e[f..n-l), optional hi */
	size_t f = (e[0] == lo) ? 1 : 0;
	size_t l = (e[n - 1] == hi) ? 1 : 0;
	/* build the event sequence and pair it up */
	size_t total = (n - f - l) + (f ? 0 : 1) + (l ? 0 : 1);
	double *t = (double *)malloc((total ? total : 1) * sizeof(double));
	if (!t) return 1;
	size_t k = 0;
	if (!f) t[k++] = lo;
	for (size_t i = f; i < n - l; ++i) t[k++] = e[i];
	if (!l) t[k++] = hi;
	int rc = 0;
	for (size_t i = 0; i + 1 < k; i += 2) rc |= ivbuf_push(out, t[i], t[i + 1]);
	free(t);
	return rc;
}

int oracle_negate3d(int nx, int ny, const uint64_t *off, const double *ev, double zlo, double zhi,
	uint64_t **out_off, double **out_ev)
{
	const int mx = nx + 2, my = ny + 2;
	const uint64_t M = (uint64_t)mx * my;
	list_t *lists = (list_t *)calloc(M, sizeof(list_t));
	if (!lists) return 1;
	ivbuf_t buf = {0, 0, 0};
	int fail = 0;
	for (int y = 0; y < my; ++y)
		for (int x = 0; x < mx; ++x) {
			uint64_t c = (uint64_t)x + (uint64_t)mx * y;
			if (x >= 1 && x <= nx && y >= 1 && y <= ny) {
				uint64_t s = (uint64_t)(x - 1) + (uint64_t)nx * (y - 1);
				fail |= negate_ray(ev + 2 * off[s], 2 * (size_t)(off[s + 1] - off[s]), zlo, zhi, &buf);
			} else {
				fail |= negate_ray(NULL, 0, zlo, zhi, &buf);
			}
			fail |= list_from(&lists[c], buf.v, buf.n);
		}
	free(buf.v);
	if (fail) return 1;
	return pack_lists(lists, M, out_off, out_ev);
}

/* negate_ray_range (MorphologyOperators.cpp:282-315), literal on the event list. */
static int negate_ray_range(const double *e, size_t n, double lo, double hi, ivbuf_t *out)
{
	out->n = 0;
	if (n == 0) return ivbuf_push(out, lo, hi);
	double *t = (double *)malloc((n + 2) * sizeof(double));
	if (!t) return 1;
	size_t b = 0, len = n; /* live range of e: e[b .. b+len) */
	size_t cf = 0, cl = 0;
	while (len && e[b] <= lo) { b++; len--; cf++; }
	size_t k = 0;
	if (cf % 2 == 0) t[k++] = lo;
	for (size_t i = 0; i < len; ++i) t[k++] = e[b + i];
	/* second loop works on the edited list t[0..k) */
	while (k && t[k - 1] >= hi) { k--; cl++; }
	if (cl % 2 == 0) t[k++] = hi;
	int rc = 0;
	for (size_t i = 0; i + 1 < k; i += 2) rc |= ivbuf_push(out, t[i], t[i + 1]);
	free(t);
	return rc;
}

/* negateInv: input is the (mx*my) bordered grid; output is (mx-2)*(my-2). */
int oracle_negate_inv3d(int mx, int my, const uint64_t *off, const double *ev, double zlo, double zhi,
	uint64_t **out_off, double **out_ev)
{
	const int nx = mx - 2, ny = my - 2;
	if (nx < 0 || ny < 0) return 1;
	const uint64_t N = (uint64_t)nx * ny;
	list_t *lists = (list_t *)calloc(N ? N : 1, sizeof(list_t));
	if (!lists) return 1;
	ivbuf_t buf = {0, 0, 0};
	int fail = 0;
	for (int y = 0; y < ny; ++y)
		for (int x = 0; x < nx; ++x) {
			uint64_t s = (uint64_t)(x + 1) + (uint64_t)mx * (y + 1);
			fail |= negate_ray_range(ev + 2 * off[s], 2 * (size_t)(off[s + 1] - off[s]), zlo, zhi, &buf);
			fail |= list_from(&lists[(uint64_t)x + (uint64_t)nx * y], buf.v, buf.n);
		}
	free(buf.v);
	if (fail) return 1;
	return pack_lists(lists, N, out_off, out_ev);
}

/* op: 0 dilation, 1 erosion, 2 opening, 3 closing; method: 0 ours, 1 brute_force.
 * zmin/zmax as the reference derives them from the volume metadata (VoronoiVorPower.cpp:28-29):
 * zmin = origin_z/spacing, zmax = zmin + 2*padding + extent_z/spacing. */
static int dilate(int method, int nx, int ny, const uint64_t *off, const double *ev, double R,
	uint64_t **oo, double **oe)
{
	return method == 0 ? oracle_dilate3d_ours(nx, ny, off, ev, R, oo, oe)
	                   : oracle_dilate3d_brute(nx, ny, off, ev, R, oo, oe);
}

static int erode(int method, int nx, int ny, const uint64_t *off, const double *ev, double R,
	double zmin, double zmax, uint64_t **oo, double **oe)
{
	/* Voronoi.cpp:8-17 */
	double z_min = zmin - 1, z_max = zmax + 1;
	uint64_t *o1 = NULL, *o2 = NULL; double *e1 = NULL, *e2 = NULL;
	if (oracle_negate3d(nx, ny, off, ev, z_min, z_max, &o1, &e1)) return 1;
	int rc = dilate(method, nx + 2, ny + 2, o1, e1, R, &o2, &e2);
	free(o1); free(e1);
	if (rc) return rc;
	rc = oracle_negate_inv3d(nx + 2, ny + 2, o2, e2, z_min + 1, z_max - 1, oo, oe);
	free(o2); free(e2);
	return rc;
}

int oracle_morph3d(int op, int method, int nx, int ny, const uint64_t *off, const double *ev, double R,
	double zmin, double zmax, uint64_t **out_off, double **out_ev)
{
	if (method != 0 && method != 1) return 2;
	if (op == 0) return dilate(method, nx, ny, off, ev, R, out_off, out_ev);
	if (op == 1) return erode(method, nx, ny, off, ev, R, zmin, zmax, out_off, out_ev);
	uint64_t *o1 = NULL; double *e1 = NULL;
	int rc;
	if (op == 2) { /* opening: erosion then dilation (offset3d.cpp:129-133) */
		if ((rc = erode(method, nx, ny, off, ev, R, zmin, zmax, &o1, &e1))) return rc;
		rc = dilate(method, nx, ny, o1, e1, R, out_off, out_ev);
	} else if (op == 3) { /* closing: dilation then erosion (offset3d.cpp:124-128) */
		if ((rc = dilate(method, nx, ny, off, ev, R, &o1, &e1))) return rc;
		rc = erode(method, nx, ny, o1, e1, R, zmin, zmax, out_off, out_ev);
	} else return 2;
	free(o1); free(e1);
	return rc;
}

/* ------------------------------------------------------------------------------------------------
 * xor (calculate_ray_xor, MorphologyOperators.cpp:334-350; removepoint :364-374; get_volume,
 * CompressedVolume.cpp:61-73). For sorted disjoint lists the chain of negate/union/negate equals the
 * symmetric difference a^b clipped the way negate_ray does; we restate it with the same primitives.
 * ---------------------------------------------------------------------------------------------- */
static int union_two(const iv_t *a, size_t na, const iv_t *b, size_t nb, ivbuf_t *out)
{
	/* unionSegs (MorphologyOperators.cpp:117-158): merge by start, addSegmentAtTheEnd semantics
	 * (strict '<' opens a new interval, hpp:10-18). */
	out->n = 0;
	size_t i = 0, j = 0;
	while (i < na || j < nb) {
		iv_t c;
		if (j < nb && (i >= na || b[j].s <= a[i].s)) c = b[j++]; else c = a[i++];
		if (out->n == 0 || out->v[out->n - 1].e < c.s) { if (ivbuf_push(out, c.s, c.e)) return 1; }
		else if (c.e > out->v[out->n - 1].e) out->v[out->n - 1].e = c.e;
	}
	return 0;
}

int oracle_xor3d(int nx, int ny, const uint64_t *off_a, const double *ev_a, const uint64_t *off_b,
	const double *ev_b, double zmin, double zmax, double spacing, double *volume,
	uint64_t **out_off, double **out_ev)
{
	const uint64_t N = (uint64_t)nx * ny;
	list_t *lists = (list_t *)calloc(N ? N : 1, sizeof(list_t));
	if (!lists) return 1;
	ivbuf_t r1 = {0,0,0}, r2 = {0,0,0}, u1 = {0,0,0}, u2 = {0,0,0}, n1 = {0,0,0}, n2 = {0,0,0}, uu = {0,0,0}, res = {0,0,0};
	int fail = 0;
	double vol = 0.0;
	for (int x = 0; x < nx && !fail; ++x)       /* same column order as CompressedVolume::get_volume */
		for (int y = 0; y < ny && !fail; ++y) {
			uint64_t c = (uint64_t)x + (uint64_t)nx * y;
			const iv_t *a = (const iv_t *)(ev_a + 2 * off_a[c]); size_t na = (size_t)(off_a[c + 1] - off_a[c]);
			const iv_t *b = (const iv_t *)(ev_b + 2 * off_b[c]); size_t nb = (size_t)(off_b[c + 1] - off_b[c]);
			fail |= negate_ray((const double *)a, 2 * na, zmin, zmax, &r1);
			fail |= negate_ray((const double *)b, 2 * nb, zmin, zmax, &r2);
			fail |= union_two(r1.v, r1.n, b, nb, &u1);
			fail |= union_two(r2.v, r2.n, a, na, &u2);
			fail |= negate_ray((const double *)u1.v, 2 * u1.n, zmin, zmax, &n1);
			fail |= negate_ray((const double *)u2.v, 2 * u2.n, zmin, zmax, &n2);
			fail |= union_two(n1.v, n1.n, n2.v, n2.n, &uu);
			res.n = 0;
			for (size_t k = 0; k < uu.n; ++k)
				if (!(uu.v[k].e - uu.v[k].s < 1e-10)) fail |= ivbuf_push(&res, uu.v[k].s, uu.v[k].e);
			for (size_t k = 0; k < res.n; ++k) vol = vol + spacing * spacing * spacing * (res.v[k].e - res.v[k].s);
			fail |= list_from(&lists[c], res.v, res.n);
		}
	free(r1.v); free(r2.v); free(u1.v); free(u2.v); free(n1.v); free(n2.v); free(uu.v); free(res.v);
	if (fail) return 1;
	*volume = vol;
	return pack_lists(lists, N, out_off, out_ev);
}

/* ------------------------------------------------------------------------------------------------
 * 2D (vor2d). Rows are the sweep axis; `width` is DoubleCompressedImage::width() (= m_XSize).
 *   dilate: out(i) = U_{|di| <= floor(R)} [max(0, a-h), min(width, b+h)], h = sqrt(R*R - di*di), candidates
 *           with a' > width or b' < 0 skipped (DoubleVoronoi.cpp:713-725; retire :610,630,646);
 *           R = r * rows (DoubleCompressedImage.cpp:685-686).
 *   erode : seeds are the complement intervals of each row with extremes -1 and width plus a full
 *           sentinel row [-1,width] at rows -1 and `rows` (DoubleVoronoi.h:123-148), R = r
 *           (DoubleCompressedImage.cpp:698-699), then negate() (:438-468).
 * ---------------------------------------------------------------------------------------------- */
static int clamp_push(ivbuf_t *buf, double y1, double y2, double h, double W)
{
	double a = y1 - h > 0 ? y1 - h : 0;
	double b = y2 + h < W ? y2 + h : W;
	if (a > W || b < 0) return 0;
	return ivbuf_push(buf, a, b);
}

static int dilate2d_R(int rows, int width, const uint64_t *off, const double *ev, double R, int complement,
	list_t *lists)
{
	const int J = (int)floor(R);
	const double W = (double)width;
	ivbuf_t buf = {0, 0, 0};
	int fail = 0;
	for (int i = 0; i < rows; ++i) {
		buf.n = 0;
		for (int di = -J; di <= J; ++di) {
			int r = i + di;
			double did = (double)di;
			double h = sqrt(R * R - did * did);
			if (!complement) {
				if (r < 0 || r >= rows) continue;
				for (uint64_t k = off[r]; k < off[r + 1]; ++k) fail |= clamp_push(&buf, ev[2 * k], ev[2 * k + 1], h, W);
			} else {
				if (r < -1 || r > rows) continue;
				if (r == -1 || r == rows) { fail |= clamp_push(&buf, -1.0, W, h, W); continue; }
				/* DoubleVoronoi.h:132-144 */
				double j2 = (double)((float)width * 1.0f);
				for (int64_t k = (int64_t)off[r + 1] - 1; k >= (int64_t)off[r]; --k) {
					fail |= clamp_push(&buf, ev[2 * k + 1], j2, h, W);
					j2 = ev[2 * k];
				}
				fail |= clamp_push(&buf, -1.0, j2, h, W);
			}
		}
		size_t m = union_sorted(buf.v, buf.n);
		fail |= list_from(&lists[i], buf.v, m);
	}
	free(buf.v);
	return fail;
}

static int negate2d_lists(int rows, int width, list_t *lists)
{
	/* DoubleCompressedImage.cpp:438-468 (note `m_XSize*1.0f`) */
	ivbuf_t buf = {0, 0, 0};
	int fail = 0;
	for (int i = 0; i < rows; ++i) {
		size_t n = 2 * lists[i].n;
		const double *e = (const double *)lists[i].v;
		if (n == 0) { buf.n = 0; fail |= ivbuf_push(&buf, 0.0, (double)((float)width * 1.0f)); }
		else fail |= negate_ray(e, n, 0.0, (double)width, &buf);
		free(lists[i].v);
		fail |= list_from(&lists[i], buf.v, buf.n);
	}
	free(buf.v);
	return fail;
}

/* op: 0 dilate, 1 erode, 2 open, 3 close, 4 negate; r exactly as passed to the member function. */
int oracle_morph2d(int op, int rows, int width, const uint64_t *off, const double *ev, double r,
	uint64_t **out_off, double **out_ev)
{
	list_t *lists = (list_t *)calloc(rows ? rows : 1, sizeof(list_t));
	if (!lists) return 1;
	int rc = 0;
	if (op == 0) rc = dilate2d_R(rows, width, off, ev, r * (double)rows, 0, lists);
	else if (op == 1) { rc = dilate2d_R(rows, width, off, ev, r, 1, lists); if (!rc) rc = negate2d_lists(rows, width, lists); }
	else if (op == 4) {
		for (int i = 0; i < rows && !rc; ++i) rc |= list_from(&lists[i], (const iv_t *)(ev + 2 * off[i]), (size_t)(off[i + 1] - off[i]));
		if (!rc) rc = negate2d_lists(rows, width, lists);
	} else if (op == 2 || op == 3) {
		free(lists);
		uint64_t *o1 = NULL; double *e1 = NULL;
		rc = oracle_morph2d(op == 2 ? 1 : 0, rows, width, off, ev, r, &o1, &e1);
		if (rc) return rc;
		rc = oracle_morph2d(op == 2 ? 0 : 1, rows, width, o1, e1, r, out_off, out_ev);
		free(o1); free(e1);
		return rc;
	} else { free(lists); return 2; }
	if (rc) return rc;
	return pack_lists(lists, (uint64_t)rows, out_off, out_ev);
}

/* ------------------------------------------------------------------------------------------------
 * Dexeliser, the step before the path: compute_sign (src/vor3d/Dexelize.cpp:166-225).
 * PARITY UNPINNED: the reference's Dexelize.cpp needs geogram (mesh container, AABB tree), which is
 * absent here, so this restatement cannot be checked against a compiled reference; the reference has
 * no test or golden vector for it either. What is restated, literally:
 *   orientation / point_in_triangle_2d  Dexelize.cpp:58-92 (after SDFGen)
 *   orient_2d_inexact                    Dexelize.cpp:107-121  (a11*a22 - a12*a21, geogram det2x2)
 *   intersect_ray_z                      Dexelize.cpp:136-162  (z = u*p1z + v*p2z + w*p3z; facets
 *                                        whose projection has zero signed area yield nothing)
 *   compute_sign                         Dexelize.cpp:186-212: centre = (x+0.5)*spacing + origin
 *                                        (CompressedVolumeBase.cpp:5-11); the AABB query with a box of
 *                                        zero xy extent hands over exactly the facets whose bounding
 *                                        box contains the centre (geogram bboxes_overlap, inclusive);
 *                                        push z / spacing; std::sort.
 * The reference keeps a list with an odd number of crossings as it is; the CSR containers hold
 * intervals, so the last crossing of such a column is dropped (voroffset_b200/cpp/vo_host.cpp does
 * the same). Columns x0 <= x < x1, y0 <= y < y1 are produced, numbered x-fastest inside that window,
 * every facet is looked at for every column (no spatial index: that is the point of a checker).
 * ---------------------------------------------------------------------------------------------- */
static int dex_orientation(double x1, double y1, double x2, double y2, double *twice_signed_area)
{
	*twice_signed_area = y1 * x2 - x1 * y2;
	if (*twice_signed_area > 0) return 1;
	else if (*twice_signed_area < 0) return -1;
	else if (y2 > y1) return 1;
	else if (y2 < y1) return -1;
	else if (x1 > x2) return 1;
	else if (x1 < x2) return -1;
	else return 0;
}

static int dex_point_in_triangle_2d(double x0, double y0, double x1, double y1, double x2, double y2,
	double x3, double y3, double *a, double *b, double *c)
{
	x1 -= x0; x2 -= x0; x3 -= x0;
	y1 -= y0; y2 -= y0; y3 -= y0;
	int signa = dex_orientation(x2, y2, x3, y3, a);
	if (signa == 0) return 0;
	int signb = dex_orientation(x3, y3, x1, y1, b);
	if (signb != signa) return 0;
	int signc = dex_orientation(x1, y1, x2, y2, c);
	if (signc != signa) return 0;
	double sum = *a + *b + *c;
	*a /= sum; *b /= sum; *c /= sum;
	return 1;
}

static int cmp_double(const void *a, const void *b)
{
	const double p = *(const double *)a, q = *(const double *)b;
	return (p > q) - (p < q);
}

typedef struct {
	uint64_t nv, nf; const double *V; const int32_t *F; const double *fbox; /* [4*nf] bx0,bx1,by0,by1 */
	double ox, oy, spacing; int x0, x1, y0; list_t *lists; atomic_int fail;
} dex_ctx_t;

static void dex_chunk(int64_t c0, int64_t c1, void *p)
{
	dex_ctx_t *k = (dex_ctx_t *)p;
	const int w = k->x1 - k->x0;
	double *z = NULL; size_t zn = 0, zcap = 0;
	int fail = 0;
	for (int64_t c = c0; c < c1 && !fail; ++c) {
		const int x = k->x0 + (int)(c % w), y = k->y0 + (int)(c / w);
		const double cx = (x + 0.5) * k->spacing + k->ox, cy = (y + 0.5) * k->spacing + k->oy;
		zn = 0;
		for (uint64_t f = 0; f < k->nf; ++f) {
			const double *b = k->fbox + 4 * f;
			if (cx < b[0] || cx > b[1] || cy < b[2] || cy > b[3]) continue;   /* the AABB query */
			const double *p1 = k->V + 3 * (uint64_t)k->F[3 * f], *p2 = k->V + 3 * (uint64_t)k->F[3 * f + 1],
			             *p3 = k->V + 3 * (uint64_t)k->F[3 * f + 2];
			double u, v, t;
			if (!dex_point_in_triangle_2d(cx, cy, p1[0], p1[1], p2[0], p2[1], p3[0], p3[1], &u, &v, &t)) continue;
			const double zz = u * p1[2] + v * p2[2] + t * p3[2];
			const double a11 = p2[0] - p1[0], a12 = p2[1] - p1[1], a21 = p3[0] - p1[0], a22 = p3[1] - p1[1];
			const double delta = a11 * a22 - a12 * a21;
			if (delta > 0 || delta < 0) {
				if (zn == zcap) {
					size_t nc = zcap ? 2 * zcap : 32;
					double *nz = (double *)realloc(z, nc * sizeof(double));
					if (!nz) { fail = 1; break; }
					z = nz; zcap = nc;
				}
				z[zn++] = zz / k->spacing;
			}
		}
		qsort(z, zn, sizeof(double), cmp_double);
		list_t *dst = &k->lists[c];
		dst->n = zn / 2; dst->v = NULL;
		if (dst->n) {
			dst->v = (iv_t *)malloc(dst->n * sizeof(iv_t));
			if (!dst->v) { fail = 1; break; }
			for (size_t i = 0; i < dst->n; ++i) { dst->v[i].s = z[2 * i]; dst->v[i].e = z[2 * i + 1]; }
		}
	}
	free(z);
	if (fail) atomic_store(&k->fail, 1);
}

int oracle_dexelize(uint64_t nv, const double *V, uint64_t nf, const int32_t *F, double ox, double oy,
	double spacing, int x0, int x1, int y0, int y1, uint64_t **out_off, double **out_ev)
{
	if (x1 < x0 || y1 < y0 || !(spacing > 0)) return 1;
	for (uint64_t i = 0; i < 3 * nf; ++i) if (F[i] < 0 || (uint64_t)F[i] >= nv) return 1;
	const uint64_t N = (uint64_t)(x1 - x0) * (uint64_t)(y1 - y0);
	list_t *lists = (list_t *)calloc(N ? N : 1, sizeof(list_t));
	double *fbox = (double *)malloc((nf ? nf : 1) * 4 * sizeof(double));
	if (!lists || !fbox) { free(lists); free(fbox); return 1; }
	for (uint64_t f = 0; f < nf; ++f) {
		double *b = fbox + 4 * f;
		b[0] = b[2] = INFINITY; b[1] = b[3] = -INFINITY;
		for (int k = 0; k < 3; ++k) {
			const double *p = V + 3 * (uint64_t)F[3 * f + k];
			if (p[0] < b[0]) b[0] = p[0];
			if (p[0] > b[1]) b[1] = p[0];
			if (p[1] < b[2]) b[2] = p[1];
			if (p[1] > b[3]) b[3] = p[1];
		}
	}
	dex_ctx_t k = { nv, nf, V, F, fbox, ox, oy, spacing, x0, x1, y0, lists, 0 };
	if (N) parallel_for((int64_t)N, 64, dex_chunk, &k);
	free(fbox);
	if (atomic_load(&k.fail)) return 1;
	return pack_lists(lists, N, out_off, out_ev);
}

/* ------------------------------------------------------------------------------------------------
 * 2D ingestion: DoubleCompressedImage::fromImage / scanLine / unionIntersections
 * (src/vor2d/DoubleCompressedImage.cpp:25-111), restated literally. Curve k = points
 * pts[2*coff[k]] .. pts[2*coff[k+1]) as (x, y) = (real, imag). Ray j (0 <= j < h) is the scan line
 * x = j (an INTEGER abscissa, not a pixel centre); it collects the y values where the closed polygon
 * crosses it, sorted per curve; curves are assumed not to overlap: the crossings of a curve are appended
 * when they start at or after the ray's last value and are put in FRONT of the ray otherwise (:88-111).
 * PINNED against the compiled reference (oracle/_ref: ref2d_from_image) by tests/test_from_image.py.
 * ---------------------------------------------------------------------------------------------- */
int oracle_from_image2d(int h, int ncurves, const uint64_t *coff, const double *pts,
	uint64_t **out_off, double **out_ev)
{
	if (h < 0 || ncurves < 0) return 1;
	list_t *lists = (list_t *)calloc(h ? (size_t)h : 1, sizeof(list_t));
	if (!lists) return 1;
	double *ray = NULL, *cross = NULL; size_t rcap = 0, ccap = 0;
	int fail = 0;
	for (int line = 0; line < h && !fail; ++line) {
		size_t rn = 0;
		const double lx = (double)line;
		for (int k = 0; k < ncurves && !fail; ++k) {
			const double *c = pts + 2 * coff[k];
			const int64_t n = (int64_t)(coff[k + 1] - coff[k]);
			size_t cn = 0;
			#define RE(i) c[2 * ((i) % n)]
			#define IM(i) c[2 * ((i) % n) + 1]
			for (int64_t i = 0; i < n && !fail; ++i) {
				double y; int have = 0;
				if (RE(i) < lx) {
					if (RE(i + 1) > lx) {
						const double s = (RE(i + 1) - lx) / (RE(i + 1) - RE(i));
						y = s * IM(i) + (1 - s) * IM(i + 1); have = 1;
					} else {
						int64_t j = 1;
						while (RE(i + j) == lx) j++;
						if (RE(i + j) > lx) { y = IM(i + j - 1); have = 1; }
					}
				} else if (RE(i) > lx) {
					if (RE(i + 1) < lx) {
						const double s = (RE(i + 1) - lx) / (RE(i + 1) - RE(i));
						y = s * IM(i) + (1 - s) * IM(i + 1); have = 1;
					} else {
						int64_t j = 1;
						while (RE(i + j) == lx) j++;
						if (RE(i + j) < lx) { y = IM(i + j - 1); have = 1; }
					}
				}
				if (have) {
					if (cn == ccap) {
						size_t nc = ccap ? 2 * ccap : 32;
						double *nz = (double *)realloc(cross, nc * sizeof(double));
						if (!nz) { fail = 1; break; }
						cross = nz; ccap = nc;
					}
					cross[cn++] = y;
				}
			}
			#undef RE
			#undef IM
			if (fail || cn == 0) continue;
			qsort(cross, cn, sizeof(double), cmp_double);
			if (rn + cn > rcap) {
				size_t nc = 2 * (rn + cn);
				double *nz = (double *)realloc(ray, nc * sizeof(double));
				if (!nz) { fail = 1; break; }
				ray = nz; rcap = nc;
			}
			if (rn == 0 || ray[rn - 1] <= cross[0]) {
				memcpy(ray + rn, cross, cn * sizeof(double));
			} else {
				memmove(ray + cn, ray, rn * sizeof(double));
				memcpy(ray, cross, cn * sizeof(double));
			}
			rn += cn;
		}
		lists[line].n = rn / 2; lists[line].v = NULL;
		if (lists[line].n) {
			lists[line].v = (iv_t *)malloc(lists[line].n * sizeof(iv_t));
			if (!lists[line].v) { fail = 1; break; }
			for (size_t i = 0; i < lists[line].n; ++i) { lists[line].v[i].s = ray[2 * i]; lists[line].v[i].e = ray[2 * i + 1]; }
		}
	}
	free(ray); free(cross);
	if (fail) return 1;
	return pack_lists(lists, (uint64_t)h, out_off, out_ev);
}
