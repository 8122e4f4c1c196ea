// TEST INFRASTRUCTURE ONLY - never linked into, imported by, or executed from the product path.
//
// extern "C" driver around the UNMODIFIED reference sources (compiled in place from
// /root/reference/src by oracle/Makefile into oracle/_ref/libvoroffset_ref.so). It marshals flat
// CSR arrays into the reference's own containers and calls the reference's own operators:
//   3D: voroffset3d::VoronoiMorphoVorPower / VoronoiMorphoBruteForce ::dilation / ::erosion
//       (src/vor3d/Voronoi.h:18,29; VoronoiVorPower.cpp:24-96; VoronoiBruteForce.cpp:16-68),
//       opening / closing composed exactly as app/cli3d/offset3d.cpp:124-133 does;
//   3D: VoronoiMorpho::calculateXor (src/vor3d/Voronoi.cpp:91-111);
//   2D: voroffset::DoubleCompressedImage::dilate/erode/open/close/negate
//       (src/vor2d/DoubleCompressedImage.cpp:438-468,680-719).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs load it.
#include "vor3d/CompressedVolume.h"
#include "vor3d/VoronoiVorPower.h"
#include "vor3d/VoronoiBruteForce.h"
#include "vor3d/MorphologyOperators.h"
#include "vor3d/HalfDilationOperator.h"
#include "vor3d/Voronoi2D.h"
#include "vor2d/DoubleCompressedImage.h"
#include "tbb/task_scheduler_init.h"

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>

namespace {

void fill_volume(vor3d::CompressedVolume &vol, int nx, int ny, const double *origin, const double *extent,
	double spacing, int padding, const uint64_t *off, const double *ev)
{
	Eigen::Vector3d o(origin[0], origin[1], origin[2]);
	Eigen::Vector3d e(extent[0], extent[1], extent[2]);
	vol.reset(o, e, spacing, padding, nx, ny);
	for (int y = 0; y < ny; ++y)
		for (int x = 0; x < nx; ++x) {
			const uint64_t c = (uint64_t)x + (uint64_t)nx * y;
			vol.at(x, y).assign(ev + 2 * off[c], ev + 2 * off[c + 1]);
		}
}

int dump_volume(const vor3d::CompressedVolume &vol, int *out_nx, int *out_ny, uint64_t **out_off, double **out_ev)
{
	const int nx = vol.gridSize()(0), ny = vol.gridSize()(1);
	const uint64_t n = (uint64_t)nx * ny;
	uint64_t *off = (uint64_t *)std::malloc((n + 1) * sizeof(uint64_t));
	if (!off) return 1;
	off[0] = 0;
	for (int y = 0; y < ny; ++y)
		for (int x = 0; x < nx; ++x) {
			const uint64_t c = (uint64_t)x + (uint64_t)nx * y;
			off[c + 1] = off[c] + vol.at(x, y).size() / 2;
		}
	double *ev = (double *)std::malloc((2 * off[n] + 2) * sizeof(double));
	if (!ev) { std::free(off); return 1; }
	for (int y = 0; y < ny; ++y)
		for (int x = 0; x < nx; ++x) {
			const uint64_t c = (uint64_t)x + (uint64_t)nx * y;
			const auto &col = vol.at(x, y);
			std::memcpy(ev + 2 * off[c], col.data(), (col.size() / 2) * 2 * sizeof(double));
		}
	*out_nx = nx; *out_ny = ny; *out_off = off; *out_ev = ev;
	return 0;
}

void set_err(char *err, int errlen, const char *msg)
{
	if (err && errlen > 0) { std::snprintf(err, (size_t)errlen, "%s", msg); }
}

} // namespace

extern "C" {

void ref_free(void *p) { std::free(p); }

// op: 0 dilation, 1 erosion, 2 opening, 3 closing.  method: 0 ours, 1 brute_force.
// threads: value handed to the (shimmed) tbb::task_scheduler_init, as offset3d.cpp:98-100 does.
// Intervals: column (x,y) = ev[2*off[x+nx*y] .. 2*off[x+nx*y+1]) as (z1,z2) pairs.
int ref3d_morph(int op, int method, int threads, int nx, int ny, const double *origin, const double *extent,
	double spacing, int padding, const uint64_t *off, const double *ev, double radius,
	int *out_nx, int *out_ny, uint64_t **out_off, double **out_ev, double *time_1, double *time_2,
	char *err, int errlen)
{
	try {
		tbb::task_scheduler_init init(threads > 0 ? threads : 1);
		vor3d::CompressedVolume input, output;
		fill_volume(input, nx, ny, origin, extent, spacing, padding, off, ev);
		std::unique_ptr<vor3d::VoronoiMorpho> m;
		if (method == 0) m = std::make_unique<vor3d::VoronoiMorphoVorPower>();
		else if (method == 1) m = std::make_unique<vor3d::VoronoiMorphoBruteForce>();
		else { set_err(err, errlen, "Invalid method"); return 2; }
		double t1 = 0, t2 = 0;
		if (op == 0) m->dilation(input, output, radius, t1, t2);
		else if (op == 1) m->erosion(input, output, radius, t1, t2);
		else if (op == 2) { vor3d::CompressedVolume tmp; m->erosion(input, tmp, radius, t1, t2); m->dilation(tmp, output, radius, t1, t2); }
		else if (op == 3) { vor3d::CompressedVolume tmp; m->dilation(input, tmp, radius, t1, t2); m->erosion(tmp, output, radius, t1, t2); }
		else { set_err(err, errlen, "Operation"); return 2; }
		if (time_1) *time_1 = t1;
		if (time_2) *time_2 = t2;
		if (dump_volume(output, out_nx, out_ny, out_off, out_ev)) { set_err(err, errlen, "out of memory"); return 3; }
		return 0;
	} catch (const std::exception &e) {
		set_err(err, errlen, e.what());
		return 1;
	}
}

// Number of pieces the reference's intermediate volume (`mid_output`, VoronoiVorPower.cpp:65) holds for
// this input and radius: runs the reference's own first pass - the loop of VoronoiVorPower.cpp:50-57
// (halfDilate forward + backward per x-slice with a VoronoiMorpho2D) followed by its unionMap - and counts.
// This is the k_mid of SURVEY.md section 8(d) that the roofline's algorithmic bytes are defined on.
int ref3d_mid_count(int nx, int ny, const double *origin, const double *extent, double spacing, int padding,
	const uint64_t *off, const double *ev, double radius, uint64_t *pieces, char *err, int errlen)
{
	try {
		vor3d::CompressedVolume input;
		fill_volume(input, nx, ny, origin, extent, spacing, padding, off, ev);
		double zmin = input.origin()(2) / input.spacing();
		double zmax = zmin + 2 * input.padding() + input.extent()(2) / input.spacing();
		vor3d::CompressedVolumeWithRadii o1, o2, mid;
		o1.reshape(nx, ny); o2.reshape(nx, ny); mid.reshape(nx, ny);
		vor3d::VoronoiMorpho2D op_x(ny, zmin, zmax, radius, input.spacing());
		for (int x = 0; x < nx; x++) {
			vor3d::halfDilate(op_x, true, input, o1, x, 0, 0, +1);
			op_x.resetData();
			vor3d::halfDilate(op_x, true, input, o2, x, ny - 1, 0, -1);
			op_x.resetData();
		}
		vor3d::unionMap(o1, o2, mid);
		uint64_t n = 0;
		for (int y = 0; y < ny; ++y) for (int x = 0; x < nx; ++x) n += mid.at(x, y).size();
		*pieces = n;
		return 0;
	} catch (const std::exception &e) {
		set_err(err, errlen, e.what());
		return 1;
	}
}

// Symmetric-difference volume of two same-grid volumes (Voronoi.cpp:91-111); also returns the xor volume.
int ref3d_xor(int nx, int ny, const double *origin, const double *extent, double spacing, int padding,
	const uint64_t *off_a, const double *ev_a, const uint64_t *off_b, const double *ev_b,
	double *volume, uint64_t **out_off, double **out_ev, char *err, int errlen)
{
	try {
		vor3d::CompressedVolume a, b, r;
		fill_volume(a, nx, ny, origin, extent, spacing, padding, off_a, ev_a);
		fill_volume(b, nx, ny, origin, extent, spacing, padding, off_b, ev_b);
		vor3d::VoronoiMorphoBruteForce m; // calculateXor lives in the base class
		*volume = m.calculateXor(a, b, r);
		int onx, ony;
		if (dump_volume(r, &onx, &ony, out_off, out_ev)) { set_err(err, errlen, "out of memory"); return 3; }
		return 0;
	} catch (const std::exception &e) {
		set_err(err, errlen, e.what());
		return 1;
	}
}

// op: 0 dilate, 1 erode, 2 open, 3 close, 4 negate. Row i = ev[2*off[i] .. 2*off[i+1]).
// `r` is passed to the member function untouched (dilate multiplies it by the row count itself,
// DoubleCompressedImage.cpp:685-686).
int ref2d_morph(int op, int rows, int width, const uint64_t *off, const double *ev, double r,
	uint64_t **out_off, double **out_ev, char *err, int errlen)
{
	try {
		voroffset::DoubleCompressedImage img(width, rows);
		for (int i = 0; i < rows; ++i) img.m_Rays[i].assign(ev + 2 * off[i], ev + 2 * off[i + 1]);
		if (op == 0) img.dilate(r);
		else if (op == 1) img.erode(r);
		else if (op == 2) img.open(r);
		else if (op == 3) img.close(r);
		else if (op == 4) img.negate();
		else { set_err(err, errlen, "Operation"); return 2; }
		uint64_t *o = (uint64_t *)std::malloc(((size_t)rows + 1) * sizeof(uint64_t));
		if (!o) { set_err(err, errlen, "out of memory"); return 3; }
		o[0] = 0;
		for (int i = 0; i < rows; ++i) o[i + 1] = o[i] + img.m_Rays[i].size() / 2;
		double *e = (double *)std::malloc((2 * o[rows] + 2) * sizeof(double));
		if (!e) { std::free(o); set_err(err, errlen, "out of memory"); return 3; }
		for (int i = 0; i < rows; ++i)
			std::memcpy(e + 2 * o[i], img.m_Rays[i].data(), (img.m_Rays[i].size() / 2) * 2 * sizeof(double));
		*out_off = o; *out_ev = e;
		return 0;
	} catch (const std::exception &e) {
		set_err(err, errlen, e.what());
		return 1;
	}
}

// DoubleCompressedImage::fromImage (src/vor2d/DoubleCompressedImage.cpp:25-111) on `ncurves` closed polygons:
// curve k = points pts[2*coff[k]] .. pts[2*coff[k+1]) as (x, y) pairs = (real, imag) of the reference's PointF.
// The image is constructed as src/vor2d/Dexelize.cpp:42 does: DoubleCompressedImage(w, h) with h rays.
int ref2d_from_image(int w, int h, int ncurves, const uint64_t *coff, const double *pts,
	uint64_t **out_off, double **out_ev, char *err, int errlen)
{
	try {
		std::vector<voroffset::Curve> curves((size_t)ncurves);
		for (int k = 0; k < ncurves; ++k)
			for (uint64_t i = coff[k]; i < coff[k + 1]; ++i) curves[k].push_back(voroffset::PointF(pts[2 * i], pts[2 * i + 1]));
		voroffset::DoubleCompressedImage img(w, h);
		img.fromImage(curves);
		const int rows = img.height();
		uint64_t *o = (uint64_t *)std::malloc(((size_t)rows + 1) * sizeof(uint64_t));
		if (!o) { set_err(err, errlen, "out of memory"); return 3; }
		o[0] = 0;
		for (int i = 0; i < rows; ++i) o[i + 1] = o[i] + img.m_Rays[i].size() / 2;
		double *e = (double *)std::malloc((2 * o[rows] + 2) * sizeof(double));
		if (!e) { std::free(o); set_err(err, errlen, "out of memory"); return 3; }
		for (int i = 0; i < rows; ++i)
			std::memcpy(e + 2 * o[i], img.m_Rays[i].data(), (img.m_Rays[i].size() / 2) * 2 * sizeof(double));
		*out_off = o; *out_ev = e;
		return 0;
	} catch (const std::exception &e) {
		set_err(err, errlen, e.what());
		return 1;
	}
}

// DoubleCompressedImage::transposeInPlace (src/vor2d/DoubleCompressedImage.cpp:478-584); the result has `width`
// rays of length `rows`.
int ref2d_transpose(int rows, int width, const uint64_t *off, const double *ev,
	uint64_t **out_off, double **out_ev, char *err, int errlen)
{
	try {
		voroffset::DoubleCompressedImage img(width, rows);
		for (int i = 0; i < rows; ++i) img.m_Rays[i].assign(ev + 2 * off[i], ev + 2 * off[i + 1]);
		img.transposeInPlace();
		const int n = img.height();
		uint64_t *o = (uint64_t *)std::malloc(((size_t)n + 1) * sizeof(uint64_t));
		if (!o) { set_err(err, errlen, "out of memory"); return 3; }
		o[0] = 0;
		for (int i = 0; i < n; ++i) o[i + 1] = o[i] + img.m_Rays[i].size() / 2;
		double *e = (double *)std::malloc((2 * o[n] + 2) * sizeof(double));
		if (!e) { std::free(o); set_err(err, errlen, "out of memory"); return 3; }
		for (int i = 0; i < n; ++i)
			std::memcpy(e + 2 * o[i], img.m_Rays[i].data(), (img.m_Rays[i].size() / 2) * 2 * sizeof(double));
		*out_off = o; *out_ev = e;
		return 0;
	} catch (const std::exception &e) {
		set_err(err, errlen, e.what());
		return 1;
	}
}

} // extern "C"
