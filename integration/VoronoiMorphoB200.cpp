#include "VoronoiMorphoB200.h"

#include <stdexcept>
#include <string>
#include <vector>

namespace
{
	// the reference reports failures by throwing std::runtime_error (src/vor3d/Common.cpp:7-17)
	void check(vo_ctx *ctx, int rc)
	{
		if (rc != VO_OK) throw std::runtime_error(std::string("voroffset_b200: ") + vo_last_error(ctx));
	}

	vo_ctx *make_ctx(int device)
	{
		vo_ctx *ctx = nullptr;
		if (vo_create(device, &ctx) != VO_OK)
			throw std::runtime_error("voroffset_b200: no usable CUDA device (there is no CPU fallback)");
		return ctx;
	}

	template <typename Lists>
	void to_csr(const Lists &lists, std::vector<uint32_t> &off, std::vector<double> &spans)
	{
		off.assign(lists.size() + 1, 0);
		size_t total = 0;
		for (size_t c = 0; c < lists.size(); ++c) { total += lists[c].size() / 2; off[c + 1] = (uint32_t)total; }
		spans.resize(2 * total + 2);
		size_t w = 0;
		for (const auto &l : lists) for (size_t k = 0; k + 1 < l.size(); k += 2) { spans[w++] = l[k]; spans[w++] = l[k + 1]; }
	}
}

namespace voroffset3d
{
	VoronoiMorphoB200::VoronoiMorphoB200(int method, int device) : m_ctx(make_ctx(device)), m_method(method) {}
	VoronoiMorphoB200::VoronoiMorphoB200(int method, int first_device, int n_gpus) : m_ctx(nullptr), m_method(method)
	{
		if (n_gpus <= 1) { m_ctx = make_ctx(first_device); return; }
		std::vector<int> devs(n_gpus);
		for (int i = 0; i < n_gpus; ++i) devs[i] = first_device + i;
		if (vo_mg_create(devs.data(), n_gpus, &m_mg) != VO_OK)
			throw std::runtime_error("voroffset_b200: cannot create the multi-GPU group (CUDA devices or libnccl.so.2 missing; there is no CPU fallback)");
	}
	VoronoiMorphoB200::~VoronoiMorphoB200() { if (m_mg) vo_mg_destroy(m_mg); else vo_destroy(m_ctx); }

	void VoronoiMorphoB200::run(int op, CompressedVolume &input, CompressedVolume &result, double radius, double &time_1, double &time_2)
	{
		const int nx = input.gridSize()(0), ny = input.gridSize()(1);
		// VoronoiVorPower.cpp:28-29 / Voronoi.cpp:10-11
		const double zmin = input.origin()(2) / input.spacing();
		const double zmax = input.origin()(2) / input.spacing() + 2 * input.padding() + input.extent()(2) / input.spacing();
		std::vector<std::vector<double>> lists((size_t)nx * ny);
		for (int y = 0; y < ny; ++y) for (int x = 0; x < nx; ++x) lists[x + (size_t)nx * y] = input.at(x, y);
		std::vector<uint32_t> off;
		std::vector<double> spans;
		to_csr(lists, off, spans);
		uint32_t *o_off = nullptr;
		double *o_spans = nullptr;
		uint64_t n = 0;
		if (m_mg) {      // y-slabs over several GPUs, NCCL halo exchange inside the library
			if (vo_mg_morph3d(m_mg, op, m_method, nx, ny, zmin, zmax, off.data(), spans.data(), radius, &o_off, &o_spans, &n, &time_1, &time_2) != VO_OK)
				throw std::runtime_error(std::string("voroffset_b200: ") + vo_mg_last_error(m_mg));
		} else
			check(m_ctx, vo_morph3d(m_ctx, op, m_method, nx, ny, zmin, zmax, off.data(), spans.data(), radius,
			                        &o_off, &o_spans, &n, &time_1, &time_2));
		// result.reset(...) exactly like VoronoiVorPower.cpp:37 / VoronoiBruteForce.cpp:20
		result.reset(input.origin(), input.extent(), input.spacing(), input.padding(), nx, ny);
		for (int y = 0; y < ny; ++y)
			for (int x = 0; x < nx; ++x) {
				const size_t c = x + (size_t)nx * y;
				result.at(x, y).assign(o_spans + 2 * (size_t)o_off[c], o_spans + 2 * (size_t)o_off[c + 1]);
			}
		vo_free(o_off);
		vo_free(o_spans);
	}

	void VoronoiMorphoB200::dilation(CompressedVolume input, CompressedVolume &result, double radius, double &time_1, double &time_2)
	{ run(VO_OP_DILATION, input, result, radius, time_1, time_2); }
	void VoronoiMorphoB200::erosion(CompressedVolume input, CompressedVolume &result, double radius, double &time_1, double &time_2)
	{ run(VO_OP_EROSION, input, result, radius, time_1, time_2); }
	void VoronoiMorphoB200::opening(CompressedVolume input, CompressedVolume &result, double radius, double &time_1, double &time_2)
	{ run(VO_OP_OPENING, input, result, radius, time_1, time_2); }
	void VoronoiMorphoB200::closing(CompressedVolume input, CompressedVolume &result, double radius, double &time_1, double &time_2)
	{ run(VO_OP_CLOSING, input, result, radius, time_1, time_2); }
}

namespace voroffset
{
	static void morph2d(DoubleCompressedImage &img, int op, double r, int device)
	{
		vo_ctx *ctx = make_ctx(device);
		std::vector<uint32_t> off;
		std::vector<double> spans;
		to_csr(img.m_Rays, off, spans);
		uint32_t *o_off = nullptr;
		double *o_spans = nullptr;
		uint64_t n = 0;
		double ms = 0;
		int rc = vo_morph2d(ctx, op, img.height(), img.width(), off.data(), spans.data(), r, &o_off, &o_spans, &n, &ms);
		if (rc != VO_OK) { std::string msg = vo_last_error(ctx); vo_destroy(ctx); throw std::runtime_error("voroffset_b200: " + msg); }
		for (int i = 0; i < img.height(); ++i)
			img.m_Rays[i].assign(o_spans + 2 * (size_t)o_off[i], o_spans + 2 * (size_t)o_off[i + 1]);
		vo_free(o_off);
		vo_free(o_spans);
		vo_destroy(ctx);
		vor_assert(img.isValid() == true);       // DoubleCompressedImage.cpp:688,702
	}
	void dilate_b200(DoubleCompressedImage &img, double r, int device) { morph2d(img, VO_OP2D_DILATE, r, device); }
	void erode_b200(DoubleCompressedImage &img, double r, int device) { morph2d(img, VO_OP2D_ERODE, r, device); }
	void close_b200(DoubleCompressedImage &img, double r, int device) { morph2d(img, VO_OP2D_CLOSE, r, device); }
	void open_b200(DoubleCompressedImage &img, double r, int device) { morph2d(img, VO_OP2D_OPEN, r, device); }
}
