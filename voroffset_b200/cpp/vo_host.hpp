// Host-side C++ mirror of the reference's containers and operator interface, dependency-free
// (no Eigen / geogram / CLI11 / json): what the re-hosted offset3d / offset2d executables are built on.
//
//   voroffset3d::CompressedVolume      <- src/vor3d/CompressedVolume.{h,cpp}, CompressedVolumeBase.{h,cpp}
//   voroffset3d::VoronoiMorpho (+ VorPower / BruteForce)  <- src/vor3d/Voronoi.h:13-44, VoronoiVorPower.h,
//                                         VoronoiBruteForce.h; every operator goes through the C ABI
//   voroffset::DoubleCompressedImage   <- src/vor2d/DoubleCompressedImage.{h,cpp} (storage, save/load,
//                                         isValid, negate/dilate/erode/open/close through the C ABI)
// Same member names and argument meaning as the reference; Eigen vectors become std::array.
#pragma once
#include <array>
#include <complex>
#include <cmath>
#include <cstdint>
#include <functional>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "voroffset_b200.h"

namespace voroffset3d
{
	typedef double Scalar;
	typedef std::array<double, 3> Vector3d;
	typedef std::array<int, 2> Vector2i;

	class CompressedVolume
	{
		Vector3d m_Origin{{0, 0, 0}}, m_Extent{{0, 0, 0}};
		Scalar m_Spacing = 1;
		int m_Padding = 0;
		Vector2i m_GridSize{{0, 0}};
		std::vector<std::vector<Scalar>> m_Data;

	public:
		CompressedVolume() = default;
		// CompressedVolume.cpp:11-23
		CompressedVolume(Vector3d origin, Vector3d extent, Scalar voxel_size, int padding)
			: m_Origin(origin), m_Extent(extent), m_Spacing(voxel_size), m_Padding(padding)
		{
			for (auto &o : m_Origin) o -= padding * voxel_size;
			m_GridSize[0] = (int)(std::ceil(extent[0] / m_Spacing) + 2 * padding);
			m_GridSize[1] = (int)(std::ceil(extent[1] / m_Spacing) + 2 * padding);
			m_Data.assign((size_t)m_GridSize[0] * m_GridSize[1], {});
		}
		// CompressedVolumeBase.cpp:13-21
		void reset(Vector3d origin, Vector3d extent, Scalar voxel_size, int padding, int xsize, int ysize)
		{
			m_Origin = origin; m_Extent = extent; m_Spacing = voxel_size; m_Padding = padding;
			reshape(xsize, ysize);
		}
		void reshape(int xsize, int ysize) { m_GridSize = {{xsize, ysize}}; m_Data.assign((size_t)xsize * ysize, {}); }
		void resize(int xsize, int ysize) { m_GridSize = {{xsize, ysize}}; m_Data.resize((size_t)xsize * ysize); }

		int numDexels() const { return m_GridSize[0] * m_GridSize[1]; }
		Vector2i gridSize() const { return m_GridSize; }
		Vector3d origin() const { return m_Origin; }
		Vector3d extent() const { return m_Extent; }
		double spacing() const { return m_Spacing; }
		int padding() const { return m_Padding; }
		std::array<double, 2> dexelCenter(int x, int y) const   // CompressedVolumeBase.cpp:5-11
		{
			return {{(x + 0.5) * m_Spacing + m_Origin[0], (y + 0.5) * m_Spacing + m_Origin[1]}};
		}
		const std::vector<Scalar> &at(int x, int y) const { return m_Data[x + (size_t)m_GridSize[0] * y]; }
		std::vector<Scalar> &at(int x, int y) { return m_Data[x + (size_t)m_GridSize[0] * y]; }

		void iterate(int i, int j, std::function<void(Scalar, Scalar)> func) const
		{
			const auto &ray = at(i, j);
			for (size_t k = 0; k + 1 < ray.size(); k += 2) func(ray[k], ray[k + 1]);
		}
		int numSegments() const
		{
			size_t n = 0;
			for (const auto &r : m_Data) n += r.size() / 2;
			return (int)n;
		}
		double get_volume() const                                  // CompressedVolume.cpp:61-73, same column order
		{
			double vol = 0;
			for (int x = 0; x < m_GridSize[0]; x++)
				for (int y = 0; y < m_GridSize[1]; y++)
					iterate(x, y, [&](Scalar a, Scalar b) { vol = vol + m_Spacing * m_Spacing * m_Spacing * (b - a); });
			return vol;
		}
		// zmin / zmax as VoronoiVorPower.cpp:28-29
		double zmin() const { return m_Origin[2] / m_Spacing; }
		double zmax() const { return m_Origin[2] / m_Spacing + 2 * m_Padding + m_Extent[2] / m_Spacing; }

		void save(std::ostream &out) const;    // CompressedVolume.cpp:133-152
		void load(std::istream &in);           // CompressedVolume.cpp:116-131

		// flat CSR <-> nested vectors (the layout of include/voroffset_b200.h)
		void to_csr(std::vector<uint32_t> &off, std::vector<double> &spans) const;
		void from_csr(const uint32_t *off, const double *spans);
	};

	// src/vor3d/Voronoi.h:13-44. Errors surface as std::runtime_error (Common.cpp:7-17).
	class VoronoiMorpho
	{
	public:
		explicit VoronoiMorpho(int method, int device = 0);
		// n_gpus > 1: the grid is cut into y-slabs over the GPUs 0 .. n_gpus-1 (vo_mg_*: NCCL halo exchange inside the
		// library); every operator below then runs on all of them
		VoronoiMorpho(int method, int first_device, int n_gpus);
		virtual ~VoronoiMorpho();
		VoronoiMorpho(const VoronoiMorpho &) = delete;
		VoronoiMorpho &operator=(const VoronoiMorpho &) = delete;
		virtual void dilation(CompressedVolume input, CompressedVolume &result, double radius, double &time_1, double &time_2);
		virtual void erosion(CompressedVolume input, CompressedVolume &result, double radius, double &time_1, double &time_2);
		void opening(CompressedVolume input, CompressedVolume &result, double radius, double &time_1, double &time_2);
		void closing(CompressedVolume input, CompressedVolume &result, double radius, double &time_1, double &time_2);
		double calculateXor(CompressedVolume voxel_1, CompressedVolume voxel_2, CompressedVolume &result);

	protected:
		void run(int op, const CompressedVolume &input, CompressedVolume &result, double radius, double &t1, double &t2);
		vo_ctx *m_ctx;
		vo_mg *m_mg = nullptr;
		int m_method;
	};
	class VoronoiMorphoVorPower : public VoronoiMorpho
	{
	public:
		explicit VoronoiMorphoVorPower(int device = 0) : VoronoiMorpho(VO_METHOD_OURS, device) {}
		VoronoiMorphoVorPower(int first_device, int n_gpus) : VoronoiMorpho(VO_METHOD_OURS, first_device, n_gpus) {}
	};
	class VoronoiMorphoBruteForce : public VoronoiMorpho
	{
	public:
		explicit VoronoiMorphoBruteForce(int device = 0) : VoronoiMorpho(VO_METHOD_BRUTE_FORCE, device) {}
		VoronoiMorphoBruteForce(int first_device, int n_gpus) : VoronoiMorpho(VO_METHOD_BRUTE_FORCE, first_device, n_gpus) {}
	};

	// Geogram-free restatement of src/vor3d/Dexelize.cpp (mesh -> dexels, dexels -> hex mesh / points).
	// device >= 0: the ray-marching loop runs on that GPU (vo_dexelize_dev); device < 0: host loop (offset3d -x noop,
	// the one mode that needs no GPU). Both give the same bits.
	CompressedVolume create_dexels(const std::string &filename, double &voxel_size, int padding = 0, int num_voxels = -1, int device = -1);
	void dexel_dump(const std::string &filename, const CompressedVolume &voxels);
}

namespace voroffset
{
	typedef double Scalar;
	typedef std::complex<double> PointF;    // src/vor2d/Common.h:15-16
	typedef std::vector<PointF> Curve;

	class DoubleCompressedImage
	{
		int m_XSize = 0;

	public:
		std::vector<std::vector<Scalar>> m_Rays;

		DoubleCompressedImage(int w = 0, int h = 0) : m_XSize(w), m_Rays(h) {}
		explicit DoubleCompressedImage(std::istream &in) { load(in); }
		int width() const { return m_XSize; }
		int height() const { return (int)m_Rays.size(); }
		void resize(int w, int h) { m_XSize = w; m_Rays.assign(h, {}); }
		void fromImage(const std::vector<Curve> &input_curves);   // DoubleCompressedImage.cpp:25-40 (vo_svg.cpp)
		void transposeInPlace();               // DoubleCompressedImage.cpp:478-584 (vo_svg.cpp)
		bool isValid() const;                  // DoubleCompressedImage.cpp:197-223
		void save(std::ostream &out) const;    // DoubleCompressedImage.cpp:145-160
		void load(std::istream &in);           // DoubleCompressedImage.cpp:162-183
		// DoubleCompressedImage.cpp:438-468, 680-719: in place, `r` as the reference takes it
		void negate();
		void dilate(double r);
		void erode(double r);
		void close(double r);
		void open(double r);

	private:
		void apply(int op, double r);
		void scanLine(std::vector<Scalar> &intersections, int line_x, const Curve &curve);          // DoubleCompressedImage.cpp:43-86
		void unionIntersections(std::vector<Scalar> &intersections_1, const std::vector<Scalar> &intersections_2);   // :88-111
	};

	// src/vor2d/Dexelize.cpp:22-47 without nanosvg (vo_svg.cpp): SVG -> contours (mm at 90 DPI) -> fromImage
	std::vector<Curve> svg_contours(const std::string &file, double &width_mm, double &height_mm);
	DoubleCompressedImage create_dexels(const std::string &file);
	void dexel_dump(const std::string &filename, const DoubleCompressedImage &dexels);   // src/vor2d/Dexelize.cpp:48-89, OBJ
}
