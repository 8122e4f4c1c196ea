// Reference-side binding: what a voroffset maintainer adds to the reference tree to route the hot path
// through libvoroffset_b200.so. Two subclasses of the reference's own operator interface
// (src/vor3d/Voronoi.h:13-44) with the SAME virtual signature, selected by method name exactly like
// app/cli3d/offset3d.cpp:104-112, and four free functions mirroring DoubleCompressedImage's operators.
//
// This file includes the REFERENCE's headers (it is compiled inside the reference's build, or - in this
// repo - by oracle/Makefile's `dropin` target against /root/reference/src). Nothing here is compute:
// it marshals vector<vector<double>> <-> flat CSR and calls the C ABI (include/voroffset_b200.h).
#pragma once
#include "vor3d/Voronoi.h"
#include "vor2d/DoubleCompressedImage.h"
#include "voroffset_b200.h"

namespace voroffset3d
{
	// Shared plumbing of the two GPU-backed operators.
	class VoronoiMorphoB200 : public VoronoiMorpho
	{
	public:
		explicit VoronoiMorphoB200(int method, int device = 0);
		// n_gpus > 1: the grid is cut into y-slabs over the GPUs first_device .. first_device + n_gpus - 1 (vo_mg_*)
		VoronoiMorphoB200(int method, int first_device, int n_gpus);
		~VoronoiMorphoB200() override;
		VoronoiMorphoB200(const VoronoiMorphoB200 &) = delete;
		VoronoiMorphoB200 &operator=(const VoronoiMorphoB200 &) = delete;

		// Voronoi.h:18 - same signature, same out-parameter meaning (time_1/time_2 in ms)
		void dilation(CompressedVolume input, CompressedVolume &result, double radius, double &time_1, double &time_2) override;
		// Voronoi.h:29 - the reference composes erosion on the host (Voronoi.cpp:8-17); here the complement,
		// the dilation and the inverse complement all run on the device in one call
		void erosion(CompressedVolume input, CompressedVolume &result, double radius, double &time_1, double &time_2) override;
		// offset3d.cpp:124-133 compositions without leaving HBM between the two primitives
		void opening(CompressedVolume input, CompressedVolume &result, double radius, double &time_1, double &time_2);
		void closing(CompressedVolume input, CompressedVolume &result, double radius, double &time_1, double &time_2);

	private:
		void run(int op, CompressedVolume &input, CompressedVolume &result, double radius, double &time_1, double &time_2);
		vo_ctx *m_ctx;
		vo_mg *m_mg = nullptr;
		int m_method;
	};

	// "ours" (VoronoiVorPower.h) and "brute_force" (VoronoiBruteForce.h) on the GPU
	class VoronoiMorphoVorPowerB200 : public VoronoiMorphoB200
	{
	public:
		explicit VoronoiMorphoVorPowerB200(int device = 0) : VoronoiMorphoB200(VO_METHOD_OURS, device) {}
		VoronoiMorphoVorPowerB200(int first_device, int n_gpus) : VoronoiMorphoB200(VO_METHOD_OURS, first_device, n_gpus) {}
	};
	class VoronoiMorphoBruteForceB200 : public VoronoiMorphoB200
	{
	public:
		explicit VoronoiMorphoBruteForceB200(int device = 0) : VoronoiMorphoB200(VO_METHOD_BRUTE_FORCE, device) {}
		VoronoiMorphoBruteForceB200(int first_device, int n_gpus) : VoronoiMorphoB200(VO_METHOD_BRUTE_FORCE, first_device, n_gpus) {}
	};
}

namespace voroffset
{
	// DoubleCompressedImage.h:108-111, in place like the member functions; `r` untouched
	// (dilate sweeps with r*rows inside the library, DoubleCompressedImage.cpp:685-686).
	void dilate_b200(DoubleCompressedImage &img, double r, int device = 0);
	void erode_b200(DoubleCompressedImage &img, double r, int device = 0);
	void close_b200(DoubleCompressedImage &img, double r, int device = 0);
	void open_b200(DoubleCompressedImage &img, double r, int device = 0);
}
