// Drop-in demonstration (test infrastructure): the reference's operators and the B200-backed subclasses
// are driven through the SAME base-class pointer (std::unique_ptr<vor3d::VoronoiMorpho>, as in
// app/cli3d/offset3d.cpp:104-136) on the same reference CompressedVolume, and the results are compared
// column by column. Built by oracle/Makefile (`dropin`) from the reference's own sources where
// /root/reference exists; the binary travels to the GPU box in oracle/_ref/.
#include "VoronoiMorphoB200.h"
#include "vor3d/VoronoiVorPower.h"
#include "vor3d/VoronoiBruteForce.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>

static vor3d::CompressedVolume make_blobs(int n, int padding)
{
	// union of a few balls, rasterised analytically at column centres (generic position by construction)
	const double sp = 1.0 / n;
	vor3d::CompressedVolume vol(Eigen::Vector3d(0, 0, 0), Eigen::Vector3d(1, 1, 1), sp, padding);
	const double c[5][4] = {{0.35, 0.4, 0.45, 0.21}, {0.62, 0.55, 0.5, 0.17}, {0.5, 0.3, 0.7, 0.12}, {0.3, 0.7, 0.3, 0.14}, {0.72, 0.28, 0.33, 0.09}};
	const int nx = vol.gridSize()(0), ny = vol.gridSize()(1);
	for (int y = 0; y < ny; ++y)
		for (int x = 0; x < nx; ++x) {
			const double px = (x + 0.5 + 0.0137) * sp + vol.origin()(0), py = (y + 0.5 + 0.0071) * sp + vol.origin()(1);
			std::vector<std::pair<double, double>> iv;
			for (auto &b : c) {
				const double d2 = b[3] * b[3] - (px - b[0]) * (px - b[0]) - (py - b[1]) * (py - b[1]);
				if (d2 > 0) iv.push_back({(b[2] - std::sqrt(d2)) / sp, (b[2] + std::sqrt(d2)) / sp});
			}
			std::sort(iv.begin(), iv.end());
			for (auto &s : iv) vol.appendSegment(x, y, s.first, s.second, 0);
		}
	return vol;
}

static int compare(const vor3d::CompressedVolume &a, const vor3d::CompressedVolume &b, bool bitwise, double tol, const char *what)
{
	if (a.gridSize()(0) != b.gridSize()(0) || a.gridSize()(1) != b.gridSize()(1)) { std::printf("%s: grid differs\n", what); return 1; }
	size_t diff_topo = 0, diff_bits = 0;
	double maxd = 0;
	for (int y = 0; y < a.gridSize()(1); ++y)
		for (int x = 0; x < a.gridSize()(0); ++x) {
			const auto &p = a.at(x, y);
			const auto &q = b.at(x, y);
			if (p.size() != q.size()) { ++diff_topo; continue; }
			for (size_t k = 0; k < p.size(); ++k) {
				if (std::memcmp(&p[k], &q[k], sizeof(double)) != 0) ++diff_bits;
				maxd = std::max(maxd, std::fabs(p[k] - q[k]));
			}
		}
	const bool ok = diff_topo == 0 && (bitwise ? diff_bits == 0 : maxd <= tol);
	std::printf("%-34s topology mismatches %zu, differing endpoints %zu, max |dz| %.3g  -> %s\n", what, diff_topo, diff_bits, maxd, ok ? "OK" : "FAIL");
	return ok ? 0 : 1;
}

int main(int argc, char **argv)
{
	const int n = argc > 1 ? std::atoi(argv[1]) : 64;
	const double radius = argc > 2 ? std::atof(argv[2]) : 5.5;
	int bad = 0;
	try {
		vor3d::CompressedVolume input = make_blobs(n, (int)std::ceil(radius) + 1);
		std::printf("grid %d x %d, %d segments, radius %g\n", input.gridSize()(0), input.gridSize()(1), input.numSegments(), radius);
		for (const std::string method : {"ours", "brute_force"}) {
			std::unique_ptr<vor3d::VoronoiMorpho> ref, gpu;            // offset3d.cpp:104-112
			if (method == "ours") { ref = std::make_unique<vor3d::VoronoiMorphoVorPower>(); gpu = std::make_unique<vor3d::VoronoiMorphoVorPowerB200>(); }
			else { ref = std::make_unique<vor3d::VoronoiMorphoBruteForce>(); gpu = std::make_unique<vor3d::VoronoiMorphoBruteForceB200>(); }
			for (const std::string op : {"dilation", "erosion", "closing", "opening"}) {
				vor3d::CompressedVolume out_ref, out_gpu, tmp;
				double t1 = 0, t2 = 0, g1 = 0, g2 = 0;
				for (int side = 0; side < 2; ++side) {                 // offset3d.cpp:116-136, verbatim structure
					auto &m = side ? gpu : ref;
					auto &out = side ? out_gpu : out_ref;
					double &a = side ? g1 : t1, &b = side ? g2 : t2;
					if (op == "erosion") m->erosion(input, out, radius, a, b);
					else if (op == "dilation") m->dilation(input, out, radius, a, b);
					else if (op == "closing") { m->dilation(input, tmp, radius, a, b); m->erosion(tmp, out, radius, a, b); }
					else { m->erosion(input, tmp, radius, a, b); m->dilation(tmp, out, radius, a, b); }
				}
				const bool prim = op == "dilation" || op == "erosion";
				bad += compare(out_ref, out_gpu, prim || method == "brute_force", 1e-11, (method + " " + op).c_str());
			}
		}
		// 2D
		voroffset::DoubleCompressedImage a(96, 64), b;
		for (int i = 0; i < 64; ++i) {
			const double w = 20 + 15 * std::sin(0.37 * i + 0.1);
			a.m_Rays[i] = {48 - w + 0.013 * i, 48 - 0.3 * w, 48 + 0.2 * w + 0.007 * i, 48 + w};
		}
		b.copyFrom(a);
		voroffset::DoubleCompressedImage c1, c2;
		c1.copyFrom(a); c2.copyFrom(a);
		c1.dilate(4.5 / 64); voroffset::dilate_b200(c2, 4.5 / 64);
		bool same = c1.m_Rays == c2.m_Rays;
		c1.erode(2.5); voroffset::erode_b200(c2, 2.5);
		same = same && c1.m_Rays == c2.m_Rays;
		std::printf("%-34s %s\n", "2D dilate + erode", same ? "bit-identical -> OK" : "FAIL");
		bad += same ? 0 : 1;
	} catch (const std::exception &e) {
		std::printf("exception: %s\n", e.what());
		return 2;
	}
	return bad ? 1 : 0;
}
