// offset2d re-hosted on voroffset_b200: the flags of app/cli2d/offset2d.cpp:30-39 (-i -o -r -e -f -t -n).
// Input: an SVG file like the reference (".svg": vo_svg.cpp reads it with nanosvg's conventions and rasterises with
// the restated DoubleCompressedImage::fromImage), the reference's own dexel text format
// (DoubleCompressedImage::save/load, ".dex") or a plain polygon file (".poly": first line "width height",
// then one closed polygon per line as "x0 y0 x1 y1 ..." in pixels), scan-converted at row centres.
// The morphology runs on the GPU through the C ABI.
#include "vo_host.hpp"

#include <algorithm>
#include <cstring>
#include <fstream>
#include <sstream>

namespace vor = voroffset;

static bool ends(const std::string &s, const char *suf) { const size_t n = std::strlen(suf); return s.size() >= n && s.compare(s.size() - n, n, suf) == 0; }

static vor::DoubleCompressedImage load_poly(const std::string &file)
{
	std::ifstream in(file);
	int w = 0, h = 0;
	in >> w >> h;
	std::string line;
	std::getline(in, line);
	vor::DoubleCompressedImage img(w, h);
	std::vector<std::vector<std::pair<double, double>>> rows(h);
	while (std::getline(in, line)) {
		std::istringstream ss(line);
		std::vector<std::pair<double, double>> p;
		double x, y;
		while (ss >> x >> y) p.push_back({x, y});
		if (p.size() < 3) continue;
		for (int i = 0; i < h; ++i) {
			const double yy = i + 0.5;
			std::vector<double> xs;
			for (size_t k = 0; k < p.size(); ++k) {
				const auto &a = p[k], &b = p[(k + 1) % p.size()];
				if ((a.second <= yy) != (b.second <= yy)) xs.push_back(a.first + (yy - a.second) / (b.second - a.second) * (b.first - a.first));
			}
			std::sort(xs.begin(), xs.end());
			for (size_t k = 0; k + 1 < xs.size(); k += 2) rows[i].push_back({std::max(0.0, xs[k]), std::min((double)w, xs[k + 1])});
		}
	}
	for (int i = 0; i < h; ++i) {
		std::sort(rows[i].begin(), rows[i].end());
		for (auto &s : rows[i]) {
			if (s.second <= s.first) continue;
			auto &r = img.m_Rays[i];
			if (!r.empty() && s.first <= r.back()) r.back() = std::max(r.back(), s.second);
			else { r.push_back(s.first); r.push_back(s.second); }
		}
	}
	return img;
}

int main(int argc, char *argv[])
{
	struct { std::string input, output = "out.dex"; double radius = 0; bool erode = false, force = false, transpose = false, negate = false; } args;
	int positional = 0;
	for (int i = 1; i < argc; ++i) {
		const std::string a = argv[i];
		auto val = [&]() -> std::string { if (i + 1 >= argc) { std::cerr << a << " needs a value\n"; std::exit(1); } return argv[++i]; };
		if (a == "-i" || a == "--input") args.input = val();
		else if (a == "-o" || a == "--output") args.output = val();
		else if (a == "-r" || a == "--radius") args.radius = std::stod(val());
		else if (a == "-e" || a == "--erode") args.erode = true;
		else if (a == "-f" || a == "--force") args.force = true;
		else if (a == "-t" || a == "--transpose") args.transpose = true;
		else if (a == "-n" || a == "--negate") args.negate = true;
		else if (a == "-h" || a == "--help") { std::cout << "Offset2D\nUsage: offset2d [-i] input [-o output] [-r radius] [-e] [-f] [-t] [-n]\n"; return 0; }
		else if (positional++ == 0) args.input = a;
		else { std::cerr << "unexpected argument " << a << "\n"; return 1; }
	}
	if (args.input.empty() || !std::ifstream(args.input)) { std::cerr << "input: File does not exist\n"; return 1; }
	try {
		vor::DoubleCompressedImage dexels;
		if (ends(args.input, ".svg")) dexels = vor::create_dexels(args.input);           // offset2d.cpp:47
		else if (ends(args.input, ".poly")) dexels = load_poly(args.input);
		else { std::ifstream in(args.input); dexels.load(in); }
		if (args.transpose) dexels.transposeInPlace();                      // offset2d.cpp:50-52
		if (args.negate) dexels.negate();                           // offset2d.cpp:53-55
		if (args.radius > 0) {                                      // offset2d.cpp:58-66
			std::cout << "-- Performing offset by radius r = " << args.radius << std::endl;
			if (args.erode) dexels.erode(args.radius);
			else dexels.dilate(args.radius);
		}
		if (std::ifstream(args.output) && !args.force) {
			std::cerr << "-- Output file already exists. Please use -f to force overwriting." << std::endl;
		} else {
			std::cout << "-- Saving" << std::endl;
			if (ends(args.output, ".obj")) vor::dexel_dump(args.output, dexels);   // offset2d.cpp:69-81
			else { std::ofstream out(args.output); dexels.save(out); }
		}
	} catch (const std::exception &e) {
		std::cerr << "error: " << e.what() << std::endl;
		return 1;
	}
	return 0;
}
