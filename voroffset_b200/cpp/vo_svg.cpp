// 2D ingestion for the re-hosted offset2d: SVG -> closed curves -> DoubleCompressedImage.
//
//  * DoubleCompressedImage::fromImage / scanLine / unionIntersections: host restatement of
//    src/vor2d/DoubleCompressedImage.cpp:25-111 (pinned against the compiled reference through
//    tests/test_from_image.py: offset2d <svg> with radius 0 vs oracle/_ref).
//  * voroffset::create_dexels(file): src/vor2d/Dexelize.cpp:22-47. The reference parses with nanosvg
//    (nsvgParseFromFile(file, "mm", 90)), which is not available here; read_svg below is a small reader that
//    follows nanosvg's conventions for the subset it understands: float coordinates, user units at 90 DPI
//    converted to millimetres, viewBox with preserveAspectRatio (default xMidYMid meet), nested transforms,
//    every segment stored as a cubic (a straight line gets control points at 1/3 and 2/3), sub-paths of a shape
//    in reverse document order, and - like Dexelize.cpp:33-37 - ALL stored points but the last of each sub-path
//    become polygon vertices (control points included). Elements: path (M L H V C S Q T Z, both cases), polygon,
//    polyline, rect (square corners), circle, ellipse, line; g with transform; defs skipped. Elliptical arcs
//    (A / a), rounded rectangles, CSS, <use> are not supported and raise. Parity with nanosvg itself is UNPINNED.
#include "vo_host.hpp"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>

namespace voroffset
{
	// ---- DoubleCompressedImage.cpp:25-40 ------------------------------------------------------------
	void DoubleCompressedImage::fromImage(const std::vector<Curve> &input_curves)
	{
		std::vector<Scalar> intersections;
		for (int j = 0; j < (int)m_Rays.size(); ++j) {
			m_Rays[j].clear();
			for (const Curve &curve : input_curves) {
				scanLine(intersections, j, curve);
				unionIntersections(m_Rays[j], intersections);
				intersections.clear();
			}
		}
	}

	// DoubleCompressedImage.cpp:43-86: crossings of the closed polygon with the line x = line_x (an integer abscissa).
	// From every vertex off the line, walk to the next vertex off the line; the polygon crosses when the two lie on
	// different sides. A direct edge is cut with the reference's interpolation (weights from the END point:
	// s = (x_next - line) / (x_next - x_i), y = s y_i + (1 - s) y_next); when vertices ON the line lie in between,
	// the crossing is the y of the last of them.
	void DoubleCompressedImage::scanLine(std::vector<Scalar> &out, int line_x, const Curve &curve)
	{
		out.clear();
		const int n = (int)curve.size();
		auto side = [&](int i) { const Scalar x = curve[i % n].real(); return x < line_x ? -1 : (x > line_x ? 1 : 0); };
		for (int i = 0; i < n; ++i) {
			const int here = side(i);
			if (here == 0) continue;
			int k = i + 1;
			while (side(k) == 0) ++k;              // (terminates: vertex i itself is off the line)
			if (side(k) == here) continue;
			if (k == i + 1) {
				const PointF &p = curve[i], &q = curve[k % n];
				const Scalar s = (q.real() - line_x) / (q.real() - p.real());
				out.push_back(s * p.imag() + (1 - s) * q.imag());
			} else {
				out.push_back(curve[(k - 1) % n].imag());
			}
		}
		std::sort(out.begin(), out.end());
	}

	// DoubleCompressedImage.cpp:88-111: curves are assumed disjoint; a curve's crossings go behind the ray's
	// content when they start at or after its last value and in front of it otherwise.
	void DoubleCompressedImage::unionIntersections(std::vector<Scalar> &a, const std::vector<Scalar> &b)
	{
		if (b.empty()) return;
		if (a.empty() || a.back() <= b.front()) a.insert(a.end(), b.begin(), b.end());
		else a.insert(a.begin(), b.begin(), b.end());
	}

	// Same result as DoubleCompressedImage::transposeInPlace (DoubleCompressedImage.cpp:478-584, the #else branch that
	// is compiled), computed on a coverage vector instead of the reference's incremental run toggling. What that
	// routine does: every event is truncated to int (it comes from the integer-pixel CompressedImage) and the (line,
	// ray) pairs are de-duplicated in a std::set; going through the lines in ascending order, the rays that have an
	// event on a line flip their coverage there; new ray ii, for every ii from one event line up to the next, lists
	// the positions where the coverage (as a function of the old ray index) changes; the last event line only
	// yields its own ray. An image without events just swaps its dimensions (the reference indexes m_Rays[-1]
	// there); an event at or beyond the old ray length raises instead of writing out of bounds.
	void DoubleCompressedImage::transposeInPlace()
	{
		const int old_rays = (int)m_Rays.size();
		std::map<int, std::set<int>> flips;                    // line -> rays with an event on it
		for (int j = 0; j < old_rays; ++j)
			for (const Scalar val : m_Rays[j]) flips[(int)val].insert(j);
		const int new_rays = m_XSize;
		m_XSize = old_rays;
		m_Rays.assign(new_rays, {});
		if (flips.empty()) return;
		std::vector<char> covered((size_t)old_rays + 1, 0);    // [old_rays] stays 0: closes a run that reaches the end
		auto boundaries = [&](std::vector<Scalar> &dst) {
			char prev = 0;
			for (int j = 0; j <= old_rays; ++j) {
				if (covered[j] != prev) dst.push_back((Scalar)j);
				prev = covered[j];
			}
		};
		int prev_line = -1;
		for (const auto &line : flips) {
			if (prev_line != -1)
				for (int ii = prev_line; ii < line.first; ++ii) boundaries(m_Rays.at(ii));
			for (int j : line.second) covered[j] ^= 1;
			prev_line = line.first;
		}
		boundaries(m_Rays.at(prev_line));
	}

	// ---- a small SVG reader with nanosvg's conventions ----------------------------------------------
	namespace
	{
		const float DPI = 90.0f;        // Dexelize.cpp:17-19
		const float K90 = 0.5522847493f;

		struct Xf { float t[6] = {1, 0, 0, 1, 0, 0}; };
		Xf mul(const Xf &a, const Xf &b)   // apply a first, then b (nanosvg nsvg__xformMultiply(a, b))
		{
			Xf r;
			r.t[0] = a.t[0] * b.t[0] + a.t[1] * b.t[2];
			r.t[2] = a.t[2] * b.t[0] + a.t[3] * b.t[2];
			r.t[4] = a.t[4] * b.t[0] + a.t[5] * b.t[2] + b.t[4];
			r.t[1] = a.t[0] * b.t[1] + a.t[1] * b.t[3];
			r.t[3] = a.t[2] * b.t[1] + a.t[3] * b.t[3];
			r.t[5] = a.t[4] * b.t[1] + a.t[5] * b.t[3] + b.t[5];
			return r;
		}

		struct SubPath { std::vector<float> pts; };
		struct Shape { std::vector<SubPath> paths; };   // in nanosvg's (reversed) order

		struct Tag { std::string name; std::vector<std::pair<std::string, std::string>> attr; bool open = false, close = false; };

		float to_px(const std::string &v, float ref_len)
		{
			const char *s = v.c_str();
			char *end = nullptr;
			const float x = (float)std::strtod(s, &end);
			std::string u(end);
			u.erase(std::remove_if(u.begin(), u.end(), [](unsigned char c) { return std::isspace(c); }), u.end());
			if (u.empty() || u == "px") return x;
			if (u == "pt") return x / 72.0f * DPI;
			if (u == "pc") return x / 6.0f * DPI;
			if (u == "mm") return x / 25.4f * DPI;
			if (u == "cm") return x / 2.54f * DPI;
			if (u == "in") return x * DPI;
			if (u == "%") return x / 100.0f * ref_len;
			throw std::runtime_error("svg: unsupported unit '" + u + "'");
		}

		std::vector<float> numbers(const std::string &s)
		{
			std::vector<float> out;
			const char *p = s.c_str();
			while (*p) {
				if (std::isdigit((unsigned char)*p) || *p == '-' || *p == '+' || *p == '.') {
					char *end = nullptr;
					out.push_back((float)std::strtod(p, &end));
					if (end == p) break;
					p = end;
				} else ++p;
			}
			return out;
		}

		Xf parse_transform(const std::string &s)
		{
			Xf acc;
			size_t i = 0;
			while (i < s.size()) {
				while (i < s.size() && !std::isalpha((unsigned char)s[i])) ++i;
				size_t j = i;
				while (j < s.size() && std::isalpha((unsigned char)s[j])) ++j;
				if (j == i) break;
				const std::string name = s.substr(i, j - i);
				const size_t a = s.find('(', j), b = s.find(')', j);
				if (a == std::string::npos || b == std::string::npos) break;
				const std::vector<float> v = numbers(s.substr(a + 1, b - a - 1));
				Xf t;
				if (name == "matrix" && v.size() == 6) for (int k = 0; k < 6; ++k) t.t[k] = v[k];
				else if (name == "translate" && !v.empty()) { t.t[4] = v[0]; t.t[5] = v.size() > 1 ? v[1] : 0.0f; }
				else if (name == "scale" && !v.empty()) { t.t[0] = v[0]; t.t[3] = v.size() > 1 ? v[1] : v[0]; }
				else if (name == "rotate" && !v.empty()) {
					const float an = v[0] / 180.0f * 3.14159265358979323846f, cs = std::cos(an), sn = std::sin(an);
					Xf r; r.t[0] = cs; r.t[1] = sn; r.t[2] = -sn; r.t[3] = cs;
					if (v.size() > 2) { Xf m1, m2; m1.t[4] = -v[1]; m1.t[5] = -v[2]; m2.t[4] = v[1]; m2.t[5] = v[2]; t = mul(mul(m1, r), m2); }
					else t = r;
				} else if (name == "skewX" && !v.empty()) t.t[2] = std::tan(v[0] / 180.0f * 3.14159265358979323846f);
				else if (name == "skewY" && !v.empty()) t.t[1] = std::tan(v[0] / 180.0f * 3.14159265358979323846f);
				else throw std::runtime_error("svg: unsupported transform '" + name + "'");
				acc = mul(t, acc);     // nsvg__xformPremultiply: the later entry applies first
				i = b + 1;
			}
			return acc;
		}

		// the point list nanosvg keeps while a sub-path is built
		struct Builder {
			std::vector<float> pts;
			std::vector<SubPath> done;
			Xf xf;
			void reset() { pts.clear(); }
			void move_to(float x, float y) { if (!pts.empty()) { pts[pts.size() - 2] = x; pts[pts.size() - 1] = y; } else { pts.push_back(x); pts.push_back(y); } }
			void line_to(float x, float y)
			{
				if (pts.empty()) return;
				const float px = pts[pts.size() - 2], py = pts[pts.size() - 1], dx = x - px, dy = y - py;
				const float c[6] = {px + dx / 3.0f, py + dy / 3.0f, x - dx / 3.0f, y - dy / 3.0f, x, y};
				pts.insert(pts.end(), c, c + 6);
			}
			void cubic_to(float a, float b, float c, float d, float x, float y) { if (pts.empty()) return; const float q[6] = {a, b, c, d, x, y}; pts.insert(pts.end(), q, q + 6); }
			void add_path(bool closed)
			{
				if (pts.size() < 8) return;
				if (closed) line_to(pts[0], pts[1]);
				if ((pts.size() / 2) % 3 != 1) return;
				SubPath sp;
				for (size_t i = 0; i + 1 < pts.size(); i += 2) {
					sp.pts.push_back(pts[i] * xf.t[0] + pts[i + 1] * xf.t[2] + xf.t[4]);
					sp.pts.push_back(pts[i] * xf.t[1] + pts[i + 1] * xf.t[3] + xf.t[5]);
				}
				done.insert(done.begin(), sp);     // nanosvg prepends to the shape's path list
			}
		};

		void parse_path_d(const std::string &d, Builder &b)
		{
			const char *p = d.c_str();
			char cmd = 0;
			float cpx = 0, cpy = 0, cpx2 = 0, cpy2 = 0;
			bool closed = false;
			std::vector<float> args;
			auto nargs = [](char c) { switch (std::tolower(c)) { case 'm': case 'l': case 't': return 2; case 'h': case 'v': return 1; case 'c': return 6; case 's': case 'q': return 4; case 'a': return 7; default: return 0; } };
			b.reset();
			while (*p) {
				while (*p && (std::isspace((unsigned char)*p) || *p == ',')) ++p;
				if (!*p) break;
				if (std::isalpha((unsigned char)*p) && *p != 'e' && *p != 'E') {
					cmd = *p++;
					args.clear();
					if (cmd == 'A' || cmd == 'a') throw std::runtime_error("svg: elliptical arcs (A / a) are not supported");
					if (cmd == 'Z' || cmd == 'z') {
						closed = true;
						if (!b.pts.empty()) { cpx = b.pts[0]; cpy = b.pts[1]; cpx2 = cpx; cpy2 = cpy; b.add_path(closed); }
						b.reset();
						b.move_to(cpx, cpy);
						closed = false;
					} else if (nargs(cmd) == 0) throw std::runtime_error(std::string("svg: unknown path command '") + cmd + "'");
					continue;
				}
				char *end = nullptr;
				const float v = (float)std::strtod(p, &end);
				if (end == p) throw std::runtime_error("svg: malformed path data");
				p = end;
				if (!cmd) throw std::runtime_error("svg: path data must start with a command");
				args.push_back(v);
				if ((int)args.size() < nargs(cmd)) continue;
				const bool rel = std::islower((unsigned char)cmd) != 0;
				switch (std::tolower(cmd)) {
				case 'm':
					if (rel) { cpx += args[0]; cpy += args[1]; } else { cpx = args[0]; cpy = args[1]; }
					if (!b.pts.empty()) b.add_path(closed);                        // commit the previous sub-path
					b.reset();
					b.move_to(cpx, cpy);
					closed = false;
					cpx2 = cpx; cpy2 = cpy;
					cmd = rel ? 'l' : 'L';                                          // further pairs are lineTo
					break;
				case 'l':
					if (rel) { cpx += args[0]; cpy += args[1]; } else { cpx = args[0]; cpy = args[1]; }
					b.line_to(cpx, cpy); cpx2 = cpx; cpy2 = cpy; break;
				case 'h':
					cpx = rel ? cpx + args[0] : args[0];
					b.line_to(cpx, cpy); cpx2 = cpx; cpy2 = cpy; break;
				case 'v':
					cpy = rel ? cpy + args[0] : args[0];
					b.line_to(cpx, cpy); cpx2 = cpx; cpy2 = cpy; break;
				case 'c': {
					const float ox = rel ? cpx : 0, oy = rel ? cpy : 0;
					b.cubic_to(ox + args[0], oy + args[1], ox + args[2], oy + args[3], ox + args[4], oy + args[5]);
					cpx2 = ox + args[2]; cpy2 = oy + args[3]; cpx = ox + args[4]; cpy = oy + args[5]; break;
				}
				case 's': {
					const float ox = rel ? cpx : 0, oy = rel ? cpy : 0;
					const float c1x = 2 * cpx - cpx2, c1y = 2 * cpy - cpy2;
					b.cubic_to(c1x, c1y, ox + args[0], oy + args[1], ox + args[2], oy + args[3]);
					cpx2 = ox + args[0]; cpy2 = oy + args[1]; cpx = ox + args[2]; cpy = oy + args[3]; break;
				}
				case 'q': {
					const float ox = rel ? cpx : 0, oy = rel ? cpy : 0;
					const float cx = ox + args[0], cy = oy + args[1], x2 = ox + args[2], y2 = oy + args[3];
					b.cubic_to(cpx + 2.0f / 3.0f * (cx - cpx), cpy + 2.0f / 3.0f * (cy - cpy), x2 + 2.0f / 3.0f * (cx - x2), y2 + 2.0f / 3.0f * (cy - y2), x2, y2);
					cpx2 = cx; cpy2 = cy; cpx = x2; cpy = y2; break;
				}
				case 't': {
					const float x2 = rel ? cpx + args[0] : args[0], y2 = rel ? cpy + args[1] : args[1];
					const float cx = 2 * cpx - cpx2, cy = 2 * cpy - cpy2;
					b.cubic_to(cpx + 2.0f / 3.0f * (cx - cpx), cpy + 2.0f / 3.0f * (cy - cpy), x2 + 2.0f / 3.0f * (cx - x2), y2 + 2.0f / 3.0f * (cy - y2), x2, y2);
					cpx2 = cx; cpy2 = cy; cpx = x2; cpy = y2; break;
				}
				}
				args.clear();
			}
			if (!b.pts.empty()) b.add_path(closed);
		}

		// minimal tag scanner: comments, processing instructions and doctype are skipped, entities are not expanded
		std::vector<Tag> scan_tags(const std::string &s)
		{
			std::vector<Tag> tags;
			size_t i = 0;
			while ((i = s.find('<', i)) != std::string::npos) {
				if (s.compare(i, 4, "<!--") == 0) { i = s.find("-->", i); if (i == std::string::npos) break; i += 3; continue; }
				if (s.compare(i, 2, "<?") == 0 || s.compare(i, 2, "<!") == 0) { i = s.find('>', i); if (i == std::string::npos) break; ++i; continue; }
				size_t j = i + 1;
				Tag t;
				if (j < s.size() && s[j] == '/') { t.close = true; ++j; }
				size_t k = j;
				while (k < s.size() && !std::isspace((unsigned char)s[k]) && s[k] != '>' && s[k] != '/') ++k;
				t.name = s.substr(j, k - j);
				const size_t colon = t.name.find(':');
				if (colon != std::string::npos) t.name = t.name.substr(colon + 1);
				// attributes
				while (k < s.size() && s[k] != '>') {
					while (k < s.size() && (std::isspace((unsigned char)s[k]) || s[k] == '/')) { ++k; }
					if (k >= s.size() || s[k] == '>') break;
					size_t e = k;
					while (e < s.size() && s[e] != '=' && s[e] != '>' && !std::isspace((unsigned char)s[e])) ++e;
					std::string key = s.substr(k, e - k);
					while (e < s.size() && std::isspace((unsigned char)s[e])) ++e;
					std::string val;
					if (e < s.size() && s[e] == '=') {
						++e;
						while (e < s.size() && std::isspace((unsigned char)s[e])) ++e;
						if (e < s.size() && (s[e] == '"' || s[e] == '\'')) {
							const char q = s[e];
							const size_t f = s.find(q, e + 1);
							if (f == std::string::npos) throw std::runtime_error("svg: unterminated attribute value");
							val = s.substr(e + 1, f - e - 1);
							e = f + 1;
						}
					}
					t.attr.push_back({key, val});
					k = e;
				}
				if (k >= s.size()) break;
				const bool self = k > 0 && s[k - 1] == '/';
				if (!t.close) { t.open = true; t.close = self; }
				tags.push_back(t);
				i = k + 1;
			}
			return tags;
		}

		const std::string *get(const Tag &t, const char *key)
		{
			for (const auto &a : t.attr) if (a.first == key) return &a.second;
			return nullptr;
		}

		struct Svg { float width = 0, height = 0; std::vector<Shape> shapes; };

		Svg read_svg(const std::string &file)
		{
			std::ifstream in(file, std::ios::binary);
			if (!in) throw std::runtime_error("Invalid input file: " + file);
			std::stringstream ss;
			ss << in.rdbuf();
			const std::vector<Tag> tags = scan_tags(ss.str());
			Svg svg;
			float vminx = 0, vminy = 0, vw = 0, vh = 0;
			int align_x = 1, align_y = 1, align_type = 1;   // 0 min, 1 mid, 2 max; type 0 none, 1 meet, 2 slice
			std::vector<Xf> xstack{Xf()};
			std::vector<bool> hidden{false};
			int defs = 0;
			bool seen_svg = false;
			for (const Tag &t : tags) {
				if (t.open) {
					Xf xf = xstack.back();
					bool hide = hidden.back();
					if (const std::string *tr = get(t, "transform")) xf = mul(parse_transform(*tr), xf);
					if (const std::string *st = get(t, "style")) { std::string c = *st; c.erase(std::remove_if(c.begin(), c.end(), [](unsigned char ch) { return std::isspace(ch); }), c.end()); if (c.find("display:none") != std::string::npos) hide = true; }
					if (const std::string *dp = get(t, "display")) if (*dp == "none") hide = true;
					if (t.name == "svg" && !seen_svg) {
						seen_svg = true;
						if (const std::string *w = get(t, "width")) svg.width = to_px(*w, 0.0f);
						if (const std::string *h = get(t, "height")) svg.height = to_px(*h, 0.0f);
						if (const std::string *vb = get(t, "viewBox")) { const auto v = numbers(*vb); if (v.size() == 4) { vminx = v[0]; vminy = v[1]; vw = v[2]; vh = v[3]; } }
						if (const std::string *pa = get(t, "preserveAspectRatio")) {
							if (pa->find("none") != std::string::npos) align_type = 0;
							else {
								if (pa->find("xMin") != std::string::npos) align_x = 0; else if (pa->find("xMax") != std::string::npos) align_x = 2;
								if (pa->find("yMin") != std::string::npos) align_y = 0; else if (pa->find("yMax") != std::string::npos) align_y = 2;
								if (pa->find("slice") != std::string::npos) align_type = 2;
							}
						}
					} else if (t.name == "defs") {
						if (!t.close) defs++;
					} else if (defs == 0 && !hide) {
						Builder b;
						b.xf = xf;
						auto num = [&](const char *k, float ref) { const std::string *v = get(t, k); return v ? to_px(*v, ref) : 0.0f; };
						if (t.name == "path") {
							if (const std::string *d = get(t, "d")) parse_path_d(*d, b);
						} else if (t.name == "polygon" || t.name == "polyline") {
							if (const std::string *pp = get(t, "points")) {
								const auto v = numbers(*pp);
								for (size_t i = 0; i + 1 < v.size(); i += 2) { if (i == 0) b.move_to(v[0], v[1]); else b.line_to(v[i], v[i + 1]); }
								b.add_path(t.name == "polygon");
							}
						} else if (t.name == "rect") {
							const float x = num("x", svg.width), y = num("y", svg.height), w = num("width", svg.width), h = num("height", svg.height);
							if (get(t, "rx") || get(t, "ry")) throw std::runtime_error("svg: rounded rectangles are not supported");
							if (w != 0.0f && h != 0.0f) { b.move_to(x, y); b.line_to(x + w, y); b.line_to(x + w, y + h); b.line_to(x, y + h); b.add_path(true); }
						} else if (t.name == "circle" || t.name == "ellipse") {
							const float cx = num("cx", svg.width), cy = num("cy", svg.height);
							const float rx = t.name == "circle" ? num("r", svg.width) : num("rx", svg.width), ry = t.name == "circle" ? rx : num("ry", svg.height);
							if (rx > 0.0f && ry > 0.0f) {
								b.move_to(cx + rx, cy);
								b.cubic_to(cx + rx, cy + ry * K90, cx + rx * K90, cy + ry, cx, cy + ry);
								b.cubic_to(cx - rx * K90, cy + ry, cx - rx, cy + ry * K90, cx - rx, cy);
								b.cubic_to(cx - rx, cy - ry * K90, cx - rx * K90, cy - ry, cx, cy - ry);
								b.cubic_to(cx + rx * K90, cy - ry, cx + rx, cy - ry * K90, cx + rx, cy);
								b.add_path(true);
							}
						} else if (t.name == "line") {
							b.move_to(num("x1", svg.width), num("y1", svg.height)); b.line_to(num("x2", svg.width), num("y2", svg.height)); b.add_path(false);
						} else if (t.name == "use") {
							throw std::runtime_error("svg: <use> is not supported");
						}
						if (!b.done.empty()) { Shape s; s.paths = b.done; svg.shapes.push_back(s); }
					}
					if (!t.close) { xstack.push_back(xf); hidden.push_back(hide); }
				} else if (t.close) {
					if (t.name == "defs" && defs > 0) defs--;
					if (xstack.size() > 1) { xstack.pop_back(); hidden.pop_back(); }
				}
			}
			if (!seen_svg) throw std::runtime_error("Invalid input file (no <svg> element): " + file);
			// nsvg__scaleToViewbox with units "mm"
			float bx0 = 1e30f, by0 = 1e30f, bx1 = -1e30f, by1 = -1e30f;
			for (const auto &s : svg.shapes) for (const auto &p : s.paths) for (size_t i = 0; i + 1 < p.pts.size(); i += 2) {
				bx0 = std::min(bx0, p.pts[i]); bx1 = std::max(bx1, p.pts[i]); by0 = std::min(by0, p.pts[i + 1]); by1 = std::max(by1, p.pts[i + 1]);
			}
			if (svg.shapes.empty()) bx0 = by0 = bx1 = by1 = 0;
			if (vw == 0) { if (svg.width > 0) vw = svg.width; else { vminx = bx0; vw = bx1 - bx0; } }
			if (vh == 0) { if (svg.height > 0) vh = svg.height; else { vminy = by0; vh = by1 - by0; } }
			if (svg.width == 0) svg.width = vw;
			if (svg.height == 0) svg.height = vh;
			float tx = -vminx, ty = -vminy;
			float sx = vw > 0 ? svg.width / vw : 0, sy = vh > 0 ? svg.height / vh : 0;
			const float us = 1.0f / (1.0f / 25.4f * DPI);
			auto view_align = [](float content, float container, int type) { return type == 0 ? 0.0f : type == 2 ? container - content : (container - content) * 0.5f; };
			if (align_type == 1) { sx = sy = std::min(sx, sy); tx += view_align(vw * sx, svg.width, align_x) / sx; ty += view_align(vh * sy, svg.height, align_y) / sy; }
			else if (align_type == 2) { sx = sy = std::max(sx, sy); tx += view_align(vw * sx, svg.width, align_x) / sx; ty += view_align(vh * sy, svg.height, align_y) / sy; }
			sx *= us; sy *= us;
			for (auto &s : svg.shapes) for (auto &p : s.paths) for (size_t i = 0; i + 1 < p.pts.size(); i += 2) { p.pts[i] = (p.pts[i] + tx) * sx; p.pts[i + 1] = (p.pts[i + 1] + ty) * sy; }
			svg.width *= us; svg.height *= us;
			return svg;
		}
	}

	// src/vor2d/Dexelize.cpp:22-47
	std::vector<Curve> svg_contours(const std::string &file, double &width_mm, double &height_mm)
	{
		const Svg svg = read_svg(file);
		std::vector<Curve> contours;
		for (const Shape &shape : svg.shapes) {
			Curve poly;
			for (const SubPath &path : shape.paths) {
				const int npts = (int)path.pts.size() / 2;
				for (int i = 0; i < npts - 1; i++) poly.push_back(PointF(path.pts[2 * i], path.pts[2 * i + 1]));   // Dexelize.cpp:35-36
			}
			contours.push_back(poly);
		}
		width_mm = svg.width; height_mm = svg.height;
		return contours;
	}

	DoubleCompressedImage create_dexels(const std::string &file)
	{
		double w = 0, h = 0;
		std::cerr << "[loading] " << file << " ... ";
		const std::vector<Curve> contours = svg_contours(file, w, h);
		std::cerr << "Read a SVG of size : " << w << " x " << h << " (" << contours.size() << " polygones)" << std::endl;
		// Dexelize.cpp:42, unit quirk included: the size already is in mm and is scaled by 25.4 / DPI once more;
		// the first constructor argument is the ray length (m_XSize), the second the number of rays
		DoubleCompressedImage dexels((int)std::ceil(h * 25.4 / 90), (int)std::ceil(w * 25.4 / 90));
		dexels.fromImage(contours);
		return dexels;
	}

	// src/vor2d/Dexelize.cpp:48-89 without geogram: one quad per interval, ray x spans [x, x+1], z = 1; vertices
	// in the reference's order (min,min) (max,min) (min,max) (max,max) and the quad (v, v+1, v+2, v+3) as it is
	// created there. compute_borders / connect / remove_isolated do not change what mesh_save writes for an OBJ.
	void dexel_dump(const std::string &filename, const DoubleCompressedImage &dexels)
	{
		std::ofstream out(filename);
		if (!out) throw std::runtime_error("cannot write " + filename);
		out << std::setprecision(17);
		size_t v = 1;
		std::ostringstream faces;
		for (int x = 0; x < dexels.height(); ++x)
			for (size_t i = 0; 2 * i + 1 < dexels.m_Rays[x].size(); ++i) {
				const double y0 = dexels.m_Rays[x][2 * i], y1 = dexels.m_Rays[x][2 * i + 1];
				out << "v " << x << ' ' << y0 << " 1\nv " << x + 1 << ' ' << y0 << " 1\nv " << x << ' ' << y1 << " 1\nv " << x + 1 << ' ' << y1 << " 1\n";
				faces << "f " << v << ' ' << v + 1 << ' ' << v + 2 << ' ' << v + 3 << '\n';
				v += 4;
			}
		out << faces.str();
	}
}
