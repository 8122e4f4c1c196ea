#include "vo_host.hpp"

#include <algorithm>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <limits>
#include <map>
#include <sstream>

namespace
{
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
	bool endswith(const std::string &s, const std::string &suf)
	{
		return s.size() >= suf.size() && 0 == s.compare(s.size() - suf.size(), suf.size(), suf);
	}
	std::string lower(std::string s) { for (auto &c : s) c = (char)std::tolower((unsigned char)c); return s; }
}

namespace voroffset3d
{
	// ---- CompressedVolume --------------------------------------------------------------------------
	void CompressedVolume::save(std::ostream &out) const
	{
		out << std::setprecision(17);
		out << m_Origin[0] << ' ' << m_Origin[1] << ' ' << m_Origin[2] << "\n";
		out << m_Extent[0] << ' ' << m_Extent[1] << ' ' << m_Extent[2] << "\n";
		out << m_GridSize[0] << ' ' << m_GridSize[1] << "\n" << m_Padding << "\n" << m_Spacing << "\n";
		for (const auto &row : m_Data) {
			out << row.size();
			for (const auto &v : row) out << ' ' << v;
			out << "\n";
		}
	}

	void CompressedVolume::load(std::istream &in)
	{
		in >> m_Origin[0] >> m_Origin[1] >> m_Origin[2];
		in >> m_Extent[0] >> m_Extent[1] >> m_Extent[2];
		in >> m_GridSize[0] >> m_GridSize[1] >> m_Padding >> m_Spacing;
		m_Data.assign((size_t)m_GridSize[0] * m_GridSize[1], {});
		for (auto &row : m_Data) {
			size_t n;
			in >> n;
			row.resize(n);
			for (auto &v : row) in >> v;
		}
		if (!in) throw std::runtime_error("Invalid volume file.");
	}

	void CompressedVolume::to_csr(std::vector<uint32_t> &off, std::vector<double> &spans) const
	{
		off.assign(m_Data.size() + 1, 0);
		size_t total = 0;
		for (size_t c = 0; c < m_Data.size(); ++c) { total += m_Data[c].size() / 2; off[c + 1] = (uint32_t)total; }
		if (total >= (1ull << 32)) throw std::runtime_error("volume has more than 2^32-1 intervals");
		spans.resize(2 * total + 2);
		size_t w = 0;
		for (const auto &l : m_Data) for (size_t k = 0; k + 1 < l.size(); k += 2) { spans[w++] = l[k]; spans[w++] = l[k + 1]; }
	}

	void CompressedVolume::from_csr(const uint32_t *off, const double *spans)
	{
		for (size_t c = 0; c < m_Data.size(); ++c) m_Data[c].assign(spans + 2 * (size_t)off[c], spans + 2 * (size_t)off[c + 1]);
	}

	// ---- VoronoiMorpho -----------------------------------------------------------------------------
	VoronoiMorpho::VoronoiMorpho(int method, int device) : m_ctx(make_ctx(device)), m_method(method) {}
	VoronoiMorpho::VoronoiMorpho(int method, int first_device, int n_gpus) : m_ctx(nullptr), m_method(method)
	{
		if (n_gpus <= 1) { m_ctx = make_ctx(first_device); return; }
		std::vector<int> devs(n_gpus);
		for (int i = 0; i < n_gpus; ++i) devs[i] = first_device + i;
		if (vo_mg_create(devs.data(), n_gpus, &m_mg) != VO_OK)
			throw std::runtime_error("voroffset_b200: cannot create a " + std::to_string(n_gpus) + "-GPU group (CUDA devices or libnccl.so.2 missing; there is no CPU fallback)");
		m_ctx = vo_mg_ctx(m_mg, 0);      // (owned by the group; used for xor)
	}
	VoronoiMorpho::~VoronoiMorpho() { if (m_mg) vo_mg_destroy(m_mg); else vo_destroy(m_ctx); }

	void VoronoiMorpho::run(int op, const CompressedVolume &input, CompressedVolume &result, double radius, double &t1, double &t2)
	{
		const int nx = input.gridSize()[0], ny = input.gridSize()[1];
		std::vector<uint32_t> off;
		std::vector<double> spans;
		input.to_csr(off, spans);
		uint32_t *o_off = nullptr;
		double *o_spans = nullptr;
		uint64_t n = 0;
		if (m_mg) {
			const int rc = vo_mg_morph3d(m_mg, op, m_method, nx, ny, input.zmin(), input.zmax(), off.data(), spans.data(), radius,
			                             &o_off, &o_spans, &n, &t1, &t2);
			if (rc != VO_OK) throw std::runtime_error(std::string("Assertion failed: ") + vo_mg_last_error(m_mg));
		} else
			check(m_ctx, vo_morph3d(m_ctx, op, m_method, nx, ny, input.zmin(), input.zmax(), off.data(), spans.data(), radius,
			                        &o_off, &o_spans, &n, &t1, &t2));
		result.reset(input.origin(), input.extent(), input.spacing(), input.padding(), nx, ny);   // VoronoiVorPower.cpp:37
		result.from_csr(o_off, o_spans);
		vo_free(o_off);
		vo_free(o_spans);
	}
	void VoronoiMorpho::dilation(CompressedVolume input, CompressedVolume &result, double radius, double &t1, double &t2) { run(VO_OP_DILATION, input, result, radius, t1, t2); }
	void VoronoiMorpho::erosion(CompressedVolume input, CompressedVolume &result, double radius, double &t1, double &t2) { run(VO_OP_EROSION, input, result, radius, t1, t2); }
	void VoronoiMorpho::opening(CompressedVolume input, CompressedVolume &result, double radius, double &t1, double &t2) { run(VO_OP_OPENING, input, result, radius, t1, t2); }
	void VoronoiMorpho::closing(CompressedVolume input, CompressedVolume &result, double radius, double &t1, double &t2) { run(VO_OP_CLOSING, input, result, radius, t1, t2); }

	double VoronoiMorpho::calculateXor(CompressedVolume a, CompressedVolume b, CompressedVolume &result)
	{
		if (a.gridSize() != b.gridSize()) throw std::runtime_error("calculateXor: the two voxels must have the same grid size");
		std::vector<uint32_t> oa, ob;
		std::vector<double> sa, sb;
		a.to_csr(oa, sa);
		b.to_csr(ob, sb);
		uint32_t *o_off = nullptr;
		double *o_spans = nullptr;
		uint64_t n = 0;
		double volume = 0;
		check(m_ctx, vo_xor3d(m_ctx, a.gridSize()[0], a.gridSize()[1], a.zmin(), a.zmax(), a.spacing(), oa.data(), sa.data(),
		                      ob.data(), sb.data(), &o_off, &o_spans, &n, &volume));
		result.reset(a.origin(), a.extent(), a.spacing(), a.padding(), a.gridSize()[0], a.gridSize()[1]);
		result.from_csr(o_off, o_spans);
		vo_free(o_off);
		vo_free(o_spans);
		return volume;
	}

	// ---- mesh I/O + dexelisation (restates src/vor3d/Dexelize.cpp without geogram) ------------------
	namespace
	{
		struct Mesh { std::vector<std::array<double, 3>> V; std::vector<std::array<int, 3>> F; };

		void add_polygon(Mesh &m, const std::vector<int> &poly)
		{
			for (size_t k = 1; k + 1 < poly.size(); ++k) m.F.push_back({{poly[0], poly[k], poly[k + 1]}});   // fan
		}

		Mesh load_obj(std::istream &in)
		{
			Mesh m;
			std::string line;
			while (std::getline(in, line)) {
				std::istringstream ss(line);
				std::string tag;
				ss >> tag;
				if (tag == "v") { std::array<double, 3> p; ss >> p[0] >> p[1] >> p[2]; m.V.push_back(p); }
				else if (tag == "f") {
					std::vector<int> poly;
					std::string tok;
					while (ss >> tok) {
						int i = std::stoi(tok.substr(0, tok.find('/')));
						poly.push_back(i > 0 ? i - 1 : (int)m.V.size() + i);
					}
					add_polygon(m, poly);
				}
			}
			return m;
		}

		Mesh load_off(std::istream &in)
		{
			Mesh m;
			std::string hdr;
			in >> hdr;
			size_t nv, nf, ne;
			in >> nv >> nf >> ne;
			m.V.resize(nv);
			for (auto &p : m.V) in >> p[0] >> p[1] >> p[2];
			for (size_t f = 0; f < nf; ++f) {
				int k;
				in >> k;
				std::vector<int> poly(k);
				for (auto &i : poly) in >> i;
				add_polygon(m, poly);
			}
			return m;
		}

		Mesh load_stl(const std::string &filename)
		{
			Mesh m;
			std::ifstream in(filename, std::ios::binary);
			char head[80];
			uint32_t ntri = 0;
			in.read(head, 80);
			in.read(reinterpret_cast<char *>(&ntri), 4);
			in.seekg(0, std::ios::end);
			const std::streamoff size = in.tellg();
			if (size == (std::streamoff)(84 + 50ull * ntri)) {          // binary
				in.seekg(84);
				for (uint32_t t = 0; t < ntri; ++t) {
					float buf[12];
					uint16_t attr;
					in.read(reinterpret_cast<char *>(buf), 48);
					in.read(reinterpret_cast<char *>(&attr), 2);
					const int b = (int)m.V.size();
					for (int k = 0; k < 3; ++k) m.V.push_back({{buf[3 + 3 * k], buf[4 + 3 * k], buf[5 + 3 * k]}});
					m.F.push_back({{b, b + 1, b + 2}});
				}
			} else {                                                   // ascii
				std::ifstream txt(filename);
				std::string tok;
				while (txt >> tok)
					if (tok == "vertex") {
						std::array<double, 3> p;
						txt >> p[0] >> p[1] >> p[2];
						m.V.push_back(p);
						if (m.V.size() % 3 == 0) { const int b = (int)m.V.size() - 3; m.F.push_back({{b, b + 1, b + 2}}); }
					}
			}
			return m;
		}

		Mesh load_ply_ascii(std::istream &in)
		{
			Mesh m;
			std::string line, tok;
			size_t nv = 0, nf = 0;
			while (std::getline(in, line)) {
				std::istringstream ss(line);
				ss >> tok;
				if (tok == "format") { ss >> tok; if (tok != "ascii") throw std::runtime_error("Invalid input mesh (only ascii PLY is supported)."); }
				else if (tok == "element") { std::string what; size_t n; ss >> what >> n; if (what == "vertex") nv = n; else if (what == "face") nf = n; }
				else if (tok == "end_header") break;
			}
			m.V.resize(nv);
			for (auto &p : m.V) { std::getline(in, line); std::istringstream ss(line); ss >> p[0] >> p[1] >> p[2]; }
			for (size_t f = 0; f < nf; ++f) {
				int k;
				in >> k;
				std::vector<int> poly(k);
				for (auto &i : poly) in >> i;
				add_polygon(m, poly);
			}
			return m;
		}

		Mesh mesh_load(const std::string &filename)
		{
			const std::string f = lower(filename);
			std::ifstream in(filename);
			if (!in) throw std::runtime_error("Invalid input mesh.");
			Mesh m;
			if (endswith(f, ".obj")) m = load_obj(in);
			else if (endswith(f, ".off")) m = load_off(in);
			else if (endswith(f, ".ply")) m = load_ply_ascii(in);
			else if (endswith(f, ".stl")) m = load_stl(filename);
			else throw std::runtime_error("Invalid input mesh.");
			if (m.V.empty() || m.F.empty()) throw std::runtime_error("Invalid input mesh.");
			for (const auto &t : m.F)
				for (int k = 0; k < 3; ++k)
					if (t[k] < 0 || (size_t)t[k] >= m.V.size()) throw std::runtime_error("Invalid input mesh (a facet names a vertex that does not exist).");
			return m;
		}

		// SOS orientation and robust point-in-triangle test, Dexelize.cpp:56-92 (after SDFGen)
		int orientation(double x1, double y1, double x2, double y2, double &twice_signed_area)
		{
			twice_signed_area = y1 * x2 - x1 * y2;
			if (twice_signed_area > 0) return 1;
			else if (twice_signed_area < 0) return -1;
			else if (y2 > y1) return 1;
			else if (y2 < y1) return -1;
			else if (x1 > x2) return 1;
			else if (x1 < x2) return -1;
			else return 0;
		}

		bool point_in_triangle_2d(double x0, double y0, double x1, double y1, double x2, double y2, double x3, double y3,
		                          double &a, double &b, double &c)
		{
			x1 -= x0; x2 -= x0; x3 -= x0;
			y1 -= y0; y2 -= y0; y3 -= y0;
			const int signa = orientation(x2, y2, x3, y3, a);
			if (signa == 0) return false;
			const int signb = orientation(x3, y3, x1, y1, b);
			if (signb != signa) return false;
			const int signc = orientation(x1, y1, x2, y2, c);
			if (signc != signa) return false;
			const double sum = a + b + c;
			a /= sum; b /= sum; c /= sum;
			return true;
		}
	}

	CompressedVolume create_dexels(const std::string &filename, double &voxel_size, int padding, int num_voxels, int device)
	{
		const Mesh M = mesh_load(filename);
		std::array<double, 3> lo = M.V[0], hi = M.V[0];
		for (const auto &p : M.V) for (int d = 0; d < 3; ++d) { lo[d] = std::min(lo[d], p[d]); hi[d] = std::max(hi[d], p[d]); }
		const Vector3d extent{{hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]}};
		if (num_voxels > 0) voxel_size = std::max(extent[0], std::max(extent[1], extent[2])) / num_voxels;   // Dexelize.cpp:259-263
		CompressedVolume dexels(lo, extent, voxel_size, padding);
		const int nx = dexels.gridSize()[0], ny = dexels.gridSize()[1];
		const double spacing = dexels.spacing();
		if (device >= 0) {
			// the ray-marching loop (compute_sign, Dexelize.cpp:166-225) on the GPU: facets are the work items there
			vo_ctx *ctx = nullptr;
			if (vo_create(device, &ctx) != VO_OK) throw std::runtime_error("vo_create failed: no usable CUDA device (there is no CPU fallback)");
			std::vector<int32_t> tris;
			tris.reserve(3 * M.F.size());
			for (const auto &t : M.F) for (int k = 0; k < 3; ++k) tris.push_back(t[k]);
			vo_dvol *dv = nullptr;
			int rc = vo_dexelize_dev(ctx, M.V.size(), M.V[0].data(), M.F.size(), tris.data(), dexels.origin()[0], dexels.origin()[1],
			                         spacing, nx, ny, &dv, nullptr);
			uint64_t n = 0;
			std::vector<uint32_t> off((size_t)nx * ny + 1);
			std::vector<double> spans;
			if (rc == VO_OK) rc = vo_dvol_info(dv, nullptr, nullptr, &n, nullptr, nullptr);
			if (rc == VO_OK) { spans.resize(2 * n + 2); rc = vo_dvol_download(ctx, dv, off.data(), spans.data()); }
			const std::string err = rc == VO_OK ? "" : vo_last_error(ctx);
			if (dv) vo_dvol_free(ctx, dv);
			vo_destroy(ctx);
			if (rc != VO_OK) throw std::runtime_error("Assertion failed: " + err);
			dexels.from_csr(off.data(), spans.data());
			return dexels;
		}
		// Host loop (offset3d -x noop, which needs no GPU). Bucket the facets by the columns their xy bounding box can
		// cover; the exact test below is the reference's AABB query: a facet is handed to intersect_ray_z when its
		// bounding box contains the column centre (Dexelize.cpp:190-207, a query box of zero extent in x and y).
		std::vector<std::vector<int>> bucket((size_t)nx * ny);
		std::vector<std::array<double, 4>> fbox(M.F.size());
		for (int f = 0; f < (int)M.F.size(); ++f) {
			const auto &t = M.F[f];
			double bx0 = 1e300, bx1 = -1e300, by0 = 1e300, by1 = -1e300;
			for (int k = 0; k < 3; ++k) {
				bx0 = std::min(bx0, M.V[t[k]][0]); bx1 = std::max(bx1, M.V[t[k]][0]);
				by0 = std::min(by0, M.V[t[k]][1]); by1 = std::max(by1, M.V[t[k]][1]);
			}
			fbox[f] = {{bx0, bx1, by0, by1}};
			const int x0 = std::max(0, (int)std::floor((bx0 - dexels.origin()[0]) / spacing - 0.5) - 1);
			const int x1 = std::min(nx - 1, (int)std::ceil((bx1 - dexels.origin()[0]) / spacing - 0.5) + 1);
			const int y0 = std::max(0, (int)std::floor((by0 - dexels.origin()[1]) / spacing - 0.5) - 1);
			const int y1 = std::min(ny - 1, (int)std::ceil((by1 - dexels.origin()[1]) / spacing - 0.5) + 1);
			for (int y = y0; y <= y1; ++y) for (int x = x0; x <= x1; ++x) bucket[x + (size_t)nx * y].push_back(f);
		}
		for (int y = 0; y < ny; ++y)
			for (int x = 0; x < nx; ++x) {
				const auto c = dexels.dexelCenter(x, y);
				std::vector<double> inter;
				for (int f : bucket[x + (size_t)nx * y]) {
					const auto &b = fbox[f];
					if (c[0] < b[0] || c[0] > b[1] || c[1] < b[2] || c[1] > b[3]) continue;
					const auto &p1 = M.V[M.F[f][0]], &p2 = M.V[M.F[f][1]], &p3 = M.V[M.F[f][2]];
					double u, v, w;
					if (point_in_triangle_2d(c[0], c[1], p1[0], p1[1], p2[0], p2[1], p3[0], p3[1], u, v, w)) {
						// intersect_ray_z (Dexelize.cpp:136-162) drops facets whose projection is flat
						const double det = (p2[0] - p1[0]) * (p3[1] - p1[1]) - (p2[1] - p1[1]) * (p3[0] - p1[0]);
						if (det != 0) inter.push_back((u * p1[2] + v * p2[2] + w * p3[2]) / spacing);
					}
				}
				std::sort(inter.begin(), inter.end());
				if (inter.size() % 2) inter.pop_back();      // open surfaces: keep the list a list of intervals
				dexels.at(x, y) = inter;
			}
		return dexels;
	}

	// The hex mesh the reference builds (Dexelize.cpp:312-351): one hexahedron with eight vertices of its own per interval;
	// since no two cells share a vertex, geogram's compute_borders() makes every face of every cell a border facet and
	// remove_isolated() removes nothing. Written in the Medit text format (what GEO::mesh_save writes for a ".mesh"
	// name): Vertices, Quadrilaterals (the border), Hexahedra (Medit corner order: bottom face counter-clockwise, then
	// the top face).
	static void hex_mesh_dump(std::ostream &out, const CompressedVolume &dexels)
	{
		const double sp = dexels.spacing();
		size_t n = 0;
		for (int y = 0; y < dexels.gridSize()[1]; ++y)
			for (int x = 0; x < dexels.gridSize()[0]; ++x) n += dexels.at(x, y).size() / 2;
		out << "MeshVersionFormatted 2\nDimension 3\nVertices\n" << 8 * n << "\n";
		for (int y = 0; y < dexels.gridSize()[1]; ++y)
			for (int x = 0; x < dexels.gridSize()[0]; ++x)
				for (size_t i = 0; 2 * i + 1 < dexels.at(x, y).size(); ++i) {
					const double z0 = dexels.at(x, y)[2 * i] * sp, z1 = dexels.at(x, y)[2 * i + 1] * sp;
					const double x0 = dexels.origin()[0] + x * sp, x1 = dexels.origin()[0] + (x + 1) * sp;
					const double y0 = dexels.origin()[1] + y * sp, y1 = dexels.origin()[1] + (y + 1) * sp;
					// (the reference's corner order, Dexelize.cpp:326-329)
					const double P[8][3] = {{x0, y0, z1}, {x1, y0, z1}, {x0, y1, z1}, {x1, y1, z1}, {x0, y0, z0}, {x1, y0, z0}, {x0, y1, z0}, {x1, y1, z0}};
					for (auto &p : P) out << p[0] << ' ' << p[1] << ' ' << p[2] << " 0\n";
				}
		out << "Quadrilaterals\n" << 6 * n << "\n";
		const int Q[6][4] = {{1, 2, 4, 3}, {5, 7, 8, 6}, {1, 5, 6, 2}, {3, 4, 8, 7}, {1, 3, 7, 5}, {2, 6, 8, 4}};   // outward normals
		for (size_t c = 0; c < n; ++c)
			for (auto &q : Q) out << 8 * c + q[0] << ' ' << 8 * c + q[1] << ' ' << 8 * c + q[2] << ' ' << 8 * c + q[3] << " 0\n";
		out << "Hexahedra\n" << n << "\n";
		const int H[8] = {5, 6, 8, 7, 1, 2, 4, 3};
		for (size_t c = 0; c < n; ++c) {
			for (int k : H) out << 8 * c + k << ' ';
			out << "0\n";
		}
		out << "End\n";
	}

	// dexels -> one box per interval (Dexelize.cpp:312-351) written as a Medit hex mesh (".mesh": cells + border facets,
	// see above) or as an OBJ quad mesh (the border facets alone), or the interval end points as an .xyz point list
	// (Dexelize.cpp:289-310). A ".vol" / ".txt" name writes the reference's own text format (CompressedVolume::save).
	void dexel_dump(const std::string &filename, const CompressedVolume &dexels)
	{
		std::ofstream out(filename);
		if (!out) throw std::runtime_error("Cannot write " + filename);
		out << std::setprecision(17);
		const std::string f = lower(filename);
		if (endswith(f, ".vol") || endswith(f, ".txt")) { dexels.save(out); return; }
		if (endswith(f, ".mesh")) { hex_mesh_dump(out, dexels); return; }
		const double sp = dexels.spacing();
		const bool points = endswith(f, ".xyz");
		size_t v = 0;
		for (int y = 0; y < dexels.gridSize()[1]; ++y)
			for (int x = 0; x < dexels.gridSize()[0]; ++x)
				for (size_t i = 0; 2 * i + 1 < dexels.at(x, y).size(); ++i) {
					const double z0 = dexels.at(x, y)[2 * i] * sp, z1 = dexels.at(x, y)[2 * i + 1] * sp;
					if (points) {
						const double px = dexels.origin()[0] + (x + 0.5) * sp, py = dexels.origin()[1] + (y + 0.5) * sp;
						out << px << ' ' << py << ' ' << z0 << "\n" << px << ' ' << py << ' ' << z1 << "\n";
						continue;
					}
					const double x0 = dexels.origin()[0] + x * sp, x1 = dexels.origin()[0] + (x + 1) * sp;
					const double y0 = dexels.origin()[1] + y * sp, y1 = dexels.origin()[1] + (y + 1) * sp;
					const double P[8][3] = {{x0, y0, z1}, {x1, y0, z1}, {x0, y1, z1}, {x1, y1, z1}, {x0, y0, z0}, {x1, y0, z0}, {x0, y1, z0}, {x1, y1, z0}};
					for (auto &p : P) out << "v " << p[0] << ' ' << p[1] << ' ' << p[2] << "\n";
					const int Q[6][4] = {{1, 2, 4, 3}, {5, 7, 8, 6}, {1, 5, 6, 2}, {3, 4, 8, 7}, {1, 3, 7, 5}, {2, 6, 8, 4}};
					for (auto &q : Q) out << "f " << v + q[0] << ' ' << v + q[1] << ' ' << v + q[2] << ' ' << v + q[3] << "\n";
					v += 8;
				}
	}
}

namespace voroffset
{
	bool DoubleCompressedImage::isValid() const
	{
		for (const auto &row : m_Rays) {
			if (row.size() % 2 != 0) { std::cerr << "row size % 2 != 0" << std::endl; return false; }
			Scalar lastEvent = -1;
			for (const auto &val : row) {
				if (val < lastEvent) { std::cerr << "val > lastEvent" << std::endl; return false; }
				lastEvent = val;
			}
			if (lastEvent > m_XSize) { std::cerr << "lastEvent <= xsize" << std::endl; return false; }
		}
		return true;
	}

	void DoubleCompressedImage::save(std::ostream &out) const
	{
		out << std::setprecision(17) << m_XSize << " " << m_Rays.size() << "\n";
		for (const auto &row : m_Rays) {
			out << row.size();
			for (const auto &v : row) out << ' ' << v;
			out << '\n';
		}
	}

	void DoubleCompressedImage::load(std::istream &in)
	{
		unsigned int num_cols;
		in >> m_XSize >> num_cols;
		m_Rays.assign(num_cols, {});
		for (auto &row : m_Rays) {
			size_t size;
			in >> size;
			if (size % 2) throw std::runtime_error("Assertion failed: size % 2 == 0");
			row.resize(size);
			for (auto &v : row) in >> v;
		}
	}

	void DoubleCompressedImage::apply(int op, double r)
	{
		vo_ctx *ctx = make_ctx(0);
		std::vector<uint32_t> off(m_Rays.size() + 1, 0);
		std::vector<double> spans;
		for (size_t i = 0; i < m_Rays.size(); ++i) {
			spans.insert(spans.end(), m_Rays[i].begin(), m_Rays[i].begin() + 2 * (m_Rays[i].size() / 2));
			off[i + 1] = (uint32_t)(spans.size() / 2);
		}
		spans.resize(spans.size() + 2);
		uint32_t *o_off = nullptr;
		double *o_spans = nullptr;
		uint64_t n = 0;
		double ms = 0;
		const int rc = vo_morph2d(ctx, op, height(), width(), off.data(), spans.data(), r, &o_off, &o_spans, &n, &ms);
		if (rc != VO_OK) { const std::string msg = vo_last_error(ctx); vo_destroy(ctx); throw std::runtime_error("voroffset_b200: " + msg); }
		for (int i = 0; i < height(); ++i) m_Rays[i].assign(o_spans + 2 * (size_t)o_off[i], o_spans + 2 * (size_t)o_off[i + 1]);
		vo_free(o_off);
		vo_free(o_spans);
		vo_destroy(ctx);
		if (op != VO_OP2D_NEGATE && !isValid()) throw std::runtime_error("Assertion failed: isValid() == true");
	}
	void DoubleCompressedImage::negate() { apply(VO_OP2D_NEGATE, 0); }
	void DoubleCompressedImage::dilate(double r) { apply(VO_OP2D_DILATE, r); }
	void DoubleCompressedImage::erode(double r) { apply(VO_OP2D_ERODE, r); }
	void DoubleCompressedImage::close(double r) { apply(VO_OP2D_CLOSE, r); }
	void DoubleCompressedImage::open(double r) { apply(VO_OP2D_OPEN, r); }
}
