// offset3d re-hosted on voroffset_b200: same flags, same operations and the same JSON keys as
// app/cli3d/offset3d.cpp:38-51,83-96,156-177 of the reference; CLI11 / nlohmann-json / geogram replaced by
// a few lines of plain C++. The morphology runs on the GPU through the C ABI (no CPU fallback).
// Input: a triangle mesh (.obj .off .stl .ply-ascii) or a volume in the reference's text format (.vol).
#include "vo_host.hpp"

#include <chrono>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <memory>
#include <thread>

namespace vor3d = voroffset3d;

static bool ends(const std::string &s, const char *suf) { const size_t n = std::strlen(suf); return s.size() >= n && s.compare(s.size() - n, n, suf) == 0; }

static int usage(const char *msg)
{
	if (msg) std::cerr << msg << "\n";
	std::cerr << "Offset3D\nUsage: offset3d [OPTIONS] input [output]\n"
	             "  input,-i,--input TEXT       Input model (required)\n"
	             "  output,-o,--output TEXT     Output model (default output.obj)\n"
	             "  -j,--json TEXT              Output json file\n"
	             "  -d,--dexels_size FLOAT      Size of a dexel (in mm)\n"
	             "  -n,--num_dexels INT         Number of dexels (-1 to use dexel size instead)\n"
	             "  -p,--padding INT            Padding (in #dexels)\n"
	             "  -t,--num_thread UINT        Number of threads (kept for compatibility; the GPU path ignores it)\n"
	             "  -r,--radius FLOAT           Dilation/erosion radius (in #dexels)\n"
	             "  -m,--method {ours,brute_force}\n"
	             "  -x,--apply {noop,dilation,erosion,closing,opening}\n"
	             "  -f,--force                  Overwrite output file\n"
	             "  -u,--radius_in_mm           Radius is given in mm instead\n"
	             "  -g,--gpus INT               Number of GPUs (default 1): the grid is cut into y-slabs, halo rows travel with NCCL\n";
	return msg ? 1 : 0;
}

int main(int argc, char *argv[])
{
	struct {
		std::string input, output_mesh = "output.obj", output_json = "", method = "ours", operation = "dilation";
		double radius = 8, dexels_size = 1;
		int padding = 0, num_dexels = 256, gpus = 1;
		unsigned int num_thread = std::max(1u, std::thread::hardware_concurrency());
		bool force = false, radius_in_mm = false;
	} args;

	int positional = 0;
	for (int i = 1; i < argc; ++i) {
		const std::string a = argv[i];
		auto val = [&](const char *name) -> std::string { if (i + 1 >= argc) { usage((std::string(name) + " needs a value").c_str()); std::exit(1); } return argv[++i]; };
		if (a == "-h" || a == "--help") return usage(nullptr);
		else if (a == "-i" || a == "--input") args.input = val("-i");
		else if (a == "-o" || a == "--output") args.output_mesh = val("-o");
		else if (a == "-j" || a == "--json") args.output_json = val("-j");
		else if (a == "-d" || a == "--dexels_size") args.dexels_size = std::stod(val("-d"));
		else if (a == "-n" || a == "--num_dexels") args.num_dexels = std::stoi(val("-n"));
		else if (a == "-p" || a == "--padding") args.padding = std::stoi(val("-p"));
		else if (a == "-t" || a == "--num_thread") args.num_thread = (unsigned)std::stoul(val("-t"));
		else if (a == "-r" || a == "--radius") args.radius = std::stod(val("-r"));
		else if (a == "-m" || a == "--method") args.method = val("-m");
		else if (a == "-x" || a == "--apply") args.operation = val("-x");
		else if (a == "-f" || a == "--force") args.force = true;
		else if (a == "-u" || a == "--radius_in_mm") args.radius_in_mm = true;
		else if (a == "-g" || a == "--gpus") args.gpus = std::stoi(val("-g"));
		else if (!a.empty() && a[0] == '-') return usage(("unknown option " + a).c_str());
		else if (positional == 0) { args.input = a; ++positional; }
		else if (positional == 1) { args.output_mesh = a; ++positional; }
		else return usage("too many positional arguments");
	}
	if (args.input.empty()) return usage("input is required");
	if (!std::ifstream(args.input)) return usage(("File does not exist: " + args.input).c_str());
	if (args.method != "ours" && args.method != "brute_force") return usage("-m must be one of {ours,brute_force}");
	const char *ops[] = {"noop", "dilation", "erosion", "closing", "opening"};
	bool okop = false;
	for (auto o : ops) okop |= args.operation == o;
	if (!okop) return usage("-x must be one of {noop,dilation,erosion,closing,opening}");

	try {
		vor3d::CompressedVolume input, output;
		double time_1 = 0, time_2 = 0;
		const auto t0 = std::chrono::steady_clock::now();

		// Load input model and dexelize (offset3d.cpp:66-70)
		if (ends(args.input, ".vol")) { std::ifstream in(args.input); input.load(in); }
		else input = vor3d::create_dexels(args.input, args.dexels_size, args.padding, args.num_dexels,
		                                  args.operation == "noop" ? -1 : 0);   // on the GPU whenever one is needed anyway

		if (args.radius_in_mm) args.radius /= input.spacing();                                    // offset3d.cpp:73-75
		std::cout << "[Stats] Grid size: " << input.gridSize()[0] << " " << input.gridSize()[1] << "\n"
		          << "[Stats] Spacing (in mm): " << input.spacing() << "\n"
		          << "[Stats] Origin (in mm): " << input.origin()[0] << " " << input.origin()[1] << " " << input.origin()[2] << "\n"
		          << "[Stats] Extent (in mm): " << input.extent()[0] << " " << input.extent()[1] << " " << input.extent()[2] << "\n"
		          << "[Stats] Radius (in #dexels): " << args.radius << "\n"
		          << "[Stats] Number of threads: " << args.num_thread << std::endl;
		const int num_segments = input.numSegments();

		// Create offset operator (offset3d.cpp:104-112)
		std::unique_ptr<vor3d::VoronoiMorpho> op;
		if (args.operation != "noop") {          // (the GPU context is only needed when something is computed)
			if (args.method == "ours") op = std::make_unique<vor3d::VoronoiMorphoVorPower>(0, std::max(1, args.gpus));
			else op = std::make_unique<vor3d::VoronoiMorphoBruteForce>(0, std::max(1, args.gpus));
		}

		// Apply operation (offset3d.cpp:116-136)
		if (args.operation == "noop") output = input;
		else if (args.operation == "erosion") op->erosion(input, output, args.radius, time_1, time_2);
		else if (args.operation == "dilation") op->dilation(input, output, args.radius, time_1, time_2);
		else if (args.operation == "closing") {
			vor3d::CompressedVolume tmp;
			op->dilation(input, tmp, args.radius, time_1, time_2);
			op->erosion(tmp, output, args.radius, time_1, time_2);
		} else if (args.operation == "opening") {
			vor3d::CompressedVolume tmp;
			op->erosion(input, tmp, args.radius, time_1, time_2);
			op->dilation(tmp, output, args.radius, time_1, time_2);
		} else throw std::invalid_argument("Operation");

		// Saving (offset3d.cpp:138-154)
		if (!args.output_mesh.empty()) {
			if (std::ifstream(args.output_mesh) && !args.force)
				std::cout << "[Save] Output mesh already exists. Please use -f to force overwriting." << std::endl;
			else vor3d::dexel_dump(args.output_mesh, output);
		}
		const double time = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();

		if (!args.output_json.empty()) {
			if (std::ifstream(args.output_json) && !args.force)
				std::cout << "[Save] Output json already exists. Please use -f to force overwriting." << std::endl;
			else {
				std::ofstream o(args.output_json);
				o << std::setprecision(17) << "{\n"
				  << "    \"grid_size\": [\n        " << input.gridSize()[0] << ",\n        " << input.gridSize()[1] << "\n    ],\n"
				  << "    \"method\": \"" << args.method << "\",\n"
				  << "    \"model_name\": \"" << args.input << "\",\n"
				  << "    \"num_dexels\": " << args.num_dexels << ",\n"
				  << "    \"num_gpus\": " << std::max(1, args.gpus) << ",\n"
				  << "    \"num_segments\": " << num_segments << ",\n"
				  << "    \"num_threads\": " << args.num_thread << ",\n"
				  << "    \"operation\": \"" << args.operation << "\",\n"
				  << "    \"padding\": " << input.padding() << ",\n"
				  << "    \"radius\": " << args.radius << ",\n"
				  << "    \"time\": " << time << ",\n"
				  << "    \"time_first_pass\": " << time_1 << ",\n"
				  << "    \"time_second_pass\": " << time_2 << ",\n"
				  << "    \"voxel_size\": " << input.spacing() << "\n}" << std::endl;
			}
		}
	} catch (const std::exception &e) {
		std::cerr << "error: " << e.what() << std::endl;
		return 1;
	}
	return 0;
}
