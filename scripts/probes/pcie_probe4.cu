// Follow-up to pcie_probe3: 8 bands, downloads enqueued up front behind the band's upload event. What separates the
// two-copies-per-band pattern (72 GB/s) from one copy per band out of one buffer (84 GB/s)?
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/pcie_probe4 scripts/probes/pcie_probe4.cu
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); std::exit(1); } } while (0)
static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main()
{
	const size_t UO = 16777232, US = 40489552, DO_ = 16777232, DS = 44647152;
	char *h_u, *h_d, *d_u, *d_d;                           // [offsets | spans] in one allocation per direction and side
	CK(cudaHostAlloc(&h_u, UO + US, 0)); CK(cudaHostAlloc(&h_d, DO_ + DS, 0));
	CK(cudaMalloc(&d_u, UO + US)); CK(cudaMalloc(&d_d, DO_ + DS));
	cudaStream_t s_in, s_in2, s_out, s_out2;
	CK(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
	CK(cudaStreamCreateWithFlags(&s_in2, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s_out2, cudaStreamNonBlocking));
	std::vector<cudaEvent_t> ev(64), ev2(64);
	for (auto &e : ev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
	for (auto &e : ev2) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
	auto cut = [](size_t n, int k, int parts) { return k >= parts ? n : (n * k / parts) & ~(size_t)15; };
	// mode 0: one copy per band and direction (a band = the same fraction of the whole allocation)
	// mode 1: two copies per band and direction on one stream (offsets, then spans)
	// mode 2: two copies per band and direction, offsets and spans on streams of their own
	// mode 3: as 1, spans first
	// dgroup: bands per download group (the group waits for its last band)
	auto run = [&](int B, int mode, int dgroup) {
		double best = 1e9;
		for (int rep = 0; rep < 10; ++rep) {
			CK(cudaDeviceSynchronize());
			const double t0 = now_ms();
			for (int b = 0; b < B; ++b) {
				if (mode == 0) CK(cudaMemcpyAsync(d_u + cut(UO + US, b, B), h_u + cut(UO + US, b, B), cut(UO + US, b + 1, B) - cut(UO + US, b, B), cudaMemcpyHostToDevice, s_in));
				else {
					auto offs = [&](cudaStream_t s) { CK(cudaMemcpyAsync(d_u + cut(UO, b, B), h_u + cut(UO, b, B), cut(UO, b + 1, B) - cut(UO, b, B), cudaMemcpyHostToDevice, s)); };
					auto spans = [&](cudaStream_t s) { CK(cudaMemcpyAsync(d_u + UO + cut(US, b, B), h_u + UO + cut(US, b, B), cut(US, b + 1, B) - cut(US, b, B), cudaMemcpyHostToDevice, s)); };
					if (mode == 1) { offs(s_in); spans(s_in); }
					if (mode == 3) { spans(s_in); offs(s_in); }
					if (mode == 2) { offs(s_in2); CK(cudaEventRecord(ev2[b], s_in2)); spans(s_in); }
				}
				CK(cudaEventRecord(ev[b], s_in));
			}
			for (int g0 = 0; g0 < B; g0 += dgroup) {
				const int g1 = std::min(B, g0 + dgroup);
				CK(cudaStreamWaitEvent(s_out, ev[g1 - 1], 0));
				if (mode == 2) { CK(cudaStreamWaitEvent(s_out, ev2[g1 - 1], 0)); CK(cudaStreamWaitEvent(s_out2, ev[g1 - 1], 0)); CK(cudaStreamWaitEvent(s_out2, ev2[g1 - 1], 0)); }
				if (mode == 0) CK(cudaMemcpyAsync(h_d + cut(DO_ + DS, g0, B), d_d + cut(DO_ + DS, g0, B), cut(DO_ + DS, g1, B) - cut(DO_ + DS, g0, B), cudaMemcpyDeviceToHost, s_out));
				else {
					auto offs = [&](cudaStream_t s) { CK(cudaMemcpyAsync(h_d + cut(DO_, g0, B), d_d + cut(DO_, g0, B), cut(DO_, g1, B) - cut(DO_, g0, B), cudaMemcpyDeviceToHost, s)); };
					auto spans = [&](cudaStream_t s) { CK(cudaMemcpyAsync(h_d + DO_ + cut(DS, g0, B), d_d + DO_ + cut(DS, g0, B), cut(DS, g1, B) - cut(DS, g0, B), cudaMemcpyDeviceToHost, s)); };
					if (mode == 1) { offs(s_out); spans(s_out); }
					if (mode == 3) { spans(s_out); offs(s_out); }
					if (mode == 2) { offs(s_out2); spans(s_out); }
				}
			}
			CK(cudaStreamSynchronize(s_out)); CK(cudaStreamSynchronize(s_out2)); CK(cudaStreamSynchronize(s_in)); CK(cudaStreamSynchronize(s_in2));
			best = std::min(best, now_ms() - t0);
		}
		std::printf("bands %2d mode %d download groups of %d: %.3f ms  %.1f GB/s\n", B, mode, dgroup, best, (UO + US + DO_ + DS) / best / 1e6);
		std::fflush(stdout);
	};
	for (int mode : {0, 1, 2, 3}) run(8, mode, 1);
	for (int mode : {0, 1}) { run(8, mode, 2); run(4, mode, 1); run(16, mode, 1); run(16, mode, 2); run(16, mode, 4); }
	return 0;
}
