// The two copies of a band (offsets, spans) as two cudaMemcpyAsync calls or as ONE cudaMemcpyBatchAsync (CUDA 12.8+):
// 8 bands, 256-byte aligned boundaries, download band k gated on upload band k.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/pcie_probe7 scripts/probes/pcie_probe7.cu
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
	const size_t UO = 16777216, US = 40489472, DO_ = 16777216, DS = 44647168;
	char *h_uo, *h_us, *h_do, *h_ds, *d_uo, *d_us, *d_do, *d_ds;
	CK(cudaHostAlloc(&h_uo, UO, 0)); CK(cudaHostAlloc(&h_us, US, 0)); CK(cudaHostAlloc(&h_do, DO_, 0)); CK(cudaHostAlloc(&h_ds, DS, 0));
	CK(cudaMalloc(&d_uo, UO)); CK(cudaMalloc(&d_us, US)); CK(cudaMalloc(&d_do, DO_)); CK(cudaMalloc(&d_ds, DS));
	cudaStream_t s_in, s_out;
	CK(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
	std::vector<cudaEvent_t> ev(64);
	for (auto &e : ev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
	auto cut = [](size_t n, int k, int parts) { return k >= parts ? n : (n * k / parts) & ~(size_t)255; };
	cudaMemcpyAttributes attr{};
	attr.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
	auto two = [&](bool batch, void *d0, void *s0, size_t n0, void *d1, void *s1, size_t n1, cudaMemcpyKind kind, cudaStream_t st) {
		if (!batch) { CK(cudaMemcpyAsync(d0, s0, n0, kind, st)); CK(cudaMemcpyAsync(d1, s1, n1, kind, st)); return; }
		void *dsts[2] = {d0, d1}, *srcs[2] = {s0, s1};
		size_t sizes[2] = {n0, n1}, idx[1] = {0}, fail = 0;
		CK(cudaMemcpyBatchAsync(dsts, srcs, sizes, 2, &attr, idx, 1, &fail, st));
	};
	for (int B : {8, 16})
	for (int batch = 0; batch < 2; ++batch) {
		double best = 1e9;
		for (int rep = 0; rep < 10; ++rep) {
			CK(cudaDeviceSynchronize());
			const double t0 = now_ms();
			for (int b = 0; b < B; ++b) {
				two(batch, d_uo + cut(UO, b, B), h_uo + cut(UO, b, B), cut(UO, b + 1, B) - cut(UO, b, B),
				    d_us + cut(US, b, B), h_us + cut(US, b, B), cut(US, b + 1, B) - cut(US, b, B), cudaMemcpyHostToDevice, s_in);
				CK(cudaEventRecord(ev[b], s_in));
			}
			for (int b = 0; b < B; ++b) {
				CK(cudaStreamWaitEvent(s_out, ev[b], 0));
				two(batch, h_do + cut(DO_, b, B), d_do + cut(DO_, b, B), cut(DO_, b + 1, B) - cut(DO_, b, B),
				    h_ds + cut(DS, b, B), d_ds + cut(DS, b, B), cut(DS, b + 1, B) - cut(DS, b, B), cudaMemcpyDeviceToHost, s_out);
			}
			CK(cudaStreamSynchronize(s_out)); CK(cudaStreamSynchronize(s_in));
			best = std::min(best, now_ms() - t0);
		}
		std::printf("bands %2d, %s: %.3f ms  %.1f GB/s\n", B, batch ? "one cudaMemcpyBatchAsync per band and direction" : "two cudaMemcpyAsync per band and direction", best, (UO + US + DO_ + DS) / best / 1e6);
		std::fflush(stdout);
	}
	return 0;
}
