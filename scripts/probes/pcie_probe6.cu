// Does the ALIGNMENT of the chunk boundaries matter? 8 chunks per direction, download chunk k gated on upload chunk k;
// boundaries rounded down to 16 / 64 / 128 / 256 / 4096 bytes, and odd boundaries (16-byte aligned, 16 mod 256).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/pcie_probe6 scripts/probes/pcie_probe6.cu
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
	const size_t UP = 57266784, DN = 61424384;
	char *d_u, *d_d, *h_u, *h_d;
	CK(cudaMalloc(&d_u, UP + 4096)); CK(cudaMalloc(&d_d, DN + 4096));
	CK(cudaHostAlloc(&h_u, UP + 4096, 0)); CK(cudaHostAlloc(&h_d, DN + 4096, 0));
	for (size_t i = 0; i < UP; i += 4096) h_u[i] = 1;
	for (size_t i = 0; i < DN; i += 4096) h_d[i] = 1;
	cudaStream_t s_in, s_out;
	CK(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
	std::vector<cudaEvent_t> ev(64);
	for (auto &e : ev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
	const int B = 8;
	for (int gated = 0; gated < 2; ++gated)
	for (size_t align : {16, 64, 128, 256, 4096, 0}) {          // 0: boundaries at 16 mod 256
		auto cut = [&](size_t n, int k) -> size_t {
			if (k <= 0) return 0;
			if (k >= B) return n;
			const size_t c = n * k / B;
			return align ? c & ~(align - 1) : (c & ~(size_t)255) + 16;
		};
		double best = 1e9;
		for (int rep = 0; rep < 10; ++rep) {
			CK(cudaDeviceSynchronize());
			const double t0 = now_ms();
			for (int b = 0; b < B; ++b) {
				CK(cudaMemcpyAsync(d_u + cut(UP, b), h_u + cut(UP, b), cut(UP, b + 1) - cut(UP, b), cudaMemcpyHostToDevice, s_in));
				CK(cudaEventRecord(ev[b], s_in));
			}
			for (int b = 0; b < B; ++b) {
				if (gated) CK(cudaStreamWaitEvent(s_out, ev[b], 0));
				CK(cudaMemcpyAsync(h_d + cut(DN, b), d_d + cut(DN, b), cut(DN, b + 1) - cut(DN, b), cudaMemcpyDeviceToHost, s_out));
			}
			CK(cudaStreamSynchronize(s_out)); CK(cudaStreamSynchronize(s_in));
			best = std::min(best, now_ms() - t0);
		}
		std::printf("%-5s boundaries aligned to %4zu bytes%s: %.3f ms  %.1f GB/s\n", gated ? "gated" : "free", align ? align : (size_t)16, align ? "" : " (16 mod 256)", best, (UP + DN) / best / 1e6);
		std::fflush(stdout);
	}
	return 0;
}
