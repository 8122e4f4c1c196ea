// Same box, same pattern (8 chunks per direction, one buffer each way): free-running vs download chunk k gated on upload
// chunk k, pinned memory allocated with default flags vs cudaHostAllocMapped vs write-combined upload buffer.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/pcie_probe5 scripts/probes/pcie_probe5.cu
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
	char *d_u, *d_d;
	CK(cudaMalloc(&d_u, UP)); CK(cudaMalloc(&d_d, DN));
	cudaStream_t s_in, s_out;
	CK(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
	std::vector<cudaEvent_t> ev(64);
	for (auto &e : ev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
	auto cut = [](size_t n, int k, int parts) { return k >= parts ? n : (n * k / parts) & ~(size_t)255; };
	const unsigned int flags[3] = {cudaHostAllocDefault, cudaHostAllocMapped, cudaHostAllocWriteCombined};
	const char *names[3] = {"default", "mapped", "write-combined (upload buffer)"};
	for (int f = 0; f < 3; ++f) {
		char *h_u, *h_d;
		CK(cudaHostAlloc(&h_u, UP, flags[f])); CK(cudaHostAlloc(&h_d, DN, f == 2 ? cudaHostAllocDefault : flags[f]));
		for (size_t i = 0; i < UP; i += 4096) h_u[i] = 1;
		for (size_t i = 0; i < DN; i += 4096) h_d[i] = 1;
		for (int B : {1, 8}) for (int gated = 0; gated < 2; ++gated) for (int lag = 0; lag < (gated ? 3 : 1); ++lag) {
			double best = 1e9;
			for (int rep = 0; rep < 10; ++rep) {
				CK(cudaDeviceSynchronize());
				const double t0 = now_ms();
				for (int b = 0; b < B; ++b) {
					CK(cudaMemcpyAsync(d_u + cut(UP, b, B), h_u + cut(UP, b, B), cut(UP, b + 1, B) - cut(UP, b, B), cudaMemcpyHostToDevice, s_in));
					CK(cudaEventRecord(ev[b], s_in));
				}
				for (int b = 0; b < B; ++b) {
					if (gated) CK(cudaStreamWaitEvent(s_out, ev[std::min(B - 1, b + lag)], 0));
					CK(cudaMemcpyAsync(h_d + cut(DN, b, B), d_d + cut(DN, b, B), cut(DN, b + 1, B) - cut(DN, b, B), cudaMemcpyDeviceToHost, s_out));
				}
				CK(cudaStreamSynchronize(s_out)); CK(cudaStreamSynchronize(s_in));
				best = std::min(best, now_ms() - t0);
			}
			std::printf("%-32s chunks %d %-5s lag %d: %.3f ms  %.1f GB/s\n", names[f], B, gated ? "gated" : "free", lag, best, (UP + DN) / best / 1e6);
			std::fflush(stdout);
		}
		CK(cudaFreeHost(h_u)); CK(cudaFreeHost(h_d));
	}
	return 0;
}
