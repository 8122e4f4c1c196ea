// PCIe duplex behaviour of the box under the transfer patterns the banded host-buffer call could use (no compute):
// copy-engine copies in n chunks per direction (free-running, or download chunk k gated on upload chunk k) and
// SM-driven copies through mapped pinned memory (a few CTAs streaming 16-byte loads / stores) for either direction.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/pcie_probe2 scripts/probes/pcie_probe2.cu
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); std::exit(1); } } while (0)

__global__ void k_copy(const uint4 *__restrict__ src, uint4 *__restrict__ dst, size_t n)
{
	const size_t stride = (size_t)gridDim.x * blockDim.x;
	size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	for (; i + 3 * stride < n; i += 4 * stride) {          // four independent 16-byte transfers in flight per thread
		const uint4 a = src[i], b = src[i + stride], c = src[i + 2 * stride], d = src[i + 3 * stride];
		dst[i] = a; dst[i + stride] = b; dst[i + 2 * stride] = c; dst[i + 3 * stride] = d;
	}
	for (; i < n; i += stride) dst[i] = src[i];
}

static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main()
{
	const size_t UP = 57266784, DN = 61424384;              // bytes of the C5 step, multiples of 16
	char *h_in, *h_out, *d_in, *d_out;
	CK(cudaHostAlloc(&h_in, UP, cudaHostAllocMapped)); CK(cudaHostAlloc(&h_out, DN, cudaHostAllocMapped));
	CK(cudaMalloc(&d_in, UP)); CK(cudaMalloc(&d_out, DN));
	for (size_t i = 0; i < UP; i += 4096) h_in[i] = 1;
	CK(cudaMemset(d_out, 1, DN));
	cudaStream_t s_in, s_out;
	CK(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
	std::vector<cudaEvent_t> ev(64);
	for (auto &e : ev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
	auto cut = [](size_t n, int k, int parts) { return (n * k / parts) & ~(size_t)255; };
	// mode of a direction: 0 none, 1 copy engine, 2 SM kernel with `ctas` CTAs
	auto run = [&](int up_mode, int dn_mode, int chunks, bool gated, int ctas) {
		double best = 1e9;
		for (int rep = 0; rep < 8; ++rep) {
			CK(cudaDeviceSynchronize());
			const double t0 = now_ms();
			for (int k = 0; k < chunks; ++k) {
				const size_t a0 = cut(UP, k, chunks), a1 = k + 1 < chunks ? cut(UP, k + 1, chunks) : UP;
				if (up_mode == 1) CK(cudaMemcpyAsync(d_in + a0, h_in + a0, a1 - a0, cudaMemcpyHostToDevice, s_in));
				if (up_mode == 2) k_copy<<<ctas, 256, 0, s_in>>>((const uint4 *)(h_in + a0), (uint4 *)(d_in + a0), (a1 - a0) / 16);
				if (gated) CK(cudaEventRecord(ev[k], s_in));
			}
			for (int k = 0; k < chunks; ++k) {
				const size_t a0 = cut(DN, k, chunks), a1 = k + 1 < chunks ? cut(DN, k + 1, chunks) : DN;
				if (gated) CK(cudaStreamWaitEvent(s_out, ev[k], 0));
				if (dn_mode == 1) CK(cudaMemcpyAsync(h_out + a0, d_out + a0, a1 - a0, cudaMemcpyDeviceToHost, s_out));
				if (dn_mode == 2) k_copy<<<ctas, 256, 0, s_out>>>((const uint4 *)(d_out + a0), (uint4 *)(h_out + a0), (a1 - a0) / 16);
			}
			CK(cudaStreamSynchronize(s_in)); CK(cudaStreamSynchronize(s_out));
			best = std::min(best, now_ms() - t0);
		}
		const double bytes = (up_mode ? UP : 0) + (dn_mode ? DN : 0);
		std::printf("up %-6s down %-6s chunks %2d %-6s ctas %3d : %.3f ms  %.1f GB/s\n", up_mode == 0 ? "-" : up_mode == 1 ? "CE" : "SM",
		            dn_mode == 0 ? "-" : dn_mode == 1 ? "CE" : "SM", chunks, gated ? "gated" : "free", ctas, best, bytes / best / 1e6);
		std::fflush(stdout);
	};
	run(1, 0, 1, false, 0); run(0, 1, 1, false, 0); run(1, 1, 1, false, 0);
	for (int c : {2, 4, 8, 16, 32}) run(1, 1, c, false, 0);
	for (int c : {4, 8, 16}) run(1, 1, c, true, 0);
	for (int c : {8}) { run(1, 0, c, false, 0); run(0, 1, c, false, 0); }
	for (int ctas : {4, 8, 16, 32, 148}) { run(2, 0, 1, false, ctas); run(0, 2, 1, false, ctas); }
	for (int ctas : {8, 16, 32}) { run(1, 2, 1, false, ctas); run(2, 1, 1, false, ctas); run(2, 2, 1, false, ctas); }
	for (int ctas : {8, 16, 32}) { run(1, 2, 8, false, ctas); run(1, 2, 8, true, ctas); run(2, 2, 8, true, ctas); run(2, 1, 8, true, ctas); }
	return 0;
}
