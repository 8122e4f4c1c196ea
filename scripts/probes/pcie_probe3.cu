// Which part of the host-buffer call's transfer pattern costs the duplex rate (no compute): per band two copies per
// direction (offsets, spans) instead of one; downloads enqueued by the host after a round trip per band (an 8-byte D2H
// copy on a control stream, cudaEventSynchronize) instead of up front behind an event.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/pcie_probe3 scripts/probes/pcie_probe3.cu
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
	const size_t UO = 16777220, US = 40489552, DO_ = 16777220, DS = 44647152;
	char *h_uo, *h_us, *h_do, *h_ds, *d_uo, *d_us, *d_do, *d_ds;
	unsigned long long *h_tot, *d_tot;
	CK(cudaHostAlloc(&h_uo, UO, 0)); CK(cudaHostAlloc(&h_us, US, 0)); CK(cudaHostAlloc(&h_do, DO_, 0)); CK(cudaHostAlloc(&h_ds, DS, 0));
	CK(cudaHostAlloc(&h_tot, 64 * 8, 0));
	CK(cudaMalloc(&d_uo, UO)); CK(cudaMalloc(&d_us, US)); CK(cudaMalloc(&d_do, DO_)); CK(cudaMalloc(&d_ds, DS)); CK(cudaMalloc(&d_tot, 64 * 8));
	CK(cudaMemset(d_tot, 0, 64 * 8));
	cudaStream_t s_in, s_out, s_ctl, s_mid;
	CK(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
	CK(cudaStreamCreateWithFlags(&s_ctl, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s_mid, cudaStreamNonBlocking));
	std::vector<cudaEvent_t> ev(64), ev2(64);
	for (auto &e : ev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
	for (auto &e : ev2) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
	auto cut = [](size_t n, int k, int parts) { return k >= parts ? n : (n * k / parts) & ~(size_t)15; };
	auto run = [&](int B, bool two, bool host_driven, bool offsets_first) {
		double best = 1e9;
		for (int rep = 0; rep < 10; ++rep) {
			CK(cudaDeviceSynchronize());
			const double t0 = now_ms();
			if (offsets_first) CK(cudaMemcpyAsync(d_uo, h_uo, UO, cudaMemcpyHostToDevice, s_in));        // all offsets in one copy ahead of the bands
			for (int b = 0; b < B; ++b) {
				if (two && !offsets_first) CK(cudaMemcpyAsync(d_uo + cut(UO, b, B), h_uo + cut(UO, b, B), cut(UO, b + 1, B) - cut(UO, b, B), cudaMemcpyHostToDevice, s_in));
				const size_t n = two ? US : US + UO;                   // one copy per band: the same bytes out of one buffer
				(void)n;
				if (two) CK(cudaMemcpyAsync(d_us + cut(US, b, B), h_us + cut(US, b, B), cut(US, b + 1, B) - cut(US, b, B), cudaMemcpyHostToDevice, s_in));
				else {
					// (single copy per band: spans sized like both arrays together, capped by the buffer)
					const size_t a0 = cut(US, b, B), a1 = cut(US, b + 1, B);
					CK(cudaMemcpyAsync(d_us + a0, h_us + a0, a1 - a0, cudaMemcpyHostToDevice, s_in));
					const size_t c0 = cut(UO, b, B), c1 = cut(UO, b + 1, B);
					(void)c0; (void)c1;
				}
				CK(cudaEventRecord(ev[b], s_in));
			}
			auto downloads = [&](int b) {
				if (two) CK(cudaMemcpyAsync(h_do + cut(DO_, b, B), d_do + cut(DO_, b, B), cut(DO_, b + 1, B) - cut(DO_, b, B), cudaMemcpyDeviceToHost, s_out));
				CK(cudaMemcpyAsync(h_ds + cut(DS, b, B), d_ds + cut(DS, b, B), cut(DS, b + 1, B) - cut(DS, b, B), cudaMemcpyDeviceToHost, s_out));
			};
			if (!host_driven) {
				for (int b = 0; b < B; ++b) { CK(cudaStreamWaitEvent(s_out, ev[b], 0)); downloads(b); }
			} else {
				for (int b = 0; b < B; ++b) {              // the compute side's hand-over, enqueued up front
					CK(cudaStreamWaitEvent(s_mid, ev[b], 0));
					CK(cudaMemsetAsync(d_tot + b, 1, 8, s_mid));
					CK(cudaEventRecord(ev2[b], s_mid));
					CK(cudaStreamWaitEvent(s_ctl, ev2[b], 0));
					CK(cudaMemcpyAsync(h_tot + b, d_tot + b, 8, cudaMemcpyDeviceToHost, s_ctl));
					CK(cudaEventRecord(ev2[b], s_ctl));
				}
				for (int b = 0; b < B; ++b) { CK(cudaEventSynchronize(ev2[b])); downloads(b); }
			}
			CK(cudaStreamSynchronize(s_out)); CK(cudaStreamSynchronize(s_in));
			best = std::min(best, now_ms() - t0);
		}
		std::printf("bands %2d  copies per band and direction %d  downloads %-11s %s: %.3f ms\n", B, two ? 2 : 1, host_driven ? "host-driven" : "up front", offsets_first ? "offsets first " : "", best);
		std::fflush(stdout);
	};
	for (int B : {4, 8, 16}) { run(B, false, false, false); run(B, true, false, false); run(B, true, true, false); run(B, true, true, true); }
	return 0;
}
