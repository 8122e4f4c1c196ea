// Minimal stand-in for <tbb/parallel_for.h>, written for this repo's test oracle only.
//
// The reference guards its multithreaded regions with `#ifdef USE_TBB` but its build never defines
// it and fetches TBB at configure time. Compiling the reference's own sources with
// -DUSE_TBB against this shim turns those dormant regions on without touching a line of them:
// parallel_for hands out [begin,end) in grain-sized chunks to std::threads (dynamic schedule, one
// body invocation per chunk, like tbb's simple partitioning of a blocked_range).
#pragma once
#include <atomic>
#include <exception>
#include <mutex>
#include <thread>
#include <vector>
#include "tbb/blocked_range.h"

namespace tbb {

inline int &shim_num_threads() { static int n = 1; return n; }

template <typename T, typename Body>
void parallel_for(const blocked_range<T> &range, const Body &body)
{
	const T b = range.begin(), e = range.end();
	if (!(b < e)) return;
	const T grain = (T)range.grainsize();
	int nt = shim_num_threads();
	const unsigned long long nchunks = ((unsigned long long)(e - b) + grain - 1) / grain;
	if ((unsigned long long)nt > nchunks) nt = (int)nchunks;
	if (nt <= 1) {
		for (T s = b; s < e; s += grain) body(blocked_range<T>(s, (e - s < grain) ? e : (T)(s + grain), grain));
		return;
	}
	std::atomic<unsigned long long> next(0);
	std::exception_ptr err;
	std::mutex err_mu;
	auto worker = [&]() {
		try {
			for (;;) {
				unsigned long long c = next.fetch_add(1);
				if (c >= nchunks) break;
				T s = (T)(b + c * grain);
				T t = (e - s < grain) ? e : (T)(s + grain);
				body(blocked_range<T>(s, t, grain));
			}
		} catch (...) {
			std::lock_guard<std::mutex> lk(err_mu);
			if (!err) err = std::current_exception();
		}
	};
	std::vector<std::thread> pool;
	for (int i = 1; i < nt; ++i) pool.emplace_back(worker);
	worker();
	for (auto &t : pool) t.join();
	if (err) std::rethrow_exception(err);
}

} // namespace tbb
