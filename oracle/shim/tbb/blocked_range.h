// Minimal stand-in for <tbb/blocked_range.h> (test oracle only; see parallel_for.h).
#pragma once
#include <cstddef>
namespace tbb {
template <typename T> class blocked_range {
	T b_, e_; std::size_t g_;
public:
	blocked_range(T b, T e, std::size_t grain = 1) : b_(b), e_(e), g_(grain ? grain : 1) {}
	T begin() const { return b_; }
	T end() const { return e_; }
	std::size_t grainsize() const { return g_; }
};
} // namespace tbb
