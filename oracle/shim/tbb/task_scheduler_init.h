// Minimal stand-in for the (removed) <tbb/task_scheduler_init.h> API (test oracle only).
#pragma once
#include "tbb/parallel_for.h"
namespace tbb {
class task_scheduler_init {
public:
	explicit task_scheduler_init(int n = -1) { if (n > 0) shim_num_threads() = n; }
};
} // namespace tbb
