// Development aid, compiled in with -DVO_KTRACE only (build.py --out build/lib_ktrace.so -DVO_KTRACE; never in the
// shipped library): every CTA (the tile kernel: every warp) of the kernels on the dilation path records
// {kernel id, SM, aux, block, start, end} (globaltimer, ns) into a device buffer -
// vo_set_option("ktrace", "<records>") allocates it, vo_set_option("ktrace_dump", "<path>") writes it as CSV and
// clears it. scripts/ktrace_view.py turns the CSV into a per-kernel / per-SM timeline: which launch holds which SMs
// when, and where SMs sit idle between the stream-ordered launches of the banded host-buffer call.
#pragma once
#include <cuda_runtime.h>

namespace vo {

enum KtId { KT_THRESH = 1, KT_ORDER_COUNT, KT_ORDER_PLACE, KT_TILE, KT_TILE_LIST, KT_PASS1, KT_PASS2_ROWS, KT_PASS2, KT_SCAN_COMPACT, KT_MARK, KT_COPY_OUT };

#ifdef VO_KTRACE
struct KTraceBuf { unsigned long long *rec; unsigned int *count; unsigned int cap; };
__constant__ KTraceBuf c_kt;

__device__ __forceinline__ unsigned long long kt_now()
{
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}
struct KtScope {
	unsigned int slot;
	__device__ __forceinline__ KtScope(int id, unsigned int aux, bool leader)
	{
		slot = 0xffffffffu;
		if (leader && c_kt.rec) {
			const unsigned int s = atomicAdd(c_kt.count, 1u);
			if (s < c_kt.cap) {
				unsigned int smid;
				asm("mov.u32 %0, %%smid;" : "=r"(smid));
				slot = s;
				c_kt.rec[4ull * s] = (unsigned long long)id | ((unsigned long long)smid << 32);
				c_kt.rec[4ull * s + 1] = (unsigned long long)aux | ((unsigned long long)blockIdx.x << 32);
				c_kt.rec[4ull * s + 3] = 0;
				c_kt.rec[4ull * s + 2] = kt_now();
			}
		}
	}
	__device__ __forceinline__ ~KtScope()
	{
		if (slot != 0xffffffffu) c_kt.rec[4ull * slot + 3] = kt_now();
	}
};
#define KT_SCOPE(id, aux, leader) KtScope kt_scope_((id), (unsigned int)(aux), (leader))
// a time stamp on a stream (copies: after a band's upload, around a band's download)
__global__ void k_kt_mark(unsigned int aux) { KT_SCOPE(KT_MARK, aux, true); }
#define KT_MARK_STREAM(aux, stream) k_kt_mark<<<1, 1, 0, (stream)>>>((unsigned int)(aux))
#else
#define KT_SCOPE(id, aux, leader)
#define KT_MARK_STREAM(aux, stream)
#endif

} // namespace vo
