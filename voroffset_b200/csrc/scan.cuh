// Exclusive prefix sum of per-list interval counts -> CSR offsets ("prefix-sum compaction").
// Three launches: per-tile reduce (u64 sums), single-block scan of the tile sums, per-tile scan + add.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vo {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;                       // per thread
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;  // 2048 counts per block

__device__ __forceinline__ unsigned long long warp_incl_scan(unsigned long long v)
{
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		unsigned long long t = __shfl_up_sync(0xffffffffu, v, d);
		if ((threadIdx.x & 31) >= d) v += t;
	}
	return v;
}

// Block-wide exclusive scan of one value per thread; returns the exclusive prefix, *total = block sum.
__device__ __forceinline__ unsigned long long block_excl_scan(unsigned long long v, unsigned long long *total)
{
	__shared__ unsigned long long wsum[32];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	unsigned long long inc = warp_incl_scan(v);
	if (lane == 31) wsum[wid] = inc;
	__syncthreads();
	if (wid == 0) {
		unsigned long long w = (lane < (int)(blockDim.x >> 5)) ? wsum[lane] : 0ull;
		unsigned long long wi = warp_incl_scan(w);
		wsum[lane] = wi - w;                       // exclusive prefix of each warp
	}
	__syncthreads();
	unsigned long long excl = wsum[wid] + inc - v;
	__shared__ unsigned long long tot_s;
	if (threadIdx.x == blockDim.x - 1) tot_s = excl + v;   // last thread: prefix + own value = block sum
	__syncthreads();
	*total = tot_s;
	__syncthreads();                                        // shared scratch may be reused by the next call
	return excl;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const uint32_t *__restrict__ cnt, uint64_t n,
                                                              unsigned long long *__restrict__ tile_sum)
{
	const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE;
	unsigned long long s = 0;
#pragma unroll
	for (int i = 0; i < SCAN_ITEMS; ++i) {
		uint64_t k = base + (uint64_t)i * SCAN_THREADS + threadIdx.x;
		if (k < n) s += cnt[k];
	}
	unsigned long long tot;
	block_excl_scan(s, &tot);
	if (threadIdx.x == 0) tile_sum[blockIdx.x] = tot;
}

// One block: in-place exclusive scan of the tile sums; tile_sum[ntiles] receives the grand total.
// base_in (optional): the scan starts from *base_in instead of 0 (offsets of a band of a larger volume);
// base_out (optional) receives the grand total as well.
__global__ void __launch_bounds__(1024) k_scan_tiles(unsigned long long *__restrict__ tile_sum, uint32_t ntiles,
                                                     const unsigned long long *base_in = nullptr, unsigned long long *base_out = nullptr)
{
	unsigned long long carry = base_in ? *base_in : 0ull;
	for (uint32_t b = 0; b < ntiles; b += blockDim.x) {
		uint32_t k = b + threadIdx.x;
		unsigned long long v = (k < ntiles) ? tile_sum[k] : 0ull;
		unsigned long long tot;
		unsigned long long ex = block_excl_scan(v, &tot);
		if (k < ntiles) tile_sum[k] = carry + ex;
		carry += tot;
	}
	if (threadIdx.x == 0) { tile_sum[ntiles] = carry; if (base_out) *base_out = carry; }
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const uint32_t *__restrict__ cnt, uint64_t n,
                                                             const unsigned long long *__restrict__ tile_sum,
                                                             uint32_t *__restrict__ off)
{
	const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
	uint32_t c[SCAN_ITEMS];
	unsigned long long s = 0;
#pragma unroll
	for (int i = 0; i < SCAN_ITEMS; ++i) {
		uint64_t k = base + i;
		c[i] = (k < n) ? cnt[k] : 0u;
		s += c[i];
	}
	unsigned long long tot;
	unsigned long long ex = block_excl_scan(s, &tot) + tile_sum[blockIdx.x];
#pragma unroll
	for (int i = 0; i < SCAN_ITEMS; ++i) {
		uint64_t k = base + i;
		if (k < n) off[k] = (uint32_t)ex;
		ex += c[i];
	}
	if (blockIdx.x == gridDim.x - 1 && threadIdx.x == blockDim.x - 1) off[n] = (uint32_t)tile_sum[gridDim.x];
}

// ---------------------------------------------------------------------------------------------------
// Single-pass form ("decoupled look-back"): ONE launch turns the staged lists into canonical CSR - the exclusive
// prefix sum of the counts AND the compaction of the spans - instead of reduce / scan / apply / compact with a host
// round trip in between (the spans buffer is sized from the previous result; a result that outgrows it falls back to
// k_compact with the offsets this kernel has already written).
// Tiles take tickets in launch order, publish their own sum at once and then look back over their predecessors'
// state words until they meet one that carries an inclusive prefix. A state word holds value, call epoch and flag
// together, so one 64-bit load / store is the whole protocol and the array never has to be cleared between calls.
// ---------------------------------------------------------------------------------------------------
constexpr unsigned long long SCAN_AGG = 1ull, SCAN_INCL = 2ull;
constexpr int SCAN_EPOCH_BITS = 22;
__device__ __forceinline__ unsigned long long scan_word(unsigned long long value, uint32_t epoch, unsigned long long flag)
{
	return (value << (SCAN_EPOCH_BITS + 2)) | ((unsigned long long)epoch << 2) | flag;
}

// Prefix of everything before `tile` (warp 0 of the block calls this; every lane returns the same value).
__device__ __forceinline__ unsigned long long scan_look_back(const volatile unsigned long long *state, uint32_t tile, uint32_t epoch)
{
	const int lane = threadIdx.x & 31;
	unsigned long long excl = 0;
	long long idx = (long long)tile - 1 - lane;
	for (;;) {
		const unsigned long long w = idx >= 0 ? state[idx] : scan_word(0ull, epoch, SCAN_INCL);
		const bool ok = (uint32_t)((w >> 2) & ((1u << SCAN_EPOCH_BITS) - 1u)) == epoch && (w & 3ull) != 0ull;
		if (!__all_sync(0xffffffffu, ok)) continue;          // a predecessor has not published yet
		const unsigned int incl = __ballot_sync(0xffffffffu, (w & 3ull) == SCAN_INCL);
		const int first = incl ? __ffs(incl) - 1 : 31;       // lanes 0 .. first contribute
		unsigned long long v = lane <= first ? (w >> (SCAN_EPOCH_BITS + 2)) : 0ull;
#pragma unroll
		for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
		excl += v;
		if (incl) return excl;
		idx -= 32;
	}
}

} // namespace vo
