// Multi-GPU form of the 3D operators behind the C ABI (vo_mg_* in include/voroffset_b200.h): the grid is cut into
// y-slabs, one per GPU, and the floor(R) boundary rows of the INPUT travel to the two neighbours with NCCL
// (ncclSend / ncclRecv over NVLink, one group per step) while pass 1 of the rows that need no halo is already running.
// This is the reference's dormant TBB decomposition (VoronoiVorPower.cpp:41-63,70-92: tasks own disjoint slices)
// stretched across devices; erosion keeps the reference's border / complement handling per slab (Voronoi.cpp:8-89:
// the one-line solid border belongs to the two edge slabs, the complement kernels are column-local).
//
// Included at the end of vo_lib.cu (same translation unit: it uses the internals of the single-GPU path).
//
// Two ways to build a group:
//   vo_mg_create(device_ids, n)               one process drives n GPUs (ncclCommInitAll), one host thread per GPU while
//                                             an operator runs: what offset3d --gpus N and VoronoiMorphoB200 use;
//   vo_mg_unique_id + vo_mg_create_rank(...)  one process per GPU (torchrun): what bench.py uses.
// NCCL is loaded at run time (dlopen "libnccl.so.2": inside a PyTorch process that is the copy torch has already
// loaded, elsewhere the system's), so the single-GPU library has no link-time dependency on it.
//
// Halo protocol, per neighbour link and step (both sides derive every decision from numbers both have seen):
//   message 1  (J nx + 1) offsets rebased to 0 + [interval count, overflow flag]           fixed size
//              FAST links add the spans, padded to the capacity agreed so far, in the same NCCL group
//   message 2  the exact spans - only on a link's first step for a grid width / radius (no capacity yet) or when a halo
//              outgrew its capacity (flag set by the sender, seen by the receiver in message 1)
//   afterwards both sides raise the link's capacities to grow(count) of what was just transmitted.
// The compute path is a local decision: with every link FAST the overlapped slab step (vo_slab_begin / finish in
// vo_lib.cu) runs pass 1 of the interior rows during message 1; otherwise the rank concatenates [prev | own | next]
// and runs the plain passes on the rows it owns. Same kernels and tables either way: the rows are bit-identical to
// the single-GPU result.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <map>
#include <thread>

namespace {

struct NcclApi {
	void *h = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	const char *(*GetErrorString)(ncclResult_t) = nullptr;
	ncclResult_t (*GetVersion)(int *) = nullptr;
	std::string err;
};

NcclApi *nccl_api()
{
	static NcclApi api;
	static std::once_flag once;
	std::call_once(once, [] {
		const char *names[] = {std::getenv("VO_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
		for (const char *n : names) {
			if (!n || !*n) continue;
			api.h = dlopen(n, RTLD_NOW | RTLD_LOCAL);
			if (api.h) break;
		}
		if (!api.h) { api.err = "libnccl.so.2 not found (set VO_NCCL_LIB)"; return; }
		auto sym = [&](const char *n) { void *p = dlsym(api.h, n); if (!p && api.err.empty()) api.err = std::string("NCCL symbol missing: ") + n; return p; };
		api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
		api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
		api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
		api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
		api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
		api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
		api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
		api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
		api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
		api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
		api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
	});
	return api.err.empty() ? &api : nullptr;
}

constexpr int MG_HDR = 2;                                   // [interval count, overflow flag] behind the offsets of a halo message

inline uint64_t mg_grow(uint64_t n) { return std::max<uint64_t>(1024, n + n / 2 + 16); }   // capacity both sides derive from a transmitted count

struct MgCaps { uint64_t out = 0, in = 0; };

struct MgLink {                                             // one neighbour
	int peer = -1;
	std::map<std::pair<int, int>, MgCaps> caps;             // (nx, floor(R)) -> capacities agreed so far (absent: first step)
	uint32_t *off_out = nullptr, *off_in = nullptr;         // message 1: offsets + header
	double2 *sp_out = nullptr, *sp_in = nullptr;            // spans (sp_out only for FAST links: exact sends read the volume itself)
	size_t off_cap = 0, sp_out_cap = 0, sp_in_cap = 0;      // allocated entries
};

struct MgStats {
	double halo_ms = 0;          // device time of the NCCL groups (events on the communication stream)
	double halo_wait_ms = 0;     // host time blocked until the halos had landed (beyond enqueueing the overlapped work)
	uint64_t halo_bytes = 0;     // bytes sent
	int messages = 0;            // NCCL groups
	int overlapped = 0;          // primitives that took the overlapped slab step
	int plain = 0;               // ... the concatenate-then-dilate path
};

struct MgRank {
	vo_ctx *ctx = nullptr;
	ncclComm_t comm = nullptr;
	cudaStream_t cs = nullptr;                              // communication stream
	cudaEvent_t ev_pack = nullptr, ev_c0 = nullptr, ev_c1 = nullptr;
	uint32_t *h_hdr = nullptr;                              // pinned: [link][in count, in flag, out count, out flag], [8] the agreement word
	unsigned int *d_agree = nullptr;                        // device word of mg_all_agree
	int rank = 0, world = 1;
	MgLink prev, next;
	MgStats stats;
};

} // namespace

struct vo_mg {
	std::vector<MgRank> ranks;                              // local ranks (all of them for a single-process group)
	int world = 1;
	bool single_process = true;
	std::string err;
};

namespace {

int mg_fail(MgRank &r, int code, const std::string &msg) { return fail(r.ctx, code, msg); }

#define VO_NCCL(r, call)                                                                                \
	do {                                                                                                \
		ncclResult_t e_ = (call);                                                                       \
		if (e_ != ncclSuccess) return mg_fail(r, VO_ERR_CUDA, std::string(#call) + ": " + nccl_api()->GetErrorString(e_)); \
	} while (0)

template <typename T> int mg_ensure(MgRank &r, T **p, size_t *cap, size_t need)
{
	if (*cap >= need && *p) return VO_OK;
	vo_ctx *ctx = r.ctx;
	if (*p) { cudaStreamSynchronize(r.cs); cudaStreamSynchronize(ctx->stream); cudaFree(*p); *p = nullptr; *cap = 0; }
	const size_t n = std::max<size_t>(need + need / 4, 1024);
	VO_CUDA(cudaMalloc((void **)p, n * sizeof(T)));
	*cap = n;
	return VO_OK;
}

void mg_free_link(MgLink &l)
{
	cudaFree(l.off_out); cudaFree(l.off_in); cudaFree(l.sp_out); cudaFree(l.sp_in);
	l = MgLink();
}

// One dilation of this rank's slab. world_act: ranks [0, world_act) take part (a grid with fewer rows than ranks x halo
// leaves the last ranks idle). clip_lo / clip_hi / unpruned: see erode_with.
// dual (mg_erode_dual): not a dilation but the erosion of `own` in dual form (vo_lib.cu: erode_dual) - the same halo of
// input rows, the mirrored intervals through pass 1, k_pass2_rows_dual on the rows this rank owns.
struct MgDual { double zmin, zmax; };

int mg_dilate(MgRank &r, int world_act, int method, const vo_dvol *own, double R, double clip_lo, double clip_hi, bool unpruned,
              vo_dvol **out, PassTimes *pt, const MgDual *dual = nullptr)
{
	vo_ctx *ctx = r.ctx;
	VO_TRY(check_radius(ctx, R));
	const int J = (int)std::floor(R), nx = own->nx, ny = own->ny;
	const bool has_prev = r.rank > 0, has_next = r.rank + 1 < world_act;
	auto local = [&](const vo_dvol *v, vo_dvol **o) {
		if (dual) {                                         // (the group has agreed that every slab qualifies)
			const int rc = erode_dual(ctx, v, dual->zmin, dual->zmax, R, o, pt);
			return rc == DUAL_NA ? fail(ctx, VO_ERR_OVERFLOW, "dual erosion: a slab did not qualify after all") : rc;
		}
		const bool saved = ctx->force_simple_pass1;
		if (unpruned) ctx->force_simple_pass1 = true;
		const int rc = dilate(ctx, method, v, R, o, pt, clip_lo, clip_hi);
		ctx->force_simple_pass1 = saved;
		return rc;
	};
	if (world_act <= 1 || J == 0 || (!has_prev && !has_next)) return local(own, out);
	NcclApi *nc = nccl_api();                               // (loaded on first use: a group of one never needs it)
	if (!nc) return mg_fail(r, VO_ERR_CUDA, "NCCL is not available");
	if (ny < J) return mg_fail(r, VO_ERR_ARG, "a slab has fewer rows than the halo (floor(radius)): use fewer GPUs");
	const size_t L = (size_t)J * nx + 1;
	const std::pair<int, int> key(nx, J);
	MgLink *links[2] = {has_prev ? &r.prev : nullptr, has_next ? &r.next : nullptr};
	bool fast[2] = {false, false}, all_fast = true;
	MgCaps caps[2];
	cudaStream_t sm = ctx->stream;
	for (int i = 0; i < 2; ++i) {
		MgLink *lk = links[i];
		if (!lk) continue;
		auto it = lk->caps.find(key);
		fast[i] = it != lk->caps.end();
		if (fast[i]) caps[i] = it->second;
		all_fast = all_fast && fast[i];
		if (lk->off_cap < L + MG_HDR) {                     // (both message buffers share one size)
			size_t c0 = lk->off_cap, c1 = lk->off_cap;
			VO_TRY(mg_ensure(r, &lk->off_out, &c0, L + MG_HDR));
			VO_TRY(mg_ensure(r, &lk->off_in, &c1, L + MG_HDR));
			lk->off_cap = std::min(c0, c1);
		}
		if (fast[i]) {
			VO_TRY(mg_ensure(r, &lk->sp_out, &lk->sp_out_cap, caps[i].out));
			VO_TRY(mg_ensure(r, &lk->sp_in, &lk->sp_in_cap, caps[i].in));
		}
		// outgoing boundary rows -> message buffers (offsets rebased, header; the spans of a FAST link when they fit)
		const unsigned long long c0 = i == 0 ? 0ull : (unsigned long long)(ny - J) * nx, c1 = i == 0 ? (unsigned long long)J * nx : (unsigned long long)ny * nx;
		k_halo_pack<<<64, 256, 0, sm>>>(own->off, own->spans, c0, c1, lk->off_out, lk->sp_out, fast[i] ? caps[i].out : 0ull);
		ctx->launches++;
	}
	VO_CUDA(cudaEventRecord(r.ev_pack, sm));
	VO_CUDA(cudaStreamWaitEvent(r.cs, r.ev_pack, 0));

	// overlapped slab step: pass 1 of the rows that need no halo runs while message 1 travels
	vo_slab *S = nullptr;
	if (all_fast && method == VO_METHOD_OURS && !unpruned && !ctx->force_simple_pass1) {
		const int rc = slab_begin(ctx, own, R, has_prev, has_next, has_prev ? caps[0].in : 0, has_next ? caps[1].in : 0,
		                          HaloOut{nullptr, nullptr, 0}, HaloOut{nullptr, nullptr, 0}, nullptr, &S, clip_lo, clip_hi,
		                          dual != nullptr, dual ? dual->zmin : 0.0, dual ? dual->zmax : 0.0);
		if (rc != VO_OK && rc != VO_ERR_ARG) return rc;     // (VO_ERR_ARG: not a case for the overlapped path)
		if (rc != VO_OK) { S = nullptr; ctx->err.clear(); }
	}
	struct SlabGuard { vo_ctx *c; vo_slab *&s; ~SlabGuard() { if (s) { cudaStreamSynchronize(c->stream); if (c->s_in) cudaStreamSynchronize(c->s_in); free_slab(s); } } } slab_guard{ctx, S};

	// ---- message 1 ----
	VO_CUDA(cudaEventRecord(r.ev_c0, r.cs));
	VO_NCCL(r, nc->GroupStart());
	for (int i = 0; i < 2; ++i) {
		MgLink *lk = links[i];
		if (!lk) continue;
		VO_NCCL(r, nc->Send(lk->off_out, (L + MG_HDR) * sizeof(uint32_t), ncclChar, lk->peer, r.comm, r.cs));
		VO_NCCL(r, nc->Recv(lk->off_in, (L + MG_HDR) * sizeof(uint32_t), ncclChar, lk->peer, r.comm, r.cs));
		r.stats.halo_bytes += (L + MG_HDR) * sizeof(uint32_t);
		if (fast[i]) {
			VO_NCCL(r, nc->Send(lk->sp_out, caps[i].out * sizeof(double2), ncclChar, lk->peer, r.comm, r.cs));
			VO_NCCL(r, nc->Recv(lk->sp_in, caps[i].in * sizeof(double2), ncclChar, lk->peer, r.comm, r.cs));
			r.stats.halo_bytes += caps[i].out * sizeof(double2);
		}
	}
	VO_NCCL(r, nc->GroupEnd());
	r.stats.messages++;
	for (int i = 0; i < 2; ++i) {
		MgLink *lk = links[i];
		if (!lk) continue;
		VO_CUDA(cudaMemcpyAsync(r.h_hdr + 4 * i, lk->off_in + L, MG_HDR * sizeof(uint32_t), cudaMemcpyDeviceToHost, r.cs));
		VO_CUDA(cudaMemcpyAsync(r.h_hdr + 4 * i + 2, lk->off_out + L, MG_HDR * sizeof(uint32_t), cudaMemcpyDeviceToHost, r.cs));
	}
	VO_CUDA(cudaEventRecord(r.ev_c1, r.cs));
	{
		const auto t0 = std::chrono::steady_clock::now();
		VO_CUDA(cudaStreamSynchronize(r.cs));
		r.stats.halo_wait_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
		float t = 0;
		if (cudaEventElapsedTime(&t, r.ev_c0, r.ev_c1) == cudaSuccess) r.stats.halo_ms += t; else cudaGetLastError();
	}
	uint64_t n_in[2] = {0, 0}, n_out[2] = {0, 0};
	bool follow_in[2] = {false, false}, follow_out[2] = {false, false}, any_follow = false, overflow = false;
	for (int i = 0; i < 2; ++i) {
		if (!links[i]) continue;
		n_in[i] = r.h_hdr[4 * i]; n_out[i] = r.h_hdr[4 * i + 2];
		// (k_halo_pack raises the flag when count > capacity; a link without capacities always needs message 2)
		follow_in[i] = (!fast[i] || r.h_hdr[4 * i + 1] != 0) && n_in[i] > 0;
		follow_out[i] = (!fast[i] || r.h_hdr[4 * i + 3] != 0) && n_out[i] > 0;
		overflow = overflow || (fast[i] && (r.h_hdr[4 * i + 1] != 0 || r.h_hdr[4 * i + 3] != 0));
		any_follow = any_follow || follow_in[i] || follow_out[i];
	}
	// ---- message 2: exact spans where message 1 could not carry them ----
	if (any_follow) {
		for (int i = 0; i < 2; ++i)
			if (links[i] && follow_in[i]) VO_TRY(mg_ensure(r, &links[i]->sp_in, &links[i]->sp_in_cap, n_in[i]));
		VO_CUDA(cudaEventRecord(r.ev_c0, r.cs));
		VO_NCCL(r, nc->GroupStart());
		for (int i = 0; i < 2; ++i) {
			MgLink *lk = links[i];
			if (!lk) continue;
			if (follow_out[i]) {                             // the boundary rows are one contiguous range of the volume's spans
				const double2 *src = own->spans + (i == 0 ? 0ull : own->nspans - n_out[i]);
				VO_NCCL(r, nc->Send(src, n_out[i] * sizeof(double2), ncclChar, lk->peer, r.comm, r.cs));
				r.stats.halo_bytes += n_out[i] * sizeof(double2);
			}
			if (follow_in[i]) VO_NCCL(r, nc->Recv(lk->sp_in, n_in[i] * sizeof(double2), ncclChar, lk->peer, r.comm, r.cs));
		}
		VO_NCCL(r, nc->GroupEnd());
		r.stats.messages++;
		VO_CUDA(cudaEventRecord(r.ev_c1, r.cs));
		const auto t0 = std::chrono::steady_clock::now();
		VO_CUDA(cudaStreamSynchronize(r.cs));
		r.stats.halo_wait_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
		float t = 0;
		if (cudaEventElapsedTime(&t, r.ev_c0, r.ev_c1) == cudaSuccess) r.stats.halo_ms += t; else cudaGetLastError();
	}
	for (int i = 0; i < 2; ++i) {
		if (!links[i]) continue;
		MgCaps &c = links[i]->caps[key];
		c.in = std::max(c.in, mg_grow(n_in[i]));
		c.out = std::max(c.out, mg_grow(n_out[i]));
	}

	// ---- compute ----
	if (S && !overflow) {
		vo_slab *s = S;
		S = nullptr;                                        // slab_finish's caller releases it
		double ms1 = 0, ms2 = 0;
		const int rc = slab_finish(s, has_prev ? r.prev.off_in : nullptr, has_prev ? r.prev.sp_in : nullptr, n_in[0],
		                           has_next ? r.next.off_in : nullptr, has_next ? r.next.sp_in : nullptr, n_in[1], out, &ms1, &ms2);
		if (rc != VO_OK) { cudaStreamSynchronize(sm); if (ctx->s_in) cudaStreamSynchronize(ctx->s_in); }
		free_slab(s);
		if (rc == VO_OK) { if (pt) { pt->ms1 = ms1; pt->ms2 = ms2; } r.stats.overlapped++; return VO_OK; }
		if (rc != VO_ERR_OVERFLOW) return rc;               // (a pool that was too small: the plain path below regrows it)
		ctx->err.clear();
	}
	// plain path: [prev halo | own rows | next halo] -> the passes on the rows this rank owns
	vo_dvol hp, hn;
	if (has_prev) { hp.nx = nx; hp.ny = J; hp.nspans = n_in[0]; hp.off = r.prev.off_in; hp.spans = r.prev.sp_in; }
	if (has_next) { hn.nx = nx; hn.ny = J; hn.nspans = n_in[1]; hn.off = r.next.off_in; hn.spans = r.next.sp_in; }
	vo_dvol *ext = nullptr, *full = nullptr;
	VO_TRY(concat_rows(ctx, has_prev ? &hp : nullptr, own, has_next ? &hn : nullptr, &ext));
	const int y0 = has_prev ? J : 0;
	int rc;
	if (dual) {
		rc = erode_dual(ctx, ext, dual->zmin, dual->zmax, R, out, pt, y0, y0 + ny);
		if (rc == DUAL_NA) rc = fail(ctx, VO_ERR_OVERFLOW, "dual erosion: a slab did not qualify after all");
	} else if (method == VO_METHOD_OURS) {
		float t1 = 0, t2 = 0;
		// (a list beyond the redo capacity: once more with the last-resort launches, like dilate() - the ranks do not talk here)
		rc = with_huge_lists(ctx, [&] {
			cudaEventRecord(ctx->ev[0], sm);
			vo_dmid *mid = nullptr;
			const bool saved = ctx->force_simple_pass1;
			if (unpruned) ctx->force_simple_pass1 = true;
			int rc1 = pass1(ctx, ext, R, &mid, clip_lo, clip_hi);
			ctx->force_simple_pass1 = saved;
			if (rc1 == VO_OK) {
				cudaEventRecord(ctx->ev[1], sm);
				rc1 = pass2(ctx, mid, y0, y0 + ny, out);
				vo_dmid_free(ctx, mid);
			}
			return rc1;
		});
		if (rc == VO_OK) {
			cudaEventRecord(ctx->ev[2], sm);
			cudaEventSynchronize(ctx->ev[2]);
			cudaEventElapsedTime(&t1, ctx->ev[0], ctx->ev[1]);
			cudaEventElapsedTime(&t2, ctx->ev[1], ctx->ev[2]);
			if (pt) { pt->ms1 = t1; pt->ms2 = t2; }
		}
	} else {
		rc = local(ext, &full);
		if (rc == VO_OK) {
			rc = vo_dvol_rows(ctx, full, y0, y0 + ny, out);
			free_dvol(ctx, full);
		}
	}
	free_dvol(ctx, ext);
	if (rc == VO_OK) r.stats.plain++;
	return rc;
}

// Does every slab of the group qualify for the dual form of the erosion? Host-side conditions (the tile kernel must be
// the one that runs) and the data (k_dual_check) go into one device word, the group takes the maximum.
int mg_dual_agree(MgRank &r, int world_act, const vo_dvol *own, double zmin, double zmax, double R, bool *all)
{
	vo_ctx *ctx = r.ctx;
	*all = false;
	const int J = (int)std::floor(R);
	const unsigned long long ncols = (unsigned long long)own->nx * own->ny;
	const double k_in = ncols ? (double)own->nspans / (double)ncols : 0.0;
	bool cand = check_radius(ctx, R) == VO_OK && ncols > 0 && own->nspans <= ncols && own->max_cnt <= 1 && own->dual_state != 2 &&
	            !ctx->force_simple_pass1 && TilePlan::fits(J, k_in) &&
	            ((double)ncols * (J + 1) * std::max(1.0, k_in) >= (double)(2ull << 20) || ctx->force_tile_pass1);
	ctx->err.clear();
	if (world_act <= 1 || J == 0) {                         // no neighbour to agree with: erode_dual finds out by itself
		*all = cand;
		return VO_OK;
	}
	if (world_act != r.world) return VO_OK;                 // (idle ranks do not take part in a collective: the general path)
	NcclApi *nc = nccl_api();
	if (!nc) return mg_fail(r, VO_ERR_CUDA, "NCCL is not available");
	if (!r.d_agree) VO_CUDA(cudaMalloc((void **)&r.d_agree, sizeof(unsigned int)));
	cudaStream_t sm = ctx->stream;
	VO_CUDA(cudaMemsetAsync(r.d_agree, 0, sizeof(unsigned int), sm));
	k_dual_check<<<blocks_for(std::max<unsigned long long>(ncols, 1), 256), 256, 0, sm>>>(own->off, own->spans, cand ? ncols : 0ull, zmin - 1, zmax + 1,
	                                                                                   cand ? 0u : 1u, r.d_agree);
	ctx->launches++;
	VO_CUDA(cudaEventRecord(r.ev_pack, sm));
	VO_CUDA(cudaStreamWaitEvent(r.cs, r.ev_pack, 0));
	VO_NCCL(r, nc->AllReduce(r.d_agree, r.d_agree, 1, ncclUint32, ncclMax, r.comm, r.cs));
	VO_CUDA(cudaMemcpyAsync(r.h_hdr + 8, r.d_agree, sizeof(unsigned int), cudaMemcpyDeviceToHost, r.cs));
	const auto t0 = std::chrono::steady_clock::now();
	VO_CUDA(cudaStreamSynchronize(r.cs));
	r.stats.halo_wait_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
	r.stats.messages++;
	*all = r.h_hdr[8] == 0;
	return VO_OK;
}

// One operator on this rank's slab (the -x switch of offset3d.cpp:116-136 per slab).
int mg_morph(MgRank &r, int world_act, int op, int method, const vo_dvol *own, double zmin, double zmax, double R, vo_dvol **out, PassTimes *pt)
{
	vo_ctx *ctx = r.ctx;
	const double ninf = -std::numeric_limits<double>::infinity(), pinf = std::numeric_limits<double>::infinity();
	auto dil = [&](const vo_dvol *v, vo_dvol **o) { return mg_dilate(r, world_act, method, v, R, ninf, pinf, false, o, pt); };
	auto ero = [&](const vo_dvol *v, vo_dvol **o) {
		// Dual form (vo_lib.cu: erode_dual) when EVERY slab qualifies: the ranks exchange input rows instead of
		// complement rows then, so they have to agree before anything travels - one word, all-reduced.
		if (method == VO_METHOD_OURS && ctx->erosion_mode != 2) {
			bool all = false;
			VO_TRY(mg_dual_agree(r, world_act, v, zmin, zmax, R, &all));
			if (all) {
				const MgDual md{zmin, zmax};
				const int rc = mg_dilate(r, world_act, method, v, R, ninf, pinf, false, o, pt, &md);
				if (rc == VO_OK) ctx->dual_erosions++;
				return rc;
			}
			if (ctx->erosion_mode == 1) return fail(ctx, VO_ERR_ARG, "erosion = dual: the input does not qualify for the dual form");
		}
		// Voronoi.cpp:18-55: one line of solid border around the GLOBAL grid - the first and the last slab own its rows
		return erode_with(ctx, v, zmin, zmax, r.rank == 0 ? 1 : 0, r.rank == world_act - 1 ? 1 : 0,
			[&](const vo_dvol *neg, double clo, double chi, bool unpruned, vo_dvol **d) { return mg_dilate(r, world_act, method, neg, R, clo, chi, unpruned, d, pt); }, o);
	};
	if (method != VO_METHOD_OURS && method != VO_METHOD_BRUTE_FORCE) return fail(ctx, VO_ERR_ARG, "Invalid method");
	switch (op) {
	case VO_OP_DILATION: return dil(own, out);
	case VO_OP_EROSION: return ero(own, out);
	case VO_OP_OPENING: {
		vo_dvol *tmp = nullptr;
		VO_TRY(ero(own, &tmp));
		const int rc = dil(tmp, out);
		free_dvol(ctx, tmp);
		return rc;
	}
	case VO_OP_CLOSING: {
		vo_dvol *tmp = nullptr;
		VO_TRY(dil(own, &tmp));
		const int rc = ero(tmp, out);
		free_dvol(ctx, tmp);
		return rc;
	}
	default: return fail(ctx, VO_ERR_ARG, "Operation");
	}
}

int mg_init_rank(MgRank &r, int device, int rank, int world)
{
	r.rank = rank; r.world = world;
	r.prev.peer = rank - 1; r.next.peer = rank + 1;
	int rc = vo_create(device, &r.ctx);
	if (rc != VO_OK) return rc;
	DeviceGuard g(device);
	bool ok = cudaStreamCreateWithFlags(&r.cs, cudaStreamNonBlocking) == cudaSuccess;
	ok = ok && cudaEventCreateWithFlags(&r.ev_pack, cudaEventDisableTiming) == cudaSuccess;
	ok = ok && cudaEventCreate(&r.ev_c0) == cudaSuccess && cudaEventCreate(&r.ev_c1) == cudaSuccess;
	ok = ok && cudaMallocHost((void **)&r.h_hdr, 16 * sizeof(uint32_t)) == cudaSuccess;
	if (!ok) { cudaGetLastError(); return VO_ERR_CUDA; }
	return VO_OK;
}

void mg_destroy_rank(MgRank &r)
{
	if (!r.ctx) return;
	DeviceGuard g(r.ctx->device);
	if (r.cs) cudaStreamSynchronize(r.cs);
	cudaStreamSynchronize(r.ctx->stream);
	if (r.comm && nccl_api()) nccl_api()->CommDestroy(r.comm);
	mg_free_link(r.prev); mg_free_link(r.next);
	if (r.h_hdr) cudaFreeHost(r.h_hdr);
	if (r.d_agree) cudaFree(r.d_agree);
	for (cudaEvent_t e : {r.ev_pack, r.ev_c0, r.ev_c1}) if (e) cudaEventDestroy(e);
	if (r.cs) cudaStreamDestroy(r.cs);
	vo_destroy(r.ctx);
	r = MgRank();
}

// fn(local index) on every local rank at once: one host thread per GPU (NCCL's one-thread-per-device model); a single
// local rank runs inline.
template <typename Fn> void mg_parallel(vo_mg *mg, Fn fn)
{
	if (mg->ranks.size() == 1) { fn(0); return; }
	std::vector<std::thread> th;
	for (size_t i = 0; i < mg->ranks.size(); ++i) th.emplace_back([&, i] { fn((int)i); });
	for (auto &t : th) t.join();
}

// rows owned by each of `world` slabs: as even as possible, earlier ranks take the remainder (voroffset_b200/slab.py: slab_bounds)
inline void mg_bounds(int ny, int world, int rank, int *y0, int *y1)
{
	const int base = ny / world, rem = ny % world;
	*y0 = rank * base + std::min(rank, rem);
	*y1 = *y0 + base + (rank < rem ? 1 : 0);
}

} // namespace

extern "C" {

int vo_mg_create(const int *device_ids, int n_dev, vo_mg **out)
{
	if (!out || n_dev < 1 || n_dev > 64) return VO_ERR_ARG;
	*out = nullptr;
	vo_mg *mg = new (std::nothrow) vo_mg();
	if (!mg) return VO_ERR_NOMEM;
	mg->world = n_dev;
	mg->single_process = true;
	mg->ranks.resize(n_dev);
	std::vector<int> devs(n_dev);
	for (int i = 0; i < n_dev; ++i) devs[i] = device_ids ? device_ids[i] : i;
	for (int i = 0; i < n_dev; ++i) {
		const int rc = mg_init_rank(mg->ranks[i], devs[i], i, n_dev);
		if (rc != VO_OK) { vo_mg_destroy(mg); return rc; }
	}
	if (n_dev > 1) {
		NcclApi *nc = nccl_api();
		std::vector<ncclComm_t> comms(n_dev, nullptr);
		if (!nc || nc->CommInitAll(comms.data(), n_dev, devs.data()) != ncclSuccess) { vo_mg_destroy(mg); return VO_ERR_CUDA; }
		for (int i = 0; i < n_dev; ++i) mg->ranks[i].comm = comms[i];
	}
	*out = mg;
	return VO_OK;
}

int vo_mg_unique_id(void *id128)
{
	NcclApi *nc = nccl_api();
	if (!nc || !id128) return nc ? VO_ERR_ARG : VO_ERR_CUDA;
	static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
	ncclUniqueId id;
	if (nc->GetUniqueId(&id) != ncclSuccess) return VO_ERR_CUDA;
	std::memcpy(id128, &id, sizeof(id));
	return VO_OK;
}

int vo_mg_create_rank(int device, int rank, int world, const void *id128, vo_mg **out)
{
	if (!out || world < 1 || rank < 0 || rank >= world || (world > 1 && !id128)) return VO_ERR_ARG;
	*out = nullptr;
	vo_mg *mg = new (std::nothrow) vo_mg();
	if (!mg) return VO_ERR_NOMEM;
	mg->world = world;
	mg->single_process = false;
	mg->ranks.resize(1);
	int rc = mg_init_rank(mg->ranks[0], device, rank, world);
	if (rc == VO_OK && world > 1) {
		NcclApi *nc = nccl_api();
		DeviceGuard g(device);
		ncclUniqueId id;
		std::memcpy(&id, id128, sizeof(id));
		if (!nc || nc->CommInitRank(&mg->ranks[0].comm, world, id, rank) != ncclSuccess) rc = VO_ERR_CUDA;
	}
	if (rc != VO_OK) { vo_mg_destroy(mg); return rc; }
	*out = mg;
	return VO_OK;
}

void vo_mg_destroy(vo_mg *mg)
{
	if (!mg) return;
	for (auto &r : mg->ranks) mg_destroy_rank(r);
	delete mg;
}

int vo_mg_world(const vo_mg *mg) { return mg ? mg->world : 0; }
int vo_mg_local_count(const vo_mg *mg) { return mg ? (int)mg->ranks.size() : 0; }
vo_ctx *vo_mg_ctx(vo_mg *mg, int local) { return (mg && local >= 0 && local < (int)mg->ranks.size()) ? mg->ranks[local].ctx : nullptr; }
int vo_mg_rank(const vo_mg *mg, int local) { return (mg && local >= 0 && local < (int)mg->ranks.size()) ? mg->ranks[local].rank : -1; }
const char *vo_mg_last_error(const vo_mg *mg)
{
	if (!mg) return "no group";
	if (!mg->err.empty()) return mg->err.c_str();
	for (auto &r : mg->ranks) if (r.ctx && !r.ctx->err.empty()) return r.ctx->err.c_str();
	return "";
}

int vo_mg_stats(const vo_mg *mg, int local, double *halo_ms, double *halo_wait_ms, uint64_t *halo_bytes, int *messages, int *overlapped, int *plain)
{
	if (!mg || local < 0 || local >= (int)mg->ranks.size()) return VO_ERR_ARG;
	const MgStats &s = mg->ranks[local].stats;
	if (halo_ms) *halo_ms = s.halo_ms;
	if (halo_wait_ms) *halo_wait_ms = s.halo_wait_ms;
	if (halo_bytes) *halo_bytes = s.halo_bytes;
	if (messages) *messages = s.messages;
	if (overlapped) *overlapped = s.overlapped;
	if (plain) *plain = s.plain;
	return VO_OK;
}

int vo_mg_morph3d_dev(vo_mg *mg, int op, int method, const vo_dvol *const *in, double zmin, double zmax, double radius,
                      vo_dvol **out, double *ms_pass1, double *ms_pass2)
{
	if (!mg || !in || !out) return VO_ERR_ARG;
	mg->err.clear();
	const int nl = (int)mg->ranks.size();
	std::vector<int> rcs(nl, VO_OK);
	std::vector<PassTimes> pts(nl);
	mg_parallel(mg, [&](int i) {
		MgRank &r = mg->ranks[i];
		DeviceGuard g(r.ctx->device);
		r.ctx->err.clear();
		r.stats = MgStats();
		out[i] = nullptr;
		if (!in[i]) { rcs[i] = fail(r.ctx, VO_ERR_ARG, "missing slab"); return; }
		rcs[i] = mg_morph(r, mg->world, op, method, in[i], zmin, zmax, radius, &out[i], &pts[i]);
		if (rcs[i] == VO_OK && cudaStreamSynchronize(r.ctx->stream) != cudaSuccess) rcs[i] = fail(r.ctx, VO_ERR_CUDA, "stream synchronisation failed");
	});
	double t1 = 0, t2 = 0;
	for (int i = 0; i < nl; ++i) {
		if (rcs[i] != VO_OK) return rcs[i];
		t1 = std::max(t1, pts[i].ms1); t2 = std::max(t2, pts[i].ms2);
	}
	if (ms_pass1) *ms_pass1 = t1;
	if (ms_pass2) *ms_pass2 = t2;
	return VO_OK;
}

// Host-buffer drop-in on a single-process group: rows are cut into slabs, uploaded, processed and downloaded per GPU.
int vo_mg_morph3d(vo_mg *mg, int op, int method, int nx, int ny, double zmin, double zmax,
                  const uint32_t *off, const double *spans, double radius,
                  uint32_t **out_off, double **out_spans, uint64_t *out_nspans, double *ms_pass1, double *ms_pass2)
{
	if (!mg || !out_off || !out_spans) return VO_ERR_ARG;
	mg->err.clear();
	if (!mg->single_process) { mg->err = "vo_mg_morph3d needs a single-process group (vo_mg_create): every rank of a multi-process group only has its own rows"; return VO_ERR_ARG; }
	vo_ctx *ctx0 = mg->ranks[0].ctx;
	ctx0->err.clear();
	if (nx < 0 || ny < 0 || !off) return fail(ctx0, VO_ERR_ARG, "bad grid / offsets");
	VO_TRY(check_dims(ctx0, nx, ny));
	const unsigned long long ncols = (unsigned long long)nx * ny;
	if (off[0] != 0) return fail(ctx0, VO_ERR_ARG, "off[0] must be 0");
	if (!offsets_sorted(off, ncols)) return fail(ctx0, VO_ERR_ARG, "offsets must be non-decreasing");
	if (off[ncols] && !spans) return fail(ctx0, VO_ERR_ARG, "spans is NULL");
	if (!(radius >= 0.0)) return fail(ctx0, VO_ERR_ARG, "radius must be in [0, 4096) dexels");
	// every active slab needs at least floor(R) rows (its halo comes from the direct neighbours only)
	const int J = (int)std::floor(std::min(radius, 4096.0));
	const int world_act = std::max(1, std::min(mg->world, ny / std::max(J, 1)));
	const int nl = mg->world;
	std::vector<int> rcs(nl, VO_OK);
	std::vector<PassTimes> pts(nl);
	std::vector<vo_dvol *> res(nl, nullptr);
	mg_parallel(mg, [&](int i) {
		MgRank &r = mg->ranks[i];
		r.stats = MgStats();
		if (i >= world_act) return;
		vo_ctx *ctx = r.ctx;
		DeviceGuard g(ctx->device);
		ctx->err.clear();
		int y0, y1;
		mg_bounds(ny, world_act, i, &y0, &y1);
		const unsigned long long c0 = (unsigned long long)y0 * nx, c1 = (unsigned long long)y1 * nx;
		vo_dvol *own = nullptr;
		rcs[i] = [&]() -> int {
			VO_TRY(new_dvol(ctx, nx, y1 - y0, &own));
			own->nspans = off[c1] - off[c0];
			VO_TRY(dalloc(ctx, &own->spans, own->nspans));
			VO_CUDA(cudaMemcpyAsync(own->off, off + c0, (c1 - c0 + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
			if (own->nspans) VO_CUDA(cudaMemcpyAsync(own->spans, spans + 2 * (size_t)off[c0], own->nspans * sizeof(double2), cudaMemcpyHostToDevice, ctx->stream));
			if (off[c0]) { k_rebase<<<blocks_for(c1 - c0 + 1, 256), 256, 0, ctx->stream>>>(own->off, c1 - c0 + 1, off[c0], 0u); ctx->launches++; }
			VO_TRY(mg_morph(r, world_act, op, method, own, zmin, zmax, radius, &res[i], &pts[i]));
			VO_CUDA(cudaStreamSynchronize(ctx->stream));
			return VO_OK;
		}();
		free_dvol(ctx, own);
	});
	auto cleanup = [&]() { for (int i = 0; i < nl; ++i) if (res[i]) { DeviceGuard g(mg->ranks[i].ctx->device); free_dvol(mg->ranks[i].ctx, res[i]); } };
	for (int i = 0; i < nl; ++i)
		if (rcs[i] != VO_OK) { mg->err = mg->ranks[i].ctx->err; cleanup(); return rcs[i]; }
	uint64_t total = 0;
	for (int i = 0; i < world_act; ++i) total += res[i]->nspans;
	if (total >= (1ull << 32)) { cleanup(); return fail(ctx0, VO_ERR_OVERFLOW, "result has more than 2^32-1 intervals"); }
	uint32_t *ho = (uint32_t *)host_block((ncols + 1) * sizeof(uint32_t));
	double *hs = (double *)host_block(std::max<uint64_t>(total, 1) * sizeof(double2));
	if (!ho || !hs) { vo_free(ho); vo_free(hs); cleanup(); return fail(ctx0, VO_ERR_NOMEM, "pinned host allocation failed"); }
	uint64_t base = 0;
	bool ok = true;
	for (int i = 0; i < world_act; ++i) {                   // downloads of all GPUs in flight together
		vo_ctx *ctx = mg->ranks[i].ctx;
		DeviceGuard g(ctx->device);
		int y0, y1;
		mg_bounds(ny, world_act, i, &y0, &y1);
		const unsigned long long c0 = (unsigned long long)y0 * nx, n = (unsigned long long)(y1 - y0) * nx;
		if (base) { k_rebase<<<blocks_for(n + 1, 256), 256, 0, ctx->stream>>>(res[i]->off, n + 1, 0u, (uint32_t)base); ctx->launches++; }
		ok = ok && cudaMemcpyAsync(ho + c0, res[i]->off, (n + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream) == cudaSuccess;
		if (res[i]->nspans) ok = ok && cudaMemcpyAsync(hs + 2 * base, res[i]->spans, res[i]->nspans * sizeof(double2), cudaMemcpyDeviceToHost, ctx->stream) == cudaSuccess;
		base += res[i]->nspans;
	}
	for (int i = 0; i < world_act; ++i) { DeviceGuard g(mg->ranks[i].ctx->device); ok = ok && cudaStreamSynchronize(mg->ranks[i].ctx->stream) == cudaSuccess; }
	cleanup();
	if (!ok) { cudaGetLastError(); vo_free(ho); vo_free(hs); return fail(ctx0, VO_ERR_CUDA, "download failed"); }
	double t1 = 0, t2 = 0;
	for (int i = 0; i < world_act; ++i) { t1 = std::max(t1, pts[i].ms1); t2 = std::max(t2, pts[i].ms2); }
	*out_off = ho; *out_spans = hs;
	if (out_nspans) *out_nspans = total;
	if (ms_pass1) *ms_pass1 = t1;
	if (ms_pass2) *ms_pass2 = t2;
	return VO_OK;
}

} // extern "C"
