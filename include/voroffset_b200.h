/*
 * voroffset_b200 - C ABI of the B200-native dexel-morphology hot path.
 *
 * This header is the drop-in boundary. Every entry point names the reference interface it replaces
 * (paths relative to the geometryprocessing/voroffset tree). Only plain pointers and sizes cross it;
 * there are no torch / C++ types in the signatures. The library (libvoroffset_b200.so) is CUDA-only:
 * there is no CPU fallback, every compute entry point fails with VO_ERR_CUDA when no sm_100 device is
 * usable.
 *
 * Data layout (same on host and in HBM): a dexel volume with nx*ny columns is a CSR pair
 *   off   : uint32_t[nx*ny + 1]   column (x,y) is list c = x + nx*y  (CompressedVolume.h:28-29)
 *   spans : double  [2 * off[nx*ny]]  (z1,z2) pairs, ascending and disjoint inside a column, in dexel
 *           units (world z / spacing, Dexelize.cpp:204).   vo_span_bytes() == 16.
 * A 2D dexel image (DoubleCompressedImage) is the same with one list per row.
 *
 * Threading: a vo_ctx owns one CUDA device, one stream and its scratch memory; it is not
 * thread-safe. Use one context per calling thread / per GPU.
 *
 * Limits the reference does not have (all reported, never silent):
 *   - CSR offsets are uint32: at most 2^32 - 9 columns and 2^32 - 1 intervals per volume (VO_ERR_OVERFLOW beyond);
 *   - 3D radii must be in [0, 4096) dexels (VO_ERR_ARG); the 2D radius is only bounded by 1e9 rows;
 *   - while a column's result is being gathered its running union may hold at most 32768 disjoint intervals at a time
 *     (VO_ERR_OVERFLOW beyond). Lists of up to 32 intervals take the fast path, up to 512 the redo launch; a dilation that
 *     meets a longer one is repeated once with the redo launches in their last-resort form (csrc/kernels.cuh: CAP_HUGE,
 *     lists in 4.6 GiB of global-memory scratch that only lives for that repeat; VO_ERR_NOMEM if it cannot be had).
 *     The raw y-slab step (vo_slab_*) and the split passes (vo_pass1_dev / vo_pass2_dev) stop at 512; the multi-GPU
 *     operators (vo_mg_*) leave their overlapped slab step for the plain passes then, which repeat in the same way;
 *   - host CSR inputs are validated (off[0] = 0, offsets non-decreasing: VO_ERR_ARG); volumes that are ALREADY in device
 *     memory (vo_dvol_from_device, the vo_*_dev entry points) are trusted.
 *
 * Errors: every function returns VO_OK (0) or a VO_ERR_* code; vo_last_error(ctx) gives the text.
 * The C++ adapters (voroffset_b200/cpp) turn a non-zero code into std::runtime_error, which is what
 * the reference's vor_assert does (src/vor3d/Common.cpp:7-17).
 */
#ifndef VOROFFSET_B200_H
#define VOROFFSET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VO_OK              0
#define VO_ERR_ARG         1   /* invalid argument (bad op / method / sizes / unsorted offsets)        */
#define VO_ERR_CUDA        2   /* CUDA runtime failure, or no usable device                            */
#define VO_ERR_NOMEM       3   /* host or device allocation failed                                     */
#define VO_ERR_OVERFLOW    4   /* interval count exceeds the uint32 CSR offsets / a list capacity      */

/* op for the 3D entry points: app/cli3d/offset3d.cpp:116-136 (-x dilation|erosion|opening|closing). */
#define VO_OP_DILATION     0
#define VO_OP_EROSION      1
#define VO_OP_OPENING      2
#define VO_OP_CLOSING      3
/* method: app/cli3d/offset3d.cpp:104-112 (-m ours|brute_force).                                      */
#define VO_METHOD_OURS         0   /* VoronoiMorphoVorPower  (src/vor3d/VoronoiVorPower.cpp:24-96)   */
#define VO_METHOD_BRUTE_FORCE  1   /* VoronoiMorphoBruteForce (src/vor3d/VoronoiBruteForce.cpp:16-100) */
/* op for the 2D entry points: DoubleCompressedImage::dilate/erode/open/close/negate
 * (src/vor2d/DoubleCompressedImage.h:96-111).                                                        */
#define VO_OP2D_DILATE     0
#define VO_OP2D_ERODE      1
#define VO_OP2D_OPEN       2
#define VO_OP2D_CLOSE      3
#define VO_OP2D_NEGATE     4

typedef struct vo_ctx  vo_ctx;
typedef struct vo_dvol vo_dvol;   /* a dexel volume resident in HBM (device CSR)                      */
typedef struct vo_dmid vo_dmid;   /* the intermediate volume between the two passes, resident in HBM:
                                     replaces CompressedVolumeWithRadii (CompressedVolumeWithRadii.h:10-38) */

/* ---- lifecycle --------------------------------------------------------------------------------- */
int         vo_create(int device, vo_ctx **out);
void        vo_destroy(vo_ctx *ctx);
const char *vo_last_error(const vo_ctx *ctx);   /* never NULL; "" when the last call succeeded         */
const char *vo_version(void);
int         vo_span_bytes(void);                /* 16: spans are (double z1, double z2)                */
void        vo_free(void *host_ptr);            /* releases a host buffer returned by vo_morph3d/2d    */
/* Stream the context launches on (a cudaStream_t), for callers that time with CUDA events.           */
void       *vo_stream(const vo_ctx *ctx);
/* Number of kernel launches issued by this context so far (bench.py's gpu_launches).                 */
uint64_t    vo_launch_count(const vo_ctx *ctx);

/* Tuning knobs (none changes a result bit). "scan" = "fused" (default: staged lists -> CSR and both complements in one
 * single-pass look-back kernel each) | "classic"; "tile_dbuf" = "auto" | "on" | "off" (double-buffered candidate staging
 * of the pass-1 tile kernel). "pass1" = "auto" (default: tile kernel when it fits) | "tile" | "simple"
 * (one thread per (x, y, class); kept as the general fallback and as an independent implementation for the tests);
 * "tile_order" = "on" | "off" (expensive pass-1 tiles first); "tile_ctas" = 1..8 (CTAs per SM of the tile kernel);
 * "block_cache" = "on" | "off" (released scratch blocks >= 1 MiB kept whole for the next call);
 * host-buffer call (vo_morph3d): "pipeline" = "on" | "off" (bands of rows uploaded, processed and downloaded
 * concurrently), "bands" = 3..64, "band_split" = 1..4, "band_free" = 0..96 SMs, "pipe_warps" = 1..64,
 * "copy_align" = 0 | 16..65536 (band copies start and end on such a boundary; default 256), "pipe_lean" = "on" | "off"
 * (bands leave out the launches that are idle for height-field-like input; a call that needed one is redone on the plain
 * path and the context remembers), "pipe_ahead" = "on" | "off" (a band's offsets download enqueued before its total is
 * known), "pipe_order_one" = "on" | "off", "copy_batch" = "on" | "off" (the two copies of a band and direction as one
 * cudaMemcpyBatchAsync), "copy_out" = 0 (default: span downloads by the copy engine) | 1..1024 (CTAs of an
 * SM-driven span download that needs no host round trip), "erosion" = "auto" | "dual" | "general", "multi_warps" = 0
 * (auto) | 1..16 (warps per CTA of the tile kernel's multi-interval launches), "cand_order" = "auto" | "column" | "layer"
 * (order in which those launches walk a tile's candidates; a speed knob, the result never depends on it), "tile_general" =
 * "auto" | "redo" | "inline" (first tile launch: classes that need the sorted-list union go to the redo launch - 20 warps
 * per SM - or are folded inline - 16 warps; auto switches to inline for a while when a call met such classes), "pass2_union" =
 * "auto" | "registers" | "lists" (pass 2 on one-interval-per-column input: running union of capacity 2 in registers, a
 * third interval goes to the redo launch; auto falls back to the list-capable kernel when many columns needed it);
 * "tile_lean" = "auto" | "on" | "off" (candidates of the tile kernel's list launches left in global memory);
 * y-slab step: "slab" = "overlap" | "serial", "slab_reserve" = SMs the interior launch leaves to the NCCL kernels (8).
 * Experiment switches of the host-buffer pipeline, all measured and left at their defaults:
 * "band_weights" = "1,2,..." (relative band heights), "pipe_ctas", "pipe_quota", "pipe_warps0", "pipe_first_full",
 * "pipe_interleave", "pipe_mid"; development builds only: "pipe_dry", "ktrace", "ktrace_dump" (-DVO_KTRACE),
 * "tile_debug" (-DVO_TILE_DEBUG). (DESIGN.md 4.1-4.3 say what each is for and what it measured.)     */
int         vo_set_option(vo_ctx *ctx, const char *key, const char *value);
/* Timing on the context's own stream (torch.cuda.Event only sees torch's streams): vo_mark records event
 * `slot` (0..7); vo_elapsed_ms waits for slot_b and returns the device time between the two marks.     */
int         vo_mark(vo_ctx *ctx, int slot);
int         vo_elapsed_ms(vo_ctx *ctx, int slot_a, int slot_b, double *ms);
/* Device time of the two dominant kernels (k_pass1, k_pass2; first launch of each) of the most recent
 * 'ours' dilation on this context, measured with CUDA events around the launches.                     */
int         vo_last_profile(vo_ctx *ctx, double *k_pass1_ms, double *k_pass2_ms);

/* ---- host-buffer drop-in: one call = upload, operator, download -------------------------------- */
/* Replaces  VoronoiMorpho::dilation / ::erosion  (src/vor3d/Voronoi.h:18,29; Voronoi.cpp:8-17) and the
 * opening / closing compositions of app/cli3d/offset3d.cpp:124-133.
 *   zmin, zmax : origin_z/spacing and zmin + 2*padding + extent_z/spacing (VoronoiVorPower.cpp:28-29);
 *                only erosion (and therefore opening / closing) reads them (Voronoi.cpp:10-11).
 *   radius     : in dexels (offset3d.cpp:73-75 has already divided by the spacing for -u).
 *   out_off / out_spans : pinned host buffers owned by the caller afterwards, release with vo_free().
 *                The output grid is nx*ny again (Voronoi.cpp:57-89 strips erosion's border).
 *   ms_pass1 / ms_pass2 : device time of the two passes in milliseconds, the time_1 / time_2
 *                out-parameters of the reference (VoronoiVorPower.cpp:66,95). For composites they hold
 *                the last primitive, like offset3d.cpp:127-133. May be NULL.                          */
int vo_morph3d(vo_ctx *ctx, int op, int method, int nx, int ny, double zmin, double zmax,
               const uint32_t *off, const double *spans, double radius,
               uint32_t **out_off, double **out_spans, uint64_t *out_nspans,
               double *ms_pass1, double *ms_pass2);

/* The same call on a WINDOW of rows: the operator runs on the grid (nx, ny) that is passed, but only the rows
 * [row0, row1) of its result are computed to the end and returned (offsets starting at 0). Meant for a y-slab of a
 * larger grid handed over together with its ghost rows - the decomposition of VoronoiVorPower.cpp:41-63,70-92 with the
 * halo cut from HOST memory, where a caller that holds the whole grid has it anyway (no device-to-device exchange):
 * `off` may point into the middle of the larger CSR (ny*nx+1 entries, off[0] != 0 allowed) and `spans` stays the
 * base of the larger span array (span k of the window is spans[2k], k counted like the offsets). The ends of the
 * passed grid are treated as the ends of the volume (erosion's border), so pass as many ghost rows as the operator
 * reaches: floor(radius) for a dilation or an erosion, twice that for an opening or a closing; none at the true ends
 * of the volume. A large 'ours' dilation runs as the banded pipeline of vo_morph3d (upload, passes and download
 * overlap), every other case as upload - operator - rows - download.                                           */
int vo_morph3d_rows(vo_ctx *ctx, int op, int method, int nx, int ny, double zmin, double zmax,
                    const uint32_t *off, const double *spans, double radius, int row0, int row1,
                    uint32_t **out_off, double **out_spans, uint64_t *out_nspans,
                    double *ms_pass1, double *ms_pass2);

/* Replaces  DoubleCompressedImage::dilate / erode / open / close / negate
 * (src/vor2d/DoubleCompressedImage.cpp:438-468,680-719). `r` is the argument of the member function,
 * untouched: dilate sweeps with R = r*rows, erode with R = r (DoubleCompressedImage.cpp:685-686,698-699). */
int vo_morph2d(vo_ctx *ctx, int op, int rows, int width, const uint32_t *off, const double *spans, double r,
               uint32_t **out_off, double **out_spans, uint64_t *out_nspans, double *ms);

/* Replaces  VoronoiMorpho::calculateXor (src/vor3d/Voronoi.cpp:91-111): symmetric difference of two
 * same-grid volumes, slivers shorter than 1e-10 dropped (MorphologyOperators.cpp:354-374); *volume is
 * spacing^3 * total length (CompressedVolume.cpp:61-73).                                              */
int vo_xor3d(vo_ctx *ctx, int nx, int ny, double zmin, double zmax, double spacing,
             const uint32_t *off_a, const double *spans_a, const uint32_t *off_b, const double *spans_b,
             uint32_t **out_off, double **out_spans, uint64_t *out_nspans, double *volume);

/* ---- device-resident API (inputs and results stay in HBM) --------------------------------------- */
/* CompressedVolume storage (src/vor3d/CompressedVolume.h:14) converted to device CSR.                 */
int  vo_dvol_upload(vo_ctx *ctx, int nx, int ny, const uint32_t *off, const double *spans, vo_dvol **out);
/* Destination buffers may be host or device memory (cudaMemcpyDefault).                               */
int  vo_dvol_download(vo_ctx *ctx, const vo_dvol *vol, uint32_t *off, double *spans);
/* Same as upload, from device pointers (e.g. buffers an NCCL halo exchange has just filled).           */
int  vo_dvol_from_device(vo_ctx *ctx, int nx, int ny, const void *d_off, const void *d_spans, uint64_t nspans,
                         vo_dvol **out);
/* Shape and raw device pointers (uint32_t* / double*) of a resident volume.                           */
int  vo_dvol_info(const vo_dvol *vol, int *nx, int *ny, uint64_t *nspans, const void **d_off, const void **d_spans);
void vo_dvol_free(vo_ctx *ctx, vo_dvol *vol);
/* Copy of rows [y0, y1) of a resident volume (a y-slab; rows are contiguous in the x-fastest layout). */
int  vo_dvol_rows(vo_ctx *ctx, const vo_dvol *vol, int y0, int y1, vo_dvol **out);
/* Rows [y0, y1) copied straight into caller-owned DEVICE buffers (d_off: (y1-y0)*nx+1 uint32, rebased to
 * start at 0; d_spans: room for cap_spans intervals). *nspans is always set; the spans are only copied
 * when they fit. This is what a halo exchange sends (voroffset_b200/slab.py).                          */
int  vo_dvol_rows_to(vo_ctx *ctx, const vo_dvol *vol, int y0, int y1, void *d_off, void *d_spans,
                     uint64_t cap_spans, uint64_t *nspans);
/* Concatenate up to three y-slabs of equal nx (NULL entries are skipped).                             */
int  vo_dvol_concat_rows(vo_ctx *ctx, const vo_dvol *a, const vo_dvol *b, const vo_dvol *c, vo_dvol **out);

/* Same operators as vo_morph3d on resident volumes.                                                   */
int  vo_morph3d_dev(vo_ctx *ctx, int op, int method, const vo_dvol *in, double zmin, double zmax,
                    double radius, vo_dvol **out, double *ms_pass1, double *ms_pass2);
int  vo_xor3d_dev(vo_ctx *ctx, const vo_dvol *a, const vo_dvol *b, double zmin, double zmax, double spacing,
                  vo_dvol **out, double *volume);

/* The two passes of 'ours' separately (what a multi-GPU driver interleaves with its halo exchange).
 * Pass 1 (x-direction): for every column and every radius class j = |dy| the union over |dx| <= reach(j)
 * of the column's neighbours capped by sqrt(r1(j)^2 - dx^2)  - the work of
 * VoronoiMorpho2D / halfDilate (Voronoi2D.cpp:591-739, HalfDilationOperator.cpp:6-29).
 * Pass 2 (y-direction): out(x,y) = union over |dy| <= floor(R) of class |dy| of column (x, y+dy) - the work
 * of SeparatePowerMorpho2D + unionMap (SeparatePower2D.cpp:215-341, HalfDilationOperator.hpp:5-16).
 * Pass 2 produces rows [y0, y1) of the mid volume's grid.                                              */
int  vo_pass1_dev(vo_ctx *ctx, const vo_dvol *in, double radius, vo_dmid **mid, double *ms);
int  vo_pass2_dev(vo_ctx *ctx, const vo_dmid *mid, int y0, int y1, vo_dvol **out, double *ms);
void vo_dmid_free(vo_ctx *ctx, vo_dmid *mid);
int  vo_dmid_info(const vo_dmid *mid, int *nx, int *ny, int *classes, uint64_t *bytes);

/* Overlapped form of the two calls above for a y-slab of a grid sharded over several GPUs (voroffset_b200/slab.py;
 * the reference's dormant TBB decomposition, VoronoiVorPower.cpp:41-63,70-92, stretched across devices).
 * vo_slab_begin starts pass 1 on the rows of `own` that do not depend on a neighbour and returns at once; the caller
 * exchanges the floor(R) boundary rows with its neighbours meanwhile (NCCL); vo_slab_finish takes the received halos
 * (device pointers: (floor(R)*nx + 1) uint32 offsets starting at 0 and n intervals each; NULL / 0 where there is no
 * neighbour), finishes pass 1, runs pass 2 on the own rows and releases the slab. cap_prev / cap_next: upper bounds
 * of the halos' interval counts, agreed with the neighbours beforehand. VO_ERR_ARG from vo_slab_begin = not a case
 * for the overlapped path (small grid, radius < 1, ...): use vo_pass1_dev / vo_pass2_dev on the concatenated rows.   */
typedef struct vo_slab vo_slab;
/* d_off_to_* / d_spans_to_* (optional): send buffers of the two neighbours. The boundary rows are packed into them on
 * the device before pass 1 starts - floor(R)*nx + 1 offsets, then [interval count, overflow flag]; the spans only
 * when they fit cap_to_* (flag 1 otherwise) - and comm_stream (a cudaStream_t) is made to wait for that packing.      */
int  vo_slab_begin(vo_ctx *ctx, const vo_dvol *own, double radius, int has_prev, int has_next,
                   uint64_t cap_prev, uint64_t cap_next,
                   void *d_off_to_prev, void *d_spans_to_prev, uint64_t cap_to_prev,
                   void *d_off_to_next, void *d_spans_to_next, uint64_t cap_to_next, void *comm_stream, vo_slab **out);
int  vo_slab_finish(vo_ctx *ctx, vo_slab *slab, const void *d_off_prev, const void *d_spans_prev, uint64_t n_prev,
                    const void *d_off_next, const void *d_spans_next, uint64_t n_next,
                    vo_dvol **out, double *ms_pass1, double *ms_pass2);
void vo_slab_abort(vo_ctx *ctx, vo_slab *slab);

/* ---- multi-GPU: one grid cut into y-slabs over several B200s, halo rows exchanged with NCCL ------------------------ */
/* Replaces the same calls as vo_morph3d (Voronoi.h:18,29; offset3d.cpp:104-136) for a grid sharded the way the
 * reference's dormant TBB regions cut it (VoronoiVorPower.cpp:41-63,70-92: tasks own disjoint slices). Slab g of G owns
 * the rows [g ny/G, (g+1) ny/G) (remainder to the first slabs); before pass 1 every slab receives the floor(radius)
 * boundary rows of its two neighbours (ncclSend / ncclRecv, csrc/vo_mg.cuh); erosion keeps the reference's one-line solid
 * border (Voronoi.cpp:18-55) on the two edge slabs. All four operations and both methods; results are bit-identical to
 * the single-GPU call. libnccl.so.2 is loaded at run time (the copy PyTorch has loaded, if any).
 *   vo_mg_create        one process drives n_dev GPUs (device_ids NULL: 0 .. n_dev-1), one host thread per GPU while an
 *                       operator runs. n_dev = 1 works without NCCL.
 *   vo_mg_unique_id / vo_mg_create_rank   one process per GPU (torchrun): rank 0 makes the 128-byte id, every rank
 *                       passes the same bytes; the group then has ONE local rank.                                        */
typedef struct vo_mg vo_mg;
int         vo_mg_create(const int *device_ids, int n_dev, vo_mg **out);
int         vo_mg_unique_id(void *id128);
int         vo_mg_create_rank(int device, int rank, int world, const void *id128, vo_mg **out);
void        vo_mg_destroy(vo_mg *mg);
int         vo_mg_world(const vo_mg *mg);                 /* slabs = GPUs of the whole group                              */
int         vo_mg_local_count(const vo_mg *mg);           /* ranks driven by this process                                 */
int         vo_mg_rank(const vo_mg *mg, int local);       /* global rank of a local one                                   */
vo_ctx     *vo_mg_ctx(vo_mg *mg, int local);              /* its context (upload / download / free of its slab volumes)   */
const char *vo_mg_last_error(const vo_mg *mg);
/* Host buffers in, host buffers out (single-process groups): the drop-in call of offset3d --gpus N.                    */
int vo_mg_morph3d(vo_mg *mg, int op, int method, int nx, int ny, double zmin, double zmax,
                  const uint32_t *off, const double *spans, double radius,
                  uint32_t **out_off, double **out_spans, uint64_t *out_nspans, double *ms_pass1, double *ms_pass2);
/* Resident slabs: in[i] / out[i] belong to local rank i (uploaded through vo_mg_ctx(mg, i)); every slab but the thinnest
 * grid needs at least floor(radius) rows. Collective: every process of the group calls it with the same op / radius.  */
int vo_mg_morph3d_dev(vo_mg *mg, int op, int method, const vo_dvol *const *in, double zmin, double zmax, double radius,
                      vo_dvol **out, double *ms_pass1, double *ms_pass2);
/* Halo traffic of the last operator on a local rank: device time of the NCCL groups, host time blocked until the halos
 * had landed (what the step stalls for beyond the overlapped pass 1), bytes sent, NCCL groups, primitives that took the
 * overlapped slab step / the concatenate-then-dilate path.                                                             */
int vo_mg_stats(const vo_mg *mg, int local, double *halo_ms, double *halo_wait_ms, uint64_t *halo_bytes,
                int *messages, int *overlapped, int *plain);

/* ---- the step before the path: mesh -> dexel volume --------------------------------------------- */
/* Replaces the ray-marching loop of vor3d::create_dexels, compute_sign (src/vor3d/Dexelize.cpp:166-225, with
 * point_in_triangle_2d :74-92 and intersect_ray_z :136-162): for every column (x,y) of an nx*ny grid the z values
 * (in dexel units, z / spacing, :204) at which the vertical line through the column centre
 * ((x+0.5)*spacing + origin_x, (y+0.5)*spacing + origin_y) crosses a facet, ascending (:210), as (z1,z2) pairs.
 *   verts : double[3*nv] xyz; tris : int32[3*nf] vertex indices (host or device memory).
 *   origin_x/y, spacing, nx, ny : the CompressedVolume's own (origin already lowered by padding*spacing,
 *                CompressedVolume.cpp:17-21); a column with an odd number of crossings (open surface) loses the last.
 * The result stays resident in HBM, ready for vo_morph3d_dev; *ms = device time.                          */
int  vo_dexelize_dev(vo_ctx *ctx, uint64_t nv, const double *verts, uint64_t nf, const int32_t *tris,
                     double origin_x, double origin_y, double spacing, int nx, int ny, vo_dvol **out, double *ms);

/* 2D on resident data (one list per row).                                                             */
int  vo_morph2d_dev(vo_ctx *ctx, int op, const vo_dvol *rows_as_vol /* nx = rows, ny = 1 */, int width,
                    double r, vo_dvol **out, double *ms);

#ifdef __cplusplus
}
#endif
#endif /* VOROFFSET_B200_H */
