/*
 * sgn_raster.h -- C ABI of libsgn_raster.so, the B200-native (sm_100a) Gaussian rasterizer hot path
 * that replaces the gsplat 0.1.x calls made by street-gaussians-ns.
 *
 * Every entry point below names the reference interface it stands in for (paths relative to
 * /root/reference/street_gaussians_ns/).  gsplat itself is an un-vendored pip dependency of the
 * reference; its semantics are restated in SURVEY.md Appendix A.
 *
 * Conventions
 *   - all data pointers are DEVICE pointers owned by the caller (e.g. torch allocations), fp32 /
 *     int32, row-major contiguous, 16-byte aligned;  `stream` is a cudaStream_t passed as void*.
 *   - structs (sgn_camera, sgn_segment, ...) are plain host structs passed by pointer; segment
 *     tables are staged to a caller-provided device buffer with sgn_upload().
 *   - every function returns 0 on success, a negative sgn_status otherwise; the message is
 *     available from sgn_last_error() (thread local).  No exceptions cross the ABI, no allocation
 *     inside the library: scratch sizes are queried, buffers are passed in.  The only process-wide state is a launch
 *     counter and one lazily created auxiliary stream per device, onto which sgn_blend_fwd / sgn_blend_bwd fork the
 *     object-class pass (event fork / join around it: calls from several host threads on different streams stay
 *     correctly ordered, they merely share that side stream).
 *   - there is NO CPU fallback: a missing device or a failed launch is an error.
 */
#ifndef SGN_RASTER_H_
#define SGN_RASTER_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGN_ABI_VERSION 2
#define SGN_MAX_FOURIER 8
#define SGN_RECORD_FLOATS 12 /* per-Gaussian projected record, see below */

typedef enum sgn_status {
    SGN_OK = 0,
    SGN_ERR_INVALID = -1,   /* bad argument (shape, alignment, block width ...) */
    SGN_ERR_CUDA = -2,      /* a CUDA runtime call or launch failed */
    SGN_ERR_WORKSPACE = -3, /* scratch buffer too small */
    SGN_ERR_OVERFLOW = -4   /* intersection count exceeds the provided capacity */
} sgn_status;

/* Per-Gaussian projected record (12 floats, 48 B, 16-byte aligned):
 *   [0] x  [1] y            pixel centre            == gsplat project_gaussians `xys`
 *   [2] a  [3] b  [4] c     conic (inverse cov2d)   == `conics`
 *   [5] opacity             sigmoid(logit)          (sgn_splatfacto.py:946-949)
 *   [6] r  [7] g  [8] b     clamp(SH+0.5, min 0)    (sgn_splatfacto.py:939-940)
 *   [9] depth               view-space z            == `depths`
 *   [10] aux (int bits)     bits 0-2: colour clamp pass mask, bit 3: object class, bit 4: visible
 *   [11] unused
 * xys / conics / depths handed back to Python are strided views of this array. */

/* One visible sub-model of the scene graph for one frame (sgn_splatfacto_scene_graph.py:332-360).
 * Rows [row0, row0+count) of the concatenated index space; concatenation order is the reference's:
 * background first, then actors in annotation order. */
typedef struct sgn_segment {
    int32_t row0;
    int32_t count;
    int32_t F;        /* fourier_features_dim of features_dc (1..8)            (:239-247) */
    int32_t cls;      /* 0 background, 1 object                                 (:364-366) */
    int32_t has_pose; /* apply R,t,q  (object2world_gs, :404-417)                          */
    int32_t chunk0;   /* index of this segment's first 128-row chunk: sum over earlier segments of ceil(count/128) */
    float R[9];       /* object->world rotation, row-major (Box.rot cast to fp32, :411)    */
    float t[3];       /* Box.center cast to fp32 (:410)                                    */
    float q[4];       /* quaternion_from_matrix(rot), wxyz (:413)                          */
    float idft[SGN_MAX_FOURIER]; /* IDFT(t, F) basis (:420-433); [1,0,..] when F == 1      */
    const float* means;         /* [count,3]          gauss_params (sgn_splatfacto.py:291-300) */
    const float* scales;        /* [count,3] log                                            */
    const float* quats;         /* [count,4] wxyz, un-normalised                            */
    const float* features_dc;   /* [count,F,3]                                              */
    const float* features_rest; /* [count,(sh_degree+1)^2-1,3]                              */
    const float* opacities;     /* [count,1] logit                                          */
} sgn_segment;

/* Gradient destinations, one per segment, same shapes as the parameters (dense: rows the
 * rasterizer never touched receive zeros, as autograd of the reference produces). */
typedef struct sgn_segment_grads {
    float* means;
    float* scales;
    float* quats;
    float* features_dc;
    float* features_rest;
    float* opacities;
} sgn_segment_grads;

/* Camera + render settings (sgn_splatfacto.py:822-841, 860-873, 934-938). */
typedef struct sgn_camera {
    float viewmat[12]; /* world->camera 3x4 row-major (viewmat[:3,:], :831-836) */
    float fx, fy, cx, cy;
    int32_t width, height;
    float cam_pos[3];  /* camera_to_worlds[:3,3] for SH view directions (:934) */
    float limx, limy;  /* 1.3*tan(fov/2), float32 as gsplat computes it */
    float clip_thresh; /* 0.01 */
    int32_t block_width;      /* config.block_width; binning semantics (tile AABB) */
    int32_t sh_degree;        /* coefficients stored */
    int32_t sh_degree_to_use; /* min(step//interval, sh_degree) train, sh_degree eval (:936-938) */
} sgn_camera;

/* Blend settings. */
typedef struct sgn_blend_opts {
    float alpha_clamp_fwd; /* 0.999  gsplat rasterize_forward  */
    float alpha_clamp_bwd; /* 0.99   gsplat rasterize_backward */
    int32_t class_streams; /* also produce objects-only / background-only accumulation
                              (get_submodel_output, sgn_splatfacto_scene_graph.py:364-366) */
    int32_t has_sky;       /* rgb = rgb*alpha + sky*(1-alpha) (sgn_splatfacto.py:971-972) */
    int32_t eval_clamp;    /* not training: rgb.clamp(0,1) (:974-975) */
    /* tuning (0 = default): list length / traversal depth up to which one warp renders a whole tile;
     * each doubling splits the tile into 2/4/8 row strips rendered by independent warps */
    int32_t split_fwd_main, split_fwd_acc, split_bwd_main, split_bwd_acc;
    /* raw mode (gsplat rasterize_gaussians semantics, Level-1 shim): no post-ops; the four blended
     * channels come back as out = sum(c*alpha*T) + T_final*background[c] in rgb[...,0:3] and depth */
    int32_t raw_mode;
    float background[4];
    int32_t tuning; /* SGN_TUNE_* bits: execution variants with identical results up to fp32 rounding */
} sgn_blend_opts;

/* main forward skips row pairs an entry cannot reach / that have fully terminated (exact no-op) */
#define SGN_TUNE_FWD_ROW_SKIP 1
/* Blackwell packed-FP32 (f32x2) slot bodies in the main forward / backward kernels */
#define SGN_TUNE_FWD_PACKED 4
#define SGN_TUNE_BWD_PACKED 8
/* the accumulation-only (object / background) kernels do NOT skip unreachable row pairs */
#define SGN_TUNE_ACC_NO_ROW_SKIP 16
/* experiment: the main forward reads its lists as materialised 48-byte staged entries moved by cp.async.bulk + mbarrier
 * (sgn_blend_fwd_out.staged must then point to 48 * M bytes of scratch) instead of gathering records per lane */
#define SGN_TUNE_FWD_TMA 32

const char* sgn_last_error(void);
int sgn_abi_version(void);
/* number of kernel launches this library has issued in this process (a CUB device-wide call counts as 1) */
long long sgn_launch_count(void);
size_t sgn_sizeof_segment(void);
size_t sgn_sizeof_segment_grads(void);
size_t sgn_sizeof_camera(void);

/* Async H2D copy of a small host table (segment / grads table) into caller-provided device memory. */
int sgn_upload(const void* host, size_t bytes, void* dev, void* stream);

/* ---- fused compose + project + SH + sigmoid -------------------------------------------------
 * Replaces, in ONE launch over all segments: get_fourier_features + object2world_gs + the six
 * torch.cat (sgn_splatfacto_scene_graph.py:332-360), exp(scales) / cat(dc,rest) / quat normalise
 * (sgn_splatfacto.py:857-858,864), gsplat project_gaussians (:860-873), view directions +
 * spherical_harmonics + clamp (:934-940) and sigmoid(opacities) (:946-949).
 * Outputs: records[N,12] (layout above), radii[N] i32, num_tiles_hit[N] i32, tile_bbox[N] (4 x u16:
 * xmin,ymin,xmax,ymax in tiles), tiles_touched[N] i32 = number of AABB tiles the Gaussian can really
 * reach (exact test, see sgn_bin_count), touch_mask[N] u32 = one bit per AABB tile when the AABB has at
 * most 32 tiles.  Work is split into 128-row chunks that never straddle segments:
 * num_chunks = sum over segments of ceil(count/128), sgn_segment.chunk0 = the segment's first chunk. */
int sgn_project_fwd(const sgn_segment* segs_dev, int nseg, int N, int num_chunks, const sgn_camera* cam,
                    float* records, int32_t* radii, int32_t* num_tiles_hit, uint16_t* tile_bbox,
                    int32_t* tiles_touched, uint32_t* touch_mask, void* stream);

/* Backward of the above.  v_records[N,12] holds the per-Gaussian cotangents in record layout
 * ([0:2] v_xy, [2:5] v_conic, [5] v_opacity, [6:9] v_rgb, [9] v_depth), as accumulated by
 * sgn_blend_bwd.  Writes dense parameter gradients for every segment.
 * Replaces gsplat project_gaussians backward + compute_sh_backward + autograd of the glue. */
int sgn_project_bwd(const sgn_segment* segs_dev, const sgn_segment_grads* grads_dev, int nseg, int N, int num_chunks,
                    const sgn_camera* cam, const float* records, const int32_t* radii,
                    const float* v_records, void* stream);
/* The same backward over the chunks [chunk_begin, chunk_end) only (rows of the concatenated row space in 128-row chunks):
 * the data-parallel step produces the gradient arena range by range so that the exchange of a finished range
 * (sgn_allreduce_sym) overlaps the production of the next one. */
int sgn_project_bwd_range(const sgn_segment* segs_dev, const sgn_segment_grads* grads_dev, int nseg, int N, int num_chunks,
                          const sgn_camera* cam, const float* records, const int32_t* radii, const float* v_records,
                          int chunk_begin, int chunk_end, void* stream);

/* ---- Level-1: gsplat 0.1.x function API on plain tensors (sgn_splatfacto.py:11-14) -------------------
 * gsplat.project_gaussians(means3d, scales, glob_scale, quats, viewmat, fx, fy, cx, cy, H, W, block_width,
 * clip_thresh) -> xys[N,2], depths[N], radii[N] i32, conics[N,3], compensation[N], num_tiles_hit[N] i32,
 * cov3d[N,6]   (call site sgn_splatfacto.py:860-873), and its backward (cotangent pointers may be NULL). */
int sgn_l1_project_fwd(int N, const float* means, const float* scales, float glob_scale, const float* quats,
                       const sgn_camera* cam, float* xys, float* depths, int32_t* radii, float* conics,
                       float* compensation, int32_t* num_tiles_hit, float* cov3d, void* stream);
int sgn_l1_project_bwd(int N, const float* means, const float* scales, float glob_scale, const float* quats,
                       const sgn_camera* cam, const int32_t* radii, const float* v_xys, const float* v_depths,
                       const float* v_conics, float* v_means, float* v_scales, float* v_quats, void* stream);
/* gsplat.spherical_harmonics(degrees_to_use, viewdirs[N,3], coeffs[N,K,3]) (call site :939): forward when
 * `colors` is non-NULL, backward (v_coeffs = Y_k * v_colors) when `v_coeffs` is non-NULL. */
int sgn_l1_sh(int N, int K, int degree, const float* viewdirs, const float* coeffs, const float* v_colors,
              float* colors, float* v_coeffs, void* stream);

/* ---- binning: cumulative intersects, key emit, radix sort, tile bin edges ---------------------
 * Replaces the inside of gsplat rasterize_gaussians: compute_cumulative_intersects,
 * map_gaussian_to_intersects, torch.sort, get_tile_bin_edges. */
/* step 0 (only when the records do not come from sgn_project_fwd, e.g. the Level-1 rasterize path):
 * per-Gaussian count of the AABB tiles it can actually reach.  Exact, conservative ellipse-vs-tile test:
 * a tile where no pixel centre can have alpha >= 1/255 is a no-op for every stream and is dropped;
 * gsplat's num_tiles_hit is NOT changed. */
int sgn_bin_count(int N, const sgn_camera* cam, const float* records, const int32_t* radii,
                  const uint16_t* tile_bbox, int32_t* tiles_touched, uint32_t* touch_mask, void* stream);
/* step 1: stable depth order of the N rows (invisible rows last) and the inclusive scan of tiles_touched IN THAT
 * ORDER (cum[rank]); order[g] = where row g's run of entries starts in the depth-ordered entry sequence (the
 * exclusive scan value of its rank); the total M is also written to *total_dev (int64). */
size_t sgn_bin_scan_scratch_bytes(int N);
int sgn_bin_scan(int N, const float* records, const int32_t* radii, const int32_t* tiles_touched,
                 int32_t* order, int32_t* cum, int64_t* total_dev, void* scratch, size_t scratch_bytes, void* stream);
/* step 2: emit (in depth order) + stable sort by tile id + bin edges for M = total intersections. */
size_t sgn_bin_sort_scratch_bytes(int64_t M);
int sgn_bin_sort(int N, int64_t M, const sgn_camera* cam, const float* records, const int32_t* radii,
                 const uint16_t* tile_bbox, const uint32_t* touch_mask, const int32_t* order, const int32_t* cum,
                 int32_t* sorted_ids /*[M]*/, int32_t* tile_bins /*[tiles,2]*/, void* scratch, size_t scratch_bytes,
                 void* stream);
/* The same step without the host knowing M: `capacity` bounds the buffers (sorted_ids[capacity], scratch of
 * sgn_bin_sort_scratch_bytes(capacity)) and the launches, the real count is read from *total_dev on the device.  gsplat reads
 * the count back every frame (`cum_tiles_hit[-1].item()`); without that read-back the host can run ahead of the GPU.  Unused
 * slots are padded with a key behind every tile (the sort runs over `capacity` items).  If *total_dev > capacity the lists
 * are truncated and *overflow_dev is set to 1 (never cleared here): the caller checks it later and grows the capacity. */
int sgn_bin_sort_capped(int N, int64_t capacity, const int64_t* total_dev, int32_t* overflow_dev, const sgn_camera* cam,
                        const float* records, const int32_t* radii, const uint16_t* tile_bbox, const uint32_t* touch_mask,
                        const int32_t* order, const int32_t* cum, int32_t* sorted_ids, int32_t* tile_bins, void* scratch,
                        size_t scratch_bytes, void* stream);
/* EXPERIMENTAL alternative to steps 1 + 2 (csrc/binning_local.cu; same lists, same order, same payload): tile histogram +
 * unordered scatter + a shared-memory sort inside every tile instead of the two device-wide radix sorts.
 *   sgn_bin_local_count: tile_count[tiles], tile_start[tiles] (exclusive scan), info_dev = {M, longest list} (int64[2]);
 *   the caller reads info_dev, and when the longest list exceeds sgn_bin_local_cap() uses steps 1 + 2 for this frame;
 *   sgn_bin_local_sort: sorted_ids[M], tile_bins[tiles,2].  Scratch: sgn_bin_local_scratch_bytes(M, tiles) (M = 0 for count). */
int sgn_bin_local_cap(void);
size_t sgn_bin_local_scratch_bytes(int64_t M, int tiles);
int sgn_bin_local_count(int N, const sgn_camera* cam, const float* records, const int32_t* radii, const uint16_t* tile_bbox,
                        const uint32_t* touch_mask, int32_t* tile_count, int32_t* tile_start, int64_t* info_dev, void* scratch,
                        size_t scratch_bytes, void* stream);
int sgn_bin_local_sort(int N, int64_t M, int longest_list, const sgn_camera* cam, const float* records, const int32_t* radii,
                       const uint16_t* tile_bbox, const uint32_t* touch_mask, const int32_t* tile_count, const int32_t* tile_start,
                       int32_t* sorted_ids, int32_t* tile_bins, int32_t* cls_ids /*[2,M] or NULL*/,
                       int32_t* cls_bins /*[2,tiles,2] or NULL: the class sub-lists of step 3, built in the same pass*/,
                       void* scratch, size_t scratch_bytes, void* stream);
/* sorted_ids payload: bits 0-30 = Gaussian row (concatenated index space), bit 31 = object class.
 * step 3 (only for the class renders): per-tile class sub-lists, a stable partition of every tile's
 * list into background entries (class 0) and object entries (class 1) -- what the reference's
 * objects-only / background-only re-renders sort and traverse (sgn_splatfacto_scene_graph.py:
 * 255-303,364-366).  cls_ids is [2,M] (class c at cls_ids + c*M), cls_bins is [2,tiles,2]. */
size_t sgn_bin_class_scratch_bytes(int tiles);
int sgn_bin_class_lists(const sgn_camera* cam, int64_t M, const int32_t* sorted_ids, const int32_t* tile_bins,
                        int32_t* cls_ids, int32_t* cls_bins, void* scratch, size_t scratch_bytes, void* stream);

/* ---- alpha blending ------------------------------------------------------------------------------
 * Forward: gsplat rasterize_forward for rgb AND the depth pass in one traversal
 * (sgn_splatfacto.py:954-996), optionally the two accumulation-only re-renders
 * (sgn_splatfacto_scene_graph.py:364-366), with the reference's post-ops fused in the epilogue
 * (:968-975, :995).  Saves raw[H,W,4] (premultiplied rgb + depth), final_T / final_idx per stream. */
typedef struct sgn_blend_fwd_out {
    float* rgb;            /* [H,W,3] final */
    float* accumulation;   /* [H,W]   1 - T */
    float* depth;          /* [H,W]   where(alpha>1e-3, d/alpha, 10) */
    float* object_acc;     /* [H,W] or NULL */
    float* background_acc; /* [H,W] or NULL */
    float* raw;            /* [H,W,4] saved for backward */
    float* final_T;        /* [3,H,W] planar: slot 0 main, 1 object, 2 background */
    int32_t* final_idx;    /* [3,H,W] */
    int32_t* tile_depth;   /* [3,tiles] entries traversed per tile (main, object, background pass); sizes the backward */
    int32_t* sched;        /* scratch of sgn_blend_sched_ints(tiles) int32, or NULL: heavy-first work lists (longest tile
                              lists are scheduled first; without it CTAs take the tiles in raster order) */
    float* staged;         /* scratch of 12 * M floats (16-byte aligned) for SGN_TUNE_FWD_TMA, or NULL */
} sgn_blend_fwd_out;

size_t sgn_blend_sched_ints(int tiles);
int sgn_blend_fwd(const sgn_camera* cam, const sgn_blend_opts* opts, const float* records,
                  const int32_t* sorted_ids, const int32_t* tile_bins, int64_t M,
                  const int32_t* cls_ids /*[2,M] or NULL*/, const int32_t* cls_bins /*[2,tiles,2] or NULL*/,
                  const float* sky /*[H,W,3] or NULL*/, const sgn_blend_fwd_out* out, void* stream);

typedef struct sgn_blend_bwd_in {
    const float* v_rgb;            /* [H,W,3] or NULL */
    const float* v_accumulation;   /* [H,W]   or NULL */
    const float* v_depth;          /* [H,W]   or NULL */
    const float* v_object_acc;     /* [H,W]   or NULL */
    const float* v_background_acc; /* [H,W]   or NULL */
    const float* raw;
    const float* final_T;
    const int32_t* final_idx;
    const int32_t* tile_depth;     /* [3,tiles] from the forward */
    int32_t* sched;                /* scratch of sgn_blend_sched_ints(tiles) int32 (may be the forward's), or NULL */
    const float* sky;              /* [H,W,3] or NULL */
    float* v_sky;                  /* [H,W,3] or NULL: gradient to the sky colour */
    /* Deterministic mode (NULL = off): the per-Gaussian gradients are accumulated in 64-bit fixed point (one rounding per
     * addend, integer adds: bit-identical totals whatever the execution order) instead of with float atomics, then written
     * to v_records.  v_fixed: [num_gaussians,12] int64, ZERO on entry; fixed_scale: one float of device scratch. */
    int64_t* v_fixed;
    float* fixed_scale;
    int64_t num_gaussians;
} sgn_blend_bwd_in;

/* Backward: gsplat rasterize_backward for all streams in one traversal.  v_records[N,12] must be
 * zero on entry; it is accumulated into (record layout, see sgn_project_bwd). */
int sgn_blend_bwd(const sgn_camera* cam, const sgn_blend_opts* opts, const float* records,
                  const int32_t* sorted_ids, const int32_t* tile_bins, int64_t M,
                  const int32_t* cls_ids /*[2,M] or NULL*/, const int32_t* cls_bins /*[2,tiles,2] or NULL*/,
                  const sgn_blend_bwd_in* in, float* v_records, void* stream);

/* Generic per-Gaussian channels (the north-star's per-Gaussian semantic logits; the reference's dormant consumer:
 * scripts/render.py:188,231-236): extra[N,C] is composited with the weights of the main render -- out[p,c] = sum_k
 * extra[k,c] alpha_k T_k over the entries the main pass blended (final_T / final_idx slot 0 of sgn_blend_fwd) -- 8 channels per
 * traversal.  What the reference would obtain from one more gsplat rasterize_gaussians(colors = logits) call.  The backward
 * ACCUMULATES into v_extra[N,C] (zero it first) and into v_records[N,12] (geometry part: call it between sgn_blend_bwd and
 * sgn_project_bwd). */
int sgn_blend_extra_fwd(const sgn_camera* cam, const sgn_blend_opts* opts, const float* records, const int32_t* sorted_ids,
                        const int32_t* tile_bins, const float* final_T, const int32_t* final_idx, const float* extra, int C,
                        float* out /*[H,W,C]*/, void* stream);
int sgn_blend_extra_bwd(const sgn_camera* cam, const sgn_blend_opts* opts, const float* records, const int32_t* sorted_ids,
                        const int32_t* tile_bins, const float* final_T, const int32_t* final_idx, const float* extra, int C,
                        const float* v_out /*[H,W,C]*/, float* v_extra /*[N,C]*/, float* v_records /*[N,12]*/, void* stream);

/* ---- loss epilogue ------------------------------------------------------------------
 * The image-space loss terms of the reference that re-read the rasterizer's outputs right after the render, and
 * their cotangents: L1 = w_l1 * mean|gt - rgb| (sgn_splatfacto.py:1079-1084; with mask: both sides times mask,
 * :1073-1076), sky = w_sky * mean(sky_mask * accumulation) (:1090-1093), entropy = w_entropy * mean(-(oa log oa +
 * (1-oa) log(1-oa))) with oa = clamp(object_acc, 1e-5, 1-1e-5) (sgn_splatfacto_scene_graph.py:386-389).
 * A term whose inputs are NULL is 0 / gets no cotangent.  gt is the float image or the uint8 one (gt = u8 / 255). */
typedef struct sgn_loss_in {
    const float* rgb;           /* [H,W,3] or NULL */
    const uint8_t* gt_u8;       /* [H,W,3] exactly one of gt_u8 / gt_f32 when rgb is given */
    const float* gt_f32;
    const float* mask;          /* [H,W,1] or NULL */
    const float* accumulation;  /* [H,W,1] or NULL */
    const uint8_t* sky_mask;    /* [H,W] 1 = sky (semantic == SKY) or NULL */
    const float* object_acc;    /* [H,W,1] or NULL */
    float w_l1, w_sky, w_entropy;
} sgn_loss_in;
size_t sgn_loss_scratch_bytes(void);
/* losses[3] (device) = {L1, sky, entropy} terms, already weighted */
int sgn_loss_fwd(int H, int W, const sgn_loss_in* in, float* losses, void* scratch, size_t scratch_bytes, void* stream);
/* cotangents of the outputs for incoming gradients grad_losses[3] (device scalars; NULL = all ones); each output may be NULL */
int sgn_loss_bwd(int H, int W, const sgn_loss_in* in, const float* grad_losses, float* v_rgb, float* v_accumulation,
                 float* v_object_acc, void* stream);

/* ---- densification statistics -------------------------------------------------------
 * What each sub-model's `after_train` accumulates after backward (sgn_splatfacto.py:513-541), for all visible
 * sub-models of the frame in one launch: grads = ||v_records[:,0:2]||; on a sub-model's first call
 * xys_grad_norm = grads, vis_counts = 1, max_2Dsize = 0 (then the max below); afterwards, for rows with
 * radii > 0: xys_grad_norm += grads, vis_counts += 1; max_2Dsize = max(max_2Dsize, radii / max(H, W)). */
typedef struct sgn_densify_segment {
    int32_t row0, count;  /* rows of this sub-model in the frame's row space */
    int32_t first;        /* 1: the sub-model's statistics are being created by this call */
    int32_t pad;
    float* xys_grad_norm; /* [count] */
    float* vis_counts;    /* [count] */
    float* max_2Dsize;    /* [count] */
} sgn_densify_segment;
size_t sgn_sizeof_densify_segment(void);
int sgn_densify_stats(const sgn_densify_segment* table_dev, int nseg, int N, const float* v_records /*[N,12]*/,
                      const int32_t* radii /*[N]*/, int height, int width, void* stream);

/* ---- fused multi-tensor Adam ------------------------------------------------------
 * torch.optim.Adam semantics (betas, eps, no weight decay, no amsgrad) for every Gaussian parameter tensor in
 * ONE launch; replaces the nine nerfstudio Adam optimizers over ~200 tensors (sgn_config.py:71-108).
 * Gradients / exp_avg / exp_avg_sq are flat arenas of 16-byte aligned slices (offsets in floats, multiples of 4):
 * the gradient arena has the layout sgn_project_bwd wrote for the frame, the moment arenas cover every tensor of
 * the model; parameters are updated in place.  The table lists the tensors that have a gradient this step.  The host fills, per tensor and
 * per step, step_size = lr / (1 - beta1^t) and sqrt_bc2 = sqrt(1 - beta2^t), both evaluated in double. */
typedef struct sgn_adam_tensor {
    float* param;
    int64_t arena_offset; /* of the tensor's moments in exp_avg / exp_avg_sq */
    int64_t grad_offset;  /* of its gradient in grad_arena: the arena of a frame only holds the sub-models visible in that
                             frame (torch.optim.Adam skips parameters whose .grad is None: no decay, no step count) */
    int64_t numel;
    int32_t chunk0; /* first block of this tensor: sum over earlier tensors of ceil(numel / sgn_adam_chunk_elems()) */
    float beta1, beta2, eps, step_size, sqrt_bc2;
    float one_minus_beta1, one_minus_beta2; /* rounded from double on the host, as torch passes them (1 - 0.999f != 0.001f) */
} sgn_adam_tensor;
size_t sgn_sizeof_adam_tensor(void);
int sgn_adam_chunk_elems(void);
int sgn_adam_step(const sgn_adam_tensor* table_dev, int ntensors, int num_chunks, const float* grad_arena,
                  float* exp_avg, float* exp_avg_sq, void* stream);

/* ---- gradient exchange of the camera-sharded data-parallel step ---------------------------------
 * The reference has no distributed code; SURVEY 8e defines the exchange: SUM (or mean) of the flat per-Gaussian gradient
 * arena over the replicas.  sgn_allreduce_sym is that exchange as ONE kernel of this library over peer-mapped
 * ("symmetric") memory -- what a training loop would otherwise hand to ncclAllReduce: rank r reduces the r-th part of
 * every listed slice (multimem.ld_reduce through `multicast`, i.e. inside the NVSwitch; or, with multicast == NULL, loads
 * from every peer pointer in rank order), multiplies by `scale` and pushes the result to all replicas (multimem.st / one
 * store per peer), in place.  `local` is this rank's arena, `peer_ptrs_dev` a DEVICE array of `world` arena base pointers
 * (this rank's own included).  Slice offsets / lengths are in floats, multiples of 4.  The caller orders the call between
 * two cross-GPU barriers on the same stream (all replicas written; all parts pushed).  Bit-identical results on all ranks. */
#define SGN_AR_MAX_SLICES 48
/* Row skipping (optional, NULL = exchange everything): a slice with slice_widths[s] > 0 holds slice_rows[s] rows of that many
 * floats, row j of the slice is entry slice_row0[s] + j of `visible_union` (uint8, 1 = some replica saw the Gaussian).  Rows
 * nobody saw carry an all-zero gradient on every replica (sgn_project_bwd writes zeros there): their sum is already in place,
 * so they are neither pulled nor pushed.  sgn_visible_flags writes this rank's flags (radii > 0) -- into symmetric memory --
 * and sgn_visible_union ORs the peers' flags (peer base + flags_byte_offset) into a local array. */
int sgn_allreduce_sym(void* local, void* multicast, const uint64_t* peer_ptrs_dev, int rank, int world, int nslices,
                      const int64_t* slice_offsets, const int64_t* slice_lengths, const int32_t* slice_widths,
                      const int64_t* slice_row0, const int64_t* slice_rows, const uint8_t* visible_union, float scale, int max_ctas,
                      void* stream);
int sgn_visible_flags(const int32_t* radii, int64_t n, uint8_t* flags, void* stream);
int sgn_visible_union(const uint64_t* peer_ptrs_dev, int64_t flags_byte_offset, int world, int64_t n, uint8_t* out, void* stream);

/* ---- refinement: split / duplicate / cull ---------------------------------------------
 * What `SplatfactoModel.refinement_after` does to ONE sub-model every `refine_every` steps
 * (sgn_splatfacto.py:550-646 with cull_gaussians :648-672, split_gaussians :674-710, dup_gaussians :712-720 and the
 * optimizer surgery dup_in_optim / remove_from_optim :459-511), as two launches instead of ~120 torch statements
 * with eight host syncs per sub-model:
 *   sgn_refine_decide: per row, from the running statistics of sgn_densify_stats, the log-scales and the opacity
 *     logit -> flags (SGN_RF_* bits, csrc/sgn_refine_rules.cuh) and four 0/1 marks the caller prefix-sums;
 *   sgn_refine_apply: rebuilds the six parameter tensors and their Adam moments in the reference's row order
 *     [surviving old rows | split samples, sample-major | duplicates]: a split sample's mean is
 *     mean + R(q) (exp(scale) * z), split rows shrink by 1/1.6 in log space, new rows start with zero moments.
 * Both are HBM-bound streaming passes (decide: 32 B read per row; apply: 3 x 4 B read + written per element). */
typedef struct sgn_refine_config {
    int32_t densify;         /* 1: split + duplicate + cull (do_densification, :563-619); 0: cull only (:620-621) */
    int32_t n_split_samples; /* config.n_split_samples (2) */
    int32_t use_screen_size; /* step < stop_screen_size_at (:575, :662) */
    int32_t cull_big;        /* step > refine_every * reset_alpha_every (:659) */
    float max_size;          /* (float)max(last_size) (:570) */
    float densify_grad_thresh, densify_size_thresh, split_screen_size;
    float cull_alpha_thresh, cull_scale_thresh, cull_screen_size;
    float inv_size_fac;      /* fp32 reciprocal of size_fac = 1.6 (:694-695) */
} sgn_refine_config;

/* Source / destination of the rebuild: the six parameter tensors in gradient-arena order (means, scales, quats,
 * features_dc, features_rest, opacities) and, when an Adam state exists, exp_avg (m) and exp_avg_sq (v) per tensor
 * (all NULL = no optimizer state).  width[k] = floats per row of tensor k (3, 3, 4, 3F, 3(K-1), 1). */
typedef struct sgn_refine_tensors {
    const float* src[6];
    float* dst[6];
    const float* src_m[6];
    float* dst_m[6];
    const float* src_v[6];
    float* dst_v[6];
    int32_t width[6];
} sgn_refine_tensors;
size_t sgn_sizeof_refine_config(void);
size_t sgn_sizeof_refine_tensors(void);
/* flags[n] u8; marks[4,n] i32 = {survives, split samples survive, duplicate survives, is split} per row.
 * xys_grad_norm / vis_counts may be NULL when !cfg->densify, max_2Dsize when !cfg->use_screen_size. */
int sgn_refine_decide(int n, const sgn_refine_config* cfg, const float* scales /*[n,3]*/, const float* opacities /*[n,1]*/,
                      const float* xys_grad_norm, const float* vis_counts, const float* max_2Dsize, uint8_t* flags,
                      int32_t* marks, void* stream);
/* scan[4,n] = inclusive prefix sums of marks along the rows; totals[4] (HOST) = their last column; samples =
 * [n_split_samples * totals[3], 3] standard-normal draws (torch.randn, :680), NULL when totals[3] == 0.
 * Destinations hold totals[0] + n_split_samples * totals[1] + totals[2] rows. */
int sgn_refine_apply(int n, const sgn_refine_config* cfg, const sgn_refine_tensors* tensors, const uint8_t* flags,
                     const int32_t* scan, const int32_t* totals, const float* samples, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SGN_RASTER_H_ */
