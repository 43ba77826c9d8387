// Per-row rules of the refinement step (second half): which rows of a sub-model are split,
// duplicated and culled, and what a surviving / new row holds.  Restates, row by row, what
// `SplatfactoModel.refinement_after` does with whole-tensor torch statements
// (street_gaussians_ns/sgn_splatfacto.py:550-646, cull_gaussians :648-672, split_gaussians :674-710,
// dup_gaussians :712-720, dup_in_optim / remove_from_optim :459-511).
//
// The functions are plain C++ marked SGN_HD so that the SAME source is (a) the body of the CUDA kernels in
// refine.cu and (b) compiled by g++ into a host harness that tests/ checks against the torch restatement of the
// reference (there is no GPU in the build container; the harness is test infrastructure, not a product path).
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/sgn_raster.h"

#ifdef __CUDACC__
#define SGN_HD __host__ __device__ __forceinline__
#else
#define SGN_HD inline
#endif

// ---- flags of one source row ------------------------------------------------------------------------------
#define SGN_RF_SPLIT 0x01       // the row is split into n_split_samples new rows (and itself removed)
#define SGN_RF_DUP 0x02         // the row is duplicated once
#define SGN_RF_KEEP_ORIG 0x04   // the row itself survives the cull
#define SGN_RF_KEEP_SPLIT 0x08  // its split samples survive the cull (all samples share scale and opacity)
#define SGN_RF_KEEP_DUP 0x10    // its duplicate survives the cull
#define SGN_RF_HIGH_GRAD 0x20   // averaged pixel-space gradient norm above densify_grad_thresh
#define SGN_RF_ALPHA 0x40       // sigmoid(opacity) < cull_alpha_thresh
#define SGN_RF_TOOBIG 0x80      // culled for its size (world scale or screen size)

// torch: `torch.log(torch.exp(scales) / size_fac)` (:694-696).  A CUDA tensor divided by a python scalar is a
// multiplication with the scalar's fp32 reciprocal (ATen BinaryDivTrueKernel), hence inv_size_fac.
SGN_HD float sgn_refine_shrunk_scale(float log_scale, float inv_size_fac) { return logf(expf(log_scale) * inv_size_fac); }

SGN_HD float sgn_refine_max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

// Decision for one row.  `max_2dsize` is only read when cfg.use_screen_size; the statistics are only read when
// cfg.densify.
SGN_HD uint8_t sgn_refine_decide_row(const sgn_refine_config& cfg, float sx, float sy, float sz, float opacity_logit,
                                     float xys_grad_norm, float vis_count, float max_2dsize) {
    const float smax = sgn_refine_max3(expf(sx), expf(sy), expf(sz));  // exp(scales).max(dim=-1) (:574)
    float smax_now = smax;  // the row's largest scale as later statements see it
    bool high = false, split = false, dup = false;
    if (cfg.densify) {
        // avg_grad_norm = (xys_grad_norm / vis_counts) * 0.5 * max(H, W) (:570), thresholded in fp32 (:571)
        const float avg = ((xys_grad_norm / vis_count) * 0.5f) * cfg.max_size;
        high = avg > cfg.densify_grad_thresh;
        split = smax > cfg.densify_size_thresh;                                   // :574
        if (cfg.use_screen_size) split = split || (max_2dsize > cfg.split_screen_size);  // :575-576
        split = split && high;                                                    // :577
        // split_gaussians rewrites self.scales[split_mask] IN PLACE (:696) before `dups` is evaluated (:582):
        // a split row whose shrunk scale falls under the size threshold is duplicated as well
        if (split)
            smax_now = sgn_refine_max3(expf(sgn_refine_shrunk_scale(sx, cfg.inv_size_fac)),
                                       expf(sgn_refine_shrunk_scale(sy, cfg.inv_size_fac)),
                                       expf(sgn_refine_shrunk_scale(sz, cfg.inv_size_fac)));
        dup = (smax_now <= cfg.densify_size_thresh) && high;                      // :582-583
    }
    // cull_gaussians over [old rows | split samples | duplicates] (:648-672).  New rows carry max_2Dsize = 0 (:592-599)
    const float sig = 1.0f / (1.0f + expf(-opacity_logit));  // torch.sigmoid
    const bool alpha = sig < cfg.cull_alpha_thresh;          // :654
    bool big_new = false, big_old = false;
    if (cfg.cull_big) {                                      // step > refine_every * reset_alpha_every (:659)
        big_new = smax_now > cfg.cull_scale_thresh;          // :661
        big_old = big_new || (cfg.use_screen_size && max_2dsize > cfg.cull_screen_size);  // :662-665
    }
    uint8_t f = 0;
    if (split) f |= SGN_RF_SPLIT;
    if (dup) f |= SGN_RF_DUP;
    if (high) f |= SGN_RF_HIGH_GRAD;
    if (alpha) f |= SGN_RF_ALPHA;
    if (big_old) f |= SGN_RF_TOOBIG;
    if (!(alpha || split || big_old)) f |= SGN_RF_KEEP_ORIG;  // split rows are pruned (:607-618)
    if (split && !(alpha || big_new)) f |= SGN_RF_KEEP_SPLIT;
    if (dup && !(alpha || big_new)) f |= SGN_RF_KEEP_DUP;
    return f;
}

// Component `c` (0..2) of the offset a split sample adds to the mean (:680-687):
//   R(q / ||q||) (exp(scales) * z),  z ~ N(0, I);  quat_to_rotmat normalises once more (gsplat _torch_impl).
SGN_HD float sgn_refine_split_offset(int c, const float* q, const float* log_scales, const float* z) {
    const float n1 = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    float a[4] = {q[0] / n1, q[1] / n1, q[2] / n1, q[3] / n1};
    const float n2 = fmaxf(sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + a[3] * a[3]), 1e-12f);  // F.normalize eps
    const float w = a[0] / n2, x = a[1] / n2, y = a[2] / n2, zq = a[3] / n2;
    float r0, r1, r2;
    if (c == 0) {
        r0 = 1.f - 2.f * (y * y + zq * zq); r1 = 2.f * (x * y - w * zq); r2 = 2.f * (x * zq + w * y);
    } else if (c == 1) {
        r0 = 2.f * (x * y + w * zq); r1 = 1.f - 2.f * (x * x + zq * zq); r2 = 2.f * (y * zq - w * x);
    } else {
        r0 = 2.f * (x * zq - w * y); r1 = 2.f * (y * zq + w * x); r2 = 1.f - 2.f * (x * x + y * y);
    }
    return r0 * (expf(log_scales[0]) * z[0]) + r1 * (expf(log_scales[1]) * z[1]) + r2 * (expf(log_scales[2]) * z[2]);
}

// Column layout of one Gaussian across its six parameter tensors (PARAM order of the gradient arena):
// means 3 | scales 3 | quats 4 | features_dc 3F | features_rest 3(K-1) | opacities 1.
#define SGN_RT_MEANS 0
#define SGN_RT_SCALES 1
#define SGN_RT_QUATS 2
#define SGN_RT_DC 3
#define SGN_RT_REST 4
#define SGN_RT_OPAC 5

SGN_HD int sgn_refine_row_width(const sgn_refine_tensors& t) {
    return t.width[0] + t.width[1] + t.width[2] + t.width[3] + t.width[4] + t.width[5];
}

// One (source row i, column c) element of the rebuild: copies the element to every output row it feeds.
// Output order is the reference's: surviving old rows, then split samples sample-major (`repeat(samps, 1)`,
// :680-698), then duplicates, each in source-row order (boolean-mask indexing keeps order).
//   scan[k*n + i]: INCLUSIVE prefix sums over the rows of k = 0 KEEP_ORIG, 1 KEEP_SPLIT, 2 KEEP_DUP, 3 SPLIT
//   totals[4]:     their last values (kept old rows, kept split rows, kept duplicates, all split rows)
//   samples:       [n_split_samples * totals[3], 3] standard normal draws, sample-major like the reference's
//                  `torch.randn((samps * n_splits, 3))`
// Adam moments (dup_in_optim / remove_from_optim): old rows keep theirs, new rows start at zero.
SGN_HD void sgn_refine_apply_elem(int64_t i, int c, int64_t n, const sgn_refine_config& cfg, const sgn_refine_tensors& t,
                                  const uint8_t* flags, const int32_t* scan, const int32_t* totals, const float* samples) {
    const uint8_t f = flags[i];
    if (!(f & (SGN_RF_KEEP_ORIG | SGN_RF_KEEP_SPLIT | SGN_RF_KEEP_DUP))) return;
    int k = 0, lc = c;
    while (lc >= t.width[k]) { lc -= t.width[k]; ++k; }
    const int w = t.width[k];
    const int64_t e = i * w + lc;
    const float val = t.src[k][e];
    const bool moments = t.src_m[k] != nullptr;
    if (f & SGN_RF_KEEP_ORIG) {
        const int64_t o = (int64_t)(scan[0 * n + i] - 1) * w + lc;
        t.dst[k][o] = val;
        if (moments) { t.dst_m[k][o] = t.src_m[k][e]; t.dst_v[k][o] = t.src_v[k][e]; }
    }
    float nv = val;
    if (k == SGN_RT_SCALES && (f & SGN_RF_SPLIT)) nv = sgn_refine_shrunk_scale(val, cfg.inv_size_fac);  // :695-696
    const int64_t base_split = totals[0];
    const int64_t base_dup = base_split + (int64_t)cfg.n_split_samples * totals[1];
    if (f & SGN_RF_KEEP_DUP) {  // a copy of the row as it stands after the in-place rescale (:716-720)
        const int64_t o = (base_dup + scan[2 * n + i] - 1) * w + lc;
        t.dst[k][o] = nv;
        if (moments) { t.dst_m[k][o] = 0.f; t.dst_v[k][o] = 0.f; }
    }
    if (f & SGN_RF_KEEP_SPLIT) {
        const int64_t rank_kept = scan[1 * n + i] - 1, rank_all = scan[3 * n + i] - 1;
        for (int s = 0; s < cfg.n_split_samples; ++s) {
            float out = nv;
            if (k == SGN_RT_MEANS) {
                const float* z = samples + ((int64_t)s * totals[3] + rank_all) * 3;
                out = sgn_refine_split_offset(lc, t.src[SGN_RT_QUATS] + i * 4, t.src[SGN_RT_SCALES] + i * 3, z) + val;
            }
            const int64_t o = (base_split + (int64_t)s * totals[1] + rank_kept) * w + lc;
            t.dst[k][o] = out;
            if (moments) { t.dst_m[k][o] = 0.f; t.dst_v[k][o] = 0.f; }
        }
    }
}
