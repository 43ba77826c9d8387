/*
 * sgn_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the street-gaussians-ns rasterizer hot path:
 *   scene-graph compose  : street_gaussians_ns/sgn_splatfacto_scene_graph.py:239-247 (Fourier DC),
 *                          :404-417 (object2world_gs), :355-360 (concat order bg, actor1, actor2 ...)
 *   pre-ops + projection : street_gaussians_ns/sgn_splatfacto.py:857-873 (exp(scales), quats/|quats|,
 *                          gsplat project_gaussians -- gsplat 0.1.x, NOT vendored in the reference;
 *                          semantics restated from SURVEY.md Appendix A.1-A.4)
 *   SH / sigmoid         : street_gaussians_ns/sgn_splatfacto.py:933-949 (gsplat spherical_harmonics, A.7)
 *   binning + sort       : gsplat rasterize_gaussians internals
 *   blend fwd / bwd      : gsplat rasterize_forward / rasterize_backward
 *   project / SH bwd     : gsplat project_gaussians_backward / compute_sh_backward (Appendix A.7-A.8)
 *
 * PARITY: the reference ships no tests and gsplat is absent from the container, so gsplat's KERNEL
 * ARITHMETIC here (projection, SH, binning, blend, and their backward: SURVEY.md Appendix A) is a restatement --
 * "parity unpinned" for those -- pinned only by (a) closed-form known answers, (b) a float64 autograd restatement
 * (oracle/oracle_torch.py) and (c) finite differences -- see tests/test_oracle.py.  Everything AROUND those kernels
 * (scene-graph compose, camera, pre-ops, view directions, SH schedule, the four rasterize calls, post-ops, side
 * effects, the backward chain) is pinned to the reference's own code: tests/test_reference_glue.py runs the reference's
 * get_outputs + autograd on the CPU with this restatement in gsplat's slots and this file reproduces its outputs and
 * every gradient tensor (tests/golden/reference_glue_tiny.npz).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * call into this file.  It is never on the product path.
 *
 * Arithmetic contract ("exact section"): everything upstream of an integer decision (radius,
 * tile AABB, depth sort key) is computed in float32 with every * + - / sqrt individually rounded
 * (build with -ffp-contract=off), in the operation order written here.  The CUDA product follows
 * the same order (compiled with --fmad=false), which makes radii / num_tiles_hit / sort order /
 * tile_bins bit-comparable.  exp() in that section is sgn_expf_spec (a fixed polynomial), not libm.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define SGN_MAX_F 8

typedef struct {
    int32_t row0, count, F, cls, has_pose, pad0;
    float R[9];
    float t[3];
    float q[4];
    float idft[SGN_MAX_F];
    const float* means;
    const float* scales;
    const float* quats;
    const float* features_dc;
    const float* features_rest;
    const float* opacities;
} oracle_segment;

typedef struct {
    float viewmat[12]; /* 3x4 row-major world->camera (OpenCV: +z forward) */
    float fx, fy, cx, cy;
    int32_t width, height;
    float cam_pos[3]; /* camera_to_worlds[:3,3] (sgn_splatfacto.py:934) */
    float limx, limy; /* 1.3 * 0.5*W/fx, 1.3 * 0.5*H/fy computed by the host in float32 */
    float clip_thresh;
    int32_t block_width;
    int32_t sh_degree;        /* coefficients stored: (sh_degree+1)^2 */
    int32_t sh_degree_to_use; /* n at sgn_splatfacto.py:936-938 */
} oracle_camera;

/* ------------------------------------------------------------------------------------------ */
/* exact-section helpers                                                                       */
/* ------------------------------------------------------------------------------------------ */

/* exp() with a fixed operation sequence (see header).  |x| is clamped to 80. */
float sgn_expf_spec(float x) {
    if (x > 80.0f) x = 80.0f;
    if (x < -80.0f) x = -80.0f;
    float n = rintf(x * 1.44269504f);
    float r = x - n * 0.693145752f;
    r = r - n * 1.42860677e-6f;
    float p = 1.98412698e-4f;
    p = p * r + 1.38888889e-3f;
    p = p * r + 8.33333333e-3f;
    p = p * r + 4.16666667e-2f;
    p = p * r + 1.66666667e-1f;
    p = p * r + 0.5f;
    p = p * r + 1.0f;
    p = p * r + 1.0f;
    return ldexpf(p, (int)n);
}

/* Sensitivity switch (tools/exp_sensitivity.py; never used by the parity tests): which exp() the exact section uses.
 * 0 = sgn_expf_spec (the contract), 1 = libm expf (glibc, < 1 ulp), 2 / 3 = the spec value moved one ulp up / down.
 * The reference calls torch.exp (sgn_splatfacto.py:857), whose CUDA kernel is within 2 ulp of the true value: modes 1-3
 * bracket how many integer decisions (radius, tile AABB) could differ from the reference for that reason alone. */
static int g_exp_mode = 0;
void sgn_oracle_set_exp_mode(int mode) { g_exp_mode = mode; }
static inline float oracle_exp(float x) {
    if (g_exp_mode == 1) return expf(x);
    const float v = sgn_expf_spec(x);
    if (g_exp_mode == 2) return nextafterf(v, INFINITY);
    if (g_exp_mode == 3) return nextafterf(v, 0.f);
    return v;
}

static inline int f2i_sat(float x) {
    if (x != x) return 0;
    if (x >= 1.0e9f) return 1000000000;
    if (x <= -1.0e9f) return -1000000000;
    return (int)x;
}
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

typedef struct {
    float mw[3];   /* world mean */
    float qr[4];   /* composed, un-normalised quaternion */
    float qnorm;   /* |qr| */
    float qn[4];   /* normalised */
    float s[3];    /* exp(log scale) */
    float Rg[9];   /* rotation of the Gaussian */
    float S[6];    /* Sigma3D upper triangle 00 01 02 11 12 22 */
    float pv[3];   /* view-space mean */
    float tx, ty;  /* fov-clamped */
    int clampx, clampy; /* -1/0/+1 */
    float T[6];    /* 2x3 J*W */
    float a, b, c; /* blurred cov2d */
    float det;
    int visible;
    float xy[2];
    float conic[3];
    int radius;
    int tmin[2], tmax[2];
} proj_state;

/* Exact-section forward for one Gaussian.  Returns st->visible. */
static int project_one(const oracle_segment* sg, int i, const oracle_camera* cam, proj_state* st) {
    const float* m = sg->means + 3 * (size_t)i;
    const float* ls = sg->scales + 3 * (size_t)i;
    const float* q = sg->quats + 4 * (size_t)i;
    const float* W = cam->viewmat;
    st->visible = 0;
    st->radius = 0;
    st->xy[0] = st->xy[1] = 0.f;
    st->conic[0] = st->conic[1] = st->conic[2] = 0.f;
    st->tmin[0] = st->tmin[1] = st->tmax[0] = st->tmax[1] = 0;
    /* object2world_gs (sgn_splatfacto_scene_graph.py:404-417) */
    if (sg->has_pose) {
        const float* R = sg->R;
        st->mw[0] = ((R[0] * m[0] + R[1] * m[1]) + R[2] * m[2]) + sg->t[0];
        st->mw[1] = ((R[3] * m[0] + R[4] * m[1]) + R[5] * m[2]) + sg->t[1];
        st->mw[2] = ((R[6] * m[0] + R[7] * m[1]) + R[8] * m[2]) + sg->t[2];
        const float aw = sg->q[0], ax = sg->q[1], ay = sg->q[2], az = sg->q[3];
        const float bw = q[0], bx = q[1], by = q[2], bz = q[3];
        st->qr[0] = ((aw * bw - ax * bx) - ay * by) - az * bz;
        st->qr[1] = ((aw * bx + ax * bw) + ay * bz) - az * by;
        st->qr[2] = ((aw * by - ax * bz) + ay * bw) + az * bx;
        st->qr[3] = ((aw * bz + ax * by) - ay * bx) + az * bw;
    } else {
        st->mw[0] = m[0]; st->mw[1] = m[1]; st->mw[2] = m[2];
        st->qr[0] = q[0]; st->qr[1] = q[1]; st->qr[2] = q[2]; st->qr[3] = q[3];
    }
    /* view transform (gsplat clip_near_plane / transform_4x3) */
    st->pv[0] = ((W[0] * st->mw[0] + W[1] * st->mw[1]) + W[2] * st->mw[2]) + W[3];
    st->pv[1] = ((W[4] * st->mw[0] + W[5] * st->mw[1]) + W[6] * st->mw[2]) + W[7];
    st->pv[2] = ((W[8] * st->mw[0] + W[9] * st->mw[1]) + W[10] * st->mw[2]) + W[11];
    if (st->pv[2] <= cam->clip_thresh) return 0;
    /* quats / |quats| (sgn_splatfacto.py:864) */
    {
        float n2 = ((st->qr[0] * st->qr[0] + st->qr[1] * st->qr[1]) + st->qr[2] * st->qr[2]) + st->qr[3] * st->qr[3];
        st->qnorm = sqrtf(n2);
        for (int k = 0; k < 4; ++k) st->qn[k] = st->qr[k] / st->qnorm;
    }
    /* exp(scales) (sgn_splatfacto.py:857) */
    for (int k = 0; k < 3; ++k) st->s[k] = oracle_exp(ls[k]);
    {
        const float w = st->qn[0], x = st->qn[1], y = st->qn[2], z = st->qn[3];
        float* R = st->Rg;
        R[0] = 1.f - 2.f * (y * y + z * z);
        R[1] = 2.f * (x * y - w * z);
        R[2] = 2.f * (x * z + w * y);
        R[3] = 2.f * (x * y + w * z);
        R[4] = 1.f - 2.f * (x * x + z * z);
        R[5] = 2.f * (y * z - w * x);
        R[6] = 2.f * (x * z - w * y);
        R[7] = 2.f * (y * z + w * x);
        R[8] = 1.f - 2.f * (x * x + y * y);
        float M[9];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) M[3 * r + c] = R[3 * r + c] * st->s[c];
        st->S[0] = (M[0] * M[0] + M[1] * M[1]) + M[2] * M[2];
        st->S[1] = (M[0] * M[3] + M[1] * M[4]) + M[2] * M[5];
        st->S[2] = (M[0] * M[6] + M[1] * M[7]) + M[2] * M[8];
        st->S[3] = (M[3] * M[3] + M[4] * M[4]) + M[5] * M[5];
        st->S[4] = (M[3] * M[6] + M[4] * M[7]) + M[5] * M[8];
        st->S[5] = (M[6] * M[6] + M[7] * M[7]) + M[8] * M[8];
    }
    /* EWA projection (Appendix A.2) */
    {
        const float z = st->pv[2];
        const float rz = 1.f / z;
        const float rz2 = rz * rz;
        float ux = st->pv[0] / z, uy = st->pv[1] / z;
        st->clampx = 0; st->clampy = 0;
        if (ux > cam->limx) { ux = cam->limx; st->clampx = 1; }
        else if (ux < -cam->limx) { ux = -cam->limx; st->clampx = -1; }
        if (uy > cam->limy) { uy = cam->limy; st->clampy = 1; }
        else if (uy < -cam->limy) { uy = -cam->limy; st->clampy = -1; }
        st->tx = z * ux;
        st->ty = z * uy;
        const float J00 = cam->fx * rz, J11 = cam->fy * rz;
        const float J02 = -((cam->fx * st->tx) * rz2);
        const float J12 = -((cam->fy * st->ty) * rz2);
        float* T = st->T;
        for (int c = 0; c < 3; ++c) {
            T[c] = J00 * W[c] + J02 * W[8 + c];
            T[3 + c] = J11 * W[4 + c] + J12 * W[8 + c];
        }
        const float* S = st->S;
        /* TS = T * Sigma (2x3) */
        float TS[6];
        TS[0] = (T[0] * S[0] + T[1] * S[1]) + T[2] * S[2];
        TS[1] = (T[0] * S[1] + T[1] * S[3]) + T[2] * S[4];
        TS[2] = (T[0] * S[2] + T[1] * S[4]) + T[2] * S[5];
        TS[3] = (T[3] * S[0] + T[4] * S[1]) + T[5] * S[2];
        TS[4] = (T[3] * S[1] + T[4] * S[3]) + T[5] * S[4];
        TS[5] = (T[3] * S[2] + T[4] * S[4]) + T[5] * S[5];
        const float c00 = (TS[0] * T[0] + TS[1] * T[1]) + TS[2] * T[2];
        const float c01 = (TS[0] * T[3] + TS[1] * T[4]) + TS[2] * T[5];
        const float c11 = (TS[3] * T[3] + TS[4] * T[4]) + TS[5] * T[5];
        st->a = c00 + 0.3f;
        st->b = c01;
        st->c = c11 + 0.3f;
    }
    /* conic + radius (Appendix A.3) */
    st->det = st->a * st->c - st->b * st->b;
    if (st->det == 0.f) return 0;
    {
        const float inv = 1.f / st->det;
        st->conic[0] = st->c * inv;
        st->conic[1] = (-st->b) * inv;
        st->conic[2] = st->a * inv;
        const float bm = 0.5f * (st->a + st->c);
        const float disc = sqrtf(fmaxf(0.1f, bm * bm - st->det));
        const float v1 = bm + disc, v2 = bm - disc;
        st->radius = f2i_sat(ceilf(3.f * sqrtf(fmaxf(v1, v2))));
    }
    /* centre + tile AABB (Appendix A.4) */
    float cxp, cyp;
    {
        const float rw = 1.f / (st->pv[2] + 1e-6f);
        cxp = (st->pv[0] * rw) * cam->fx + cam->cx;
        cyp = (st->pv[1] * rw) * cam->fy + cam->cy;
        const float bw = (float)cam->block_width;
        const int tiles_x = (cam->width + cam->block_width - 1) / cam->block_width;
        const int tiles_y = (cam->height + cam->block_width - 1) / cam->block_width;
        const float tcx = cxp / bw, tcy = cyp / bw, tr = (float)st->radius / bw;
        st->tmin[0] = imin(imax(0, f2i_sat(tcx - tr)), tiles_x);
        st->tmax[0] = imin(imax(0, f2i_sat((tcx + tr) + 1.f)), tiles_x);
        st->tmin[1] = imin(imax(0, f2i_sat(tcy - tr)), tiles_y);
        st->tmax[1] = imin(imax(0, f2i_sat((tcy + tr) + 1.f)), tiles_y);
    }
    const int area = (st->tmax[0] - st->tmin[0]) * (st->tmax[1] - st->tmin[1]);
    if (area <= 0) { st->radius = 0; return 0; }
    st->xy[0] = cxp;
    st->xy[1] = cyp;
    st->visible = 1;
    return 1;
}

/* ------------------------------------------------------------------------------------------ */
/* SH (Appendix A.7)                                                                           */
/* ------------------------------------------------------------------------------------------ */
static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                               0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                               -0.5900435899266435f};

static void sh_basis(int deg, float x, float y, float z, float* Y /*16*/) {
    for (int k = 0; k < 16; ++k) Y[k] = 0.f;
    Y[0] = SH_C0;
    if (deg < 1) return;
    Y[1] = -SH_C1 * y; Y[2] = SH_C1 * z; Y[3] = -SH_C1 * x;
    if (deg < 2) return;
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    Y[4] = SH_C2[0] * xy; Y[5] = SH_C2[1] * yz; Y[6] = SH_C2[2] * (2.f * zz - xx - yy);
    Y[7] = SH_C2[3] * xz; Y[8] = SH_C2[4] * (xx - yy);
    if (deg < 3) return;
    Y[9] = SH_C3[0] * y * (3.f * xx - yy);
    Y[10] = SH_C3[1] * xy * z;
    Y[11] = SH_C3[2] * y * (4.f * zz - xx - yy);
    Y[12] = SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
    Y[13] = SH_C3[4] * x * (4.f * zz - xx - yy);
    Y[14] = SH_C3[5] * z * (xx - yy);
    Y[15] = SH_C3[6] * x * (xx - 3.f * yy);
}

static void view_dir(const float* mw, const float* cam_pos, float* d) {
    d[0] = mw[0] - cam_pos[0]; d[1] = mw[1] - cam_pos[1]; d[2] = mw[2] - cam_pos[2];
    const float n = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    d[0] /= n; d[1] /= n; d[2] /= n;
}

/* ------------------------------------------------------------------------------------------ */
/* 1. compose + project + SH + sigmoid                                                         */
/* ------------------------------------------------------------------------------------------ */
/* Outputs are dense over the concatenated row space [0, N).  Invisible rows: xys=0, depths=0,
 * radii=0, num_tiles_hit=0 (gsplat zero-initialises its outputs), conics as gsplat leaves them
 * (written before the tile-area test).  rgbs/opac are computed for every row (the reference
 * evaluates SH and sigmoid on all Gaussians). */
int sgn_oracle_project(const oracle_segment* segs, int nseg, const oracle_camera* cam,
                       float* xys, float* depths, int32_t* radii, float* conics,
                       int32_t* num_tiles_hit, float* rgbs, float* rgb_pre, float* opac,
                       int32_t* tile_bbox /* [N,4] xmin,ymin,xmax,ymax or NULL */) {
    const int K = (cam->sh_degree + 1) * (cam->sh_degree + 1);
    for (int s = 0; s < nseg; ++s) {
        const oracle_segment* sg = &segs[s];
#pragma omp parallel for schedule(static)
        for (int i = 0; i < sg->count; ++i) {
            const size_t g = (size_t)sg->row0 + i;
            proj_state st;
            const int vis = project_one(sg, i, cam, &st);
            xys[2 * g] = st.xy[0]; xys[2 * g + 1] = st.xy[1];
            depths[g] = vis ? st.pv[2] : 0.f;
            radii[g] = st.radius;
            conics[3 * g] = st.conic[0]; conics[3 * g + 1] = st.conic[1]; conics[3 * g + 2] = st.conic[2];
            num_tiles_hit[g] = vis ? (st.tmax[0] - st.tmin[0]) * (st.tmax[1] - st.tmin[1]) : 0;
            if (tile_bbox) {
                tile_bbox[4 * g] = st.tmin[0]; tile_bbox[4 * g + 1] = st.tmin[1];
                tile_bbox[4 * g + 2] = st.tmax[0]; tile_bbox[4 * g + 3] = st.tmax[1];
            }
            /* colour: Fourier DC (scene graph :239-247) + SH + 0.5, clamp min 0 (:939-940) */
            float c0[3] = {0.f, 0.f, 0.f};
            for (int f = 0; f < sg->F; ++f)
                for (int ch = 0; ch < 3; ++ch)
                    c0[ch] += sg->features_dc[((size_t)i * sg->F + f) * 3 + ch] * sg->idft[f];
            float out[3];
            if (cam->sh_degree > 0) {
                float d[3], Y[16];
                view_dir(st.mw, cam->cam_pos, d);
                sh_basis(cam->sh_degree_to_use, d[0], d[1], d[2], Y);
                const int Kuse = (cam->sh_degree_to_use + 1) * (cam->sh_degree_to_use + 1);
                for (int ch = 0; ch < 3; ++ch) {
                    float acc = Y[0] * c0[ch];
                    for (int k = 1; k < Kuse && k < K; ++k)
                        acc += Y[k] * sg->features_rest[((size_t)i * (K - 1) + (k - 1)) * 3 + ch];
                    out[ch] = acc + 0.5f;
                }
                for (int ch = 0; ch < 3; ++ch) {
                    if (rgb_pre) rgb_pre[3 * g + ch] = out[ch];
                    rgbs[3 * g + ch] = out[ch] > 0.f ? out[ch] : 0.f;
                }
            } else { /* sgn_splatfacto.py:942 : sigmoid(colors[:,0,:]) */
                for (int ch = 0; ch < 3; ++ch) {
                    if (rgb_pre) rgb_pre[3 * g + ch] = c0[ch];
                    rgbs[3 * g + ch] = 1.f / (1.f + expf(-c0[ch]));
                }
            }
            opac[g] = 1.f / (1.f + expf(-sg->opacities[i]));
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* 2. binning + sort (Appendix A.5)                                                            */
/* ------------------------------------------------------------------------------------------ */
static int cmp_u64(const void* a, const void* b) {
    const uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

/* Returns M.  sorted_ids must hold sum(num_tiles_hit) entries; tile_bins is [tiles,2].
 * Order == stable sort of (tile<<32 | depth_bits) with emission order (gaussian index, then
 * row-major tile) as the tie break == sort by (tile, depth_bits, gaussian index). */
int64_t sgn_oracle_bin_sort(int N, const float* xys, const float* depths, const int32_t* radii,
                            const int32_t* num_tiles_hit, int width, int height, int block_width,
                            int32_t* sorted_ids, int32_t* tile_bins) {
    const int tiles_x = (width + block_width - 1) / block_width;
    const int tiles_y = (height + block_width - 1) / block_width;
    const int ntiles = tiles_x * tiles_y;
    int64_t* count = (int64_t*)calloc((size_t)ntiles + 1, sizeof(int64_t));
    const float bw = (float)block_width;
    /* pass 1: recompute tile bbox exactly as map_gaussian_to_intersects does */
    int32_t* bbox = (int32_t*)malloc(sizeof(int32_t) * 4 * (size_t)N);
#pragma omp parallel for schedule(static)
    for (int g = 0; g < N; ++g) {
        bbox[4 * g] = bbox[4 * g + 1] = bbox[4 * g + 2] = bbox[4 * g + 3] = 0;
        if (radii[g] <= 0) continue;
        const float tcx = xys[2 * g] / bw, tcy = xys[2 * g + 1] / bw, tr = (float)radii[g] / bw;
        bbox[4 * g] = imin(imax(0, f2i_sat(tcx - tr)), tiles_x);
        bbox[4 * g + 2] = imin(imax(0, f2i_sat((tcx + tr) + 1.f)), tiles_x);
        bbox[4 * g + 1] = imin(imax(0, f2i_sat(tcy - tr)), tiles_y);
        bbox[4 * g + 3] = imin(imax(0, f2i_sat((tcy + tr) + 1.f)), tiles_y);
    }
    int64_t M = 0;
    for (int g = 0; g < N; ++g) {
        if (radii[g] <= 0) continue;
        for (int ty = bbox[4 * g + 1]; ty < bbox[4 * g + 3]; ++ty)
            for (int tx = bbox[4 * g]; tx < bbox[4 * g + 2]; ++tx) count[ty * tiles_x + tx + 1]++;
        M += num_tiles_hit[g];
    }
    for (int t = 0; t < ntiles; ++t) count[t + 1] += count[t];
    int64_t total = count[ntiles];
    if (total != M) { free(count); free(bbox); return -1; } /* num_tiles_hit inconsistent */
    uint64_t* keys = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(total > 0 ? total : 1));
    int64_t* cursor = (int64_t*)malloc(sizeof(int64_t) * (size_t)ntiles);
    memcpy(cursor, count, sizeof(int64_t) * (size_t)ntiles);
    for (int g = 0; g < N; ++g) {
        if (radii[g] <= 0) continue;
        uint32_t dbits;
        memcpy(&dbits, &depths[g], 4);
        const uint64_t k = ((uint64_t)dbits << 32) | (uint32_t)g;
        for (int ty = bbox[4 * g + 1]; ty < bbox[4 * g + 3]; ++ty)
            for (int tx = bbox[4 * g]; tx < bbox[4 * g + 2]; ++tx) keys[cursor[ty * tiles_x + tx]++] = k;
    }
#pragma omp parallel for schedule(dynamic, 16)
    for (int t = 0; t < ntiles; ++t) {
        const int64_t b = count[t], e = count[t + 1];
        /* depth bits of a positive float order like the float; gsplat casts to int32 then int64 */
        if (e - b > 1) qsort(keys + b, (size_t)(e - b), sizeof(uint64_t), cmp_u64);
        for (int64_t k = b; k < e; ++k) sorted_ids[k] = (int32_t)(keys[k] & 0xffffffffu);
        tile_bins[2 * t] = (int32_t)b;
        tile_bins[2 * t + 1] = (int32_t)e;
    }
    free(keys); free(cursor); free(count); free(bbox);
    return total;
}

/* ------------------------------------------------------------------------------------------ */
/* 3. blend forward (Appendix A.6)                                                             */
/* ------------------------------------------------------------------------------------------ */
/* colors[N,C] ; out_img[H,W,C] ; final_T[H,W] ; final_idx[H,W] ; fragile[H,W] (optional).
 * cls_filter < 0: all Gaussians; otherwise only those with gauss_cls[g]==cls_filter (this is the
 * objects-only / background-only re-render of get_submodel_output, scene graph :255-303: the
 * subset keeps its relative order, so filtering the merged sorted list is equivalent).
 * fragile: set when a skip/stop decision is within a relative margin of flipping -- such pixels
 * are excluded from max-abs parity checks (documented in DESIGN.md). */
int sgn_oracle_blend_fwd(int width, int height, int block_width, int C,
                         const int32_t* sorted_ids, const int32_t* tile_bins,
                         const float* xys, const float* conics, const float* colors,
                         const float* opac, const float* background, float alpha_clamp,
                         const int32_t* gauss_cls, int cls_filter,
                         float* out_img, float* final_T, int32_t* final_idx, uint8_t* fragile,
                         float margin) {
    const int tiles_x = (width + block_width - 1) / block_width;
    const int tiles_y = (height + block_width - 1) / block_width;
#pragma omp parallel for schedule(dynamic, 4)
    for (int t = 0; t < tiles_x * tiles_y; ++t) {
        const int ty = t / tiles_x, tx = t % tiles_x;
        const int b = tile_bins[2 * t], e = tile_bins[2 * t + 1];
        float pix[64];
        for (int ly = 0; ly < block_width; ++ly) {
            const int i = ty * block_width + ly;
            if (i >= height) break;
            for (int lx = 0; lx < block_width; ++lx) {
                const int j = tx * block_width + lx;
                if (j >= width) break;
                const float px = (float)j + 0.5f, py = (float)i + 0.5f;
                float T = 1.f;
                int cur = 0, frag = 0;
                for (int c = 0; c < C; ++c) pix[c] = 0.f;
                for (int k = b; k < e; ++k) {
                    const int g = sorted_ids[k];
                    if (cls_filter >= 0 && gauss_cls[g] != cls_filter) continue;
                    const float dx = xys[2 * g] - px, dy = xys[2 * g + 1] - py;
                    const float ca = conics[3 * g], cb = conics[3 * g + 1], cc = conics[3 * g + 2];
                    const float sigma = 0.5f * (ca * dx * dx + cc * dy * dy) + cb * dx * dy;
                    const float raw = opac[g] * expf(-sigma);
                    const float alpha = fminf(alpha_clamp, raw);
                    if (fabsf(raw * 255.f - 1.f) < margin || fabsf(sigma) < 1e-7f) frag = 1;
                    if (sigma < 0.f || alpha < 1.f / 255.f) continue;
                    const float nT = T * (1.f - alpha);
                    if (fabsf(nT - 1e-4f) < 1e-4f * margin) frag = 1;
                    if (nT <= 1e-4f) break;
                    const float vis = alpha * T;
                    for (int c = 0; c < C; ++c) pix[c] += colors[(size_t)g * C + c] * vis;
                    T = nT;
                    cur = k;
                }
                const size_t p = (size_t)i * width + j;
                final_T[p] = T;
                final_idx[p] = cur;
                if (fragile) fragile[p] = (uint8_t)frag;
                for (int c = 0; c < C; ++c) out_img[p * C + c] = pix[c] + T * background[c];
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* 4. blend backward (Appendix A.6) -- per-pixel back-to-front, as gsplat rasterize_backward    */
/* ------------------------------------------------------------------------------------------ */
/* v_xy[N,2], v_conic[N,3], v_colors[N,C], v_opac[N] are ACCUMULATED INTO (caller zeroes). */
int sgn_oracle_blend_bwd(int width, int height, int block_width, int C,
                         const int32_t* sorted_ids, const int32_t* tile_bins,
                         const float* xys, const float* conics, const float* colors,
                         const float* opac, const float* background, float alpha_clamp_bwd,
                         const int32_t* gauss_cls, int cls_filter,
                         const float* final_T, const int32_t* final_idx,
                         const float* v_out_img, const float* v_out_alpha,
                         double* v_xy, double* v_conic, double* v_colors, double* v_opac) {
    const int tiles_x = (width + block_width - 1) / block_width;
    const int tiles_y = (height + block_width - 1) / block_width;
#pragma omp parallel for schedule(dynamic, 4)
    for (int t = 0; t < tiles_x * tiles_y; ++t) {
        const int ty = t / tiles_x, tx = t % tiles_x;
        const int b = tile_bins[2 * t], e = tile_bins[2 * t + 1];
        if (e <= b) continue;
        const int len = e - b;
        const int stride = 6 + C;
        float* acc = (float*)calloc((size_t)len * stride, sizeof(float));
        float buffer[64];
        for (int ly = 0; ly < block_width; ++ly) {
            const int i = ty * block_width + ly;
            if (i >= height) break;
            for (int lx = 0; lx < block_width; ++lx) {
                const int j = tx * block_width + lx;
                if (j >= width) break;
                const float px = (float)j + 0.5f, py = (float)i + 0.5f;
                const size_t p = (size_t)i * width + j;
                const float T_final = final_T[p];
                float T = T_final;
                const int bin_final = final_idx[p];
                const float* vo = v_out_img + p * C;
                const float voa = v_out_alpha ? v_out_alpha[p] : 0.f;
                for (int c = 0; c < C; ++c) buffer[c] = 0.f;
                for (int k = bin_final; k >= b; --k) {
                    const int g = sorted_ids[k];
                    if (cls_filter >= 0 && gauss_cls[g] != cls_filter) continue;
                    const float dx = xys[2 * g] - px, dy = xys[2 * g + 1] - py;
                    const float ca = conics[3 * g], cb = conics[3 * g + 1], cc = conics[3 * g + 2];
                    const float sigma = 0.5f * (ca * dx * dx + cc * dy * dy) + cb * dx * dy;
                    const float vis = expf(-sigma);
                    const float o = opac[g];
                    const float alpha = fminf(alpha_clamp_bwd, o * vis);
                    if (sigma < 0.f || alpha < 1.f / 255.f) continue;
                    const float ra = 1.f / (1.f - alpha);
                    T *= ra;
                    const float fac = alpha * T;
                    float v_alpha = 0.f;
                    float* a = acc + (size_t)(k - b) * stride;
                    for (int c = 0; c < C; ++c) {
                        const float col = colors[(size_t)g * C + c];
                        a[6 + c] += fac * vo[c];
                        v_alpha += (col * T - buffer[c] * ra) * vo[c];
                        v_alpha += -T_final * ra * background[c] * vo[c];
                        buffer[c] += col * fac;
                    }
                    v_alpha += T_final * ra * voa;
                    const float v_sigma = -o * vis * v_alpha;
                    a[0] += v_sigma * (ca * dx + cb * dy);
                    a[1] += v_sigma * (cb * dx + cc * dy);
                    a[2] += 0.5f * v_sigma * dx * dx;
                    a[3] += v_sigma * dx * dy;
                    a[4] += 0.5f * v_sigma * dy * dy;
                    a[5] += vis * v_alpha;
                }
            }
        }
        for (int k = 0; k < len; ++k) {
            const int g = sorted_ids[b + k];
            const float* a = acc + (size_t)k * stride;
#pragma omp atomic
            v_xy[2 * (size_t)g] += a[0];
#pragma omp atomic
            v_xy[2 * (size_t)g + 1] += a[1];
#pragma omp atomic
            v_conic[3 * (size_t)g] += a[2];
#pragma omp atomic
            v_conic[3 * (size_t)g + 1] += a[3];
#pragma omp atomic
            v_conic[3 * (size_t)g + 2] += a[4];
#pragma omp atomic
            v_opac[g] += a[5];
            for (int c = 0; c < C; ++c) {
#pragma omp atomic
                v_colors[(size_t)g * C + c] += a[6 + c];
            }
        }
        free(acc);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* 5. SH + project + compose backward (Appendix A.7-A.8 and autograd of the reference glue)    */
/* ------------------------------------------------------------------------------------------ */
/* Inputs: per-Gaussian cotangents v_xy[N,2], v_depth[N] (may be NULL), v_conic[N,3], v_rgb[N,3]
 * (w.r.t. the clamped colour), v_opac[N] (w.r.t. sigmoid output).  rgb_pre from the forward.
 * Outputs (dense, per segment, same layout as the parameters): g_means, g_scales, g_quats,
 * g_dc, g_rest, g_opac -- arrays of nseg pointers.  Rows with radii==0 get zero geometry grads
 * (gsplat's project backward returns early for radii<=0) but still get colour/opacity grads
 * (which are zero anyway because the rasterizer never touched them). */
int sgn_oracle_project_bwd(const oracle_segment* segs, int nseg, const oracle_camera* cam,
                           const int32_t* radii, const float* rgb_pre,
                           const float* v_xy, const float* v_depth, const float* v_conic,
                           const float* v_rgb, const float* v_opac,
                           float** g_means, float** g_scales, float** g_quats, float** g_dc,
                           float** g_rest, float** g_opac) {
    const int K = (cam->sh_degree + 1) * (cam->sh_degree + 1);
    const float* W = cam->viewmat;
    for (int s = 0; s < nseg; ++s) {
        const oracle_segment* sg = &segs[s];
#pragma omp parallel for schedule(static)
        for (int i = 0; i < sg->count; ++i) {
            const size_t g = (size_t)sg->row0 + i;
            float* gm = g_means[s] + 3 * (size_t)i;
            float* gs = g_scales[s] + 3 * (size_t)i;
            float* gq = g_quats[s] + 4 * (size_t)i;
            float* gd = g_dc[s] + (size_t)i * sg->F * 3;
            float* gr = g_rest[s] + (size_t)i * (K - 1) * 3;
            for (int k = 0; k < 3; ++k) { gm[k] = 0.f; gs[k] = 0.f; }
            for (int k = 0; k < 4; ++k) gq[k] = 0.f;
            /* opacity: sigmoid backward */
            {
                const float o = 1.f / (1.f + expf(-sg->opacities[i]));
                g_opac[s][i] = v_opac[g] * o * (1.f - o);
            }
            proj_state st;
            const int vis = project_one(sg, i, cam, &st);
            /* colour */
            {
                float vc[3];
                if (cam->sh_degree > 0) {
                    float d[3], Y[16];
                    if (!vis) { /* project_one bails before mw is final only when clipped: mw is always set */ }
                    view_dir(st.mw, cam->cam_pos, d);
                    sh_basis(cam->sh_degree_to_use, d[0], d[1], d[2], Y);
                    const int Kuse = (cam->sh_degree_to_use + 1) * (cam->sh_degree_to_use + 1);
                    for (int ch = 0; ch < 3; ++ch) vc[ch] = rgb_pre[3 * g + ch] >= 0.f ? v_rgb[3 * g + ch] : 0.f;
                    for (int k = 1; k < K; ++k)
                        for (int ch = 0; ch < 3; ++ch) gr[(size_t)(k - 1) * 3 + ch] = (k < Kuse) ? Y[k] * vc[ch] : 0.f;
                    for (int ch = 0; ch < 3; ++ch) vc[ch] *= Y[0];
                } else {
                    for (int ch = 0; ch < 3; ++ch) {
                        const float sgm = 1.f / (1.f + expf(-rgb_pre[3 * g + ch]));
                        vc[ch] = v_rgb[3 * g + ch] * sgm * (1.f - sgm);
                    }
                    for (int k = 1; k < K; ++k)
                        for (int ch = 0; ch < 3; ++ch) gr[(size_t)(k - 1) * 3 + ch] = 0.f;
                }
                for (int f = 0; f < sg->F; ++f)
                    for (int ch = 0; ch < 3; ++ch) gd[(size_t)f * 3 + ch] = sg->idft[f] * vc[ch];
            }
            if (!vis || radii[g] <= 0) continue;
            /* ---- geometry ---- */
            const float fx = cam->fx, fy = cam->fy;
            float vpv[3];
            {
                const float rw = 1.f / (st.pv[2] + 1e-6f);
                const float vx = fx * v_xy[2 * g], vy = fy * v_xy[2 * g + 1];
                vpv[0] = vx * rw;
                vpv[1] = vy * rw;
                vpv[2] = -(vx * st.pv[0] + vy * st.pv[1]) * rw * rw;
                if (v_depth) vpv[2] += v_depth[g];
            }
            /* conic -> cov2d : v_Sigma = -X G X */
            float vA, vB, vC; /* d/d a, d/d b (single parameter), d/d c */
            {
                const float X0 = st.conic[0], X1 = st.conic[1], X2 = st.conic[2];
                const float G0 = v_conic[3 * g], G1 = 0.5f * v_conic[3 * g + 1], G2 = v_conic[3 * g + 2];
                /* XG */
                const float a00 = X0 * G0 + X1 * G1, a01 = X0 * G1 + X1 * G2;
                const float a10 = X1 * G0 + X2 * G1, a11 = X1 * G1 + X2 * G2;
                const float s00 = -(a00 * X0 + a01 * X1);
                const float s01 = -(a00 * X1 + a01 * X2);
                const float s10 = -(a10 * X0 + a11 * X1);
                const float s11 = -(a10 * X1 + a11 * X2);
                vA = s00; vB = s01 + s10; vC = s11;
            }
            /* cov = T S T^T ; G2 = [[vA, vB/2],[vB/2, vC]] */
            const float g00 = vA, g01 = 0.5f * vB, g11 = vC;
            const float* T = st.T;
            float Sf[9] = {st.S[0], st.S[1], st.S[2], st.S[1], st.S[3], st.S[4], st.S[2], st.S[4], st.S[5]};
            /* GT = G2 * T (2x3) */
            float GT[6];
            for (int c = 0; c < 3; ++c) {
                GT[c] = g00 * T[c] + g01 * T[3 + c];
                GT[3 + c] = g01 * T[c] + g11 * T[3 + c];
            }
            /* v_S = T^T G2 T (3x3) */
            float vS[9];
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) vS[3 * r + c] = T[r] * GT[c] + T[3 + r] * GT[3 + c];
            /* v_T = 2 * G2 T S (2x3) */
            float vT[6];
            for (int r = 0; r < 2; ++r)
                for (int c = 0; c < 3; ++c)
                    vT[3 * r + c] = 2.f * (GT[3 * r] * Sf[c] + GT[3 * r + 1] * Sf[3 + c] + GT[3 * r + 2] * Sf[6 + c]);
            /* v_J = v_T W^T (only entries 00, 02, 11, 12 matter) */
            const float vJ00 = vT[0] * W[0] + vT[1] * W[1] + vT[2] * W[2];
            const float vJ02 = vT[0] * W[8] + vT[1] * W[9] + vT[2] * W[10];
            const float vJ11 = vT[3] * W[4] + vT[4] * W[5] + vT[5] * W[6];
            const float vJ12 = vT[3] * W[8] + vT[4] * W[9] + vT[5] * W[10];
            {
                const float rz = 1.f / st.pv[2], rz2 = rz * rz, rz3 = rz2 * rz;
                const float vtx = -fx * rz2 * vJ02;
                const float vty = -fy * rz2 * vJ12;
                const float vtz = -fx * rz2 * vJ00 - fy * rz2 * vJ11 + 2.f * fx * st.tx * rz3 * vJ02 +
                                  2.f * fy * st.ty * rz3 * vJ12;
                /* fov clamp: tx = z*clamp(x/z) */
                if (st.clampx == 0) vpv[0] += vtx; else vpv[2] += (st.clampx > 0 ? cam->limx : -cam->limx) * vtx;
                if (st.clampy == 0) vpv[1] += vty; else vpv[2] += (st.clampy > 0 ? cam->limy : -cam->limy) * vty;
                vpv[2] += vtz;
            }
            /* v_mw = W^T v_pv */
            float vmw[3];
            for (int c = 0; c < 3; ++c) vmw[c] = W[c] * vpv[0] + W[4 + c] * vpv[1] + W[8 + c] * vpv[2];
            if (sg->has_pose) {
                const float* R = sg->R;
                for (int c = 0; c < 3; ++c) gm[c] = R[c] * vmw[0] + R[3 + c] * vmw[1] + R[6 + c] * vmw[2];
            } else {
                for (int c = 0; c < 3; ++c) gm[c] = vmw[c];
            }
            /* Sigma = M M^T ; v_M = 2 v_S M */
            float M[9], vM[9];
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) M[3 * r + c] = st.Rg[3 * r + c] * st.s[c];
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c)
                    vM[3 * r + c] = 2.f * (vS[3 * r] * M[c] + vS[3 * r + 1] * M[3 + c] + vS[3 * r + 2] * M[6 + c]);
            float vR[9];
            for (int c = 0; c < 3; ++c) {
                const float vs = st.Rg[c] * vM[c] + st.Rg[3 + c] * vM[3 + c] + st.Rg[6 + c] * vM[6 + c];
                gs[c] = vs * st.s[c]; /* through exp */
                for (int r = 0; r < 3; ++r) vR[3 * r + c] = vM[3 * r + c] * st.s[c];
            }
            /* quat_to_rotmat vjp */
            float vqn[4];
            {
                const float w = st.qn[0], x = st.qn[1], y = st.qn[2], z = st.qn[3];
                vqn[0] = 2.f * (x * (vR[7] - vR[5]) + y * (vR[2] - vR[6]) + z * (vR[3] - vR[1]));
                vqn[1] = 2.f * (-2.f * x * (vR[4] + vR[8]) + y * (vR[1] + vR[3]) + z * (vR[2] + vR[6]) + w * (vR[7] - vR[5]));
                vqn[2] = 2.f * (x * (vR[1] + vR[3]) - 2.f * y * (vR[0] + vR[8]) + z * (vR[5] + vR[7]) + w * (vR[2] - vR[6]));
                vqn[3] = 2.f * (x * (vR[2] + vR[6]) + y * (vR[5] + vR[7]) - 2.f * z * (vR[0] + vR[4]) + w * (vR[3] - vR[1]));
            }
            /* normalisation backward */
            float vqr[4];
            {
                const float dot = vqn[0] * st.qn[0] + vqn[1] * st.qn[1] + vqn[2] * st.qn[2] + vqn[3] * st.qn[3];
                for (int k = 0; k < 4; ++k) vqr[k] = (vqn[k] - st.qn[k] * dot) / st.qnorm;
            }
            if (sg->has_pose) {
                const float aw = sg->q[0], ax = sg->q[1], ay = sg->q[2], az = sg->q[3];
                gq[0] = aw * vqr[0] + ax * vqr[1] + ay * vqr[2] + az * vqr[3];
                gq[1] = -ax * vqr[0] + aw * vqr[1] + az * vqr[2] - ay * vqr[3];
                gq[2] = -ay * vqr[0] - az * vqr[1] + aw * vqr[2] + ax * vqr[3];
                gq[3] = -az * vqr[0] + ay * vqr[1] - ax * vqr[2] + aw * vqr[3];
            } else {
                for (int k = 0; k < 4; ++k) gq[k] = vqr[k];
            }
        }
    }
    return 0;
}

int sgn_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void sgn_oracle_set_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------------------------ */
/* For every entry of the (gsplat AABB) per-tile lists: can ANY pixel centre of that tile accept */
/* it (sigma >= 0 and alpha >= 1/255)?  Used to check that the product's exact tile culling only */
/* drops entries that are no-ops for every pixel.                                               */
/* ------------------------------------------------------------------------------------------ */
int sgn_oracle_entry_any_valid(int width, int height, int block_width, const int32_t* sorted_ids,
                               const int32_t* tile_bins, const float* xys, const float* conics,
                               const float* opac, uint8_t* any_valid) {
    const int tiles_x = (width + block_width - 1) / block_width;
    const int tiles_y = (height + block_width - 1) / block_width;
#pragma omp parallel for schedule(dynamic, 4)
    for (int t = 0; t < tiles_x * tiles_y; ++t) {
        const int ty = t / tiles_x, tx = t % tiles_x;
        for (int k = tile_bins[2 * t]; k < tile_bins[2 * t + 1]; ++k) {
            const int g = sorted_ids[k];
            const float ca = conics[3 * g], cb = conics[3 * g + 1], cc = conics[3 * g + 2];
            int any = 0;
            for (int ly = 0; ly < block_width && !any; ++ly) {
                const int i = ty * block_width + ly;
                if (i >= height) break;
                for (int lx = 0; lx < block_width; ++lx) {
                    const int j = tx * block_width + lx;
                    if (j >= width) break;
                    const float dx = xys[2 * g] - ((float)j + 0.5f), dy = xys[2 * g + 1] - ((float)i + 0.5f);
                    const float sigma = 0.5f * (ca * dx * dx + cc * dy * dy) + cb * dx * dy;
                    if (sigma >= 0.f && opac[g] * expf(-sigma) >= 1.f / 255.f) { any = 1; break; }
                }
            }
            any_valid[k] = (uint8_t)any;
        }
    }
    return 0;
}
