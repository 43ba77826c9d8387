// Exact-section projection math shared by the forward and backward per-Gaussian kernels.
//
// Arithmetic contract: this translation unit is compiled with --fmad=false, default -prec-div /
// -prec-sqrt (IEEE), so every * + - / sqrt below is individually rounded, in the order written.
// tests/ compare the integer outputs (radii, tile AABB, num_tiles_hit, sort order) BIT-EXACTLY
// against the CPU oracle, which states the same sequence independently.
//
// Semantics: gsplat 0.1.x project_gaussians preceded by the reference's
// compose / pre-ops (street_gaussians_ns/sgn_splatfacto_scene_graph.py:404-417,
// street_gaussians_ns/sgn_splatfacto.py:857,864).
#pragma once
#include "sgn_common.cuh"

struct SgnProj {
    float mw[3];
    float qr[4];
    float qnorm;
    float qn[4];
    float s[3];
    float Rg[9];
    float S[6];
    float pv[3];
    float tx, ty;
    int clampx, clampy;
    float T[6];
    float a, b, c;
    float conic[3];
    float xy[2];
    int radius;
    int tmin[2], tmax[2];
    bool visible;
};

// exp() as a fixed sequence of IEEE operations (same constants as oracle/sgn_oracle.c).
__device__ __forceinline__ float sgn_expf_exact(float x) {
    x = fminf(fmaxf(x, -80.0f), 80.0f);
    const float n = rintf(x * 1.44269504f);
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

__device__ __forceinline__ int sgn_f2i_sat(float x) {
    if (x != x) return 0;
    if (x >= 1.0e9f) return 1000000000;
    if (x <= -1.0e9f) return -1000000000;
    return (int)x;
}

// Returns st.visible.  `m`, `ls`, `q` are this Gaussian's raw parameters.
// log_scales: `ls` holds log-scales (the model's parameters) -> exp is applied here; otherwise `ls` holds
// activated scales (gsplat's project_gaussians argument) multiplied by glob_scale.
__device__ __forceinline__ bool sgn_project_exact(const sgn_segment& sg, const sgn_camera& cam, const float m[3],
                                                  const float ls[3], const float q[4], SgnProj& st,
                                                  const bool log_scales = true, const float glob_scale = 1.f) {
    const float* W = cam.viewmat;
    st.visible = false;
    st.radius = 0;
    st.xy[0] = st.xy[1] = 0.f;
    st.conic[0] = st.conic[1] = st.conic[2] = 0.f;
    st.tmin[0] = st.tmin[1] = st.tmax[0] = st.tmax[1] = 0;
    st.clampx = st.clampy = 0;
    if (sg.has_pose) {
        const float* R = sg.R;
        st.mw[0] = ((R[0] * m[0] + R[1] * m[1]) + R[2] * m[2]) + sg.t[0];
        st.mw[1] = ((R[3] * m[0] + R[4] * m[1]) + R[5] * m[2]) + sg.t[1];
        st.mw[2] = ((R[6] * m[0] + R[7] * m[1]) + R[8] * m[2]) + sg.t[2];
        const float aw = sg.q[0], ax = sg.q[1], ay = sg.q[2], az = sg.q[3];
        const float bw = q[0], bx = q[1], by = q[2], bz = q[3];
        st.qr[0] = ((aw * bw - ax * bx) - ay * by) - az * bz;
        st.qr[1] = ((aw * bx + ax * bw) + ay * bz) - az * by;
        st.qr[2] = ((aw * by - ax * bz) + ay * bw) + az * bx;
        st.qr[3] = ((aw * bz + ax * by) - ay * bx) + az * bw;
    } else {
        st.mw[0] = m[0]; st.mw[1] = m[1]; st.mw[2] = m[2];
        st.qr[0] = q[0]; st.qr[1] = q[1]; st.qr[2] = q[2]; st.qr[3] = q[3];
    }
    st.pv[0] = ((W[0] * st.mw[0] + W[1] * st.mw[1]) + W[2] * st.mw[2]) + W[3];
    st.pv[1] = ((W[4] * st.mw[0] + W[5] * st.mw[1]) + W[6] * st.mw[2]) + W[7];
    st.pv[2] = ((W[8] * st.mw[0] + W[9] * st.mw[1]) + W[10] * st.mw[2]) + W[11];
    if (st.pv[2] <= cam.clip_thresh) return false;
    {
        const float n2 = ((st.qr[0] * st.qr[0] + st.qr[1] * st.qr[1]) + st.qr[2] * st.qr[2]) + st.qr[3] * st.qr[3];
        st.qnorm = sqrtf(n2);
#pragma unroll
        for (int k = 0; k < 4; ++k) st.qn[k] = st.qr[k] / st.qnorm;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) st.s[k] = log_scales ? sgn_expf_exact(ls[k]) : ls[k] * glob_scale;
    {
        const float w = st.qn[0], x = st.qn[1], y = st.qn[2], z = st.qn[3];
        float* R = st.Rg;
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
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) M[3 * r + c] = R[3 * r + c] * st.s[c];
        st.S[0] = (M[0] * M[0] + M[1] * M[1]) + M[2] * M[2];
        st.S[1] = (M[0] * M[3] + M[1] * M[4]) + M[2] * M[5];
        st.S[2] = (M[0] * M[6] + M[1] * M[7]) + M[2] * M[8];
        st.S[3] = (M[3] * M[3] + M[4] * M[4]) + M[5] * M[5];
        st.S[4] = (M[3] * M[6] + M[4] * M[7]) + M[5] * M[8];
        st.S[5] = (M[6] * M[6] + M[7] * M[7]) + M[8] * M[8];
    }
    {
        const float z = st.pv[2];
        const float rz = 1.f / z;
        const float rz2 = rz * rz;
        float ux = st.pv[0] / z, uy = st.pv[1] / z;
        if (ux > cam.limx) { ux = cam.limx; st.clampx = 1; }
        else if (ux < -cam.limx) { ux = -cam.limx; st.clampx = -1; }
        if (uy > cam.limy) { uy = cam.limy; st.clampy = 1; }
        else if (uy < -cam.limy) { uy = -cam.limy; st.clampy = -1; }
        st.tx = z * ux;
        st.ty = z * uy;
        const float J00 = cam.fx * rz, J11 = cam.fy * rz;
        const float J02 = -((cam.fx * st.tx) * rz2);
        const float J12 = -((cam.fy * st.ty) * rz2);
        float* T = st.T;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            T[c] = J00 * W[c] + J02 * W[8 + c];
            T[3 + c] = J11 * W[4 + c] + J12 * W[8 + c];
        }
        const float* S = st.S;
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
        st.a = c00 + 0.3f;
        st.b = c01;
        st.c = c11 + 0.3f;
    }
    const float det = st.a * st.c - st.b * st.b;
    if (det == 0.f) return false;
    {
        const float inv = 1.f / det;
        st.conic[0] = st.c * inv;
        st.conic[1] = (-st.b) * inv;
        st.conic[2] = st.a * inv;
        const float bm = 0.5f * (st.a + st.c);
        const float disc = sqrtf(fmaxf(0.1f, bm * bm - det));
        const float v1 = bm + disc, v2 = bm - disc;
        st.radius = sgn_f2i_sat(ceilf(3.f * sqrtf(fmaxf(v1, v2))));
    }
    float cxp, cyp;
    {
        const float rw = 1.f / (st.pv[2] + 1e-6f);
        cxp = (st.pv[0] * rw) * cam.fx + cam.cx;
        cyp = (st.pv[1] * rw) * cam.fy + cam.cy;
        const float bw = (float)cam.block_width;
        const int tiles_x = (cam.width + cam.block_width - 1) / cam.block_width;
        const int tiles_y = (cam.height + cam.block_width - 1) / cam.block_width;
        const float tcx = cxp / bw, tcy = cyp / bw, tr = (float)st.radius / bw;
        st.tmin[0] = min(max(0, sgn_f2i_sat(tcx - tr)), tiles_x);
        st.tmax[0] = min(max(0, sgn_f2i_sat((tcx + tr) + 1.f)), tiles_x);
        st.tmin[1] = min(max(0, sgn_f2i_sat(tcy - tr)), tiles_y);
        st.tmax[1] = min(max(0, sgn_f2i_sat((tcy + tr) + 1.f)), tiles_y);
    }
    const int area = (st.tmax[0] - st.tmin[0]) * (st.tmax[1] - st.tmin[1]);
    if (area <= 0) { st.radius = 0; return false; }
    st.xy[0] = cxp;
    st.xy[1] = cyp;
    st.visible = true;
    return true;
}

// ---- SH basis (gsplat "poly" SH, Appendix A.7); not part of the exact section -------------------
__device__ __forceinline__ void sgn_sh_basis(int deg, float x, float y, float z, float Y[16]) {
#pragma unroll
    for (int k = 0; k < 16; ++k) Y[k] = 0.f;
    Y[0] = 0.28209479177387814f;
    if (deg < 1) return;
    const float C1 = 0.4886025119029199f;
    Y[1] = -C1 * y; Y[2] = C1 * z; Y[3] = -C1 * x;
    if (deg < 2) return;
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    Y[4] = 1.0925484305920792f * xy;
    Y[5] = -1.0925484305920792f * yz;
    Y[6] = 0.31539156525252005f * (2.f * zz - xx - yy);
    Y[7] = -1.0925484305920792f * xz;
    Y[8] = 0.5462742152960396f * (xx - yy);
    if (deg < 3) return;
    Y[9] = -0.5900435899266435f * y * (3.f * xx - yy);
    Y[10] = 2.890611442640554f * xy * z;
    Y[11] = -0.4570457994644658f * y * (4.f * zz - xx - yy);
    Y[12] = 0.3731763325901154f * z * (2.f * zz - 3.f * xx - 3.f * yy);
    Y[13] = -0.4570457994644658f * x * (4.f * zz - xx - yy);
    Y[14] = 1.445305721320277f * z * (xx - yy);
    Y[15] = -0.5900435899266435f * x * (xx - 3.f * yy);
}
