// Exact-section projection math shared by the forward and backward per-Gaussian kernels.
//
// Arithmetic contract: every * + - / sqrt of the exact section goes through the `xf` wrapper below (explicit
// round-to-nearest intrinsics, never contracted), in the order written -- independent of the translation unit's --fmad.
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

// xf: a float whose * + - / are the explicit round-to-nearest intrinsics, which the compiler never contracts into an FMA.
// Every operand of the exact section is an xf, so the section is individually rounded in ANY translation unit, whatever
// its --fmad setting (the colour / gradient code around it is free to use FMAs).
struct xf {
    float v;
    __device__ __forceinline__ xf() {}
    __device__ __forceinline__ xf(float x) : v(x) {}
};
__device__ __forceinline__ xf operator*(xf a, xf b) { return xf(__fmul_rn(a.v, b.v)); }
__device__ __forceinline__ xf operator+(xf a, xf b) { return xf(__fadd_rn(a.v, b.v)); }
__device__ __forceinline__ xf operator-(xf a, xf b) { return xf(__fsub_rn(a.v, b.v)); }
__device__ __forceinline__ xf operator/(xf a, xf b) { return xf(__fdiv_rn(a.v, b.v)); }
__device__ __forceinline__ xf operator-(xf a) { return xf(-a.v); }
__device__ __forceinline__ xf xsqrt(xf a) { return xf(__fsqrt_rn(a.v)); }
__device__ __forceinline__ xf xmax(xf a, xf b) { return xf(fmaxf(a.v, b.v)); }

// exp() as a fixed sequence of IEEE operations (same constants as oracle/sgn_oracle.c).
__device__ __forceinline__ float sgn_expf_exact(float x_) {
    const xf x(fminf(fmaxf(x_, -80.0f), 80.0f));
    const xf n(rintf((x * 1.44269504f).v));
    xf r = x - n * 0.693145752f;
    r = r - n * 1.42860677e-6f;
    xf p(1.98412698e-4f);
    p = p * r + 1.38888889e-3f;
    p = p * r + 8.33333333e-3f;
    p = p * r + 4.16666667e-2f;
    p = p * r + 1.66666667e-1f;
    p = p * r + 0.5f;
    p = p * r + 1.0f;
    p = p * r + 1.0f;
    return ldexpf(p.v, (int)n.v);
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
__device__ __forceinline__ bool sgn_project_exact(const sgn_segment& sg, const sgn_camera& cam, const float m_[3],
                                                  const float ls[3], const float q_[4], SgnProj& st,
                                                  const bool log_scales = true, const float glob_scale = 1.f) {
    xf W[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) W[k] = xf(cam.viewmat[k]);
    st.visible = false;
    st.radius = 0;
    st.xy[0] = st.xy[1] = 0.f;
    st.conic[0] = st.conic[1] = st.conic[2] = 0.f;
    st.tmin[0] = st.tmin[1] = st.tmax[0] = st.tmax[1] = 0;
    st.clampx = st.clampy = 0;
    const xf m[3] = {xf(m_[0]), xf(m_[1]), xf(m_[2])};
    xf mw[3], qr[4];
    if (sg.has_pose) {
        xf R[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) R[k] = xf(sg.R[k]);
        mw[0] = ((R[0] * m[0] + R[1] * m[1]) + R[2] * m[2]) + xf(sg.t[0]);
        mw[1] = ((R[3] * m[0] + R[4] * m[1]) + R[5] * m[2]) + xf(sg.t[1]);
        mw[2] = ((R[6] * m[0] + R[7] * m[1]) + R[8] * m[2]) + xf(sg.t[2]);
        const xf aw(sg.q[0]), ax(sg.q[1]), ay(sg.q[2]), az(sg.q[3]);
        const xf bw(q_[0]), bx(q_[1]), by(q_[2]), bz(q_[3]);
        qr[0] = ((aw * bw - ax * bx) - ay * by) - az * bz;
        qr[1] = ((aw * bx + ax * bw) + ay * bz) - az * by;
        qr[2] = ((aw * by - ax * bz) + ay * bw) + az * bx;
        qr[3] = ((aw * bz + ax * by) - ay * bx) + az * bw;
    } else {
        mw[0] = m[0]; mw[1] = m[1]; mw[2] = m[2];
        qr[0] = xf(q_[0]); qr[1] = xf(q_[1]); qr[2] = xf(q_[2]); qr[3] = xf(q_[3]);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) st.mw[k] = mw[k].v;
#pragma unroll
    for (int k = 0; k < 4; ++k) st.qr[k] = qr[k].v;
    xf pv[3];
    pv[0] = ((W[0] * mw[0] + W[1] * mw[1]) + W[2] * mw[2]) + W[3];
    pv[1] = ((W[4] * mw[0] + W[5] * mw[1]) + W[6] * mw[2]) + W[7];
    pv[2] = ((W[8] * mw[0] + W[9] * mw[1]) + W[10] * mw[2]) + W[11];
#pragma unroll
    for (int k = 0; k < 3; ++k) st.pv[k] = pv[k].v;
    if (pv[2].v <= cam.clip_thresh) return false;
    xf qn[4];
    {
        const xf n2 = ((qr[0] * qr[0] + qr[1] * qr[1]) + qr[2] * qr[2]) + qr[3] * qr[3];
        const xf qnorm = xsqrt(n2);
        st.qnorm = qnorm.v;
#pragma unroll
        for (int k = 0; k < 4; ++k) { qn[k] = qr[k] / qnorm; st.qn[k] = qn[k].v; }
    }
    xf sc[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        sc[k] = log_scales ? xf(sgn_expf_exact(ls[k])) : xf(ls[k]) * xf(glob_scale);
        st.s[k] = sc[k].v;
    }
    xf S[6];
    {
        const xf w = qn[0], x = qn[1], y = qn[2], z = qn[3];
        xf R[9];
        R[0] = xf(1.f) - xf(2.f) * (y * y + z * z);
        R[1] = xf(2.f) * (x * y - w * z);
        R[2] = xf(2.f) * (x * z + w * y);
        R[3] = xf(2.f) * (x * y + w * z);
        R[4] = xf(1.f) - xf(2.f) * (x * x + z * z);
        R[5] = xf(2.f) * (y * z - w * x);
        R[6] = xf(2.f) * (x * z - w * y);
        R[7] = xf(2.f) * (y * z + w * x);
        R[8] = xf(1.f) - xf(2.f) * (x * x + y * y);
        xf M[9];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) { st.Rg[3 * r + c] = R[3 * r + c].v; M[3 * r + c] = R[3 * r + c] * sc[c]; }
        S[0] = (M[0] * M[0] + M[1] * M[1]) + M[2] * M[2];
        S[1] = (M[0] * M[3] + M[1] * M[4]) + M[2] * M[5];
        S[2] = (M[0] * M[6] + M[1] * M[7]) + M[2] * M[8];
        S[3] = (M[3] * M[3] + M[4] * M[4]) + M[5] * M[5];
        S[4] = (M[3] * M[6] + M[4] * M[7]) + M[5] * M[8];
        S[5] = (M[6] * M[6] + M[7] * M[7]) + M[8] * M[8];
#pragma unroll
        for (int k = 0; k < 6; ++k) st.S[k] = S[k].v;
    }
    xf a, b, c;
    {
        const xf z = pv[2];
        const xf rz = xf(1.f) / z;
        const xf rz2 = rz * rz;
        xf ux = pv[0] / z, uy = pv[1] / z;
        if (ux.v > cam.limx) { ux = xf(cam.limx); st.clampx = 1; }
        else if (ux.v < -cam.limx) { ux = xf(-cam.limx); st.clampx = -1; }
        if (uy.v > cam.limy) { uy = xf(cam.limy); st.clampy = 1; }
        else if (uy.v < -cam.limy) { uy = xf(-cam.limy); st.clampy = -1; }
        const xf tx = z * ux, ty = z * uy;
        st.tx = tx.v;
        st.ty = ty.v;
        const xf fx(cam.fx), fy(cam.fy);
        const xf J00 = fx * rz, J11 = fy * rz;
        const xf J02 = -((fx * tx) * rz2);
        const xf J12 = -((fy * ty) * rz2);
        xf T[6];
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
            T[cc] = J00 * W[cc] + J02 * W[8 + cc];
            T[3 + cc] = J11 * W[4 + cc] + J12 * W[8 + cc];
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) st.T[k] = T[k].v;
        xf TS[6];
        TS[0] = (T[0] * S[0] + T[1] * S[1]) + T[2] * S[2];
        TS[1] = (T[0] * S[1] + T[1] * S[3]) + T[2] * S[4];
        TS[2] = (T[0] * S[2] + T[1] * S[4]) + T[2] * S[5];
        TS[3] = (T[3] * S[0] + T[4] * S[1]) + T[5] * S[2];
        TS[4] = (T[3] * S[1] + T[4] * S[3]) + T[5] * S[4];
        TS[5] = (T[3] * S[2] + T[4] * S[4]) + T[5] * S[5];
        const xf c00 = (TS[0] * T[0] + TS[1] * T[1]) + TS[2] * T[2];
        const xf c01 = (TS[0] * T[3] + TS[1] * T[4]) + TS[2] * T[5];
        const xf c11 = (TS[3] * T[3] + TS[4] * T[4]) + TS[5] * T[5];
        a = c00 + xf(0.3f);
        b = c01;
        c = c11 + xf(0.3f);
        st.a = a.v; st.b = b.v; st.c = c.v;
    }
    const xf det = a * c - b * b;
    if (det.v == 0.f) return false;
    {
        const xf inv = xf(1.f) / det;
        st.conic[0] = (c * inv).v;
        st.conic[1] = ((-b) * inv).v;
        st.conic[2] = (a * inv).v;
        const xf bm = xf(0.5f) * (a + c);
        const xf disc = xsqrt(xmax(xf(0.1f), bm * bm - det));
        const xf v1 = bm + disc, v2 = bm - disc;
        st.radius = sgn_f2i_sat(ceilf((xf(3.f) * xsqrt(xmax(v1, v2))).v));
    }
    xf cxp, cyp;
    {
        const xf rw = xf(1.f) / (pv[2] + xf(1e-6f));
        cxp = (pv[0] * rw) * xf(cam.fx) + xf(cam.cx);
        cyp = (pv[1] * rw) * xf(cam.fy) + xf(cam.cy);
        const xf bw((float)cam.block_width);
        const int tiles_x = (cam.width + cam.block_width - 1) / cam.block_width;
        const int tiles_y = (cam.height + cam.block_width - 1) / cam.block_width;
        const xf tcx = cxp / bw, tcy = cyp / bw, tr = xf((float)st.radius) / bw;
        st.tmin[0] = min(max(0, sgn_f2i_sat((tcx - tr).v)), tiles_x);
        st.tmax[0] = min(max(0, sgn_f2i_sat(((tcx + tr) + xf(1.f)).v)), tiles_x);
        st.tmin[1] = min(max(0, sgn_f2i_sat((tcy - tr).v)), tiles_y);
        st.tmax[1] = min(max(0, sgn_f2i_sat(((tcy + tr) + xf(1.f)).v)), tiles_y);
    }
    const int area = (st.tmax[0] - st.tmin[0]) * (st.tmax[1] - st.tmin[1]);
    if (area <= 0) { st.radius = 0; return false; }
    st.xy[0] = cxp.v;
    st.xy[1] = cyp.v;
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
