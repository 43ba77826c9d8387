// Exact (conservative) ellipse-vs-tile test shared by the projection kernel (which counts the tiles a
// Gaussian really reaches) and the key-emit kernel.  Every arithmetic step uses explicit,
// un-contractable intrinsics, so the decision is bit-identical in every translation unit regardless of
// its --fmad setting.
#pragma once
#include "sgn_common.cuh"

// Exact (conservative) tile culling.
//
// gsplat lists a Gaussian in every tile of the AABB of its 3-sigma radius, but a pixel only ever
// uses it when alpha = min(clamp, o*exp(-sigma)) >= 1/255, i.e. sigma <= tau = ln(255*o)
// (SURVEY.md Appendix A.6).  A tile whose pixel centres ALL have sigma > tau is a no-op for every
// stream, forward and backward, so it is dropped from the lists here (about half of the entries on
// the synthetic street scenes).  num_tiles_hit -- the value gsplat reports -- is untouched.
// The test minimises the (convex) quadratic form over the rectangle of the tile's pixel centres; a
// margin covers float rounding in both this test and the blend kernels' own evaluation, so no pair
// the blend would accept is ever dropped.  tests/ check "dropped => no valid pixel" against the oracle.
struct TouchCtx {
    float gx, gy, a, b, c, tau;
    float nbc, nba;  // -b/c, -b/a: minimiser slopes along the rectangle edges
    int always;      // degenerate conic: keep every AABB tile
};

__device__ __forceinline__ TouchCtx make_touch_ctx(const float4 r0, const float4 r1) {
    TouchCtx t;
    t.gx = r0.x; t.gy = r0.y; t.a = r0.z; t.b = r0.w; t.c = r1.x;
    const float o = r1.y;
    t.tau = __logf(__fmul_rn(255.f, o));
    t.always = (!(t.a > 0.f && t.c > 0.f && __fsub_rn(__fmul_rn(t.a, t.c), __fmul_rn(t.b, t.b)) > 0.f) || !(t.tau == t.tau)) ? 1 : 0;
    t.nbc = t.always ? 0.f : __fdiv_rn(-t.b, t.c);
    t.nba = t.always ? 0.f : __fdiv_rn(-t.b, t.a);
    return t;
}

__device__ __forceinline__ float touch_q(const TouchCtx& t, float dx, float dy, float& mag) {
    // explicit, un-contractable operations: count_tiles_kernel and emit_keys_kernel must take
    // bit-identical decisions
    const float qa = __fmul_rn(__fmul_rn(0.5f * t.a, dx), dx), qc = __fmul_rn(__fmul_rn(0.5f * t.c, dy), dy);
    const float qb = __fmul_rn(__fmul_rn(t.b, dx), dy);
    const float s = __fadd_rn(qa, qc);
    mag = __fadd_rn(s, fabsf(qb));
    return __fadd_rn(s, qb);
}

// does the Gaussian reach any pixel centre of tile (tx,ty)?  (pixel centres: 16*tx+0.5 ... +15.5, clipped to the image)
__device__ __forceinline__ bool tile_touched(const TouchCtx& t, int tx, int ty, int width, int height, int bw) {
    if (t.always) return true;
    if (t.tau < 0.f) return false;  // opacity < 1/255: alpha can never reach 1/255
    const float x0 = __fsub_rn((float)(tx * bw) + 0.5f, t.gx), x1 = __fsub_rn(fminf((float)(tx * bw + bw), (float)width) - 0.5f, t.gx);
    const float y0 = __fsub_rn((float)(ty * bw) + 0.5f, t.gy), y1 = __fsub_rn(fminf((float)(ty * bw + bw), (float)height) - 0.5f, t.gy);
    if (x0 <= 0.f && x1 >= 0.f && y0 <= 0.f && y1 >= 0.f) return true;  // centre inside the rectangle
    // The centre lies outside the rectangle, so the minimum of the convex form over it sits on an edge FACING the
    // centre: for an interior minimum on the edge y = y1 the gradient there is (0, g) with g < 0, and convexity
    // (0 = q(centre) >= q(p) - g*p_y) forces p_y < 0, i.e. the centre above that edge; likewise for the other three.
    // At most one vertical and one horizontal edge qualify.
    float best = 3.4e38f, best_mag = 0.f, mag, q;
    if (x0 > 0.f || x1 < 0.f) {  // vertical edge nearest the centre: minimise over dy
        const float xe = x0 > 0.f ? x0 : x1;
        q = touch_q(t, xe, fminf(fmaxf(__fmul_rn(t.nbc, xe), y0), y1), mag);
        best = q; best_mag = mag;
    }
    if (y0 > 0.f || y1 < 0.f) {  // horizontal edge nearest the centre: minimise over dx
        const float ye = y0 > 0.f ? y0 : y1;
        q = touch_q(t, fminf(fmaxf(__fmul_rn(t.nba, ye), x0), x1), ye, mag);
        if (q < best) { best = q; best_mag = mag; }
    }
    return best <= __fadd_rn(__fadd_rn(t.tau, 1e-3f), __fmul_rn(8e-6f, best_mag));
}


// Gaussians whose AABB spans more than COOP_AREA tiles are handled by the whole warp (32 tiles per
// step) after the per-thread pass, so one huge splat does not serialise its warp.
#define COOP_AREA 32

__device__ __forceinline__ TouchCtx shfl_ctx(const TouchCtx& t, int src) {
    TouchCtx r;
    r.gx = __shfl_sync(0xffffffffu, t.gx, src); r.gy = __shfl_sync(0xffffffffu, t.gy, src);
    r.a = __shfl_sync(0xffffffffu, t.a, src); r.b = __shfl_sync(0xffffffffu, t.b, src);
    r.c = __shfl_sync(0xffffffffu, t.c, src); r.tau = __shfl_sync(0xffffffffu, t.tau, src);
    r.nbc = __shfl_sync(0xffffffffu, t.nbc, src); r.nba = __shfl_sync(0xffffffffu, t.nba, src);
    r.always = __shfl_sync(0xffffffffu, t.always, src);
    return r;
}

// Warp-collective: every lane passes its own Gaussian (vis=false for idle lanes).  Returns the number of
// AABB tiles the lane's Gaussian reaches; for AABBs of at most 32 tiles `mask` holds one bit per AABB tile
// (row-major), so the emit kernel does not have to repeat the test.
__device__ __forceinline__ int count_touched_tiles(bool vis, const TouchCtx& t, ushort4 bb, int width, int height, int bw,
                                                   uint32_t& mask) {
    const int lane = threadIdx.x & 31;
    const int bwid = bb.z - bb.x, area = bwid * (bb.w - bb.y);
    int n = 0;
    mask = 0;
    // AABBs of at most COOP_AREA tiles: the tiles of all 32 lanes are laid end to end and tested 32 at a time, whoever owns
    // them (profiles/r02: one lane looping over its own AABB left 5 of 32 lanes active in a loop that was 37 % of
    // project_fwd's instructions).  Tile f of the flattened sequence belongs to the last lane whose exclusive prefix is <= f.
    const int mine = (vis && area <= COOP_AREA) ? area : 0;
    int pre = mine;  // inclusive scan
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, pre, o);
        if (lane >= o) pre += v;
    }
    const int total = __shfl_sync(0xffffffffu, pre, 31);
    pre -= mine;  // exclusive
    for (int base = 0; base < total; base += 32) {
        const int f = base + lane;
        int o = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            const int cand = o + step;  // <= 31
            const int pc = __shfl_sync(0xffffffffu, pre, cand);
            if (pc <= f) o = cand;
        }
        const TouchCtx c = shfl_ctx(t, o);
        const int x0 = __shfl_sync(0xffffffffu, (int)bb.x, o), y0 = __shfl_sync(0xffffffffu, (int)bb.y, o);
        const int w = __shfl_sync(0xffffffffu, bwid, o);
        const int ti = f - __shfl_sync(0xffffffffu, pre, o);
        // ti < 32, w in [1, 32]: the quotient through a float reciprocal is exact for these magnitudes
        const int row = w > 0 ? (int)(((float)ti + 0.5f) / (float)w) : 0;
        const bool ok = (f < total) && tile_touched(c, x0 + (ti - row * w), y0 + row, width, height, bw);
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        const int lo = max(pre, base), hi = min(pre + mine, base + 32);
        if (hi > lo) {
            const unsigned bits = (bal >> (lo - base)) & ((hi - lo) >= 32 ? 0xffffffffu : ((1u << (hi - lo)) - 1u));
            mask |= bits << (lo - pre);
        }
    }
    n = __popc(mask);
    unsigned big = __ballot_sync(0xffffffffu, vis && area > COOP_AREA);
    while (big) {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        const TouchCtx c = shfl_ctx(t, src);
        const int x0 = __shfl_sync(0xffffffffu, (int)bb.x, src), y0 = __shfl_sync(0xffffffffu, (int)bb.y, src);
        const int w = __shfl_sync(0xffffffffu, bwid, src), ar = __shfl_sync(0xffffffffu, area, src);
        int cnt = 0;
        for (int base = 0; base < ar; base += 32) {
            const int ti = base + lane;
            const bool ok = (ti < ar) && tile_touched(c, x0 + ti % w, y0 + ti / w, width, height, bw);
            cnt += __popc(__ballot_sync(0xffffffffu, ok));
        }
        if (lane == src) n = cnt;
    }
    return n;
}
