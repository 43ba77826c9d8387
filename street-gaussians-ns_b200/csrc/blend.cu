// Front-to-back alpha compositing, forward and backward, for 16x16 tiles.
//
// One traversal of each tile's depth-sorted list produces what the reference obtains from three
// gsplat rasterize_gaussians calls (street_gaussians_ns/sgn_splatfacto.py:954-996 rgb+alpha and the
// depth pass; street_gaussians_ns/sgn_splatfacto_scene_graph.py:366 background-only accumulation): the
// per-pair alpha is evaluated once and fed to two transmittance streams (main / background), each
// with gsplat's own skip and termination rules.  The objects-only
// accumulation (:364-365) runs over the compacted per-tile object sub-lists (binning.cu), which is
// what the reference's subset re-render sees.  The reference's post-ops (:968-975, :995) run in
// the epilogue.
//
// Sorted payloads carry the Gaussian row in bits 0-30 and the object-class flag in bit 31.
#include <cooperative_groups.h>
#include <cooperative_groups/reduce.h>

#include "sgn_common.cuh"

namespace cg = cooperative_groups;

#define BLEND_THREADS 256
#define ALPHA_MIN (1.f / 255.f)
#define T_STOP 1e-4f
#define ID_MASK 0x7fffffff

// saved per-pixel state is planar: slot 0 main, 1 object, 2 background
#define SLOT_MAIN 0
#define SLOT_OBJ 1
#define SLOT_BG 2

struct BlendFwdParams {
    int width, height, tiles_x;
    float clamp_fwd;
    int has_sky, eval_clamp;
    const float4* records;
    const int32_t* sorted_ids;
    const int2* tile_bins;
    const int32_t* obj_ids;
    const int2* obj_bins;
    const float* sky;
    float* rgb;
    float* acc;
    float* depth;
    float* obj_acc;
    float* bg_acc;
    float4* raw;
    float* final_T;      // [3][H*W]
    int32_t* final_idx;  // [3][H*W]
};

// sigma with a fixed operation sequence so the forward and backward kernels take identical
// skip decisions on identical inputs.
__device__ __forceinline__ float sgn_sigma(float ca, float cb, float cc, float dx, float dy) {
    const float t0 = __fmul_rn(ca, __fmul_rn(dx, dx));
    const float t1 = __fmaf_rn(cc, __fmul_rn(dy, dy), t0);
    return __fmaf_rn(cb, __fmul_rn(dx, dy), __fmul_rn(0.5f, t1));
}

template <bool BG>
__global__ void __launch_bounds__(BLEND_THREADS) blend_fwd_kernel(const BlendFwdParams p) {
    __shared__ float4 sA[BLEND_THREADS];  // x y ca cb
    __shared__ float4 sB[BLEND_THREADS];  // cc opac r g
    __shared__ float2 sC[BLEND_THREADS];  // b depth
    __shared__ int sId[BLEND_THREADS];
    const int tile = blockIdx.x;
    const int tx = tile % p.tiles_x, ty = tile / p.tiles_x;
    const int tr = threadIdx.x;
    const int j = tx * SGN_TILE + (tr & 15), i = ty * SGN_TILE + (tr >> 4);
    const float px = (float)j + 0.5f, py = (float)i + 0.5f;
    const bool inside = (i < p.height) && (j < p.width);
    const int2 range = p.tile_bins[tile];
    const int num_batches = (range.y - range.x + BLEND_THREADS - 1) / BLEND_THREADS;

    float T = 1.f, Tb = 1.f;
    int idx = 0, idxb = 0;
    bool done = !inside, doneb = !inside || !BG;
    float4 pix = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int b = 0; b < num_batches; ++b) {
        if (__syncthreads_count(done && doneb) >= BLEND_THREADS) break;
        const int batch_start = range.x + BLEND_THREADS * b;
        const int k = batch_start + tr;
        if (k < range.y) {
            const int id = p.sorted_ids[k];
            const float4* rec = p.records + 3 * (size_t)(id & ID_MASK);
            sA[tr] = __ldg(rec);
            sB[tr] = __ldg(rec + 1);
            const float4 c = __ldg(rec + 2);
            sC[tr] = make_float2(c.x, c.y);
            sId[tr] = id;
        }
        __syncthreads();
        const int batch_size = min(BLEND_THREADS, range.y - batch_start);
        for (int t = 0; t < batch_size && !(done && doneb); ++t) {
            const float4 A = sA[t];
            const float4 B = sB[t];
            const float dx = A.x - px, dy = A.y - py;
            const float sigma = sgn_sigma(A.z, A.w, B.x, dx, dy);
            const float alpha = fminf(p.clamp_fwd, B.y * __expf(-sigma));
            if (sigma < 0.f || alpha < ALPHA_MIN) continue;
            const float om = 1.f - alpha;
            if (!done) {
                const float nT = T * om;
                if (nT <= T_STOP) done = true;
                else {
                    const float2 Cc = sC[t];
                    const float vis = alpha * T;
                    pix.x += B.z * vis; pix.y += B.w * vis; pix.z += Cc.x * vis; pix.w += Cc.y * vis;
                    T = nT;
                    idx = batch_start + t;
                }
            }
            if (BG) {
                if (!doneb && sId[t] >= 0) {
                    const float nT = Tb * om;
                    if (nT <= T_STOP) doneb = true; else { Tb = nT; idxb = batch_start + t; }
                }
            }
        }
    }
    if (!inside) return;
    const size_t P = (size_t)p.width * p.height;
    const size_t pid = (size_t)i * p.width + j;
    const float alpha = 1.f - T;
    p.raw[pid] = pix;
    // post-ops (sgn_splatfacto.py:968-975): clamp(max=1), sky blend (premultiplied rgb times alpha again), eval clamp
    float r = fminf(pix.x, 1.f), g = fminf(pix.y, 1.f), bl = fminf(pix.z, 1.f);
    if (p.has_sky) {
        const float* s = p.sky + 3 * pid;
        r = r * alpha + s[0] * (1.f - alpha);
        g = g * alpha + s[1] * (1.f - alpha);
        bl = bl * alpha + s[2] * (1.f - alpha);
    }
    if (p.eval_clamp) {
        r = fminf(fmaxf(r, 0.f), 1.f); g = fminf(fmaxf(g, 0.f), 1.f); bl = fminf(fmaxf(bl, 0.f), 1.f);
    }
    p.rgb[3 * pid] = r; p.rgb[3 * pid + 1] = g; p.rgb[3 * pid + 2] = bl;
    p.acc[pid] = alpha;
    p.depth[pid] = alpha > 1e-3f ? pix.w / alpha : 10.f;  // sgn_splatfacto.py:995
    p.final_T[SLOT_MAIN * P + pid] = T;
    p.final_idx[SLOT_MAIN * P + pid] = idx;
    if (BG) {
        p.final_T[SLOT_BG * P + pid] = Tb;
        p.final_idx[SLOT_BG * P + pid] = idxb;
        p.bg_acc[pid] = 1.f - Tb;
    }
}

// accumulation-only pass over per-tile sub-lists (objects-only render)
__global__ void __launch_bounds__(BLEND_THREADS) acc_fwd_kernel(const BlendFwdParams p) {
    __shared__ float4 sA[BLEND_THREADS];
    __shared__ float2 sB[BLEND_THREADS];  // cc opac
    const int tile = blockIdx.x;
    const int tx = tile % p.tiles_x, ty = tile / p.tiles_x;
    const int tr = threadIdx.x;
    const int j = tx * SGN_TILE + (tr & 15), i = ty * SGN_TILE + (tr >> 4);
    const float px = (float)j + 0.5f, py = (float)i + 0.5f;
    const bool inside = (i < p.height) && (j < p.width);
    const int2 range = p.obj_bins[tile];
    const int num_batches = (range.y - range.x + BLEND_THREADS - 1) / BLEND_THREADS;
    float T = 1.f;
    int idx = 0;
    bool done = !inside;
    for (int b = 0; b < num_batches; ++b) {
        if (__syncthreads_count(done) >= BLEND_THREADS) break;
        const int batch_start = range.x + BLEND_THREADS * b;
        const int k = batch_start + tr;
        if (k < range.y) {
            const float4* rec = p.records + 3 * (size_t)(p.obj_ids[k] & ID_MASK);
            sA[tr] = __ldg(rec);
            const float4 B = __ldg(rec + 1);
            sB[tr] = make_float2(B.x, B.y);
        }
        __syncthreads();
        const int batch_size = min(BLEND_THREADS, range.y - batch_start);
        for (int t = 0; t < batch_size && !done; ++t) {
            const float4 A = sA[t];
            const float2 B = sB[t];
            const float dx = A.x - px, dy = A.y - py;
            const float sigma = sgn_sigma(A.z, A.w, B.x, dx, dy);
            const float alpha = fminf(p.clamp_fwd, B.y * __expf(-sigma));
            if (sigma < 0.f || alpha < ALPHA_MIN) continue;
            const float nT = T * (1.f - alpha);
            if (nT <= T_STOP) done = true; else { T = nT; idx = batch_start + t; }
        }
    }
    if (!inside) return;
    const size_t P = (size_t)p.width * p.height;
    const size_t pid = (size_t)i * p.width + j;
    p.final_T[SLOT_OBJ * P + pid] = T;
    p.final_idx[SLOT_OBJ * P + pid] = idx;
    p.obj_acc[pid] = 1.f - T;
}

static int check_cam(const sgn_camera* cam) {
    SGN_REQUIRE(cam, "null camera");
    SGN_REQUIRE(cam->block_width == SGN_TILE, "the fused blend kernels require block_width == 16 (got %d)", cam->block_width);
    SGN_REQUIRE(cam->width > 0 && cam->height > 0, "empty image");
    return SGN_OK;
}

extern "C" int sgn_blend_fwd(const sgn_camera* cam, const sgn_blend_opts* opts, const float* records,
                             const int32_t* sorted_ids, const int32_t* tile_bins, const int32_t* obj_ids,
                             const int32_t* obj_bins, const float* sky, const sgn_blend_fwd_out* out, void* stream) {
    if (int rc = check_cam(cam)) return rc;
    SGN_REQUIRE(opts && records && tile_bins && out, "sgn_blend_fwd: null pointer");
    SGN_REQUIRE(out->rgb && out->accumulation && out->depth && out->raw && out->final_T && out->final_idx,
                "sgn_blend_fwd: null output");
    SGN_REQUIRE(!opts->class_streams || (out->object_acc && out->background_acc && obj_ids && obj_bins),
                "class_streams needs object_acc/background_acc outputs and the object sub-lists");
    SGN_REQUIRE(!opts->has_sky || sky, "has_sky set but sky is null");
    SGN_REQUIRE(sgn_aligned16(records) && sgn_aligned16(out->raw), "records / raw must be 16-byte aligned");
    BlendFwdParams p;
    p.width = cam->width; p.height = cam->height;
    p.tiles_x = (cam->width + SGN_TILE - 1) / SGN_TILE;
    const int tiles_y = (cam->height + SGN_TILE - 1) / SGN_TILE;
    p.clamp_fwd = opts->alpha_clamp_fwd;
    p.has_sky = opts->has_sky; p.eval_clamp = opts->eval_clamp;
    p.records = reinterpret_cast<const float4*>(records);
    p.sorted_ids = sorted_ids;
    p.tile_bins = reinterpret_cast<const int2*>(tile_bins);
    p.obj_ids = obj_ids;
    p.obj_bins = reinterpret_cast<const int2*>(obj_bins);
    p.sky = sky;
    p.rgb = out->rgb; p.acc = out->accumulation; p.depth = out->depth;
    p.obj_acc = out->object_acc; p.bg_acc = out->background_acc;
    p.raw = reinterpret_cast<float4*>(out->raw);
    p.final_T = out->final_T; p.final_idx = out->final_idx;
    const int tiles = p.tiles_x * tiles_y;
    if (opts->class_streams) {
        blend_fwd_kernel<true><<<tiles, BLEND_THREADS, 0, (cudaStream_t)stream>>>(p);
        SGN_CHECK_LAUNCH("blend_fwd_kernel<bg>");
        acc_fwd_kernel<<<tiles, BLEND_THREADS, 0, (cudaStream_t)stream>>>(p);
        SGN_CHECK_LAUNCH("acc_fwd_kernel");
    } else {
        blend_fwd_kernel<false><<<tiles, BLEND_THREADS, 0, (cudaStream_t)stream>>>(p);
        SGN_CHECK_LAUNCH("blend_fwd_kernel");
    }
    return SGN_OK;
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
struct BlendBwdParams {
    int width, height, tiles_x;
    float clamp_bwd;
    int has_sky, eval_clamp;
    const float4* records;
    const int32_t* sorted_ids;
    const int2* tile_bins;
    const int32_t* obj_ids;
    const int2* obj_bins;
    const float* v_rgb;
    const float* v_acc;
    const float* v_depth;
    const float* v_obj;
    const float* v_bg;
    const float4* raw;
    const float* final_T;
    const int32_t* final_idx;
    const float* sky;
    float* v_sky;
    float* v_records;
};

__device__ __forceinline__ int block_max(int v, int* s_max, cg::thread_block_tile<32>& warp, int tr) {
    const int w = cg::reduce(warp, v, cg::greater<int>());
    if (warp.thread_rank() == 0) s_max[tr >> 5] = w;
    __syncthreads();
    int m = s_max[0];
#pragma unroll
    for (int k = 1; k < BLEND_THREADS / 32; ++k) m = max(m, s_max[k]);
    return m;
}

template <bool BG>
__global__ void __launch_bounds__(BLEND_THREADS) blend_bwd_kernel(const BlendBwdParams p) {
    __shared__ float4 sA[BLEND_THREADS];
    __shared__ float4 sB[BLEND_THREADS];
    __shared__ float2 sC[BLEND_THREADS];
    __shared__ int sId[BLEND_THREADS];
    __shared__ int s_max[BLEND_THREADS / 32];
    auto block = cg::this_thread_block();
    cg::thread_block_tile<32> warp = cg::tiled_partition<32>(block);
    const int tile = blockIdx.x;
    const int tx = tile % p.tiles_x, ty = tile / p.tiles_x;
    const int tr = threadIdx.x;
    const int j = tx * SGN_TILE + (tr & 15), i = ty * SGN_TILE + (tr >> 4);
    const float px = (float)j + 0.5f, py = (float)i + 0.5f;
    const bool inside = (i < p.height) && (j < p.width);
    const int2 range = p.tile_bins[tile];
    const size_t P = (size_t)p.width * p.height;

    // ---- per-pixel prologue: cotangents of the RAW blend outputs from those of the final outputs
    float4 vo = make_float4(0.f, 0.f, 0.f, 0.f);  // d/d raw rgb, d/d raw depth
    float voa = 0.f, vbg = 0.f;
    float Tf = 1.f, Tfb = 1.f;
    int idx = -1, idxb = -1;
    if (inside) {
        const size_t pid = (size_t)i * p.width + j;
        Tf = p.final_T[SLOT_MAIN * P + pid];
        idx = p.final_idx[SLOT_MAIN * P + pid];
        if (BG && p.v_bg) {
            Tfb = p.final_T[SLOT_BG * P + pid];
            idxb = p.final_idx[SLOT_BG * P + pid];
            vbg = p.v_bg[pid];
        }
        const float alpha = 1.f - Tf;
        const float4 raw = p.raw[pid];
        if (p.v_acc) voa = p.v_acc[pid];
        if (p.v_rgb) {
            float v[3] = {p.v_rgb[3 * pid], p.v_rgb[3 * pid + 1], p.v_rgb[3 * pid + 2]};
            const float rr[3] = {raw.x, raw.y, raw.z};
            float vraw[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float cl = fminf(rr[c], 1.f);
                float fin = cl;
                float s = 0.f;
                if (p.has_sky) { s = p.sky[3 * pid + c]; fin = cl * alpha + s * (1.f - alpha); }
                if (p.eval_clamp && (fin < 0.f || fin > 1.f)) v[c] = 0.f;
                if (p.has_sky) {
                    voa += v[c] * (cl - s);
                    if (p.v_sky) p.v_sky[3 * pid + c] = v[c] * (1.f - alpha);
                    vraw[c] = (rr[c] <= 1.f) ? v[c] * alpha : 0.f;
                } else {
                    vraw[c] = (rr[c] <= 1.f) ? v[c] : 0.f;
                }
            }
            vo.x = vraw[0]; vo.y = vraw[1]; vo.z = vraw[2];
        }
        if (p.v_depth && alpha > 1e-3f) {
            const float vd = p.v_depth[pid];
            vo.w = vd / alpha;
            voa += -vd * raw.w / (alpha * alpha);
        }
    }
    if (range.y <= range.x) return;  // empty tile (after the prologue: v_sky is written for every pixel)
    // a pair at sorted position k matters to this pixel iff k <= kmax
    const int kmax = BG ? max(idx, idxb) : idx;
    const int warp_kmax = cg::reduce(warp, kmax, cg::greater<int>());
    const int block_kmax = block_max(kmax, s_max, warp, tr);
    const int range_end = min(range.y, block_kmax + 1);
    if (range_end <= range.x) return;
    const int num_batches = (range_end - range.x + BLEND_THREADS - 1) / BLEND_THREADS;

    float T = Tf;
    float4 buffer = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int b = 0; b < num_batches; ++b) {
        __syncthreads();
        const int batch_end = range_end - 1 - BLEND_THREADS * b;
        const int batch_size = min(BLEND_THREADS, batch_end + 1 - range.x);
        const int kk = batch_end - tr;
        if (kk >= range.x) {
            const int id = p.sorted_ids[kk];
            sId[tr] = id;
            const float4* rec = p.records + 3 * (size_t)(id & ID_MASK);
            sA[tr] = __ldg(rec);
            sB[tr] = __ldg(rec + 1);
            const float4 c = __ldg(rec + 2);
            sC[tr] = make_float2(c.x, c.y);
        }
        __syncthreads();
        for (int t = max(0, batch_end - warp_kmax); t < batch_size; ++t) {
            const int k = batch_end - t;
            const float4 A = sA[t];
            const float4 B = sB[t];
            bool valid = inside && (k <= kmax);
            float alpha = 0.f, vis = 0.f, dx = 0.f, dy = 0.f;
            if (valid) {
                dx = A.x - px; dy = A.y - py;
                const float sigma = sgn_sigma(A.z, A.w, B.x, dx, dy);
                vis = __expf(-sigma);
                alpha = fminf(p.clamp_bwd, B.y * vis);
                if (sigma < 0.f || alpha < ALPHA_MIN) valid = false;
            }
            if (!warp.any(valid)) continue;
            float l_xy0 = 0.f, l_xy1 = 0.f, l_c0 = 0.f, l_c1 = 0.f, l_c2 = 0.f, l_o = 0.f;
            float l_r = 0.f, l_g = 0.f, l_b = 0.f, l_d = 0.f;
            if (valid) {
                const float ra = 1.f / (1.f - alpha);
                float v_alpha = 0.f;
                if (k <= idx) {
                    const float2 Cc = sC[t];
                    T *= ra;
                    const float fac = alpha * T;
                    l_r = fac * vo.x; l_g = fac * vo.y; l_b = fac * vo.z; l_d = fac * vo.w;
                    v_alpha += (B.z * T - buffer.x * ra) * vo.x;
                    v_alpha += (B.w * T - buffer.y * ra) * vo.y;
                    v_alpha += (Cc.x * T - buffer.z * ra) * vo.z;
                    v_alpha += (Cc.y * T - buffer.w * ra) * vo.w;
                    v_alpha += Tf * ra * voa;
                    buffer.x += B.z * fac; buffer.y += B.w * fac; buffer.z += Cc.x * fac; buffer.w += Cc.y * fac;
                }
                if (BG) {
                    if (sId[t] >= 0 && k <= idxb) v_alpha += Tfb * ra * vbg;
                }
                const float v_sigma = -B.y * vis * v_alpha;
                l_xy0 = v_sigma * (A.z * dx + A.w * dy);
                l_xy1 = v_sigma * (A.w * dx + B.x * dy);
                l_c0 = 0.5f * v_sigma * dx * dx;
                l_c1 = v_sigma * dx * dy;
                l_c2 = 0.5f * v_sigma * dy * dy;
                l_o = vis * v_alpha;
            }
            l_xy0 = cg::reduce(warp, l_xy0, cg::plus<float>());
            l_xy1 = cg::reduce(warp, l_xy1, cg::plus<float>());
            l_c0 = cg::reduce(warp, l_c0, cg::plus<float>());
            l_c1 = cg::reduce(warp, l_c1, cg::plus<float>());
            l_c2 = cg::reduce(warp, l_c2, cg::plus<float>());
            l_o = cg::reduce(warp, l_o, cg::plus<float>());
            l_r = cg::reduce(warp, l_r, cg::plus<float>());
            l_g = cg::reduce(warp, l_g, cg::plus<float>());
            l_b = cg::reduce(warp, l_b, cg::plus<float>());
            l_d = cg::reduce(warp, l_d, cg::plus<float>());
            if (warp.thread_rank() == 0) {
                float* dst = p.v_records + (size_t)(sId[t] & ID_MASK) * SGN_RECORD_FLOATS;
                atomicAdd(dst + 0, l_xy0); atomicAdd(dst + 1, l_xy1);
                atomicAdd(dst + 2, l_c0); atomicAdd(dst + 3, l_c1); atomicAdd(dst + 4, l_c2);
                atomicAdd(dst + 5, l_o);
                atomicAdd(dst + 6, l_r); atomicAdd(dst + 7, l_g); atomicAdd(dst + 8, l_b);
                atomicAdd(dst + 9, l_d);
            }
        }
    }
}

// backward of the accumulation-only pass: out = 1 - T_final  =>  v_alpha_k = T_final * ra_k * v_out
__global__ void __launch_bounds__(BLEND_THREADS) acc_bwd_kernel(const BlendBwdParams p) {
    __shared__ float4 sA[BLEND_THREADS];
    __shared__ float2 sB[BLEND_THREADS];
    __shared__ int sId[BLEND_THREADS];
    __shared__ int s_max[BLEND_THREADS / 32];
    auto block = cg::this_thread_block();
    cg::thread_block_tile<32> warp = cg::tiled_partition<32>(block);
    const int tile = blockIdx.x;
    const int tx = tile % p.tiles_x, ty = tile / p.tiles_x;
    const int tr = threadIdx.x;
    const int j = tx * SGN_TILE + (tr & 15), i = ty * SGN_TILE + (tr >> 4);
    const float px = (float)j + 0.5f, py = (float)i + 0.5f;
    const bool inside = (i < p.height) && (j < p.width);
    const int2 range = p.obj_bins[tile];
    if (range.y <= range.x) return;
    const size_t P = (size_t)p.width * p.height;
    float Tf = 1.f, vout = 0.f;
    int idx = -1;
    if (inside) {
        const size_t pid = (size_t)i * p.width + j;
        Tf = p.final_T[SLOT_OBJ * P + pid];
        idx = p.final_idx[SLOT_OBJ * P + pid];
        vout = p.v_obj[pid];
    }
    const int warp_kmax = cg::reduce(warp, idx, cg::greater<int>());
    const int block_kmax = block_max(idx, s_max, warp, tr);
    const int range_end = min(range.y, block_kmax + 1);
    if (range_end <= range.x) return;
    const int num_batches = (range_end - range.x + BLEND_THREADS - 1) / BLEND_THREADS;
    for (int b = 0; b < num_batches; ++b) {
        __syncthreads();
        const int batch_end = range_end - 1 - BLEND_THREADS * b;
        const int batch_size = min(BLEND_THREADS, batch_end + 1 - range.x);
        const int kk = batch_end - tr;
        if (kk >= range.x) {
            const int id = p.obj_ids[kk];
            sId[tr] = id;
            const float4* rec = p.records + 3 * (size_t)(id & ID_MASK);
            sA[tr] = __ldg(rec);
            const float4 B = __ldg(rec + 1);
            sB[tr] = make_float2(B.x, B.y);
        }
        __syncthreads();
        for (int t = max(0, batch_end - warp_kmax); t < batch_size; ++t) {
            const int k = batch_end - t;
            const float4 A = sA[t];
            const float2 B = sB[t];
            bool valid = inside && (k <= idx);
            float alpha = 0.f, vis = 0.f, dx = 0.f, dy = 0.f;
            if (valid) {
                dx = A.x - px; dy = A.y - py;
                const float sigma = sgn_sigma(A.z, A.w, B.x, dx, dy);
                vis = __expf(-sigma);
                alpha = fminf(p.clamp_bwd, B.y * vis);
                if (sigma < 0.f || alpha < ALPHA_MIN) valid = false;
            }
            if (!warp.any(valid)) continue;
            float l_xy0 = 0.f, l_xy1 = 0.f, l_c0 = 0.f, l_c1 = 0.f, l_c2 = 0.f, l_o = 0.f;
            if (valid) {
                const float ra = 1.f / (1.f - alpha);
                const float v_alpha = Tf * ra * vout;
                const float v_sigma = -B.y * vis * v_alpha;
                l_xy0 = v_sigma * (A.z * dx + A.w * dy);
                l_xy1 = v_sigma * (A.w * dx + B.x * dy);
                l_c0 = 0.5f * v_sigma * dx * dx;
                l_c1 = v_sigma * dx * dy;
                l_c2 = 0.5f * v_sigma * dy * dy;
                l_o = vis * v_alpha;
            }
            l_xy0 = cg::reduce(warp, l_xy0, cg::plus<float>());
            l_xy1 = cg::reduce(warp, l_xy1, cg::plus<float>());
            l_c0 = cg::reduce(warp, l_c0, cg::plus<float>());
            l_c1 = cg::reduce(warp, l_c1, cg::plus<float>());
            l_c2 = cg::reduce(warp, l_c2, cg::plus<float>());
            l_o = cg::reduce(warp, l_o, cg::plus<float>());
            if (warp.thread_rank() == 0) {
                float* dst = p.v_records + (size_t)(sId[t] & ID_MASK) * SGN_RECORD_FLOATS;
                atomicAdd(dst + 0, l_xy0); atomicAdd(dst + 1, l_xy1);
                atomicAdd(dst + 2, l_c0); atomicAdd(dst + 3, l_c1); atomicAdd(dst + 4, l_c2);
                atomicAdd(dst + 5, l_o);
            }
        }
    }
}

extern "C" int sgn_blend_bwd(const sgn_camera* cam, const sgn_blend_opts* opts, const float* records,
                             const int32_t* sorted_ids, const int32_t* tile_bins, const int32_t* obj_ids,
                             const int32_t* obj_bins, const sgn_blend_bwd_in* in, float* v_records, void* stream) {
    if (int rc = check_cam(cam)) return rc;
    SGN_REQUIRE(opts && records && tile_bins && in && v_records, "sgn_blend_bwd: null pointer");
    SGN_REQUIRE(in->raw && in->final_T && in->final_idx, "sgn_blend_bwd: saved forward state missing");
    SGN_REQUIRE(!opts->has_sky || in->sky, "has_sky set but sky is null");
    SGN_REQUIRE(!(in->v_object_acc || in->v_background_acc) || opts->class_streams,
                "cotangents for object_acc/background_acc need class_streams");
    SGN_REQUIRE(!in->v_object_acc || (obj_ids && obj_bins), "v_object_acc needs the object sub-lists");
    SGN_REQUIRE(sgn_aligned16(records) && sgn_aligned16(in->raw), "records / raw must be 16-byte aligned");
    BlendBwdParams p;
    p.width = cam->width; p.height = cam->height;
    p.tiles_x = (cam->width + SGN_TILE - 1) / SGN_TILE;
    const int tiles_y = (cam->height + SGN_TILE - 1) / SGN_TILE;
    p.clamp_bwd = opts->alpha_clamp_bwd;
    p.has_sky = opts->has_sky; p.eval_clamp = opts->eval_clamp;
    p.records = reinterpret_cast<const float4*>(records);
    p.sorted_ids = sorted_ids;
    p.tile_bins = reinterpret_cast<const int2*>(tile_bins);
    p.obj_ids = obj_ids;
    p.obj_bins = reinterpret_cast<const int2*>(obj_bins);
    p.v_rgb = in->v_rgb; p.v_acc = in->v_accumulation; p.v_depth = in->v_depth;
    p.v_obj = in->v_object_acc; p.v_bg = in->v_background_acc;
    p.raw = reinterpret_cast<const float4*>(in->raw);
    p.final_T = in->final_T; p.final_idx = in->final_idx;
    p.sky = in->sky; p.v_sky = in->v_sky;
    p.v_records = v_records;
    const int tiles = p.tiles_x * tiles_y;
    if (opts->class_streams && in->v_background_acc) {
        blend_bwd_kernel<true><<<tiles, BLEND_THREADS, 0, (cudaStream_t)stream>>>(p);
        SGN_CHECK_LAUNCH("blend_bwd_kernel<bg>");
    } else {
        blend_bwd_kernel<false><<<tiles, BLEND_THREADS, 0, (cudaStream_t)stream>>>(p);
        SGN_CHECK_LAUNCH("blend_bwd_kernel");
    }
    if (in->v_object_acc) {
        acc_bwd_kernel<<<tiles, BLEND_THREADS, 0, (cudaStream_t)stream>>>(p);
        SGN_CHECK_LAUNCH("acc_bwd_kernel");
    }
    return SGN_OK;
}
