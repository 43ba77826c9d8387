// Tile binning: cumulative intersects, (tile|depth) key emit, radix sort, per-tile bin edges.
// Semantics: gsplat 0.1.x rasterize_gaussians internals; sort order ==
// stable sort of (tile_id << 32 | float_bits(depth)) with emission order as the tie break.
#include <cub/cub.cuh>

#include "sgn_common.cuh"

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ------------------------------------------------------------------------------------------------
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
    t.tau = __logf(255.f * o);
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
    const float nbc = t.nbc, nba = t.nba;
    float best = 3.4e38f, best_mag = 0.f, mag, q;
    // edges x = x0, x = x1: minimise over dy
    q = touch_q(t, x0, fminf(fmaxf(__fmul_rn(nbc, x0), y0), y1), mag); if (q < best) { best = q; best_mag = mag; }
    q = touch_q(t, x1, fminf(fmaxf(__fmul_rn(nbc, x1), y0), y1), mag); if (q < best) { best = q; best_mag = mag; }
    // edges y = y0, y = y1: minimise over dx
    q = touch_q(t, fminf(fmaxf(__fmul_rn(nba, y0), x0), x1), y0, mag); if (q < best) { best = q; best_mag = mag; }
    q = touch_q(t, fminf(fmaxf(__fmul_rn(nba, y1), x0), x1), y1, mag); if (q < best) { best = q; best_mag = mag; }
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

__global__ void __launch_bounds__(256)
count_tiles_kernel(int N, int width, int height, int bw, const float4* __restrict__ records, const int32_t* __restrict__ radii,
                   const ushort4* __restrict__ tile_bbox, int32_t* __restrict__ tiles_touched) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool vis = (g < N) && radii[g] > 0;
    ushort4 bb = make_ushort4(0, 0, 0, 0);
    TouchCtx t = {};
    if (vis) {
        bb = tile_bbox[g];
        t = make_touch_ctx(records[3 * (size_t)g], records[3 * (size_t)g + 1]);
    }
    const int bwid = bb.z - bb.x, area = bwid * (bb.w - bb.y);
    int n = 0;
    if (vis && area <= COOP_AREA) {
        for (int ty = bb.y; ty < bb.w; ++ty)
            for (int tx = bb.x; tx < bb.z; ++tx) n += tile_touched(t, tx, ty, width, height, bw) ? 1 : 0;
    }
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
    if (g < N) tiles_touched[g] = n;
}

__global__ void write_total_kernel(const int32_t* __restrict__ cum, int N, int64_t* __restrict__ total) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *total = (N > 0) ? (int64_t)cum[N - 1] : 0;
}

extern "C" size_t sgn_bin_scan_scratch_bytes(int N) {
    size_t temp = 0;
    cub::DeviceScan::InclusiveSum(nullptr, temp, (const int32_t*)nullptr, (int32_t*)nullptr, N > 0 ? N : 1);
    return align_up(temp, 256) + 256 + align_up(sizeof(int32_t) * (size_t)(N > 0 ? N : 1), 256);
}

extern "C" int sgn_bin_scan(int N, const sgn_camera* cam, const float* records, const int32_t* radii,
                            const uint16_t* tile_bbox, int32_t* cum, int64_t* total_dev, void* scratch,
                            size_t scratch_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SGN_REQUIRE(cam && records && radii && tile_bbox && cum && total_dev && scratch, "sgn_bin_scan: null pointer");
    if (scratch_bytes < sgn_bin_scan_scratch_bytes(N)) {
        sgn_set_error("sgn_bin_scan: scratch too small (%zu < %zu)", scratch_bytes, sgn_bin_scan_scratch_bytes(N));
        return SGN_ERR_WORKSPACE;
    }
    if (N > 0) {
        const size_t cnt_bytes = align_up(sizeof(int32_t) * (size_t)N, 256);
        int32_t* touched = (int32_t*)scratch;
        count_tiles_kernel<<<(N + 255) / 256, 256, 0, stream>>>(N, cam->width, cam->height, cam->block_width,
                                                                reinterpret_cast<const float4*>(records), radii,
                                                                reinterpret_cast<const ushort4*>(tile_bbox), touched);
        SGN_CHECK_LAUNCH("count_tiles_kernel");
        size_t temp = scratch_bytes - cnt_bytes;
        SGN_CHECK_CUDA(cub::DeviceScan::InclusiveSum((char*)scratch + cnt_bytes, temp, touched, cum, N, stream));
        sgn_count_launch(1);
    }
    write_total_kernel<<<1, 32, 0, stream>>>(cum, N, total_dev);
    SGN_CHECK_LAUNCH("write_total_kernel");
    return SGN_OK;
}

// ------------------------------------------------------------------------------------------------
// key emit: one thread per Gaussian (map_gaussian_to_intersects), only tiles that pass tile_touched
__global__ void __launch_bounds__(256)
emit_keys_kernel(int N, int tiles_x, int width, int height, int bw, const float4* __restrict__ records,
                 const int32_t* __restrict__ radii, const ushort4* __restrict__ tile_bbox, const int32_t* __restrict__ cum,
                 uint64_t* __restrict__ keys, int32_t* __restrict__ vals) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool vis = (g < N) && radii[g] > 0;
    ushort4 bb = make_ushort4(0, 0, 0, 0);
    TouchCtx t = {};
    uint32_t dbits = 0;
    int32_t payload = 0;
    int cur = 0, end = 0;
    if (vis) {
        bb = tile_bbox[g];
        const float4 r0 = records[3 * (size_t)g], r1 = records[3 * (size_t)g + 1], r2 = records[3 * (size_t)g + 2];
        t = make_touch_ctx(r0, r1);
        dbits = (uint32_t)__float_as_int(r2.y);
        // payload: Gaussian row in the low 31 bits, object-class flag in bit 31 (no gather needed later)
        payload = g | ((__float_as_int(r2.z) & SGN_AUX_OBJECT) ? (int32_t)0x80000000 : 0);
        cur = (g == 0) ? 0 : cum[g - 1];
        end = cum[g];
    }
    const int bwid = bb.z - bb.x, area = bwid * (bb.w - bb.y);
    if (vis && area <= COOP_AREA) {
        for (int ty = bb.y; ty < bb.w; ++ty) {
            for (int tx = bb.x; tx < bb.z; ++tx) {
                if (!tile_touched(t, tx, ty, width, height, bw)) continue;
                if (cur >= end) break;  // cannot happen (same test as count_tiles_kernel); never overrun the slot range
                keys[cur] = ((uint64_t)(ty * tiles_x + tx) << 32) | (uint64_t)dbits;
                vals[cur] = payload;
                ++cur;
            }
        }
    }
    unsigned big = __ballot_sync(0xffffffffu, vis && area > COOP_AREA);
    while (big) {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        const TouchCtx c = shfl_ctx(t, src);
        const int x0 = __shfl_sync(0xffffffffu, (int)bb.x, src), y0 = __shfl_sync(0xffffffffu, (int)bb.y, src);
        const int w = __shfl_sync(0xffffffffu, bwid, src), ar = __shfl_sync(0xffffffffu, area, src);
        const uint32_t db = __shfl_sync(0xffffffffu, dbits, src);
        const int32_t pl = __shfl_sync(0xffffffffu, payload, src);
        int pos = __shfl_sync(0xffffffffu, cur, src);
        const int lim = __shfl_sync(0xffffffffu, end, src);
        for (int base = 0; base < ar; base += 32) {
            const int ti = base + lane;
            const int tx = x0 + ti % w, ty = y0 + ti / w;
            const bool ok = (ti < ar) && tile_touched(c, tx, ty, width, height, bw);
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            const int my = pos + __popc(m & ((1u << lane) - 1u));
            if (ok && my < lim) {
                keys[my] = ((uint64_t)(ty * tiles_x + tx) << 32) | (uint64_t)db;
                vals[my] = pl;
            }
            pos += __popc(m);
        }
    }
}

__global__ void __launch_bounds__(256)
bin_edges_kernel(int64_t M, const uint64_t* __restrict__ keys_sorted, int32_t* __restrict__ tile_bins) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const int32_t cur = (int32_t)(keys_sorted[i] >> 32);
    if (i == 0) tile_bins[2 * cur] = 0;
    else {
        const int32_t prev = (int32_t)(keys_sorted[i - 1] >> 32);
        if (prev != cur) {
            tile_bins[2 * prev + 1] = (int32_t)i;
            tile_bins[2 * cur] = (int32_t)i;
        }
    }
    if (i == M - 1) tile_bins[2 * cur + 1] = (int32_t)M;
}

struct SortLayout {
    size_t keys_in, keys_out, vals_in, temp, temp_bytes, total;
};

static SortLayout sort_layout(int64_t M) {
    SortLayout L;
    const size_t m = (size_t)(M > 0 ? M : 1);
    size_t temp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, temp, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const int32_t*)nullptr,
                                    (int32_t*)nullptr, (int64_t)m, 0, 64);
    L.keys_in = 0;
    L.keys_out = align_up(L.keys_in + m * 8, 256);
    L.vals_in = align_up(L.keys_out + m * 8, 256);
    L.temp = align_up(L.vals_in + m * 4, 256);
    L.temp_bytes = temp;
    L.total = align_up(L.temp + temp, 256);
    return L;
}

extern "C" size_t sgn_bin_sort_scratch_bytes(int64_t M) { return sort_layout(M).total; }

extern "C" int sgn_bin_sort(int N, int64_t M, const sgn_camera* cam, const float* records, const int32_t* radii,
                            const uint16_t* tile_bbox, const int32_t* cum, int32_t* sorted_ids, int32_t* tile_bins,
                            void* scratch, size_t scratch_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SGN_REQUIRE(cam && records && radii && tile_bbox && cum && tile_bins && scratch, "sgn_bin_sort: null pointer");
    SGN_REQUIRE(M >= 0 && M < ((int64_t)1 << 31), "sgn_bin_sort: M=%lld out of the int32 range gsplat's cum_tiles_hit supports", (long long)M);
    const int bw = cam->block_width;
    const int tiles_x = (cam->width + bw - 1) / bw, tiles_y = (cam->height + bw - 1) / bw;
    const int tiles = tiles_x * tiles_y;
    SGN_CHECK_CUDA(cudaMemsetAsync(tile_bins, 0, sizeof(int32_t) * 2 * (size_t)tiles, stream));
    if (M == 0 || N == 0) return SGN_OK;
    SGN_REQUIRE(sorted_ids, "sgn_bin_sort: sorted_ids is null");
    const SortLayout L = sort_layout(M);
    if (scratch_bytes < L.total) {
        sgn_set_error("sgn_bin_sort: scratch too small (%zu < %zu)", scratch_bytes, L.total);
        return SGN_ERR_WORKSPACE;
    }
    char* base = (char*)scratch;
    uint64_t* keys_in = (uint64_t*)(base + L.keys_in);
    uint64_t* keys_out = (uint64_t*)(base + L.keys_out);
    int32_t* vals_in = (int32_t*)(base + L.vals_in);
    emit_keys_kernel<<<(N + 255) / 256, 256, 0, stream>>>(N, tiles_x, cam->width, cam->height, bw,
                                                          reinterpret_cast<const float4*>(records), radii,
                                                          reinterpret_cast<const ushort4*>(tile_bbox), cum, keys_in, vals_in);
    SGN_CHECK_LAUNCH("emit_keys_kernel");
    int tile_bits = 1;
    while ((1 << tile_bits) < tiles) ++tile_bits;
    size_t temp = L.temp_bytes;
    SGN_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(base + L.temp, temp, keys_in, keys_out, vals_in, sorted_ids, M, 0,
                                                   32 + tile_bits, stream));
    sgn_count_launch(1);
    bin_edges_kernel<<<(unsigned)((M + 255) / 256), 256, 0, stream>>>(M, keys_out, tile_bins);
    SGN_CHECK_LAUNCH("bin_edges_kernel");
    return SGN_OK;
}

// ------------------------------------------------------------------------------------------------
// per-tile class sub-lists: stable partition of every tile's list into its object entries and its
// background entries.  The objects-only / background-only accumulation renders of the reference
// (get_submodel_output, scene graph :255-303,364-366) see exactly these entries in exactly this order.
// Layout: class c (0 = background, 1 = object) owns cls_ids[c*M ...) and cls_bins[c][tiles][2].
__global__ void __launch_bounds__(256)
class_count_kernel(int tiles, const int2* __restrict__ tile_bins, const int32_t* __restrict__ sorted_ids,
                   int32_t* __restrict__ counts /*[2][tiles]*/) {
    const int tile = blockIdx.x;
    const int2 range = tile_bins[tile];
    int c = 0;
    for (int k = range.x + threadIdx.x; k < range.y; k += blockDim.x) c += (sorted_ids[k] < 0) ? 1 : 0;
    typedef cub::BlockReduce<int, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    const int total = BR(tmp).Sum(c);
    if (threadIdx.x == 0) {
        counts[tiles + tile] = total;
        counts[tile] = (range.y - range.x) - total;
    }
}

__global__ void __launch_bounds__(256)
class_compact_kernel(int tiles, int64_t M, const int2* __restrict__ tile_bins, const int32_t* __restrict__ sorted_ids,
                     const int32_t* __restrict__ offsets /* exclusive scan over [2][tiles] */,
                     const int32_t* __restrict__ counts, int32_t* __restrict__ cls_ids, int2* __restrict__ cls_bins) {
    const int tile = blockIdx.x;
    const int2 range = tile_bins[tile];
    // offsets run over the concatenation [background tiles..., object tiles...]; class 1's base within its own
    // array is offsets - (total background) = offsets - offsets[tiles]
    const int base0 = offsets[tile];
    const int base1 = offsets[tiles + tile] - offsets[tiles];
    if (threadIdx.x == 0) {
        cls_bins[tile] = make_int2(base0, base0 + counts[tile]);
        cls_bins[tiles + tile] = make_int2(base1, base1 + counts[tiles + tile]);
    }
    typedef cub::BlockScan<int, 256> BS;
    __shared__ typename BS::TempStorage tmp;
    int run0 = 0, run1 = 0;
    for (int k0 = range.x; k0 < range.y; k0 += blockDim.x) {
        const int k = k0 + threadIdx.x;
        const bool in = k < range.y;
        const int id = in ? sorted_ids[k] : 0;
        const int flag = (in && id < 0) ? 1 : 0;
        int pos, total;
        BS(tmp).ExclusiveSum(flag, pos, total);
        if (in) {
            if (flag) cls_ids[M + base1 + run1 + pos] = id;
            else cls_ids[base0 + run0 + (threadIdx.x - pos)] = id;
        }
        run1 += total;
        run0 += min((int)blockDim.x, range.y - k0) - total;
        __syncthreads();
    }
}

extern "C" size_t sgn_bin_class_scratch_bytes(int tiles) {
    size_t temp = 0;
    const int n = 2 * (tiles > 0 ? tiles : 1);
    cub::DeviceScan::ExclusiveSum(nullptr, temp, (const int32_t*)nullptr, (int32_t*)nullptr, n);
    return align_up(temp, 256) + 2 * align_up(sizeof(int32_t) * (size_t)n, 256);
}

extern "C" int sgn_bin_class_lists(const sgn_camera* cam, int64_t M, const int32_t* sorted_ids, const int32_t* tile_bins,
                                   int32_t* cls_ids, int32_t* cls_bins, void* scratch, size_t scratch_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SGN_REQUIRE(cam && tile_bins && cls_ids && cls_bins && scratch, "sgn_bin_class_lists: null pointer");
    const int bw = cam->block_width;
    const int tiles = ((cam->width + bw - 1) / bw) * ((cam->height + bw - 1) / bw);
    if (scratch_bytes < sgn_bin_class_scratch_bytes(tiles)) {
        sgn_set_error("sgn_bin_class_lists: scratch too small");
        return SGN_ERR_WORKSPACE;
    }
    char* base = (char*)scratch;
    const size_t arr = align_up(sizeof(int32_t) * 2 * (size_t)tiles, 256);
    int32_t* counts = (int32_t*)base;
    int32_t* offsets = (int32_t*)(base + arr);
    void* temp = base + 2 * arr;
    size_t temp_bytes = scratch_bytes - 2 * arr;
    class_count_kernel<<<tiles, 256, 0, stream>>>(tiles, reinterpret_cast<const int2*>(tile_bins), sorted_ids, counts);
    SGN_CHECK_LAUNCH("class_count_kernel");
    SGN_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(temp, temp_bytes, counts, offsets, 2 * tiles, stream));
    sgn_count_launch(1);
    class_compact_kernel<<<tiles, 256, 0, stream>>>(tiles, M, reinterpret_cast<const int2*>(tile_bins), sorted_ids, offsets, counts,
                                                    cls_ids, reinterpret_cast<int2*>(cls_bins));
    SGN_CHECK_LAUNCH("class_compact_kernel");
    return SGN_OK;
}
