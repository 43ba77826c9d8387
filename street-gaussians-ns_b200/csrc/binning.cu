// Tile binning: per-Gaussian touched-tile counts -> depth order -> key emit -> stable sort by tile ->
// per-tile bin edges (-> class sub-lists).
// Semantics: gsplat 0.1.x rasterize_gaussians internals: every tile's list is
// ordered like a stable sort of (tile_id << 32 | float_bits(depth)) with emission order (Gaussian index)
// as the tie break.  That order is produced in two cheaper stable steps instead of one 46-bit sort of
// the M intersections:
//   1. the N Gaussians are stably sorted by depth bits (32-bit keys, N elements);
//   2. intersections are emitted in that order and stably sorted by their 14-bit tile id only.
// Step 2 moves 12 B per entry twice instead of 24 B six times.
#include <cub/cub.cuh>

#include "sgn_touch.cuh"

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ------------------------------------------------------------------------------------------------
// standalone touched-tile count (the fused path gets it from project_fwd; the Level-1 rasterize path,
// which starts from plain xys/conics tensors, calls this)
__global__ void __launch_bounds__(256)
count_tiles_kernel(int N, int width, int height, int bw, const float4* __restrict__ records, const int32_t* __restrict__ radii,
                   const ushort4* __restrict__ tile_bbox, int32_t* __restrict__ tiles_touched, uint32_t* __restrict__ touch_mask) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const bool vis = (g < N) && radii[g] > 0;
    ushort4 bb = make_ushort4(0, 0, 0, 0);
    TouchCtx t = {};
    if (vis) {
        bb = tile_bbox[g];
        t = make_touch_ctx(records[3 * (size_t)g], records[3 * (size_t)g + 1]);
    }
    uint32_t mask;
    const int n = count_touched_tiles(vis, t, bb, width, height, bw, mask);
    if (g < N) { tiles_touched[g] = n; touch_mask[g] = mask; }
}

extern "C" int sgn_bin_count(int N, const sgn_camera* cam, const float* records, const int32_t* radii,
                             const uint16_t* tile_bbox, int32_t* tiles_touched, uint32_t* touch_mask, void* stream) {
    SGN_RANGE("sgn_bin_count");
    SGN_REQUIRE(cam && records && radii && tile_bbox && tiles_touched && touch_mask, "sgn_bin_count: null pointer");
    if (N == 0) return SGN_OK;
    count_tiles_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        N, cam->width, cam->height, cam->block_width, reinterpret_cast<const float4*>(records), radii,
        reinterpret_cast<const ushort4*>(tile_bbox), tiles_touched, touch_mask);
    SGN_CHECK_LAUNCH("count_tiles_kernel");
    return SGN_OK;
}

// ------------------------------------------------------------------------------------------------
// step 1: depth order + inclusive scan of the touched-tile counts in that order
__global__ void __launch_bounds__(256)
depth_keys_kernel(int N, const float4* __restrict__ records, const int32_t* __restrict__ radii, uint32_t* __restrict__ keys,
                  int32_t* __restrict__ vals) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= N) return;
    // invisible rows sort last (no positive float has all bits set) and emit nothing
    keys[g] = radii[g] > 0 ? (uint32_t)__float_as_int(records[3 * (size_t)g + 2].y) : 0xffffffffu;
    vals[g] = g;
}

struct PermutedCount {
    const int32_t* order;
    const int32_t* counts;
    __host__ __device__ int32_t operator()(int i) const { return counts[order[i]]; }
};

// start[g] = where the run of Gaussian g begins in the depth-ordered entry sequence: the exclusive scan value of its
// depth rank, scattered back to row order so that the emit kernel reads it coalesced (no rank -> cum gathers)
__global__ void __launch_bounds__(256)
start_offsets_kernel(int N, const int32_t* __restrict__ rows_by_depth, const int32_t* __restrict__ cum, int32_t* __restrict__ start) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) start[rows_by_depth[i]] = (i == 0) ? 0 : cum[i - 1];
}

__global__ void write_total_kernel(const int32_t* __restrict__ cum, int N, int64_t* __restrict__ total) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *total = (N > 0) ? (int64_t)cum[N - 1] : 0;
}

struct ScanLayout {
    size_t keys_in, keys_out, vals_in, vals_out, temp, temp_bytes, total;
};
static ScanLayout scan_layout(int N) {
    ScanLayout L;
    const size_t n = (size_t)(N > 0 ? N : 1);
    size_t t1 = 0, t2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, t1, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const int32_t*)nullptr,
                                    (int32_t*)nullptr, (int)n, 0, 32);
    cub::DeviceScan::InclusiveSum(nullptr, t2, (const int32_t*)nullptr, (int32_t*)nullptr, (int)n);
    L.keys_in = 0;
    L.keys_out = align_up(n * 4, 256);
    L.vals_in = L.keys_out + align_up(n * 4, 256);
    L.vals_out = L.vals_in + align_up(n * 4, 256);
    L.temp = L.vals_out + align_up(n * 4, 256);
    L.temp_bytes = t1 > t2 ? t1 : t2;
    L.total = L.temp + align_up(L.temp_bytes, 256) + 256;
    return L;
}

extern "C" size_t sgn_bin_scan_scratch_bytes(int N) { return scan_layout(N).total; }

extern "C" int sgn_bin_scan(int N, const float* records, const int32_t* radii, const int32_t* tiles_touched,
                            int32_t* order /* out: rank[g], the position of row g in depth order */, int32_t* cum,
                            int64_t* total_dev, void* scratch, size_t scratch_bytes, void* stream_) {
    SGN_RANGE("sgn_bin_scan");
    cudaStream_t stream = (cudaStream_t)stream_;
    SGN_REQUIRE(records && radii && tiles_touched && order && cum && total_dev && scratch, "sgn_bin_scan: null pointer");
    const ScanLayout L = scan_layout(N);
    if (scratch_bytes < L.total) {
        sgn_set_error("sgn_bin_scan: scratch too small (%zu < %zu)", scratch_bytes, L.total);
        return SGN_ERR_WORKSPACE;
    }
    if (N > 0) {
        char* base = (char*)scratch;
        uint32_t* keys_in = (uint32_t*)(base + L.keys_in);
        uint32_t* keys_out = (uint32_t*)(base + L.keys_out);
        int32_t* vals_in = (int32_t*)(base + L.vals_in);
        depth_keys_kernel<<<(N + 255) / 256, 256, 0, stream>>>(N, reinterpret_cast<const float4*>(records), radii, keys_in, vals_in);
        SGN_CHECK_LAUNCH("depth_keys_kernel");
        int32_t* sorted_rows = (int32_t*)(base + L.vals_out);
        size_t temp = L.temp_bytes;
        SGN_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(base + L.temp, temp, keys_in, keys_out, vals_in, sorted_rows, N, 0, 32, stream));
        sgn_count_launch(1);
        // counts scanned IN DEPTH ORDER: cum[r] = number of entries emitted by the first r+1 rows of that order
        cub::CountingInputIterator<int> idx(0);
        cub::TransformInputIterator<int32_t, PermutedCount, cub::CountingInputIterator<int>> it(idx, PermutedCount{sorted_rows, tiles_touched});
        temp = L.temp_bytes;
        SGN_CHECK_CUDA(cub::DeviceScan::InclusiveSum(base + L.temp, temp, it, cum, N, stream));
        sgn_count_launch(1);
        start_offsets_kernel<<<(N + 255) / 256, 256, 0, stream>>>(N, sorted_rows, cum, order);
        SGN_CHECK_LAUNCH("start_offsets_kernel");
    }
    write_total_kernel<<<1, 32, 0, stream>>>(cum, N, total_dev);
    SGN_CHECK_LAUNCH("write_total_kernel");
    return SGN_OK;
}

// ------------------------------------------------------------------------------------------------
// step 2: key emit (thread g handles Gaussian g -- coalesced reads -- and writes its run of entries at the
// offset of its depth rank, so the emitted sequence is in depth order), stable sort by tile, bin edges
__global__ void __launch_bounds__(256)
emit_keys_kernel(int N, int tiles_x, int width, int height, int bw, const float4* __restrict__ records,
                 const int32_t* __restrict__ radii, const ushort4* __restrict__ tile_bbox,
                 const uint32_t* __restrict__ touch_mask, const int32_t* __restrict__ start, int end,
                 uint16_t* __restrict__ keys, int32_t* __restrict__ vals) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool vis = (g < N) && radii[g] > 0;
    ushort4 bb = make_ushort4(0, 0, 0, 0);
    TouchCtx t = {};
    int32_t payload = 0;
    uint32_t mask = 0;
    int cur = 0;  // `end` (= M) only guards the buffers: the counts come from the same test as the counting pass
    if (vis) {
        bb = tile_bbox[g];
        // payload: Gaussian row in the low 31 bits, object-class flag in bit 31 (no gather needed later)
        payload = g | ((__float_as_int(records[3 * (size_t)g + 2].z) & SGN_AUX_OBJECT) ? (int32_t)0x80000000 : 0);
        mask = touch_mask[g];
        cur = start[g];
    }
    const int bwid = bb.z - bb.x, area = bwid * (bb.w - bb.y);
    if (vis && area <= COOP_AREA) {
        // the projection kernel already decided every tile of a small AABB: replay its bit mask, one AABB row
        // at a time (no per-tile division by the AABB width)
        const unsigned row_bits = bwid >= 32 ? 0xffffffffu : ((1u << bwid) - 1u);
        for (int ty = bb.y; ty < bb.w && mask && cur < end; ++ty) {
            unsigned rm = mask & row_bits;
            mask = bwid >= 32 ? 0u : (mask >> bwid);
            const int base = ty * tiles_x + bb.x;
            while (rm && cur < end) {
                const int bit = __ffs(rm) - 1;
                rm &= rm - 1;
                keys[cur] = (uint16_t)(base + bit);
                vals[cur] = payload;
                ++cur;
            }
        }
    }
    unsigned big = __ballot_sync(0xffffffffu, vis && area > COOP_AREA);
    if (big) {
        if (vis && area > COOP_AREA) t = make_touch_ctx(records[3 * (size_t)g], records[3 * (size_t)g + 1]);
        while (big) {
            const int src = __ffs(big) - 1;
            big &= big - 1;
            const TouchCtx c = shfl_ctx(t, src);
            const int x0 = __shfl_sync(0xffffffffu, (int)bb.x, src), y0 = __shfl_sync(0xffffffffu, (int)bb.y, src);
            const int w = __shfl_sync(0xffffffffu, bwid, src), ar = __shfl_sync(0xffffffffu, area, src);
            const int32_t pl = __shfl_sync(0xffffffffu, payload, src);
            int pos = __shfl_sync(0xffffffffu, cur, src);
            const int lim = end;
            for (int base = 0; base < ar; base += 32) {
                const int ti = base + lane;
                const int tx = x0 + ti % w, ty = y0 + ti / w;
                const bool ok = (ti < ar) && tile_touched(c, tx, ty, width, height, bw);
                const unsigned m = __ballot_sync(0xffffffffu, ok);
                const int my = pos + __popc(m & ((1u << lane) - 1u));
                if (ok && my < lim) {  // my >= lim cannot happen: same test as the counting pass
                    keys[my] = (uint16_t)(ty * tiles_x + tx);
                    vals[my] = pl;
                }
                pos += __popc(m);
            }
        }
    }
}

__global__ void __launch_bounds__(256)
bin_edges_kernel(int64_t M, const uint16_t* __restrict__ keys_sorted, int32_t* __restrict__ tile_bins) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const int32_t cur = (int32_t)keys_sorted[i];
    if (i == 0) tile_bins[2 * cur] = 0;
    else {
        const int32_t prev = (int32_t)keys_sorted[i - 1];
        if (prev != cur) {
            tile_bins[2 * prev + 1] = (int32_t)i;
            tile_bins[2 * cur] = (int32_t)i;
        }
    }
    if (i == M - 1) tile_bins[2 * cur + 1] = (int32_t)M;
}

// capacity-bounded form: the number of entries stays on the device (no host read-back between the scan and the sort).
// Slots [min(total, cap), cap) are filled with a key that sorts behind every tile, so the sort runs over `cap` items.
__global__ void __launch_bounds__(256)
pad_keys_kernel(int64_t cap, const int64_t* __restrict__ total, uint16_t sentinel, uint16_t* __restrict__ keys, int32_t* __restrict__ vals) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t m = min(*total, cap);
    if (i >= m && i < cap) { keys[i] = sentinel; vals[i] = 0; }
}
__global__ void __launch_bounds__(256)
bin_edges_capped_kernel(int64_t cap, const int64_t* __restrict__ total, const uint16_t* __restrict__ keys_sorted, int32_t* __restrict__ tile_bins,
                        int32_t* __restrict__ overflow) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t M = min(*total, cap);
    if (i == 0 && *total > cap) *overflow = 1;  // the lists of this frame are truncated: the host finds out with the next frame
    if (i >= M) return;
    const int32_t cur = (int32_t)keys_sorted[i];
    if (i == 0) tile_bins[2 * cur] = 0;
    else {
        const int32_t prev = (int32_t)keys_sorted[i - 1];
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
    cub::DeviceRadixSort::SortPairs(nullptr, temp, (const uint16_t*)nullptr, (uint16_t*)nullptr, (const int32_t*)nullptr,
                                    (int32_t*)nullptr, (int64_t)m, 0, 16);
    L.keys_in = 0;
    L.keys_out = align_up(L.keys_in + m * 2, 256);
    L.vals_in = align_up(L.keys_out + m * 2, 256);
    L.temp = align_up(L.vals_in + m * 4, 256);
    L.temp_bytes = temp;
    L.total = align_up(L.temp + temp, 256);
    return L;
}

extern "C" size_t sgn_bin_sort_scratch_bytes(int64_t M) { return sort_layout(M).total; }

static int bin_sort_impl(int N, int64_t M, const int64_t* total_dev, int32_t* overflow_dev, const sgn_camera* cam, const float* records,
                         const int32_t* radii, const uint16_t* tile_bbox, const uint32_t* touch_mask, const int32_t* order, const int32_t* cum,
                         int32_t* sorted_ids, int32_t* tile_bins, void* scratch, size_t scratch_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SGN_REQUIRE(cam && records && radii && tile_bbox && touch_mask && order && cum && tile_bins && scratch,
                "sgn_bin_sort: null pointer");
    SGN_REQUIRE(M >= 0 && M < ((int64_t)1 << 31), "sgn_bin_sort: M=%lld out of the int32 range gsplat's cum_tiles_hit supports", (long long)M);
    const int bw = cam->block_width;
    const int tiles_x = (cam->width + bw - 1) / bw, tiles_y = (cam->height + bw - 1) / bw;
    const int tiles = tiles_x * tiles_y;
    SGN_REQUIRE(tiles <= 65536, "sgn_bin_sort: more than 65536 tiles (16-bit tile keys)");
    SGN_CHECK_CUDA(cudaMemsetAsync(tile_bins, 0, sizeof(int32_t) * 2 * (size_t)tiles, stream));
    if (M == 0 || N == 0) return SGN_OK;
    SGN_REQUIRE(sorted_ids, "sgn_bin_sort: sorted_ids is null");
    SGN_REQUIRE(!total_dev || overflow_dev, "sgn_bin_sort_capped: overflow flag is null");
    const SortLayout L = sort_layout(M);
    if (scratch_bytes < L.total) {
        sgn_set_error("sgn_bin_sort: scratch too small (%zu < %zu)", scratch_bytes, L.total);
        return SGN_ERR_WORKSPACE;
    }
    char* base = (char*)scratch;
    uint16_t* keys_in = (uint16_t*)(base + L.keys_in);
    uint16_t* keys_out = (uint16_t*)(base + L.keys_out);
    int32_t* vals_in = (int32_t*)(base + L.vals_in);
    emit_keys_kernel<<<(N + 255) / 256, 256, 0, stream>>>(N, tiles_x, cam->width, cam->height, bw,
                                                          reinterpret_cast<const float4*>(records), radii,
                                                          reinterpret_cast<const ushort4*>(tile_bbox), touch_mask, order, (int)M,
                                                          keys_in, vals_in);
    SGN_CHECK_LAUNCH("emit_keys_kernel");
    int tile_bits = 1;
    while ((1 << tile_bits) < tiles + (total_dev ? 1 : 0)) ++tile_bits;  // capped: one more key value, the padding sentinel
    if (total_dev) {
        SGN_REQUIRE(tile_bits <= 16, "sgn_bin_sort_capped: %d tiles leave no 16-bit key for the padding", tiles);
        pad_keys_kernel<<<(unsigned)((M + 255) / 256), 256, 0, stream>>>(M, total_dev, (uint16_t)((1u << tile_bits) - 1u), keys_in, vals_in);
        SGN_CHECK_LAUNCH("pad_keys_kernel");
    }
    size_t temp = L.temp_bytes;
    SGN_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(base + L.temp, temp, keys_in, keys_out, vals_in, sorted_ids, M, 0, tile_bits, stream));
    sgn_count_launch(1);
    if (total_dev) {
        bin_edges_capped_kernel<<<(unsigned)((M + 255) / 256), 256, 0, stream>>>(M, total_dev, keys_out, tile_bins, overflow_dev);
        SGN_CHECK_LAUNCH("bin_edges_capped_kernel");
    } else {
        bin_edges_kernel<<<(unsigned)((M + 255) / 256), 256, 0, stream>>>(M, keys_out, tile_bins);
        SGN_CHECK_LAUNCH("bin_edges_kernel");
    }
    return SGN_OK;
}

extern "C" int sgn_bin_sort(int N, int64_t M, const sgn_camera* cam, const float* records, const int32_t* radii,
                            const uint16_t* tile_bbox, const uint32_t* touch_mask, const int32_t* order, const int32_t* cum,
                            int32_t* sorted_ids, int32_t* tile_bins, void* scratch, size_t scratch_bytes, void* stream_) {
    SGN_RANGE("sgn_bin_sort");
    return bin_sort_impl(N, M, nullptr, nullptr, cam, records, radii, tile_bbox, touch_mask, order, cum, sorted_ids, tile_bins, scratch,
                         scratch_bytes, stream_);
}

extern "C" int sgn_bin_sort_capped(int N, int64_t capacity, const int64_t* total_dev, int32_t* overflow_dev, const sgn_camera* cam,
                                   const float* records, const int32_t* radii, const uint16_t* tile_bbox, const uint32_t* touch_mask,
                                   const int32_t* order, const int32_t* cum, int32_t* sorted_ids, int32_t* tile_bins, void* scratch,
                                   size_t scratch_bytes, void* stream_) {
    SGN_RANGE("sgn_bin_sort_capped");
    SGN_REQUIRE(total_dev && capacity > 0, "sgn_bin_sort_capped: needs the device-side entry count and a positive capacity");
    return bin_sort_impl(N, capacity, total_dev, overflow_dev, cam, records, radii, tile_bbox, touch_mask, order, cum, sorted_ids, tile_bins,
                         scratch, scratch_bytes, stream_);
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
    SGN_RANGE("sgn_bin_class_lists");
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
