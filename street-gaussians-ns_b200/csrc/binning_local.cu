// Tile binning, variant "local": per-tile lists built by a tile histogram + an unordered scatter, then sorted INSIDE
// each tile by a CTA in shared memory -- instead of two device-wide radix sorts (binning.cu: N rows by depth, then M
// entries by tile).
//
// EXPERIMENTAL (opt-in, SGN_BIN_LOCAL=1 in raster.py): written without GPU access at the end of round 1; the default path
// is binning.cu.  Same contract: every tile's list is ordered like a stable sort of (tile << 32 | float_bits(depth)) with
// the Gaussian row as the tie break, the payload is row | object_class << 31, the same exact
// tile test decides which (tile, Gaussian) pairs exist (sgn_touch.cuh) -- so tests/test_gpu_parity.py's order-equality
// checks apply unchanged.
//
// Why: the device-wide sorts cost ~0.38 ms of a 2.1 ms step on config 3 for 1.3 M rows and 4.5 M entries -- sizes at
// which CUB's passes are latency-bound (5-9 % of HBM peak) -- while the ordering problem is local: a tile's list averages
// ~470 entries (max 4909), which a CTA sorts in shared memory in a few microseconds.  Traffic: 8 B per entry written by
// the scatter, read + 4 B written by the sort; no pass over the N rows besides the two that replay the touch masks.
//
//   local_hist_kernel     per Gaussian (mask replay / warp-cooperative for big AABBs): atomicAdd on its tiles' counters
//   CUB ExclusiveSum      9600 tile counts -> tile starts;  tile_info_kernel: total M and the longest list
//   local_scatter_kernel  same traversal: key = depth_bits << 32 | row << 1 | class at start[tile] + atomicAdd(cursor[tile])
//   local_sort_kernel     CTA per tile: bitonic sort of the 64-bit keys in shared memory, writes payloads + bin edges and,
//                         while the list is still in shared memory, its background / object class sub-lists
//
// Lists longer than sgn_bin_local_cap() do not fit the shared-memory sort: the caller reads the longest list together
// with M (the one read-back the path has anyway) and uses the device-wide path for such a frame.
#include <cub/cub.cuh>

#include "sgn_touch.cuh"

#define LOCAL_CAP 8192        // entries a CTA sorts in shared memory (64 KB of 8-byte keys)
#define LOCAL_SMALL 1024      // size class of the first launch (8 KB, 256 threads)

static inline size_t align_up_l(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Visits every tile the lane's Gaussian reaches, with exactly the decisions of count_touched_tiles / emit_keys_kernel:
// AABBs of at most COOP_AREA tiles replay the bit mask project_fwd stored, larger ones are tested by the whole warp.
// `put(tile, key)` is called by the lane that owns the tile test, with the SOURCE Gaussian's key.
template <class Put>
__device__ __forceinline__ void for_each_touched_tile(bool vis, ushort4 bb, uint32_t mask, const float4* __restrict__ records, size_t g,
                                                      int tiles_x, int width, int height, int bw, unsigned long long key, Put put) {
    const int lane = threadIdx.x & 31;
    const int bwid = bb.z - bb.x, area = bwid * (bb.w - bb.y);
    if (vis && area <= COOP_AREA) {
        const unsigned row_bits = bwid >= 32 ? 0xffffffffu : ((1u << bwid) - 1u);
        for (int ty = bb.y; ty < bb.w && mask; ++ty) {
            unsigned rm = mask & row_bits;
            mask = bwid >= 32 ? 0u : (mask >> bwid);
            const int base = ty * tiles_x + bb.x;
            while (rm) {
                const int bit = __ffs(rm) - 1;
                rm &= rm - 1;
                put(base + bit, key);
            }
        }
    }
    unsigned big = __ballot_sync(0xffffffffu, vis && area > COOP_AREA);
    if (big) {
        TouchCtx t = {};
        if (vis && area > COOP_AREA) t = make_touch_ctx(records[3 * g], records[3 * g + 1]);
        while (big) {
            const int src = __ffs(big) - 1;
            big &= big - 1;
            const TouchCtx c = shfl_ctx(t, src);
            const int x0 = __shfl_sync(0xffffffffu, (int)bb.x, src), y0 = __shfl_sync(0xffffffffu, (int)bb.y, src);
            const int w = __shfl_sync(0xffffffffu, bwid, src), ar = __shfl_sync(0xffffffffu, area, src);
            const unsigned long long k = __shfl_sync(0xffffffffu, key, src);
            for (int base = 0; base < ar; base += 32) {
                const int ti = base + lane;
                if (ti < ar) {
                    const int tx = x0 + ti % w, ty = y0 + ti / w;
                    if (tile_touched(c, tx, ty, width, height, bw)) put(ty * tiles_x + tx, k);
                }
            }
        }
    }
}

struct LocalRow {
    bool vis;
    ushort4 bb;
    uint32_t mask;
    unsigned long long key;
};

__device__ __forceinline__ LocalRow load_row(int g, int N, const float4* __restrict__ records, const int32_t* __restrict__ radii,
                                             const ushort4* __restrict__ tile_bbox, const uint32_t* __restrict__ touch_mask) {
    LocalRow r;
    r.vis = (g < N) && radii[g] > 0;
    r.bb = make_ushort4(0, 0, 0, 0);
    r.mask = 0;
    r.key = 0;
    if (r.vis) {
        r.bb = tile_bbox[g];
        r.mask = touch_mask[g];
        const float4 r2 = records[3 * (size_t)g + 2];
        const unsigned depth_bits = (unsigned)__float_as_int(r2.y);                    // positive floats order like their bits
        const unsigned cls = (__float_as_int(r2.z) & SGN_AUX_OBJECT) ? 1u : 0u;
        r.key = ((unsigned long long)depth_bits << 32) | ((unsigned long long)(unsigned)g << 1) | cls;  // (depth, row): the stable order
    }
    return r;
}

__global__ void __launch_bounds__(256)
local_hist_kernel(int N, int tiles_x, int width, int height, int bw, const float4* __restrict__ records,
                  const int32_t* __restrict__ radii, const ushort4* __restrict__ tile_bbox,
                  const uint32_t* __restrict__ touch_mask, int32_t* __restrict__ tile_count) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const LocalRow r = load_row(g, N, records, radii, tile_bbox, touch_mask);
    for_each_touched_tile(r.vis, r.bb, r.mask, records, (size_t)(g < N ? g : 0), tiles_x, width, height, bw, r.key,
                          [&](int tile, unsigned long long) { atomicAdd(tile_count + tile, 1); });
}

__global__ void __launch_bounds__(256)
tile_info_kernel(int tiles, const int32_t* __restrict__ tile_count, const int32_t* __restrict__ tile_start, int64_t* __restrict__ info) {
    // info[0] = M (total entries), info[1] = longest list; info[1] must be zero on entry
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    int c = (t < tiles) ? tile_count[t] : 0;
    typedef cub::BlockReduce<int, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    const int m = BR(tmp).Reduce(c, cub::Max());
    if (threadIdx.x == 0 && m > 0) atomicMax((unsigned long long*)(info + 1), (unsigned long long)m);
    if (t == tiles - 1) info[0] = (int64_t)tile_start[t] + tile_count[t];
}

__global__ void __launch_bounds__(256)
local_scatter_kernel(int N, int tiles_x, int width, int height, int bw, const float4* __restrict__ records,
                     const int32_t* __restrict__ radii, const ushort4* __restrict__ tile_bbox,
                     const uint32_t* __restrict__ touch_mask, const int32_t* __restrict__ tile_start, int32_t* __restrict__ cursor,
                     long long capacity, unsigned long long* __restrict__ keys) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const LocalRow r = load_row(g, N, records, radii, tile_bbox, touch_mask);
    for_each_touched_tile(r.vis, r.bb, r.mask, records, (size_t)(g < N ? g : 0), tiles_x, width, height, bw, r.key,
                          [&](int tile, unsigned long long k) {
                              const long long pos = (long long)tile_start[tile] + atomicAdd(cursor + tile, 1);
                              if (pos < capacity) keys[pos] = k;  // cannot overflow: same decisions as the histogram pass
                          });
}

// CTA per tile; tiles whose padded length is outside (LO, HI] leave immediately (two launches cover the two size classes)
template <int HI, int LO, int THREADS>
__global__ void __launch_bounds__(THREADS)
local_sort_kernel(int tiles, long long M, const int32_t* __restrict__ tile_count, const int32_t* __restrict__ tile_start,
                  const unsigned long long* __restrict__ keys, int32_t* __restrict__ sorted_ids, int2* __restrict__ tile_bins,
                  int32_t* __restrict__ cls_ids /*[2,M] or null*/, int2* __restrict__ cls_bins /*[2,tiles] or null*/) {
    extern __shared__ unsigned long long s_keys[];
    const int tile = blockIdx.x;
    const int n = tile_count[tile];
    if (n <= LO || n > HI) {
        if (LO == 0 && n == 0 && threadIdx.x == 0) {
            tile_bins[tile] = make_int2(0, 0);  // as the device-wide path leaves empty tiles
            if (cls_bins) { cls_bins[tile] = make_int2(0, 0); cls_bins[tiles + tile] = make_int2(0, 0); }
        }
        return;
    }
    const int start = tile_start[tile];
    int P = 1;
    while (P < n) P <<= 1;
    for (int i = threadIdx.x; i < P; i += THREADS) s_keys[i] = (i < n) ? keys[(size_t)start + i] : 0xffffffffffffffffull;
    __syncthreads();
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < (P >> 1); i += THREADS) {
                const int lo = ((i & ~(j - 1)) << 1) | (i & (j - 1));
                const int hi = lo | j;
                const bool up = (lo & k) == 0;
                const unsigned long long a = s_keys[lo], b = s_keys[hi];
                if ((a > b) == up) { s_keys[lo] = b; s_keys[hi] = a; }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < n; i += THREADS) {
        const unsigned low = (unsigned)s_keys[i];
        sorted_ids[(size_t)start + i] = (int32_t)((low >> 1) | ((low & 1u) << 31));
    }
    if (threadIdx.x == 0) tile_bins[tile] = make_int2(start, start + n);
    if (cls_ids == nullptr) return;
    // Per-tile class sub-lists (what sgn_bin_class_lists builds with a count, a device-wide scan and a compaction): the stable
    // partition of the sorted list into background (class 0) and object (class 1) entries.  Each class array has room for M
    // entries, so a tile's sub-lists simply live at the tile's own offset in their array -- no global offsets needed.
    typedef cub::BlockScan<int, THREADS> BS;
    __shared__ typename BS::TempStorage scan_tmp;
    int run0 = 0, run1 = 0;
    for (int k0 = 0; k0 < n; k0 += THREADS) {
        const int k = k0 + threadIdx.x;
        const bool in = k < n;
        const unsigned low = in ? (unsigned)s_keys[k] : 0u;
        const int flag = (in && (low & 1u)) ? 1 : 0;
        int pos, total;
        BS(scan_tmp).ExclusiveSum(flag, pos, total);
        if (in) {
            const int32_t payload = (int32_t)((low >> 1) | ((low & 1u) << 31));
            if (flag) cls_ids[(size_t)M + start + run1 + pos] = payload;
            else cls_ids[(size_t)start + run0 + ((int)threadIdx.x - pos)] = payload;
        }
        run1 += total;
        run0 += min(THREADS, n - k0) - total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        cls_bins[tile] = make_int2(start, start + run0);
        cls_bins[tiles + tile] = make_int2(start, start + run1);
    }
}

struct LocalLayout {
    size_t cursor, keys, temp, temp_bytes, total;
};
static LocalLayout local_layout(int64_t M, int tiles) {
    LocalLayout L;
    const size_t m = (size_t)(M > 0 ? M : 1), t = (size_t)(tiles > 0 ? tiles : 1);
    size_t temp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, temp, (const int32_t*)nullptr, (int32_t*)nullptr, (int)t);
    L.cursor = 0;
    L.keys = align_up_l(t * 4, 256);
    L.temp = L.keys + align_up_l(m * 8, 256);
    L.temp_bytes = temp;
    L.total = L.temp + align_up_l(temp, 256);
    return L;
}

extern "C" int sgn_bin_local_cap(void) { return LOCAL_CAP; }
extern "C" size_t sgn_bin_local_scratch_bytes(int64_t M, int tiles) { return local_layout(M, tiles).total; }

static int local_tiles(const sgn_camera* cam, int& tiles_x) {
    const int bw = cam->block_width;
    tiles_x = (cam->width + bw - 1) / bw;
    return tiles_x * ((cam->height + bw - 1) / bw);
}

extern "C" int sgn_bin_local_count(int N, const sgn_camera* cam, const float* records, const int32_t* radii, const uint16_t* tile_bbox,
                                   const uint32_t* touch_mask, int32_t* tile_count, int32_t* tile_start, int64_t* info_dev,
                                   void* scratch, size_t scratch_bytes, void* stream_) {
    SGN_RANGE("sgn_bin_local_count");
    cudaStream_t stream = (cudaStream_t)stream_;
    SGN_REQUIRE(cam && records && radii && tile_bbox && touch_mask && tile_count && tile_start && info_dev && scratch,
                "sgn_bin_local_count: null pointer");
    SGN_REQUIRE(N >= 0 && N < (1 << 30), "sgn_bin_local_count: N=%d outside [0, 2^30) (row << 1 | class must fit 32 bits)", N);
    int tiles_x;
    const int tiles = local_tiles(cam, tiles_x);
    const LocalLayout L = local_layout(0, tiles);
    if (scratch_bytes < L.total) {
        sgn_set_error("sgn_bin_local_count: scratch too small (%zu < %zu)", scratch_bytes, L.total);
        return SGN_ERR_WORKSPACE;
    }
    SGN_CHECK_CUDA(cudaMemsetAsync(tile_count, 0, sizeof(int32_t) * (size_t)tiles, stream));
    SGN_CHECK_CUDA(cudaMemsetAsync(info_dev, 0, 2 * sizeof(int64_t), stream));
    if (N > 0) {
        local_hist_kernel<<<(N + 255) / 256, 256, 0, stream>>>(N, tiles_x, cam->width, cam->height, cam->block_width,
                                                              reinterpret_cast<const float4*>(records), radii,
                                                              reinterpret_cast<const ushort4*>(tile_bbox), touch_mask, tile_count);
        SGN_CHECK_LAUNCH("local_hist_kernel");
    }
    size_t temp = L.temp_bytes;
    SGN_CHECK_CUDA(cub::DeviceScan::ExclusiveSum((char*)scratch + L.temp, temp, tile_count, tile_start, tiles, stream));
    sgn_count_launch(1);
    tile_info_kernel<<<(tiles + 255) / 256, 256, 0, stream>>>(tiles, tile_count, tile_start, info_dev);
    SGN_CHECK_LAUNCH("tile_info_kernel");
    return SGN_OK;
}

extern "C" int sgn_bin_local_sort(int N, int64_t M, int longest_list, const sgn_camera* cam, const float* records, const int32_t* radii,
                                  const uint16_t* tile_bbox, const uint32_t* touch_mask, const int32_t* tile_count,
                                  const int32_t* tile_start, int32_t* sorted_ids, int32_t* tile_bins, int32_t* cls_ids,
                                  int32_t* cls_bins, void* scratch, size_t scratch_bytes, void* stream_) {
    SGN_RANGE("sgn_bin_local_sort");
    cudaStream_t stream = (cudaStream_t)stream_;
    SGN_REQUIRE(cam && records && radii && tile_bbox && touch_mask && tile_count && tile_start && tile_bins && scratch,
                "sgn_bin_local_sort: null pointer");
    SGN_REQUIRE(M >= 0 && M < ((int64_t)1 << 31), "sgn_bin_local_sort: M=%lld out of range", (long long)M);
    SGN_REQUIRE(longest_list >= 0 && longest_list <= LOCAL_CAP,
                "sgn_bin_local_sort: a tile lists %d entries, more than the %d a CTA sorts in shared memory: use sgn_bin_scan / sgn_bin_sort "
                "for this frame", longest_list, LOCAL_CAP);
    int tiles_x;
    const int tiles = local_tiles(cam, tiles_x);
    const LocalLayout L = local_layout(M, tiles);
    if (scratch_bytes < L.total) {
        sgn_set_error("sgn_bin_local_sort: scratch too small (%zu < %zu)", scratch_bytes, L.total);
        return SGN_ERR_WORKSPACE;
    }
    SGN_REQUIRE((cls_ids == nullptr) == (cls_bins == nullptr), "sgn_bin_local_sort: cls_ids and cls_bins go together");
    if (M == 0 || N == 0) {
        SGN_CHECK_CUDA(cudaMemsetAsync(tile_bins, 0, sizeof(int32_t) * 2 * (size_t)tiles, stream));
        if (cls_bins) SGN_CHECK_CUDA(cudaMemsetAsync(cls_bins, 0, sizeof(int32_t) * 4 * (size_t)tiles, stream));
        return SGN_OK;
    }
    SGN_REQUIRE(sorted_ids, "sgn_bin_local_sort: sorted_ids is null");
    char* base = (char*)scratch;
    int32_t* cursor = (int32_t*)(base + L.cursor);
    unsigned long long* keys = (unsigned long long*)(base + L.keys);
    SGN_CHECK_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int32_t) * (size_t)tiles, stream));
    local_scatter_kernel<<<(N + 255) / 256, 256, 0, stream>>>(N, tiles_x, cam->width, cam->height, cam->block_width,
                                                             reinterpret_cast<const float4*>(records), radii,
                                                             reinterpret_cast<const ushort4*>(tile_bbox), touch_mask, tile_start, cursor,
                                                             (long long)M, keys);
    SGN_CHECK_LAUNCH("local_scatter_kernel");
    local_sort_kernel<LOCAL_SMALL, 0, 256><<<tiles, 256, LOCAL_SMALL * 8, stream>>>(tiles, (long long)M, tile_count, tile_start, keys,
                                                                                    sorted_ids, reinterpret_cast<int2*>(tile_bins), cls_ids,
                                                                                    reinterpret_cast<int2*>(cls_bins));
    SGN_CHECK_LAUNCH("local_sort_kernel<small>");
    // > 48 KB of dynamic shared memory needs the opt-in (per device; the call is a few hundred nanoseconds)
    SGN_CHECK_CUDA(cudaFuncSetAttribute(local_sort_kernel<LOCAL_CAP, LOCAL_SMALL, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        LOCAL_CAP * 8));
    local_sort_kernel<LOCAL_CAP, LOCAL_SMALL, 1024><<<tiles, 1024, LOCAL_CAP * 8, stream>>>(tiles, (long long)M, tile_count, tile_start,
                                                                                          keys, sorted_ids, reinterpret_cast<int2*>(tile_bins),
                                                                                          cls_ids, reinterpret_cast<int2*>(cls_bins));
    SGN_CHECK_LAUNCH("local_sort_kernel<large>");
    return SGN_OK;
}
