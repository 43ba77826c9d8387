// Refinement of one sub-model: split / duplicate / cull (second half).
//
// The reference runs `refinement_after` per sub-model every `refine_every` steps as ~120 whole-tensor torch
// statements (street_gaussians_ns/sgn_splatfacto.py:550-646): masks, `.sum().item()` host syncs, `param[mask]`,
// `.repeat`, `torch.cat` of every parameter, and the same surgery again on the Adam state of every group
// (:459-511) -- each parameter row is read and written about five times.  Here the per-row rules
// (sgn_refine_rules.cuh) run as two streaming kernels:
//
//   refine_decide_kernel  thread per row: statistics + log-scales + opacity -> flag byte + four 0/1 marks
//   (the caller prefix-sums the marks, reads four totals, draws the split samples)
//   refine_apply_kernel   thread per (row, column) of the 56+3F floats of a Gaussian: copies the element -- and its
//                         two Adam moments -- to every output row it feeds, in the reference's output order
//
// Both are HBM-bound: decide reads 32 B per row; apply moves each surviving element once (3 x 4 B read, 3 x 4 B
// written).  Compiled with --fmad=false: the rules mirror torch's separately rounded elementwise kernels.
#include "sgn_common.cuh"
#include "sgn_refine_rules.cuh"

#define REFINE_THREADS 256

__global__ void __launch_bounds__(REFINE_THREADS)
refine_decide_kernel(int n, const sgn_refine_config cfg, const float* __restrict__ scales, const float* __restrict__ opacities,
                     const float* __restrict__ xys_grad_norm, const float* __restrict__ vis_counts,
                     const float* __restrict__ max_2Dsize, uint8_t* __restrict__ flags, int32_t* __restrict__ marks) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float sx = scales[3 * (size_t)i], sy = scales[3 * (size_t)i + 1], sz = scales[3 * (size_t)i + 2];
    const float g = cfg.densify ? xys_grad_norm[i] : 0.f;
    const float c = cfg.densify ? vis_counts[i] : 1.f;
    const float m = cfg.use_screen_size ? max_2Dsize[i] : 0.f;
    const uint8_t f = sgn_refine_decide_row(cfg, sx, sy, sz, opacities[i], g, c, m);
    flags[i] = f;
    marks[i] = (f & SGN_RF_KEEP_ORIG) ? 1 : 0;
    marks[(size_t)n + i] = (f & SGN_RF_KEEP_SPLIT) ? 1 : 0;
    marks[2 * (size_t)n + i] = (f & SGN_RF_KEEP_DUP) ? 1 : 0;
    marks[3 * (size_t)n + i] = (f & SGN_RF_SPLIT) ? 1 : 0;
}

struct refine_totals {
    int32_t v[4];
};

__global__ void __launch_bounds__(REFINE_THREADS)
refine_apply_kernel(int n, int row_width, const sgn_refine_config cfg, const sgn_refine_tensors t,
                    const uint8_t* __restrict__ flags, const int32_t* __restrict__ scan, const refine_totals totals,
                    const float* __restrict__ samples) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)n * row_width;
    if (e >= total) return;
    long long i;
    int c;
    if (total < (1ll << 32)) {  // the usual case: a 32-bit divide instead of the 64-bit software routine
        const unsigned e32 = (unsigned)e, w32 = (unsigned)row_width;
        const unsigned i32 = e32 / w32;
        i = i32;
        c = (int)(e32 - i32 * w32);
    } else {
        i = e / row_width;
        c = (int)(e - i * row_width);
    }
    sgn_refine_apply_elem(i, c, n, cfg, t, flags, scan, totals.v, samples);
}

extern "C" size_t sgn_sizeof_refine_config(void) { return sizeof(sgn_refine_config); }
extern "C" size_t sgn_sizeof_refine_tensors(void) { return sizeof(sgn_refine_tensors); }

static int refine_check_config(const sgn_refine_config* cfg, const char* who) {
    SGN_REQUIRE(cfg != nullptr, "%s: null config", who);
    SGN_REQUIRE(cfg->n_split_samples >= 1 && cfg->n_split_samples <= 16, "%s: n_split_samples=%d out of range [1,16]", who,
                cfg->n_split_samples);
    SGN_REQUIRE(cfg->inv_size_fac > 0.f && cfg->max_size > 0.f, "%s: inv_size_fac and max_size must be positive", who);
    return SGN_OK;
}

extern "C" int sgn_refine_decide(int n, const sgn_refine_config* cfg, const float* scales, const float* opacities,
                                 const float* xys_grad_norm, const float* vis_counts, const float* max_2Dsize, uint8_t* flags,
                                 int32_t* marks, void* stream) {
    SGN_RANGE("sgn_refine_decide");
    SGN_REQUIRE(n >= 0, "sgn_refine_decide: n=%d", n);
    if (int rc = refine_check_config(cfg, "sgn_refine_decide")) return rc;
    if (n == 0) return SGN_OK;
    SGN_REQUIRE(scales && opacities && flags && marks, "sgn_refine_decide: null pointer");
    SGN_REQUIRE(!cfg->densify || (xys_grad_norm && vis_counts), "sgn_refine_decide: densify needs xys_grad_norm and vis_counts");
    SGN_REQUIRE(!cfg->use_screen_size || max_2Dsize, "sgn_refine_decide: use_screen_size needs max_2Dsize");
    refine_decide_kernel<<<(n + REFINE_THREADS - 1) / REFINE_THREADS, REFINE_THREADS, 0, (cudaStream_t)stream>>>(
        n, *cfg, scales, opacities, xys_grad_norm, vis_counts, max_2Dsize, flags, marks);
    SGN_CHECK_LAUNCH("refine_decide_kernel");
    return SGN_OK;
}

extern "C" int sgn_refine_apply(int n, const sgn_refine_config* cfg, const sgn_refine_tensors* tensors, const uint8_t* flags,
                                const int32_t* scan, const int32_t* totals, const float* samples, void* stream) {
    SGN_RANGE("sgn_refine_apply");
    SGN_REQUIRE(n >= 0, "sgn_refine_apply: n=%d", n);
    if (int rc = refine_check_config(cfg, "sgn_refine_apply")) return rc;
    SGN_REQUIRE(tensors && totals, "sgn_refine_apply: null tensors / totals");
    if (n == 0) return SGN_OK;
    SGN_REQUIRE(flags && scan, "sgn_refine_apply: null flags / scan");
    SGN_REQUIRE(totals[0] >= 0 && totals[1] >= 0 && totals[2] >= 0 && totals[3] >= totals[1] && totals[0] <= n && totals[3] <= n &&
                    totals[2] <= n,
                "sgn_refine_apply: inconsistent totals {%d,%d,%d,%d} for n=%d", totals[0], totals[1], totals[2], totals[3], n);
    SGN_REQUIRE(totals[1] == 0 || samples, "sgn_refine_apply: split rows need samples");
    const long long out_rows = (long long)totals[0] + (long long)cfg->n_split_samples * totals[1] + totals[2];
    SGN_REQUIRE(out_rows < (1ll << 31), "sgn_refine_apply: %lld output rows exceed the int32 row space", out_rows);
    bool any_m = false, all_m = true;
    for (int k = 0; k < 6; ++k) {
        SGN_REQUIRE(tensors->width[k] >= (k == SGN_RT_DC || k == SGN_RT_REST ? 0 : 1) && tensors->width[k] <= 3 * 64,
                    "sgn_refine_apply: width[%d]=%d", k, tensors->width[k]);
        SGN_REQUIRE(tensors->width[k] == 0 || (tensors->src[k] && (out_rows == 0 || tensors->dst[k])),
                    "sgn_refine_apply: null parameter pointer %d", k);
        const bool m = tensors->src_m[k] && tensors->src_v[k] && (out_rows == 0 || (tensors->dst_m[k] && tensors->dst_v[k]));
        const bool none = !tensors->src_m[k] && !tensors->src_v[k];
        SGN_REQUIRE(m || none, "sgn_refine_apply: tensor %d has a partial optimizer state", k);
        if (tensors->width[k]) { any_m = any_m || m; all_m = all_m && m; }
    }
    SGN_REQUIRE(!any_m || all_m, "sgn_refine_apply: optimizer state must be given for all tensors or none");
    SGN_REQUIRE(tensors->width[SGN_RT_MEANS] == 3 && tensors->width[SGN_RT_SCALES] == 3 && tensors->width[SGN_RT_QUATS] == 4 &&
                    tensors->width[SGN_RT_OPAC] == 1,
                "sgn_refine_apply: means/scales/quats/opacities rows must be 3/3/4/1 floats");
    if (out_rows == 0) return SGN_OK;
    const int row_width = sgn_refine_row_width(*tensors);
    refine_totals tt;
    for (int k = 0; k < 4; ++k) tt.v[k] = totals[k];
    const long long elems = (long long)n * row_width;
    const long long blocks = (elems + REFINE_THREADS - 1) / REFINE_THREADS;
    SGN_REQUIRE(blocks < (1ll << 31), "sgn_refine_apply: grid too large");
    refine_apply_kernel<<<(unsigned)blocks, REFINE_THREADS, 0, (cudaStream_t)stream>>>(n, row_width, *cfg, *tensors, flags, scan, tt,
                                                                                       samples);
    SGN_CHECK_LAUNCH("refine_apply_kernel");
    return SGN_OK;
}
