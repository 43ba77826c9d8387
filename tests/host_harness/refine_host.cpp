// TEST INFRASTRUCTURE (not a product path): compiles the SGN_HD row rules of csrc/sgn_refine_rules.cuh -- the bodies
// of the CUDA kernels in csrc/refine.cu -- with g++ and runs them as plain host loops over HOST pointers, behind the
// same two entry-point signatures as the C ABI (include/sgn_raster.h).  tests/test_refine.py checks them against
// the torch restatement of the reference (oracle/oracle_refine.py): the build container has no GPU, so this is
// how the arithmetic and the index logic of the refinement kernels are verified before they reach a B200.
#include "../../street-gaussians-ns_b200/csrc/sgn_refine_rules.cuh"

extern "C" {

size_t sgn_sizeof_refine_config(void) { return sizeof(sgn_refine_config); }
size_t sgn_sizeof_refine_tensors(void) { return sizeof(sgn_refine_tensors); }

int sgn_refine_decide(int n, const sgn_refine_config* cfg, const float* scales, const float* opacities,
                      const float* xys_grad_norm, const float* vis_counts, const float* max_2Dsize, uint8_t* flags,
                      int32_t* marks, void* /*stream*/) {
    for (int i = 0; i < n; ++i) {
        const float g = cfg->densify ? xys_grad_norm[i] : 0.f;
        const float c = cfg->densify ? vis_counts[i] : 1.f;
        const float m = cfg->use_screen_size ? max_2Dsize[i] : 0.f;
        const uint8_t f = sgn_refine_decide_row(*cfg, scales[3 * (size_t)i], scales[3 * (size_t)i + 1], scales[3 * (size_t)i + 2],
                                                opacities[i], g, c, m);
        flags[i] = f;
        marks[i] = (f & SGN_RF_KEEP_ORIG) ? 1 : 0;
        marks[(size_t)n + i] = (f & SGN_RF_KEEP_SPLIT) ? 1 : 0;
        marks[2 * (size_t)n + i] = (f & SGN_RF_KEEP_DUP) ? 1 : 0;
        marks[3 * (size_t)n + i] = (f & SGN_RF_SPLIT) ? 1 : 0;
    }
    return 0;
}

int sgn_refine_apply(int n, const sgn_refine_config* cfg, const sgn_refine_tensors* tensors, const uint8_t* flags,
                     const int32_t* scan, const int32_t* totals, const float* samples, void* /*stream*/) {
    const int row_width = sgn_refine_row_width(*tensors);
    for (int64_t i = 0; i < n; ++i)
        for (int c = 0; c < row_width; ++c) sgn_refine_apply_elem(i, c, n, *cfg, *tensors, flags, scan, totals, samples);
    return 0;
}

const char* sgn_last_error(void) { return ""; }
}
