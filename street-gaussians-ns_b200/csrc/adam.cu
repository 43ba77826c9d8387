// Fused multi-tensor Adam over the flat gradient arena.
//
// The reference optimises the six Gaussian parameter groups of every sub-model with nine torch Adam
// optimizers (street_gaussians_ns/sgn_config.py:71-108: lr per group, eps 1e-15, betas (0.9, 0.999), no weight
// decay), i.e. ~200 parameter tensors per step.  Here ONE launch walks all of them: gradients and both
// moments live in flat arenas with the layout of the project backward's gradient arena, parameters stay
// where the model keeps them (a pointer table).  Semantics: torch.optim.Adam (dense: moments decay for
// Gaussians that were not visible, as in the reference).  HBM-bound: 28 B per element.
#include "sgn_common.cuh"

#define ADAM_THREADS 256
#define ADAM_CHUNK 4096  // elements per block

__global__ void __launch_bounds__(ADAM_THREADS)
adam_kernel(const sgn_adam_tensor* __restrict__ table, int ntensors, const float* __restrict__ grads,
            float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq) {
    extern __shared__ int s_chunk0[];
    for (int i = threadIdx.x; i < ntensors; i += blockDim.x) s_chunk0[i] = table[i].chunk0;
    __syncthreads();
    int lo = 0, hi = ntensors - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (s_chunk0[mid] <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    const sgn_adam_tensor t = table[lo];
    const long long e0 = (long long)(blockIdx.x - t.chunk0) * ADAM_CHUNK;
    const long long n = t.numel;
    float* __restrict__ p = t.param;
    const float* __restrict__ g = grads + t.grad_offset;
    float* __restrict__ m = exp_avg + t.arena_offset;
    float* __restrict__ v = exp_avg_sq + t.arena_offset;
    const float b1 = t.beta1, b2 = t.beta2, step_size = t.step_size, sqrt_bc2 = t.sqrt_bc2, eps = t.eps;
    const float w1 = t.one_minus_beta1, w2 = t.one_minus_beta2;
    const bool vec = ((reinterpret_cast<uintptr_t>(p) & 15u) == 0) && ((t.arena_offset & 3) == 0) && ((t.grad_offset & 3) == 0);
#pragma unroll
    for (int it = 0; it < ADAM_CHUNK / (ADAM_THREADS * 4); ++it) {
        const long long e = e0 + ((long long)it * ADAM_THREADS + threadIdx.x) * 4;
        if (e >= n) break;
        if (vec && e + 4 <= n) {
            const float4 gg = *reinterpret_cast<const float4*>(g + e);
            float4 mm = *reinterpret_cast<const float4*>(m + e);
            float4 vv = *reinterpret_cast<const float4*>(v + e);
            float4 pp = *reinterpret_cast<const float4*>(p + e);
            float* ga = (float*)&gg; float* ma = (float*)&mm; float* va = (float*)&vv; float* pa = (float*)&pp;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                ma[k] = ma[k] + w1 * (ga[k] - ma[k]);        // exp_avg.lerp_(grad, 1 - beta1)
                va[k] = va[k] * b2 + w2 * ga[k] * ga[k];     // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
                const float denom = sqrtf(va[k]) / sqrt_bc2 + eps;
                pa[k] = pa[k] - step_size * (ma[k] / denom);
            }
            *reinterpret_cast<float4*>(m + e) = mm;
            *reinterpret_cast<float4*>(v + e) = vv;
            *reinterpret_cast<float4*>(p + e) = pp;
        } else {
            for (int k = 0; k < 4 && e + k < n; ++k) {
                const float gk = g[e + k];
                const float mk = m[e + k] + w1 * (gk - m[e + k]);
                const float vk = v[e + k] * b2 + w2 * gk * gk;
                m[e + k] = mk; v[e + k] = vk;
                p[e + k] = p[e + k] - step_size * (mk / (sqrtf(vk) / sqrt_bc2 + eps));
            }
        }
    }
}

extern "C" size_t sgn_sizeof_adam_tensor(void) { return sizeof(sgn_adam_tensor); }
extern "C" int sgn_adam_chunk_elems(void) { return ADAM_CHUNK; }

extern "C" int sgn_adam_step(const sgn_adam_tensor* table_dev, int ntensors, int num_chunks, const float* grad_arena,
                             float* exp_avg, float* exp_avg_sq, void* stream) {
    SGN_RANGE("sgn_adam_step");
    SGN_REQUIRE(table_dev && grad_arena && exp_avg && exp_avg_sq, "sgn_adam_step: null pointer");
    SGN_REQUIRE(ntensors >= 1 && ntensors <= 8192, "sgn_adam_step: ntensors=%d out of range [1,8192]", ntensors);
    SGN_REQUIRE(sgn_aligned16(grad_arena) && sgn_aligned16(exp_avg) && sgn_aligned16(exp_avg_sq), "arenas must be 16-byte aligned");
    if (num_chunks <= 0) return SGN_OK;
    adam_kernel<<<num_chunks, ADAM_THREADS, ntensors * sizeof(int), (cudaStream_t)stream>>>(table_dev, ntensors, grad_arena,
                                                                                          exp_avg, exp_avg_sq);
    SGN_CHECK_LAUNCH("adam_kernel");
    return SGN_OK;
}
