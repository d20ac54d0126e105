// kodb_sgd_step — the optimizer step of the reference's train loop: torch.optim.SGD(momentum,
// nesterov=True) over the three parameter groups SmartOptimizer builds (kod/nn/optim/smart.py:11-60: biases without
// decay, weights with decay, norm weights without decay), with the per-group learning rate / momentum the warm-up
// updater interpolates every step (kod/lightning/experiments/yv5_baseline/warmup.py:11-58).
//
// B200-first layout: all parameters, gradients and momentum buffers live in three flat fp32 arenas, partitioned by
// group, so the whole step is ONE streaming launch (read p, g, m; write p, m, and optionally g = 0 for the next
// backward: the separate zero_grad pass disappears) instead of ~180 small per-tensor kernel groups.  Hyper-parameters
// are read from a small DEVICE array (the host refreshes it asynchronously), so the launch is CUDA-graph capturable
// and never synchronises.  HBM-bound: 24 B per parameter (20 B without the gradient reset).
//
// Arithmetic = torch's `_single_tensor_sgd`, op for op, with the fused multiply-adds ATen's `add(alpha=)` kernels use:
//     d   = g + wd * p                  grad.add(param, alpha=weight_decay)        (skipped when wd == 0)
//     m   = mu * m + (1 - damp) * d     buf.mul_(momentum).add_(d, alpha=1 - dampening)   (m == 0 before the first step,
//                                        which reproduces torch's buf = clone(d))
//     d   = d + mu * m  (nesterov)  |  m
//     p   = p - lr * d                  param.add_(d, alpha=-lr)
#include "kodb_common.cuh"

namespace kodb {

constexpr int SGD_MAX_GROUPS = 8;

struct SgdParams {
    float* p;
    float* g;
    float* m;
    long long seg_begin[SGD_MAX_GROUPS + 1];  // float offsets of the groups (multiples of 4)
    int num_groups, nesterov, zero_grads;
    const float* hyper;  // device [num_groups][4]: lr, momentum, weight_decay, dampening
};

__device__ __forceinline__ float sgd_one(float p, float g, float& m, float lr, float mu, float wd, float one_minus_damp,
                                         bool nesterov) {
    float d = g;
    if (wd != 0.0f) d = fmaf(wd, p, d);
    if (mu != 0.0f) {
        m = fmaf(one_minus_damp, d, mu * m);  // mul_ rounds, then add_(d, alpha) is one fused multiply-add
        d = nesterov ? fmaf(mu, m, d) : m;
    }
    return fmaf(-lr, d, p);
}

__global__ void __launch_bounds__(256) sgd_step_kernel(const __grid_constant__ SgdParams q) {
    const long long gsize = (long long)gridDim.x * blockDim.x;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (int grp = 0; grp < q.num_groups; ++grp) {
        const float lr = __ldg(q.hyper + 4 * grp), mu = __ldg(q.hyper + 4 * grp + 1), wd = __ldg(q.hyper + 4 * grp + 2);
        const float omd = 1.0f - __ldg(q.hyper + 4 * grp + 3);
        const long long i0 = q.seg_begin[grp] >> 2, i1 = q.seg_begin[grp + 1] >> 2;
        float4* p4 = reinterpret_cast<float4*>(q.p);
        float4* g4 = reinterpret_cast<float4*>(q.g);
        float4* m4 = reinterpret_cast<float4*>(q.m);
        const bool nest = q.nesterov != 0;
        // two float4 per thread and round: all six loads are in flight before the first use
        for (long long i = i0 + gtid; i < i1; i += 2 * gsize) {
            const long long j = i + gsize;
            const bool two = j < i1;
            float4 p0 = p4[i], g0 = __ldcs(g4 + i), m0 = mu != 0.0f ? m4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 p1 = p0, g1 = g0, m1 = m0;
            if (two) { p1 = p4[j]; g1 = __ldcs(g4 + j); if (mu != 0.0f) m1 = m4[j]; }
            p0.x = sgd_one(p0.x, g0.x, m0.x, lr, mu, wd, omd, nest); p0.y = sgd_one(p0.y, g0.y, m0.y, lr, mu, wd, omd, nest);
            p0.z = sgd_one(p0.z, g0.z, m0.z, lr, mu, wd, omd, nest); p0.w = sgd_one(p0.w, g0.w, m0.w, lr, mu, wd, omd, nest);
            p4[i] = p0;
            if (mu != 0.0f) m4[i] = m0;
            if (q.zero_grads) __stcs(g4 + i, make_float4(0.f, 0.f, 0.f, 0.f));
            if (two) {
                p1.x = sgd_one(p1.x, g1.x, m1.x, lr, mu, wd, omd, nest); p1.y = sgd_one(p1.y, g1.y, m1.y, lr, mu, wd, omd, nest);
                p1.z = sgd_one(p1.z, g1.z, m1.z, lr, mu, wd, omd, nest); p1.w = sgd_one(p1.w, g1.w, m1.w, lr, mu, wd, omd, nest);
                p4[j] = p1;
                if (mu != 0.0f) m4[j] = m1;
                if (q.zero_grads) __stcs(g4 + j, make_float4(0.f, 0.f, 0.f, 0.f));
            }
        }
    }
}

}  // namespace kodb

extern "C" int kodb_sgd_step(float* params, float* grads, float* momentum, const int64_t* seg_begin_h, int32_t num_groups,
                             const float* hyper, int32_t nesterov, int32_t zero_grads, void* stream) {
    using namespace kodb;
    if (!params || !grads || !momentum || !seg_begin_h || !hyper || num_groups < 1 || num_groups > SGD_MAX_GROUPS)
        return KODB_ERR_INVALID_ARGUMENT;
    if ((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) | reinterpret_cast<uintptr_t>(momentum)) & 15u)
        return KODB_ERR_INVALID_ARGUMENT;
    SgdParams q;
    q.p = params; q.g = grads; q.m = momentum;
    for (int i = 0; i <= num_groups; ++i) {
        if (seg_begin_h[i] < 0 || (seg_begin_h[i] & 3) || (i && seg_begin_h[i] < seg_begin_h[i - 1])) return KODB_ERR_INVALID_ARGUMENT;
        q.seg_begin[i] = seg_begin_h[i];
    }
    q.num_groups = num_groups;
    q.nesterov = nesterov;
    q.zero_grads = zero_grads;
    q.hyper = hyper;
    const long long n4 = (q.seg_begin[num_groups] - q.seg_begin[0]) >> 2;
    if (n4 == 0) return KODB_OK;
    // one wave of 148 SMs x 8 CTAs; small models get fewer CTAs (two float4 per thread and round)
    long long blocks = (n4 + 2 * 256 - 1) / (2 * 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    sgd_step_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(q);
    KODB_CHECK_LAUNCH();
    return KODB_OK;
}
