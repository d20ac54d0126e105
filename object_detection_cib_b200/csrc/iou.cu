// kodb_iou_aligned — IoUCalculator.__call__ (kod/core/bbox/iou.py:249-268) as a standalone elementwise op:
// the aligned IoU family compute_iou / compute_giou / compute_diou / compute_ciou (iou.py:77-95, :136-161, :164-197,
// :200-246) on two (N, 4) xyxy box arrays, forward and -- when gradient buffers are given -- the analytic backward
// with respect to BOTH box arrays (what autograd derives from the reference's op chain: ties of max/min split the
// gradient evenly, clamp(0) passes it at 0, CIoU's alpha is a constant, iou.py:241-242).
//
// The fused train tail (train_tail.cu::iou_metric) evaluates the same expressions on decoded head logits; this kernel
// serves callers that use the calculator on its own.  One thread per box pair; fp32, no FMA contraction (-fmad=false),
// every operation in the reference's order.  Memory-bound: 32 B read + 4 B written per pair (+ 32 B of gradients).
#include "kodb_common.cuh"

namespace kodb {

__device__ __forceinline__ float wi_gt(float a, float b) { return a > b ? 1.0f : (a == b ? 0.5f : 0.0f); }
__device__ __forceinline__ float wi_lt(float a, float b) { return a < b ? 1.0f : (a == b ? 0.5f : 0.0f); }
__device__ __forceinline__ float sgn_f(float a) { return a > 0.0f ? 1.0f : (a < 0.0f ? -1.0f : 0.0f); }

// value of the metric; da / dg = d(value)/d(box a) / d(box g) (xyxy) when WITH_GRAD
template <bool WITH_GRAD>
__device__ __forceinline__ float iou_family(int type, float eps, const float4 a, const float4 g, float4* da, float4* dg) {
    const float x1 = a.x, y1 = a.y, x2 = a.z, y2 = a.w;
    const float x1g = g.x, y1g = g.y, x2g = g.z, y2g = g.w;
    const float tix = fminf(x2, x2g) - fmaxf(x1, x1g), tiy = fminf(y2, y2g) - fmaxf(y1, y1g);
    const float iw = fmaxf(tix, 0.0f), ih = fmaxf(tiy, 0.0f);
    const float inter = iw * ih;
    const float w1 = x2 - x1, h1 = y2 - y1, w2 = x2g - x1g, h2 = y2g - y1g;
    const float uni = (w1 * h1 + w2 * h2) - inter;
    const float ue = uni + eps;
    const float iou = inter / ue;
    float val = iou;

    float g_uni = -inter / (ue * ue);
    const float g_inter_direct = 1.0f / ue;
    float g_cw = 0.0f, g_ch = 0.0f, g_w1 = 0.0f, g_h1 = 0.0f, g_w2 = 0.0f, g_h2 = 0.0f;
    float ax1 = 0.0f, ay1 = 0.0f, ax2 = 0.0f, ay2 = 0.0f;  // d/d(box a)
    float bx1 = 0.0f, by1 = 0.0f, bx2 = 0.0f, by2 = 0.0f;  // d/d(box g)

    if (type != KODB_IOU_IOU) {
        const float cw = fmaxf(x2, x2g) - fminf(x1, x1g), ch = fmaxf(y2, y2g) - fminf(y1, y1g);
        if (type == KODB_IOU_GIOU) {
            const float hull = cw * ch;
            const float num = hull - uni, den = hull + eps;
            val = iou - fabsf(num) / fabsf(den);
            if (WITH_GRAD) {
                const float dnum = sgn_f(num) / fabsf(den);
                const float dden = -fabsf(num) * sgn_f(den) / (den * den);
                const float g_hull = -(dnum + dden);
                g_uni += dnum;
                g_cw += g_hull * ch;
                g_ch += g_hull * cw;
            }
        } else {
            const float diag2 = cw * cw + ch * ch;
            const float dx = (x1 + x2) / 2.0f - (x1g + x2g) / 2.0f, dy = (y1 + y2) / 2.0f - (y1g + y2g) / 2.0f;
            const float cd2 = dx * dx + dy * dy;
            const float de = diag2 + eps;
            const float D = cd2 / de;
            val = iou - D;
            if (WITH_GRAD) {
                const float g_cd2 = -1.0f / de, g_diag2 = cd2 / (de * de);
                ax1 += g_cd2 * dx; ax2 += g_cd2 * dx; ay1 += g_cd2 * dy; ay2 += g_cd2 * dy;
                bx1 -= g_cd2 * dx; bx2 -= g_cd2 * dx; by1 -= g_cd2 * dy; by2 -= g_cd2 * dy;
                g_cw += g_diag2 * 2.0f * cw;
                g_ch += g_diag2 * 2.0f * ch;
            }
            if (type == KODB_IOU_CIOU) {
                const float kk = (float)(4.0 / (3.141592653589793 * 3.141592653589793));
                const float h1e = h1 + eps, h2e = h2 + eps;
                const float r1 = w1 / h1e, r2 = w2 / h2e;
                const float delta = atanf(r2) - atanf(r1);
                const float v = kk * (delta * delta);
                const float alpha = v / ((1.0f - iou) + v + eps);
                val = (iou - D) - alpha * v;
                if (WITH_GRAD) {
                    const float gd = alpha * kk * 2.0f * delta;  // -d(val)/d(delta)
                    const float g_r1 = gd / (1.0f + r1 * r1), g_r2 = -gd / (1.0f + r2 * r2);
                    g_w1 += g_r1 / h1e;
                    g_h1 += -g_r1 * w1 / (h1e * h1e);
                    g_w2 += g_r2 / h2e;
                    g_h2 += -g_r2 * w2 / (h2e * h2e);
                }
            }
        }
        if (WITH_GRAD) {
            ax2 += g_cw * wi_gt(x2, x2g); bx2 += g_cw * wi_gt(x2g, x2);
            ax1 -= g_cw * wi_lt(x1, x1g); bx1 -= g_cw * wi_lt(x1g, x1);
            ay2 += g_ch * wi_gt(y2, y2g); by2 += g_ch * wi_gt(y2g, y2);
            ay1 -= g_ch * wi_lt(y1, y1g); by1 -= g_ch * wi_lt(y1g, y1);
        }
    }
    if (WITH_GRAD) {
        const float g_inter = g_inter_direct - g_uni;  // union = area1 + area2 - inter
        g_w1 += g_uni * h1; g_h1 += g_uni * w1;
        g_w2 += g_uni * h2; g_h2 += g_uni * w2;
        const float g_tix = (tix >= 0.0f) ? g_inter * ih : 0.0f;
        const float g_tiy = (tiy >= 0.0f) ? g_inter * iw : 0.0f;
        ax2 += g_tix * wi_lt(x2, x2g); bx2 += g_tix * wi_lt(x2g, x2);
        ax1 -= g_tix * wi_gt(x1, x1g); bx1 -= g_tix * wi_gt(x1g, x1);
        ay2 += g_tiy * wi_lt(y2, y2g); by2 += g_tiy * wi_lt(y2g, y2);
        ay1 -= g_tiy * wi_gt(y1, y1g); by1 -= g_tiy * wi_gt(y1g, y1);
        ax2 += g_w1; ax1 -= g_w1; ay2 += g_h1; ay1 -= g_h1;
        bx2 += g_w2; bx1 -= g_w2; by2 += g_h2; by1 -= g_h2;
        *da = make_float4(ax1, ay1, ax2, ay2);
        *dg = make_float4(bx1, by1, bx2, by2);
    }
    return val;
}

template <bool WITH_GRAD>
__global__ void __launch_bounds__(256) iou_aligned_kernel(const float4* __restrict__ b1, const float4* __restrict__ b2,
                                                          int64_t n, int type, float eps, float* __restrict__ out,
                                                          const float* __restrict__ upstream, float4* __restrict__ g1,
                                                          float4* __restrict__ g2) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float4 da, dg;
        const float v = iou_family<WITH_GRAD>(type, eps, __ldg(b1 + i), __ldg(b2 + i), &da, &dg);
        if (out) out[i] = v;
        if (WITH_GRAD) {
            const float u = upstream ? upstream[i] : 1.0f;
            if (g1) g1[i] = make_float4(u * da.x, u * da.y, u * da.z, u * da.w);
            if (g2) g2[i] = make_float4(u * dg.x, u * dg.y, u * dg.z, u * dg.w);
        }
    }
}

}  // namespace kodb

extern "C" int kodb_iou_aligned(const float* boxes1, const float* boxes2, int64_t n, int32_t iou_type, float eps,
                                float* out, const float* upstream, float* grad1, float* grad2, void* stream) {
    using namespace kodb;
    if (n < 0 || (n > 0 && (!boxes1 || !boxes2)) || iou_type < KODB_IOU_IOU || iou_type > KODB_IOU_CIOU)
        return KODB_ERR_INVALID_ARGUMENT;
    if (!out && !grad1 && !grad2) return KODB_ERR_INVALID_ARGUMENT;
    if (n == 0) return KODB_OK;
    if ((reinterpret_cast<uintptr_t>(boxes1) | reinterpret_cast<uintptr_t>(boxes2) | reinterpret_cast<uintptr_t>(grad1) |
         reinterpret_cast<uintptr_t>(grad2)) & 15u)
        return KODB_ERR_INVALID_ARGUMENT;  // (N, 4) fp32 rows are read / written as float4
    const int64_t blocks64 = (n + 255) / 256;
    const unsigned blocks = (unsigned)(blocks64 < 148 * 16 ? blocks64 : 148 * 16);
    cudaStream_t st = (cudaStream_t)stream;
    if (grad1 || grad2)
        iou_aligned_kernel<true><<<blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(boxes1),
                                                         reinterpret_cast<const float4*>(boxes2), n, iou_type, eps, out,
                                                         upstream, reinterpret_cast<float4*>(grad1),
                                                         reinterpret_cast<float4*>(grad2));
    else
        iou_aligned_kernel<false><<<blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(boxes1),
                                                          reinterpret_cast<const float4*>(boxes2), n, iou_type, eps, out,
                                                          nullptr, nullptr, nullptr);
    KODB_CHECK_LAUNCH();
    return KODB_OK;
}
