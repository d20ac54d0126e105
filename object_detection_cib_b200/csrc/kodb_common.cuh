// Shared device/host helpers for libkodb (sm_100a).  Compiled with -fmad=false: every fp32 expression
// below rounds exactly like the reference's separate ATen ops; FMAs are written explicitly (fmaf)
// only where bit-parity with the reference is not required.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/kodb.h"

#define KODB_WARP 32
#define KODB_FULL_MASK 0xffffffffu

#define KODB_CHECK_LAUNCH()                              \
    do {                                                 \
        cudaError_t e__ = cudaGetLastError();            \
        if (e__ != cudaSuccess) return KODB_ERR_LAUNCH;  \
    } while (0)

namespace kodb {

// ---- small math -----------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoid_f(float x) {
    // 1/(1+exp(-x)) with the accurate expf and IEEE division (torch.sigmoid semantics).
    return 1.0f / (1.0f + expf(-x));
}

// softplus(x) = -log(sigmoid(-x)) = max(x,0) + log1p(exp(-|x|));  also returns sigmoid(x).
__device__ __forceinline__ float softplus_sigmoid(float x, float* sig) {
    float e = expf(-fabsf(x));
    float inv = 1.0f / (1.0f + e);
    *sig = x >= 0.0f ? inv : e * inv;
    return fmaxf(x, 0.0f) + log1pf(e);
}

// ---- streaming-rate math for the dense passes -----------------------------------------------------------
// The dense kernels touch every head logit once; the libm-accurate expf/log1pf/IEEE-division chain
// (~100 instructions per logit) made them issue-bound.  These versions cost ~10-30 instructions and stay
// within ~3 ulp: exp via ex2.approx with the rounding error of x*log2(e) compensated (2 FMAs), reciprocal
// via the correctly-rounded __frcp_rn, log1p on [0,1] via a degree-9 polynomial (max rel. error 1.7e-7).
__device__ __forceinline__ float ex2_approx(float y) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(y));
    return r;
}
__device__ __forceinline__ float exp_fast(float x) {  // |x| <= 87 (callers clamp)
    const float L2E = 1.44269502162933349609375f, L2E_LO = 1.92596299e-8f;
    const float y = x * L2E;
    float r = fmaf(x, L2E, -y);  // exact residual of the rounded product
    r = fmaf(x, L2E_LO, r);
    const float e = ex2_approx(y);
    return fmaf(e, r * 0.693147182464599609375f, e);
}
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// Eval-side sigmoid: decode, the candidate filter's exact tests, the listed scores and the box decode all use THIS
// function, and it is torch's CUDA sigmoid bit for bit -- ATen's sigmoid_kernel_cuda computes
// 1.0f / (1.0f + std::exp(-x)) in fp32 with libdevice's expf and an IEEE division -- so heads -> keep indices match the
// reference executed on the GPU exactly (layers.py:57-89).  The streaming pre-tests of the candidate kernel (logit
// windows) stay approximate; every decision inside a window is re-taken with this function.
__device__ __forceinline__ float sigmoid_eval(float x) { return 1.0f / (1.0f + expf(-x)); }
// softplus(x) and sigmoid(x) for the dense objectness pass
__device__ __forceinline__ float softplus_sigmoid_fast(float x, float* sig) {
    const float e = exp_fast(-fminf(fabsf(x), 87.0f));  // (0, 1]
    const float inv = __frcp_rn(1.0f + e);
    *sig = x >= 0.0f ? inv : e * inv;
    float q = -0.003256378462538123f;
    q = fmaf(q, e, 0.019907161593437195f);
    q = fmaf(q, e, -0.057064201682806015f);
    q = fmaf(q, e, 0.10614264756441116f);
    q = fmaf(q, e, -0.15311862528324127f);
    q = fmaf(q, e, 0.19678117334842682f);
    q = fmaf(q, e, -0.24954558908939362f);
    q = fmaf(q, e, 0.33330005407333374f);
    q = fmaf(q, e, -0.4999990463256836f);
    q = fmaf(q, e, 1.0f);
    return fmaf(q, e, fmaxf(x, 0.0f));  // max(x,0) + log1p(e)
}

// torch.remainder(x, 1) for floats: fmod then shift negatives by +1 (ATen remainder kernel).
__device__ __forceinline__ float remainder_one(float x) {
    float r = fmodf(x, 1.0f);
    if (r < 0.0f) r += 1.0f;
    return r;
}

// ---- ground-truth rows (yv5.py:68-121): xyxy pixels -> fp32 (cx,cy,w,h) in grid units --------------
struct GtRow {
    float cx, cy, w, h;
    int label;
};

// `inv_stride` is the python float 1/stride; fp32 boxes multiply by its fp32 rounding (torch casts the
// scalar to the tensor dtype), fp64 boxes by the double itself, then the row is stored as fp32.
__device__ __forceinline__ GtRow load_gt_row(const void* boxes, int boxes_dtype, const int64_t* labels, int n,
                                             double inv_stride_d) {
    GtRow r;
    if (boxes_dtype == KODB_BOXES_F64) {
        const double* b = reinterpret_cast<const double*>(boxes) + 4 * (int64_t)n;
        double x1 = b[0], y1 = b[1], x2 = b[2], y2 = b[3];
        double s = inv_stride_d;
        r.cx = (float)(((x1 + x2) / 2.0) * s);
        r.cy = (float)(((y1 + y2) / 2.0) * s);
        r.w = (float)((x2 - x1) * s);
        r.h = (float)((y2 - y1) * s);
    } else {
        const float4 b = reinterpret_cast<const float4*>(boxes)[n];
        const float inv_stride = (float)inv_stride_d;
        r.cx = ((b.x + b.z) / 2.0f) * inv_stride;
        r.cy = ((b.y + b.w) / 2.0f) * inv_stride;
        r.w = (b.z - b.x) * inv_stride;
        r.h = (b.w - b.y) * inv_stride;
    }
    // the reference stores the label in its fp32 row buffer and casts back with .long()
    r.label = (int)(float)labels[n];
    return r;
}

// anchor filter (yv5.py:160-176): max(max(rw,1/rw), max(rh,1/rh)) < thr, torch.max propagates NaN
__device__ __forceinline__ bool anchor_keeps(float w, float h, float aw, float ah, float thr) {
    float rw = w / aw, rh = h / ah;
    float iw = 1.0f / rw, ih = 1.0f / rh;
    float mw = (rw != rw || iw != iw) ? __int_as_float(0x7fc00000) : fmaxf(rw, iw);
    float mh = (rh != rh || ih != ih) ? __int_as_float(0x7fc00000) : fmaxf(rh, ih);
    float m = (mw != mw || mh != mh) ? __int_as_float(0x7fc00000) : fmaxf(mw, mh);
    return m < thr;
}

// neighbour masks (yv5.py:178-205) as bits: bit0 = base cell (always), bit1 = j, bit2 = k, bit3 = l, bit4 = m
__device__ __forceinline__ unsigned neighbour_bits(float cx, float cy, float fm_w, float fm_h) {
    float ix = fm_w - cx, iy = fm_h - cy;
    unsigned bits = 1u;
    if (remainder_one(cx) < 0.5f && cx > 1.0f) bits |= 2u;
    if (remainder_one(cy) < 0.5f && cy > 1.0f) bits |= 4u;
    if (remainder_one(ix) < 0.5f && ix > 1.0f) bits |= 8u;
    if (remainder_one(iy) < 0.5f && iy > 1.0f) bits |= 16u;
    return bits;
}

__device__ __forceinline__ float offset_x(int o) { return o == 1 ? 0.5f : (o == 3 ? -0.5f : 0.0f); }
__device__ __forceinline__ float offset_y(int o) { return o == 2 ? 0.5f : (o == 4 ? -0.5f : 0.0f); }

// float -> int64 like tensor.long() (truncation toward zero); the value range here is tiny
__device__ __forceinline__ long long trunc_ll(float v) { return (long long)v; }

__device__ __forceinline__ long long clamp_ll(long long v, long long lo, long long hi) {
    return v < lo ? lo : (v > hi ? hi : v);
}

// ---- eval decode of one anchor-cell (layers.py:55-63 + torchvision _box_cxcywh_to_xyxy) ----------
__device__ __forceinline__ float4 decode_box(float p0, float p1, float p2, float p3, float gx, float gy, float stride,
                                             float aw_px, float ah_px) {
    float cx = ((sigmoid_eval(p0) * 2.0f + gx) - 0.5f) * stride;
    float cy = ((sigmoid_eval(p1) * 2.0f + gy) - 0.5f) * stride;
    float sw = sigmoid_eval(p2) * 2.0f, sh = sigmoid_eval(p3) * 2.0f;
    float w = (sw * sw) * aw_px;
    float h = (sh * sh) * ah_px;
    float hw = 0.5f * w, hh = 0.5f * h;
    return make_float4(cx - hw, cy - hh, cx + hw, cy + hh);
}

// ---- asynchronous global -> shared copies (LDGSTS) -------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async16_s(unsigned smem_dst, const void* gmem_src) {  // shared-window address
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
// streaming variant: the source is read once, so its L2 lines are the first to go (keeps what other kernels of the
// pipeline still need -- candidate lists, class masks -- in the L2)
__device__ __forceinline__ unsigned long long l2_evict_first_policy(bool streaming = true) {
    unsigned long long pol;
    if (streaming) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void cp_async16_s_ef(unsigned smem_dst, const void* gmem_src, unsigned long long pol) {
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gmem_src), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---- reductions -----------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(KODB_FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(KODB_FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ int warp_excl_scan_i(int v, int lane) {
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(KODB_FULL_MASK, x, o);
        if (lane >= o) x += y;
    }
    return x - v;
}

// Block-wide exclusive scan of one int per thread (blockDim.x multiple of 32, <= 1024).
// `warp_tot` is shared scratch of >= 33 ints.  Returns the exclusive prefix; *total = block sum.
__device__ __forceinline__ int block_excl_scan(int v, int* warp_tot, int* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    int ex = warp_excl_scan_i(v, lane);
    __syncthreads();  // protect warp_tot reuse
    if (lane == 31) warp_tot[wid] = ex + v;
    __syncthreads();
    if (wid == 0) {
        int t = lane < nw ? warp_tot[lane] : 0;
        int e = warp_excl_scan_i(t, lane);
        warp_tot[lane] = e;
        if (lane == 31) warp_tot[32] = e + t;
    }
    __syncthreads();
    *total = warp_tot[32];
    return ex + warp_tot[wid];
}

}  // namespace kodb
