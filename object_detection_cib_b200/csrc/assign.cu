// kodb_assign — Yolov5LabelAssigner.__call__ (kod/core/label_assignment/yv5.py:298-319) materialised in
// the reference's row order: offset-major [base, j, k, l, m], then anchor, then target (flat batch order).
//
// One CTA per pyramid level.  The (offset, anchor, target) index space is walked in chunks of blockDim
// with an order-preserving compaction (warp shuffle scan + block scan), so the int64 index tensors come
// out bit-identical to the reference's two boolean-mask selects (yv5.py:176, :201).
#include "kodb_common.cuh"

namespace kodb {

struct AssignParams {
    kodb_level_cfg lv[KODB_NUM_LEVELS];
    kodb_assign_out out[KODB_NUM_LEVELS];
    const void* boxes;
    const int64_t* labels;
    const int32_t* img_offsets;
    int boxes_dtype;
    int T, B;
    float thr;
    int cap;
    int32_t* counts;
};

__device__ __forceinline__ int image_of_target(const int32_t* __restrict__ off, int B, int n) {
    // largest b with off[b] <= n  (empty images have off[b] == off[b+1] and are skipped naturally)
    int lo = 0, hi = B;  // invariant: off[lo] <= n < off[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (off[mid] <= n) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(1024, 1) assign_kernel(const __grid_constant__ AssignParams p) {
    __shared__ int scan_tmp[33];
    const int L = blockIdx.x;
    const kodb_level_cfg& lv = p.lv[L];
    const kodb_assign_out& out = p.out[L];
    const int A = lv.num_anchors, T = p.T;
    const double inv_stride = 1.0 / (double)lv.stride;
    const long long total = 5LL * A * T;
    int base = 0;
    for (long long start = 0; start < total; start += blockDim.x) {
        const long long i = start + threadIdx.x;
        int flag = 0, o = 0, k = 0, n = 0;
        GtRow g;
        if (i < total) {
            n = (int)(i % T);
            k = (int)((i / T) % A);
            o = (int)(i / ((long long)T * A));
            g = load_gt_row(p.boxes, p.boxes_dtype, p.labels, n, inv_stride);
            if (anchor_keeps(g.w, g.h, lv.anchor_grid[k][0], lv.anchor_grid[k][1], p.thr)) {
                unsigned bits = neighbour_bits(g.cx, g.cy, lv.fm_w, lv.fm_h);
                flag = (bits >> o) & 1u;
            }
        }
        int tot;
        const int pos = base + block_excl_scan(flag, scan_tmp, &tot);
        if (flag && pos < p.cap) {
            const long long gi = trunc_ll(g.cx - offset_x(o));
            const long long gj = trunc_ll(g.cy - offset_y(o));
            out.samples[pos] = image_of_target(p.img_offsets, p.B, n);
            out.anchors[pos] = k;
            out.grid_y[pos] = clamp_ll(gj, 0, lv.H - 1);
            out.grid_x[pos] = clamp_ll(gi, 0, lv.W - 1);
            out.labels[pos] = g.label;
            reinterpret_cast<float4*>(out.gt_boxes)[pos] = make_float4(g.cx - (float)gi, g.cy - (float)gj, g.w, g.h);
            reinterpret_cast<float2*>(out.anchor_wh)[pos] = make_float2(lv.anchor_grid[k][0], lv.anchor_grid[k][1]);
        }
        base += tot;
    }
    if (threadIdx.x == 0) p.counts[L] = base;
}

}  // namespace kodb

extern "C" int kodb_assign(const kodb_level_cfg* levels_h, const kodb_targets* t, float threshold,
                           const kodb_assign_out* out_h, int32_t cap, int32_t* counts, void* stream) {
    if (!levels_h || !t || !out_h || !counts || cap < 0) return KODB_ERR_INVALID_ARGUMENT;
    kodb::AssignParams p;
    for (int l = 0; l < KODB_NUM_LEVELS; ++l) {
        if (levels_h[l].num_anchors < 1 || levels_h[l].num_anchors > KODB_MAX_ANCHORS || levels_h[l].stride < 1)
            return KODB_ERR_INVALID_ARGUMENT;
        p.lv[l] = levels_h[l];
        p.out[l] = out_h[l];
    }
    p.boxes = t->boxes;
    p.labels = t->labels;
    p.img_offsets = t->img_offsets;
    p.boxes_dtype = t->boxes_dtype;
    p.T = t->num_targets;
    p.B = t->batch;
    p.thr = threshold;
    p.cap = cap;
    p.counts = counts;
    if (p.T < 0 || p.B < 0) return KODB_ERR_INVALID_ARGUMENT;
    kodb::assign_kernel<<<KODB_NUM_LEVELS, 1024, 0, (cudaStream_t)stream>>>(p);
    KODB_CHECK_LAUNCH();
    return KODB_OK;
}
