// kodb_train_tail — Yolov5Loss.forward (kod/lightning/experiments/yv5_baseline/loss.py:166-248) fused
// with its backward and with the target assignment it starts from (yv5.py:207-296).
//
// ONE launch with two kinds of CTAs (CTA-level specialisation):
//
//  * dense CTAs (blockIdx < dense_grid) — pure HBM streaming over the nine gradient tensors (the dense dY the 1x1
//    head convs consume), from an atomic chunk queue.  The objectness chunks lead the queue: read every objectness
//    logit ONCE (float4), write dobj = s*sigmoid(x) (the BCE gradient for target 0), keep the chunk's fp64 softplus
//    sum at a fixed place.  Then the zero-fill of the box / class planes -- large chunks first, small ones for the
//    last quarter of the bytes, so that the stream ends with fine-grained work.  They depend on nothing, never wait,
//    and carry ~all algorithmic bytes, 4*AC*(1+5+C) per image.  The first B of them also count N_l (the number of
//    assigned rows per level, needed by every gradient scale).
//
//  * sparse CTAs (one per (image, level), scheduled after the dense ones, resident from the start) — re-derive the
//    level's assigned rows of the image from its ground-truth list (order-preserving compaction: for a fixed anchor
//    the reference's row order is (offset, target)), gather 4 box + 1 obj + C cls logits per row, run the IoU-family
//    forward + analytic backward (iou.py:77-246), resolve duplicate cells (summed gradients; last-writer-wins
//    objectness target, index_put_ semantics; the target is NOT detached, loss.py:109-118), evaluate the
//    class BCE with pos_weight (loss.py:155-160) and publish their fp64 loss partials.  Their scattered HBM round
//    trips happen while the dense CTAs are still on the (light) objectness chunks; the rest is arithmetic and shared
//    memory, which the store flood does not slow down; the patch VALUES are computed ahead too.  Then every sparse
//    CTA waits for ONE word, the completion count of the dense CTAs, and stores its patches on top of the dense
//    output.  The CTA that published its partials last reduces all fp64 partials in a fixed order
//    (deterministic losses): its loads go out the moment the dense CTAs are through and are consumed behind its
//    patch stores.  Dense CTAs never wait and have the lowest block indices (dispatched first), so the
//    waits cannot deadlock; they are bounded all the same (status word, NaN losses).
//
// Head tensors that are not flat float4-divisible arrays (13x13 grids with odd batches, strided views) take a
// two-launch fallback with no in-kernel waiting at all: a grid-stride dense pass, then the sparse CTAs patch
// (stream order is the dependency).
//
// No host sync, no data-dependent shapes, CUDA-graph capturable.
#include <stdlib.h>

#include <mutex>

#include "kodb_common.cuh"

namespace kodb {

constexpr int TD_THREADS = 256;
constexpr int TS_THREADS = 256;
constexpr int TD_MAX_GRID = 148 * 8;
constexpr int TT_MAX_PARTIALS = 16384;  // fp64 objectness partials in the workspace ((chunk, warp) pairs; fallback: (CTA, level))

constexpr unsigned KODB_TT_BAD_LABEL = 1u;  // a target label outside [0, C): the reference raises IndexError (loss.py:145-152)
constexpr unsigned KODB_TT_TIMEOUT = 2u;    // a bounded wait expired (forward-progress assumption violated)

struct TrainParams {
    kodb_head_level head[KODB_NUM_LEVELS];
    kodb_grad_level grad[KODB_NUM_LEVELS];
    kodb_level_cfg lv[KODB_NUM_LEVELS];
    const void* boxes;
    const int64_t* labels;
    const int32_t* img_offsets;
    int boxes_dtype, T, B;
    int C, iou_type;
    float eps, thr;
    double lam_loc, lam_obj, lam_cls;  // effective lambdas (image-size / class-count scaling folded in)
    const float* pos_weight;
    float up_h[3];
    const float* up_dev;
    int skip_if_equal, has_grads;
    float* losses;
    int32_t* n_assigned;
    int contiguous[KODB_NUM_LEVELS];  // planar tensors of the level are flat arrays (fast path)
    int row_cap;                      // row capacity of a sparse CTA: 5 * A * max targets per image
    int cls_in_smem;                  // class-BCE gradients of the rows fit in shared memory
    int n_sparse;                     // 3 * B
    int dense_grid;
    int main_launch;                  // 1: the single-launch kernel (sparse CTAs wait for the dense ones); 0: fallback
    double* dense_partials;   // queue: softplus sums per (objectness chunk, warp); fallback: [dense_grid][3]
    double* dense_sums;       // queue: the three per-level totals of dense_partials (pre-reduced mid-stream)
    double* sparse_partials;  // [B*3][3]  (loc, obj correction, cls) per (image, level)
    // control words (workspace header), all zero on entry / exit
    unsigned int* queue;        // chunk tickets
    unsigned int* nl_ready;     // images whose N_l contribution has been added
    unsigned int* obj_done;     // objectness chunks written (gradients + fp64 partial)
    unsigned int* dense_done;   // dense CTAs that are through
    unsigned int* sparse_exit;  // sparse CTAs that published their partials
    unsigned int* sparse_exit2; // sparse CTAs that are completely done
    int* nl_pub;                // [3] N_l
    unsigned int* status;       // [0] KODB_TT_* bits of this launch, [1] sticky copy for kodb_train_status
    // Chunk queue of the dense stream (all levels flat).  Order: the objectness chunks of all levels, then the
    // zero-fill of the virtual concatenation  box0 | cls0 | box1 | cls1 | box2 | cls2  (float4 units) -- large chunks
    // up to z_split, small ones behind it, so that the stream ends with fine-grained work.
    int use_queue;
    int zero_delay_ns;                         // dense CTAs hold their first zero-fill chunk back until this long after their start
    int ch_obj;                                // float4 per objectness chunk
    int n_obj_chunks[KODB_NUM_LEVELS];
    int n_obj_total;
    long long zseg_begin[2 * KODB_NUM_LEVELS + 1];
    int ch_z1, ch_z2;                          // float4 per large / small zero-fill chunk
    long long z_split;                         // virtual offset where the small chunks begin (multiple of ch_z1)
    int n_z1, n_z2;
    int n_chunks;
    unsigned char* rows_global;                // row records of the sparse CTAs when they do not fit in shared memory
    long long rows_global_stride;
    long long* trace;         // debug: per-CTA phase stamps (globaltimer ns), NULL in production
};

__device__ __forceinline__ long long gtime() {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define TT_TRACE(slot)                                                                    \
    do {                                                                                  \
        if (p.trace && threadIdx.x == 0) p.trace[(int64_t)blockIdx.x * 8 + (slot)] = gtime(); \
    } while (0)

__device__ __forceinline__ bool load_upstream(const TrainParams& p, float up[3]) {
    if (p.up_dev) {
        up[0] = p.up_dev[0]; up[1] = p.up_dev[1]; up[2] = p.up_dev[2];
        if (p.skip_if_equal && up[0] == p.up_h[0] && up[1] == p.up_h[1] && up[2] == p.up_h[2]) return false;
    } else {
        up[0] = p.up_h[0]; up[1] = p.up_h[1]; up[2] = p.up_h[2];
    }
    return true;
}

__device__ __forceinline__ double block_sum_d(double v, double* sm /* >= 32 */) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) r += sm[i];  // fixed order
    return r;
}

__device__ __forceinline__ void st_zero4(float* p) {
    // streaming store: the gradients are consumed by a later kernel, not re-read here
    __stcs(reinterpret_cast<float4*>(p), make_float4(0.f, 0.f, 0.f, 0.f));
}

// ---- bounded waits ---------------------------------------------------------------------------------------------
// Everything a sparse CTA waits for is produced by dense CTAs, which never wait and precede the waiters in dispatch
// order, so the waits terminate.  They are bounded all the same: after ~2 s the waiter raises KODB_TT_TIMEOUT and
// carries on (the losses become NaN where still possible) instead of hanging the GPU.
constexpr long long TT_WAIT_NS = 2000000000ll;

__device__ __forceinline__ unsigned ld_acquire(const unsigned int* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void spin_until_ge(const unsigned int* word, unsigned want, unsigned int* status) {
    if (ld_acquire(word) >= want) return;
    const long long t0 = gtime();
    for (unsigned n = 1;; ++n) {
        __nanosleep(64);
        if (ld_acquire(word) >= want) return;
        if ((n & 1023u) == 0u && gtime() - t0 > TT_WAIT_NS) {
            atomicOr(status, KODB_TT_TIMEOUT);
            return;
        }
    }
}

// =========================================================================================================
// 1. sparse rows
// =========================================================================================================
__device__ __forceinline__ float w_gt(float a, float b) { return a > b ? 1.0f : (a == b ? 0.5f : 0.0f); }
__device__ __forceinline__ float w_lt(float a, float b) { return a < b ? 1.0f : (a == b ? 0.5f : 0.0f); }
__device__ __forceinline__ float sgn(float a) { return a > 0.0f ? 1.0f : (a < 0.0f ? -1.0f : 0.0f); }

// Value of the configured IoU metric for one assigned row and d(value)/d(box logits).
// Forward follows iou.py op for op (fp32, no FMA contraction); backward is the analytic gradient of
// exactly that expression with alpha treated as a constant (iou.py:241-242).
__device__ __forceinline__ float iou_metric(int type, float eps, const float p[4], float aw, float ah, float gx, float gy,
                                            float gw, float gh, float d[4]) {
    const float sx = sigmoid_f(p[0]), sy = sigmoid_f(p[1]), sw = sigmoid_f(p[2]), sh = sigmoid_f(p[3]);
    const float px = sx * 2.0f - 0.5f, py = sy * 2.0f - 0.5f;
    const float tw = sw * 2.0f, th = sh * 2.0f;
    const float pw = (tw * tw) * aw, ph = (th * th) * ah;
    const float x1 = px - 0.5f * pw, y1 = py - 0.5f * ph, x2 = px + 0.5f * pw, y2 = py + 0.5f * ph;
    const float x1g = gx - 0.5f * gw, y1g = gy - 0.5f * gh, x2g = gx + 0.5f * gw, y2g = gy + 0.5f * gh;

    const float tix = fminf(x2, x2g) - fmaxf(x1, x1g), tiy = fminf(y2, y2g) - fmaxf(y1, y1g);
    const float iw = fmaxf(tix, 0.0f), ih = fmaxf(tiy, 0.0f);
    const float inter = iw * ih;
    const float w1 = x2 - x1, h1 = y2 - y1, w2 = x2g - x1g, h2 = y2g - y1g;
    const float uni = (w1 * h1 + w2 * h2) - inter;
    const float ue = uni + eps;
    const float iou = inter / ue;
    float val = iou;

    // coefficients of d(val) on the intermediates
    float g_uni = -inter / (ue * ue);  // through iou
    float g_inter_direct = 1.0f / ue;
    float g_cw = 0.0f, g_ch = 0.0f, g_w1 = 0.0f, g_h1 = 0.0f;
    float gx1 = 0.0f, gy1 = 0.0f, gx2 = 0.0f, gy2 = 0.0f;

    if (type != KODB_IOU_IOU) {
        const float cw = fmaxf(x2, x2g) - fminf(x1, x1g), ch = fmaxf(y2, y2g) - fminf(y1, y1g);
        if (type == KODB_IOU_GIOU) {
            const float hull = cw * ch;
            const float num = hull - uni, den = hull + eps;
            val = iou - fabsf(num) / fabsf(den);
            const float dnum = sgn(num) / fabsf(den);
            const float dden = -fabsf(num) * sgn(den) / (den * den);
            const float g_hull = -(dnum + dden);
            g_uni += dnum;
            g_cw += g_hull * ch;
            g_ch += g_hull * cw;
        } else {
            const float diag2 = cw * cw + ch * ch;
            const float dx = (x1 + x2) / 2.0f - (x1g + x2g) / 2.0f, dy = (y1 + y2) / 2.0f - (y1g + y2g) / 2.0f;
            const float cd2 = dx * dx + dy * dy;
            const float de = diag2 + eps;
            const float D = cd2 / de;
            val = iou - D;
            const float g_cd2 = -1.0f / de, g_diag2 = cd2 / (de * de);
            gx1 += g_cd2 * dx;  // 2*dx * 1/2
            gx2 += g_cd2 * dx;
            gy1 += g_cd2 * dy;
            gy2 += g_cd2 * dy;
            g_cw += g_diag2 * 2.0f * cw;
            g_ch += g_diag2 * 2.0f * ch;
            if (type == KODB_IOU_CIOU) {
                const float kk = (float)(4.0 / (3.141592653589793 * 3.141592653589793));
                const float h1e = h1 + eps;
                const float r = w1 / h1e;
                const float delta = atanf(w2 / (h2 + eps)) - atanf(r);
                const float v = kk * (delta * delta);
                const float alpha = v / ((1.0f - iou) + v + eps);
                val = (iou - D) - alpha * v;
                const float g_r = alpha * kk * 2.0f * delta / (1.0f + r * r);  // -alpha * dv/ddelta * ddelta/dr, ddelta/dr = -1/(1+r^2)
                g_w1 += g_r / h1e;
                g_h1 += -g_r * w1 / (h1e * h1e);
            }
        }
        gx2 += g_cw * w_gt(x2, x2g);
        gx1 -= g_cw * w_lt(x1, x1g);
        gy2 += g_ch * w_gt(y2, y2g);
        gy1 -= g_ch * w_lt(y1, y1g);
    }
    // union = area1 + area2 - inter
    const float g_inter = g_inter_direct - g_uni;
    g_w1 += g_uni * h1;
    g_h1 += g_uni * w1;
    // inter = clamp(tix,0) * clamp(tiy,0)
    const float g_tix = (tix >= 0.0f) ? g_inter * ih : 0.0f;
    const float g_tiy = (tiy >= 0.0f) ? g_inter * iw : 0.0f;
    gx2 += g_tix * w_lt(x2, x2g);
    gx1 -= g_tix * w_gt(x1, x1g);
    gy2 += g_tiy * w_lt(y2, y2g);
    gy1 -= g_tiy * w_gt(y1, y1g);
    // w1 = x2 - x1, h1 = y2 - y1
    gx2 += g_w1;
    gx1 -= g_w1;
    gy2 += g_h1;
    gy1 -= g_h1;
    const float g_px = gx1 + gx2, g_py = gy1 + gy2;
    const float g_pw = 0.5f * (gx2 - gx1), g_ph = 0.5f * (gy2 - gy1);
    d[0] = g_px * 2.0f * (sx * (1.0f - sx));
    d[1] = g_py * 2.0f * (sy * (1.0f - sy));
    d[2] = g_pw * (4.0f * tw * aw) * (sw * (1.0f - sw));
    d[3] = g_ph * (4.0f * th * ah) * (sh * (1.0f - sh));
    return val;
}



// ---- losses: fp64 partials reduced in a fixed order (deterministic, independent of scheduling) ----------------------
// Queue mode, stage 1 (mid-stream, off the critical path): one sparse CTA folds the (objectness chunk, warp) softplus
// sums into three per-level totals as soon as every dense CTA has left the objectness chunks.
__device__ __noinline__ void reduce_dense_partials(const TrainParams& p, double* s_red9) {
    const int tid = threadIdx.x;
    constexpr int W = TD_THREADS / 32, U = 8;
    const int n = p.n_obj_total * W, e0 = p.n_obj_chunks[0] * W, e1 = e0 + p.n_obj_chunks[1] * W;
    double v0 = 0.0, v1 = 0.0, v2 = 0.0;
    for (int i0 = 0; i0 < n; i0 += U * TS_THREADS) {  // U independent L2 loads in flight per thread, then the adds in order
        double a[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = i0 + u * TS_THREADS + tid;
            a[u] = i < n ? __ldcg(p.dense_partials + i) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = i0 + u * TS_THREADS + tid;
            if (i < e0) v0 += a[u]; else if (i < e1) v1 += a[u]; else if (i < n) v2 += a[u];
        }
    }
    v0 = warp_sum(v0); v1 = warp_sum(v1); v2 = warp_sum(v2);
    __syncthreads();
    if ((tid & 31) == 0) { s_red9[(tid >> 5) * 3] = v0; s_red9[(tid >> 5) * 3 + 1] = v1; s_red9[(tid >> 5) * 3 + 2] = v2; }
    __syncthreads();
    if (tid < 3) {
        double t = 0.0;
        for (int w = 0; w < TS_THREADS / 32; ++w) t += s_red9[w * 3 + tid];  // warp order
        p.dense_sums[tid] = t;
    }
    __syncthreads();
}

// Stage 2 (the CTA that published its partials last): nine column sums -- (term, level) -- over the B sparse triples.
// Thread (column, slice) of the first 72 adds the images slice, slice + 8, ... in order; nine threads add the eight
// slices in order; three threads form the losses.  Split in two so that the caller can put independent work between the
// loads and their first use: fin_load issues every global load (one L2 round trip), fin_reduce consumes them.
constexpr int FIN_SLICES = 8, FIN_U = 8;
struct FinLoads {
    double a[FIN_U];  // sparse partials of images slice, slice + 8, ... (first FIN_U of them)
    double dsum;      // per-level objectness total (queue mode)
    int nl;           // N_l of the column's level
    unsigned status;
};
__device__ __forceinline__ void fin_load(const TrainParams& p, FinLoads& f) {
    const int tid = threadIdx.x;
    const int col = tid % 9, slice = tid / 9, term = col / 3, l = col % 3;
    f.nl = __ldcg(&p.nl_pub[l]);
    f.status = __ldcg(p.status);
    f.dsum = p.use_queue ? __ldcg(p.dense_sums + l) : 0.0;
#pragma unroll
    for (int u = 0; u < FIN_U; ++u) {
        const int b = slice + FIN_SLICES * u;
        f.a[u] = (slice < FIN_SLICES && b < p.B) ? __ldcg(p.sparse_partials + ((size_t)b * KODB_NUM_LEVELS + l) * 3 + term) : 0.0;
    }
}
__device__ __noinline__ void fin_reduce(const TrainParams& p, const FinLoads& f, double* s_red9 /* [8*9 + 9] */) {
    const int tid = threadIdx.x, C = p.C;
    const int col = tid % 9, slice = tid / 9, term = col / 3, l = col % 3;
    if (slice < FIN_SLICES) {
        double t = 0.0;
#pragma unroll
        for (int u = 0; u < FIN_U; ++u) t += f.a[u];  // (absent images contribute +0.0)
        for (int b = slice + FIN_SLICES * FIN_U; b < p.B; b += FIN_SLICES)  // batches larger than 64 images
            t += __ldcg(p.sparse_partials + ((size_t)b * KODB_NUM_LEVELS + l) * 3 + term);
        if (!p.use_queue && term == 1)  // fallback: (CTA, level) softplus sums of the dense launch
            for (int i = slice; i < p.dense_grid; i += FIN_SLICES) t += __ldcg(p.dense_partials + (size_t)i * 3 + l);
        s_red9[slice * 9 + col] = t;
    }
    __syncthreads();
    if (tid < 9) {  // one thread per (term, level): slice totals in slice order, then the level's mean
        double t = 0.0;
        for (int w = 0; w < FIN_SLICES; ++w) t += s_red9[w * 9 + tid];
        if (term == 1 && p.use_queue) t += f.dsum;  // sparse corrections first, then the dense softplus total
        const kodb_level_cfg& q = p.lv[l];
        const int nl = f.nl;
        const double nobj = (double)p.B * q.num_anchors * q.H * q.W;
        // 0/0 -> NaN: mean over an empty level (loss.py:96)
        s_red9[72 + tid] = term == 0 ? t / (double)nl : (term == 1 ? (double)q.obj_ratio * (t / nobj) : t / ((double)nl * (double)C));
        if (term == 0 && p.n_assigned) p.n_assigned[l] = nl;
    }
    __syncthreads();
    if (tid < 3) {
        const double lam = tid == 0 ? p.lam_loc : (tid == 1 ? p.lam_obj : p.lam_cls);
        double r = lam * ((s_red9[72 + 3 * tid] + s_red9[72 + 3 * tid + 1]) + s_red9[72 + 3 * tid + 2]);
        // a label outside [0, C) or an expired wait must not train silently
        if (f.status != 0u) r = __longlong_as_double(0x7ff8000000000000ll);
        p.losses[tid] = (float)r;
    }
}

// the last CTA of the last launch re-arms the control words (one thread)
__device__ __forceinline__ void rearm(const TrainParams& p) {
    p.status[1] = *p.status;  // sticky copy for the host (kodb_train_status); overwritten by every launch
    *p.status = 0u;
    *p.queue = 0u;
    *p.nl_ready = 0u;
    *p.obj_done = 0u;
    *p.dense_done = 0u;
    *p.sparse_exit = 0u;
    *p.sparse_exit2 = 0u;
    p.nl_pub[0] = 0; p.nl_pub[1] = 0; p.nl_pub[2] = 0;
}

// row record of the sparse CTAs (structure of arrays, capacity p.row_cap; shared memory or a workspace slice)
struct SparseRows {
    int* id;       // anchor * HW + cell
    int* label;
    float* val;    // IoU-family metric of the row
    float* xo;     // objectness logit at the cell
    float4* d;     // d(metric)/d(box logits), unscaled
    int* flags;    // bit0: first row of its cell ("leader"), bit1: last row of its cell (index_put_ winner)
    int* next;     // next row with the same cell (-1: none): the duplicate chain, in reference row order
    float* gcls;   // [row_cap * C] unscaled class-BCE gradient per (row, class) when cls_in_smem
};
constexpr size_t TT_ROW_BYTES = 6 * sizeof(int) + sizeof(float4);

__device__ __forceinline__ SparseRows carve_rows(unsigned char* smem_raw, int cap) {
    SparseRows r;
    r.d = reinterpret_cast<float4*>(smem_raw);
    r.id = reinterpret_cast<int*>(r.d + cap);
    r.label = r.id + cap;
    r.val = reinterpret_cast<float*>(r.label + cap);
    r.xo = r.val + cap;
    r.flags = reinterpret_cast<int*>(r.xo + cap);
    r.next = r.flags + cap;
    r.gcls = reinterpret_cast<float*>(r.next + cap);
    return r;
}

// unscaled class-BCE gradient of one (row, class): (1 - t) - lw * (1 - sigmoid(q))   
__device__ __forceinline__ float cls_grad_unscaled(float q, int label, int c, const float* pos_weight) {
    float sig;
    softplus_sigmoid(-q, &sig);  // sig = 1 - sigmoid(q)
    const float pw = pos_weight ? pos_weight[c] : 1.0f;
    const float tgt = (label == c) ? 1.0f : 0.0f;
    const float lw = 1.0f + (pw - 1.0f) * tgt;
    return (1.0f - tgt) - lw * sig;
}

// Both launches patch the gradient planes directly.  In the main kernel (p.main_launch) the patch waits for the dense
// CTAs of the same launch; in the fallback the dense pass is a launch of its own and stream order is the dependency.
__device__ __forceinline__ void sparse_cta(const TrainParams& p, const float up[3], unsigned char* smem_raw, int* scan_tmp,
                                           double* red_tmp, double* s_red9, int& s_last) {
    const int tid = threadIdx.x;
    const int sid = p.main_launch ? (int)blockIdx.x - p.dense_grid : (int)blockIdx.x;
    const int b = sid / KODB_NUM_LEVELS, L = sid - b * KODB_NUM_LEVELS;
    TT_TRACE(0);

    const kodb_level_cfg& lv = p.lv[L];
    const kodb_head_level& hd = p.head[L];
    const kodb_grad_level& gd = p.grad[L];
    const int W = lv.W, H = lv.H, HW = H * W, A = lv.num_anchors, C = p.C;
    const double inv_stride = 1.0 / (double)lv.stride;
    // row records: shared memory, or (crowded images: more rows than fit) this CTA's slice of the workspace
    const SparseRows R_ = carve_rows(p.rows_global ? p.rows_global + (size_t)sid * p.rows_global_stride : smem_raw, p.row_cap);

    double acc_loc = 0.0, acc_obj = 0.0, acc_cls = 0.0;

    // ---- rows of (image b, level L): index i = (k*5 + o)*Tb + n, i.e. (offset, target) order per anchor ----
    const int t0 = p.img_offsets[b], Tb = p.img_offsets[b + 1] - t0;
    const int total = 5 * A * Tb;
    int R = 0;
    bool bad_label = false;
    for (int start = 0; start < total; start += TS_THREADS) {
        const int i = start + tid;
        int flag = 0, cell = 0, k = 0;
        GtRow g;
        float gxf = 0.f, gyf = 0.f;
        if (i < total) {
            const int ko = i / Tb, n = t0 + (i - ko * Tb);
            k = ko / 5;
            const int o = ko - k * 5;
            g = load_gt_row(p.boxes, p.boxes_dtype, p.labels, n, inv_stride);
            if (g.label < 0 || g.label >= C) bad_label = true;  // loss.py:145-152 would raise IndexError
            if (anchor_keeps(g.w, g.h, lv.anchor_grid[k][0], lv.anchor_grid[k][1], p.thr) &&
                ((neighbour_bits(g.cx, g.cy, lv.fm_w, lv.fm_h) >> o) & 1u)) {
                const long long gi = trunc_ll(g.cx - offset_x(o)), gj = trunc_ll(g.cy - offset_y(o));
                cell = (int)(clamp_ll(gj, 0, H - 1) * W + clamp_ll(gi, 0, W - 1));
                flag = 1;
                gxf = g.cx - (float)gi;  // the gt box uses the UNclamped cell (yv5.py:281)
                gyf = g.cy - (float)gj;
            }
        }
        int tot;
        const int pos = R + block_excl_scan(flag, scan_tmp, &tot);
        TT_TRACE(1);
        if (flag) {
            const float aw = lv.anchor_grid[k][0], ah = lv.anchor_grid[k][1];
            const int64_t bo = (int64_t)b * hd.box_sb + (int64_t)k * hd.box_sa + cell;
            // every scattered HBM read of the row is issued here, in ONE round trip, while the dense CTAs are still on the
            // light objectness chunks: the class logits are parked in shared memory for the class phase further down
            const float* cq = hd.cls + (int64_t)b * hd.cls_sb + (int64_t)k * hd.cls_sa + cell;
            float pl[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) pl[q] = hd.box[bo + q * hd.box_sp];
            const float xo = hd.obj[(int64_t)b * hd.obj_sb + (int64_t)k * hd.obj_sa + cell];
            if (p.cls_in_smem) {
                for (int c0 = 0; c0 < C; c0 += 16) {  // up to 16 loads in flight, then the stores: one round trip per batch
                    float qv[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) qv[u] = c0 + u < C ? cq[(int64_t)(c0 + u) * hd.cls_sp] : 0.0f;
#pragma unroll
                    for (int u = 0; u < 16; ++u)
                        if (c0 + u < C) R_.gcls[(size_t)pos * C + c0 + u] = qv[u];
                }
            } else {
                for (int c = 0; c < C; ++c) asm volatile("prefetch.global.L2 [%0];" ::"l"(cq + (int64_t)c * hd.cls_sp));
            }
            float d[4];
            const float val = iou_metric(p.iou_type, p.eps, pl, aw, ah, gxf, gyf, g.w, g.h, d);
            acc_loc += (double)(1.0f - val);
            R_.id[pos] = k * HW + cell;
            R_.label[pos] = g.label;
            R_.val[pos] = val;
            R_.xo[pos] = xo;
            R_.d[pos] = make_float4(d[0], d[1], d[2], d[3]);
        }
        R += tot;
    }
    __syncthreads();
    TT_TRACE(4);
    // duplicate structure: leader / last flags and the chain to the next row of the same cell
    // (no early exits: the loads of a row's scan are independent and pipeline instead of forming a dependent chain)
    for (int r = tid; r < R; r += TS_THREADS) {
        const int id = R_.id[r];
        int before = 0, nx = 0x7fffffff;
#pragma unroll 4
        for (int q = 0; q < R; ++q) {
            const bool same = R_.id[q] == id;
            before |= (same && q < r) ? 1 : 0;
            nx = (same && q > r && q < nx) ? q : nx;
        }
        const int fl = (before ? 0 : 1) | (nx == 0x7fffffff ? 2 : 0);
        if (nx == 0x7fffffff) nx = -1;
        R_.flags[r] = fl;
        R_.next[r] = nx;
        if (fl & 2) acc_obj -= (double)(fmaxf(R_.val[r], 0.0f) * R_.xo[r]);  // BCE(x,t) = softplus(x) - t*x, last writer
    }
    TT_TRACE(5);
    // classification: one thread per (row, class); value now, gradient kept unscaled in shared memory
    for (int idx = tid; idx < R * C; idx += TS_THREADS) {
        const int r = idx / C, c = idx - r * C;
        const int id = R_.id[r];
        const int k = id / HW, cell = id - k * HW;
        const float q = p.cls_in_smem ? R_.gcls[idx]
                                      : hd.cls[(int64_t)b * hd.cls_sb + (int64_t)k * hd.cls_sa + (int64_t)c * hd.cls_sp + cell];
        float sig;
        const float spn = softplus_sigmoid(-q, &sig);  // -log sigmoid(q); sig = 1 - sigmoid(q)
        const float pw = p.pos_weight ? p.pos_weight[c] : 1.0f;
        const float tgt = (R_.label[r] == c) ? 1.0f : 0.0f;
        const float lw = 1.0f + (pw - 1.0f) * tgt;
        acc_cls += (double)((1.0f - tgt) * q + lw * spn);
        if (p.cls_in_smem) R_.gcls[idx] = (1.0f - tgt) - lw * sig;
    }
    TT_TRACE(2);

    // ---- loss partials -> workspace (fixed place per (image, level): the reduction order never depends on timing)
    const double pl_ = block_sum_d(acc_loc, red_tmp);
    const double po_ = block_sum_d(acc_obj, red_tmp);
    const double pc_ = block_sum_d(acc_cls, red_tmp);
    if (__syncthreads_or(bad_label) && tid == 0) atomicOr(p.status, KODB_TT_BAD_LABEL);
    if (tid == 0) {
        double* dst = p.sparse_partials + ((size_t)b * KODB_NUM_LEVELS + L) * 3;
        dst[0] = pl_; dst[1] = po_; dst[2] = pc_;
    }
    const double n_obj = (double)p.B * A * H * W;
    const float s_obj = (float)((double)up[1] * p.lam_obj * (double)lv.obj_ratio / n_obj);

    // ---- publish; the CTA that publishes last reduces the losses ---------------------------------------------------
    const unsigned nblocks = (unsigned)p.n_sparse;
    if (p.main_launch && sid == 0) {
        // sparse CTA 0 first folds the objectness partials (complete once every dense CTA has left the objectness
        // chunks, a few microseconds into the kernel); its own publication below then covers the three totals as well
        if (tid == 0) spin_until_ge(p.obj_done, (unsigned)p.dense_grid, p.status);
        __syncthreads();
        reduce_dense_partials(p, s_red9);
    }
    if (tid == 0) {
        __threadfence();
        const unsigned prev = atomicAdd(p.sparse_exit, 1u);
        __threadfence();  // (acquire side of the publication chain for the CTA that turns out to be last)
        s_last = (prev == nblocks - 1);
    }
    __syncthreads();
    TT_TRACE(3);
    if (p.main_launch) {
        // N_l is counted by the first B dense CTAs in their first microsecond
        if (tid == 0) spin_until_ge(p.nl_ready, (unsigned)p.B, p.status);
        __syncthreads();
    }
    const int NL = (int)__ldcg(&p.nl_pub[L]);
    const float s_loc = (float)((double)up[0] * p.lam_loc / (double)NL);
    const float s_cls = (float)((double)up[2] * p.lam_cls / ((double)NL * (double)C));
    if (p.has_grads) {
        // Patch values first (they only need N_l), parked in the row arrays that are dead by now: the box patch of a cell
        // in its leader's d, the objectness patch of a cell in its last row's label slot, the class patches in the leader's
        // gcls row -- so that nothing but stores is left once the dense CTAs are through.  (Rows that spilled to the
        // workspace / class gradients that are recomputed take the plain path below.)
        const bool pre = p.cls_in_smem != 0;
        if (pre) {
            for (int r = tid; r < R; r += TS_THREADS) {
                const int fl = R_.flags[r];
                if (fl & 2) R_.label[r] = __float_as_int(s_obj * (sigmoid_f(R_.xo[r]) - fmaxf(R_.val[r], 0.0f)));
                if (fl & 1) {
                    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int q = r; q >= 0; q = R_.next[q]) {  // every duplicate receives (and contributes) gradient
                        const float v = R_.val[q];
                        // dL/d(val): localisation mean(1 - val) and the non-detached objectness target
                        const float gv = -s_loc + (v >= 0.0f ? -R_.xo[q] * s_obj : 0.0f);
                        const float4 d4 = R_.d[q];
                        s.x += gv * d4.x; s.y += gv * d4.y; s.z += gv * d4.z; s.w += gv * d4.w;
                    }
                    R_.d[r] = s;  // (only this thread ever reads a leader's d again)
                }
            }
            for (int idx = tid; idx < R * C; idx += TS_THREADS) {
                const int r = idx / C, c = idx - r * C;
                if (!(R_.flags[r] & 1)) continue;
                float s = 0.0f;
                for (int q = r; q >= 0; q = R_.next[q]) s += R_.gcls[q * C + c] * s_cls;
                R_.gcls[idx] = s;  // (a leader's row is read by its own (row, class) threads only)
            }
        }
    }
    // ONE word says that every dense CTA is through: the patches must land on top of the dense output, and the
    // objectness partials of the loss reduction are complete.  (Forward-only calls: only the reducing CTA waits.)
    if (p.main_launch && (p.has_grads || s_last)) {
        if (tid == 0) spin_until_ge(p.dense_done, (unsigned)p.dense_grid, p.status);
    }
    __syncthreads();
    TT_TRACE(6);
    // Losses: the CTA that published its partials last reduces everything in a fixed order.  Its loads go out now and are
    // consumed behind its patch stores: one L2 round trip, overlapped with the patch.
    FinLoads fin;
    if (s_last) fin_load(p, fin);
    if (p.has_grads) {
        const bool pre = p.cls_in_smem != 0;
        for (int r = tid; r < R; r += TS_THREADS) {
            const int id = R_.id[r], fl = R_.flags[r];
            const int k = id / HW, cell = id - k * HW;
            if (fl & 2)  // index_put_ last-writer-wins target (loss.py:113-118)
                gd.obj[(int64_t)b * gd.obj_sb + (int64_t)k * gd.obj_sa + cell] =
                    pre ? __int_as_float(R_.label[r]) : s_obj * (sigmoid_f(R_.xo[r]) - fmaxf(R_.val[r], 0.0f));
            if (fl & 1) {
                float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
                if (pre) {
                    s = R_.d[r];
                } else {
                    for (int q = r; q >= 0; q = R_.next[q]) {
                        const float v = R_.val[q];
                        const float gv = -s_loc + (v >= 0.0f ? -R_.xo[q] * s_obj : 0.0f);
                        const float4 d4 = R_.d[q];
                        s.x += gv * d4.x; s.y += gv * d4.y; s.z += gv * d4.z; s.w += gv * d4.w;
                    }
                }
                float* gb = gd.box + (int64_t)b * gd.box_sb + (int64_t)k * gd.box_sa + cell;
                gb[0] = s.x; gb[gd.box_sp] = s.y; gb[2 * gd.box_sp] = s.z; gb[3 * gd.box_sp] = s.w;
            }
        }
        for (int idx = tid; idx < R * C; idx += TS_THREADS) {
            const int r = idx / C, c = idx - r * C;
            if (!(R_.flags[r] & 1)) continue;
            const int id = R_.id[r];
            const int k = id / HW, cell = id - k * HW;
            float s = 0.0f;
            if (pre) {
                s = R_.gcls[idx];
            } else {  // row list too large for shared-memory gradients: recompute from the logit
                const float qv = hd.cls[(int64_t)b * hd.cls_sb + (int64_t)k * hd.cls_sa + (int64_t)c * hd.cls_sp + cell];
                for (int q = r; q >= 0; q = R_.next[q]) s += cls_grad_unscaled(qv, R_.label[q], c, p.pos_weight) * s_cls;
            }
            gd.cls[(int64_t)b * gd.cls_sb + (int64_t)k * gd.cls_sa + (int64_t)c * gd.cls_sp + cell] = s;
        }
    }
    TT_TRACE(7);
    if (s_last) {
        fin_reduce(p, fin, s_red9);
        TT_TRACE(2);
    }
    // ---- the sparse CTA that finishes last re-arms the control words for the next launch: by then every dense CTA is
    // through as well (main launch: the sparse CTAs waited for them; fallback: an earlier launch)
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned prev2 = atomicAdd(p.sparse_exit2, 1u);
        if (prev2 == nblocks - 1) rearm(p);
    }
}

// =========================================================================================================
// 2. dense stream
// =========================================================================================================
// Dense stream over a chunk queue (order: see TrainParams).  An objectness chunk reads its logits (all loads of the
// thread issued first), writes dobj = s*sigmoid(x) and stores the fp64 softplus sums of its eight warps at fixed places
// (dense_partials[chunk][warp]) -- the finaliser adds the partials in (chunk, warp) order, so the loss does not depend
// on which CTA took which chunk.  The next ticket is requested before the current chunk is processed (the atomic's round trip is
// hidden).
__device__ __forceinline__ void dense_queue(const TrainParams& p, const float up[3], double* red_tmp, int* ticket_sm) {
    constexpr int OBJ_PER_THREAD = 4, ZERO_PER_THREAD = 8;  // float4 per thread and round (chunk sizes are multiples)
    const int tid = threadIdx.x;
    const long long z_total = p.zseg_begin[2 * KODB_NUM_LEVELS];
    // the first chunk of a CTA is its block index (no atomic); tickets from the counter start behind those
    unsigned next = blockIdx.x;
    bool delayed = false, obj_reported = false;
    const long long t_start = gtime();
    for (;;) {
        __syncthreads();  // the previous round's readers of ticket_sm are done
        if (tid == 0) {
            ticket_sm[0] = (int)next;
            if ((int)next < p.n_chunks) next = (unsigned)p.dense_grid + atomicAdd(p.queue, 1u);  // prefetch the next ticket
        }
        __syncthreads();
        const int c = ticket_sm[0];
        if (c >= p.n_chunks) {
            if (!obj_reported && tid == 0) {
                __threadfence();
                atomicAdd(p.obj_done, 1u);
            }
            break;
        }
        if (c < p.n_obj_total) {
            int L = 0, cl = c;
            while (cl >= p.n_obj_chunks[L]) { cl -= p.n_obj_chunks[L]; ++L; }
            const kodb_level_cfg& lv = p.lv[L];
            const int64_t n_obj = (int64_t)p.B * lv.num_anchors * lv.H * lv.W;
            const int64_t n4 = n_obj >> 2, i0 = (int64_t)cl * p.ch_obj;
            const int64_t i1 = i0 + p.ch_obj < n4 ? i0 + p.ch_obj : n4;
            const float s_obj = (float)((double)up[1] * p.lam_obj * (double)lv.obj_ratio / (double)n_obj);
            const float4* src = reinterpret_cast<const float4*>(p.head[L].obj);
            float4* dst = reinterpret_cast<float4*>(p.grad[L].obj);
            float a0 = 0.0f;
            for (int64_t base = i0; base < i1; base += OBJ_PER_THREAD * TD_THREADS) {
                float4 x[OBJ_PER_THREAD];
#pragma unroll
                for (int u = 0; u < OBJ_PER_THREAD; ++u) {
                    const int64_t i = base + u * TD_THREADS + tid;
                    x[u] = i < i1 ? __ldcs(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < OBJ_PER_THREAD; ++u) {
                    const int64_t i = base + u * TD_THREADS + tid;
                    if (i < i1) {
                        float s0, s1, s2, s3;
                        const float v0 = softplus_sigmoid_fast(x[u].x, &s0), v1 = softplus_sigmoid_fast(x[u].y, &s1);
                        const float v2 = softplus_sigmoid_fast(x[u].z, &s2), v3 = softplus_sigmoid_fast(x[u].w, &s3);
                        a0 += (v0 + v1) + (v2 + v3);
                        if (p.has_grads) __stcs(dst + i, make_float4(s_obj * s0, s_obj * s1, s_obj * s2, s_obj * s3));
                    }
                }
            }
            // fp64 partial per (chunk, warp) at a fixed place: no block barrier on the streaming path; published by the
            // release at the end of this CTA (the finaliser runs behind the completion count of all dense CTAs)
            const double sum = warp_sum((double)a0);
            if ((tid & 31) == 0) p.dense_partials[(size_t)c * (TD_THREADS / 32) + (tid >> 5)] = sum;
            continue;
        }
        if (!obj_reported) {
            // this CTA's objectness chunks (gradients + fp64 partials) are out: one release per CTA, at the hand-over
            // from the light objectness phase to the zero-fill
            obj_reported = true;
            if (tid == 0) {
                __threadfence();
                atomicAdd(p.obj_done, 1u);
            }
        }
        if (p.zero_delay_ns > 0 && !delayed) {
            // the scattered reads of the sparse CTAs are issued in the first microseconds; under the store flood each of
            // them takes several times longer, so the flood is held back for a moment (time-based: no dependency, no
            // deadlock, whatever else shares the GPU)
            delayed = true;
            while (gtime() - t_start < p.zero_delay_ns) __nanosleep(100);
        }
        const int cz = c - p.n_obj_total;
        const long long lo = cz < p.n_z1 ? (long long)cz * p.ch_z1 : p.z_split + (long long)(cz - p.n_z1) * p.ch_z2;
        const long long hi0 = lo + (cz < p.n_z1 ? p.ch_z1 : p.ch_z2);
        const long long hi = hi0 < z_total ? hi0 : z_total;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        int sg = 0;
        while (p.zseg_begin[sg + 1] <= lo) ++sg;
        for (long long a = lo; a < hi; ++sg) {  // a chunk may straddle the end of a tensor
            const long long e = p.zseg_begin[sg + 1] < hi ? p.zseg_begin[sg + 1] : hi;
            float4* dst = reinterpret_cast<float4*>((sg & 1) ? p.grad[sg >> 1].cls : p.grad[sg >> 1].box) + (a - p.zseg_begin[sg]);
            const int n = (int)(e - a);
            for (int base = tid; base < n; base += ZERO_PER_THREAD * TD_THREADS) {
#pragma unroll
                for (int u = 0; u < ZERO_PER_THREAD; ++u)
                    if (base + u * TD_THREADS < n) __stcs(dst + base + u * TD_THREADS, z);
            }
            a = e;
        }
    }
}

// N_l = number of assigned rows per level over the whole batch: CTA i counts image i, i + grid, ... (a handful of
// targets each) and accumulates with integer atomics
__device__ __forceinline__ void count_assigned(const TrainParams& p, int* cnt_tmp, int first, int stride) {
    for (int img = first; img < p.B; img += stride) {
        if (threadIdx.x < KODB_NUM_LEVELS) cnt_tmp[threadIdx.x] = 0;
        __syncthreads();
        const int t0 = p.img_offsets[img], Tb = p.img_offsets[img + 1] - t0;
        for (int i = threadIdx.x; i < Tb * KODB_NUM_LEVELS; i += TD_THREADS) {
            const int n = t0 + i / KODB_NUM_LEVELS, l = i % KODB_NUM_LEVELS;
            const kodb_level_cfg& lv = p.lv[l];
            GtRow g = load_gt_row(p.boxes, p.boxes_dtype, p.labels, n, 1.0 / (double)lv.stride);
            const int nb = __popc(neighbour_bits(g.cx, g.cy, lv.fm_w, lv.fm_h));
            int c = 0;
            for (int k = 0; k < lv.num_anchors; ++k)
                if (anchor_keeps(g.w, g.h, lv.anchor_grid[k][0], lv.anchor_grid[k][1], p.thr)) c += nb;
            if (c) atomicAdd(&cnt_tmp[l], c);
        }
        __syncthreads();
        if (threadIdx.x < KODB_NUM_LEVELS && cnt_tmp[threadIdx.x]) atomicAdd(&p.nl_pub[threadIdx.x], cnt_tmp[threadIdx.x]);
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();  // the three level counters before the "image counted" word
            atomicAdd(p.nl_ready, 1u);
        }
    }
}

__device__ __forceinline__ void dense_cta_queue(const TrainParams& p, const float up[3], double* red_tmp, int* ticket_sm) {
    TT_TRACE(0);
    count_assigned(p, ticket_sm, blockIdx.x, p.dense_grid);
    dense_queue(p, up, red_tmp, ticket_sm);
    __syncthreads();  // every thread's stores are issued
    if (threadIdx.x == 0) {
        __threadfence();  // release: stores of this CTA visible device-wide before the counter moves
        atomicAdd(p.dense_done, 1u);
    }
    TT_TRACE(1);
}

__global__ void __launch_bounds__(TS_THREADS, 4) train_tail_kernel(const __grid_constant__ TrainParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int scan_tmp[33];
    __shared__ double red_tmp[32];
    __shared__ double s_red9[(TS_THREADS / 32) * 9 + 9];
    __shared__ int s_last;
    float up[3];
    // Programmatic dependent launch (sm_90+): let the next kernel of the stream become resident while this one runs, and
    // do not touch global memory before the previous kernel of the stream has completed and flushed.  Both are no-ops
    // when the launches are not programmatically serialised.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (!load_upstream(p, up)) return;
    if ((int)blockIdx.x < p.dense_grid) dense_cta_queue(p, up, red_tmp, scan_tmp);
    else sparse_cta(p, up, smem_raw, scan_tmp, red_tmp, s_red9, s_last);
}

// ---- fallback (tensors that are not flat float4 arrays): dense grid-stride pass, then the sparse launch -------------
__device__ __forceinline__ void zero_flat(float* ptr, int64_t n, int64_t gtid, int64_t gsize) {
    const int64_t n4 = n >> 2;
    for (int64_t i = gtid; i < n4; i += gsize) st_zero4(ptr + 4 * i);
    for (int64_t i = (n4 << 2) + gtid; i < n; i += gsize) ptr[i] = 0.0f;
}

// generic (strided) addressing of element i of a logical (B, A, P, HW) tensor
__device__ __forceinline__ int64_t strided_off(int64_t i, int HW, int P, int A, int64_t sb, int64_t sa, int64_t sp) {
    const int cell = (int)(i % HW);
    int64_t r = i / HW;
    const int pl = (int)(r % P);
    r /= P;
    const int a = (int)(r % A);
    const int64_t b = r / A;
    return b * sb + a * sa + (int64_t)pl * sp + cell;
}

__global__ void __launch_bounds__(TD_THREADS, 4) train_dense_fallback_kernel(const __grid_constant__ TrainParams p) {
    __shared__ int cnt_tmp[KODB_NUM_LEVELS];
    __shared__ double red_tmp[32];
    float up[3];
    if (!load_upstream(p, up)) return;
    count_assigned(p, cnt_tmp, blockIdx.x, gridDim.x);
    const int64_t gtid = (int64_t)blockIdx.x * TD_THREADS + threadIdx.x;
    const int64_t gsize = (int64_t)gridDim.x * TD_THREADS;
    double acc[KODB_NUM_LEVELS];
#pragma unroll
    for (int L = 0; L < KODB_NUM_LEVELS; ++L) {
        const kodb_level_cfg& lv = p.lv[L];
        const kodb_head_level& hd = p.head[L];
        const kodb_grad_level& gd = p.grad[L];
        const int HW = lv.H * lv.W, A = lv.num_anchors;
        const int64_t n_obj = (int64_t)p.B * A * HW;
        const float s_obj = (float)((double)up[1] * p.lam_obj * (double)lv.obj_ratio / (double)n_obj);
        float a0 = 0.0f;
        if (p.contiguous[L]) {
            // objectness: read once, BCE(x, 0) value + gradient
            const int64_t n4 = n_obj >> 2;
            for (int64_t i = gtid; i < n4; i += gsize) {
                const float4 x = __ldcs(reinterpret_cast<const float4*>(hd.obj) + i);
                float4 g;
                float s0, s1, s2, s3;
                const float v0 = softplus_sigmoid_fast(x.x, &s0), v1 = softplus_sigmoid_fast(x.y, &s1);
                const float v2 = softplus_sigmoid_fast(x.z, &s2), v3 = softplus_sigmoid_fast(x.w, &s3);
                a0 += (v0 + v1) + (v2 + v3);
                if (p.has_grads) {
                    g.x = s_obj * s0; g.y = s_obj * s1; g.z = s_obj * s2; g.w = s_obj * s3;
                    __stcs(reinterpret_cast<float4*>(gd.obj) + i, g);
                }
            }
            for (int64_t i = (n4 << 2) + gtid; i < n_obj; i += gsize) {
                float sg;
                a0 += softplus_sigmoid_fast(hd.obj[i], &sg);
                if (p.has_grads) gd.obj[i] = s_obj * sg;
            }
            if (p.has_grads) {
                zero_flat(gd.box, n_obj * 4, gtid, gsize);
                zero_flat(gd.cls, n_obj * p.C, gtid, gsize);
            }
        } else {
            for (int64_t i = gtid; i < n_obj; i += gsize) {
                float sg;
                a0 += softplus_sigmoid_fast(hd.obj[strided_off(i, HW, 1, A, hd.obj_sb, hd.obj_sa, 0)], &sg);
                if (p.has_grads) gd.obj[strided_off(i, HW, 1, A, gd.obj_sb, gd.obj_sa, 0)] = s_obj * sg;
            }
            if (p.has_grads) {
                for (int64_t i = gtid; i < n_obj * 4; i += gsize)
                    gd.box[strided_off(i, HW, 4, A, gd.box_sb, gd.box_sa, gd.box_sp)] = 0.0f;
                for (int64_t i = gtid; i < n_obj * p.C; i += gsize)
                    gd.cls[strided_off(i, HW, p.C, A, gd.cls_sb, gd.cls_sa, gd.cls_sp)] = 0.0f;
            }
        }
        acc[L] = (double)a0;
    }
#pragma unroll
    for (int L = 0; L < KODB_NUM_LEVELS; ++L) {
        const double s = block_sum_d(acc[L], red_tmp);
        if (threadIdx.x == 0) p.dense_partials[(size_t)blockIdx.x * 3 + L] = s;
    }
}

__global__ void __launch_bounds__(TS_THREADS, 4) train_sparse_fallback_kernel(const __grid_constant__ TrainParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int scan_tmp[33];
    __shared__ double red_tmp[32];
    __shared__ double s_red9[(TS_THREADS / 32) * 9 + 9];
    __shared__ int s_last;
    float up[3];
    if (!load_upstream(p, up)) return;
    sparse_cta(p, up, smem_raw, scan_tmp, red_tmp, s_red9, s_last);
}

static long long* g_tt_trace = nullptr;
static bool aligned16(const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15u) == 0; }

}  // namespace kodb

// debug hook (not part of include/kodb.h): device buffer of (grid)*8 int64 phase stamps per CTA
extern "C" void kodb_debug_set_train_trace(void* ptr) { kodb::g_tt_trace = reinterpret_cast<long long*>(ptr); }

static size_t tt_align(size_t v) { return (v + 255) & ~(size_t)255; }

namespace kodb {
constexpr size_t TT_SMEM_ROWS_MAX = 40 * 1024;  // row records beyond this go to the workspace (keeps the dense CTAs' occupancy)

static size_t tt_row_cap(int num_anchors, int max_per_image) {
    return (size_t)5 * num_anchors * (max_per_image > 0 ? max_per_image : 1);
}
// bytes of workspace row records per sparse CTA, 0 when the rows fit in shared memory
static size_t tt_rows_spill(int num_anchors, int max_per_image) {
    const size_t cap = tt_row_cap(num_anchors, max_per_image);
    return cap * TT_ROW_BYTES > TT_SMEM_ROWS_MAX ? tt_align(cap * TT_ROW_BYTES) : 0;
}
}  // namespace kodb

extern "C" int64_t kodb_train_workspace_bytes(int32_t batch, int32_t max_targets_per_image, int32_t max_anchors,
                                              int32_t num_classes) {
    const size_t b = (size_t)(batch > 0 ? batch : 1);
    const int a = max_anchors > 0 ? max_anchors : KODB_MAX_ANCHORS;
    (void)num_classes;  // (reserved: the class-gradient rows are recomputed, not stored, when the rows spill)
    return (int64_t)(256 + tt_align((size_t)kodb::TT_MAX_PARTIALS * sizeof(double)) + tt_align(b * 9 * sizeof(double)) +
                     b * KODB_NUM_LEVELS * kodb::tt_rows_spill(a, max_targets_per_image));
}

// Re-arm a workspace after a failed or aborted launch (or before first use instead of a memset by the caller):
// asynchronous zero-fill on `stream`.  Works for every kodb workspace (train, eval, AP).
extern "C" int kodb_workspace_reset(void* workspace, int64_t workspace_bytes, void* stream) {
    if (!workspace || workspace_bytes < 0) return KODB_ERR_INVALID_ARGUMENT;
    if (cudaMemsetAsync(workspace, 0, (size_t)workspace_bytes, (cudaStream_t)stream) != cudaSuccess) return KODB_ERR_LAUNCH;
    return KODB_OK;
}

// Status bits of the most recent kodb_train_tail launch on this workspace (device word -> *status_h; synchronises
// `stream`): bit 0 = a target label outside [0, C) (the reference raises IndexError, loss.py:145-152; the losses
// of that launch are NaN), bit 1 = a bounded wait inside the kernel expired.
extern "C" int kodb_train_status(const void* workspace, int32_t* status_h, void* stream) {
    if (!workspace || !status_h) return KODB_ERR_INVALID_ARGUMENT;
    unsigned v = 0;
    if (cudaMemcpyAsync(&v, reinterpret_cast<const char*>(workspace) + 68, sizeof(v), cudaMemcpyDeviceToHost,
                        (cudaStream_t)stream) != cudaSuccess ||
        cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess)
        return KODB_ERR_LAUNCH;
    *status_h = (int32_t)v;
    return KODB_OK;
}

extern "C" int kodb_train_tail(const kodb_head_level* heads_h, const kodb_level_cfg* levels_h,
                               const kodb_targets* t, const kodb_loss_cfg* cfg, const float* pos_weight,
                               const float* upstream_h, const float* upstream_dev, int32_t skip_if_equal,
                               const kodb_grad_level* grads_h, float* losses, int32_t* n_assigned,
                               void* workspace, int64_t workspace_bytes, void* stream) {
    using namespace kodb;
    if (!heads_h || !levels_h || !t || !cfg || !upstream_h || !losses || !workspace) return KODB_ERR_INVALID_ARGUMENT;
    if (cfg->num_classes < 1 || cfg->num_classes > KODB_MAX_CLASSES) return KODB_ERR_INVALID_ARGUMENT;
    if (cfg->iou_type < KODB_IOU_IOU || cfg->iou_type > KODB_IOU_CIOU) return KODB_ERR_INVALID_ARGUMENT;
    if (t->batch < 1 || t->batch > 65535 || t->num_targets < 0 || t->max_per_image < 0) return KODB_ERR_INVALID_ARGUMENT;
    TrainParams p;
    int amax = 1;
    for (int l = 0; l < KODB_NUM_LEVELS; ++l) {
        if (levels_h[l].num_anchors < 1 || levels_h[l].num_anchors > KODB_MAX_ANCHORS) return KODB_ERR_INVALID_ARGUMENT;
        if (levels_h[l].H < 1 || levels_h[l].W < 1) return KODB_ERR_INVALID_ARGUMENT;
        p.head[l] = heads_h[l];
        p.lv[l] = levels_h[l];
        if (grads_h) p.grad[l] = grads_h[l]; else p.grad[l] = kodb_grad_level{};
        if (levels_h[l].num_anchors > amax) amax = levels_h[l].num_anchors;
    }
    if (workspace_bytes < kodb_train_workspace_bytes(t->batch, t->max_per_image, amax, cfg->num_classes)) return KODB_ERR_CAPACITY;
    p.boxes = t->boxes;
    p.labels = t->labels;
    p.img_offsets = t->img_offsets;
    p.boxes_dtype = t->boxes_dtype;
    p.T = t->num_targets;
    p.B = t->batch;
    p.C = cfg->num_classes;
    p.iou_type = cfg->iou_type;
    p.eps = cfg->eps;
    p.thr = cfg->assign_threshold;
    // loss.py:231-237: python-float scaling of the lambdas
    p.lam_loc = (double)cfg->lambda_localization;
    p.lam_obj = (double)cfg->lambda_objectness * (((double)cfg->img_w / 640.0) * ((double)cfg->img_w / 640.0));
    p.lam_cls = (double)cfg->lambda_classification * ((double)cfg->num_classes / 80.0);
    p.pos_weight = pos_weight;
    for (int i = 0; i < 3; ++i) p.up_h[i] = upstream_h[i];
    p.up_dev = upstream_dev;
    p.skip_if_equal = skip_if_equal;
    p.has_grads = grads_h != nullptr;
    p.losses = losses;
    p.n_assigned = n_assigned;
    p.n_sparse = p.B * KODB_NUM_LEVELS;
    for (int l = 0; l < KODB_NUM_LEVELS; ++l) {
        const int64_t HW = (int64_t)p.lv[l].H * p.lv[l].W;
        const int A = p.lv[l].num_anchors;
        const kodb_head_level& h = p.head[l];
        bool ok = h.obj_sa == HW && h.obj_sb == A * HW && aligned16(h.obj);
        if (ok && p.has_grads) {
            const kodb_grad_level& g = p.grad[l];
            ok = g.obj_sa == HW && g.obj_sb == A * HW && g.box_sp == HW && g.box_sa == 4 * HW && g.box_sb == 4 * A * HW &&
                 g.cls_sp == HW && g.cls_sa == (int64_t)p.C * HW && g.cls_sb == (int64_t)p.C * A * HW && aligned16(g.obj) &&
                 aligned16(g.box) && aligned16(g.cls);
        }
        p.contiguous[l] = ok ? 1 : 0;
    }
    // chunk queue of the dense stream (needs every level flat and float4-divisible)
    p.use_queue = 1;
    for (int l = 0; l < KODB_NUM_LEVELS; ++l) {
        const int64_t n_obj = (int64_t)p.B * p.lv[l].num_anchors * p.lv[l].H * p.lv[l].W;
        if (!p.contiguous[l] || (n_obj & 3)) p.use_queue = 0;
    }
    // launch geometry / tuning knobs cached per process; shared by every host thread that calls into the library
    static std::mutex cache_mu;
    std::lock_guard<std::mutex> cache_lock(cache_mu);
    static int c_sms[64] = {0};
    static size_t c_smem[2][64] = {{0}, {0}};
    static int per_sm = -1, tail_pct = -1, ch2_mult = 0, zero_delay = 0;
    if (per_sm < 0) {  // tuning knobs; defaults chosen on B200
        const char* e = getenv("KODB_TRAIN_DENSE_PER_SM");
        per_sm = e ? atoi(e) : 0;  // 0: chosen per launch below
        if (per_sm < 0 || per_sm > 8) per_sm = 0;
        e = getenv("KODB_TRAIN_TAIL_PCT");
        tail_pct = e ? atoi(e) : 25;
        if (tail_pct < 0 || tail_pct > 100) tail_pct = 25;
        e = getenv("KODB_TRAIN_ZERO_DELAY_NS");
        zero_delay = e ? atoi(e) : 0;
        if (zero_delay < 0 || zero_delay > 100000) zero_delay = 0;
        e = getenv("KODB_TRAIN_TAIL_CHUNK");  // float4 per thread of a small chunk
        ch2_mult = e ? atoi(e) : 4;
        if (ch2_mult < 1 || ch2_mult > 16) ch2_mult = 4;
    }
    p.n_chunks = p.n_obj_total = p.n_z1 = p.n_z2 = 0;
    p.ch_obj = 4 * TD_THREADS;         // 16 KB of logits in, 16 KB of gradients out
    p.ch_z1 = 16 * TD_THREADS;         // 64 KB: few enough tickets for one atomic counter ...
    p.ch_z2 = ch2_mult * TD_THREADS;   // ... 16 KB for the last part of the stream: fine-grained balance at the end
    p.z_split = 0;
    p.zero_delay_ns = zero_delay;
    for (int i = 0; i <= 2 * KODB_NUM_LEVELS; ++i) p.zseg_begin[i] = 0;
    for (int l = 0; l < KODB_NUM_LEVELS; ++l) p.n_obj_chunks[l] = 0;
    if (p.use_queue) {
        int64_t tot4 = 0;
        for (int l = 0; l < KODB_NUM_LEVELS; ++l) tot4 += ((int64_t)p.B * p.lv[l].num_anchors * p.lv[l].H * p.lv[l].W) >> 2;
        // objectness chunks are the fp64 partials of the finaliser: keep their number inside the workspace
        const int64_t max_chunks = TT_MAX_PARTIALS / (TD_THREADS / 32) - KODB_NUM_LEVELS;
        while ((tot4 + p.ch_obj - 1) / p.ch_obj > max_chunks) p.ch_obj *= 2;
        long long z = 0;
        for (int l = 0; l < KODB_NUM_LEVELS; ++l) {
            const int64_t n_obj = (int64_t)p.B * p.lv[l].num_anchors * p.lv[l].H * p.lv[l].W;
            const int64_t n4 = n_obj >> 2;
            p.n_obj_chunks[l] = (int)((n4 + p.ch_obj - 1) / p.ch_obj);
            p.n_obj_total += p.n_obj_chunks[l];
            p.zseg_begin[2 * l] = z;
            if (p.has_grads) z += n_obj;     // box: 4 planes of n_obj floats = n_obj float4
            p.zseg_begin[2 * l + 1] = z;
            if (p.has_grads) z += n4 * p.C;  // cls: C planes
        }
        p.zseg_begin[2 * KODB_NUM_LEVELS] = z;
        for (;;) {
            p.z_split = (z / 100 * (100 - tail_pct)) / p.ch_z1 * p.ch_z1;
            p.n_z1 = (int)(p.z_split / p.ch_z1);
            p.n_z2 = (int)((z - p.z_split + p.ch_z2 - 1) / p.ch_z2);
            if ((long long)p.n_obj_total + p.n_z1 + p.n_z2 < (1ll << 30)) break;
            p.ch_z1 *= 2;
            p.ch_z2 *= 2;
        }
        p.n_chunks = p.n_obj_total + p.n_z1 + p.n_z2;
    }
    p.row_cap = (int)tt_row_cap(amax, t->max_per_image);
    const size_t spill = tt_rows_spill(amax, t->max_per_image);
    size_t smem = spill ? 0 : (size_t)p.row_cap * TT_ROW_BYTES;
    const size_t cls_bytes = (size_t)p.row_cap * p.C * sizeof(float);
    p.cls_in_smem = (!spill && smem + cls_bytes <= 46 * 1024) ? 1 : 0;  // (+ ~1 KB of static shared memory: no opt-in needed)
    if (p.cls_in_smem) smem += cls_bytes;
    // control words at fixed offsets, so a cached workspace stays valid when the batch size changes
    char* w = reinterpret_cast<char*>(workspace);
    // (words that are polled sit in 32-byte sectors of their own, away from the ones hammered by atomics)
    p.queue = reinterpret_cast<unsigned int*>(w);
    p.nl_pub = reinterpret_cast<int*>(w + 32);
    p.status = reinterpret_cast<unsigned int*>(w + 64);
    p.sparse_exit = reinterpret_cast<unsigned int*>(w + 96);
    p.sparse_exit2 = reinterpret_cast<unsigned int*>(w + 100);
    p.nl_ready = reinterpret_cast<unsigned int*>(w + 128);
    p.obj_done = reinterpret_cast<unsigned int*>(w + 160);
    p.dense_done = reinterpret_cast<unsigned int*>(w + 192);
    p.dense_sums = reinterpret_cast<double*>(w + 224);  // 3 doubles (written before they are read in every launch)  // [0] live, [1] sticky copy of the last launch (kodb_train_status)
    w += 256;
    p.dense_partials = reinterpret_cast<double*>(w);
    w += tt_align((size_t)TT_MAX_PARTIALS * sizeof(double));
    p.sparse_partials = reinterpret_cast<double*>(w);
    w += tt_align((size_t)p.B * 9 * sizeof(double));
    p.rows_global = spill ? reinterpret_cast<unsigned char*>(w) : nullptr;
    p.rows_global_stride = (long long)spill;

    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return KODB_ERR_UNSUPPORTED;
    if (c_sms[dev] == 0) {
        int sms = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        c_sms[dev] = sms > 0 ? sms : 148;
    }
    p.trace = g_tt_trace;
    cudaStream_t st = (cudaStream_t)stream;
    if (p.use_queue) {
        if (smem > 46 * 1024 && c_smem[0][dev] < smem) {
            if (cudaFuncSetAttribute(train_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
                return KODB_ERR_LAUNCH;
            c_smem[0][dev] = smem;
        }
        // dense CTAs first (lowest block ids are dispatched first and never wait), then one sparse CTA per
        // (image, level).  per_sm dense CTAs per SM (chunk queue: two per SM already reach the fill rate of the HBM)
        // leave room for the sparse ones and for the other kernels of a pipelined step to run concurrently.
        // measured on B200: with many sparse CTAs around (B = 64: 192 of them) a third dense CTA per SM shortens the
        // stream (5.6 vs 5.0 TB/s); with few (B = 8 at 1280^2) two per SM stream better
        int grid_dense = c_sms[dev] * (per_sm ? per_sm : (p.n_sparse >= 96 ? 3 : 2));
        // every sparse CTA should be resident from the start (its row phase overlaps the objectness chunks, and it
        // must not queue up behind dense CTAs): leave it a slot out of the 4 CTAs/SM the launch bounds allow
        const int slots = c_sms[dev] * 4;
        if (grid_dense > slots - p.n_sparse) grid_dense = slots - p.n_sparse;
        if (grid_dense < c_sms[dev]) grid_dense = c_sms[dev];
        if (grid_dense > TD_MAX_GRID) grid_dense = TD_MAX_GRID;
        p.dense_grid = grid_dense;
        p.main_launch = 1;
        static int use_pdl = -1;
        if (use_pdl < 0) {
            const char* e = getenv("KODB_PDL");  // programmatic dependent launch: measured slower on B200 (early-resident
                                                 // CTAs of the next launch get in the way); off unless KODB_PDL=1
            use_pdl = e ? (atoi(e) != 0) : 0;
        }
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3((unsigned)(grid_dense + p.n_sparse));
        lc.blockDim = dim3(TS_THREADS);
        lc.dynamicSmemBytes = smem;
        lc.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        lc.attrs = attr;
        lc.numAttrs = use_pdl ? 1 : 0;
        if (cudaLaunchKernelEx(&lc, train_tail_kernel, p) != cudaSuccess) {
            cudaGetLastError();
            return KODB_ERR_LAUNCH;
        }
        return KODB_OK;
    }
    // fallback: two launches, stream order is the only dependency
    if (smem > 46 * 1024 && c_smem[1][dev] < smem) {
        if (cudaFuncSetAttribute(train_sparse_fallback_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return KODB_ERR_LAUNCH;
        c_smem[1][dev] = smem;
    }
    int grid_dense = c_sms[dev] * 4;
    if (grid_dense > TD_MAX_GRID) grid_dense = TD_MAX_GRID;
    p.dense_grid = grid_dense;
    p.main_launch = 0;
    train_dense_fallback_kernel<<<grid_dense, TD_THREADS, 0, st>>>(p);
    KODB_CHECK_LAUNCH();
    train_sparse_fallback_kernel<<<p.n_sparse, TS_THREADS, smem, st>>>(p);
    KODB_CHECK_LAUNCH();
    return KODB_OK;
}
