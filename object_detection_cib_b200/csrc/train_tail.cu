// kodb_train_tail — Yolov5Loss.forward (kod/lightning/experiments/yv5_baseline/loss.py:166-248) fused
// with its backward and with the target assignment it starts from (yv5.py:207-296).
//
// ONE launch with two kinds of CTAs (CTA-level specialisation):
//
//  * dense CTAs (blockIdx < dense_grid) — pure HBM streaming over the nine gradient tensors (the dense dY
//    the 1x1 head convs consume): read every objectness logit ONCE (float4), write
//    dobj = s*sigmoid(x) (the BCE gradient for target 0) and accumulate softplus(x) (the BCE value for
//    target 0); zero-fill the dbox / dcls planes.  Grid-stride float4 loops over contiguous planar
//    tensors, streaming stores, no dependence on the targets.  They carry ~all algorithmic bytes,
//    4*AC*(1+5+C) per image.  When done, each publishes its fp64 partial and bumps a completion counter.
//
//  * sparse CTAs (one per (image, level), scheduled after the dense ones) — re-derive the level's assigned
//    rows of the image from its ground-truth list (order-preserving compaction: for a fixed anchor the
//    reference's row order is (offset, target)), gather 4 box + 1 obj + C cls logits per row, run the
//    IoU-family forward + analytic backward (iou.py:77-246).  All of that overlaps the dense stream.
//    Only then do they wait for the dense completion counter (dense CTAs never wait on anybody and are
//    scheduled first, so the spin cannot deadlock) and patch the assigned cells on top of the dense
//    output: summed duplicate gradients, last-writer-wins objectness target (index_put_ semantics,
//    SURVEY A.3; the target is NOT detached, loss.py:109-118), class BCE with pos_weight
//    (loss.py:155-160).  N_l (needed by every gradient scale) is recomputed per CTA from the flat target
//    list.  The last sparse CTA reduces all fp64 partials in a fixed order -> deterministic losses.
//
// No host sync, no data-dependent shapes, CUDA-graph capturable.
#include <stdlib.h>

#include "kodb_common.cuh"

namespace kodb {

constexpr int TD_THREADS = 256;
constexpr int TS_THREADS = 256;
constexpr int TD_MAX_GRID = 148 * 8;

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
    int row_cap;                      // shared-memory row capacity of the sparse CTAs
    int cls_in_smem;                  // class-BCE gradients of the rows fit in shared memory
    int dense_grid;
    double* dense_partials;   // [TD_MAX_GRID][3]   softplus sums per level
    double* sparse_partials;  // [B*3][3]           (loc, obj correction, cls) per (image, level)
    unsigned int* counter;    // sparse CTAs that published their loss partials; zero on entry, zero on exit
    unsigned int* counter2;   // sparse CTAs that are completely done; zero on entry, zero on exit
    int* nl_pub;              // [3] N_l published by the b == 0 CTAs for the finaliser
    unsigned int* dense_done; // completion counter of the dense CTAs; zero on entry / exit
    // chunk queue of the dense stream (all levels contiguous): chunk ids are handed out by an atomic counter, so a
    // CTA that is served slower by the memory system simply takes fewer chunks
    int use_queue;
    unsigned int* queue;                       // zero on entry / exit
    int ch_obj, ch_zero;                       // float4 per objectness chunk / per zero-fill chunk
    int n_obj_chunks[KODB_NUM_LEVELS];         // chunks of the level's objectness plane
    int n_box_chunks[KODB_NUM_LEVELS], n_cls_chunks[KODB_NUM_LEVELS];
    int n_chunks;                              // all chunks
    int n_partials;                            // fp64 softplus partials the finaliser sums (queue: one per obj chunk)
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

// =========================================================================================================
// 1. dense stream
// =========================================================================================================
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

// Dense stream over a chunk queue.  Chunk order: per level [objectness chunks | box zero-fill | cls zero-fill].  An
// objectness chunk reads its logits (all loads of the thread issued first), writes dobj = s*sigmoid(x) and stores
// the chunk's fp64 softplus sum at a fixed place (dense_partials[chunk]) -- the finaliser adds the partials in chunk
// order, so the loss does not depend on which CTA took which chunk.  The next ticket is requested before the
// current chunk is processed (the atomic's round trip is hidden).
__device__ __forceinline__ void dense_queue(const TrainParams& p, const float up[3], double* red_tmp, int* ticket_sm) {
    constexpr int OBJ_PER_THREAD = 4, ZERO_PER_THREAD = 8;  // float4 per thread and chunk (chunk sizes are multiples)
    const int tid = threadIdx.x;
    // the first chunk of a CTA is its block index (no atomic); tickets from the counter start behind those
    unsigned next = blockIdx.x;
    for (;;) {
        __syncthreads();  // the previous round's readers of ticket_sm are done
        if (tid == 0) {
            ticket_sm[0] = (int)next;
            if ((int)next < p.n_chunks) next = (unsigned)p.dense_grid + atomicAdd(p.queue, 1u);  // prefetch the next ticket
        }
        __syncthreads();
        int c = ticket_sm[0];
        if (c >= p.n_chunks) break;
        int L = 0, kind = 0;  // kind 0: obj, 1: box, 2: cls
        for (;;) {
            if (c < p.n_obj_chunks[L]) { kind = 0; break; }
            c -= p.n_obj_chunks[L];
            if (c < p.n_box_chunks[L]) { kind = 1; break; }
            c -= p.n_box_chunks[L];
            if (c < p.n_cls_chunks[L]) { kind = 2; break; }
            c -= p.n_cls_chunks[L];
            ++L;
        }
        const kodb_level_cfg& lv = p.lv[L];
        const int64_t n_obj = (int64_t)p.B * lv.num_anchors * lv.H * lv.W;
        if (kind == 0) {
            const int64_t n4 = n_obj >> 2, i0 = (int64_t)c * p.ch_obj;
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
            int pidx = c;  // index of this objectness chunk among all objectness chunks
            for (int l = 0; l < L; ++l) pidx += p.n_obj_chunks[l];
            const double sum = block_sum_d((double)a0, red_tmp);
            if (tid == 0) p.dense_partials[pidx] = sum;
        } else {
            const int64_t n4 = kind == 1 ? n_obj : (n_obj >> 2) * p.C;  // float4 of the box / cls planes
            float4* dst = reinterpret_cast<float4*>(kind == 1 ? p.grad[L].box : p.grad[L].cls);
            const int64_t i0 = (int64_t)c * p.ch_zero;
            const int64_t i1 = i0 + p.ch_zero < n4 ? i0 + p.ch_zero : n4;
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int64_t base = i0 + tid; base < i1; base += ZERO_PER_THREAD * TD_THREADS) {
#pragma unroll
                for (int u = 0; u < ZERO_PER_THREAD; ++u)
                    if (base + u * TD_THREADS < i1) __stcs(dst + base + u * TD_THREADS, z);
            }
        }
    }
}

__device__ __forceinline__ void dense_cta(const TrainParams& p, const float up[3], double* red_tmp, int* cnt_tmp) {
    TT_TRACE(0);
    // N_l = number of assigned rows per level over the whole batch: the first B dense CTAs count one image
    // each (a handful of targets) and accumulate with integer atomics; the sparse CTAs read the totals after
    // the dense-completion wait, so no extra barrier is needed.
    for (int img = blockIdx.x; img < p.B; img += p.dense_grid) {
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
    }
    if (p.use_queue) {
        dense_queue(p, up, red_tmp, cnt_tmp);
        __syncthreads();  // every thread's stores are issued
        if (threadIdx.x == 0) {
            __threadfence();  // release: stores of this CTA visible device-wide before the counter moves
            atomicAdd(p.dense_done, 1u);
        }
        TT_TRACE(1);
        return;
    }
    const int64_t gtid = (int64_t)blockIdx.x * TD_THREADS + threadIdx.x;
    const int64_t gsize = (int64_t)p.dense_grid * TD_THREADS;
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
    __syncthreads();  // every thread's stores are issued
    if (threadIdx.x == 0) {
        __threadfence();  // release: stores of this CTA visible device-wide before the counter moves
        atomicAdd(p.dense_done, 1u);
    }
    TT_TRACE(1);
}

// =========================================================================================================
// 2. sparse rows
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


// shared-memory row record of the sparse CTAs (structure of arrays, capacity p.row_cap)
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

__device__ __forceinline__ void sparse_cta(const TrainParams& p, const float up[3], unsigned char* smem_raw,
                                           int* scan_tmp, double* red_tmp, int& s_last) {
    const int tid = threadIdx.x;
    const int sid = blockIdx.x - p.dense_grid;
    const int b = sid / KODB_NUM_LEVELS, L = sid - b * KODB_NUM_LEVELS;
    TT_TRACE(0);

    const kodb_level_cfg& lv = p.lv[L];
    const kodb_head_level& hd = p.head[L];
    const kodb_grad_level& gd = p.grad[L];
    const int W = lv.W, H = lv.H, HW = H * W, A = lv.num_anchors, C = p.C;
    const double inv_stride = 1.0 / (double)lv.stride;
    const SparseRows R_ = carve_rows(smem_raw, p.row_cap);

    double acc_loc = 0.0, acc_obj = 0.0, acc_cls = 0.0;

    // ---- rows of (image b, level L): index i = (k*5 + o)*Tb + n, i.e. (offset, target) order per anchor ----
    const int t0 = p.img_offsets[b], Tb = p.img_offsets[b + 1] - t0;
    const int total = 5 * A * Tb;
    int R = 0;
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
        if (flag) {
            const float aw = lv.anchor_grid[k][0], ah = lv.anchor_grid[k][1];
            const int64_t bo = (int64_t)b * hd.box_sb + (int64_t)k * hd.box_sa + cell;
            float pl[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) pl[q] = hd.box[bo + q * hd.box_sp];
            const float xo = hd.obj[(int64_t)b * hd.obj_sb + (int64_t)k * hd.obj_sa + cell];
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
    // duplicate structure: leader / last flags and the chain to the next row of the same cell
    for (int r = tid; r < R; r += TS_THREADS) {
        const int id = R_.id[r];
        int fl = 3, nx = -1;
        for (int q = 0; q < r; ++q)
            if (R_.id[q] == id) { fl &= ~1; break; }
        for (int q = r + 1; q < R; ++q)
            if (R_.id[q] == id) { fl &= ~2; nx = q; break; }
        R_.flags[r] = fl;
        R_.next[r] = nx;
        if (fl & 2) acc_obj -= (double)(fmaxf(R_.val[r], 0.0f) * R_.xo[r]);  // BCE(x,t) = softplus(x) - t*x, last writer
    }
    // classification: one thread per (row, class); value now, gradient kept unscaled in shared memory
    for (int idx = tid; idx < R * C; idx += TS_THREADS) {
        const int r = idx / C, c = idx - r * C;
        const int id = R_.id[r];
        const int k = id / HW, cell = id - k * HW;
        const float q = hd.cls[(int64_t)b * hd.cls_sb + (int64_t)k * hd.cls_sa + (int64_t)c * hd.cls_sp + cell];
        float sig;
        const float spn = softplus_sigmoid(-q, &sig);  // -log sigmoid(q); sig = 1 - sigmoid(q)
        const float pw = p.pos_weight ? p.pos_weight[c] : 1.0f;
        const float tgt = (R_.label[r] == c) ? 1.0f : 0.0f;
        const float lw = 1.0f + (pw - 1.0f) * tgt;
        acc_cls += (double)((1.0f - tgt) * q + lw * spn);
        if (p.cls_in_smem) R_.gcls[idx] = (1.0f - tgt) - lw * sig;
    }
    TT_TRACE(2);

    // ---- loss partials -> workspace now (they do not depend on the dense stream); the CTA that publishes last
    // becomes the finaliser and reduces everything in a fixed order while the others are still patching ---------
    const double pl_ = block_sum_d(acc_loc, red_tmp);
    const double po_ = block_sum_d(acc_obj, red_tmp);
    const double pc_ = block_sum_d(acc_cls, red_tmp);
    const unsigned nblocks = gridDim.x - p.dense_grid;
    if (tid == 0) {
        double* dst = p.sparse_partials + ((size_t)b * KODB_NUM_LEVELS + L) * 3;
        dst[0] = pl_; dst[1] = po_; dst[2] = pc_;
        __threadfence();
        const unsigned prev = atomicAdd(p.counter, 1u);
        s_last = (prev == nblocks - 1);
    }
    __syncthreads();
    // ---- wait for the dense stream: everything above overlapped it; the patches below must land on top ----
    if (tid == 0) {
        // plain (L2) loads: polling with an atomic would queue up behind the dense CTAs' increments of the same word
        while (*(const volatile unsigned int*)p.dense_done < (unsigned)p.dense_grid) __nanosleep(200);
        __threadfence();  // acquire (also orders the N_l counters the dense CTAs accumulated)
    }
    __syncthreads();
    TT_TRACE(3);
    const int NL = (int)__ldcg(&p.nl_pub[L]);
    const double n_obj = (double)p.B * A * H * W;
    const float s_loc = (float)((double)up[0] * p.lam_loc / (double)NL);
    const float s_obj = (float)((double)up[1] * p.lam_obj * (double)lv.obj_ratio / n_obj);
    const float s_cls = (float)((double)up[2] * p.lam_cls / ((double)NL * (double)C));

    // ---- patch the assigned cells (the dense CTAs have written s*sigmoid / zeros) ------------------------
    if (p.has_grads) {
        for (int r = tid; r < R; r += TS_THREADS) {
            const int id = R_.id[r], fl = R_.flags[r];
            const int k = id / HW, cell = id - k * HW;
            if (fl & 2)  // index_put_ last-writer-wins target (loss.py:113-118)
                gd.obj[(int64_t)b * gd.obj_sb + (int64_t)k * gd.obj_sa + cell] =
                    s_obj * (sigmoid_f(R_.xo[r]) - fmaxf(R_.val[r], 0.0f));
            if (fl & 1) {
                float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int q = r; q >= 0; q = R_.next[q]) {  // every duplicate receives (and contributes) gradient
                    const float v = R_.val[q];
                    // dL/d(val): localisation mean(1 - val) and the non-detached objectness target
                    const float gv = -s_loc + (v >= 0.0f ? -R_.xo[q] * s_obj : 0.0f);
                    const float4 d4 = R_.d[q];
                    s.x += gv * d4.x; s.y += gv * d4.y; s.z += gv * d4.z; s.w += gv * d4.w;
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
            if (p.cls_in_smem) {
                for (int q = r; q >= 0; q = R_.next[q]) s += R_.gcls[q * C + c] * s_cls;
            } else {  // row list too large for shared-memory gradients: recompute from the logit
                const float qv = hd.cls[(int64_t)b * hd.cls_sb + (int64_t)k * hd.cls_sa + (int64_t)c * hd.cls_sp + cell];
                float sig;
                softplus_sigmoid(-qv, &sig);
                const float pw = p.pos_weight ? p.pos_weight[c] : 1.0f;
                for (int q = r; q >= 0; q = R_.next[q]) {
                    const float tgt = (R_.label[q] == c) ? 1.0f : 0.0f;
                    const float lw = 1.0f + (pw - 1.0f) * tgt;
                    s += ((1.0f - tgt) - lw * sig) * s_cls;
                }
            }
            gd.cls[(int64_t)b * gd.cls_sb + (int64_t)k * gd.cls_sa + (int64_t)c * gd.cls_sp + cell] = s;
        }
    }
    TT_TRACE(4);

    TT_TRACE(5);
    if (s_last) {
    // (the partials were released with a fence before the counter moved and are read through L2 below)
    // all loads first (independent, L2), then ONE block reduction of the nine sums in a fixed order
    double v[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) v[i] = 0.0;
    for (int i = tid; i < p.B * KODB_NUM_LEVELS; i += TS_THREADS) {
        const int l = i % KODB_NUM_LEVELS;
        const double* src = p.sparse_partials + (size_t)i * 3;
        const double a0 = __ldcg(src), a1 = __ldcg(src + 1), a2 = __ldcg(src + 2);
#pragma unroll
        for (int q = 0; q < KODB_NUM_LEVELS; ++q)
            if (q == l) { v[q] += a0; v[3 + q] += a1; v[6 + q] += a2; }
    }
    if (p.use_queue) {
        const int e0 = p.n_obj_chunks[0], e1 = e0 + p.n_obj_chunks[1];
        for (int i = tid; i < p.n_partials; i += TS_THREADS) {
            const double a = __ldcg(p.dense_partials + i);
            if (i < e0) v[3] += a; else if (i < e1) v[4] += a; else v[5] += a;
        }
    } else {
        for (int i = tid; i < p.dense_grid; i += TS_THREADS) {
            const double* src = p.dense_partials + (size_t)i * 3;
            v[3] += __ldcg(src); v[4] += __ldcg(src + 1); v[5] += __ldcg(src + 2);
        }
    }
    double* red9 = reinterpret_cast<double*>(smem_raw);  // rows are dead; >= 8 warps * 9 doubles available
#pragma unroll
    for (int i = 0; i < 9; ++i) v[i] = warp_sum(v[i]);
    __syncthreads();
    if ((tid & 31) == 0) {
#pragma unroll
        for (int i = 0; i < 9; ++i) red9[(tid >> 5) * 9 + i] = v[i];
    }
    __syncthreads();
    if (tid == 0) {
        double tot[9];
        for (int i = 0; i < 9; ++i) {
            double s = 0.0;
            for (int w = 0; w < TS_THREADS / 32; ++w) s += red9[w * 9 + i];
            tot[i] = s;
        }
        double loc = 0.0, obj = 0.0, cls = 0.0;
        for (int l = 0; l < KODB_NUM_LEVELS; ++l) {
            const kodb_level_cfg& q = p.lv[l];
            const int nl = __ldcg(&p.nl_pub[l]);
            const double nobj = (double)p.B * q.num_anchors * q.H * q.W;
            loc += tot[l] / (double)nl;  // 0/0 -> NaN: mean over an empty level (loss.py:96)
            obj += (double)q.obj_ratio * (tot[3 + l] / nobj);
            cls += tot[6 + l] / ((double)nl * (double)C);
            if (p.n_assigned) p.n_assigned[l] = nl;
        }
        p.losses[0] = (float)(p.lam_loc * loc);
        p.losses[1] = (float)(p.lam_obj * obj);
        p.losses[2] = (float)(p.lam_cls * cls);
    }
    }
    TT_TRACE(6);
    // ---- the CTA that finishes last (patches written, losses out) resets the control words for the next launch
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned prev2 = atomicAdd(p.counter2, 1u);
        if (prev2 == nblocks - 1) {
            *p.counter = 0u;
            *p.counter2 = 0u;
            *p.dense_done = 0u;
            *p.queue = 0u;
            p.nl_pub[0] = 0; p.nl_pub[1] = 0; p.nl_pub[2] = 0;
        }
    }
}

__global__ void __launch_bounds__(TS_THREADS, 4) train_tail_kernel(const __grid_constant__ TrainParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int scan_tmp[33];
    __shared__ double red_tmp[32];
    __shared__ int s_last;
    float up[3];
    if (!load_upstream(p, up)) return;
    if ((int)blockIdx.x < p.dense_grid) dense_cta(p, up, red_tmp, scan_tmp);
    else sparse_cta(p, up, smem_raw, scan_tmp, red_tmp, s_last);
}

static long long* g_tt_trace = nullptr;
static bool aligned16(const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15u) == 0; }

}  // namespace kodb

// debug hook (not part of include/kodb.h): device buffer of (grid)*8 int64 phase stamps per CTA
extern "C" void kodb_debug_set_train_trace(void* ptr) { kodb::g_tt_trace = reinterpret_cast<long long*>(ptr); }

static size_t tt_align(size_t v) { return (v + 255) & ~(size_t)255; }

extern "C" int64_t kodb_train_workspace_bytes(int32_t batch, int32_t) {
    return (int64_t)(tt_align((size_t)kodb::TD_MAX_GRID * 3 * sizeof(double)) +
                     tt_align((size_t)(batch > 0 ? batch : 1) * 9 * sizeof(double)) + 256);
}

extern "C" int kodb_train_tail(const kodb_head_level* heads_h, const kodb_level_cfg* levels_h,
                               const kodb_targets* t, const kodb_loss_cfg* cfg, const float* pos_weight,
                               const float* upstream_h, const float* upstream_dev, int32_t skip_if_equal,
                               const kodb_grad_level* grads_h, float* losses, int32_t* n_assigned,
                               void* workspace, int64_t workspace_bytes, void* stream) {
    using namespace kodb;
    if (!heads_h || !levels_h || !t || !cfg || !upstream_h || !losses || !workspace) return KODB_ERR_INVALID_ARGUMENT;
    if (workspace_bytes < kodb_train_workspace_bytes(t->batch, t->max_per_image)) return KODB_ERR_CAPACITY;
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
    p.n_chunks = p.n_partials = 0;
    p.ch_obj = 4 * TD_THREADS;
    p.ch_zero = 16 * TD_THREADS;  // 64 KB: few enough tickets for one atomic counter, fine enough to balance
    if (p.use_queue) {
        int64_t tot4 = 0;
        for (int l = 0; l < KODB_NUM_LEVELS; ++l) tot4 += ((int64_t)p.B * p.lv[l].num_anchors * p.lv[l].H * p.lv[l].W) >> 2;
        // objectness chunks are the fp64 partials of the finaliser: keep their number inside the workspace
        const int64_t max_partials = (int64_t)TD_MAX_GRID * 3 - KODB_NUM_LEVELS;
        while ((tot4 + p.ch_obj - 1) / p.ch_obj > max_partials) p.ch_obj *= 2;
        for (int l = 0; l < KODB_NUM_LEVELS; ++l) {
            const int64_t n_obj = (int64_t)p.B * p.lv[l].num_anchors * p.lv[l].H * p.lv[l].W;
            const int64_t n4 = n_obj >> 2;
            p.n_obj_chunks[l] = (int)((n4 + p.ch_obj - 1) / p.ch_obj);
            p.n_box_chunks[l] = p.has_grads ? (int)((n_obj + p.ch_zero - 1) / p.ch_zero) : 0;
            p.n_cls_chunks[l] = p.has_grads ? (int)((n4 * p.C + p.ch_zero - 1) / p.ch_zero) : 0;
            p.n_partials += p.n_obj_chunks[l];
            p.n_chunks += p.n_obj_chunks[l] + p.n_box_chunks[l] + p.n_cls_chunks[l];
        }
    }
    p.row_cap = 5 * amax * (t->max_per_image > 0 ? t->max_per_image : 1);
    size_t smem = (size_t)p.row_cap * (6 * sizeof(int) + sizeof(float4));
    if (smem > 160 * 1024) return KODB_ERR_CAPACITY;  // > ~270 boxes in ONE image (reference data has <= ~72)
    const size_t cls_bytes = (size_t)p.row_cap * p.C * sizeof(float);
    p.cls_in_smem = (smem + cls_bytes <= 96 * 1024) ? 1 : 0;
    if (p.cls_in_smem) smem += cls_bytes;
    if (smem < 8 * 9 * sizeof(double)) smem = 8 * 9 * sizeof(double);  // the finaliser's reduction scratch
    // fixed offsets first (counter / N_l), so a cached workspace stays valid when the batch size changes
    char* w = reinterpret_cast<char*>(workspace);
    p.counter = reinterpret_cast<unsigned int*>(w);
    p.nl_pub = reinterpret_cast<int*>(w + 64);
    p.dense_done = reinterpret_cast<unsigned int*>(w + 128);
    p.queue = reinterpret_cast<unsigned int*>(w + 192);
    p.counter2 = reinterpret_cast<unsigned int*>(w + 200);
    w += 256;
    p.dense_partials = reinterpret_cast<double*>(w);
    w += tt_align((size_t)TD_MAX_GRID * 3 * sizeof(double));
    p.sparse_partials = reinterpret_cast<double*>(w);

    // launch geometry cached per device (keeps the call capturable into a CUDA graph after one warm-up)
    static int c_sms[64] = {0};
    static size_t c_smem[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return KODB_ERR_UNSUPPORTED;
    if (c_sms[dev] == 0) {
        int sms = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        c_sms[dev] = sms > 0 ? sms : 148;
    }
    if (smem > 48 * 1024 && c_smem[dev] < smem) {
        if (cudaFuncSetAttribute(train_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return KODB_ERR_LAUNCH;
        c_smem[dev] = smem;
    }
    // dense CTAs first (lowest block ids are scheduled first and never wait), then one sparse CTA per
    // (image, level).  Two dense CTAs per SM (chunk queue: 296 CTAs already reach the fill rate of the HBM) leave room
    // for the sparse ones and for the other kernels of a pipelined step to run concurrently.
    static int per_sm = 0;
    if (per_sm == 0) {
        const char* e = getenv("KODB_TRAIN_DENSE_PER_SM");  // tuning knob; default chosen on B200
        per_sm = e ? atoi(e) : 2;
        if (per_sm < 1 || per_sm > 8) per_sm = 2;
    }
    int grid_dense = c_sms[dev] * per_sm;
    if (p.use_queue) {
        // every sparse CTA should be resident from the start (its row phase overlaps the dense stream, and it must
        // not queue up behind dense CTAs): leave it a slot out of the 4 CTAs/SM the kernel's launch bounds allow
        const int slots = c_sms[dev] * 4, n_sparse = p.B * KODB_NUM_LEVELS;
        if (grid_dense > slots - n_sparse) grid_dense = slots - n_sparse;
        if (grid_dense < c_sms[dev]) grid_dense = c_sms[dev];
    }
    if (grid_dense > TD_MAX_GRID) grid_dense = TD_MAX_GRID;
    p.dense_grid = grid_dense;
    p.trace = g_tt_trace;
    cudaStream_t st = (cudaStream_t)stream;
    train_tail_kernel<<<grid_dense + p.B * KODB_NUM_LEVELS, TS_THREADS, smem, st>>>(p);
    KODB_CHECK_LAUNCH();
    return KODB_OK;
}
