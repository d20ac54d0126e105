// kodb_train_tail — Yolov5Loss.forward (kod/lightning/experiments/yv5_baseline/loss.py:166-248) fused
// with its backward and with the target assignment it starts from (yv5.py:207-296).
//
// One persistent kernel.  The nine gradient tensors are cut into tiles = (level, image, anchor, <=256
// contiguous cells of the H*W plane) x all 5+C channels.  A CTA owns a tile completely:
//   1. it re-derives, from the image's ground-truth list, the assigned rows that fall into the tile
//      (order-preserving compaction, key = (offset, target) = the reference's row order for a fixed anchor);
//   2. per row: gathers 4 box + 1 obj logits from their planes, IoU-family forward + analytic backward
//      (iou.py:77-246), objectness target t = clamp(iou, 0) (NOT detached, loss.py:109-118);
//   3. streams the tile: reads the obj logits once (softplus / sigmoid -> BCE sum and dobj) and
//      zero-fills the 4 + C box/cls gradient planes (the dense dY the 1x1 head convs consume);
//   4. after a barrier, patches the assigned cells: summed duplicate gradients, last-writer-wins
//      objectness target (index_put_ semantics), class BCE with pos_weight (loss.py:155-160).
// Loss partials are reduced in fp64 per CTA and finalised by the last CTA in a fixed order
// (deterministic).  No host sync, no data-dependent shapes; N_l is recomputed per CTA from the targets.
#include "kodb_common.cuh"

namespace kodb {

constexpr int TT_THREADS = 256;
constexpr int TT_TILE = 256;  // max cells per tile

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
    int tiles_per_plane[KODB_NUM_LEVELS], tile_cells[KODB_NUM_LEVELS], tile_begin[KODB_NUM_LEVELS + 1];
    int vec_ok[KODB_NUM_LEVELS];
    int row_cap;  // shared-memory row capacity (>= 5 * max targets per image)
    double* partials;       // [gridDim.x][9]
    unsigned int* counter;  // zero on entry, zero on exit
};

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

__device__ __forceinline__ void tile_decode(const TrainParams& p, int t, int& L, int& b, int& a, int& c0, int& ncell) {
    L = t >= p.tile_begin[2] ? 2 : (t >= p.tile_begin[1] ? 1 : 0);
    int r = t - p.tile_begin[L];
    const int tpp = p.tiles_per_plane[L];
    const int chunk = r % tpp;
    r /= tpp;
    const int A = p.lv[L].num_anchors;
    a = r % A;
    b = r / A;
    const int HW = p.lv[L].H * p.lv[L].W;
    c0 = chunk * p.tile_cells[L];
    ncell = min(p.tile_cells[L], HW - c0);
}

__device__ __forceinline__ double block_sum_d(double v, double* sm /* >= 8 */) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) r += sm[i];  // fixed order
    return r;
}

__global__ void __launch_bounds__(TT_THREADS) train_tail_kernel(const __grid_constant__ TrainParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int scan_tmp[33];
    __shared__ double red_tmp[TT_THREADS / 32];
    __shared__ int s_NL[KODB_NUM_LEVELS];
    __shared__ int s_last;

    const int tid = threadIdx.x;
    float up[3];
    if (p.up_dev) {
        up[0] = p.up_dev[0]; up[1] = p.up_dev[1]; up[2] = p.up_dev[2];
        if (p.skip_if_equal && up[0] == p.up_h[0] && up[1] == p.up_h[1] && up[2] == p.up_h[2]) return;
    } else {
        up[0] = p.up_h[0]; up[1] = p.up_h[1]; up[2] = p.up_h[2];
    }

    // shared row records (SoA), capacity row_cap
    int* r_cell = reinterpret_cast<int*>(smem_raw);
    int* r_label = r_cell + p.row_cap;
    float* r_t = reinterpret_cast<float*>(r_label + p.row_cap);
    float* r_x = r_t + p.row_cap;                               // obj logit at the cell
    float4* r_g = reinterpret_cast<float4*>(r_x + p.row_cap);   // combined d/d(box logits)

    // ---- N_l: number of assigned rows per level (needed up-front: every gradient is scaled by 1/N_l) ----
    {
        int cnt[KODB_NUM_LEVELS] = {0, 0, 0};
        for (int n = tid; n < p.T; n += TT_THREADS) {
#pragma unroll
            for (int L = 0; L < KODB_NUM_LEVELS; ++L) {
                const kodb_level_cfg& lv = p.lv[L];
                GtRow g = load_gt_row(p.boxes, p.boxes_dtype, p.labels, n, 1.0 / (double)lv.stride);
                int nb = __popc(neighbour_bits(g.cx, g.cy, lv.fm_w, lv.fm_h));
                for (int k = 0; k < lv.num_anchors; ++k)
                    if (anchor_keeps(g.w, g.h, lv.anchor_grid[k][0], lv.anchor_grid[k][1], p.thr)) cnt[L] += nb;
            }
        }
#pragma unroll
        for (int L = 0; L < KODB_NUM_LEVELS; ++L) {
            int w = warp_sum_i(cnt[L]);
            __syncthreads();
            if ((tid & 31) == 0) scan_tmp[tid >> 5] = w;
            __syncthreads();
            if (tid == 0) {
                int s = 0;
                for (int i = 0; i < TT_THREADS / 32; ++i) s += scan_tmp[i];
                s_NL[L] = s;
            }
        }
        __syncthreads();
        if (blockIdx.x == 0 && tid < KODB_NUM_LEVELS && p.n_assigned) p.n_assigned[tid] = s_NL[tid];
    }

    double acc_loc[KODB_NUM_LEVELS] = {0, 0, 0}, acc_obj[KODB_NUM_LEVELS] = {0, 0, 0}, acc_cls[KODB_NUM_LEVELS] = {0, 0, 0};
    const int C = p.C;
    const int num_tiles = p.tile_begin[KODB_NUM_LEVELS];

    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int L, b, a, c0, ncell;
        tile_decode(p, t, L, b, a, c0, ncell);
        const kodb_level_cfg& lv = p.lv[L];
        const kodb_head_level& hd = p.head[L];
        const kodb_grad_level& gd = p.grad[L];
        const int W = lv.W, H = lv.H;
        const int NL = s_NL[L];
        const double n_obj = (double)p.B * lv.num_anchors * H * W;
        const float s_loc = (float)((double)up[0] * p.lam_loc / (double)NL);
        const float s_obj = (float)((double)up[1] * p.lam_obj * (double)lv.obj_ratio / n_obj);
        const float s_cls = (float)((double)up[2] * p.lam_cls / ((double)NL * (double)C));
        const float aw = lv.anchor_grid[a][0], ah = lv.anchor_grid[a][1];
        const double inv_stride = 1.0 / (double)lv.stride;

        // ---- 1. rows of this tile, in (offset, target) order ----------------------------------------
        const int t0 = p.img_offsets[b], Tb = p.img_offsets[b + 1] - t0;
        int R = 0;
        for (int start = 0; start < 5 * Tb; start += TT_THREADS) {
            const int i = start + tid;
            int flag = 0, cell = 0;
            GtRow g;
            float gxf = 0.f, gyf = 0.f;
            if (i < 5 * Tb) {
                const int o = i / Tb, n = t0 + (i - o * Tb);
                g = load_gt_row(p.boxes, p.boxes_dtype, p.labels, n, inv_stride);
                if (anchor_keeps(g.w, g.h, aw, ah, p.thr) && ((neighbour_bits(g.cx, g.cy, lv.fm_w, lv.fm_h) >> o) & 1u)) {
                    const long long gi = trunc_ll(g.cx - offset_x(o)), gj = trunc_ll(g.cy - offset_y(o));
                    cell = (int)(clamp_ll(gj, 0, H - 1) * W + clamp_ll(gi, 0, W - 1));
                    flag = (cell >= c0 && cell < c0 + ncell);
                    gxf = g.cx - (float)gi;  // gt box uses the UNclamped cell (yv5.py:281)
                    gyf = g.cy - (float)gj;
                }
            }
            int tot;
            const int pos = R + block_excl_scan(flag, scan_tmp, &tot);
            if (flag) {
                // ---- 2. per-row IoU forward/backward ------------------------------------------------
                const int64_t bo = (int64_t)b * hd.box_sb + (int64_t)a * hd.box_sa + cell;
                float pl[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) pl[q] = hd.box[bo + q * hd.box_sp];
                const float xo = hd.obj[(int64_t)b * hd.obj_sb + (int64_t)a * hd.obj_sa + cell];
                float d[4];
                const float val = iou_metric(p.iou_type, p.eps, pl, aw, ah, gxf, gyf, g.w, g.h, d);
                acc_loc[L] += (double)(1.0f - val);
                // dL/d(val): localisation mean(1 - val) and the non-detached objectness target
                const float gv = -s_loc + (val >= 0.0f ? -xo * s_obj : 0.0f);
                r_cell[pos] = cell;
                r_label[pos] = g.label;
                r_t[pos] = fmaxf(val, 0.0f);
                r_x[pos] = xo;
                r_g[pos] = make_float4(gv * d[0], gv * d[1], gv * d[2], gv * d[3]);
            }
            R += tot;
        }

        // ---- 3. dense stream of the tile -------------------------------------------------------------
        if (tid < ncell) {
            const int64_t oo = (int64_t)b * hd.obj_sb + (int64_t)a * hd.obj_sa + c0 + tid;
            float sig;
            const float sp = softplus_sigmoid(hd.obj[oo], &sig);
            acc_obj[L] += (double)sp;
            if (p.has_grads) gd.obj[(int64_t)b * gd.obj_sb + (int64_t)a * gd.obj_sa + c0 + tid] = s_obj * sig;
        }
        if (p.has_grads) {
            const int planes = 4 + C;
            if (p.vec_ok[L]) {
                const int nv = ncell >> 2;
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int idx = tid; idx < planes * nv; idx += TT_THREADS) {
                    const int pl = idx / nv, v = idx - pl * nv;
                    float* base = pl < 4 ? gd.box + (int64_t)b * gd.box_sb + (int64_t)a * gd.box_sa + (int64_t)pl * gd.box_sp
                                         : gd.cls + (int64_t)b * gd.cls_sb + (int64_t)a * gd.cls_sa + (int64_t)(pl - 4) * gd.cls_sp;
                    reinterpret_cast<float4*>(base + c0)[v] = z;
                }
            } else {
                for (int idx = tid; idx < planes * ncell; idx += TT_THREADS) {
                    const int pl = idx / ncell, v = idx - pl * ncell;
                    float* base = pl < 4 ? gd.box + (int64_t)b * gd.box_sb + (int64_t)a * gd.box_sa + (int64_t)pl * gd.box_sp
                                         : gd.cls + (int64_t)b * gd.cls_sb + (int64_t)a * gd.cls_sa + (int64_t)(pl - 4) * gd.cls_sp;
                    base[c0 + v] = 0.0f;
                }
            }
        }

        // ---- 4. patch the assigned cells --------------------------------------------------------------
        if (R > 0) {
            __syncthreads();  // rows visible; tile's dense stores ordered before the patches
            // box + objectness: one thread per row
            for (int r = tid; r < R; r += TT_THREADS) {
                const int cell = r_cell[r];
                bool leader = true, last = true;
                for (int q = 0; q < r; ++q) leader &= (r_cell[q] != cell);
                for (int q = r + 1; q < R; ++q) last &= (r_cell[q] != cell);
                if (last) {
                    // index_put_ last-writer-wins target (loss.py:113-118): BCE(x,t) = softplus(x) - t*x
                    acc_obj[L] -= (double)(r_t[r] * r_x[r]);
                    if (p.has_grads) {
                        const float sig = sigmoid_f(r_x[r]);
                        gd.obj[(int64_t)b * gd.obj_sb + (int64_t)a * gd.obj_sa + cell] = s_obj * (sig - r_t[r]);
                    }
                }
                if (leader && p.has_grads) {
                    float4 s = r_g[r];
                    for (int q = r + 1; q < R; ++q)
                        if (r_cell[q] == cell) { const float4 o4 = r_g[q]; s.x += o4.x; s.y += o4.y; s.z += o4.z; s.w += o4.w; }
                    float* gb = gd.box + (int64_t)b * gd.box_sb + (int64_t)a * gd.box_sa + cell;
                    gb[0] = s.x; gb[gd.box_sp] = s.y; gb[2 * gd.box_sp] = s.z; gb[3 * gd.box_sp] = s.w;
                }
            }
            // classification: one thread per (row, class)
            for (int idx = tid; idx < R * C; idx += TT_THREADS) {
                const int r = idx / C, c = idx - r * C;
                const int cell = r_cell[r];
                const float q = hd.cls[(int64_t)b * hd.cls_sb + (int64_t)a * hd.cls_sa + (int64_t)c * hd.cls_sp + cell];
                float sig;
                const float spn = softplus_sigmoid(-q, &sig);  // -log sigmoid(q); sig = sigmoid(-q) = 1 - sigmoid(q)
                const float pw = p.pos_weight ? p.pos_weight[c] : 1.0f;
                const float tgt = (r_label[r] == c) ? 1.0f : 0.0f;
                const float lw = 1.0f + (pw - 1.0f) * tgt;
                acc_cls[L] += (double)((1.0f - tgt) * q + lw * spn);
                if (p.has_grads) {
                    bool leader = true;
                    for (int k = 0; k < r; ++k) leader &= (r_cell[k] != cell);
                    if (leader) {
                        float s = ((1.0f - tgt) - lw * sig) * s_cls;
                        for (int k = r + 1; k < R; ++k)
                            if (r_cell[k] == cell) {
                                const float t2 = (r_label[k] == c) ? 1.0f : 0.0f;
                                const float lw2 = 1.0f + (pw - 1.0f) * t2;
                                s += ((1.0f - t2) - lw2 * sig) * s_cls;
                            }
                        gd.cls[(int64_t)b * gd.cls_sb + (int64_t)a * gd.cls_sa + (int64_t)c * gd.cls_sp + cell] = s;
                    }
                }
            }
            __syncthreads();  // rows are overwritten by the next tile
        }
    }

    // ---- loss partials -> workspace; the last CTA finalises in a fixed order ---------------------------
    double part[9];
#pragma unroll
    for (int L = 0; L < KODB_NUM_LEVELS; ++L) {
        part[L] = block_sum_d(acc_loc[L], red_tmp);
        part[3 + L] = block_sum_d(acc_obj[L], red_tmp);
        part[6 + L] = block_sum_d(acc_cls[L], red_tmp);
    }
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < 9; ++i) p.partials[(size_t)blockIdx.x * 9 + i] = part[i];
        __threadfence();
        const unsigned prev = atomicAdd(p.counter, 1u);
        s_last = (prev == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double tot[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        double v = 0.0;
        for (int c = tid; c < (int)gridDim.x; c += TT_THREADS) v += __ldcg(&p.partials[(size_t)c * 9 + i]);
        tot[i] = block_sum_d(v, red_tmp);
    }
    if (tid == 0) {
        double loc = 0.0, obj = 0.0, cls = 0.0;
        for (int L = 0; L < KODB_NUM_LEVELS; ++L) {
            const kodb_level_cfg& lv = p.lv[L];
            const double n_obj = (double)p.B * lv.num_anchors * lv.H * lv.W;
            const double nl = (double)s_NL[L];
            loc += tot[L] / nl;  // 0/0 -> NaN: mean over an empty level (loss.py:96)
            obj += (double)lv.obj_ratio * (tot[3 + L] / n_obj);
            cls += tot[6 + L] / (nl * (double)C);
        }
        p.losses[0] = (float)(p.lam_loc * loc);
        p.losses[1] = (float)(p.lam_obj * obj);
        p.losses[2] = (float)(p.lam_cls * cls);
        *p.counter = 0u;
    }
}

static int plan_tiles(TrainParams& p) {
    int begin = 0;
    for (int l = 0; l < KODB_NUM_LEVELS; ++l) {
        const int HW = p.lv[l].H * p.lv[l].W;
        if (HW <= 0) return KODB_ERR_INVALID_ARGUMENT;
        const int tpp = (HW + TT_TILE - 1) / TT_TILE;
        int tc = (HW + tpp - 1) / tpp;
        tc = (tc + 3) & ~3;
        if (tc > TT_TILE) tc = TT_TILE;
        p.tiles_per_plane[l] = (HW + tc - 1) / tc;
        p.tile_cells[l] = tc;
        p.tile_begin[l] = begin;
        begin += p.tiles_per_plane[l] * p.lv[l].num_anchors * p.B;
    }
    p.tile_begin[KODB_NUM_LEVELS] = begin;
    return KODB_OK;
}

static bool aligned16(const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15u) == 0; }

}  // namespace kodb

static const int kTrainMaxGrid = 148 * 8;

extern "C" int64_t kodb_train_workspace_bytes(int32_t, int32_t) {
    return (int64_t)kTrainMaxGrid * 9 * sizeof(double) + 256;
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
    if (t->batch < 1 || t->num_targets < 0 || t->max_per_image < 0) return KODB_ERR_INVALID_ARGUMENT;
    TrainParams p;
    for (int l = 0; l < KODB_NUM_LEVELS; ++l) {
        if (levels_h[l].num_anchors < 1 || levels_h[l].num_anchors > KODB_MAX_ANCHORS) return KODB_ERR_INVALID_ARGUMENT;
        p.head[l] = heads_h[l];
        p.lv[l] = levels_h[l];
        if (grads_h) p.grad[l] = grads_h[l]; else p.grad[l] = kodb_grad_level{};
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
    int rc = plan_tiles(p);
    if (rc != KODB_OK) return rc;
    for (int l = 0; l < KODB_NUM_LEVELS; ++l) {
        const int HW = p.lv[l].H * p.lv[l].W;
        bool ok = (HW % 4 == 0) && p.has_grads;
        if (ok) {
            const kodb_grad_level& g = p.grad[l];
            ok = aligned16(g.box) && aligned16(g.cls) && g.box_sb % 4 == 0 && g.box_sa % 4 == 0 && g.box_sp % 4 == 0 &&
                 g.cls_sb % 4 == 0 && g.cls_sa % 4 == 0 && g.cls_sp % 4 == 0;
        }
        p.vec_ok[l] = ok ? 1 : 0;
    }
    p.row_cap = 5 * (t->max_per_image > 0 ? t->max_per_image : 1);
    const size_t smem = (size_t)p.row_cap * (4 * sizeof(int) + sizeof(float4));
    if (smem > 200 * 1024) return KODB_ERR_CAPACITY;  // > ~5000 boxes in ONE image
    p.partials = reinterpret_cast<double*>(workspace);
    p.counter = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(workspace) + (size_t)kTrainMaxGrid * 9 * sizeof(double));

    // launch geometry is cached per device (no attribute / occupancy queries on the steady-state path,
    // which also keeps the call capturable into a CUDA graph after one warm-up launch)
    static int c_sms[64] = {0}, c_per_sm[64] = {0};
    static size_t c_smem[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return KODB_ERR_UNSUPPORTED;
    if (c_sms[dev] == 0 || c_smem[dev] != smem) {
        int sms = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
        if (smem > 48 * 1024 &&
            cudaFuncSetAttribute(train_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return KODB_ERR_LAUNCH;
        int occ = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, train_tail_kernel, TT_THREADS, smem);
        if (occ < 1) occ = 1;
        if (occ > 8) occ = 8;
        c_sms[dev] = sms;
        c_per_sm[dev] = occ;
        c_smem[dev] = smem;
    }
    const int num_sms = c_sms[dev], per_sm = c_per_sm[dev];
    int grid = num_sms * per_sm;
    if (grid > kTrainMaxGrid) grid = kTrainMaxGrid;
    const int num_tiles = p.tile_begin[KODB_NUM_LEVELS];
    if (grid > num_tiles) grid = num_tiles;
    if (grid < 1) grid = 1;
    train_tail_kernel<<<grid, TT_THREADS, smem, (cudaStream_t)stream>>>(p);
    KODB_CHECK_LAUNCH();
    return KODB_OK;
}
