// Eval tail: confidence filter -> multi-label candidates -> class-offset greedy NMS -> per-class IoU matching.
//
//   kodb_eval_tail  = validation_step's get_detections + non_max_suppression (exp.py:140-154) fused: the
//                     (B, AC, 5+C) tensor of kod/lightning/experiments/yv5_baseline/layers.py:127-155 is never
//                     written; candidates are produced straight from the planar head logits.
//   kodb_nms        = non_max_suppression (kod/core/nms.py:9-75) on an already decoded (B, N, 5+C) tensor.
//   kodb_map_match  = the per-(image, class) greedy matching of pycocotools' COCOeval.evaluateImg that
//                     PyCOCOMAPEvalCallback drives (pycoco_map_eval.py:50-106), on the kept detections.
//
// Two kernels per call:
//  1. cand_kernel   (grid = tiles x images): one thread per anchor-cell reads obj and the C class scores
//     once (planes are cell-contiguous -> coalesced), keeps (row, class) pairs with obj > thr and
//     obj*cls > thr, and appends them as (score, key = row*C + class) to the image's candidate arrays.
//     A tile's candidates are contiguous and in key order (block scan), tiles land in arbitrary order
//     (one atomicAdd per tile); per-tile (start, count) let the reference's row-major candidate index be
//     reconstructed for the <= 300 survivors only.  Boxes are NOT decoded here.
//  2. select_nms_kernel (one 1024-thread CTA per image): exact top-K extraction in (score desc, key asc)
//     order by histogram radix-select over a 64-bit composite key (robust to ties / skewed scores: bins
//     are refined until a batch fits shared memory), bitonic sort of the batch, box decode for the batch
//     only, greedy NMS in sub-blocks of 32 with warp-ballot suppression masks against the kept list,
//     early exit at max_det keeps (exact: the first 300 keeps depend only on the scanned prefix).
//     fp32 IoU arithmetic is torchvision's CPU kernel op for op (offset boxes, ratio vs DOUBLE threshold).
//     The kept detections then go through the fp64 COCO matching while still in shared memory.
#include "kodb_common.cuh"

namespace kodb {

constexpr int EV_TILE = 256;
constexpr int NB = 4096;         // histogram bins per refinement level
constexpr int K_CAP = 2048;      // sort capacity (batch size upper bound)
constexpr int K_TARGET = 1024;   // preferred batch size
constexpr int NMS_THREADS = 1024;
constexpr int MAX_DET_CAP = 1024;
constexpr int GT_CAP = 2048;     // max ground-truth boxes of ONE image in the matching stage
constexpr unsigned long long U64MAX = 0xffffffffffffffffull;

struct EvalGeom {
    kodb_head_level head[KODB_NUM_LEVELS];
    kodb_level_cfg lv[KODB_NUM_LEVELS];
    int row_base[KODB_NUM_LEVELS + 1];
    int tiles_per_plane[KODB_NUM_LEVELS], tile_cells[KODB_NUM_LEVELS], tile_begin[KODB_NUM_LEVELS + 1];
};

struct EvalParams {
    EvalGeom g;
    const float* det;  // interleaved input (planar == 0)
    int64_t rows;      // rows per image
    int planar, B, C;
    kodb_nms_cfg cfg;
    // workspace
    float* cand_score;
    uint32_t* cand_key;
    int64_t cand_stride;
    int* img_count;   // [B]   zero on entry / exit
    int* tile_start;  // [B][tiles_per_image]
    int* tile_count;  // [B][tiles_per_image]
    int tiles_per_image;
    kodb_nms_out out;
    // matching
    int do_match;
    const void* gt_boxes;
    const int64_t* gt_labels;
    const int32_t* gt_offsets;
    int gt_dtype;
    kodb_match_out m;
};

// ---- score <-> sortable key ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t desc_key(float s) {
    const uint32_t b = __float_as_uint(s);
    const uint32_t u = b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
    return ~u;  // ascending desc_key == descending score
}
__device__ __forceinline__ float score_of_key(uint32_t k) {
    const uint32_t u = ~k;
    const uint32_t b = (u & 0x80000000u) ? (u ^ 0x80000000u) : ~u;
    return __uint_as_float(b);
}
__device__ __forceinline__ unsigned long long make_ck(float s, uint32_t key) {
    return ((unsigned long long)desc_key(s) << 32) | key;
}

// ---- geometry helpers -----------------------------------------------------------------------------------
__device__ __forceinline__ void planar_tile(const EvalGeom& g, int t, int& L, int& a, int& c0, int& ncell) {
    L = t >= g.tile_begin[2] ? 2 : (t >= g.tile_begin[1] ? 1 : 0);
    const int r = t - g.tile_begin[L];
    const int tpp = g.tiles_per_plane[L];
    a = r / tpp;
    c0 = (r - a * tpp) * g.tile_cells[L];
    ncell = min(g.tile_cells[L], g.lv[L].H * g.lv[L].W - c0);
}
__device__ __forceinline__ void planar_row(const EvalGeom& g, int row, int& L, int& a, int& cell) {
    L = row >= g.row_base[2] ? 2 : (row >= g.row_base[1] ? 1 : 0);
    const int r = row - g.row_base[L];
    const int HW = g.lv[L].H * g.lv[L].W;
    a = r / HW;
    cell = r - a * HW;
}
__device__ __forceinline__ int tile_of_row(const EvalParams& p, int row) {
    if (!p.planar) return row / EV_TILE;
    int L, a, cell;
    planar_row(p.g, row, L, a, cell);
    return p.g.tile_begin[L] + a * p.g.tiles_per_plane[L] + cell / p.g.tile_cells[L];
}
__device__ __forceinline__ float4 fetch_box(const EvalParams& p, int b, int row) {
    if (!p.planar) {
        const float* r = p.det + ((int64_t)b * p.rows + row) * (5 + p.C);
        return make_float4(r[0], r[1], r[2], r[3]);
    }
    int L, a, cell;
    planar_row(p.g, row, L, a, cell);
    const kodb_level_cfg& lv = p.g.lv[L];
    const kodb_head_level& hd = p.g.head[L];
    const int gy = cell / lv.W, gx = cell - gy * lv.W;
    const int64_t bo = (int64_t)b * hd.box_sb + (int64_t)a * hd.box_sa + cell;
    return decode_box(hd.box[bo], hd.box[bo + hd.box_sp], hd.box[bo + 2 * hd.box_sp], hd.box[bo + 3 * hd.box_sp], (float)gx,
                      (float)gy, (float)lv.stride, lv.anchor_px[a][0], lv.anchor_px[a][1]);
}

// =========================================================================================================
// 1. candidates
// =========================================================================================================
template <bool PLANAR>
__global__ void __launch_bounds__(EV_TILE) cand_kernel(const __grid_constant__ EvalParams p) {
    __shared__ int scan_tmp[33];
    __shared__ int s_base;
    const int b = blockIdx.y, t = blockIdx.x, tid = threadIdx.x;
    const int C = p.C;
    const float thr = p.cfg.conf_thres;
    int row0, ncell;
    const float* obj_ptr;
    const float* cls_ptr;
    int64_t cls_step;
    if (PLANAR) {
        int L, a, c0;
        planar_tile(p.g, t, L, a, c0, ncell);
        const kodb_head_level& hd = p.g.head[L];
        row0 = p.g.row_base[L] + a * (p.g.lv[L].H * p.g.lv[L].W) + c0;
        obj_ptr = hd.obj + (int64_t)b * hd.obj_sb + (int64_t)a * hd.obj_sa + c0 + tid;
        cls_ptr = hd.cls + (int64_t)b * hd.cls_sb + (int64_t)a * hd.cls_sa + c0 + tid;
        cls_step = hd.cls_sp;
    } else {
        row0 = t * EV_TILE;
        ncell = (int)min((int64_t)EV_TILE, p.rows - row0);
        const float* r = p.det + ((int64_t)b * p.rows + row0 + tid) * (5 + C);
        obj_ptr = r + 4;
        cls_ptr = r + 5;
        cls_step = 1;
    }
    float o = 0.0f;
    bool active = false;
    if (tid < ncell) {
        o = PLANAR ? sigmoid_f(*obj_ptr) : *obj_ptr;
        active = o > thr;  // nms.py:33
    }
    unsigned bits[KODB_MAX_CLASSES / 32] = {0u, 0u, 0u, 0u};
    int cnt = 0;
    if (__any_sync(KODB_FULL_MASK, active)) {
#pragma unroll
        for (int w = 0; w < KODB_MAX_CLASSES / 32; ++w) {
            const unsigned allowed = p.cfg.class_mask[w];
            for (int jj = 0; jj < 32; ++jj) {
                const int j = w * 32 + jj;
                if (j >= C) break;
                if (!((allowed >> jj) & 1u)) continue;  // `classes` filter (nms.py:54-55)
                if (active) {
                    const float c = PLANAR ? sigmoid_f(cls_ptr[(int64_t)j * cls_step]) : cls_ptr[(int64_t)j * cls_step];
                    const float conf = c * o;  // nms.py:40
                    if (conf > thr) { bits[w] |= 1u << jj; ++cnt; }
                }
            }
        }
    }
    int tot;
    const int pos0 = block_excl_scan(cnt, scan_tmp, &tot);
    if (tid == 0) {
        const int base = tot ? atomicAdd(&p.img_count[b], tot) : 0;
        s_base = base;
        p.tile_start[(int64_t)b * p.tiles_per_image + t] = base;
        p.tile_count[(int64_t)b * p.tiles_per_image + t] = tot;
    }
    __syncthreads();
    if (cnt) {
        float* sc = p.cand_score + (int64_t)b * p.cand_stride + s_base + pos0;
        uint32_t* ky = p.cand_key + (int64_t)b * p.cand_stride + s_base + pos0;
        const uint32_t kbase = (uint32_t)(row0 + tid) * (uint32_t)C;
        int q = 0;
#pragma unroll
        for (int w = 0; w < KODB_MAX_CLASSES / 32; ++w) {
            unsigned m = bits[w];
            while (m) {
                const int jj = __ffs(m) - 1;
                m &= m - 1;
                const int j = w * 32 + jj;
                const float c = PLANAR ? sigmoid_f(cls_ptr[(int64_t)j * cls_step]) : cls_ptr[(int64_t)j * cls_step];
                sc[q] = c * o;
                ky[q] = kbase + j;
                ++q;
            }
        }
    }
}

// =========================================================================================================
// 2. per-image select + NMS (+ matching)
// =========================================================================================================
struct NmsSmem {
    int hist[NB];
    unsigned long long sort_ck[K_CAP];
    uint32_t sort_pos[K_CAP];
    float bx1[K_CAP], by1[K_CAP], bx2[K_CAP], by2[K_CAP], barea[K_CAP];
    float kx1[MAX_DET_CAP], ky1[MAX_DET_CAP], kx2[MAX_DET_CAP], ky2[MAX_DET_CAP], karea[MAX_DET_CAP];
    unsigned long long kept_ck[MAX_DET_CAP];
    long long kept_idx[MAX_DET_CAP];
    int scan_tmp[33];
    int alive[32];
    unsigned mask[32];
    int sel_bin, sel_p, sel_prev;
    int cnt, nk;
};

__device__ __forceinline__ bool suppresses(float ax1, float ay1, float ax2, float ay2, float aarea, float bx1, float by1,
                                           float bx2, float by2, float barea, double thr) {
    // torchvision nms_kernel.cpp: fp32 throughout, ratio compared with the double threshold
    const float xx1 = fmaxf(ax1, bx1), yy1 = fmaxf(ay1, by1), xx2 = fminf(ax2, bx2), yy2 = fminf(ay2, by2);
    const float w = fmaxf(0.0f, xx2 - xx1), h = fmaxf(0.0f, yy2 - yy1);
    const float inter = w * h;
    const float ovr = inter / (aarea + barea - inter);
    return (double)ovr > thr;
}

// ---- COCO matching of the kept detections of one image (block-wide) -------------------------------------
// dets: raw boxes/score/class in shared memory; result bit t of tp[d] = TP at iou_thrs[t].
struct MatchSmem {
    float x1[MAX_DET_CAP], y1[MAX_DET_CAP], x2[MAX_DET_CAP], y2[MAX_DET_CAP], score[MAX_DET_CAP];
    int cls[MAX_DET_CAP];
    int tp[MAX_DET_CAP];
    unsigned char taken[KODB_NUM_IOU_THRS][GT_CAP];
};

__device__ __forceinline__ void load_gt_xyxy(const void* boxes, int dtype, int g, double& l, double& t, double& r, double& b) {
    if (dtype == KODB_BOXES_F64) {
        const double* q = reinterpret_cast<const double*>(boxes) + 4 * (int64_t)g;
        l = q[0]; t = q[1]; r = q[2]; b = q[3];
    } else {
        const float* q = reinterpret_cast<const float*>(boxes) + 4 * (int64_t)g;
        l = q[0]; t = q[1]; r = q[2]; b = q[3];
    }
}

__device__ void match_image(MatchSmem& ms, int nk, int C, const void* gt_boxes, const int64_t* gt_labels, int g0, int ng,
                            int gt_dtype, const kodb_match_out& mo, int b, int max_det) {
    const int tid = threadIdx.x;
    for (int i = tid; i < nk; i += blockDim.x) ms.tp[i] = 0;
    for (int i = tid; i < KODB_NUM_IOU_THRS * ng; i += blockDim.x) ms.taken[i / ng][i % ng] = 0;
    for (int g = tid; g < ng; g += blockDim.x) {
        const long long lab = gt_labels[g0 + g];
        if (lab >= 0 && lab < C) atomicAdd(reinterpret_cast<unsigned long long*>(mo.gt_count) + lab, 1ull);
    }
    __syncthreads();
    // one thread per (class, threshold): sequential greedy matching in detection (score) order
    for (int w = tid; w < C * KODB_NUM_IOU_THRS; w += blockDim.x) {
        const int c = w / KODB_NUM_IOU_THRS, t = w - c * KODB_NUM_IOU_THRS;
        const double thr = fmin(mo.iou_thrs[t], 1.0 - 1e-10);
        for (int d = 0; d < nk; ++d) {
            if (ms.cls[d] != c) continue;
            const double dl = ms.x1[d], dt = ms.y1[d];
            const double dw = (double)ms.x2[d] - dl, dh = (double)ms.y2[d] - dt;
            const double da = dw * dh;
            double best = thr;
            int m = -1;
            for (int g = 0; g < ng; ++g) {
                if (gt_labels[g0 + g] != c || ms.taken[t][g]) continue;
                double gl, gt_, gr, gb;
                load_gt_xyxy(gt_boxes, gt_dtype, g0 + g, gl, gt_, gr, gb);
                const double gw = gr - gl, gh = gb - gt_;
                // maskApi.c::bbIou on [x,y,w,h]
                const double iw = fmin(dl + dw, gl + gw) - fmax(dl, gl);
                if (iw <= 0) { if (0.0 < best) continue; }
                const double ih = fmin(dt + dh, gt_ + gh) - fmax(dt, gt_);
                double iou = 0.0;
                if (iw > 0 && ih > 0) {
                    const double i = iw * ih;
                    iou = i / (da + gw * gh - i);
                }
                if (iou < best) continue;
                best = iou;
                m = g;
            }
            if (m >= 0) {
                ms.taken[t][m] = 1;
                atomicOr(&ms.tp[d], 1 << t);
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < max_det; i += blockDim.x) {
        const int64_t o = (int64_t)b * max_det + i;
        const bool v = i < nk;
        mo.rec_score[o] = v ? ms.score[i] : 0.0f;
        mo.rec_cls[o] = v ? ms.cls[i] : -1;
        mo.rec_tp[o] = v ? (uint8_t)ms.tp[i] : (uint8_t)0;
    }
}

__global__ void __launch_bounds__(NMS_THREADS, 1) select_nms_kernel(const __grid_constant__ EvalParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    NmsSmem& s = *reinterpret_cast<NmsSmem*>(smem_raw);
    MatchSmem& ms = *reinterpret_cast<MatchSmem*>(smem_raw);  // aliases NmsSmem: used after NMS is finished

    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int C = p.C;
    const int n = p.img_count[b];
    const float* score = p.cand_score + (int64_t)b * p.cand_stride;
    const uint32_t* key = p.cand_key + (int64_t)b * p.cand_stride;
    const int max_det = p.cfg.max_det;
    const bool rank_mode = n > p.cfg.max_nms;  // nms.py:61-63: candidates re-ordered by confidence, top max_nms kept
    const int limit = rank_mode ? p.cfg.max_nms : n;
    const double iou_thr = p.cfg.iou_thres;
    const float max_wh = p.cfg.max_wh;
    int* tcount = p.tile_count + (int64_t)b * p.tiles_per_image;
    const int* tstart = p.tile_start + (int64_t)b * p.tiles_per_image;

    if (tid == 0) { s.nk = 0; }
    // exclusive prefix of the per-tile candidate counts (tile order == row order) for index reconstruction
    {
        int carry = 0;
        for (int t0 = 0; t0 < p.tiles_per_image; t0 += NMS_THREADS) {
            const int t = t0 + tid;
            const int v = t < p.tiles_per_image ? tcount[t] : 0;
            int tot;
            const int ex = block_excl_scan(v, s.scan_tmp, &tot);
            if (t < p.tiles_per_image) tcount[t] = carry + ex;
            carry += tot;
        }
    }
    __syncthreads();

    int nk = 0, processed = 0;
    unsigned long long lo = 0ull;
    const unsigned long long rlo0 = (unsigned long long)desc_key(1.0f) << 32;

    while (nk < max_det && processed < limit) {
        // ------------------------------------------------------------------ select the next batch [lo, hi)
        unsigned long long glo = lo, ghi = U64MAX, rlo = rlo0, hi_batch = U64MAX;
        int shift = 47, m = 0;
        for (;;) {
            for (int i = tid; i < NB; i += NMS_THREADS) s.hist[i] = 0;
            if (tid == 0) s.sel_bin = -1;
            __syncthreads();
            for (int i = tid; i < n; i += NMS_THREADS) {
                const unsigned long long ck = make_ck(score[i], key[i]);
                if (ck >= glo && ck < ghi) {
                    const unsigned long long d = ck < rlo ? 0ull : ((ck - rlo) >> shift);
                    atomicAdd(&s.hist[d > NB - 1 ? NB - 1 : (int)d], 1);
                }
            }
            __syncthreads();
            const int c0 = s.hist[4 * tid], c1 = s.hist[4 * tid + 1], c2 = s.hist[4 * tid + 2], c3 = s.hist[4 * tid + 3];
            int total;
            const int ex = block_excl_scan(c0 + c1 + c2 + c3, s.scan_tmp, &total);
            const int p0 = ex + c0, p1 = p0 + c1, p2 = p1 + c2, p3 = p2 + c3;
            if (p0 >= K_TARGET && ex < K_TARGET) { s.sel_bin = 4 * tid; s.sel_p = p0; s.sel_prev = ex; }
            else if (p1 >= K_TARGET && p0 < K_TARGET) { s.sel_bin = 4 * tid + 1; s.sel_p = p1; s.sel_prev = p0; }
            else if (p2 >= K_TARGET && p1 < K_TARGET) { s.sel_bin = 4 * tid + 2; s.sel_p = p2; s.sel_prev = p1; }
            else if (p3 >= K_TARGET && p2 < K_TARGET) { s.sel_bin = 4 * tid + 3; s.sel_p = p3; s.sel_prev = p2; }
            __syncthreads();
            const int d = s.sel_bin;
            if (d < 0) { hi_batch = ghi; m = total; break; }  // fewer than K_TARGET left in range: take them all
            const unsigned long long blo = d == 0 ? glo : max(glo, rlo + ((unsigned long long)d << shift));
            const unsigned long long bhi = d == NB - 1 ? ghi : min(ghi, rlo + ((unsigned long long)(d + 1) << shift));
            if (s.sel_p <= K_CAP) { hi_batch = bhi; m = s.sel_p; break; }
            if (s.sel_prev > 0) { hi_batch = blo; m = s.sel_prev; break; }
            // a single bin overflows the sort capacity: refine inside it
            glo = blo; ghi = bhi; rlo = blo;
            const unsigned long long span = ghi - glo - 1ull;
            const int nbits = span ? 64 - __clzll((long long)span) : 0;
            shift = nbits > 12 ? nbits - 12 : 0;
            __syncthreads();
        }
        if (m == 0) break;  // nothing left
        // ------------------------------------------------------------------ collect + sort
        if (tid == 0) s.cnt = 0;
        __syncthreads();
        for (int i = tid; i < n; i += NMS_THREADS) {
            const unsigned long long ck = make_ck(score[i], key[i]);
            if (ck >= lo && ck < hi_batch) {
                const int slot = atomicAdd(&s.cnt, 1);
                s.sort_ck[slot] = ck;
                s.sort_pos[slot] = (uint32_t)i;
            }
        }
        __syncthreads();
        m = s.cnt;  // == the histogram count
        int npad = 32;
        while (npad < m) npad <<= 1;
        for (int i = m + tid; i < npad; i += NMS_THREADS) { s.sort_ck[i] = U64MAX; s.sort_pos[i] = 0; }
        __syncthreads();
        for (int k = 2; k <= npad; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < npad; i += NMS_THREADS) {
                    const int ixj = i ^ j;
                    if (ixj > i) {
                        const unsigned long long a = s.sort_ck[i], c = s.sort_ck[ixj];
                        const bool up = (i & k) == 0;
                        if ((a > c) == up) {
                            s.sort_ck[i] = c; s.sort_ck[ixj] = a;
                            const uint32_t pa = s.sort_pos[i]; s.sort_pos[i] = s.sort_pos[ixj]; s.sort_pos[ixj] = pa;
                        }
                    }
                }
                __syncthreads();
            }
        }
        int mm = m;
        if (processed + mm > limit) mm = limit - processed;  // only the top max_nms enter NMS
        // ------------------------------------------------------------------ decode the batch (offset boxes)
        for (int e = tid; e < mm; e += NMS_THREADS) {
            const uint32_t k32 = (uint32_t)s.sort_ck[e];
            const int row = (int)(k32 / (uint32_t)C), j = (int)(k32 - (uint32_t)row * (uint32_t)C);
            const float4 bx = fetch_box(p, b, row);
            const float off = (float)j * max_wh;  // nms.py:66
            const float x1 = bx.x + off, y1 = bx.y + off, x2 = bx.z + off, y2 = bx.w + off;
            s.bx1[e] = x1; s.by1[e] = y1; s.bx2[e] = x2; s.by2[e] = y2;
            s.barea[e] = (x2 - x1) * (y2 - y1);
        }
        __syncthreads();
        // ------------------------------------------------------------------ greedy NMS, 32 candidates per step
        for (int s0 = 0; s0 < mm && nk < max_det; s0 += 32) {
            const int e = s0 + wid;  // warp `wid` owns candidate e
            bool sup = false;
            unsigned msk = 0u;
            if (e < mm) {
                const float cx1 = s.bx1[e], cy1 = s.by1[e], cx2 = s.bx2[e], cy2 = s.by2[e], ca = s.barea[e];
                for (int q = lane; q < nk; q += 32)
                    sup |= suppresses(s.kx1[q], s.ky1[q], s.kx2[q], s.ky2[q], s.karea[q], cx1, cy1, cx2, cy2, ca, iou_thr);
                const int o = s0 + lane;
                bool hit = false;
                if (lane < wid && o < mm)
                    hit = suppresses(s.bx1[o], s.by1[o], s.bx2[o], s.by2[o], s.barea[o], cx1, cy1, cx2, cy2, ca, iou_thr);
                msk = __ballot_sync(KODB_FULL_MASK, hit);
            }
            const bool any = __any_sync(KODB_FULL_MASK, sup);
            if (lane == 0) { s.alive[wid] = (e < mm) && !any; s.mask[wid] = msk; }
            __syncthreads();
            if (wid == 0) {
                unsigned keptbits = 0u;
                if (lane == 0) {
                    for (int w = 0; w < 32; ++w)
                        if (s.alive[w] && !(s.mask[w] & keptbits)) keptbits |= 1u << w;
                }
                keptbits = __shfl_sync(KODB_FULL_MASK, keptbits, 0);
                if ((keptbits >> lane) & 1u) {
                    const int slot = nk + __popc(keptbits & ((1u << lane) - 1u));
                    if (slot < max_det) {
                        const int o = s0 + lane;
                        s.kx1[slot] = s.bx1[o]; s.ky1[slot] = s.by1[o]; s.kx2[slot] = s.bx2[o]; s.ky2[slot] = s.by2[o];
                        s.karea[slot] = s.barea[o];
                        const unsigned long long ck = s.sort_ck[o];
                        s.kept_ck[slot] = ck;
                        long long idx;
                        if (rank_mode) {
                            idx = processed + o;
                        } else {
                            const uint32_t k32 = (uint32_t)ck;
                            const int tl = tile_of_row(p, (int)(k32 / (uint32_t)C));
                            idx = (long long)tcount[tl] + ((long long)s.sort_pos[o] - (long long)tstart[tl]);
                        }
                        s.kept_idx[slot] = idx;
                    }
                }
                if (lane == 0) s.nk = min(max_det, nk + __popc(keptbits));
            }
            __syncthreads();
            nk = s.nk;
        }
        processed += m;
        lo = hi_batch;
        if (hi_batch == U64MAX) break;
    }
    __syncthreads();

    // ---------------------------------------------------------------------- outputs
    // stage the kept detections (un-offset boxes, nms.py:73) in registers, then reuse shared memory for matching
    float4 rb = make_float4(0.f, 0.f, 0.f, 0.f);
    float rs = 0.0f;
    int rc = -1;
    long long ridx = 0;
    if (tid < nk) {
        const unsigned long long ck = s.kept_ck[tid];
        const uint32_t k32 = (uint32_t)ck;
        const int row = (int)(k32 / (uint32_t)C);
        rc = (int)(k32 - (uint32_t)row * (uint32_t)C);
        rs = score_of_key((uint32_t)(ck >> 32));
        rb = fetch_box(p, b, row);
        ridx = s.kept_idx[tid];
    }
    if (tid < max_det) {
        float* d = p.out.dets + ((int64_t)b * max_det + tid) * 6;
        const bool v = tid < nk;
        d[0] = rb.x; d[1] = rb.y; d[2] = rb.z; d[3] = rb.w; d[4] = rs; d[5] = v ? (float)rc : 0.0f;
        p.out.keep_idx[(int64_t)b * max_det + tid] = v ? ridx : -1;
    }
    if (tid == 0) {
        p.out.counts[b] = nk;
        if (p.out.num_candidates) p.out.num_candidates[b] = n;
        p.img_count[b] = 0;  // workspace reusable without a memset
    }
    if (p.do_match) {
        __syncthreads();  // everyone is done with NmsSmem
        if (tid < nk) {
            ms.x1[tid] = rb.x; ms.y1[tid] = rb.y; ms.x2[tid] = rb.z; ms.y2[tid] = rb.w;
            ms.score[tid] = rs; ms.cls[tid] = rc;
        }
        __syncthreads();
        const int g0 = p.gt_offsets[b], ng = p.gt_offsets[b + 1] - g0;
        match_image(ms, nk, C, p.gt_boxes, p.gt_labels, g0, ng, p.gt_dtype, p.m, b, max_det);
    }
}

// standalone matching on NMS output (B, max_det, 6)
__global__ void __launch_bounds__(NMS_THREADS, 1)
match_kernel(const float* __restrict__ dets, const int32_t* __restrict__ counts, int C, int max_det, const void* gt_boxes,
             const int64_t* gt_labels, const int32_t* gt_offsets, int gt_dtype, const __grid_constant__ kodb_match_out mo) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MatchSmem& ms = *reinterpret_cast<MatchSmem*>(smem_raw);
    const int b = blockIdx.x, tid = threadIdx.x;
    const int nk = min(counts[b], max_det);
    if (tid < nk) {
        const float* d = dets + ((int64_t)b * max_det + tid) * 6;
        ms.x1[tid] = d[0]; ms.y1[tid] = d[1]; ms.x2[tid] = d[2]; ms.y2[tid] = d[3];
        ms.score[tid] = d[4]; ms.cls[tid] = (int)d[5];
    }
    __syncthreads();
    const int g0 = gt_offsets[b], ng = gt_offsets[b + 1] - g0;
    match_image(ms, nk, C, gt_boxes, gt_labels, g0, ng, gt_dtype, mo, b, max_det);
}

// ---------------------------------------------------------------------------------------------------------
static int plan_geom(EvalGeom& g, const kodb_head_level* heads_h, const kodb_level_cfg* levels_h) {
    int rows = 0, tiles = 0;
    for (int l = 0; l < KODB_NUM_LEVELS; ++l) {
        g.head[l] = heads_h[l];
        g.lv[l] = levels_h[l];
        const int HW = levels_h[l].H * levels_h[l].W;
        if (HW <= 0 || levels_h[l].num_anchors < 1 || levels_h[l].num_anchors > KODB_MAX_ANCHORS) return -1;
        g.row_base[l] = rows;
        rows += levels_h[l].num_anchors * HW;
        const int tpp = (HW + EV_TILE - 1) / EV_TILE;
        g.tiles_per_plane[l] = tpp;
        g.tile_cells[l] = (HW + tpp - 1) / tpp;
        g.tile_begin[l] = tiles;
        tiles += tpp * levels_h[l].num_anchors;
    }
    g.row_base[KODB_NUM_LEVELS] = rows;
    g.tile_begin[KODB_NUM_LEVELS] = tiles;
    return tiles;
}

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

static int64_t ws_bytes(int batch, int64_t rows, int C, int tiles_per_image) {
    const size_t cap = (size_t)rows * C;
    return (int64_t)(align256((size_t)batch * 4) + 2 * align256((size_t)batch * tiles_per_image * 4) +
                     2 * align256((size_t)batch * cap * 4));
}

static int carve_ws(EvalParams& p, void* ws, int64_t ws_bytes_avail) {
    const size_t cap = (size_t)p.rows * p.C;
    if (ws_bytes_avail < ws_bytes(p.B, p.rows, p.C, p.tiles_per_image)) return KODB_ERR_CAPACITY;
    char* w = reinterpret_cast<char*>(ws);
    p.img_count = reinterpret_cast<int*>(w);
    w += align256((size_t)p.B * 4);
    p.tile_start = reinterpret_cast<int*>(w);
    w += align256((size_t)p.B * p.tiles_per_image * 4);
    p.tile_count = reinterpret_cast<int*>(w);
    w += align256((size_t)p.B * p.tiles_per_image * 4);
    p.cand_score = reinterpret_cast<float*>(w);
    w += align256((size_t)p.B * cap * 4);
    p.cand_key = reinterpret_cast<uint32_t*>(w);
    p.cand_stride = (int64_t)cap;
    return KODB_OK;
}

static int check_common(const kodb_nms_cfg* cfg, const kodb_nms_out* out, int batch, int64_t rows) {
    if (!cfg || !out || !out->dets || !out->keep_idx || !out->counts || batch < 1 || batch > 65535 || rows < 1)
        return KODB_ERR_INVALID_ARGUMENT;
    if (cfg->num_classes < 1 || cfg->num_classes > KODB_MAX_CLASSES) return KODB_ERR_INVALID_ARGUMENT;
    if (cfg->max_det < 1 || cfg->max_det > MAX_DET_CAP || cfg->max_nms < 1) return KODB_ERR_INVALID_ARGUMENT;
    if ((double)rows * cfg->num_classes >= 4294967295.0) return KODB_ERR_UNSUPPORTED;
    return KODB_OK;
}

static int setup_match(EvalParams& p, const kodb_targets* gt, const kodb_match_out* m) {
    p.do_match = 0;
    if (!gt) return KODB_OK;
    if (!m || !m->rec_score || !m->rec_cls || !m->rec_tp || !m->gt_count) return KODB_ERR_INVALID_ARGUMENT;
    if (gt->batch != p.B) return KODB_ERR_INVALID_ARGUMENT;
    if (gt->max_per_image > GT_CAP) return KODB_ERR_CAPACITY;
    p.do_match = 1;
    p.gt_boxes = gt->boxes;
    p.gt_labels = gt->labels;
    p.gt_offsets = gt->img_offsets;
    p.gt_dtype = gt->boxes_dtype;
    p.m = *m;
    return KODB_OK;
}

static int launch_select(const EvalParams& p, cudaStream_t st) {
    const size_t smem = sizeof(NmsSmem) > sizeof(MatchSmem) ? sizeof(NmsSmem) : sizeof(MatchSmem);
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        if (cudaFuncSetAttribute(select_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return KODB_ERR_LAUNCH;
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    select_nms_kernel<<<p.B, NMS_THREADS, smem, st>>>(p);
    KODB_CHECK_LAUNCH();
    return KODB_OK;
}

}  // namespace kodb

extern "C" int64_t kodb_eval_workspace_bytes(int32_t batch, int64_t rows_per_image, int32_t num_classes) {
    // tiles per image <= rows/1 in the worst planar case; bound it by one tile per 1 row-plane remainder:
    // planar tiles = sum_l A*ceil(HW/256) <= rows/256 + 3*KODB_MAX_ANCHORS
    const int tiles = (int)((rows_per_image + kodb::EV_TILE - 1) / kodb::EV_TILE) + KODB_NUM_LEVELS * KODB_MAX_ANCHORS;
    return kodb::ws_bytes(batch, rows_per_image, num_classes, tiles);
}

extern "C" int kodb_eval_tail(const kodb_head_level* heads_h, const kodb_level_cfg* levels_h, int32_t batch,
                              const kodb_nms_cfg* cfg_h, const kodb_nms_out* out_h, const kodb_targets* gt_h,
                              const kodb_match_out* match_h, void* workspace, int64_t workspace_bytes, void* stream) {
    using namespace kodb;
    if (!heads_h || !levels_h || !workspace) return KODB_ERR_INVALID_ARGUMENT;
    EvalParams p;
    const int tiles = plan_geom(p.g, heads_h, levels_h);
    if (tiles < 0) return KODB_ERR_INVALID_ARGUMENT;
    p.det = nullptr;
    p.rows = p.g.row_base[KODB_NUM_LEVELS];
    int rc = check_common(cfg_h, out_h, batch, p.rows);
    if (rc != KODB_OK) return rc;
    p.planar = 1;
    p.B = batch;
    p.C = cfg_h->num_classes;
    p.cfg = *cfg_h;
    p.out = *out_h;
    p.tiles_per_image = tiles;
    if ((rc = carve_ws(p, workspace, workspace_bytes)) != KODB_OK) return rc;
    if ((rc = setup_match(p, gt_h, match_h)) != KODB_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int stages = cfg_h->stage_mask ? cfg_h->stage_mask : 3;
    if (stages & 1) {
        cand_kernel<true><<<dim3(tiles, batch), EV_TILE, 0, st>>>(p);
        KODB_CHECK_LAUNCH();
    }
    return (stages & 2) ? launch_select(p, st) : KODB_OK;
}

extern "C" int kodb_nms(const float* detections, int32_t batch, int64_t rows, const kodb_nms_cfg* cfg_h,
                        const kodb_nms_out* out_h, const kodb_targets* gt_h, const kodb_match_out* match_h,
                        void* workspace, int64_t workspace_bytes, void* stream) {
    using namespace kodb;
    if (!detections || !workspace) return KODB_ERR_INVALID_ARGUMENT;
    int rc = check_common(cfg_h, out_h, batch, rows);
    if (rc != KODB_OK) return rc;
    EvalParams p;
    p.g = EvalGeom{};
    p.det = detections;
    p.rows = rows;
    p.planar = 0;
    p.B = batch;
    p.C = cfg_h->num_classes;
    p.cfg = *cfg_h;
    p.out = *out_h;
    p.tiles_per_image = (int)((rows + EV_TILE - 1) / EV_TILE);
    if ((rc = carve_ws(p, workspace, workspace_bytes)) != KODB_OK) return rc;
    if ((rc = setup_match(p, gt_h, match_h)) != KODB_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int stages = cfg_h->stage_mask ? cfg_h->stage_mask : 3;
    if (stages & 1) {
        cand_kernel<false><<<dim3(p.tiles_per_image, batch), EV_TILE, 0, st>>>(p);
        KODB_CHECK_LAUNCH();
    }
    return (stages & 2) ? launch_select(p, st) : KODB_OK;
}

extern "C" int kodb_map_match(const float* dets, const int32_t* counts, int32_t batch, int32_t max_det,
                              int32_t num_classes, const kodb_targets* gt_h, const kodb_match_out* match_h, void* stream) {
    using namespace kodb;
    if (!dets || !counts || !gt_h || !match_h || batch < 1 || max_det < 1 || max_det > MAX_DET_CAP || num_classes < 1 ||
        num_classes > KODB_MAX_CLASSES)
        return KODB_ERR_INVALID_ARGUMENT;
    if (gt_h->batch != batch) return KODB_ERR_INVALID_ARGUMENT;
    if (gt_h->max_per_image > GT_CAP) return KODB_ERR_CAPACITY;
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        if (cudaFuncSetAttribute(match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MatchSmem)) != cudaSuccess)
            return KODB_ERR_LAUNCH;
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    match_kernel<<<batch, NMS_THREADS, sizeof(MatchSmem), (cudaStream_t)stream>>>(
        dets, counts, num_classes, max_det, gt_h->boxes, gt_h->labels, gt_h->img_offsets, gt_h->boxes_dtype, *match_h);
    KODB_CHECK_LAUNCH();
    return KODB_OK;
}
