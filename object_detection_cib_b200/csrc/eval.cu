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
//  1. Candidates.  Fused path (planar head, planes divisible into float4): cand_floor_kernel, one WARP per tile of
//     128 cells of one (level, anchor) plane -- see the comment in front of it: the complete candidate set
//     (obj > thr and obj*cls > thr, nms.py:33-48) goes into 16-bit class masks, and only the candidates at or above
//     a per-image score floor (measured by the image's sample CTA) are listed as (score, key = row*C + class).
//     Decoded (B, N, 5+C) input (kodb_nms) and odd grids: cand_kernel, one thread per row / cell, lists everything.
//     Chunks and the entries inside a chunk land in arbitrary order: order is re-established later from the keys.
//     A 512-bin score histogram per image is accumulated on the side.  Boxes are NOT decoded here.
//  2. select_nms_kernel (one 512-thread CTA per image, two per SM): exact extraction of the next-best batch in
//     (score desc, key asc) order -- cut taken from the coarse histogram (restricted to the bins the list is complete
//     for); if one bin overflows the sort capacity (ties, skewed scores) it is refined by a radix-select over the
//     64-bit composite key -- candidates collected straight into per-bin segments and ordered by counting (bitonic
//     fallback), box decode for the batch only, class-bucketed greedy NMS with bit-mask tiles (generic 64-wide step
//     path when classes could interact), early exit at max_det keeps (exact: the first 300 keeps depend only on the
//     scanned prefix).  fp32 IoU arithmetic is torchvision's CPU kernel op for op (offset boxes, ratio vs DOUBLE
//     threshold).  The reference's candidate index of each survivor (its rank among all keys of the image) is rebuilt
//     from per-tile counts + the class masks (fused path) or a count inside its own chunk.  If the list runs dry
//     before max_det survivors, a completion pass lists the rest of the image from the masks.  The kept detections
//     then go through the fp64 COCO matching while still in shared memory.
#include <stddef.h>
#include <stdlib.h>

#include <mutex>

#include "kodb_common.cuh"
#include "tma_common.cuh"

namespace kodb {

constexpr int CAND_THREADS = 256;
constexpr int CGROUP = 16;       // classes per pass of the candidate kernel
constexpr int HB = 512;          // coarse score-histogram bins per image (32 per octave below 1.0)
constexpr int NMS_THREADS = 512;  // two select CTAs fit on an SM (registers, shared memory), or one next to other kernels
constexpr int LOG_NB = 11;
constexpr int NB = 1 << LOG_NB;  // fine histogram bins per refinement level (4 per thread)
constexpr int K_CAP = NMS_THREADS;  // sort capacity = one element per thread (batch size upper bound)
constexpr int K_TARGET0 = 352;   // preferred size of the first batch (max_det 300 + slack); doubled for later batches
constexpr int NMS_STEP = NMS_THREADS / 8;  // candidates resolved per step of the generic NMS path (4 per warp)
constexpr int MAX_DET_CAP = NMS_THREADS;  // max_det <= 512 (the reference uses 300)
constexpr int GT_CAP = NMS_THREADS;  // max ground-truth boxes of ONE image in the matching stage
constexpr int CHUNK_CAP = 4096;  // max candidate chunks (tiles x class groups) per image
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
    int vec;           // cells per thread in the candidate kernel (4: float4 planes, 1: generic)
    int ngroups;       // ceil(C / CGROUP)
    kodb_nms_cfg cfg;
    // workspace
    float* cand_score;
    uint32_t* cand_key;
    int64_t cand_stride;
    int* img_count;    // [B]       zero on entry / exit
    int* hist;         // [B][HB]   zero on entry / exit
    int* chunk_start;  // [B][chunks_per_image]
    int* chunk_count;  // [B][chunks_per_image]
    int* chunk_prefix; // [B][chunks_per_image]  written and read by the select kernel
    int tiles_per_image, chunks_per_image;
    // floor mode (planar float4 path): only candidates at or above a per-image score floor are listed; every
    // candidate is recorded as a bit in `mask` (keep indices, the rare completion pass)
    int floor_mode;
    int ns, sample_stride;     // sample tiles per image: tiles 0, S, 2S, ... (one per warp of the image's sample CTA)
    unsigned magic_b, magic_s1; // ceil(2^32 / B), ceil(2^32 / (S - 1)) when exact over the grid's index range, else 0
    int hot_target;            // estimated number of listed candidates per image the floor aims at
    // tile loads by TMA (flat planar head tensors): per level (cell, plane, image) maps of the objectness and class logits,
    // boxes of 128 cells x 1 plane / 128 cells x NQ planes; 0: per-lane cp.async copies (any strides)
    int use_tma;
    CUtensorMap tma_obj[KODB_NUM_LEVELS], tma_cls[KODB_NUM_LEVELS];
    int l2_stream;             // class planes are loaded with the L2 evict-first policy (tuning knob KODB_EVAL_L2_STREAM)
    int floor_polls;           // bounded wait (x100 ns) for the floor at a tile's second chance; 0 in production (no gain
                               // measured on B200), large in the tests that must exercise floors on small batches
    int* floor_word;           // [B]  (floor bin + 1), 0 = not published yet; zero on entry / exit
    unsigned short* mask;      // [B][ngroups][rows]  bit u: class 16*group + u of the row is a candidate
    int64_t mask_stride;       // elements per image
    kodb_nms_out out;
    // matching
    int do_match;
    const void* gt_boxes;
    const int64_t* gt_labels;
    const int32_t* gt_offsets;
    int gt_dtype;
    kodb_match_out m;
    long long* trace;  // debug: per-image phase timestamps (clock64), NULL in production
};

#define KODB_TRACE(id)                                                                  \
    do {                                                                                \
        if (p.trace && tid == 0 && trace_n < 62) {                                      \
            p.trace[(int64_t)b * 128 + 2 + 2 * trace_n] = (id);                         \
            p.trace[(int64_t)b * 128 + 3 + 2 * trace_n] = clock64();                    \
            ++trace_n;                                                                  \
            p.trace[(int64_t)b * 128] = trace_n;                                        \
        }                                                                               \
    } while (0)

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
// coarse bin: 0 = scores >= 1.0, then 32 bins per octave downwards; negatives / tiny scores -> HB-1
__device__ __forceinline__ int coarse_bin(float s) {
    const uint32_t b = __float_as_uint(s);
    if (b >> 31) return HB - 1;
    const int d = (int)(0x3F800000u >> 18) - (int)(b >> 18);
    return d < 0 ? 0 : (d > HB - 1 ? HB - 1 : d);
}
// exclusive upper bound, in composite-key space, of coarse bins 0..d
__device__ __forceinline__ unsigned long long coarse_hi(int d) {
    if (d >= HB - 1) return U64MAX;
    const uint32_t lowest_bits = ((0x3F800000u >> 18) - (uint32_t)d) << 18;  // smallest score bits inside bin d
    return (unsigned long long)(0x80000000u - lowest_bits) << 32;
}

// ---- geometry helpers -----------------------------------------------------------------------------------
__device__ __forceinline__ void planar_tile(const EvalGeom& g, int t, int& L, int& a, int& c0, int& ncell) {
    L = t >= g.tile_begin[2] ? 2 : (t >= g.tile_begin[1] ? 1 : 0);
    int r = t - g.tile_begin[L];
    const int tpp = g.tiles_per_plane[L];
    a = 0;
    while (r >= tpp) { r -= tpp; ++a; }  // a < KODB_MAX_ANCHORS: cheaper than an integer division
    c0 = r * g.tile_cells[L];
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
    if (!p.planar) return row / CAND_THREADS;
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
    return decode_box(__ldcs(hd.box + bo), __ldcs(hd.box + bo + hd.box_sp), __ldcs(hd.box + bo + 2 * hd.box_sp),
                      __ldcs(hd.box + bo + 3 * hd.box_sp), (float)gx,
                      (float)gy, (float)lv.stride, lv.anchor_px[a][0], lv.anchor_px[a][1]);
}

// L2 prefetch of the four box logits (or the decoded row) a later fetch_box(p, b, row) will read
__device__ __forceinline__ void prefetch_box(const EvalParams& p, int b, int row) {
    if (!p.planar) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.det + ((int64_t)b * p.rows + row) * (5 + p.C)));
        return;
    }
    int L, a, cell;
    planar_row(p.g, row, L, a, cell);
    const kodb_head_level& hd = p.g.head[L];
    const float* q = hd.box + (int64_t)b * hd.box_sb + (int64_t)a * hd.box_sa + cell;
#pragma unroll
    for (int k = 0; k < 4; ++k) asm volatile("prefetch.global.L2 [%0];" ::"l"(q + k * hd.box_sp));
}

// =========================================================================================================
// 1. candidates
// =========================================================================================================
template <int VEC>
struct VecLoad;
template <>
struct VecLoad<4> {
    static __device__ __forceinline__ void ld(const float* p, float v[4]) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    }
};
template <>
struct VecLoad<1> {
    static __device__ __forceinline__ void ld(const float* p, float v[1]) { v[0] = __ldg(p); }
};

template <bool PLANAR, int VEC>
__global__ void __launch_bounds__(CAND_THREADS) cand_kernel(const __grid_constant__ EvalParams p) {
    __shared__ unsigned short s_list[CAND_THREADS * VEC * CGROUP];
    __shared__ float s_o[CAND_THREADS * VEC];  // sigmoid(obj) of the tile's cells (exact values for pass 2)
    __shared__ int s_hist[HB];
    __shared__ int s_cnt, s_base;
    const int b = blockIdx.y, t = blockIdx.x, tid = threadIdx.x;
    const int C = p.C;
    const float thr = p.cfg.conf_thres;
    int row0, ncell;
    const float* obj_base;  // address of the tile's first obj value
    const float* cls_base;  // address of the tile's first class-0 value
    int64_t cls_step, cell_step;
    if (PLANAR) {
        int L, a, c0;
        planar_tile(p.g, t, L, a, c0, ncell);
        const kodb_head_level& hd = p.g.head[L];
        row0 = p.g.row_base[L] + a * (p.g.lv[L].H * p.g.lv[L].W) + c0;
        obj_base = hd.obj + (int64_t)b * hd.obj_sb + (int64_t)a * hd.obj_sa + c0;
        cls_base = hd.cls + (int64_t)b * hd.cls_sb + (int64_t)a * hd.cls_sa + c0;
        cls_step = hd.cls_sp;
        cell_step = 1;
    } else {
        row0 = t * CAND_THREADS;
        ncell = (int)min((int64_t)CAND_THREADS, p.rows - row0);
        const float* r = p.det + ((int64_t)b * p.rows + row0) * (5 + C);
        obj_base = r + 4;
        cls_base = r + 5;
        cls_step = 1;
        cell_step = 5 + C;
    }
    for (int i = tid; i < HB; i += CAND_THREADS) s_hist[i] = 0;
    const int cell0 = tid * VEC;          // first local cell of this thread
    const bool in_range = cell0 < ncell;  // tiles are multiples of VEC cells when VEC == 4
    // Per cell: the objectness o (row filter, nms.py:33) and -- for the planar input -- the class-LOGIT window
    // around  logit(thr / o):  sigmoid(q) * o > thr  <=>  q > logit(thr / o)  in exact arithmetic.  Logits
    // outside the +-2e-3 window are decided by one comparison; only the (rare) ones inside pay for the exact
    // fp32 product, so the decision is always the one the decode kernel's arithmetic would give.
    float o[VEC], t_lo[VEC], t_hi[VEC];
    bool any_active = false;
    {
        float x[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) x[v] = 0.0f;
        if (in_range) VecLoad<VEC>::ld(obj_base + (int64_t)cell0 * cell_step, x);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const float ov = PLANAR ? sigmoid_eval(x[v]) : x[v];
            const bool act = in_range && ov > thr;
            o[v] = act ? ov : -1.0f;  // -1 marks "row filtered out"
            if (in_range) s_o[cell0 + v] = ov;
            any_active |= act;
            if (PLANAR) {
                float lo_ = 3.0e38f, hi_ = 3.0e38f;  // inactive cell: nothing passes
                if (act && thr > 0.0f) {
                    const float r = __fdividef(thr, ov);  // in (0, 1)
                    const float tau = __logf(__fdividef(r, 1.0f - r));
                    lo_ = tau - 2.0e-3f;
                    hi_ = tau + 2.0e-3f;
                    if (!(r < 0.99f)) { lo_ = -3.0e38f; hi_ = 3.0e38f; }  // ill-conditioned near r -> 1: exact test
                } else if (act) {  // thr <= 0: every class of an active row goes to the exact test
                    lo_ = -3.0e38f;
                    hi_ = 3.0e38f;
                }
                t_lo[v] = lo_;
                t_hi[v] = hi_;
            }
        }
    }
    for (int grp = 0; grp < p.ngroups; ++grp) {
        const int j0 = grp * CGROUP, jn = min(C - j0, CGROUP);
        if (tid == 0) s_cnt = 0;
        __syncthreads();
        // flagged (class, cell) pairs of this thread: bit (class_in_group * VEC + v)
        unsigned long long bits = 0ull;
        if (any_active) {
            const unsigned allowed = (p.cfg.class_mask[j0 >> 5] >> (j0 & 31)) & 0xffffu;  // `classes` filter (nms.py:54-55)
            for (int jb = 0; jb < jn; jb += 4) {
                float q[4][VEC];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (jb + u < jn) VecLoad<VEC>::ld(cls_base + (int64_t)(j0 + jb + u) * cls_step + (int64_t)cell0 * cell_step, q[u]);
                unsigned m16 = 0u;  // bit (u * VEC + v), compile-time positions
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (jb + u < jn) {
#pragma unroll
                        for (int v = 0; v < VEC; ++v) {
                            if (PLANAR) {
                                // inactive cells carry t_lo = +inf, so one comparison rejects ~90% of the pairs
                                if (q[u][v] > t_lo[v] && ((q[u][v] > t_hi[v]) || (sigmoid_eval(q[u][v]) * o[v] > thr)))
                                    m16 |= 1u << (u * VEC + v);
                            } else {
                                if (o[v] >= 0.0f && (q[u][v] * o[v] > thr)) m16 |= 1u << (u * VEC + v);  // nms.py:40, :47
                            }
                        }
                    }
                }
                if (allowed != 0xffffu) {  // rare: the `classes` argument was given
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (!((allowed >> (jb + u)) & 1u)) m16 &= ~(((1u << VEC) - 1u) << (u * VEC));
                }
                bits |= (unsigned long long)m16 << (jb * VEC);
            }
        }
        int slot = 0;
        const int mine = __popcll(bits);
        if (mine) slot = atomicAdd(&s_cnt, mine);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            unsigned w32 = half ? (unsigned)(bits >> 32) : (unsigned)bits;
            while (w32) {
                const int bi = __ffs(w32) - 1 + 32 * half;
                w32 &= w32 - 1;
                s_list[slot++] = (unsigned short)(((cell0 + (bi % VEC)) << 4) | (bi / VEC));
            }
        }
        __syncthreads();
        const int n = s_cnt;
        const int64_t chunk = (int64_t)b * p.chunks_per_image + (int64_t)t * p.ngroups + grp;
        if (tid == 0) {
            const int base = n ? atomicAdd(&p.img_count[b], n) : 0;
            s_base = base;
            p.chunk_start[chunk] = base;
            p.chunk_count[chunk] = n;
        }
        __syncthreads();
        float* sc = p.cand_score + (int64_t)b * p.cand_stride + s_base;
        uint32_t* ky = p.cand_key + (int64_t)b * p.cand_stride + s_base;
        for (int i = tid; i < n; i += CAND_THREADS) {
            const int pid = s_list[i];
            const int cl = pid >> 4, j = j0 + (pid & 15);
            const float xc = __ldg(cls_base + (int64_t)j * cls_step + (int64_t)cl * cell_step);
            // the decode kernel's arithmetic: sigmoid_eval(cls) * sigmoid_eval(obj)
            const float conf = (PLANAR ? sigmoid_eval(xc) : xc) * s_o[cl];
            sc[i] = conf;
            ky[i] = (uint32_t)(row0 + cl) * (uint32_t)C + (uint32_t)j;
            atomicAdd(&s_hist[coarse_bin(conf)], 1);
        }
        __syncthreads();
    }
    for (int i = tid; i < HB; i += CAND_THREADS)
        if (s_hist[i]) atomicAdd(&p.hist[(int64_t)b * HB + i], s_hist[i]);
}

// ---------------------------------------------------------------------------------------------------------
// 1b. warp-autonomous candidate kernel for the planar float4 case (the shipped fused path)
// ---------------------------------------------------------------------------------------------------------
// One WARP owns a tile of up to 128 contiguous cells of one (level, anchor) plane: lane l holds cells 4l..4l+3.
// The tile is brought into a warp-private shared stash in one go (TMA bulk tensor copies, or cp.async for odd strides),
// so one DRAM round trip covers the tile; there is no block-level barrier on the data path.  The per-pair
// decision is one subtraction whose sign bit is funnel-shifted into a mask:  sigmoid(q)*o > thr  <=>  q > tau(o)
// with tau = logit(thr/o) in exact arithmetic; pairs with |q - tau| <= TAU_WINDOW (a few per thousand) are
// re-decided with the decode kernel's exact fp32 product, so the candidate set is the one kodb_decode + kodb_nms
// would produce.  Flagged pairs become 16-bit ids in a warp-private shared list, then the warp writes them --
// balanced, one pair per lane -- as (score, key) into one contiguous chunk claimed with a single atomicAdd.
constexpr int WT_CELLS = 128;
constexpr int CW_WARPS = 4;   // 128-thread CTAs (32 KB smem at 10 classes): pack next to the select kernel's 100 KB CTAs
constexpr float TAU_WINDOW = 2.0e-4f;  // >> the ~1.5e-5 combined error of tau and of the fp32 product for thr/o < 0.9

__device__ __forceinline__ unsigned long long expand_nibbles(unsigned a16) {  // bit u -> bits 4u..4u+3
    unsigned long long x = a16 & 0xffffu;
    x = (x | (x << 24)) & 0x000000ff000000ffull;
    x = (x | (x << 12)) & 0x000f000f000f000full;
    x = (x | (x << 6)) & 0x0303030303030303ull;
    x = (x | (x << 3)) & 0x1111111111111111ull;
    return x * 15ull;
}

constexpr int HOT_TARGET = 2048;

// host-side caches (function attributes per device, tuning knobs) are shared by every host thread that calls into the library
static std::mutex g_eval_host_mu;

// smallest positive score of coarse bin d (d < HB - 1)
__device__ __forceinline__ float coarse_lower_edge(int d) { return __uint_as_float(((0x3F800000u >> 18) - (uint32_t)d) << 18); }

template <int NQ>
struct CandWarpSmem {  // one per warp: the data path of the kernel has no block-level barrier
    float q[NQ][WT_CELLS];               // class logits of the tile (stash: candidates are re-read by id)
    float o[WT_CELLS];                   // sigmoid(obj) per cell
    unsigned short list[WT_CELLS * NQ];  // ids of the (cell, class) pairs to list
};

// Candidate kernel of the fused path.  One WARP owns a tile of up to 128 contiguous cells of one (level, anchor)
// plane: lane l holds cells 4l..4l+3.  No block-level barrier, no waiting on other CTAs.
//  * The tile (128 cells x 1 objectness plane + 128 cells x NQ class planes) arrives in the warp's shared stash by two
//    TMA bulk tensor copies on a per-warp mbarrier (class planes with the L2 evict-first hint); head tensors that are
//    not flat planar arrays use per-lane cp.async 16-byte copies instead.  Either way: one DRAM round trip per tile.
//  * The per-pair decision is one subtraction whose sign bit is funnel-shifted into a mask:
//    sigmoid(q)*o > thr  <=>  q > tau(o) = logit(thr/o) in exact arithmetic; pairs with |q - tau| <= TAU_WINDOW (a
//    few per thousand) are re-decided with the decode kernel's exact fp32 product, so the candidate set is the one
//    kodb_decode + kodb_nms would produce.  The set is stored as one 16-bit class mask per cell.
//  * Only the candidates that can matter for the first max_det survivors are LISTED (score, key): those whose
//    score reaches a per-image floor.  The floor is measured, not guessed: the sample tiles (every 16th tile, the
//    lowest block indices, so they run first) list all their candidates and feed a sample histogram; the warp
//    finishing the image's last sample tile picks the histogram bin above which an estimated hot_target
//    candidates lie and publishes it.  A tile that finds the floor published lists only the pairs whose score
//    bin reaches it (a second sign mask against logit(floor/o)); a tile that comes too early lists everything, like
//    a sample tile, and says so in its tile flag.  The select kernel completes the list from the masks in the rare
//    case it runs dry.  Results never depend on the floor or on timing; only the amount of listing work does.
template <int NQ>
__device__ __forceinline__ unsigned fix_borderline(unsigned cmv, unsigned border, const float (*sq)[WT_CELLS], int cell, float ov,
                                                    float thr, int floor_bin) {
    while (border) {
        const int u = __ffs(border) - 1;
        border &= border - 1;
        cmv &= ~(1u << u);
        const float conf = sigmoid_eval(sq[u][cell]) * ov;  // the decode kernel's arithmetic
        if (ov > thr && conf > thr && coarse_bin(conf) <= floor_bin) cmv |= 1u << u;
    }
    return cmv;
}

template <int NQ>
__global__ void __launch_bounds__(CW_WARPS * 32, 6) cand_floor_kernel(const __grid_constant__ EvalParams p) {
    extern __shared__ __align__(128) unsigned char cw_smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    // (128-byte alignment of the stashes: destination of the TMA tile loads)
    // (offset arithmetic on the shared array itself, not an integer round trip: the compiler keeps the shared address
    // space and emits LDS / STS instead of generic loads and stores)
    unsigned char* cw_smem = cw_smem_raw + ((128u - ((unsigned)__cvta_generic_to_shared(cw_smem_raw) & 127u)) & 127u);
    CandWarpSmem<NQ>& sm = reinterpret_cast<CandWarpSmem<NQ>*>(cw_smem)[wid];
    __shared__ __align__(8) uint64_t s_tile_bar[CW_WARPS];  // one mbarrier per warp: arrival of the warp's tile (TMA path)
    // work order: CTA b < B is image b's sample CTA (one sample tile per warp, tiles 0, S, 2S, ...; it runs first and
    // keeps the sample histogram in shared memory); the other CTAs take the remaining tiles, image index fastest, so
    // that neighbouring warps work on different images and the per-image atomics do not pile up on one address
    __shared__ int s_shist[HB];
    const int T = p.tiles_per_image, S = p.sample_stride, ns = p.ns, nr = T - ns, B = p.B;
    const bool sample = (int)blockIdx.x < B;
    int b, t;
    bool active = true;
    if (sample) {
        b = blockIdx.x;
        t = wid * S;
        active = wid < ns;
        for (int i = tid; i < HB; i += CW_WARPS * 32) s_shist[i] = 0;
        __syncthreads();
    } else {
        const int g2 = ((int)blockIdx.x - B) * CW_WARPS + wid;
        if (g2 >= nr * B) return;
        // (divisions by multiply-high with host-checked magic numbers; 0 = not exact for this geometry, divide)
        const int ri = p.magic_b ? (int)__umulhi((unsigned)g2, p.magic_b) : g2 / B;
        b = g2 - ri * B;
        const int rq = p.magic_s1 ? (int)__umulhi((unsigned)ri, p.magic_s1) : ri / (S - 1);
        t = ri + rq + 1;  // ri-th tile index that is not a multiple of S (S >= 2 whenever nr > 0)
    }
    if (!active) {  // idle warp of a sample CTA: only the block barrier in front of the floor computation
        __syncthreads();
        return;
    }
    const int C = p.C;
    const float thr = p.cfg.conf_thres;
    int L, a, c0, ncell;
    planar_tile(p.g, t, L, a, c0, ncell);
    const kodb_head_level& hd = p.g.head[L];
    const int row0 = p.g.row_base[L] + a * (p.g.lv[L].H * p.g.lv[L].W) + c0;
    // (tile base addresses: needed by the cp.async path only -- computed where they are used, so that the TMA path,
    // which addresses the tile by coordinates, does not pay for the 64-bit arithmetic)
    auto obj_base_of = [&]() { return hd.obj + (int64_t)b * hd.obj_sb + (int64_t)a * hd.obj_sa + c0; };
    auto cls_base_of = [&]() { return hd.cls + (int64_t)b * hd.cls_sb + (int64_t)a * hd.cls_sa + c0; };
    const int cell0 = lane * 4;
    const bool in_range = cell0 < ncell;  // tiles are multiples of 4 cells
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    constexpr unsigned CELL_MASK = (1u << NQ) - 1u;  // the NQ class bits of one cell
    const unsigned sq0 = (unsigned)__cvta_generic_to_shared(&sm.q[0][cell0]);  // class u of the stash: + u * 512 bytes
    const unsigned long long l2pol = l2_evict_first_policy(p.l2_stream != 0);
    const bool tma = p.use_tma != 0;
    uint64_t* tbar = &s_tile_bar[wid];
    if (tma) {
        // TMA: ONE bulk tensor copy brings the tile's objectness logits (128 cells x 1 plane) and one its class logits
        // (128 cells x NQ planes) into the warp's stash -- no per-lane addresses, no per-plane copy instructions; cells
        // beyond the plane and planes beyond the last anchor's arrive as zeros.  The class planes are read once (L2
        // evict-first); the objectness plane is read by the loss too (normal policy).
        if (lane == 0) {
            mbar_init(tbar, 1);
            mbar_fence_init();
            mbar_expect_tx(tbar, (uint32_t)((NQ + 1) * WT_CELLS * sizeof(float)));
            tma_load_3d(&sm.o[0], &p.tma_obj[L], tbar, c0, a, b);
            tma_load_3d_hint(&sm.q[0][0], &p.tma_cls[L], tbar, c0, a * C, b, l2pol);
        }
        __syncwarp();
    } else {
        // Asynchronous 16-byte copies global -> the warp's shared stash (LDGSTS): all of the tile's bytes are in flight
        // at once without holding registers.  The obj logits land in sm.o and are replaced by their sigmoid below.
        const int64_t cls_step = hd.cls_sp;
        if (in_range) cp_async16(&sm.o[cell0], obj_base_of() + cell0);  // (normal policy: the loss reads the obj planes too)
        else *reinterpret_cast<float4*>(&sm.o[cell0]) = zero4;
        const int nld = in_range ? min(C, CGROUP) : 0;
        const float* cp = cls_base_of() + cell0;
        if (nld == NQ) {  // the usual case: every class slot of the group is real -- no per-copy predicate
#pragma unroll
            for (int u = 0; u < NQ; ++u) cp_async16_s_ef(sq0 + u * (WT_CELLS * 4), cp + (int64_t)u * cls_step, l2pol);
        } else {
#pragma unroll
            for (int u = 0; u < NQ; ++u) {
                if (u < nld) cp_async16_s_ef(sq0 + u * (WT_CELLS * 4), cp, l2pol);
                else *reinterpret_cast<float4*>(&sm.q[u][cell0]) = zero4;
                cp += cls_step;
            }
        }
        cp_async_commit();
    }
    // the image's floor, if its sample CTA is through already (the read overlaps the copies above)
    int floor_bin = HB - 1;  // HB - 1: list everything
    if (!sample) {
        const int w = *(const volatile int*)(p.floor_word + b);
        if (w > 0) floor_bin = w - 1;
    }
    bool full = floor_bin >= HB - 1;
    // per cell: objectness, row filter (nms.py:33), the class-logit threshold of the candidate test and the one of
    // the floor test (branch-free)
    float tau[4], tauf[4], oa[4];  // oa: objectness of an active cell, -1 otherwise
    unsigned force = 0u, forcef = 0u;  // cells whose passing pairs all take the exact test
    // floor test: score = sigmoid(q) * o reaches the floor's lower edge fl only if o >= fl; same construction as the
    // candidate test with fl in place of thr
    auto floor_cells = [&]() {
        const float fl = coarse_lower_edge(floor_bin);
        forcef = 0u;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const float rf = fl * rcp_approx(oa[v]);
            const bool richc = oa[v] >= fl, fastf = richc && rf < 0.9f;
            const float tf = 0.693147182464599609375f * lg2_approx(rf * rcp_approx(1.0f - rf));
            tauf[v] = richc ? (fastf ? tf : 2.19f) : 3.0e38f;
            if (richc && !fastf) forcef |= 1u << v;
        }
    };
    unsigned tile_phase = 0;
    if (tma) {
        mbar_wait(tbar, tile_phase);  // (acquire: the async-proxy writes of the copy are visible to every lane)
        tile_phase ^= 1u;
    } else {
        cp_async_wait_all();  // every lane reads back only what it copied itself until the __syncwarp further down
    }
    {
        const float4 xo = *reinterpret_cast<const float4*>(&sm.o[cell0]);
        const float x[4] = {xo.x, xo.y, xo.z, xo.w};
        float ov4[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const float ov = sigmoid_eval(x[v]);
            ov4[v] = ov;
            const bool act = in_range && ov > thr;
            oa[v] = act ? ov : -1.0f;
            const float r = thr * rcp_approx(ov);  // thr / o, in (0, 1) for an active cell
            const bool fast = act && thr > 0.0f && r < 0.9f;
            const float tv = 0.693147182464599609375f * lg2_approx(r * rcp_approx(1.0f - r));  // logit(thr / o)
            // r >= 0.9 (o barely above thr; logit ill-conditioned): the true threshold is >= logit(0.9) = 2.197, so
            // 2.19 is a safe lower bound -- the (very few) pairs above it are re-decided exactly.  thr <= 0: all are.
            const float slow_tau = thr > 0.0f ? 2.19f : -3.0e38f;
            tau[v] = act ? (fast ? tv : slow_tau) : 3.0e38f;  // inactive cell: nothing passes
            tauf[v] = 3.0e38f;
            if (act && !fast) force |= 1u << v;
        }
        *reinterpret_cast<float4*>(&sm.o[cell0]) = make_float4(ov4[0], ov4[1], ov4[2], ov4[3]);
    }
    if (!full) floor_cells();
    float* sc_img = p.cand_score + (int64_t)b * p.cand_stride;
    uint32_t* ky_img = p.cand_key + (int64_t)b * p.cand_stride;
    int* hist_img = p.hist + (int64_t)b * HB;
    for (int grp = 0; grp < p.ngroups; ++grp) {
        const int j0 = grp * CGROUP, jn = min(C - j0, CGROUP);
        if (grp > 0) {
            __syncwarp();  // the previous group's readers of the stash are done
            if (tma) {
                if (lane == 0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy reads before the async-proxy overwrite
                    mbar_expect_tx(tbar, (uint32_t)(NQ * WT_CELLS * sizeof(float)));
                    tma_load_3d_hint(&sm.q[0][0], &p.tma_cls[L], tbar, c0, a * C + j0, b, l2pol);
                }
                mbar_wait(tbar, tile_phase);
                tile_phase ^= 1u;
            } else {
                const int64_t cls_step = hd.cls_sp;
                const float* cp = cls_base_of() + (int64_t)j0 * cls_step + cell0;
                const int nld = in_range ? jn : 0;
#pragma unroll
                for (int u = 0; u < NQ; ++u) {
                    if (u < nld) cp_async16_s_ef(sq0 + u * (WT_CELLS * 4), cp, l2pol);
                    else *reinterpret_cast<float4*>(&sm.q[u][cell0]) = zero4;
                    cp += cls_step;
                }
                cp_async_commit();
                cp_async_wait_all();
            }
        }
        // pair (u, v): the sign bit of (tau[v] - q[u][v]) is funnel-shifted into one mask per cell (four
        // independent dependency chains); class u ends at bit NQ-1-u of the chain.  Same for the floor test.
        unsigned mv[4] = {0u, 0u, 0u, 0u}, mf[4] = {0u, 0u, 0u, 0u};
        float amin = 3.0e38f, aminf = 3.0e38f;
        if (full) {
#pragma unroll
            for (int u = 0; u < NQ; ++u) {
                const float4 q4 = *reinterpret_cast<const float4*>(&sm.q[u][cell0]);
                const float qq[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const float d = tau[v] - qq[v];
                    mv[v] = __funnelshift_l(__float_as_uint(d), mv[v], 1);
                    amin = fminf(amin, fabsf(d));
                }
            }
        } else {
#pragma unroll
            for (int u = 0; u < NQ; ++u) {
                const float4 q4 = *reinterpret_cast<const float4*>(&sm.q[u][cell0]);
                const float qq[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const float d = tau[v] - qq[v], df = tauf[v] - qq[v];
                    mv[v] = __funnelshift_l(__float_as_uint(d), mv[v], 1);
                    mf[v] = __funnelshift_l(__float_as_uint(df), mf[v], 1);
                    amin = fminf(amin, fabsf(d));
                    aminf = fminf(aminf, fabsf(df));
                }
            }
        }
        unsigned cm[4], hm[4];  // bit u: pair (class j0+u, cell cell0+v) is a candidate / reaches the floor
        unsigned fany = 0u;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            cm[v] = __brev(mv[v]) >> (32 - NQ);
            hm[v] = __brev(mf[v]) >> (32 - NQ);
            if ((force >> v) & 1u) fany |= cm[v];
        }
        if (amin <= TAU_WINDOW || fany) {  // rare: re-decide the borderline pairs with the decode kernel's arithmetic
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                unsigned border = ((force >> v) & 1u) ? cm[v] : 0u;  // forced cell: every pair above the conservative bound
#pragma unroll
                for (int u = 0; u < NQ; ++u)
                    if (fabsf(tau[v] - sm.q[u][cell0 + v]) <= TAU_WINDOW) border |= 1u << u;  // (own stash: no sync needed)
                if (border) cm[v] = fix_borderline<NQ>(cm[v], border, sm.q, cell0 + v, oa[v], thr, HB - 1);
            }
        }
        // classes beyond the group's end and the `classes` filter (nms.py:54-55)
        const unsigned am = (p.cfg.class_mask[j0 >> 5] >> (j0 & 31)) & ((1u << jn) - 1u) & CELL_MASK;
        int mine = 0;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            cm[v] &= am;
            mine += __popc(cm[v]);
        }
        const int n_pass = warp_sum_i(mine);
        if (in_range)  // the candidate set of the tile: one class mask per cell
            *reinterpret_cast<uint2*>(p.mask + (int64_t)b * p.mask_stride + (int64_t)grp * p.rows + row0 + cell0) =
                make_uint2(cm[0] | (cm[1] << 16), cm[2] | (cm[3] << 16));
        if (full && !sample) {
            // second chance: the floor may have been published while this tile was streaming; a first-wave tile
            // gives the sample CTA a few microseconds (bounded: nothing here depends on another CTA's progress)
            int w = 0;
            if (lane == 0) {
                const volatile int* fw = p.floor_word + b;
                w = *fw;
                for (int tries = 0; w == 0 && tries < p.floor_polls; ++tries) {
                    __nanosleep(100);
                    w = *fw;
                }
            }
            w = __shfl_sync(KODB_FULL_MASK, w, 0);
            if (w > 0 && w - 1 < HB - 1) {
                floor_bin = w - 1;
                full = false;
                floor_cells();
#pragma unroll
                for (int u = 0; u < NQ; ++u) {
                    const float4 q4 = *reinterpret_cast<const float4*>(&sm.q[u][cell0]);
                    const float qq[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        const float df = tauf[v] - qq[v];
                        mf[v] = __funnelshift_l(__float_as_uint(df), mf[v], 1);
                        aminf = fminf(aminf, fabsf(df));
                    }
                }
#pragma unroll
                for (int v = 0; v < 4; ++v) hm[v] = __brev(mf[v]) >> (32 - NQ);
            }
        }
        if (!full) {
            unsigned fanyf = 0u;
#pragma unroll
            for (int v = 0; v < 4; ++v)
                if ((forcef >> v) & 1u) fanyf |= hm[v];
            if (aminf <= TAU_WINDOW || fanyf) {
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    unsigned borderf = ((forcef >> v) & 1u) ? hm[v] : 0u;
#pragma unroll
                    for (int u = 0; u < NQ; ++u)
                        if (fabsf(tauf[v] - sm.q[u][cell0 + v]) <= TAU_WINDOW) borderf |= 1u << u;
                    if (borderf) hm[v] = fix_borderline<NQ>(hm[v], borderf, sm.q, cell0 + v, oa[v], thr, floor_bin);
                }
            }
#pragma unroll
            for (int v = 0; v < 4; ++v) hm[v] &= cm[v];
        }
        if (lane == 0) {
            const int64_t chunk = (int64_t)b * p.chunks_per_image + (int64_t)t * p.ngroups + grp;
            p.chunk_count[chunk] = n_pass;
            p.chunk_start[chunk] = full ? HB - 1 : floor_bin;  // the floor this chunk was listed with (HB - 1: everything)
        }
        // pairs to list; the space is claimed now and the reply of the atomic is only needed after the id list is built
        const int hot = __popc(hm[0]) + __popc(hm[1]) + __popc(hm[2]) + __popc(hm[3]);
        const int n = full ? n_pass : warp_sum_i(hot);
        int base = 0;
        if (lane == 0 && n) base = atomicAdd(&p.img_count[b], n);
        if (full) {
            // all candidates of the tile; fully predicated (the lanes' counts differ by 2x and more)
            const int slot = warp_excl_scan_i(mine, lane);
            unsigned short* lp = sm.list + slot;
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const int idb = (cell0 + v) << 4;
#pragma unroll
                for (int u = 0; u < NQ; ++u)
                    if (cm[v] & (1u << u)) *lp++ = (unsigned short)(idb | u);
            }
        } else {
            // the few pairs that reach the floor: ONE loop over the lane's four cell masks (most lanes have none)
            int slot = warp_excl_scan_i(hot, lane);
            // (two 32-bit halves: 32-bit find-first-set / clear-lowest are one instruction each, the 64-bit forms four)
            unsigned h01 = hm[0] | (hm[1] << 16), h23 = hm[2] | (hm[3] << 16);
            while (h01) {
                const int bi = __ffs((int)h01) - 1;
                h01 &= h01 - 1;
                sm.list[slot++] = (unsigned short)(((cell0 + (bi >> 4)) << 4) | (bi & 15));
            }
            while (h23) {
                const int bi = __ffs((int)h23) - 1;
                h23 &= h23 - 1;
                sm.list[slot++] = (unsigned short)(((cell0 + 2 + (bi >> 4)) << 4) | (bi & 15));
            }
        }
        __syncwarp();  // list, q stash and o are visible to the whole warp
        if (sample) {
            // sample tile: the histograms first, so that the floor is out as early as possible
#pragma unroll 1
            for (int i0 = 0; i0 < n; i0 += 32) {
                const int i = i0 + lane;
                if (i < n) {
                    const int pid = sm.list[i];
                    const int bin = coarse_bin(sigmoid_eval(sm.q[pid & 15][pid >> 4]) * sm.o[pid >> 4]);
                    atomicAdd(hist_img + bin, 1);
                    atomicAdd(&s_shist[bin], 1);
                }
            }
            if (grp == p.ngroups - 1) {
                __syncthreads();  // all sample tiles of the image are in the shared histogram
                if (wid == 0) {
                    // sample counts scale by T / ns to the whole image; first bin whose cumulative estimate reaches
                    // the target (bins are in descending score order)
                    const long long need = ((long long)p.hot_target * ns + T - 1) / T;
                    int part = 0;
                    int hv[HB / 32];
#pragma unroll
                    for (int i = 0; i < HB / 32; ++i) {
                        hv[i] = s_shist[lane * (HB / 32) + i];
                        part += hv[i];
                    }
                    int cum = warp_excl_scan_i(part, lane);
                    int found = HB;
#pragma unroll
                    for (int i = 0; i < HB / 32; ++i) {
                        cum += hv[i];
                        if (found == HB && cum >= need) found = lane * (HB / 32) + i;
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) found = min(found, __shfl_xor_sync(KODB_FULL_MASK, found, o));
                    const int fb = found >= HB - 1 ? HB - 1 : found;
                    if (lane == 0) atomicExch(&p.floor_word[b], fb + 1);
                }
            }
        }
        base = __shfl_sync(KODB_FULL_MASK, base, 0);
        float* scp = sc_img + base;
        uint32_t* kyp = ky_img + base;
        const uint32_t key0 = (uint32_t)row0 * (uint32_t)C + (uint32_t)j0;
#pragma unroll 1
        for (int i0 = 0; i0 < n; i0 += 32) {
            const int i = i0 + lane;
            if (i < n) {
                const int pid = sm.list[i];
                const int cl = pid >> 4, u = pid & 15;
                const float conf = sigmoid_eval(sm.q[u][cl]) * sm.o[cl];  // the decode kernel's arithmetic
                scp[i] = conf;
                kyp[i] = key0 + (uint32_t)cl * (uint32_t)C + (uint32_t)u;
                if (!sample) atomicAdd(hist_img + coarse_bin(conf), 1);
            }
        }
    }
}

template <int NQ>
static int launch_cand_warp(const EvalParams& p, dim3 grid, cudaStream_t st) {
    const size_t smem = sizeof(CandWarpSmem<NQ>) * CW_WARPS + 128;  // (+ alignment of the stashes)
    static bool attr_set[64] = {false};
    std::lock_guard<std::mutex> host_lock(g_eval_host_mu);
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        if (cudaFuncSetAttribute(cand_floor_kernel<NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return KODB_ERR_LAUNCH;
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    cand_floor_kernel<NQ><<<grid, CW_WARPS * 32, smem, st>>>(p);
    KODB_CHECK_LAUNCH();
    return KODB_OK;
}

// =========================================================================================================
// 2. per-image select + NMS (+ matching)
// =========================================================================================================
constexpr int CLS_CAP = 256;             // class-bucketed NMS: max candidates of ONE class in a batch
constexpr int CLS_WORDS = CLS_CAP / 32;  // suppression-mask words per candidate

struct NmsSmem {
    int hist[NB];  // fine histogram of the radix refinement; aliased by the per-(warp, class) counts of the NMS phase
    int chist[HB];
    unsigned long long sort_ck[K_CAP];
    float bx1[K_CAP], by1[K_CAP], bx2[K_CAP], by2[K_CAP], barea[K_CAP];      // batch, score order
    float4 cbox[K_CAP];  // batch, class-major order (x1, y1, x2, y2 with the class offset)
    float carea[K_CAP];
    unsigned pm[K_CAP * CLS_WORDS];  // pm[pos][J]: earlier same-class candidates 32J..32J+31 that suppress `pos`
    unsigned char pre[K_CAP], keepf[K_CAP];
    int ccount[KODB_MAX_CLASSES], cbase[KODB_MAX_CLASSES], tstart[KODB_MAX_CLASSES + 1];
    float kx1[MAX_DET_CAP], ky1[MAX_DET_CAP], kx2[MAX_DET_CAP], ky2[MAX_DET_CAP], karea[MAX_DET_CAP];
    int kcls[MAX_DET_CAP];
    int kc_cnt[KODB_MAX_CLASSES], kc_base[KODB_MAX_CLASSES];  // survivors bucketed by class (rebuilt per batch) ...
    unsigned short kc_idx[MAX_DET_CAP];                        // ... so a candidate is only tested against its own class
    unsigned long long kept_ck[MAX_DET_CAP];
    long long kept_idx[MAX_DET_CAP];
    int scan_tmp[33];
    int alive[NMS_STEP];
    unsigned mask[NMS_STEP][NMS_STEP / 32];
    int bstart[HB], bfill[HB];  // batch-relative start / fill count of every coarse bin of the current batch
    int sel_bin, sel_p, sel_prev;
    int cnt, nk, cls_over;
};
constexpr int RANK_CAP = 192;  // bins up to this population are ordered by counting; a heavier bin falls back to the bitonic sort
static_assert(NB == 4 * NMS_THREADS, "the radix refinement reads four fine bins per thread");
static_assert(GT_CAP <= NMS_THREADS, "the select kernel stages one ground-truth box per thread");
static_assert(NB >= (NMS_THREADS / 32) * KODB_MAX_CLASSES, "per-(warp, class) counts alias the fine histogram");

__device__ __forceinline__ bool suppresses(float ax1, float ay1, float ax2, float ay2, float aarea, float bx1, float by1,
                                           float bx2, float by2, float barea, double thr) {
    // torchvision nms_kernel.cpp: fp32 throughout, ratio compared with the double threshold
    const float xx1 = fmaxf(ax1, bx1), yy1 = fmaxf(ay1, by1), xx2 = fminf(ax2, bx2), yy2 = fminf(ay2, by2);
    const float w = fmaxf(0.0f, xx2 - xx1), h = fmaxf(0.0f, yy2 - yy1);
    // disjoint boxes (every cross-class pair, thanks to the class offset): inter == 0 -> ovr is 0 (or NaN),
    // never > thr for thr >= 0; skip the division
    if (thr >= 0.0 && (w == 0.0f || h == 0.0f)) return false;
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
    unsigned long long gmask[MAX_DET_CAP];  // ground truths (< 64) of the detection's class overlapping it at all
};
// ground truth of the image, staged once at kernel start (its own region: it must survive the NMS phase)
struct GtSmem {
    double gx[GT_CAP], gy[GT_CAP], gw[GT_CAP], gh[GT_CAP];  // [x, y, w, h] in fp64, like pycocotools
    int glabel[GT_CAP];
    unsigned char taken[KODB_NUM_IOU_THRS][GT_CAP];
};

constexpr size_t kNmsRegion =
    ((sizeof(NmsSmem) > sizeof(MatchSmem) ? sizeof(NmsSmem) : sizeof(MatchSmem)) + 15) & ~(size_t)15;
constexpr size_t kMatchRegion = (sizeof(MatchSmem) + 15) & ~(size_t)15;
static_assert(kNmsRegion + sizeof(GtSmem) <= 113 * 1024, "two select CTAs per SM");

__device__ __forceinline__ void load_gt_xyxy(const void* boxes, int dtype, int g, double& l, double& t, double& r, double& b) {
    if (dtype == KODB_BOXES_F64) {
        const double* q = reinterpret_cast<const double*>(boxes) + 4 * (int64_t)g;
        l = q[0]; t = q[1]; r = q[2]; b = q[3];
    } else {
        const float* q = reinterpret_cast<const float*>(boxes) + 4 * (int64_t)g;
        l = q[0]; t = q[1]; r = q[2]; b = q[3];
    }
}

__device__ __forceinline__ void stage_gt(GtSmem& gs, int C, const void* gt_boxes, const int64_t* gt_labels, int g0, int ng,
                                         int gt_dtype, const kodb_match_out& mo) {
    for (int i = threadIdx.x; i < KODB_NUM_IOU_THRS * ng; i += blockDim.x) gs.taken[i / ng][i % ng] = 0;
    for (int g = threadIdx.x; g < ng; g += blockDim.x) {
        const long long lab = gt_labels[g0 + g];
        if (lab >= 0 && lab < C) atomicAdd(reinterpret_cast<unsigned long long*>(mo.gt_count) + lab, 1ull);
        double l, t, r, bt;
        load_gt_xyxy(gt_boxes, gt_dtype, g0 + g, l, t, r, bt);
        gs.gx[g] = l; gs.gy[g] = t; gs.gw[g] = r - l; gs.gh[g] = bt - t;  // [l, t, r-l, b-t]
        gs.glabel[g] = (int)lab;
    }
}

// Ground truth of one image as the matching stage sees it: in shared memory (up to GT_CAP boxes, the fused path) or in a
// global scratch (crowded images, match_crowded_kernel) -- same code, two storages.
struct GtInSmem {
    GtSmem& s;
    __device__ __forceinline__ double gx(int g) const { return s.gx[g]; }
    __device__ __forceinline__ double gy(int g) const { return s.gy[g]; }
    __device__ __forceinline__ double gw(int g) const { return s.gw[g]; }
    __device__ __forceinline__ double gh(int g) const { return s.gh[g]; }
    __device__ __forceinline__ int glabel(int g) const { return s.glabel[g]; }
    __device__ __forceinline__ unsigned char& taken(int t, int g) const { return s.taken[t][g]; }
};
struct GtInGlobal {  // [4][cap] doubles, [cap] labels, [KODB_NUM_IOU_THRS][cap] flags
    double* xywh;
    int* label;
    unsigned char* flags;
    int cap;
    __device__ __forceinline__ double gx(int g) const { return xywh[g]; }
    __device__ __forceinline__ double gy(int g) const { return xywh[cap + g]; }
    __device__ __forceinline__ double gw(int g) const { return xywh[2 * cap + g]; }
    __device__ __forceinline__ double gh(int g) const { return xywh[3 * cap + g]; }
    __device__ __forceinline__ int glabel(int g) const { return label[g]; }
    __device__ __forceinline__ unsigned char& taken(int t, int g) const { return flags[(size_t)t * cap + g]; }
};

// maskApi.c::bbIou on [x,y,w,h] in fp64
template <class GT>
__device__ __forceinline__ double bb_iou(double dl, double dt, double dw, double dh, double da, const GT& gs, int g) {
    const double gl = gs.gx(g), gt_ = gs.gy(g), gw = gs.gw(g), gh = gs.gh(g);
    const double iw = fmin(dl + dw, gl + gw) - fmax(dl, gl);
    const double ih = fmin(dt + dh, gt_ + gh) - fmax(dt, gt_);
    if (iw > 0 && ih > 0) {
        const double i = iw * ih;
        return i / (da + gw * gh - i);
    }
    return 0.0;
}

// Greedy COCO matching of the kept detections of one image (block-wide).  `gs` must have been staged.
template <class GT>
__device__ void match_image(MatchSmem& ms, const GT gs, int nk, int C, int ng, const kodb_match_out& mo, int b, int max_det) {
    const int tid = threadIdx.x;
    double thr_min = 1.0;
#pragma unroll
    for (int t = 0; t < KODB_NUM_IOU_THRS; ++t) thr_min = fmin(thr_min, fmin(mo.iou_thrs[t], 1.0 - 1e-10));
    // phase A (parallel over detections): which ground truths of its class does a detection overlap at all?
    // Most detections overlap none and are false positives at every threshold; they drop out here.
    const bool small = ng <= 64;
    for (int d = tid; d < nk; d += blockDim.x) {
        ms.tp[d] = 0;
        unsigned long long mk = 0ull;
        if (small) {
            const int c = ms.cls[d];
            const double dl = ms.x1[d], dt = ms.y1[d];
            const double dw = (double)ms.x2[d] - dl, dh = (double)ms.y2[d] - dt;
            const double da = dw * dh;
            for (int g = 0; g < ng; ++g)
                if (gs.glabel(g) == c && bb_iou(dl, dt, dw, dh, da, gs, g) >= thr_min) mk |= 1ull << g;
        }
        ms.gmask[d] = mk;
    }
    __syncthreads();
    // phase B: one WARP per class (no divergence between classes): lane = (GT slice 0..7) x (IoU threshold 0..3).
    // The class's detections that overlap something are visited sequentially in score order (the greedy
    // dependency); the 8 slices scan the free ground truths in parallel and a shuffle reduction picks the
    // pycocotools winner: largest IoU >= thr, ties -> the LATER ground truth.
    const int lane = tid & 31, wid = tid >> 5, nwarps = blockDim.x >> 5;
    const int t = lane & 3, sub = lane >> 2;
    const double thr = fmin(mo.iou_thrs[t], 1.0 - 1e-10);
    for (int c = wid; c < C; c += nwarps) {
        bool has_gt = false;
        for (int g = lane; g < ng; g += 32) has_gt |= (gs.glabel(g) == c);
        if (!__any_sync(KODB_FULL_MASK, has_gt)) continue;  // nothing to match: all detections of c are FP
        for (int d0 = 0; d0 < nk; d0 += 32) {
            const int dd = d0 + lane;
            const bool cand = dd < nk && ms.cls[dd] == c && (!small || ms.gmask[dd] != 0ull);
            unsigned dm = __ballot_sync(KODB_FULL_MASK, cand);
            while (dm) {
                const int d = d0 + __ffs(dm) - 1;
                dm &= dm - 1;
                const double dl = ms.x1[d], dt = ms.y1[d];
                const double dw = (double)ms.x2[d] - dl, dh = (double)ms.y2[d] - dt;
                const double da = dw * dh;
                const unsigned long long mk = small ? ms.gmask[d] : ~0ull;
                double best = -1.0;
                int m = -1;
                for (int g = sub; g < ng; g += 8) {
                    if (small ? !((mk >> g) & 1ull) : (gs.glabel(g) != c)) continue;
                    if (gs.taken(t, g)) continue;
                    const double iou = bb_iou(dl, dt, dw, dh, da, gs, g);
                    if (iou < thr || iou < best) continue;
                    best = iou;  // ascending g inside a slice: a later equal IoU replaces the earlier one
                    m = g;
                }
                if (!__any_sync(KODB_FULL_MASK, m >= 0)) continue;
#pragma unroll
                for (int off = 4; off < 32; off <<= 1) {
                    const double ob = __shfl_xor_sync(KODB_FULL_MASK, best, off);
                    const int om = __shfl_xor_sync(KODB_FULL_MASK, m, off);
                    if (ob > best || (ob == best && om > m)) { best = ob; m = om; }
                }
                if (m >= 0 && sub == 0) {
                    gs.taken(t, m) = 1;
                    atomicOr(&ms.tp[d], 1 << t);
                }
                __syncwarp();
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

// Floor mode, rare: the listed candidates ran dry before max_det survivors.  Lists every candidate of the image
// that the candidate kernel left out (pairs of tiles not listed in full whose score lies in a bin below the floor), straight
// from the class masks and the head logits, appends them to the image's list and completes the histogram.
// Four warps, one tile at a time each; scratch on top of the batch arrays (free between batches).  Block-wide.
__device__ unsigned int g_completions = 0u;  // debug statistic: completion passes since the last query

__device__ int complete_list(const EvalParams& p, NmsSmem& s, int b, int n0) {
    // every warp of the CTA takes half tiles (64 cells, two per lane); scratch = the batch arrays bx1 .. pm, which are
    // consecutive members of NmsSmem and free between batches
    constexpr int HALF = WT_CELLS / 2;
    constexpr int CL_WARPS = NMS_THREADS / 32;
    constexpr size_t PER_WARP = HALF * CGROUP * sizeof(unsigned short) + HALF * sizeof(float);
    static_assert(offsetof(NmsSmem, pm) + sizeof(s.pm) - offsetof(NmsSmem, bx1) >= CL_WARPS * PER_WARP, "scratch of the completion pass");
    static_assert(offsetof(NmsSmem, bx1) < offsetof(NmsSmem, pm) && offsetof(NmsSmem, bx1) % 16 == 0, "batch arrays are laid out in order");
    if (threadIdx.x == 0) atomicAdd(&g_completions, 1u);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int C = p.C;
    if (tid == 0) s.cnt = 0;
    __syncthreads();
    {
        unsigned char* scratch = reinterpret_cast<unsigned char*>(s.bx1) + (size_t)wid * PER_WARP;
        float* so = reinterpret_cast<float*>(scratch);
        unsigned short* list = reinterpret_cast<unsigned short*>(scratch + HALF * sizeof(float));
        float* sc_img = p.cand_score + (int64_t)b * p.cand_stride + n0;
        uint32_t* ky_img = p.cand_key + (int64_t)b * p.cand_stride + n0;
        const unsigned short* mask_img = p.mask + (int64_t)b * p.mask_stride;
        for (int unit = wid; unit < 2 * p.tiles_per_image; unit += CL_WARPS) {
            const int t = unit >> 1, h0 = (unit & 1) * HALF;
            const int* flags = p.chunk_start + (int64_t)b * p.chunks_per_image + (int64_t)t * p.ngroups;
            int L, a, c0, ncell;
            planar_tile(p.g, t, L, a, c0, ncell);
            if (h0 >= ncell) continue;
            bool any_cold_chunk = false;
            for (int g = 0; g < p.ngroups; ++g) any_cold_chunk |= flags[g] < HB - 1;
            if (!any_cold_chunk) continue;  // every chunk of the tile is listed in full already
            const kodb_head_level& hd = p.g.head[L];
            const int row0 = p.g.row_base[L] + a * (p.g.lv[L].H * p.g.lv[L].W) + c0 + h0;  // first row of the half tile
            const int nh = min(HALF, ncell - h0);
            const float* obj_base = hd.obj + (int64_t)b * hd.obj_sb + (int64_t)a * hd.obj_sa + c0 + h0;
            const float* cls_base = hd.cls + (int64_t)b * hd.cls_sb + (int64_t)a * hd.cls_sa + c0 + h0;
            __syncwarp();  // the previous unit's readers of the scratch are done
            for (int i = lane; i < nh; i += 32) so[i] = sigmoid_eval(obj_base[i]);
            for (int g = 0; g < p.ngroups; ++g) {
                const int floor_bin = flags[g];  // the floor the candidate kernel listed this chunk with
                if (floor_bin >= HB - 1) continue;  // listed in full already
                const int j0 = g * CGROUP;
                const unsigned short* mk = mask_img + (int64_t)g * p.rows + row0;
                int mine = 0;
                unsigned m2[2];
#pragma unroll
                for (int v = 0; v < 2; ++v) {
                    const int cl = 2 * lane + v;
                    m2[v] = cl < nh ? mk[cl] : 0u;
                    mine += __popc(m2[v]);
                }
                int slot = warp_excl_scan_i(mine, lane);
                const int n = __shfl_sync(KODB_FULL_MASK, slot + mine, 31);
                __syncwarp();  // the previous group's readers of the list are done
#pragma unroll
                for (int v = 0; v < 2; ++v) {
                    unsigned m = m2[v];
                    while (m) {
                        const int u = __ffs(m) - 1;
                        m &= m - 1;
                        list[slot++] = (unsigned short)(((2 * lane + v) << 4) | u);
                    }
                }
                __syncwarp();
                for (int i0 = 0; i0 < n; i0 += 32) {
                    const int i = i0 + lane;
                    bool cold = false;
                    float conf = 0.0f;
                    uint32_t k32 = 0u;
                    int bin = 0;
                    if (i < n) {
                        const int pid = list[i];
                        const int cl = pid >> 4, u = pid & 15;
                        conf = sigmoid_eval(cls_base[(int64_t)(j0 + u) * hd.cls_sp + cl]) * so[cl];  // the decode kernel's arithmetic
                        bin = coarse_bin(conf);
                        cold = bin > floor_bin;
                        k32 = (uint32_t)(row0 + cl) * (uint32_t)C + (uint32_t)(j0 + u);
                    }
                    const unsigned bal = __ballot_sync(KODB_FULL_MASK, cold);
                    if (bal) {
                        int base = 0;
                        if (lane == 0) base = atomicAdd(&s.cnt, __popc(bal));
                        base = __shfl_sync(KODB_FULL_MASK, base, 0);
                        if (cold) {
                            const int pos = base + __popc(bal & ((1u << lane) - 1u));
                            sc_img[pos] = conf;
                            ky_img[pos] = k32;
                            atomicAdd(&s.chist[bin], 1);
                        }
                    }
                }
            }
        }
    }
    __threadfence_block();
    __syncthreads();
    return n0 + s.cnt;
}

__global__ void __launch_bounds__(NMS_THREADS, 2) select_nms_kernel(const __grid_constant__ EvalParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    NmsSmem& s = *reinterpret_cast<NmsSmem*>(smem_raw);
    MatchSmem& ms = *reinterpret_cast<MatchSmem*>(smem_raw);  // aliases NmsSmem: used after NMS is finished
    GtSmem& gs = *reinterpret_cast<GtSmem*>(smem_raw + kNmsRegion);

    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int C = p.C;
    const bool floor_mode = p.floor_mode != 0;
    const int* ccount = p.chunk_count + (int64_t)b * p.chunks_per_image;
    const int* cstart = p.chunk_start + (int64_t)b * p.chunks_per_image;
    int* cprefix = p.chunk_prefix + (int64_t)b * p.chunks_per_image;  // exclusive prefix of the chunk counts (written below)
    // every independent global load of the prologue is issued before the first use (one L2 round trip, not five)
    int n = p.img_count[b];  // listed candidates (floor mode: a subset of all candidates until the completion pass)
    const int hist_v = tid < HB ? p.hist[(int64_t)b * HB + tid] : 0;
    const int cc0 = tid < p.chunks_per_image ? ccount[tid] : 0;
    int cfloor = HB - 1;  // floor mode: the lowest-numbered floor bin any chunk of the image was listed with
    if (floor_mode)
        for (int t = tid; t < p.chunks_per_image; t += NMS_THREADS) cfloor = min(cfloor, cstart[t]);
    int g0 = 0, gt_n = 0;
    if (p.do_match) {
        g0 = p.gt_offsets[b];
        gt_n = p.gt_offsets[b + 1] - g0;
    }
    const float* score = p.cand_score + (int64_t)b * p.cand_stride;
    const uint32_t* key = p.cand_key + (int64_t)b * p.cand_stride;
    const int max_det = p.cfg.max_det;
    const unsigned short* mask_img = p.mask + (int64_t)b * p.mask_stride;
    const double iou_thr = p.cfg.iou_thres;
    const float max_wh = p.cfg.max_wh;

    int trace_n = 0;
    KODB_TRACE(0);
    if (p.trace && tid == 0) {
        long long gt0;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt0));
        p.trace[(int64_t)b * 128 + 1] = gt0;
    }
    if (tid == 0) { s.nk = 0; }
    // ground truth of the image: loaded now (one box per thread), stored to shared memory after the scans below
    long long gt_lab = -1;
    double gt_l = 0.0, gt_t = 0.0, gt_r = 0.0, gt_b = 0.0;
    if (tid < gt_n) {
        gt_lab = p.gt_labels[g0 + tid];
        load_gt_xyxy(p.gt_boxes, p.gt_dtype, g0 + tid, gt_l, gt_t, gt_r, gt_b);
    }
    // Class-bucketed NMS is exact iff boxes of different classes cannot intersect after the class offset: every
    // raw coordinate inside a window no wider than max_wh, class offsets j*max_wh exact in fp32 (rounding of
    // x + offset is monotone, so disjoint stays disjoint), and a threshold >= 0 (zero overlap never suppresses).
    const float win_lo = -0.25f * max_wh, win_hi = 0.75f * max_wh;
    bool cls_static = iou_thr >= 0.0 && max_wh > 0.0f && max_wh < 3.0e38f && (double)win_hi - (double)win_lo <= (double)max_wh;
    const float thr_rd = __double2float_rd(iou_thr);
    const float thr_hi = thr_rd * 1.000001f, thr_lo = thr_rd * 0.999999f;  // guard band of the approximate ratio test
    cls_static = __syncthreads_and(tid >= C || (double)((float)tid * max_wh) == (double)tid * (double)max_wh) && cls_static;
    // coarse score histogram of the image (accumulated by the candidate kernel); reset for the next launch
    if (tid < HB) {
        s.chist[tid] = hist_v;
        p.hist[(int64_t)b * HB + tid] = 0;
    }
    // exclusive prefix of the per-chunk candidate counts (chunk order == key order of the tiles); in floor mode the
    // chunk counts are the only place that knows the image's total number of candidates
    int n_total = n;
    if (floor_mode || n <= p.cfg.max_nms) {
        int carry = 0;
        for (int t0 = 0; t0 < p.chunks_per_image; t0 += NMS_THREADS) {
            const int t = t0 + tid;
            const int v = t0 == 0 ? cc0 : (t < p.chunks_per_image ? ccount[t] : 0);
            int tot;
            const int ex = block_excl_scan(v, s.scan_tmp, &tot);
            if (t < p.chunks_per_image) cprefix[t] = carry + ex;
            carry += tot;
        }
        if (floor_mode) n_total = carry;
    }
    const bool rank_mode = n_total > p.cfg.max_nms;  // nms.py:61-63: candidates re-ordered by confidence, top max_nms kept
    const int limit = rank_mode ? p.cfg.max_nms : n_total;
    // floor mode: bins 0..avail-1 of the histogram are complete (every candidate of those bins is listed)
    // Every chunk records the floor it was listed with (a tile that ran before the floor was published lists
    // everything), so the complete bins follow from the chunks themselves, whatever the timing was.
    int avail = HB;
    if (floor_mode) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cfloor = min(cfloor, __shfl_xor_sync(KODB_FULL_MASK, cfloor, o));
        if (lane == 0) s.scan_tmp[wid] = cfloor;  // (free between the scans above and the cut below)
        __syncthreads();
        cfloor = s.scan_tmp[lane < NMS_THREADS / 32 ? lane : 0];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cfloor = min(cfloor, __shfl_xor_sync(KODB_FULL_MASK, cfloor, o));
        avail = cfloor >= HB - 1 ? HB : cfloor + 1;
        if (tid == 0) p.floor_word[b] = 0;  // reusable without a memset
    }
    if (p.do_match) {  // (GT_CAP == NMS_THREADS: one ground-truth box per thread)
        for (int i = tid; i < KODB_NUM_IOU_THRS * gt_n; i += NMS_THREADS) gs.taken[i / gt_n][i % gt_n] = 0;
        if (tid < gt_n) {
            if (gt_lab >= 0 && gt_lab < C) atomicAdd(reinterpret_cast<unsigned long long*>(p.m.gt_count) + gt_lab, 1ull);
            gs.gx[tid] = gt_l; gs.gy[tid] = gt_t; gs.gw[tid] = gt_r - gt_l; gs.gh[tid] = gt_b - gt_t;  // [l, t, r-l, b-t]
            gs.glabel[tid] = (int)gt_lab;
        }
    }
    __syncthreads();
    KODB_TRACE(1);

    int nk = 0, processed = 0, k_target = K_TARGET0;
    unsigned long long lo = 0ull;  // composite keys below `lo` have been consumed
    int dnext = 0;                 // first coarse bin not consumed yet
    bool in_fine = false;          // bin `dnext` is being consumed in pieces (it overflowed the sort capacity)

    while (nk < max_det && processed < limit) {
        unsigned long long hi_batch = U64MAX;
        int m = 0;
        bool binned = false;
        if (!in_fine) {
            // ------------------------------------------------------------ cut from the coarse histogram
            if (tid == 0) s.sel_bin = -1;
            const int v = (tid < avail && tid >= dnext) ? s.chist[tid] : 0;
            int total;
            const int ex = block_excl_scan(v, s.scan_tmp, &total);  // (barriers inside order sel_bin's reset)
            if (ex < k_target && ex + v >= k_target) { s.sel_bin = tid; s.sel_p = ex + v; s.sel_prev = ex; }
            if (tid < HB) { s.bstart[tid] = ex; s.bfill[tid] = 0; }
            __syncthreads();
            if (total == 0) {
                if (avail == HB) break;  // nothing left
                // the listed candidates ran dry before max_det survivors: list the rest of the image from the masks
                n = complete_list(p, s, b, n);
                avail = HB;
                continue;
            }
            const int d = s.sel_bin, dfirst = dnext;
            if (d < 0) { m = total; hi_batch = coarse_hi(avail - 1); dnext = avail; }     // fewer than k_target left: all
            else if (s.sel_p <= K_CAP) { m = s.sel_p; hi_batch = coarse_hi(d); dnext = d + 1; }
            else if (s.sel_prev > 0) { m = s.sel_prev; hi_batch = coarse_hi(d - 1); dnext = d; }
            else { in_fine = true; dnext = d; }  // bins before d are empty; d alone overflows the sort capacity
            // whole-bin batch: candidates are collected straight into per-bin segments and ordered by counting
            // (the barrier also orders the reads of sel_* above before their next reset)
            const int heavy = __syncthreads_or(tid >= dfirst && tid < dnext && v > RANK_CAP);
            binned = !in_fine && !heavy;
        }
        if (in_fine) {
            binned = false;
            // ------------------------------------------------------------ radix-select inside coarse bin `dnext`
            const unsigned long long fine_hi = coarse_hi(dnext);
            unsigned long long glo = lo, ghi = fine_hi, rlo = lo;
            int shift;
            {
                const unsigned long long span = ghi - glo - 1ull;
                const int nbits = span ? 64 - __clzll((long long)span) : 0;
                shift = nbits > LOG_NB ? nbits - LOG_NB : 0;
            }
            for (;;) {
                for (int i = tid; i < NB; i += NMS_THREADS) s.hist[i] = 0;
                if (tid == 0) s.sel_bin = -1;
                __syncthreads();
                for (int i = tid; i < n; i += NMS_THREADS) {
                    const unsigned long long ck = make_ck(score[i], key[i]);
                    if (ck >= glo && ck < ghi) {
                        const unsigned long long d = (ck - rlo) >> shift;
                        atomicAdd(&s.hist[d > NB - 1 ? NB - 1 : (int)d], 1);
                    }
                }
                __syncthreads();
                const int c0 = s.hist[4 * tid], c1 = s.hist[4 * tid + 1], c2 = s.hist[4 * tid + 2], c3 = s.hist[4 * tid + 3];
                int total;
                const int ex = block_excl_scan(c0 + c1 + c2 + c3, s.scan_tmp, &total);
                const int p0 = ex + c0, p1 = p0 + c1, p2 = p1 + c2, p3 = p2 + c3;
                if (p0 >= k_target && ex < k_target) { s.sel_bin = 4 * tid; s.sel_p = p0; s.sel_prev = ex; }
                else if (p1 >= k_target && p0 < k_target) { s.sel_bin = 4 * tid + 1; s.sel_p = p1; s.sel_prev = p0; }
                else if (p2 >= k_target && p1 < k_target) { s.sel_bin = 4 * tid + 2; s.sel_p = p2; s.sel_prev = p1; }
                else if (p3 >= k_target && p2 < k_target) { s.sel_bin = 4 * tid + 3; s.sel_p = p3; s.sel_prev = p2; }
                __syncthreads();
                const int d = s.sel_bin;
                if (d < 0) { hi_batch = ghi; m = total; break; }  // fewer than k_target left in range: take them all
                const unsigned long long blo = d == 0 ? glo : max(glo, rlo + ((unsigned long long)d << shift));
                const unsigned long long bhi = d == NB - 1 ? ghi : min(ghi, rlo + ((unsigned long long)(d + 1) << shift));
                if (s.sel_p <= K_CAP) { hi_batch = bhi; m = s.sel_p; break; }
                if (s.sel_prev > 0) { hi_batch = blo; m = s.sel_prev; break; }
                // a single fine bin overflows too: refine inside it
                glo = blo; ghi = bhi; rlo = blo;
                const unsigned long long span = ghi - glo - 1ull;
                const int nbits = span ? 64 - __clzll((long long)span) : 0;
                shift = nbits > LOG_NB ? nbits - LOG_NB : 0;
                __syncthreads();
            }
            if (hi_batch == fine_hi) { in_fine = false; dnext += 1; }  // the coarse bin is exhausted by this batch
            if (m == 0) { lo = hi_batch; continue; }
        }
        KODB_TRACE(2);
        // ------------------------------------------------------------------ collect [lo, hi_batch)
        if (tid == 0) s.cnt = 0;
        __syncthreads();
        if ((uint32_t)lo == 0u && (hi_batch == U64MAX || (uint32_t)hi_batch == 0u)) {
            // both bounds are score-bin boundaries: decide on the score alone, read the key of the selected only.
            // float4 loads, four in flight per thread; pushes are warp-aggregated (one shared atomic per warp).
            const uint32_t dlo = (uint32_t)(lo >> 32), dhi = (uint32_t)(hi_batch >> 32);
            const bool open_hi = hi_batch == U64MAX;
            const float4* score4 = reinterpret_cast<const float4*>(score);  // cand_stride is a multiple of 4
            const int n4 = (n + 3) >> 2;
            const float ninf = __int_as_float(0xff800000);
            for (int r0 = 0; r0 < n4; r0 += 4 * NMS_THREADS) {
                float4 sv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = r0 + u * NMS_THREADS + tid;
                    sv[u] = i < n4 ? __ldcg(score4 + i) : make_float4(ninf, ninf, ninf, ninf);
                }
                unsigned take = 0u;  // bit 4u + w
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = 4 * (r0 + u * NMS_THREADS + tid);
                    const float vv[4] = {sv[u].x, sv[u].y, sv[u].z, sv[u].w};
#pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        const uint32_t dk = desc_key(vv[w]);
                        if (i + w < n && dk >= dlo && (open_hi || dk < dhi)) take |= 1u << (4 * u + w);
                    }
                }
                if (binned) {
                    if (take) {  // into the candidate's own bin segment (order inside a segment is arbitrary)
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int i = 4 * (r0 + u * NMS_THREADS + tid);
                            const float vv[4] = {sv[u].x, sv[u].y, sv[u].z, sv[u].w};
#pragma unroll
                            for (int w = 0; w < 4; ++w)
                                if ((take >> (4 * u + w)) & 1u) {
                                    const int d = coarse_bin(vv[w]);
                                    const int pos = s.bstart[d] + atomicAdd(&s.bfill[d], 1);
                                    s.sort_ck[pos] = ((unsigned long long)desc_key(vv[w]) << 32) | __ldcg(key + i + w);
                                }
                        }
                    }
                } else if (__ballot_sync(KODB_FULL_MASK, take != 0u)) {  // one shared atomic per warp and round
                    const int mine = __popc(take);
                    const int ex = warp_excl_scan_i(mine, lane);
                    int base = 0;
                    if (lane == 31) base = atomicAdd(&s.cnt, ex + mine);
                    int pos = __shfl_sync(KODB_FULL_MASK, base, 31) + ex;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = 4 * (r0 + u * NMS_THREADS + tid);
                        const float vv[4] = {sv[u].x, sv[u].y, sv[u].z, sv[u].w};
#pragma unroll
                        for (int w = 0; w < 4; ++w)
                            if ((take >> (4 * u + w)) & 1u)
                                s.sort_ck[pos++] = ((unsigned long long)desc_key(vv[w]) << 32) | __ldcg(key + i + w);
                    }
                }
            }
        } else {
            for (int i = tid; i < n; i += NMS_THREADS) {
                const unsigned long long ck = make_ck(score[i], key[i]);
                if (ck >= lo && ck < hi_batch) s.sort_ck[atomicAdd(&s.cnt, 1)] = ck;
            }
        }
        __syncthreads();
        KODB_TRACE(3);
        if (!binned) m = s.cnt;  // == the histogram count
        // the box logits of the batch have not been touched yet (DRAM): start fetching them while the batch is sorted
        if (tid < m) prefetch_box(p, b, (int)((uint32_t)s.sort_ck[tid] / (uint32_t)C));
        if (binned) {
            // bins are already in place (descending score = ascending bin); order every segment by counting the
            // smaller composite keys inside it (keys are unique)
            unsigned long long ck = 0ull;
            int dst = 0;
            if (tid < m) {
                ck = s.sort_ck[tid];
                const int d = coarse_bin(score_of_key((uint32_t)(ck >> 32)));
                const int s0 = s.bstart[d], s1 = s0 + s.chist[d];
                int r = 0;
                for (int q = s0; q < s1; ++q) r += s.sort_ck[q] < ck;
                dst = s0 + r;
            }
            __syncthreads();
            if (tid < m) s.sort_ck[dst] = ck;
            __syncthreads();
        } else
        // bitonic sort, FOUR elements per thread (element i = 4t + r) on the first npad/4 threads only: distances
        // 1, 2 are exchanges inside the thread, 4..64 warp shuffles, >= 128 go through shared memory behind a named
        // barrier of just the participating warps (3 such stages for 512 elements, instead of 10 block barriers)
        {
            int npad = 128;
            while (npad < m) npad <<= 1;
            const int nthr = npad >> 2;
            if (tid < nthr) {
                unsigned long long v[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) v[r] = 4 * tid + r < m ? s.sort_ck[4 * tid + r] : U64MAX;
                for (int k = 2; k <= npad; k <<= 1) {
                    for (int j = k >> 1; j > 0; j >>= 1) {
                        if (j >= 4) {
                            const int jt = j >> 2;  // partner thread distance
                            unsigned long long o[4];
                            if (jt >= 32) {
#pragma unroll
                                for (int r = 0; r < 4; ++r) s.sort_ck[4 * tid + r] = v[r];
                                asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");
#pragma unroll
                                for (int r = 0; r < 4; ++r) o[r] = s.sort_ck[4 * (tid ^ jt) + r];
                                asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");
                            } else {
#pragma unroll
                                for (int r = 0; r < 4; ++r) o[r] = __shfl_xor_sync(KODB_FULL_MASK, v[r], jt);
                            }
                            const bool lower = (tid & jt) == 0;
#pragma unroll
                            for (int r = 0; r < 4; ++r) {
                                const bool asc = ((4 * tid + r) & k) == 0;
                                if ((o[r] < v[r]) == (lower == asc) && o[r] != v[r]) v[r] = o[r];
                            }
                        } else {
                            // partner inside the thread: r ^ j (compile-time register indices)
                            if (j == 2) {
                                const bool a0 = ((4 * tid) & k) == 0;  // elements 0..3 share bit k for k >= 4; k == 2 never has j == 2
                                if ((v[0] > v[2]) == a0) { const unsigned long long x = v[0]; v[0] = v[2]; v[2] = x; }
                                if ((v[1] > v[3]) == a0) { const unsigned long long x = v[1]; v[1] = v[3]; v[3] = x; }
                            } else {
                                const bool a0 = ((4 * tid) & k) == 0, a2 = ((4 * tid + 2) & k) == 0;
                                if ((v[0] > v[1]) == a0) { const unsigned long long x = v[0]; v[0] = v[1]; v[1] = x; }
                                if ((v[2] > v[3]) == a2) { const unsigned long long x = v[2]; v[2] = v[3]; v[3] = x; }
                            }
                        }
                    }
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) s.sort_ck[4 * tid + r] = v[r];
            }
            __syncthreads();
        }
        KODB_TRACE(4);
        int mm = m;
        if (processed + mm > limit) mm = limit - processed;  // only the top max_nms enter NMS
        // ------------------------------------------------------------------ decode the batch (offset boxes)
        // one candidate per thread (mm <= K_CAP == NMS_THREADS); `bad`: a raw coordinate outside the window in which
        // boxes of different classes provably cannot overlap after the class offset (nms.py:66)
        int my_c = -1;
        bool bad = false;
        if (tid < mm) {
            const uint32_t k32 = (uint32_t)s.sort_ck[tid];
            const int row = (int)(k32 / (uint32_t)C), j = (int)(k32 - (uint32_t)row * (uint32_t)C);
            const float4 bx = fetch_box(p, b, row);
            const float off = (float)j * max_wh;  // nms.py:66
            const float x1 = bx.x + off, y1 = bx.y + off, x2 = bx.z + off, y2 = bx.w + off;
            s.bx1[tid] = x1; s.by1[tid] = y1; s.bx2[tid] = x2; s.by2[tid] = y2;
            s.barea[tid] = (x2 - x1) * (y2 - y1);
            my_c = j;
            bad = !(bx.x >= win_lo && bx.x <= win_hi && bx.y >= win_lo && bx.y <= win_hi && bx.z >= win_lo && bx.z <= win_hi &&
                    bx.w >= win_lo && bx.w <= win_hi);
        }
        if (cls_static) {
            for (int i = tid; i < (NMS_THREADS / 32) * C; i += NMS_THREADS) s.hist[i] = 0;  // per-(warp, class) counts
            if (tid < C) s.kc_cnt[tid] = 0;
        }
        if (__syncthreads_or(bad)) cls_static = false;  // sticky: a kept out-of-window box may reach across classes later
        bool cls_path = cls_static;
        KODB_TRACE(5);
        // ------------------------------------------------------------------ greedy NMS
        // Class-bucketed path.  After the class offset, boxes of different classes never intersect, so greedy NMS
        // factorises per class; only the cut at max_det keeps couples the classes, and it is taken afterwards in
        // score order.  (1) stable bucketing of the batch by class (match.any + per-warp counts); (2) per class,
        // 32x32 tiles of the lower-triangular "j suppresses i" matrix, one tile per warp, as bit masks; (3) one warp
        // per class resolves the greedy recurrence 32 candidates at a time (fixed-point iteration: its only fixed
        // point is the greedy solution); (4) survivors are appended in score order.  Falls back to the generic
        // step loop below when a class has more than CLS_CAP candidates in the batch or the window test failed.
        int my_pos = 0;
        if (cls_path) {
            const unsigned peers = __match_any_sync(KODB_FULL_MASK, my_c);
            const int rank = __popc(peers & ((1u << lane) - 1u));
            if (my_c >= 0 && rank == 0) s.hist[wid * C + my_c] = __popc(peers);
            int kept_c = 0, kept_r = 0;  // survivors of earlier batches -> per-class lists (unordered inside a class)
            if (tid < nk) {
                kept_c = s.kcls[tid];
                kept_r = atomicAdd(&s.kc_cnt[kept_c], 1);
            }
            __syncthreads();
            if (wid == NMS_THREADS / 32 - 1 && nk > 0) {
                int run = 0;
                for (int c0 = 0; c0 < C; c0 += 32) {
                    const int c = c0 + lane;
                    const int v = c < C ? s.kc_cnt[c] : 0;
                    const int ex = warp_excl_scan_i(v, lane);
                    if (c < C) s.kc_base[c] = run + ex;
                    run += __shfl_sync(KODB_FULL_MASK, ex + v, 31);
                }
            }
            for (int c = wid; c < C; c += NMS_THREADS / 32) {  // exclusive prefix over the warps, per class (lane = warp)
                const int v = lane < NMS_THREADS / 32 ? s.hist[lane * C + c] : 0;
                const int ex = warp_excl_scan_i(v, lane);
                if (lane < NMS_THREADS / 32) s.hist[lane * C + c] = ex;
                if (lane == 31) s.ccount[c] = ex + v;
            }
            __syncthreads();
            if (tid < nk) s.kc_idx[s.kc_base[kept_c] + kept_r] = (unsigned short)tid;
            if (wid == 0) {  // class bases, tile starts, capacity check
                int run_c = 0, run_t = 0;
                bool over = false;
                for (int c0 = 0; c0 < C; c0 += 32) {
                    const int c = c0 + lane;
                    const int cnt = c < C ? s.ccount[c] : 0;
                    const int nb = (cnt + 31) >> 5, tl = nb * (nb + 1) / 2;
                    const int ex = warp_excl_scan_i(cnt, lane), tex = warp_excl_scan_i(tl, lane);
                    if (c < C) { s.cbase[c] = run_c + ex; s.tstart[c] = run_t + tex; }
                    run_c += __shfl_sync(KODB_FULL_MASK, ex + cnt, 31);
                    run_t += __shfl_sync(KODB_FULL_MASK, tex + tl, 31);
                    over |= cnt > CLS_CAP;
                }
                over = __any_sync(KODB_FULL_MASK, over);
                if (lane == 0) { s.tstart[C] = run_t; s.cls_over = over; }
            }
            __syncthreads();
            KODB_TRACE(10);
            cls_path = !s.cls_over;
            if (cls_path) {
                if (my_c >= 0) {
                    my_pos = s.cbase[my_c] + s.hist[wid * C + my_c] + rank;
                    const float x1 = s.bx1[tid], y1 = s.by1[tid], x2 = s.bx2[tid], y2 = s.by2[tid], ar = s.barea[tid];
                    s.cbox[my_pos] = make_float4(x1, y1, x2, y2);
                    s.carea[my_pos] = ar;
                    bool dead = false;  // suppressed by a survivor of an earlier batch (same class only)
                    if (nk > 0) {
                        const int q0 = s.kc_base[my_c], q1 = q0 + s.kc_cnt[my_c];
                        for (int qi = q0; qi < q1 && !dead; ++qi) {
                            const int q = s.kc_idx[qi];
                            dead = suppresses(s.kx1[q], s.ky1[q], s.kx2[q], s.ky2[q], s.karea[q], x1, y1, x2, y2, ar, iou_thr);
                        }
                    }
                    s.pre[my_pos] = dead;
                }
                __syncthreads();
                KODB_TRACE(11);
                const int T = s.tstart[C];
                for (int tt = wid; tt < T; tt += NMS_THREADS / 32) {
                    int c = 0;
                    {
                        int hi_ = C - 1;
                        while (c < hi_) {  // largest c with tstart[c] <= tt
                            const int mid = (c + hi_ + 1) >> 1;
                            if (s.tstart[mid] <= tt) c = mid; else hi_ = mid - 1;
                        }
                    }
                    const int k = tt - s.tstart[c];
                    int I = 0;
                    while ((I + 1) * (I + 2) / 2 <= k) ++I;
                    const int J = k - I * (I + 1) / 2;
                    const int base = s.cbase[c], nc = s.ccount[c];
                    const int i = 32 * I + lane;
                    const bool vi = i < nc;
                    const int pi = base + (vi ? i : 0);
                    const float4 bi = s.cbox[pi];
                    const float ia = s.carea[pi];
                    const int jend = min(32, min(nc, 32 * I + 31) - 32 * J);  // j < i <= 32I+31
                    // Branch-free, unrolled: the iterations are independent, so their latency chains overlap.
                    // (double)ovr > thr  <=>  ovr > thr_rd (thr rounded DOWN to fp32).  The IEEE division is only
                    // executed when the approximate ratio is within 1e-6 (relative) of the threshold.
                    unsigned word = 0u;
                    const int pj0 = base + 32 * J;
#pragma unroll 4
                    for (int jj = 0; jj < jend; ++jj) {
                        const float4 bj = s.cbox[pj0 + jj];
                        const float xx1 = fmaxf(bj.x, bi.x), yy1 = fmaxf(bj.y, bi.y);
                        const float xx2 = fminf(bj.z, bi.z), yy2 = fminf(bj.w, bi.w);
                        const float w = fmaxf(0.0f, xx2 - xx1), h = fmaxf(0.0f, yy2 - yy1);
                        const float inter = w * h;
                        const float den = s.carea[pj0 + jj] + ia - inter;
                        const float qa = inter * rcp_approx(den);
                        bool hit = qa > thr_hi;
                        if (!hit && !(qa < thr_lo)) hit = inter / den > thr_rd;  // rare (also NaN / degenerate boxes)
                        word |= (unsigned)hit << jj;
                    }
                    if (I == J) word &= (1u << lane) - 1u;  // diagonal tile: only earlier candidates
                    if (vi) s.pm[(size_t)pi * CLS_WORDS + J] = word;
                }
                __syncthreads();
                KODB_TRACE(12);
                for (int c = wid; c < C; c += NMS_THREADS / 32) {
                    const int nc = s.ccount[c];
                    if (nc == 0) continue;
                    const int base = s.cbase[c], nb = (nc + 31) >> 5;
                    unsigned kw[CLS_WORDS];
#pragma unroll
                    for (int I = 0; I < CLS_WORDS; ++I) {
                        kw[I] = 0u;
                        if (I < nb) {
                            const int i = 32 * I + lane;
                            const bool vi = i < nc;
                            const int pi = base + (vi ? i : 0);
                            bool dead = !vi || s.pre[pi];
#pragma unroll
                            for (int J = 0; J < CLS_WORDS; ++J)
                                if (J < I) dead |= (s.pm[(size_t)pi * CLS_WORDS + J] & kw[J]) != 0u;
                            const unsigned diag = vi ? s.pm[(size_t)pi * CLS_WORDS + I] : 0u;
                            unsigned cur = __ballot_sync(KODB_FULL_MASK, !dead);
                            for (int it = 0; it < 33; ++it) {
                                const unsigned nxt = __ballot_sync(KODB_FULL_MASK, !dead && !(diag & cur));
                                if (nxt == cur) break;
                                cur = nxt;
                            }
                            kw[I] = cur;
                            if (vi) s.keepf[pi] = (cur >> lane) & 1u;
                        }
                    }
                }
                __syncthreads();
                KODB_TRACE(13);
                const int kf = (my_c >= 0 && s.keepf[my_pos]) ? 1 : 0;
                int tot;
                const int ex = block_excl_scan(kf, s.scan_tmp, &tot);
                if (kf) {
                    const int slot = nk + ex;
                    if (slot < max_det) {
                        s.kx1[slot] = s.bx1[tid]; s.ky1[slot] = s.by1[tid]; s.kx2[slot] = s.bx2[tid]; s.ky2[slot] = s.by2[tid];
                        s.karea[slot] = s.barea[tid];
                        s.kcls[slot] = my_c;
                        s.kept_ck[slot] = s.sort_ck[tid];
                        s.kept_idx[slot] = processed + tid;  // rank in the confidence-sorted list (rank mode)
                    }
                }
                nk = min(max_det, nk + tot);
                __syncthreads();
            }
        }
        if (!cls_path) {
            // Generic path, NMS_STEP candidates per step.
            // Warp w owns candidates 4w..4w+3 of the step.  (1) against the kept list: each kept box is loaded once
            // and tested against the warp's four candidates; (2) against the earlier candidates of the same step: a
            // 128x128 bit matrix built with ballots; (3) warp 0 resolves the step's triangular dependency by
            // fixed-point iteration (exact: the only fixed point is the greedy solution) and appends the survivors.
            for (int s0 = 0; s0 < mm && nk < max_det; s0 += NMS_STEP) {
                float cb[4][5];
                bool valid[4], sup[4];
    #pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int e = s0 + 4 * wid + r;
                    valid[r] = e < mm;
                    sup[r] = false;
                    const int ee = valid[r] ? e : 0;
                    cb[r][0] = s.bx1[ee]; cb[r][1] = s.by1[ee]; cb[r][2] = s.bx2[ee]; cb[r][3] = s.by2[ee]; cb[r][4] = s.barea[ee];
                }
                if (valid[0]) {
                    for (int q = lane; q < nk; q += 32) {
                        const float k1 = s.kx1[q], k2 = s.ky1[q], k3 = s.kx2[q], k4 = s.ky2[q], ka = s.karea[q];
    #pragma unroll
                        for (int r = 0; r < 4; ++r)
                            sup[r] |= suppresses(k1, k2, k3, k4, ka, cb[r][0], cb[r][1], cb[r][2], cb[r][3], cb[r][4], iou_thr);
                    }
                }
                unsigned msk[4][NMS_STEP / 32];
    #pragma unroll
                for (int i = 0; i < NMS_STEP / 32; ++i) {
                    const int oi = 32 * i + lane;  // index inside the step of the "other" candidate
                    const int o = s0 + oi;
                    const bool ov = o < mm && 32 * i < 4 * wid + 4;  // only earlier candidates matter
                    const int oo = ov ? o : 0;
                    const float o1 = s.bx1[oo], o2 = s.by1[oo], o3 = s.bx2[oo], o4 = s.by2[oo], oa = s.barea[oo];
    #pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const bool hit = ov && valid[r] && oi < 4 * wid + r &&
                                         suppresses(o1, o2, o3, o4, oa, cb[r][0], cb[r][1], cb[r][2], cb[r][3], cb[r][4], iou_thr);
                        msk[r][i] = __ballot_sync(KODB_FULL_MASK, hit);
                    }
                }
    #pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const bool any = __any_sync(KODB_FULL_MASK, sup[r]);
                    if (lane == 0) {
                        s.alive[4 * wid + r] = valid[r] && !any;
    #pragma unroll
                        for (int i = 0; i < NMS_STEP / 32; ++i) s.mask[4 * wid + r][i] = msk[r][i];
                    }
                }
                __syncthreads();
                if (wid == 0) {
                    // lane l looks after candidates l, l+32, l+64, l+96, so that ballot word i covers 32i..32i+31
                    bool al[NMS_STEP / 32];
                    unsigned mk[NMS_STEP / 32][NMS_STEP / 32], kept[NMS_STEP / 32];
    #pragma unroll
                    for (int i = 0; i < NMS_STEP / 32; ++i) {
                        al[i] = s.alive[32 * i + lane] != 0;
    #pragma unroll
                        for (int j = 0; j < NMS_STEP / 32; ++j) mk[i][j] = s.mask[32 * i + lane][j];
                        kept[i] = __ballot_sync(KODB_FULL_MASK, al[i]);
                    }
                    for (int it = 0; it <= NMS_STEP; ++it) {
                        unsigned nb[NMS_STEP / 32];
                        bool same = true;
    #pragma unroll
                        for (int i = 0; i < NMS_STEP / 32; ++i) {
                            unsigned hit = 0u;
    #pragma unroll
                            for (int j = 0; j < NMS_STEP / 32; ++j) hit |= mk[i][j] & kept[j];
                            nb[i] = __ballot_sync(KODB_FULL_MASK, al[i] && !hit);
                            same &= nb[i] == kept[i];
                        }
    #pragma unroll
                        for (int i = 0; i < NMS_STEP / 32; ++i) kept[i] = nb[i];
                        if (same) break;
                    }
                    int before = 0;
    #pragma unroll
                    for (int i = 0; i < NMS_STEP / 32; ++i) {
                        if ((kept[i] >> lane) & 1u) {
                            const int slot = nk + before + __popc(kept[i] & ((1u << lane) - 1u));
                            if (slot < max_det) {
                                const int o = s0 + 32 * i + lane;
                                s.kx1[slot] = s.bx1[o]; s.ky1[slot] = s.by1[o]; s.kx2[slot] = s.bx2[o]; s.ky2[slot] = s.by2[o];
                                s.karea[slot] = s.barea[o];
                                s.kept_ck[slot] = s.sort_ck[o];
                                s.kcls[slot] = (int)((uint32_t)s.sort_ck[o] % (uint32_t)C);
                                s.kept_idx[slot] = processed + o;  // rank in the confidence-sorted list (rank mode)
                            }
                        }
                        before += __popc(kept[i]);
                    }
                    if (lane == 0) s.nk = min(max_det, nk + before);
                }
                __syncthreads();
                nk = s.nk;
            }
        }
        KODB_TRACE(6);
        processed += m;
        lo = hi_batch;
        k_target = min(2 * k_target, K_CAP / 2);
        if (hi_batch == U64MAX) break;
    }
    __syncthreads();

    // ---------------------------------------------------------------------- reference candidate index
    // nms.py:47: candidates are listed in (row, class) order, i.e. by ascending key.  index(key) = number of
    // candidates of the image with a smaller key = [chunks of earlier tiles] + [smaller keys inside the
    // chunks of its own tile].  One warp per survivor.
    if (!rank_mode && floor_mode) {
        // from the class masks: candidates of earlier rows of the tile + earlier classes of the same row
        for (int q = wid; q < nk; q += NMS_THREADS / 32) {
            const uint32_t k32 = (uint32_t)s.kept_ck[q];
            const int row = (int)(k32 / (uint32_t)C), j = (int)(k32 - (uint32_t)row * (uint32_t)C);
            const int tl = tile_of_row(p, row);
            int L, a, c0, ncell;
            planar_tile(p.g, tl, L, a, c0, ncell);
            const int row0 = p.g.row_base[L] + a * (p.g.lv[L].H * p.g.lv[L].W) + c0;
            const int r = row - row0;
            int cnt = 0;
            for (int g = 0; g < p.ngroups; ++g) {
                const unsigned short* mk = mask_img + (int64_t)g * p.rows + row0;
                const int jl = j - g * CGROUP;  // classes of this group below j
                const unsigned below = jl >= CGROUP ? 0xffffu : (jl <= 0 ? 0u : ((1u << jl) - 1u));
                for (int i = lane; i <= r; i += 32) {
                    const unsigned m16 = mk[i];
                    cnt += __popc(i < r ? m16 : (m16 & below));
                }
            }
            cnt = warp_sum_i(cnt);
            if (lane == 0) s.kept_idx[q] = (long long)__ldcg(cprefix + tl * p.ngroups) + cnt;
        }
        __syncthreads();
    } else if (!rank_mode) {
        for (int q = wid; q < nk; q += NMS_THREADS / 32) {
            const uint32_t k32 = (uint32_t)s.kept_ck[q];
            const int tl = tile_of_row(p, (int)(k32 / (uint32_t)C));
            int cnt = 0;
            for (int g = 0; g < p.ngroups; ++g) {
                const int ch = tl * p.ngroups + g;
                const uint32_t* kk = key + cstart[ch];
                const int nc = ccount[ch];
                for (int i = lane; i < nc; i += 32) cnt += (kk[i] < k32);
            }
            cnt = warp_sum_i(cnt);
            if (lane == 0) s.kept_idx[q] = (long long)__ldcg(cprefix + tl * p.ngroups) + cnt;
        }
        __syncthreads();
    }

    // ---------------------------------------------------------------------- outputs
    // stage the kept detections (un-offset boxes, nms.py:73) in registers, then reuse shared memory for matching
    float4 rb = make_float4(0.f, 0.f, 0.f, 0.f);
    float rs = 0.0f;
    int rc = -1;
    long long ridx = 0, rkey = -1;
    if (tid < nk) {
        const unsigned long long ck = s.kept_ck[tid];
        const uint32_t k32 = (uint32_t)ck;
        rkey = (long long)k32;
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
        if (p.out.keep_key) p.out.keep_key[(int64_t)b * max_det + tid] = rkey;
    }
    if (tid == 0) {
        p.out.counts[b] = nk;
        if (p.out.num_candidates) p.out.num_candidates[b] = n_total;
        p.img_count[b] = 0;  // workspace reusable without a memset
    }
    KODB_TRACE(7);
    if (p.do_match) {
        __syncthreads();  // everyone is done with NmsSmem
        if (tid < nk) {
            ms.x1[tid] = rb.x; ms.y1[tid] = rb.y; ms.x2[tid] = rb.z; ms.y2[tid] = rb.w;
            ms.score[tid] = rs; ms.cls[tid] = rc;
        }
        __syncthreads();
        match_image(ms, GtInSmem{gs}, nk, C, gt_n, p.m, b, max_det);
    }
    KODB_TRACE(8);
    if (p.trace && tid == 0) {
        long long gt1;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt1));
        p.trace[(int64_t)b * 128 + 127] = gt1;
    }
}

// standalone matching on NMS output (B, max_det, 6)
__global__ void __launch_bounds__(NMS_THREADS, 1)
match_kernel(const float* __restrict__ dets, const int32_t* __restrict__ counts, int C, int max_det, const void* gt_boxes,
             const int64_t* gt_labels, const int32_t* gt_offsets, int gt_dtype, const __grid_constant__ kodb_match_out mo) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MatchSmem& ms = *reinterpret_cast<MatchSmem*>(smem_raw);
    GtSmem& gs = *reinterpret_cast<GtSmem*>(smem_raw + kMatchRegion);
    const int b = blockIdx.x, tid = threadIdx.x;
    const int nk = min(counts[b], max_det);
    const int g0 = gt_offsets[b], ng = gt_offsets[b + 1] - g0;
    stage_gt(gs, C, gt_boxes, gt_labels, g0, ng, gt_dtype, mo);
    if (tid < nk) {
        const float* d = dets + ((int64_t)b * max_det + tid) * 6;
        ms.x1[tid] = d[0]; ms.y1[tid] = d[1]; ms.x2[tid] = d[2]; ms.y2[tid] = d[3];
        ms.score[tid] = d[4]; ms.cls[tid] = (int)d[5];
    }
    __syncthreads();
    match_image(ms, GtInSmem{gs}, nk, C, ng, mo, b, max_det);
}

// Crowded images (more than GT_CAP ground-truth boxes in one image): the same matching with the ground truth staged in a
// global scratch instead of shared memory -- slower (L2 round trips in the greedy loop), no limit.  One CTA per image.
__global__ void __launch_bounds__(NMS_THREADS, 1)
match_crowded_kernel(const float* __restrict__ dets, const int32_t* __restrict__ counts, int C, int max_det, const void* gt_boxes,
                     const int64_t* gt_labels, const int32_t* gt_offsets, int gt_dtype, const __grid_constant__ kodb_match_out mo,
                     unsigned char* scratch, size_t per_image, int cap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MatchSmem& ms = *reinterpret_cast<MatchSmem*>(smem_raw);
    const int b = blockIdx.x, tid = threadIdx.x;
    const int nk = min(counts[b], max_det);
    const int g0 = gt_offsets[b], ng = gt_offsets[b + 1] - g0;
    unsigned char* base = scratch + (size_t)b * per_image;
    GtInGlobal gs;
    gs.xywh = reinterpret_cast<double*>(base);
    gs.label = reinterpret_cast<int*>(base + (size_t)4 * cap * sizeof(double));
    gs.flags = base + (size_t)4 * cap * sizeof(double) + (size_t)cap * sizeof(int);
    gs.cap = cap;
    for (int i = tid; i < KODB_NUM_IOU_THRS * cap; i += blockDim.x) gs.flags[i] = 0;
    for (int g = tid; g < ng; g += blockDim.x) {
        const long long lab = gt_labels[g0 + g];
        if (lab >= 0 && lab < C) atomicAdd(reinterpret_cast<unsigned long long*>(mo.gt_count) + lab, 1ull);
        double l, t, r, bt;
        load_gt_xyxy(gt_boxes, gt_dtype, g0 + g, l, t, r, bt);
        gs.xywh[g] = l; gs.xywh[cap + g] = t; gs.xywh[2 * cap + g] = r - l; gs.xywh[3 * cap + g] = bt - t;  // [l, t, r-l, b-t]
        gs.label[g] = (int)lab;
    }
    if (tid < nk) {
        const float* d = dets + ((int64_t)b * max_det + tid) * 6;
        ms.x1[tid] = d[0]; ms.y1[tid] = d[1]; ms.x2[tid] = d[2]; ms.y2[tid] = d[3];
        ms.score[tid] = d[4]; ms.cls[tid] = (int)d[5];
    }
    __threadfence_block();
    __syncthreads();
    match_image(ms, gs, nk, C, ng, mo, b, max_det);
}

// stream-ordered scratch for the crowded path (rare: allocated per call, freed behind the kernel)
static int launch_match_crowded(const float* dets, const int32_t* counts, int batch, int max_det, int C, const kodb_targets* gt,
                                const kodb_match_out* m, cudaStream_t st) {
    const int cap = (gt->max_per_image + 7) / 8 * 8;
    const size_t per_image = ((size_t)cap * (4 * sizeof(double) + sizeof(int) + KODB_NUM_IOU_THRS) + 15) & ~(size_t)15;
    void* scratch = nullptr;
    if (cudaMallocAsync(&scratch, per_image * batch, st) != cudaSuccess) {
        (void)cudaGetLastError();
        return KODB_ERR_CAPACITY;
    }
    static bool attr_set[64] = {false};
    std::lock_guard<std::mutex> host_lock(g_eval_host_mu);
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        if (cudaFuncSetAttribute(match_crowded_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMatchRegion) != cudaSuccess) {
            cudaFreeAsync(scratch, st);
            return KODB_ERR_LAUNCH;
        }
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    match_crowded_kernel<<<batch, NMS_THREADS, kMatchRegion, st>>>(dets, counts, C, max_det, gt->boxes, gt->labels, gt->img_offsets,
                                                                   gt->boxes_dtype, *m, (unsigned char*)scratch, per_image, cap);
    const cudaError_t e = cudaGetLastError();
    cudaFreeAsync(scratch, st);
    return e == cudaSuccess ? KODB_OK : KODB_ERR_LAUNCH;
}

// ---------------------------------------------------------------------------------------------------------
static bool aligned16e(const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15u) == 0; }

// tile plan of the planar candidate kernel; returns tiles per image and whether float4 planes are usable
static int plan_geom(EvalGeom& g, const kodb_head_level* heads_h, const kodb_level_cfg* levels_h, int C, int* vec_out) {
    int vec = 4;
    for (int l = 0; l < KODB_NUM_LEVELS; ++l) {
        const kodb_head_level& h = heads_h[l];
        const int HW = levels_h[l].H * levels_h[l].W;
        if (HW <= 0 || levels_h[l].num_anchors < 1 || levels_h[l].num_anchors > KODB_MAX_ANCHORS) return -1;
        if (HW % 4 || h.obj_sb % 4 || h.obj_sa % 4 || h.cls_sb % 4 || h.cls_sa % 4 || h.cls_sp % 4 || !aligned16e(h.obj) ||
            !aligned16e(h.cls))
            vec = 1;
    }
    (void)C;
    int rows = 0, tiles = 0;
    const int cap = vec == 4 ? WT_CELLS : CAND_THREADS;  // warp tiles (cand_floor_kernel) / block tiles (cand_kernel<.,1>)
    for (int l = 0; l < KODB_NUM_LEVELS; ++l) {
        g.head[l] = heads_h[l];
        g.lv[l] = levels_h[l];
        const int HW = levels_h[l].H * levels_h[l].W;
        g.row_base[l] = rows;
        rows += levels_h[l].num_anchors * HW;
        const int tpp = (HW + cap - 1) / cap;
        int tc = (HW + tpp - 1) / tpp;
        tc = (tc + vec - 1) / vec * vec;  // multiple of the vector width (HW % 4 == 0 when vec == 4)
        g.tiles_per_plane[l] = (HW + tc - 1) / tc;
        g.tile_cells[l] = tc;
        g.tile_begin[l] = tiles;
        tiles += g.tiles_per_plane[l] * levels_h[l].num_anchors;
    }
    g.row_base[KODB_NUM_LEVELS] = rows;
    g.tile_begin[KODB_NUM_LEVELS] = tiles;
    *vec_out = vec;
    return tiles;
}

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

static size_t cand_cap(int64_t rows, int C) { return ((size_t)rows * C + 3) & ~(size_t)3; }  // float4-aligned per image

static size_t mask_cap(int64_t rows, int C) { return (((size_t)rows + 3) & ~(size_t)3) * ((C + CGROUP - 1) / CGROUP); }

static int64_t ws_bytes(int batch, int64_t rows, int C, int chunks_per_image) {
    const size_t cap = cand_cap(rows, C);
    return (int64_t)(2 * align256((size_t)batch * 4) + align256((size_t)batch * HB * 4) +
                     3 * align256((size_t)batch * chunks_per_image * 4) + 2 * align256((size_t)batch * cap * 4) +
                     align256((size_t)batch * mask_cap(rows, C) * 2));
}

static int carve_ws(EvalParams& p, void* ws, int64_t ws_bytes_avail) {
    const size_t cap = cand_cap(p.rows, p.C);
    if (ws_bytes_avail < ws_bytes(p.B, p.rows, p.C, p.chunks_per_image)) return KODB_ERR_CAPACITY;
    char* w = reinterpret_cast<char*>(ws);
    // the zero-on-entry arrays come first, at offsets that depend on the batch size only (include/kodb.h: a
    // workspace is reusable across calls with the same batch size)
    p.img_count = reinterpret_cast<int*>(w);
    w += align256((size_t)p.B * 4);
    p.hist = reinterpret_cast<int*>(w);
    w += align256((size_t)p.B * HB * 4);
    p.floor_word = reinterpret_cast<int*>(w);
    w += align256((size_t)p.B * 4);
    p.chunk_start = reinterpret_cast<int*>(w);
    w += align256((size_t)p.B * p.chunks_per_image * 4);
    p.chunk_count = reinterpret_cast<int*>(w);
    w += align256((size_t)p.B * p.chunks_per_image * 4);
    p.chunk_prefix = reinterpret_cast<int*>(w);
    w += align256((size_t)p.B * p.chunks_per_image * 4);
    p.cand_score = reinterpret_cast<float*>(w);
    w += align256((size_t)p.B * cap * 4);
    p.cand_key = reinterpret_cast<uint32_t*>(w);
    w += align256((size_t)p.B * cap * 4);
    p.cand_stride = (int64_t)cap;
    p.mask = reinterpret_cast<unsigned short*>(w);
    p.mask_stride = (int64_t)mask_cap(p.rows, p.C);
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

// *crowded: an image has more ground-truth boxes than the select kernel stages in shared memory; the matching then runs
// as its own launch behind the select kernel (match_crowded_kernel) instead of inside it
static int setup_match(EvalParams& p, const kodb_targets* gt, const kodb_match_out* m, bool* crowded) {
    p.do_match = 0;
    *crowded = false;
    if (!gt) return KODB_OK;
    if (!m || !m->rec_score || !m->rec_cls || !m->rec_tp || !m->gt_count) return KODB_ERR_INVALID_ARGUMENT;
    if (gt->batch != p.B) return KODB_ERR_INVALID_ARGUMENT;
    if (gt->max_per_image > GT_CAP) {
        *crowded = true;
        return KODB_OK;
    }
    p.do_match = 1;
    p.gt_boxes = gt->boxes;
    p.gt_labels = gt->labels;
    p.gt_offsets = gt->img_offsets;
    p.gt_dtype = gt->boxes_dtype;
    p.m = *m;
    return KODB_OK;
}

static long long* g_trace = nullptr;
static int g_hot_target = HOT_TARGET;
static int g_floor_polls = 0;
static int g_l2_stream = 1;

// m = ceil(2^32 / d) with umulhi(n, m) == n / d for every n <= n_max, or 0 if that does not hold
static unsigned magic_div(unsigned d, uint64_t n_max) {
    if (d < 2) return 0u;
    const uint64_t m = ((1ull << 32) + d - 1) / d;
    const uint64_t e = m * d - (1ull << 32);  // 0 <= e < d;  exact while n * e < 2^32
    if (m >= (1ull << 32) || n_max * e >= (1ull << 32)) return 0u;
    return (unsigned)m;
}

static int launch_select(const EvalParams& p_in, cudaStream_t st) {
    EvalParams p = p_in;
    p.trace = g_trace;
    const size_t smem = kNmsRegion + sizeof(GtSmem);
    static bool attr_set[64] = {false};
    std::lock_guard<std::mutex> host_lock(g_eval_host_mu);
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

// debug hook (not part of include/kodb.h): device buffer of B*128 int64 receiving per-image phase stamps
extern "C" void kodb_debug_set_nms_trace(void* ptr) { kodb::g_trace = reinterpret_cast<long long*>(ptr); }
// debug hook: the number of listed candidates per image the score floor aims at (small values force the
// select kernel's completion pass; used by the parity tests).  <= 0 restores the default.
// debug hook: number of completion passes (images whose listed candidates ran dry) since the last call; synchronises
extern "C" int kodb_debug_take_completions(void) {
    unsigned int v = 0u, z = 0u;
    if (cudaMemcpyFromSymbol(&v, kodb::g_completions, sizeof(v)) != cudaSuccess) return -1;
    cudaMemcpyToSymbol(kodb::g_completions, &z, sizeof(z));
    return (int)v;
}
// debug hook: let every tile wait (bounded, n x 100 ns) for its image's floor instead of listing everything when it
// comes first -- makes small test batches use the floors deterministically.  0 restores the default.
extern "C" void kodb_debug_set_floor_polls(int n) { kodb::g_floor_polls = n > 0 ? n : 0; }
extern "C" void kodb_debug_set_hot_target(int n) { kodb::g_hot_target = n > 0 ? n : kodb::HOT_TARGET; }

// Tensor maps for the candidate kernel's tile loads: possible when every level's objectness and class tensors are flat planar
// arrays [B][A][P][H*W] (the Detect head's own layout); anything else keeps the per-lane cp.async copies.
static void setup_cand_tma(kodb::EvalParams& p, int nq) {
    using namespace kodb;
    p.use_tma = 0;
    static int knob = -1;
    std::lock_guard<std::mutex> host_lock(g_eval_host_mu);
    if (knob < 0) {
        const char* e = getenv("KODB_EVAL_TMA");
        knob = e ? (atoi(e) != 0) : 1;
    }
    EncodeTiledFn encode = knob ? tensormap_encoder() : nullptr;
    if (!encode) return;
    for (int l = 0; l < KODB_NUM_LEVELS; ++l) {
        const kodb_head_level& h = p.g.head[l];
        const int64_t HW = (int64_t)p.g.lv[l].H * p.g.lv[l].W, A = p.g.lv[l].num_anchors, C = p.C;
        if (HW % 4 || h.obj_sa != HW || h.obj_sb != A * HW || h.cls_sp != HW || h.cls_sa != C * HW || h.cls_sb != A * C * HW) return;
        if ((reinterpret_cast<uintptr_t>(h.obj) | reinterpret_cast<uintptr_t>(h.cls)) & 15u) return;
        const cuuint32_t estr[3] = {1, 1, 1};
        {
            const cuuint64_t dims[3] = {(cuuint64_t)HW, (cuuint64_t)A, (cuuint64_t)p.B};
            const cuuint64_t strides[2] = {(cuuint64_t)HW * 4, (cuuint64_t)(A * HW) * 4};
            const cuuint32_t box[3] = {(cuuint32_t)WT_CELLS, 1, 1};
            if (encode(&p.tma_obj[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(h.obj), dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return;
        }
        {
            const cuuint64_t dims[3] = {(cuuint64_t)HW, (cuuint64_t)(A * C), (cuuint64_t)p.B};
            const cuuint64_t strides[2] = {(cuuint64_t)HW * 4, (cuuint64_t)(A * C * HW) * 4};
            const cuuint32_t box[3] = {(cuuint32_t)WT_CELLS, (cuuint32_t)nq, 1};
            if (encode(&p.tma_cls[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(h.cls), dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return;
        }
    }
    p.use_tma = 1;
}

extern "C" int64_t kodb_eval_workspace_bytes(int32_t batch, int64_t rows_per_image, int32_t num_classes) {
    // upper bound of chunks per image: tiles <= ceil(rows/256) + one ragged tile per (level, anchor) plane
    const int tiles = (int)((rows_per_image + kodb::WT_CELLS - 1) / kodb::WT_CELLS) + KODB_NUM_LEVELS * KODB_MAX_ANCHORS;
    const int groups = (num_classes + kodb::CGROUP - 1) / kodb::CGROUP;
    return kodb::ws_bytes(batch, rows_per_image, num_classes, tiles * groups);
}

extern "C" int kodb_eval_tail(const kodb_head_level* heads_h, const kodb_level_cfg* levels_h, int32_t batch,
                              const kodb_nms_cfg* cfg_h, const kodb_nms_out* out_h, const kodb_targets* gt_h,
                              const kodb_match_out* match_h, void* workspace, int64_t workspace_bytes, void* stream) {
    using namespace kodb;
    if (!heads_h || !levels_h || !workspace || !cfg_h) return KODB_ERR_INVALID_ARGUMENT;
    EvalParams p;
    int vec = 1;
    const int tiles = plan_geom(p.g, heads_h, levels_h, cfg_h->num_classes, &vec);
    if (tiles < 0) return KODB_ERR_INVALID_ARGUMENT;
    p.det = nullptr;
    p.rows = p.g.row_base[KODB_NUM_LEVELS];
    int rc = check_common(cfg_h, out_h, batch, p.rows);
    if (rc != KODB_OK) return rc;
    p.planar = 1;
    p.B = batch;
    p.C = cfg_h->num_classes;
    p.vec = vec;
    p.ngroups = (p.C + CGROUP - 1) / CGROUP;
    p.cfg = *cfg_h;
    p.out = *out_h;
    p.tiles_per_image = tiles;
    p.chunks_per_image = tiles * p.ngroups;
    p.trace = nullptr;
    p.floor_mode = vec == 4;
    p.use_tma = 0;
    p.sample_stride = (tiles + CW_WARPS - 1) / CW_WARPS;  // at most one sample tile per warp of the sample CTA
    p.ns = (tiles + p.sample_stride - 1) / p.sample_stride;
    p.magic_b = magic_div((unsigned)batch, (uint64_t)tiles * batch + CW_WARPS);
    p.magic_s1 = p.sample_stride > 1 ? magic_div((unsigned)(p.sample_stride - 1), (uint64_t)tiles) : 0u;
    static bool env_read = false;
    if (std::lock_guard<std::mutex> host_lock(g_eval_host_mu); !env_read) {  // tuning knob (like KODB_TRAIN_DENSE_PER_SM); the debug hook overrides it
        env_read = true;
        const char* e = getenv("KODB_HOT_TARGET");
        if (e && atoi(e) > 0) g_hot_target = atoi(e);
        e = getenv("KODB_EVAL_L2_STREAM");
        if (e) g_l2_stream = atoi(e) != 0;
    }
    p.hot_target = g_hot_target;
    p.floor_polls = g_floor_polls;
    p.l2_stream = g_l2_stream;
    if (p.chunks_per_image > CHUNK_CAP) return KODB_ERR_UNSUPPORTED;
    if ((rc = carve_ws(p, workspace, workspace_bytes)) != KODB_OK) return rc;
    bool crowded = false;
    if ((rc = setup_match(p, gt_h, match_h, &crowded)) != KODB_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int stages = cfg_h->stage_mask ? cfg_h->stage_mask : 3;
    if (stages & 1) {
        if (vec == 4) {
            const dim3 grid((unsigned)(batch + ((int64_t)(tiles - p.ns) * batch + CW_WARPS - 1) / CW_WARPS));
            const int nq = p.C < CGROUP ? p.C : CGROUP;
            setup_cand_tma(p, nq <= 4 ? 4 : nq <= 8 ? 8 : nq <= 10 ? 10 : nq <= 12 ? 12 : 16);
            rc = nq <= 4 ? launch_cand_warp<4>(p, grid, st) : nq <= 8 ? launch_cand_warp<8>(p, grid, st)
                 : nq <= 10 ? launch_cand_warp<10>(p, grid, st) : nq <= 12 ? launch_cand_warp<12>(p, grid, st)
                                                                            : launch_cand_warp<16>(p, grid, st);
            if (rc != KODB_OK) return rc;
        } else {
            cand_kernel<true, 1><<<dim3(tiles, batch), CAND_THREADS, 0, st>>>(p);
            KODB_CHECK_LAUNCH();
        }
    }
    if (!(stages & 2)) return KODB_OK;
    if ((rc = launch_select(p, st)) != KODB_OK || !crowded) return rc;
    return launch_match_crowded(p.out.dets, p.out.counts, p.B, p.cfg.max_det, p.C, gt_h, match_h, st);
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
    p.vec = 1;
    p.ngroups = (p.C + CGROUP - 1) / CGROUP;
    p.cfg = *cfg_h;
    p.out = *out_h;
    p.tiles_per_image = (int)((rows + CAND_THREADS - 1) / CAND_THREADS);
    p.chunks_per_image = p.tiles_per_image * p.ngroups;
    p.trace = nullptr;
    p.floor_mode = 0;
    p.use_tma = 0;
    p.ns = 0;
    p.sample_stride = 1;
    p.magic_b = p.magic_s1 = 0u;
    p.hot_target = 0;
    p.floor_polls = 0;
    p.l2_stream = 0;
    if (p.chunks_per_image > CHUNK_CAP) return KODB_ERR_UNSUPPORTED;
    if ((rc = carve_ws(p, workspace, workspace_bytes)) != KODB_OK) return rc;
    bool crowded = false;
    if ((rc = setup_match(p, gt_h, match_h, &crowded)) != KODB_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int stages = cfg_h->stage_mask ? cfg_h->stage_mask : 3;
    if (stages & 1) {
        cand_kernel<false, 1><<<dim3(p.tiles_per_image, batch), CAND_THREADS, 0, st>>>(p);
        KODB_CHECK_LAUNCH();
    }
    if (!(stages & 2)) return KODB_OK;
    if ((rc = launch_select(p, st)) != KODB_OK || !crowded) return rc;
    return launch_match_crowded(p.out.dets, p.out.counts, p.B, p.cfg.max_det, p.C, gt_h, match_h, st);
}

extern "C" int kodb_map_match(const float* dets, const int32_t* counts, int32_t batch, int32_t max_det,
                              int32_t num_classes, const kodb_targets* gt_h, const kodb_match_out* match_h, void* stream) {
    using namespace kodb;
    if (!dets || !counts || !gt_h || !match_h || batch < 1 || max_det < 1 || max_det > MAX_DET_CAP || num_classes < 1 ||
        num_classes > KODB_MAX_CLASSES)
        return KODB_ERR_INVALID_ARGUMENT;
    if (gt_h->batch != batch) return KODB_ERR_INVALID_ARGUMENT;
    if (gt_h->max_per_image > GT_CAP)
        return launch_match_crowded(dets, counts, batch, max_det, num_classes, gt_h, match_h, (cudaStream_t)stream);
    static bool attr_set[64] = {false};
    std::lock_guard<std::mutex> host_lock(g_eval_host_mu);
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        if (cudaFuncSetAttribute(match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kMatchRegion + sizeof(GtSmem))) != cudaSuccess)
            return KODB_ERR_LAUNCH;
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    match_kernel<<<batch, NMS_THREADS, kMatchRegion + sizeof(GtSmem), (cudaStream_t)stream>>>(
        dets, counts, num_classes, max_det, gt_h->boxes, gt_h->labels, gt_h->img_offsets, gt_h->boxes_dtype, *match_h);
    KODB_CHECK_LAUNCH();
    return KODB_OK;
}
static_assert(sizeof(kodb::EvalParams) <= 4000, "kernel parameter space");
