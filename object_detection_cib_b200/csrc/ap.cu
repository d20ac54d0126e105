// kodb_ap_accumulate — epoch-end COCO accumulation (pycocotools COCOeval.accumulate as driven by
// PyCOCOMAPEvalCallback.on_validation_epoch_end, kod/lightning/callbacks/pycoco_map_eval.py:108-125).
//
// Input: the per-detection records the matching stage appended for every validation image
// (score, class, TP bits per IoU threshold; image-major, detection order = NMS order) and the per-class
// ground-truth counts.  Output: AP[t][k] in fp64 (101-point interpolated), NaN for classes without GT.
//
// Kernel 1 counts records per class.  Kernel 2 runs one CTA per class:
//   a. order-preserving compaction of the class's records (block scan over the slot stream);
//   b. stable LSD radix sort by descending score (4 x 8-bit passes, ranks from warp match + per-warp digit
//      tables) == numpy argsort(-score, kind='mergesort');
//   c. per IoU threshold: cumulative TP (block scan with carry), precision tp/(tp+fp+eps), right-to-left
//      running max (suffix-max scan), then the 101 recall thresholds by binary search (searchsorted left).
#include "kodb_common.cuh"

namespace kodb {

constexpr int AP_THREADS = 1024;
constexpr int AP_WARPS = AP_THREADS / 32;

struct ApParams {
    const float* rec_score;
    const int32_t* rec_cls;
    const uint8_t* rec_tp;
    const int32_t* counts;
    int64_t num_images;
    int max_det, C;
    const int64_t* gt_count;
    double* ap;
    // workspace
    int* cls_count;       // [C]  zero on entry / exit
    uint32_t* key[2];     // [N]
    uint8_t* tpb[2];      // [N]
    int32_t* tpcum;       // [N]
    double* env;          // [N]
};

__device__ __forceinline__ uint32_t ap_desc_key(float s) {
    const uint32_t b = __float_as_uint(s);
    const uint32_t u = b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
    return ~u;
}

__global__ void ap_count_kernel(const __grid_constant__ ApParams p) {
    const int64_t total = p.num_images * p.max_det;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t img = i / p.max_det;
        const int d = (int)(i - img * p.max_det);
        if (d < p.counts[img]) {
            const int c = p.rec_cls[i];
            if (c >= 0 && c < p.C) atomicAdd(&p.cls_count[c], 1);
        }
    }
}

struct ApSmem {
    int scan_tmp[33];
    int warp_digit[AP_WARPS][256];
    int digit_base[256];
    int chunk_tot[256];
    double dscan[AP_WARPS + 1];
    double q[101];
};

__device__ __forceinline__ double block_suffix_max(double v, double* sm /* AP_WARPS+1 */, double carry) {
    // returns max(v_i, v_{i+1}, ..., v_{last}, carry) for thread i; sm[AP_WARPS] = block max incl. carry
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double y = __shfl_down_sync(KODB_FULL_MASK, x, o);
        if (lane + o < 32) x = fmax(x, y);
    }
    __syncthreads();
    if (lane == 0) sm[wid] = x;
    __syncthreads();
    double tail = carry;
    for (int w = AP_WARPS - 1; w > wid; --w) tail = fmax(tail, sm[w]);
    if (threadIdx.x == 0) {
        double all = carry;
        for (int w = 0; w < AP_WARPS; ++w) all = fmax(all, sm[w]);
        sm[AP_WARPS] = all;
    }
    __syncthreads();
    return fmax(x, tail);
}

__global__ void __launch_bounds__(AP_THREADS, 1) ap_class_kernel(const __grid_constant__ ApParams p) {
    __shared__ ApSmem s;
    const int k = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int seg0 = 0;
    for (int c = 0; c < k; ++c) seg0 += p.cls_count[c];
    const int nd = p.cls_count[k];
    const long long npig = p.gt_count[k];
    uint32_t* key0 = p.key[0] + seg0;
    uint32_t* key1 = p.key[1] + seg0;
    uint8_t* tp0 = p.tpb[0] + seg0;
    uint8_t* tp1 = p.tpb[1] + seg0;

    if (npig <= 0) {  // class without ground truth: excluded from every mean (precision = -1 in pycocotools)
        if (tid < KODB_NUM_IOU_THRS) p.ap[(int64_t)tid * p.C + k] = __longlong_as_double(0x7ff8000000000000ll);
        return;
    }
    // ---- a. ordered compaction of this class's records --------------------------------------------------
    {
        const int64_t total = p.num_images * p.max_det;
        int base = 0;
        for (int64_t start = 0; start < total; start += AP_THREADS) {
            const int64_t i = start + tid;
            int flag = 0;
            if (i < total) {
                const int64_t img = i / p.max_det;
                const int d = (int)(i - img * p.max_det);
                flag = (d < p.counts[img]) && (p.rec_cls[i] == k);
            }
            int tot;
            const int pos = base + block_excl_scan(flag, s.scan_tmp, &tot);
            if (flag) {
                key0[pos] = ap_desc_key(p.rec_score[i]);
                tp0[pos] = p.rec_tp[i];
            }
            base += tot;
        }
    }
    __syncthreads();
    // ---- b. stable LSD radix sort (ascending desc_key) ----------------------------------------------------
    uint32_t* ksrc = key0; uint32_t* kdst = key1;
    uint8_t* tsrc = tp0; uint8_t* tdst = tp1;
    for (int pass = 0; pass < 4; ++pass) {
        const int sh = pass * 8;
        // digit histogram
        for (int i = tid; i < 256; i += AP_THREADS) s.digit_base[i] = 0;
        __syncthreads();
        for (int i = tid; i < nd; i += AP_THREADS) atomicAdd(&s.digit_base[(ksrc[i] >> sh) & 255], 1);
        __syncthreads();
        if (wid == 0) {  // exclusive prefix over 256 digits
            int carry = 0;
            for (int c0 = 0; c0 < 256; c0 += 32) {
                const int v = s.digit_base[c0 + lane];
                const int ex = warp_excl_scan_i(v, lane);
                s.digit_base[c0 + lane] = carry + ex;
                carry += __shfl_sync(KODB_FULL_MASK, ex + v, 31);
            }
        }
        __syncthreads();
        for (int start = 0; start < nd; start += AP_THREADS) {
            const int i = start + tid;
            const bool valid = i < nd;
            const uint32_t kv = valid ? ksrc[i] : 0u;
            const uint8_t tv = valid ? tsrc[i] : (uint8_t)0;
            const int dg = valid ? (int)((kv >> sh) & 255) : 256 + lane;  // invalid lanes never match anyone
            for (int j = lane; j < 256; j += 32) s.warp_digit[wid][j] = 0;
            __syncwarp();
            const unsigned peers = __match_any_sync(KODB_FULL_MASK, dg);
            const int rank_in_warp = __popc(peers & ((1u << lane) - 1u));
            if (valid && rank_in_warp == 0) s.warp_digit[wid][dg] = __popc(peers);
            __syncthreads();
            // per digit: exclusive prefix over warps; the chunk total advances the running base afterwards
            if (tid < 256) {
                int run = 0;
                for (int w = 0; w < AP_WARPS; ++w) {
                    const int c = s.warp_digit[w][tid];
                    s.warp_digit[w][tid] = run;
                    run += c;
                }
                s.chunk_tot[tid] = run;
            }
            __syncthreads();
            if (valid) {
                const int pos = s.digit_base[dg] + s.warp_digit[wid][dg] + rank_in_warp;
                kdst[pos] = kv;
                tdst[pos] = tv;
            }
            __syncthreads();
            if (tid < 256) s.digit_base[tid] += s.chunk_tot[tid];
            __syncthreads();
        }
        uint32_t* tk = ksrc; ksrc = kdst; kdst = tk;
        uint8_t* tt = tsrc; tsrc = tdst; tdst = tt;
        __syncthreads();
    }
    // after 4 passes the sorted data is back in (ksrc, tsrc) == (key0, tp0)
    int32_t* tpcum = p.tpcum + seg0;
    double* env = p.env + seg0;
    const double eps = 2.220446049250313e-16;  // np.spacing(1)
    for (int t = 0; t < KODB_NUM_IOU_THRS; ++t) {
        // ---- c1. cumulative TP
        int carry = 0;
        for (int start = 0; start < nd; start += AP_THREADS) {
            const int i = start + tid;
            const int v = (i < nd) ? ((tsrc[i] >> t) & 1) : 0;
            int tot;
            const int ex = block_excl_scan(v, s.scan_tmp, &tot);
            if (i < nd) tpcum[i] = carry + ex + v;
            carry += tot;
        }
        __syncthreads();
        // ---- c2. precision + suffix max, from the last chunk backwards
        double cmax = -1.0;
        const int nchunks = (nd + AP_THREADS - 1) / AP_THREADS;
        for (int ch = nchunks - 1; ch >= 0; --ch) {
            const int i = ch * AP_THREADS + tid;
            double pr = -1.0;
            if (i < nd) {
                const double tp = (double)tpcum[i];
                const double fp = (double)(i + 1 - tpcum[i]);
                pr = tp / (fp + tp + eps);
            }
            const double e = block_suffix_max(pr, s.dscan, cmax);
            if (i < nd) env[i] = e;
            cmax = s.dscan[AP_WARPS];
            __syncthreads();
        }
        // ---- c3. 101 recall thresholds (np.linspace(0,1,101)), searchsorted(side='left')
        if (tid < 101) {
            const double R = tid == 100 ? 1.0 : (double)tid * 0.01;
            int lo = 0, hi = nd;  // first i with tpcum[i]/npig >= R
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if ((double)tpcum[mid] / (double)npig >= R) hi = mid; else lo = mid + 1;
            }
            s.q[tid] = lo < nd ? env[lo] : 0.0;
        }
        __syncthreads();
        if (tid == 0) {
            double sum = 0.0;
            for (int r = 0; r < 101; ++r) sum += s.q[r];
            p.ap[(int64_t)t * p.C + k] = sum / 101.0;
        }
        __syncthreads();
    }
}

__global__ void ap_reset_kernel(int* cls_count, int C) {
    if ((int)threadIdx.x < C) cls_count[threadIdx.x] = 0;
}

static size_t ap_align(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace kodb

extern "C" int64_t kodb_ap_workspace_bytes(int64_t num_images, int32_t max_det, int32_t num_classes) {
    using namespace kodb;
    const size_t N = (size_t)num_images * max_det;
    return (int64_t)(ap_align((size_t)num_classes * 4) + 2 * ap_align(N * 4) + 2 * ap_align(N) + ap_align(N * 4) +
                     ap_align(N * 8));
}

extern "C" int kodb_ap_accumulate(const float* rec_score, const int32_t* rec_cls, const uint8_t* rec_tp,
                                  const int32_t* counts, int64_t num_images, int32_t max_det, int32_t num_classes,
                                  const int64_t* gt_count, double* ap, void* workspace, int64_t workspace_bytes,
                                  void* stream) {
    using namespace kodb;
    if (!rec_score || !rec_cls || !rec_tp || !counts || !gt_count || !ap || !workspace || num_images < 0 || max_det < 1 ||
        num_classes < 1 || num_classes > KODB_MAX_CLASSES)
        return KODB_ERR_INVALID_ARGUMENT;
    if (workspace_bytes < kodb_ap_workspace_bytes(num_images, max_det, num_classes)) return KODB_ERR_CAPACITY;
    if ((double)num_images * max_det >= 2147483647.0) return KODB_ERR_UNSUPPORTED;
    const size_t N = (size_t)num_images * max_det;
    ApParams p;
    p.rec_score = rec_score; p.rec_cls = rec_cls; p.rec_tp = rec_tp; p.counts = counts;
    p.num_images = num_images; p.max_det = max_det; p.C = num_classes; p.gt_count = gt_count; p.ap = ap;
    char* w = reinterpret_cast<char*>(workspace);
    p.cls_count = reinterpret_cast<int*>(w); w += ap_align((size_t)num_classes * 4);
    p.key[0] = reinterpret_cast<uint32_t*>(w); w += ap_align(N * 4);
    p.key[1] = reinterpret_cast<uint32_t*>(w); w += ap_align(N * 4);
    p.tpb[0] = reinterpret_cast<uint8_t*>(w); w += ap_align(N);
    p.tpb[1] = reinterpret_cast<uint8_t*>(w); w += ap_align(N);
    p.tpcum = reinterpret_cast<int32_t*>(w); w += ap_align(N * 4);
    p.env = reinterpret_cast<double*>(w);
    cudaStream_t st = (cudaStream_t)stream;
    if (N > 0) {
        int blocks = (int)((N + 255) / 256);
        if (blocks > 148 * 16) blocks = 148 * 16;
        ap_count_kernel<<<blocks, 256, 0, st>>>(p);
        KODB_CHECK_LAUNCH();
    }
    ap_class_kernel<<<num_classes, AP_THREADS, 0, st>>>(p);
    KODB_CHECK_LAUNCH();
    ap_reset_kernel<<<1, KODB_MAX_CLASSES, 0, st>>>(p.cls_count, num_classes);
    KODB_CHECK_LAUNCH();
    return KODB_OK;
}
