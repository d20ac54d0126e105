/*
 * Oracle (TEST INFRASTRUCTURE, not product code): plain-C restatement of the greedy NMS that
 * `kod.core.nms.non_max_suppression` (kod/core/nms.py:66-73) delegates to `torchvision.ops.nms`.
 *
 * torchvision is a third-party dependency of the reference (requirements.txt pins
 * torchvision==0.15.2; this image ships 0.26.0) whose sources are not under /root/reference.
 * Published algorithm of its CPU kernel (torchvision/csrc/ops/cpu/nms_kernel.cpp), restated:
 *   - areas = (x2-x1)*(y2-y1) in the box dtype (fp32);
 *   - order = stable descending sort of the scores (ties: lower index first);
 *   - scan in that order; a box that is not suppressed is kept and suppresses every later box j with
 *       inter/(area_i + area_j - inter) > iou_threshold,
 *     inter = max(0, min(x2)-max(x1)) * max(0, min(y2)-max(y1)); the fp32 ratio is compared with the
 *     DOUBLE threshold (so ovr == (float)0.6 is suppressed at threshold 0.6);
 *   - no early exit: the full O(n^2) scan, exactly what the reference pays on the CPU.
 * Pinned against torchvision.ops.nms itself in tests/test_oracle.py and by tests/golden/nms_*.npz.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off so no FMA contraction changes the fp32 bits).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float s; int64_t i; } kod_pair;

static void merge_sort_desc(kod_pair* a, kod_pair* tmp, int64_t n) {
    /* bottom-up stable merge sort, descending by score; equal scores keep index order */
    for (int64_t w = 1; w < n; w *= 2) {
        for (int64_t lo = 0; lo < n; lo += 2 * w) {
            int64_t mid = lo + w < n ? lo + w : n;
            int64_t hi = lo + 2 * w < n ? lo + 2 * w : n;
            int64_t p = lo, q = mid, o = lo;
            while (p < mid && q < hi) {
                if (a[q].s > a[p].s) tmp[o++] = a[q++]; else tmp[o++] = a[p++];
            }
            while (p < mid) tmp[o++] = a[p++];
            while (q < hi) tmp[o++] = a[q++];
        }
        memcpy(a, tmp, (size_t)n * sizeof(kod_pair));
    }
}

/* boxes: n x 4 (x1,y1,x2,y2) fp32, already class-offset; keep: capacity n.  Returns #kept. */
int64_t kod_oracle_nms_greedy(const float* boxes, const float* scores, int64_t n,
                              double iou_threshold, int64_t* keep) {
    if (n <= 0) return 0;
    kod_pair* ord = (kod_pair*)malloc((size_t)n * sizeof(kod_pair));
    kod_pair* tmp = (kod_pair*)malloc((size_t)n * sizeof(kod_pair));
    float* area = (float*)malloc((size_t)n * sizeof(float));
    unsigned char* dead = (unsigned char*)calloc((size_t)n, 1);
    for (int64_t i = 0; i < n; ++i) {
        ord[i].s = scores[i];
        ord[i].i = i;
        area[i] = (boxes[4 * i + 2] - boxes[4 * i + 0]) * (boxes[4 * i + 3] - boxes[4 * i + 1]);
    }
    merge_sort_desc(ord, tmp, n);
    int64_t nk = 0;
    for (int64_t a = 0; a < n; ++a) {
        int64_t i = ord[a].i;
        if (dead[i]) continue;
        keep[nk++] = i;
        float ix1 = boxes[4 * i], iy1 = boxes[4 * i + 1], ix2 = boxes[4 * i + 2], iy2 = boxes[4 * i + 3];
        float ia = area[i];
        for (int64_t b = a + 1; b < n; ++b) {
            int64_t j = ord[b].i;
            if (dead[j]) continue;
            float xx1 = ix1 > boxes[4 * j] ? ix1 : boxes[4 * j];
            float yy1 = iy1 > boxes[4 * j + 1] ? iy1 : boxes[4 * j + 1];
            float xx2 = ix2 < boxes[4 * j + 2] ? ix2 : boxes[4 * j + 2];
            float yy2 = iy2 < boxes[4 * j + 3] ? iy2 : boxes[4 * j + 3];
            float w = xx2 - xx1; if (!(w > 0.0f)) w = 0.0f;
            float h = yy2 - yy1; if (!(h > 0.0f)) h = 0.0f;
            float inter = w * h;
            float ovr = inter / (ia + area[j] - inter);
            if ((double)ovr > iou_threshold) dead[j] = 1;
        }
    }
    free(ord); free(tmp); free(area); free(dead);
    return nk;
}

/*
 * pycocotools maskApi.c::bbIou restated for non-crowd boxes [x,y,w,h] in fp64
 * (third-party: pycocotools 2.0.x, not under /root/reference -> parity unpinned).
 * out[d*ng + g] = IoU(det d, gt g).
 */
void kod_oracle_bbiou(const double* dt, int64_t nd, const double* gt, int64_t ng, double* out) {
    for (int64_t g = 0; g < ng; ++g) {
        const double* G = gt + 4 * g;
        double ga = G[2] * G[3];
        for (int64_t d = 0; d < nd; ++d) {
            const double* D = dt + 4 * d;
            double da = D[2] * D[3];
            double w = (D[0] + D[2] < G[0] + G[2] ? D[0] + D[2] : G[0] + G[2]) - (D[0] > G[0] ? D[0] : G[0]);
            double h = (D[1] + D[3] < G[1] + G[3] ? D[1] + D[3] : G[1] + G[3]) - (D[1] > G[1] ? D[1] : G[1]);
            double o = 0.0;
            if (w > 0 && h > 0) {
                double i = w * h;
                double u = da + ga - i;
                o = i / u;
            }
            out[d * ng + g] = o;
        }
    }
}
