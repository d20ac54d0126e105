// kodb_decode — DefaultYolov5Experiment.get_detections (kod/lightning/experiments/yv5_baseline/exp.py:70-102):
// sigmoid/grid/anchor decode of the three head levels into one (B, AC, 5+C) tensor.
//
// The head is planar ([B][A][P][H][W], cells contiguous), the output is row-interleaved (5+C floats per
// anchor-cell), so each CTA transposes a tile of 128 anchor-cells through shared memory: coalesced plane
// reads (thread = cell), coalesced row writes (the tile is one contiguous span of the output).
// The fused eval tail (eval.cu) never materialises this tensor; this kernel exists for the reference's
// `get_detections` API and shares `decode_box` with the fused path, so both are bit-identical.
#include "kodb_common.cuh"

namespace kodb {

constexpr int DEC_TILE = 128;

struct DecodeParams {
    kodb_head_level head[KODB_NUM_LEVELS];
    kodb_level_cfg lv[KODB_NUM_LEVELS];
    int row_base[KODB_NUM_LEVELS + 1];
    int tiles_per_plane[KODB_NUM_LEVELS], tile_begin[KODB_NUM_LEVELS + 1];
    int B, C, pitch;  // pitch: padded smem row stride (odd -> conflict-free)
    float* out;
};

__global__ void __launch_bounds__(DEC_TILE) decode_kernel(const __grid_constant__ DecodeParams p) {
    extern __shared__ float tile[];
    const int b = blockIdx.y;
    int t = blockIdx.x;
    const int L = t >= p.tile_begin[2] ? 2 : (t >= p.tile_begin[1] ? 1 : 0);
    t -= p.tile_begin[L];
    const kodb_level_cfg& lv = p.lv[L];
    const kodb_head_level& hd = p.head[L];
    const int tpp = p.tiles_per_plane[L];
    const int a = t / tpp, c0 = (t - a * tpp) * DEC_TILE;
    const int HW = lv.H * lv.W;
    const int ncell = min(DEC_TILE, HW - c0);
    const int cols = 5 + p.C;
    const int tid = threadIdx.x;
    if (tid < ncell) {
        const int cell = c0 + tid;
        const int gy = cell / lv.W, gx = cell - gy * lv.W;
        const int64_t bo = (int64_t)b * hd.box_sb + (int64_t)a * hd.box_sa + cell;
        const float4 bx = decode_box(hd.box[bo], hd.box[bo + hd.box_sp], hd.box[bo + 2 * hd.box_sp], hd.box[bo + 3 * hd.box_sp],
                                     (float)gx, (float)gy, (float)lv.stride, lv.anchor_px[a][0], lv.anchor_px[a][1]);
        float* row = tile + tid * p.pitch;
        row[0] = bx.x; row[1] = bx.y; row[2] = bx.z; row[3] = bx.w;
        row[4] = sigmoid_eval(hd.obj[(int64_t)b * hd.obj_sb + (int64_t)a * hd.obj_sa + cell]);
        const int64_t co = (int64_t)b * hd.cls_sb + (int64_t)a * hd.cls_sa + cell;
        for (int j = 0; j < p.C; ++j) row[5 + j] = sigmoid_eval(hd.cls[co + (int64_t)j * hd.cls_sp]);
    }
    __syncthreads();
    const int64_t row0 = (int64_t)p.row_base[L] + (int64_t)a * HW + c0;
    float* dst = p.out + ((int64_t)b * p.row_base[KODB_NUM_LEVELS] + row0) * cols;
    if (p.pitch == cols) {  // odd row length: the tile is one contiguous run in shared memory too -- a flat copy
        for (int i = tid; i < ncell * cols; i += DEC_TILE) dst[i] = tile[i];
    } else {
        for (int i = tid; i < ncell * cols; i += DEC_TILE) {
            const int r = i / cols, c = i - r * cols;
            dst[i] = tile[r * p.pitch + c];
        }
    }
}

}  // namespace kodb

extern "C" int kodb_decode(const kodb_head_level* heads_h, const kodb_level_cfg* levels_h, int32_t batch,
                           int32_t num_classes, float* detections, void* stream) {
    using namespace kodb;
    if (!heads_h || !levels_h || !detections || batch < 1 || num_classes < 1 || num_classes > KODB_MAX_CLASSES)
        return KODB_ERR_INVALID_ARGUMENT;
    DecodeParams p;
    int rows = 0, tiles = 0;
    for (int l = 0; l < KODB_NUM_LEVELS; ++l) {
        p.head[l] = heads_h[l];
        p.lv[l] = levels_h[l];
        const int HW = levels_h[l].H * levels_h[l].W;
        if (HW <= 0 || levels_h[l].num_anchors < 1 || levels_h[l].num_anchors > KODB_MAX_ANCHORS)
            return KODB_ERR_INVALID_ARGUMENT;
        p.row_base[l] = rows;
        rows += levels_h[l].num_anchors * HW;
        p.tiles_per_plane[l] = (HW + DEC_TILE - 1) / DEC_TILE;
        p.tile_begin[l] = tiles;
        tiles += p.tiles_per_plane[l] * levels_h[l].num_anchors;
    }
    p.row_base[KODB_NUM_LEVELS] = rows;
    p.tile_begin[KODB_NUM_LEVELS] = tiles;
    p.B = batch;
    p.C = num_classes;
    p.pitch = (5 + num_classes) | 1;
    p.out = detections;
    const size_t smem = (size_t)DEC_TILE * p.pitch * sizeof(float);
    if (smem > 48 * 1024) cudaFuncSetAttribute(decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    decode_kernel<<<dim3(tiles, batch), DEC_TILE, smem, (cudaStream_t)stream>>>(p);
    KODB_CHECK_LAUNCH();
    return KODB_OK;
}
