// kodb_head_forward / kodb_head_detections — the Detect head of the reference (kod/nn/heads/yolov5.py:12-178): per pyramid
// level three 1x1 convolutions (box: A*4, objectness: A, class: A*C output channels, with bias) of the same (B, K, H, W)
// feature map, each followed by Rearrange("b (a p) h w -> b a h w p") -- i.e. three planar [B][A*P][H][W] tensors
//  -- and, on the eval side, get_detections (exp.py:70-102) of those outputs.
//
// Here: ONE persistent tcgen05 kernel for all three levels.  Per level the three convolutions are one GEMM
//     D[cell][n] = sum_k X[b][k][cell] * W[n][k]          cell = 0..H*W-1 (M), n = 0..A*(5+C)-1 (N), k = 0..K-1
// with the feature map read ONCE (the reference reads it three times) straight from its NCHW layout:
//   * A operand = X tile, 128 cells x 32 channels, "MN-major" (cells contiguous): TMA boxes of 32 cells x 32 channels with
//     the 128-byte swizzle / 32-byte atom -- the only legal shared-memory layout of an MN-major tf32 operand (UMMA layout
//     type SWIZZLE_128B_BASE32B: atoms of 4 channel rows x 128 B) -- no transposition anywhere;
//   * B operand = the concatenated weights [n][k] (K-major, 128-byte swizzle), a 32-channel slab per stage, rows beyond
//     A*(5+C) zero-filled by the TMA bounds check;
//   * tf32 MMA (fp32 operands, fp32 accumulation -- the precision class of torch's default conv policy
//     torch.backends.cudnn.allow_tf32 = True), accumulators in TMEM: 128 lanes (cells) x N columns, double buffered;
//   * warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread) + TMEM owner, warps 2..9 =
//     epilogue, two per TMEM lane quadrant (tcgen05.ld 32x32b: thread = cell, registers = channels):
//       planar  (<false>): + bias; one 16-byte shared-memory table read gives a channel's plane pointer; 32 lanes store 128
//                contiguous bytes of ONE output plane -- precisely the planar layout the detection tail wants;
//       decode  (<true>):  + bias, sigmoid / grid / anchor decode, rows of (x1, y1, x2, y2, obj, cls...) staged per warp and
//                written as contiguous runs of the (B, AC, 5+C) detection tensor; the logits never reach memory;
//   * scheduling units of equal load (16 k-blocks), heavy levels first, static round-robin over two CTAs per SM.
// HBM-bound by construction (AI ~ 17 flop/B at K=128, N=45): 4*K bytes in + 4*N bytes out per cell.
// Levels the TMA / tcgen05 path cannot take (H*W % 4, in_channels % 32, > 256 output channels) run on head_generic_kernel.
#include <stdlib.h>

#include <mutex>

#include "kodb_common.cuh"
#include "tma_common.cuh"

namespace kodb {

constexpr int HD_TILE_M = 128;   // cells per tile (UMMA M)
constexpr int HD_KB = 32;        // channels per pipeline stage (one 128-byte swizzle row of fp32)
constexpr int HD_UMMA_K = 8;     // tf32: 32 bytes of K per instruction
constexpr int HD_STAGES = 4;
constexpr int HD_EPI_WARPS = 8;  // two per TMEM lane quadrant: even / odd 16-column chunks
constexpr int HD_THREADS = 32 * (2 + HD_EPI_WARPS);  // warps: TMA, MMA, 8 x epilogue
constexpr int HD_UNIT_KB = 16;   // scheduling unit: tiles worth 16 k-blocks of loads (4 tiles at K=128, 2 at 256, 1 at >= 512)
constexpr int HD_A_STAGE_BYTES = HD_TILE_M * HD_KB * 4;  // 16 KB: four slabs of 32 cells x 32 channels
constexpr int HD_A_SBO = 512;    // A operand: bytes between two 4-channel swizzle atoms
constexpr int HD_A_KADV = 1024;  // A operand: bytes per MMA (8 channels)

// where one output channel of a level goes: base = its plane in image 0, images img_stride elements apart
struct __align__(16) HeadChannel {
    float* base;
    uint32_t img_stride;
    float bias;
};

struct HeadLevelParams {
    int K, HW, B;
    int tiles_per_image, tiles;
    int unit_begin, per_unit;  // this level's scheduling units start at unit_begin; per_unit tiles each
    int n1, n2, n3;      // channel boundaries: box [0, n1), objectness [n1, n2), class [n2, n3)
    float* out[3];       // planar [B][channels][HW]
    const float* bias[3];
    // decode epilogue (kodb_head_detections): geometry of the level in the (B, AC, 5+C) detection tensor
    int row_base, W;
    float stride;
    float anchor_px[KODB_MAX_ANCHORS][2];
};

struct HeadParams {
    CUtensorMap tma_x[KODB_NUM_LEVELS];  // (cell, channel, image) fp32, box 32 x 32 x 1, 128B swizzle
    CUtensorMap tma_w[KODB_NUM_LEVELS];  // (channel, n) fp32, box 32 x n_pad, 128B swizzle
    HeadLevelParams lv[KODB_NUM_LEVELS];
    int n_pad;           // UMMA N: A*(5+C) rounded up to 16
    int acc_cols;        // TMEM columns per accumulator (power of two >= n_pad)
    int b_stage_bytes;   // n_pad * 128
    int total_units;
    int order[KODB_NUM_LEVELS];  // levels by descending K: the heavy tiles lead the static round-robin
    // decode epilogue: instead of the planar logits the epilogue writes get_detections' rows (exp.py:70-102)
    float* det;          // (B, rows_total, 5+C) or null
    int rows_total, A, C;
};

// ---- PTX wrappers (mbarrier / TMA: tma_common.cuh) ---------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem desc] * B[smem desc], tf32 inputs, fp32 accumulate, issued by ONE thread for the CTA
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all MMAs issued so far by this thread have completed -> one arrival on the mbarrier (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&v)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld1(uint32_t taddr, uint32_t& v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
}
// Decode epilogue arithmetic: the logits come out of a tf32 GEMM (|error| ~ 2e-3), so the eval tail's bit-exact sigmoid
// (libm expf + IEEE division, ~28 instructions) would buy nothing here; ex2.approx + rcp.approx (~1e-6 relative) keeps the
// epilogue off the critical path.  Same expression tree as decode_box (kodb_common.cuh) otherwise.
__device__ __forceinline__ float sigmoid_epi(float x) { return rcp_approx(1.0f + ex2_approx(-1.44269502162933349609375f * x)); }
__device__ __forceinline__ float4 decode_box_epi(float p0, float p1, float p2, float p3, float gx, float gy, float stride,
                                                 float aw_px, float ah_px) {
    const float cx = ((sigmoid_epi(p0) * 2.0f + gx) - 0.5f) * stride;
    const float cy = ((sigmoid_epi(p1) * 2.0f + gy) - 0.5f) * stride;
    const float sw = sigmoid_epi(p2) * 2.0f, sh = sigmoid_epi(p3) * 2.0f;
    const float w = (sw * sw) * aw_px, h = (sh * sh) * ah_px;
    const float hw = 0.5f * w, hh = 0.5f * h;
    return make_float4(cx - hw, cy - hh, cx + hw, cy + hh);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address, leading / stride byte offsets in
// 16-byte units, version 1 (sm_100), layout type SWIZZLE_128B
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;  // version
    d |= (uint64_t)layout << 61;  // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
    return d;
}

// =========================================================================================================
// scheduling units in the order levels are listed in p.order (descending K); unit u -> level, tiles [t0, t1)
__device__ __forceinline__ int unit_level(const HeadParams& p, int u) {
    return u >= p.lv[p.order[2]].unit_begin ? p.order[2] : (u >= p.lv[p.order[1]].unit_begin ? p.order[1] : p.order[0]);
}

template <bool DECODE>
__global__ void __launch_bounds__(HD_THREADS, 2) head_forward_kernel(const __grid_constant__ HeadParams p) {
    extern __shared__ __align__(1024) uint8_t hd_smem_raw[];
    // 1024-byte alignment of every operand stage (128B swizzle atoms are 1 KB)
    // (offset arithmetic on the shared array, not an integer round trip: keeps the shared address space -> LDS / STS)
    uint8_t* smem = hd_smem_raw + ((1024u - (smem_u32(hd_smem_raw) & 1023u)) & 1023u);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + HD_STAGES * HD_A_STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + HD_STAGES * p.b_stage_bytes);
    uint64_t* full_bar = bars;                    // [HD_STAGES]  TMA -> MMA
    uint64_t* empty_bar = bars + HD_STAGES;       // [HD_STAGES]  MMA -> TMA
    uint64_t* tfull_bar = bars + 2 * HD_STAGES;   // [2]          MMA -> epilogue
    uint64_t* tempty_bar = bars + 2 * HD_STAGES + 2;  // [2]      epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * HD_STAGES + 4);
    HeadChannel* chan = reinterpret_cast<HeadChannel*>(bars + 2 * HD_STAGES + 6);  // [KODB_NUM_LEVELS][n_pad]
    // decode epilogue: per epilogue warp a 32 rows x 16 columns staging tile (pitch 17: conflict-free both ways)
    float* det_stage = reinterpret_cast<float*>(chan + KODB_NUM_LEVELS * p.n_pad);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int l = 0; l < KODB_NUM_LEVELS; ++l) {
            tma_prefetch_desc(&p.tma_x[l]);
            tma_prefetch_desc(&p.tma_w[l]);
        }
        for (int s = 0; s < HD_STAGES; ++s) {
            mbar_init(full_bar + s, 1);
            mbar_init(empty_bar + s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar + a, 1);
            mbar_init(tempty_bar + a, HD_EPI_WARPS);  // one arrival per epilogue warp
        }
        mbar_fence_init();
    }
    if (warp == 1) {  // TMEM: two accumulators of acc_cols columns; this warp owns allocation and release
        const uint32_t ncols = 2u * (uint32_t)p.acc_cols;
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // the output channel table: the epilogue's whole address computation is one 16-byte shared-memory read per store
    for (int i = threadIdx.x; i < KODB_NUM_LEVELS * p.n_pad; i += HD_THREADS) {
        const int l = i / p.n_pad, n = i - l * p.n_pad;
        const HeadLevelParams& L = p.lv[l];
        HeadChannel c{nullptr, 0u, 0.f};
        if (L.tiles > 0 && n < L.n3) {
            const int ti = n < L.n1 ? 0 : (n < L.n2 ? 1 : 2);
            const int nb = ti == 0 ? 0 : (ti == 1 ? L.n1 : L.n2);
            const int ch = ti == 0 ? L.n1 : (ti == 1 ? L.n2 - L.n1 : L.n3 - L.n2);
            c.base = L.out[ti] ? L.out[ti] + (size_t)(n - nb) * L.HW : nullptr;
            c.img_stride = (uint32_t)(ch * L.HW);
            c.bias = __ldg(L.bias[ti] + (n - nb));
        }
        chan[i] = c;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t tx = (uint32_t)(HD_A_STAGE_BYTES + p.b_stage_bytes);
            for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
                const int l = unit_level(p, u);
                const HeadLevelParams& L = p.lv[l];
                const int t0 = (u - L.unit_begin) * L.per_unit, t1 = min(t0 + L.per_unit, L.tiles);
                for (int tt = t0; tt < t1; ++tt) {
                    const int b = tt / L.tiles_per_image, c0 = (tt - b * L.tiles_per_image) * HD_TILE_M;
                    for (int kb = 0; kb < L.K / HD_KB; ++kb) {
                        mbar_wait(empty_bar + stage, phase ^ 1u);
                        mbar_expect_tx(full_bar + stage, tx);
                        uint8_t* a_dst = smem_a + stage * HD_A_STAGE_BYTES;
#pragma unroll
                        for (int j = 0; j < HD_TILE_M / 32; ++j)  // slab j: cells c0 + 32j .. + 31 (zero-filled beyond H*W)
                            tma_load_3d(a_dst + j * (HD_KB * 128), &p.tma_x[l], full_bar + stage, c0 + 32 * j, kb * HD_KB, b);
                        tma_load_2d(smem_b + stage * p.b_stage_bytes, &p.tma_w[l], full_bar + stage, kb * HD_KB, 0);
                        if (++stage == HD_STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            // instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B tf32, A MN-major, B K-major, N, M = 128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | ((uint32_t)(p.n_pad >> 3) << 17) |
                                   ((uint32_t)(HD_TILE_M >> 4) << 24);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
                const int l = unit_level(p, u);
                const HeadLevelParams& L = p.lv[l];
                const int t0 = (u - L.unit_begin) * L.per_unit, t1 = min(t0 + L.per_unit, L.tiles);
                const int nkb = L.K / HD_KB;
                for (int tt = t0; tt < t1; ++tt) {
                    mbar_wait(tempty_bar + acc, acc_phase ^ 1u);  // the epilogue has drained this accumulator
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.acc_cols);
                    for (int kb = 0; kb < nkb; ++kb) {
                        mbar_wait(full_bar + stage, phase);
                        tc_fence_after();
                        const uint32_t a_addr = smem_u32(smem_a + stage * HD_A_STAGE_BYTES);
                        const uint32_t b_addr = smem_u32(smem_b + stage * p.b_stage_bytes);
#pragma unroll
                        for (int k = 0; k < HD_KB / HD_UMMA_K; ++k) {
                            // A (MN-major, 128B swizzle / 32B atom): atoms of 4 channel rows x 128 B; one instruction takes 8
                            //   channels = 2 atoms HD_A_SBO apart; the four 32-cell slabs are HD_KB*128 bytes apart (LBO).
                            // B (K-major, 128B swizzle): rows of 128 bytes, 8-row groups 1 KB apart (SBO); 8 channels = 32
                            //   bytes along the row.
                            const uint64_t a_desc = umma_smem_desc(a_addr + k * HD_A_KADV, HD_KB * 128, HD_A_SBO, 1);
                            const uint64_t b_desc = umma_smem_desc(b_addr + k * 32, 16, 1024, 2);
                            umma_tf32(d_tmem, a_desc, b_desc, idesc, (kb | k) != 0 ? 1u : 0u);
                        }
                        umma_commit(empty_bar + stage);  // the stage's operands may be overwritten once these MMAs are done
                        if (kb == nkb - 1) umma_commit(tfull_bar + acc);  // ... and the accumulator is complete
                        if (++stage == HD_STAGES) { stage = 0; phase ^= 1u; }
                    }
                    if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
                }
            }
        }
    } else {
        // ===== epilogue: TMEM -> registers -> (+ bias) -> planar global stores =====
        const int q = warp & 3;             // the TMEM lane quadrant this warp may read
        const int half = (warp - 2) >> 2;   // 16-column chunks: even ones to the first warp of the quadrant, odd to the second
        const int row = 32 * q + lane;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
            const int l = unit_level(p, u);
            const HeadLevelParams& L = p.lv[l];
            const int t0 = (u - L.unit_begin) * L.per_unit, t1 = min(t0 + L.per_unit, L.tiles);
            const HeadChannel* tab = chan + l * p.n_pad;
            const int n3 = L.n3;
            for (int tt = t0; tt < t1; ++tt) {
                const int b = tt / L.tiles_per_image, c0 = (tt - b * L.tiles_per_image) * HD_TILE_M;
                const int cell = c0 + row;
                const bool live = cell < L.HW;
                mbar_wait(tfull_bar + acc, acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(acc * p.acc_cols);
                if constexpr (DECODE) {
                    // ---- decode epilogue: rows of get_detections (x1, y1, x2, y2, sigmoid(obj), sigmoid(cls...)) ----
                    // anchors alternate between the two warps of a lane quadrant; a row's 5+C columns go out in chunks of
                    // 16 through the warp's staging tile, so that global stores are 64-byte runs of consecutive rows
                    const int A = p.A, C = p.C, cols = 5 + C;
                    float* st = det_stage + (warp - 2) * (32 * 17);
                    const int gy = cell / L.W, gx = cell - gy * L.W;
                    const int nrow = min(32, L.HW - (c0 + 32 * q));  // live rows of this warp (<= 0: none)
                    // work items = (anchor, chunk of 16 columns), alternating between the two warps of the quadrant (measured:
                    // 8-column chunks balance 10-class rows 3:3 instead of 2:1 but cost more than they save, 114 -> 126 us)
                    constexpr int Wc = 16;
                    const int nch = (cols + Wc - 1) / Wc;
                    for (int it = half; it < A * nch; it += 2) {
                        const int a = it / nch, j0 = (it - a * nch) * Wc;
                        const int cls0 = 5 * A + a * C;  // first class column of the anchor in the accumulator
                        float* dst = p.det + ((size_t)b * p.rows_total + L.row_base + (size_t)a * L.HW + c0 + 32 * q) * cols;
                        const int nc = min(Wc, cols - j0);
                        float val[16];
                        const int col = j0 == 0 ? cls0 : cls0 + j0 - 5;  // first class column of this chunk
                        const int ncls = j0 == 0 ? nc - 5 : nc;
                        uint32_t v[16];
                        if (col + 16 <= p.acc_cols) {
                            tmem_ld16(taddr + col, v);
                        } else {  // the 16-column read would leave the accumulator: column by column
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                v[i] = 0u;
                                if (i < ncls) tmem_ld1(taddr + col + i, v[i]);
                            }
                        }
                        if (j0 == 0) {
                            uint32_t vb[4], vo;
                            tmem_ld4(taddr + 4 * a, vb);
                            tmem_ld1(taddr + 4 * A + a, vo);
                            tmem_ld_wait();
                            const float4 bx = decode_box_epi(__uint_as_float(vb[0]) + tab[4 * a].bias, __uint_as_float(vb[1]) + tab[4 * a + 1].bias,
                                                             __uint_as_float(vb[2]) + tab[4 * a + 2].bias, __uint_as_float(vb[3]) + tab[4 * a + 3].bias,
                                                             (float)gx, (float)gy, L.stride, L.anchor_px[a][0], L.anchor_px[a][1]);
                            val[0] = bx.x; val[1] = bx.y; val[2] = bx.z; val[3] = bx.w;
                            val[4] = sigmoid_epi(__uint_as_float(vo) + tab[4 * A + a].bias);
#pragma unroll
                            for (int i = 0; i < 11; ++i)
                                val[5 + i] = i < ncls ? sigmoid_epi(__uint_as_float(v[i]) + tab[min(col + i, n3 - 1)].bias) : 0.0f;
                        } else {
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                val[i] = i < ncls ? sigmoid_epi(__uint_as_float(v[i]) + tab[min(col + i, n3 - 1)].bias) : 0.0f;
                        }
                        __syncwarp();  // the previous chunk's readers of the staging tile are done
                        // a row that fits one chunk (5 + C <= 16): staged at pitch = row length, so the warp's 32 rows are one
                        // contiguous run in shared AND global memory -- a flat, fully coalesced copy without index arithmetic
                        const bool whole = nc == cols;
                        const int pitch = whole ? cols : 17;
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (i < nc) st[lane * pitch + i] = val[i];
                        __syncwarp();
                        if (whole) {
                            for (int i = lane; i < nrow * cols; i += 32) dst[i] = st[i];
                        } else if (nc == 16) {
                            for (int i = lane; i < nrow * 16; i += 32) dst[(size_t)(i >> 4) * cols + j0 + (i & 15)] = st[(i >> 4) * 17 + (i & 15)];
                        } else {
                            for (int i = lane; i < nrow * nc; i += 32) {
                                const int rr = i / nc, cc = i - rr * nc;
                                dst[(size_t)rr * cols + j0 + cc] = st[rr * 17 + cc];
                            }
                        }
                    }
                    // every tcgen05.ld of this warp has completed (wait::ld above): hand the accumulator back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty_bar + acc);
                    if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
                    continue;
                }
                int n0 = 16 * half;
                if (n0 >= n3) {  // nothing to read for this warp: hand the accumulator back at once
                    tc_fence_before();
                    if (lane == 0) mbar_arrive(tempty_bar + acc);
                }
                for (; n0 < n3; n0 += 32) {
                    uint32_t v[16];
                    tmem_ld16(taddr + n0, v);
                    tmem_ld_wait();
                    if (n0 + 32 >= n3) {
                        // this warp's last chunk is in registers: hand the accumulator back before the stores go out
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(tempty_bar + acc);
                    }
                    if (live) {
                        if (n0 + 16 <= n3) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const HeadChannel c = tab[n0 + i];  // one broadcast 16-byte read
                                // 32 lanes -> 32 consecutive cells of one output plane: a 128-byte store
                                c.base[(size_t)b * c.img_stride + cell] = __uint_as_float(v[i]) + c.bias;
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                if (n0 + i < n3) {
                                    const HeadChannel c = tab[n0 + i];
                                    c.base[(size_t)b * c.img_stride + cell] = __uint_as_float(v[i]) + c.bias;
                                }
                            }
                        }
                    }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        const uint32_t ncols = 2u * (uint32_t)p.acc_cols;
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
    }
}

// =========================================================================================================
// Levels the TMA / tcgen05 path does not take (H*W not a multiple of 4 -- the 13x13 level of a 416 input --, in_channels
// not a multiple of 32, more than 256 output channels, misaligned tensors): a plain SIMT kernel, thread = cell, 16 output
// channels per pass with their weight rows staged in shared memory, fp32 multiply-add (exact fp32, no tf32 rounding).
// These levels are small; the kernel only has to be correct and coalesced.
constexpr int HG_THREADS = 128;
constexpr int HG_NCH = 16;     // output channels per pass
constexpr int HG_KCH = 128;    // input channels staged per round (HG_NCH * HG_KCH floats = 8 KB)

struct HeadGenericParams {
    const float* x;
    const float* w[3];     // (rows_i, K)
    const float* bias[3];
    float* out[3];
    int rows[3];           // A*4, A, A*C
    int K, HW;
};

__global__ void __launch_bounds__(HG_THREADS) head_generic_kernel(const __grid_constant__ HeadGenericParams p) {
    __shared__ float sw[HG_NCH][HG_KCH];
    const int b = blockIdx.y, cell = blockIdx.x * HG_THREADS + threadIdx.x;
    const bool live = cell < p.HW;
    const float* xb = p.x + (size_t)b * p.K * p.HW + (live ? cell : 0);
    for (int ti = 0; ti < 3; ++ti) {
        for (int n0 = 0; n0 < p.rows[ti]; n0 += HG_NCH) {
            const int nn = min(HG_NCH, p.rows[ti] - n0);
            float acc[HG_NCH];
#pragma unroll
            for (int u = 0; u < HG_NCH; ++u) acc[u] = 0.0f;
            for (int k0 = 0; k0 < p.K; k0 += HG_KCH) {
                const int kk = min(HG_KCH, p.K - k0);
                __syncthreads();  // the previous round's readers are done
                for (int i = threadIdx.x; i < HG_NCH * HG_KCH; i += HG_THREADS) {
                    const int u = i / HG_KCH, k = i - u * HG_KCH;
                    sw[u][k] = (u < nn && k < kk) ? __ldg(p.w[ti] + (size_t)(n0 + u) * p.K + k0 + k) : 0.0f;
                }
                __syncthreads();
                if (live) {
                    for (int k = 0; k < kk; ++k) {
                        const float xv = __ldg(xb + (size_t)(k0 + k) * p.HW);  // 32 lanes: 128 contiguous bytes
#pragma unroll
                        for (int u = 0; u < HG_NCH; ++u) acc[u] = fmaf(xv, sw[u][k], acc[u]);
                    }
                }
            }
            if (live) {
#pragma unroll
                for (int u = 0; u < HG_NCH; ++u)
                    if (u < nn) p.out[ti][((size_t)b * p.rows[ti] + n0 + u) * p.HW + cell] = acc[u] + __ldg(p.bias[ti] + n0 + u);
            }
        }
    }
}

// ---- host -----------------------------------------------------------------------------------------------------------
}  // namespace kodb

// det == nullptr: planar logits (Yolov5Head.forward); det != nullptr: decoded detections (get_detections of those logits),
// level_cfg_h gives each level's stride / anchors / grid
static int head_launch(const kodb_head_conv_level* levels_h, int32_t num_levels, int32_t batch, int32_t num_anchors,
                       int32_t num_classes, const kodb_level_cfg* level_cfg_h, float* det, void* stream) {
    using namespace kodb;
    if (!levels_h || num_levels < 1 || num_levels > KODB_NUM_LEVELS || batch < 1 || num_anchors < 1 || num_anchors > KODB_MAX_ANCHORS || num_classes < 1 ||
        num_classes > KODB_MAX_CLASSES)
        return KODB_ERR_INVALID_ARGUMENT;
    const int n3 = num_anchors * (5 + num_classes);
    const int n_pad = (n3 + 15) / 16 * 16;
    EncodeTiledFn encode = tensormap_encoder();
    cudaStream_t st = (cudaStream_t)stream;
    HeadParams p;
    p.n_pad = n_pad;
    p.acc_cols = 32;
    while (p.acc_cols < n_pad) p.acc_cols *= 2;
    p.b_stage_bytes = n_pad * 128;
    // A operand (MN-major tf32): the only legal shared-memory layout is the 128-byte swizzle with a 32-byte atom
    // (cutlass sm100_common.inl:92; UMMA layout type SWIZZLE_128B_BASE32B)
    int tiles = 0, n_tc = 0, det_rows = 0;
    for (int li = 0; li < num_levels; ++li) {
        const kodb_head_conv_level& h = levels_h[li];
        if (!h.x || !h.weight || !h.bias_box || !h.bias_obj || !h.bias_cls || h.in_channels < 1 || h.H < 1 || h.W < 1)
            return KODB_ERR_INVALID_ARGUMENT;
        if (!det && (!h.out_box || !h.out_obj || !h.out_cls)) return KODB_ERR_INVALID_ARGUMENT;
        const int HW = h.H * h.W;
        // the tensor-core path needs: 32-channel k-blocks, 16-byte global strides for the TMA (H*W % 4), 16-byte aligned
        // bases, one UMMA tile in N (A*(5+C) <= 256: three anchors up to 80 classes)
        const bool tc = encode && n_pad <= 256 && h.in_channels % HD_KB == 0 && HW % 4 == 0 &&
                        !((reinterpret_cast<uintptr_t>(h.x) | reinterpret_cast<uintptr_t>(h.weight)) & 15u);
        if (!tc && det) return KODB_ERR_UNSUPPORTED;  // the decode epilogue exists on the tensor-core path only
        if (!tc) {  // small / odd level: the SIMT kernel
            HeadGenericParams g;
            g.x = h.x;
            g.K = h.in_channels;
            g.HW = HW;
            g.rows[0] = num_anchors * 4; g.rows[1] = num_anchors; g.rows[2] = num_anchors * num_classes;
            g.w[0] = h.weight;
            g.w[1] = g.w[0] + (size_t)g.rows[0] * g.K;
            g.w[2] = g.w[1] + (size_t)g.rows[1] * g.K;
            g.bias[0] = h.bias_box; g.bias[1] = h.bias_obj; g.bias[2] = h.bias_cls;
            g.out[0] = h.out_box; g.out[1] = h.out_obj; g.out[2] = h.out_cls;
            if (batch > 65535) return KODB_ERR_UNSUPPORTED;
            head_generic_kernel<<<dim3((HW + HG_THREADS - 1) / HG_THREADS, batch), HG_THREADS, 0, st>>>(g);
            KODB_CHECK_LAUNCH();
            continue;
        }
        const int l = n_tc++;
        HeadLevelParams& L = p.lv[l];
        L.K = h.in_channels;
        L.HW = HW;
        L.B = batch;
        L.tiles_per_image = (HW + HD_TILE_M - 1) / HD_TILE_M;
        L.tiles = L.tiles_per_image * batch;
        L.per_unit = HD_UNIT_KB / (L.K / HD_KB) > 0 ? HD_UNIT_KB / (L.K / HD_KB) : 1;
        tiles += L.tiles;
        L.n1 = num_anchors * 4;
        L.n2 = L.n1 + num_anchors;
        L.n3 = n3;
        L.out[0] = det ? nullptr : h.out_box; L.out[1] = det ? nullptr : h.out_obj; L.out[2] = det ? nullptr : h.out_cls;
        L.bias[0] = h.bias_box; L.bias[1] = h.bias_obj; L.bias[2] = h.bias_cls;
        L.row_base = det_rows;
        L.W = h.W;
        L.stride = 0.f;
        if (det) {
            const kodb_level_cfg& c = level_cfg_h[li];
            if (c.H != h.H || c.W != h.W || c.num_anchors != num_anchors) return KODB_ERR_INVALID_ARGUMENT;
            L.stride = (float)c.stride;
            for (int a = 0; a < num_anchors; ++a) { L.anchor_px[a][0] = c.anchor_px[a][0]; L.anchor_px[a][1] = c.anchor_px[a][1]; }
            det_rows += num_anchors * HW;
        }
        {   // feature map (B, K, H, W) as (cell, channel, image)
            const cuuint64_t dims[3] = {(cuuint64_t)HW, (cuuint64_t)h.in_channels, (cuuint64_t)batch};
            const cuuint64_t strides[2] = {(cuuint64_t)HW * 4, (cuuint64_t)HW * 4 * h.in_channels};
            const cuuint32_t box[3] = {32, (cuuint32_t)HD_KB, 1};
            const cuuint32_t estr[3] = {1, 1, 1};
            if (encode(&p.tma_x[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(h.x), dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return KODB_ERR_INVALID_ARGUMENT;
        }
        {   // concatenated weights (n, K) as (channel, n); rows beyond n3 are out of bounds -> zeros
            const cuuint64_t dims[2] = {(cuuint64_t)h.in_channels, (cuuint64_t)n3};
            const cuuint64_t strides[1] = {(cuuint64_t)h.in_channels * 4};
            const cuuint32_t box[2] = {(cuuint32_t)HD_KB, (cuuint32_t)n_pad};
            const cuuint32_t estr[2] = {1, 1};
            if (encode(&p.tma_w[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(h.weight), dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return KODB_ERR_INVALID_ARGUMENT;
        }
    }
    if (n_tc == 0) return KODB_OK;
    for (int l = n_tc; l < KODB_NUM_LEVELS; ++l) {  // absent levels own no tiles
        p.lv[l] = HeadLevelParams{};
        p.tma_x[l] = p.tma_x[0];
        p.tma_w[l] = p.tma_w[0];
    }
    // scheduling units, heavy levels first
    for (int l = 0; l < KODB_NUM_LEVELS; ++l) p.order[l] = l;
    for (int i = 0; i < KODB_NUM_LEVELS; ++i)
        for (int j = i + 1; j < KODB_NUM_LEVELS; ++j)
            if (p.lv[p.order[j]].K > p.lv[p.order[i]].K) { const int t = p.order[i]; p.order[i] = p.order[j]; p.order[j] = t; }
    int units = 0;
    for (int i = 0; i < KODB_NUM_LEVELS; ++i) {
        HeadLevelParams& L = p.lv[p.order[i]];
        L.unit_begin = units;
        if (L.tiles > 0) units += (L.tiles + L.per_unit - 1) / L.per_unit;
    }
    p.total_units = units;
    p.det = det;
    p.rows_total = det_rows;
    p.A = num_anchors;
    p.C = num_classes;
    const size_t smem = 1024 + (size_t)HD_STAGES * (HD_A_STAGE_BYTES + p.b_stage_bytes) + 256 +
                        (size_t)KODB_NUM_LEVELS * n_pad * sizeof(HeadChannel) + (det ? (size_t)HD_EPI_WARPS * 32 * 17 * sizeof(float) : 0);
    static std::mutex mu;
    static size_t smem_set[64] = {0};
    static int sms[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return KODB_ERR_UNSUPPORTED;
    {
        std::lock_guard<std::mutex> lock(mu);
        if (smem_set[dev] < smem) {
            if (cudaFuncSetAttribute(head_forward_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
                cudaFuncSetAttribute(head_forward_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
                return KODB_ERR_LAUNCH;
            smem_set[dev] = smem;
        }
        if (!sms[dev]) {
            int n = 0;
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
            sms[dev] = n > 0 ? n : 148;
        }
    }
    // persistent: as many CTAs as fit (shared memory: two per SM with the small class head, one with 80 classes)
    const int per_sm = smem <= 112 * 1024 ? 2 : 1;
    int grid = sms[dev] * per_sm;
    if (grid > units) grid = units;
    if (det) head_forward_kernel<true><<<grid, HD_THREADS, smem, st>>>(p);
    else head_forward_kernel<false><<<grid, HD_THREADS, smem, st>>>(p);
    KODB_CHECK_LAUNCH();
    return KODB_OK;
}

extern "C" int kodb_head_forward(const kodb_head_conv_level* levels_h, int32_t num_levels, int32_t batch, int32_t num_anchors,
                                 int32_t num_classes, void* stream) {
    return head_launch(levels_h, num_levels, batch, num_anchors, num_classes, nullptr, nullptr, stream);
}

extern "C" int kodb_head_detections(const kodb_head_conv_level* levels_h, const kodb_level_cfg* level_cfg_h, int32_t num_levels,
                                    int32_t batch, int32_t num_anchors, int32_t num_classes, float* detections, void* stream) {
    if (!level_cfg_h || !detections) return KODB_ERR_INVALID_ARGUMENT;
    return head_launch(levels_h, num_levels, batch, num_anchors, num_classes, level_cfg_h, detections, stream);
}
