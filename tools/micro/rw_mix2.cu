// Micro-benchmark 2: how do a latency-bound read stream and a write stream share the HBM, as a function of how the writer
// is throttled (CTA count) and of the store flavour (st / st.cs / st.cg / st.wt / TMA bulk store)?
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
__global__ void rd(const float4* p, size_t n4, float* out) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    float a = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = __ldcs(p + i);
        a += v.x + v.y + v.z + v.w;
    }
    if (a == 123.456f) out[0] = a;
}
template <int MODE>
__global__ void wr(float4* p, size_t n4) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        if (MODE == 0) p[i] = z;
        else if (MODE == 1) __stcs(p + i, z);
        else if (MODE == 2) __stcg(p + i, z);
        else __stwt(p + i, z);
    }
}
// TMA bulk stores from a zeroed 16 KB block of shared memory; one issuing thread per CTA
__global__ void wr_bulk(float4* p, size_t n4) {
    __shared__ __align__(128) float4 zeros[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) zeros[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t src = (uint32_t)__cvta_generic_to_shared(zeros);
        int pending = 0;
        for (size_t c = blockIdx.x; c * 1024 < n4; c += gridDim.x) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p + c * 1024), "r"(src), "r"(16384) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (++pending >= 8) { asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory"); pending = 4; }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}
int main() {
    const size_t bytes = 96ull << 20, n4 = bytes / 16;
    const int NB = 4;
    float4 *rb[NB], *wb[NB];
    float* out;
    cudaMalloc(&out, 4);
    for (int i = 0; i < NB; ++i) { cudaMalloc(&rb[i], bytes); cudaMalloc(&wb[i], bytes); cudaMemset(rb[i], 0, bytes); }
    cudaStream_t s1, s2;
    cudaStreamCreate(&s1); cudaStreamCreate(&s2);
    cudaEvent_t e0, e1, j;
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&j);
    const int R = 30;
    auto time = [&](int gr, auto launch_w) {
        float ms = 0;
        for (int rep = 0; rep < 2; ++rep) {
            cudaDeviceSynchronize();
            cudaEventRecord(e0, s1);
            cudaStreamWaitEvent(s2, e0);
            for (int i = 0; i < R; ++i) {
                if (gr) rd<<<gr, 256, 0, s1>>>(rb[i % NB], n4, out);
                launch_w(wb[i % NB]);
            }
            cudaEventRecord(j, s2);
            cudaStreamWaitEvent(s1, j);
            cudaEventRecord(e1, s1);
            cudaDeviceSynchronize();
            cudaEventElapsedTime(&ms, e0, e1);
        }
        return ms / R * 1e3;
    };
    auto none = [&](float4*) {};
    for (int gr : {296, 592}) {
        const float tr = time(gr, none);
        printf("reader grid %d alone: %6.2f us (%5.0f GB/s)\n", gr, tr, bytes / tr / 1e3);
        auto report = [&](const char* name, int gw, auto lw) {
            const float tw = time(0, lw), tb = time(gr, lw);
            printf("  writer %-10s grid %4d: alone %6.2f us (%5.0f GB/s)  both %6.2f us (%5.0f GB/s total)  sum %6.2f  max %6.2f\n", name, gw, tw,
                   bytes / tw / 1e3, tb, 2 * bytes / tb / 1e3, tr + tw, tr > tw ? tr : tw);
        };
        for (int gw : {37, 74, 148, 296, 1184}) {
            report("st", gw, [&](float4* p) { wr<0><<<gw, 256, 0, s2>>>(p, n4); });
            report("st.cs", gw, [&](float4* p) { wr<1><<<gw, 256, 0, s2>>>(p, n4); });
        }
        report("st.cg", 296, [&](float4* p) { wr<2><<<296, 256, 0, s2>>>(p, n4); });
        report("st.wt", 296, [&](float4* p) { wr<3><<<296, 256, 0, s2>>>(p, n4); });
        for (int gw : {74, 148, 296}) report("tma-bulk", gw, [&](float4* p) { wr_bulk<<<gw, 128, 0, s2>>>(p, n4); });
        report("memsetAsync", 0, [&](float4* p) { cudaMemsetAsync(p, 0, bytes, s2); });
        report("memsetD32", 0, [&](float4* p) { cudaMemsetAsync(p, 0, bytes / 2, s2); cudaMemsetAsync((char*)p + bytes / 2, 0, bytes / 2, s2); });
    }
    return 0;
}
