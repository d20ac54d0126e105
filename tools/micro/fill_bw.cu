// Micro-benchmark: achievable HBM write bandwidth for a ~100 MB fill (rotating over buffers larger than L2).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void fill(float4* p, size_t n4) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        if (MODE == 0) p[i] = z;
        else if (MODE == 1) __stcs(p + i, z);
        else __stcg(p + i, z);
    }
}
template <int MODE>
__global__ void fill_chunk(float4* p, size_t n4) {  // each CTA owns a contiguous chunk
    const size_t per = (n4 + gridDim.x - 1) / gridDim.x;
    const size_t b0 = per * blockIdx.x, b1 = b0 + per < n4 ? b0 + per : n4;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (size_t i = b0 + threadIdx.x; i < b1; i += blockDim.x) {
        if (MODE == 0) p[i] = z; else __stcs(p + i, z);
    }
}
int main() {
    const size_t bytes = 97ull << 20, n4 = bytes / 16;
    const int NB = 4;
    float4* buf[NB];
    for (int i = 0; i < NB; ++i) cudaMalloc(&buf[i], bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char* name, auto launch) {
        for (int i = 0; i < 8; ++i) launch(buf[i % NB]);
        cudaDeviceSynchronize();
        const int R = 40;
        cudaEventRecord(e0);
        for (int i = 0; i < R; ++i) launch(buf[i % NB]);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-34s %7.2f us  %7.1f GB/s\n", name, ms / R * 1e3, bytes / (ms / R * 1e-3) / 1e9);
    };
    run("memsetAsync", [&](float4* p) { cudaMemsetAsync(p, 0, bytes); });
    for (int g : {148 * 2, 148 * 4, 148 * 8, 148 * 16}) {
        char nm[64];
        snprintf(nm, 64, "st      grid %d x256", g); run(nm, [&](float4* p) { fill<0><<<g, 256>>>(p, n4); });
        snprintf(nm, 64, "st.cs   grid %d x256", g); run(nm, [&](float4* p) { fill<1><<<g, 256>>>(p, n4); });
        snprintf(nm, 64, "st.cg   grid %d x256", g); run(nm, [&](float4* p) { fill<2><<<g, 256>>>(p, n4); });
        snprintf(nm, 64, "chunk st grid %d x256", g); run(nm, [&](float4* p) { fill_chunk<0><<<g, 256>>>(p, n4); });
        snprintf(nm, 64, "chunk st.cs grid %d x256", g); run(nm, [&](float4* p) { fill_chunk<1><<<g, 256>>>(p, n4); });
    }
    run("st grid 148x1024", [&](float4* p) { fill<0><<<148, 1024>>>(p, n4); });
    run("st grid 296x1024", [&](float4* p) { fill<0><<<296, 1024>>>(p, n4); });
    return 0;
}
