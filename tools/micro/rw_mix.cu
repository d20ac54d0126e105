// Micro-benchmark: do an HBM read stream and an HBM write stream overlap when they run as two concurrent kernels?
// (the candidate kernel of step i+1 reads 71 MB while the train tail of step i writes 97 MB)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void rd(const float4* p, size_t n4, float* out) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    float a = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = __ldcs(p + i);
        a += v.x + v.y + v.z + v.w;
    }
    if (a == 123.456f) out[0] = a;
}
__global__ void wr(float4* p, size_t n4) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) __stcs(p + i, z);
}
__global__ void cp(const float4* s, float4* d, size_t n4) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) __stcs(d + i, __ldcs(s + i));
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
    const int R = 40;
    for (int gr : {148 * 2, 148 * 4, 148 * 8}) for (int gw : {148 * 2, 148 * 4, 148 * 8}) {
        auto time = [&](bool do_r, bool do_w) {
            for (int rep = 0; rep < 2; ++rep) {
                cudaDeviceSynchronize();
                cudaEventRecord(e0, s1);
                cudaStreamWaitEvent(s2, e0);
                for (int i = 0; i < R; ++i) {
                    if (do_r) rd<<<gr, 256, 0, s1>>>(rb[i % NB], n4, out);
                    if (do_w) wr<<<gw, 256, 0, s2>>>(wb[i % NB], n4);
                }
                cudaEventRecord(j, s2);
                cudaStreamWaitEvent(s1, j);
                cudaEventRecord(e1, s1);
                cudaDeviceSynchronize();
            }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            return ms / R * 1e3;
        };
        const float tr = time(true, false), tw = time(false, true), tb = time(true, true);
        printf("read grid %4d: %6.2f us (%5.0f GB/s)  write grid %4d: %6.2f us (%5.0f GB/s)  both: %6.2f us (%5.0f GB/s total; sum of singles %6.2f)\n",
               gr, tr, bytes / tr / 1e3, gw, tw, bytes / tw / 1e3, tb, 2 * bytes / tb / 1e3, tr + tw);
    }
    for (int g : {148 * 4, 148 * 8}) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaDeviceSynchronize();
            cudaEventRecord(e0, s1);
            for (int i = 0; i < R; ++i) cp<<<g, 256, 0, s1>>>(rb[i % NB], wb[i % NB], n4);
            cudaEventRecord(e1, s1);
            cudaDeviceSynchronize();
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("copy grid %4d: %6.2f us (%5.0f GB/s read+write)\n", g, ms / R * 1e3, 2 * bytes / (ms / R * 1e-3) / 1e9);
    }
    return 0;
}
