// Micro-benchmark: does "compute data compression" (cuMemCreate with CU_MEM_ALLOCATION_COMP_GENERIC) raise the fill rate
// of all-zero data on this GPU?  Fill / read-back of a normal and of a compressible 96 MB buffer (rotating over 4 buffers).
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#define CK(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char* s_; cuGetErrorString(r_, &s_); printf("%s failed: %s\n", #x, s_); return 1; } } while (0)
__global__ void wr(float4* p, size_t n4, float v) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const float4 z = make_float4(v, v, v, v);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) __stcs(p + i, z);
}
__global__ void wr_rand(float4* p, size_t n4) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        unsigned h = (unsigned)i * 2654435761u;
        __stcs(p + i, make_float4(__uint_as_float(h), __uint_as_float(h * 31u), __uint_as_float(h * 131u), __uint_as_float(h ^ 0x5bd1e995u)));
    }
}
__global__ void rd(const float4* p, size_t n4, float* out) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    float a = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) { const float4 v = __ldcs(p + i); a += v.x + v.y + v.z + v.w; }
    if (a == 123.456f) out[0] = a;
}
int main() {
    cudaFree(0);
    CUdevice dev; CK(cuDeviceGet(&dev, 0));
    int comp = 0; CK(cuDeviceGetAttribute(&comp, CU_DEVICE_ATTRIBUTE_GENERIC_COMPRESSION_SUPPORTED, dev));
    printf("generic compression supported: %d\n", comp);
    const size_t bytes = 96ull << 20, n4 = bytes / 16;
    const int NB = 4;
    float4 *nb[NB], *cb[NB];
    float* out; cudaMalloc(&out, 4);
    for (int i = 0; i < NB; ++i) cudaMalloc(&nb[i], bytes);
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = 0;
    prop.allocFlags.compressionType = CU_MEM_ALLOCATION_COMP_GENERIC;
    size_t gran = 0; CK(cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
    const size_t sz = (bytes + gran - 1) / gran * gran;
    for (int i = 0; i < NB; ++i) {
        CUmemGenericAllocationHandle h; CK(cuMemCreate(&h, sz, &prop, 0));
        CUmemAllocationProp got = {}; CK(cuMemGetAllocationPropertiesFromHandle(&got, h));
        if (i == 0) printf("granularity %zu, compressionType of the allocation: %d\n", gran, (int)got.allocFlags.compressionType);
        CUdeviceptr va; CK(cuMemAddressReserve(&va, sz, 0, 0, 0));
        CK(cuMemMap(va, sz, 0, h, 0));
        CUmemAccessDesc ad = {}; ad.location = prop.location; ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        CK(cuMemSetAccess(va, sz, &ad, 1));
        cb[i] = (float4*)va;
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char* name, auto launch) {
        for (int i = 0; i < 8; ++i) launch(i % NB);
        cudaDeviceSynchronize();
        const int R = 40;
        cudaEventRecord(e0);
        for (int i = 0; i < R; ++i) launch(i % NB);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-44s %7.2f us  %7.1f GB/s\n", name, ms / R * 1e3, bytes / (ms / R * 1e-3) / 1e9);
    };
    run("zero fill, normal memory", [&](int i) { wr<<<592, 256>>>(nb[i], n4, 0.f); });
    run("zero fill, compressible memory", [&](int i) { wr<<<592, 256>>>(cb[i], n4, 0.f); });
    run("read back zeros, normal", [&](int i) { rd<<<1184, 256>>>(nb[i], n4, out); });
    run("read back zeros, compressible", [&](int i) { rd<<<1184, 256>>>(cb[i], n4, out); });
    run("random fill, normal memory", [&](int i) { wr_rand<<<592, 256>>>(nb[i], n4); });
    run("random fill, compressible memory", [&](int i) { wr_rand<<<592, 256>>>(cb[i], n4); });
    run("read back random, compressible", [&](int i) { rd<<<1184, 256>>>(cb[i], n4, out); });
    run("zero fill again, compressible memory", [&](int i) { wr<<<592, 256>>>(cb[i], n4, 0.f); });
    // a read stream (own buffers, random data, normal memory) next to the zero fill: normal vs compressible destination
    float4* rbuf[NB];
    for (int i = 0; i < NB; ++i) { cudaMalloc(&rbuf[i], bytes); wr_rand<<<592, 256>>>(rbuf[i], n4); }
    cudaStream_t s1, s2; cudaStreamCreate(&s1); cudaStreamCreate(&s2);
    cudaEvent_t j; cudaEventCreate(&j);
    auto both = [&](const char* name, int gr, float4** dst) {
        float ms = 0;
        const int R = 30;
        for (int rep = 0; rep < 2; ++rep) {
            cudaDeviceSynchronize();
            cudaEventRecord(e0, s1);
            cudaStreamWaitEvent(s2, e0);
            for (int i = 0; i < R; ++i) {
                rd<<<gr, 256, 0, s1>>>(rbuf[i % NB], n4, out);
                wr<<<592, 256, 0, s2>>>(dst[i % NB], n4, 0.f);
            }
            cudaEventRecord(j, s2);
            cudaStreamWaitEvent(s1, j);
            cudaEventRecord(e1, s1);
            cudaDeviceSynchronize();
            cudaEventElapsedTime(&ms, e0, e1);
        }
        printf("%-44s %7.2f us per (96 MB read + 96 MB zero fill)\n", name, ms / R * 1e3);
    };
    run("read random alone, grid 296", [&](int i) { rd<<<296, 256>>>(rbuf[i], n4, out); });
    run("read random alone, grid 1184", [&](int i) { rd<<<1184, 256>>>(rbuf[i], n4, out); });
    both("reader 296 + zero fill normal", 296, nb);
    both("reader 296 + zero fill compressible", 296, cb);
    both("reader 1184 + zero fill normal", 1184, nb);
    both("reader 1184 + zero fill compressible", 1184, cb);
    return 0;
}
