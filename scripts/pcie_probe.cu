// H2D bandwidth probe: contiguous copy vs strided 2-D copies that skip the padding of 32-byte pcl::PointXYZI records
// (width 16 or 12 bytes out of a 32-byte pitch), and a packing kernel reading mapped pinned memory directly.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__global__ void pack_from_host(const float4* __restrict__ src32, float4* __restrict__ dst, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src32[i * 2];  // first 16 bytes of every 32-byte record
}
int main() {
    const long long n = 120000LL * 256;  // 256 frames
    void *h = nullptr, *d = nullptr;
    CK(cudaHostAlloc(&h, n * 32, cudaHostAllocMapped));
    CK(cudaMalloc(&d, n * 32));
    memset(h, 1, n * 32);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    auto timeit = [&](const char* name, auto fn, double bytes_useful) {
        fn(); cudaDeviceSynchronize();
        cudaEventRecord(a);
        for (int r = 0; r < 3; r++) fn();
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); ms /= 3;
        printf("%-34s %8.2f ms  %7.1f GB/s useful  %8.0f frames/s\n", name, ms, bytes_useful / ms / 1e6, 256.0 / (ms * 1e-3));
        return 0;
    };
    timeit("contiguous 32 B records", [&] { cudaMemcpyAsync(d, h, n * 32, cudaMemcpyHostToDevice, 0); }, n * 32.0);
    timeit("contiguous 16 B records", [&] { cudaMemcpyAsync(d, h, n * 16, cudaMemcpyHostToDevice, 0); }, n * 16.0);
    timeit("2D width 16 of pitch 32", [&] { cudaMemcpy2DAsync(d, 16, h, 32, 16, n, cudaMemcpyHostToDevice, 0); }, n * 16.0);
    timeit("2D width 12 of pitch 32", [&] { cudaMemcpy2DAsync(d, 12, h, 32, 12, n, cudaMemcpyHostToDevice, 0); }, n * 12.0);
    void* hd = nullptr;
    CK(cudaHostGetDevicePointer(&hd, h, 0));
    timeit("kernel reads mapped pinned (16/32)", [&] { pack_from_host<<<(unsigned)((n + 255) / 256), 256>>>((const float4*)hd, (float4*)d, n); }, n * 16.0);
    return 0;
}
