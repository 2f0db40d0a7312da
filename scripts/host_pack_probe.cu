// Host-side probe of the host-buffer pipeline's two resources on the GPU box: packing rate of T host threads (scalar vs AVX-512,
// mld_host_pack.cpp) on pinned memory, the H2D link alone, and both at once.
//   nvcc -O2 -o build/host_pack_probe scripts/host_pack_probe.cu mono_lidar_depth_b200/csrc/mld_host_pack.o
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <atomic>
#include <functional>
#include <immintrin.h>
#include <cuda_runtime.h>
#include "../mono_lidar_depth_b200/csrc/mld_host_pack.h"
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
static void pack_scalar(const unsigned char* p, float* q, long long n, int stride) {
    for (long long i = 0; i < n; i++, p += stride, q += 3) { const float* f = (const float*)p; q[0] = f[0]; q[1] = f[1]; q[2] = f[2]; }
}
int main(int argc, char** argv) {
    const long long N = 120000, FR = 768;
    unsigned char* src; float* stage; void* d;
    CK(cudaHostAlloc((void**)&src, N * FR * 32, cudaHostAllocDefault));
    CK(cudaHostAlloc((void**)&stage, N * FR * 12, cudaHostAllocDefault));
    CK(cudaMalloc(&d, N * FR * 32));
    for (long long i = 0; i < N * FR * 8; i++) ((float*)src)[i] = (float)(i & 1023);
    memset(stage, 0, N * FR * 12);
    printf("hardware_concurrency %u, pack level %d\n", std::thread::hardware_concurrency(), mld_host_pack_level());
    auto pack_all = [&](int T, int variant, int stride) {
        std::atomic<long long> next{0};
        const long long piece = 15008, pieces = (N * FR + piece - 1) / piece;
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++) th.emplace_back([&] {
            for (;;) { long long i = next.fetch_add(1); if (i >= pieces) break;
                long long lo = i * piece, c = std::min(piece, N * FR - lo);
                if (variant) mld_host_pack_xyz(src + lo * stride, stride, stage + lo * 3, c, 0); else pack_scalar(src + lo * stride, stage + lo * 3, c, stride); }
        });
        for (auto& x : th) x.join();
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    };
    for (int T : {4, 8, 14, 16}) for (int v : {0, 1}) for (int stride : {32, 16}) {
        pack_all(T, v, stride);
        double s = pack_all(T, v, stride);
        printf("pack only: T=%2d %s stride %d: %6.1f k frames/s (%5.1f GB/s read)\n", T, v ? "dispatch" : "plain  ", stride, FR / s / 1e3, N * FR * stride / s / 1e9);
    }
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    auto copy_rate = [&](const void* h, size_t bytes, int reps) { cudaEventRecord(a); for (int r = 0; r < reps; r++) cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, 0); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); return bytes * (double)reps / ms / 1e6; };
    copy_rate(src, N * FR * 32, 1);
    printf("H2D alone: %5.1f GB/s (32-byte records), %5.1f GB/s (12-byte staging)\n", copy_rate(src, N * FR * 32, 2), copy_rate(stage, N * FR * 12, 4));
    for (int T : {8, 14}) for (int v : {0, 1}) {
        std::atomic<bool> stop{false};
        double gbs = 0;
        std::thread cp([&] { int n = 0; auto t0 = std::chrono::steady_clock::now(); while (!stop.load()) { cudaMemcpyAsync(d, stage, N * FR * 12, cudaMemcpyHostToDevice, 0); cudaStreamSynchronize(0); n++; }
            gbs = n * (double)(N * FR * 12) / std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / 1e9; });
        double s = 0; for (int r = 0; r < 3; r++) s += pack_all(T, v, 32);
        stop = true; cp.join();
        printf("pack (T=%2d %s, 32-byte) under a running H2D of the staging buffer: %6.1f k frames/s packed, link %5.1f GB/s (= %5.1f k frames/s of 12-byte points)\n", T, v ? "dispatch" : "plain  ", 3 * FR / s / 1e3, gbs, gbs * 1e9 / (N * 12) / 1e3);
    }
    // ring pipeline: R staging buffers of S frames each, packed by T threads (a spinning pool), copied as soon as they are full
    struct Pool {
        int T; std::vector<std::thread> th; std::atomic<long long> next{0}, done{0}; std::atomic<int> gen{0}; std::atomic<bool> quit{false};
        long long items = 0; std::function<void(long long)> fn;
        explicit Pool(int t) : T(t) { for (int i = 1; i < T; i++) th.emplace_back([this] { int seen = 0; while (!quit.load()) { if (gen.load(std::memory_order_acquire) != seen) { seen = gen.load(); work(); } else _mm_pause(); } }); }
        ~Pool() { quit = true; for (auto& x : th) x.join(); }
        void work() { for (;;) { long long i = next.fetch_add(1); if (i >= items) break; fn(i); done.fetch_add(1); } }
        void run(long long n, std::function<void(long long)> f) { fn = f; items = n; done = 0; next = 0; gen.fetch_add(1, std::memory_order_release); work(); while (done.load() < n) _mm_pause(); }
    };
    for (int T : {14}) {
        Pool pool(T);
        for (int cached : {0, 1}) for (int S : {1, 2, 4, 8, 32}) for (int R : {3, 6}) {
            float* ring; CK(cudaHostAlloc((void**)&ring, (size_t)R * S * N * 12, cudaHostAllocDefault));
            memset(ring, 0, (size_t)R * S * N * 12);
            std::vector<cudaEvent_t> ev(R); for (auto& e : ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
            double best = 0;
            for (int rep = 0; rep < 3; rep++) {
                auto t0 = std::chrono::steady_clock::now();
                for (long long i = 0; i * S < FR; i++) {
                    const int b = (int)(i % R);
                    if (i >= R) cudaEventSynchronize(ev[b]);
                    float* buf = ring + (size_t)b * S * N * 3;
                    const long long pts = S * N, piece = 7504, pieces = (pts + piece - 1) / piece;
                    const unsigned char* sp = src + (size_t)i * S * N * 32;
                    pool.run(pieces, [&](long long k) { long long lo = k * piece, c = std::min(piece, pts - lo); mld_host_pack_xyz(sp + lo * 32, 32, buf + lo * 3, c, cached); });
                    cudaMemcpyAsync((char*)d + (size_t)i * S * N * 12, buf, (size_t)S * N * 12, cudaMemcpyHostToDevice, 0);
                    cudaEventRecord(ev[b], 0);
                }
                cudaStreamSynchronize(0);
                double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                best = std::max(best, FR / sec);
            }
            printf("ring pipeline T=%d %s stores, %2d frames per buffer x %d buffers (%5.1f MB): %6.1f k frames/s\n", T, cached ? "plain" : "NT   ", S, R, R * S * N * 12 / 1e6, best / 1e3);
            for (auto& e : ev) cudaEventDestroy(e);
            cudaFreeHost(ring);
        }
    }
    return 0;
}
