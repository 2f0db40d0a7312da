// mld_host_pack.cpp -- host side of the host-buffer pipeline (mld_process_frames_host): strip point records down to 12-byte xyz
// in a pinned staging buffer. Only x, y, z of a record are ever used (DepthEstimator.cpp:169 casts topRows<3>), the pipeline is
// PCIe bound, and a 32-byte pcl::PointXYZI record carries 20 bytes of padding and intensity.
//
// A core's copy rate is bounded by its outstanding cache misses, so the loop has to be cheap per byte and keep the prefetchers
// fed: with AVX-512, 16 records are loaded as whole cache lines, squeezed with two rounds of two-source permutes
// (vpermt2ps) into three full 64-byte lines and written with non-temporal stores (no read-for-ownership of the staging buffer);
// measured 1.6-2.2x the scalar loop per thread on the Sapphire Rapids hosts of this pool. Plain C++ (no CUDA): compiled by the host
// compiler with per-function target attributes and selected at run time, so the library loads on hosts without AVX-512.
#include <immintrin.h>
#include <stdint.h>

#include "mld_host_pack.h"

namespace {

// NT: non-temporal stores (no read-for-ownership of the staging buffer, no cache pollution) -- what the pipeline uses. Plain
// stores into a small staging ring that could stay in the last-level cache until the copy engine has read it were measured with
// scripts/host_pack_probe.cu: slower at every ring size (12.9-19.4 k against 14.3-20.1 k frames/s), so only the probe asks for them.
template <bool NT>
void pack_scalar(const unsigned char* p, float* q, long long n, int stride_bytes) {
    for (long long i = 0; i < n; i++, p += stride_bytes, q += 3) {
        const int* f = reinterpret_cast<const int*>(p);
        _mm_prefetch(reinterpret_cast<const char*>(p) + 1024, _MM_HINT_NTA);
        if (NT) {
            _mm_stream_si32(reinterpret_cast<int*>(q), f[0]);
            _mm_stream_si32(reinterpret_cast<int*>(q) + 1, f[1]);
            _mm_stream_si32(reinterpret_cast<int*>(q) + 2, f[2]);
        } else {
            reinterpret_cast<int*>(q)[0] = f[0];
            reinterpret_cast<int*>(q)[1] = f[1];
            reinterpret_cast<int*>(q)[2] = f[2];
        }
    }
}

// records until q sits on a 64-byte line (12 i = -offset mod 64 has a solution i < 16 for every float-aligned q)
long long head_records(const float* q, long long n) {
    long long i = 0;
    while (i < n && ((reinterpret_cast<uintptr_t>(q) + 12u * (uintptr_t)i) & 63u) != 0) i++;
    return i;
}

template <bool NT>
__attribute__((target("avx512f"))) void pack32_avx512(const unsigned char* p, float* q, long long n) {
    const long long h = head_records(q, n);
    pack_scalar<NT>(p, q, h, 32);
    p += h * 32; q += h * 3; n -= h;
    // zmm = two records (x y z . i . . . | x y z . i . . .); four records -> 12 floats, then 16 records -> 3 lines
    const __m512i i4 = _mm512_setr_epi32(0, 1, 2, 8, 9, 10, 16, 17, 18, 24, 25, 26, 0, 0, 0, 0);
    const __m512i o0 = _mm512_setr_epi32(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 16, 17, 18, 19);
    const __m512i o1 = _mm512_setr_epi32(4, 5, 6, 7, 8, 9, 10, 11, 16, 17, 18, 19, 20, 21, 22, 23);
    const __m512i o2 = _mm512_setr_epi32(8, 9, 10, 11, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27);
    long long i = 0;
    for (; i + 16 <= n; i += 16, p += 512, q += 48) {
        for (int k = 0; k < 8; k++) _mm_prefetch(reinterpret_cast<const char*>(p) + 1024 + 64 * k, _MM_HINT_NTA);
        const __m512 s0 = _mm512_loadu_ps(p), s1 = _mm512_loadu_ps(p + 64), s2 = _mm512_loadu_ps(p + 128), s3 = _mm512_loadu_ps(p + 192);
        const __m512 s4 = _mm512_loadu_ps(p + 256), s5 = _mm512_loadu_ps(p + 320), s6 = _mm512_loadu_ps(p + 384), s7 = _mm512_loadu_ps(p + 448);
        const __m512 p0 = _mm512_permutex2var_ps(s0, i4, s1), p1 = _mm512_permutex2var_ps(s2, i4, s3);
        const __m512 p2 = _mm512_permutex2var_ps(s4, i4, s5), p3 = _mm512_permutex2var_ps(s6, i4, s7);
        const __m512 r0 = _mm512_permutex2var_ps(p0, o0, p1), r1 = _mm512_permutex2var_ps(p1, o1, p2), r2 = _mm512_permutex2var_ps(p2, o2, p3);
        if (NT) {
            _mm512_stream_ps(q, r0); _mm512_stream_ps(q + 16, r1); _mm512_stream_ps(q + 32, r2);
        } else {
            _mm512_store_ps(q, r0); _mm512_store_ps(q + 16, r1); _mm512_store_ps(q + 32, r2);
        }
    }
    pack_scalar<NT>(p, q, n - i, 32);
}

template <bool NT>
__attribute__((target("avx512f"))) void pack16_avx512(const unsigned char* p, float* q, long long n) {
    const long long h = head_records(q, n);
    pack_scalar<NT>(p, q, h, 16);
    p += h * 16; q += h * 3; n -= h;
    // zmm = four float4 records; 16 records (4 lines) -> 3 lines
    const __m512i o0 = _mm512_setr_epi32(0, 1, 2, 4, 5, 6, 8, 9, 10, 12, 13, 14, 16, 17, 18, 20);
    const __m512i o1 = _mm512_setr_epi32(5, 6, 8, 9, 10, 12, 13, 14, 16, 17, 18, 20, 21, 22, 24, 25);
    const __m512i o2 = _mm512_setr_epi32(10, 12, 13, 14, 16, 17, 18, 20, 21, 22, 24, 25, 26, 28, 29, 30);
    long long i = 0;
    for (; i + 16 <= n; i += 16, p += 256, q += 48) {
        for (int k = 0; k < 4; k++) _mm_prefetch(reinterpret_cast<const char*>(p) + 1024 + 64 * k, _MM_HINT_NTA);
        const __m512 s0 = _mm512_loadu_ps(p), s1 = _mm512_loadu_ps(p + 64), s2 = _mm512_loadu_ps(p + 128), s3 = _mm512_loadu_ps(p + 192);
        const __m512 r0 = _mm512_permutex2var_ps(s0, o0, s1), r1 = _mm512_permutex2var_ps(s1, o1, s2), r2 = _mm512_permutex2var_ps(s2, o2, s3);
        if (NT) {
            _mm512_stream_ps(q, r0); _mm512_stream_ps(q + 16, r1); _mm512_stream_ps(q + 32, r2);
        } else {
            _mm512_store_ps(q, r0); _mm512_store_ps(q + 16, r1); _mm512_store_ps(q + 32, r2);
        }
    }
    pack_scalar<NT>(p, q, n - i, 16);
}

}  // namespace

int mld_host_pack_level() {
    static const int level = __builtin_cpu_supports("avx512f") ? 512 : 0;
    return level;
}

namespace {
template <bool NT>
void pack_dispatch(const unsigned char* p, int stride_bytes, float* dst, long long n) {
    if (mld_host_pack_level() == 512 && stride_bytes == 32)
        pack32_avx512<NT>(p, dst, n);
    else if (mld_host_pack_level() == 512 && stride_bytes == 16)
        pack16_avx512<NT>(p, dst, n);
    else
        pack_scalar<NT>(p, dst, n, stride_bytes);
}
}  // namespace

void mld_host_pack_xyz(const void* src, int stride_bytes, float* dst, long long n, int cached_stores) {
    const unsigned char* p = static_cast<const unsigned char*>(src);
    if (cached_stores)
        pack_dispatch<false>(p, stride_bytes, dst, n);
    else
        pack_dispatch<true>(p, stride_bytes, dst, n);
    _mm_sfence();
}
