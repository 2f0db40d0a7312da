// mld_hash.h -- counter-based RNG (splitmix64 finaliser) shared by the RANSAC kernels and the synthetic generator;
// compiles as plain C++ (libmld_synth.so) and as CUDA.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define MLD_HD __host__ __device__ __forceinline__
#else
#define MLD_HD inline
#endif

MLD_HD uint64_t mld_mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
MLD_HD uint64_t mld_hash3(uint64_t seed, uint64_t a, uint64_t b, uint64_t c) {
    return mld_mix64(mld_mix64(mld_mix64(mld_mix64(seed) ^ a) ^ b) ^ c);
}
