// mld_host_pack.h -- host-side record packing of the host-buffer pipeline (mld_host_pack.cpp; plain C++, no CUDA)
#pragma once
extern "C" {
// n records of stride_bytes (first 12 bytes = x, y, z as float) at src -> n x 3 floats at dst; cached_stores = 0: written past
// the cache (non-temporal), 1: plain stores (a staging ring meant to stay in the last-level cache)
void mld_host_pack_xyz(const void* src, int stride_bytes, float* dst, long long n, int cached_stores);
// 512: the AVX-512 paths (16- and 32-byte records) are in use, 0: scalar loop
int mld_host_pack_level();
}
