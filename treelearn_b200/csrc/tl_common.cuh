// Shared helpers for the treelearn_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/treelearn_b200.h"

namespace tl {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define TL_CUDA_CHECK(expr)                                                                         \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            tl::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));     \
            return TL_ERR_CUDA;                                                                     \
        }                                                                                           \
    } while (0)

#define TL_LAUNCH_CHECK()                                                                           \
    do {                                                                                            \
        tl::count_launch();                                                                         \
        cudaError_t _e = cudaGetLastError();                                                        \
        if (_e != cudaSuccess) {                                                                    \
            tl::set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return TL_ERR_CUDA;                                                                     \
        }                                                                                           \
    } while (0)

#define TL_REQUIRE(cond, ...)            \
    do {                                 \
        if (!(cond)) {                   \
            tl::set_error(__VA_ARGS__);  \
            return TL_ERR_ARG;           \
        }                                \
    } while (0)

static inline int64_t pad_rows(int64_t n) { return (n + TL_TILE_ROWS - 1) / TL_TILE_ROWS * TL_TILE_ROWS; }
static inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

// carve typed sub-buffers out of a caller workspace
struct Carver {
    char* base;
    size_t off, cap;
    Carver(void* p, size_t c) : base((char*)p), off(0), cap(c) {}
    template <typename T>
    T* take(size_t count) {
        size_t bytes = align_up(count * sizeof(T));
        T* r = (T*)(base + off);
        off += bytes;
        return r;
    }
    bool ok() const { return off <= cap; }
};

// ---- 3 x 16-bit Morton keys: key = batch << 48 | interleave(x,y,z); low 3 bits = (x&1,y&1,z&1) = kappa
constexpr int kCoordBits = 16;
constexpr uint64_t kMortonMask = (1ull << 48) - 1;
constexpr uint64_t kEmptyKey = ~0ull;

__host__ __device__ __forceinline__ uint64_t part1by2(uint64_t x) {
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}
__host__ __device__ __forceinline__ uint32_t compact1by2(uint64_t x) {
    x &= 0x1249249249249249ull;
    x = (x ^ (x >> 2)) & 0x10c30c30c30c30c3ull;
    x = (x ^ (x >> 4)) & 0x100f00f00f00f00full;
    x = (x ^ (x >> 8)) & 0x1f0000ff0000ffull;
    x = (x ^ (x >> 16)) & 0x1f00000000ffffull;
    x = (x ^ (x >> 32)) & 0x1fffffull;
    return (uint32_t)x;
}
__host__ __device__ __forceinline__ uint64_t make_key(uint32_t b, uint32_t x, uint32_t y, uint32_t z) {
    return ((uint64_t)b << 48) | (part1by2(x) << 2) | (part1by2(y) << 1) | part1by2(z);
}
__host__ __device__ __forceinline__ void split_key(uint64_t key, int& b, int& x, int& y, int& z) {
    b = (int)(key >> 48);
    uint64_t m = key & kMortonMask;
    x = (int)compact1by2(m >> 2);
    y = (int)compact1by2(m >> 1);
    z = (int)compact1by2(m);
}

__device__ __forceinline__ uint64_t hash64(uint64_t k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return k;
}

// open-addressing table: keys (u64, kEmptyKey = free) + vals (i32)
__device__ __forceinline__ void hash_insert(uint64_t* tkeys, int32_t* tvals, uint64_t mask, uint64_t key, int32_t val) {
    uint64_t slot = hash64(key) & mask;
    while (true) {
        unsigned long long prev = atomicCAS((unsigned long long*)&tkeys[slot], (unsigned long long)kEmptyKey,
                                            (unsigned long long)key);
        if (prev == kEmptyKey || prev == key) {
            tvals[slot] = val;
            return;
        }
        slot = (slot + 1) & mask;
    }
}
__device__ __forceinline__ int32_t hash_find(const uint64_t* __restrict__ tkeys, const int32_t* __restrict__ tvals,
                                             uint64_t mask, uint64_t key) {
    uint64_t slot = hash64(key) & mask;
    while (true) {
        uint64_t k = __ldg(&tkeys[slot]);
        if (k == key) return __ldg(&tvals[slot]);
        if (k == kEmptyKey) return -1;
        slot = (slot + 1) & mask;
    }
}

static inline uint64_t table_capacity(int64_t n) {
    uint64_t c = 1024;
    while (c < (uint64_t)(2 * n + 2)) c <<= 1;
    return c;
}

}  // namespace tl
