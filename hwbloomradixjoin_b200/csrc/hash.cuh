// hash.cuh -- K0: device hash library, bit-exact with the reference's hash.c / spooky.c.
// Arithmetic re-derived from SURVEY.md Appendix A/C; each function names the reference lines it matches.
#pragma once
#include <cstdint>

namespace hwbrj {

// ---- CRC-32C (hash.c:6-10: _mm_crc32_u32(seed,key)) -------------------------------------------------
// Castagnoli polynomial, reflected (0x82F63B78), register initialised with `seed`, 4 key bytes LSB
// first, no final xor. Message length == register width, so crc = clock^32(seed ^ key), a GF(2)-linear map of
// x = seed ^ key: it decomposes into table lookups over any split of x (nibbles below).
__host__ __device__ inline uint32_t crc32c_bitwise(uint32_t seed, uint32_t key) {
    uint32_t crc = seed ^ key;
#pragma unroll
    for (int i = 0; i < 32; i++) crc = (crc >> 1) ^ (0x82F63B78u & (0u - (crc & 1u)));
    return crc;
}

// Byte tables T[j][b] = clock^32(b << 8j); crc = T[0][x&255] ^ T[1][(x>>8)&255] ^ T[2][(x>>16)&255] ^ T[3][x>>24].
// Kernels keep the 4 KB in shared memory (per-lane random indices would serialise constant memory). Measured: a
// bank-conflict-free variant with 8 nibble tables replicated per bank (16 KB) was SLOWER (15.2 vs 13.4 ms for the
// BLOCKED C1 probe): the probe kernels are instruction-issue bound, not shared-memory bound, and nibbles double the
// lookup instructions.
constexpr int kCrcSmemWords = 4 * 256;
struct CrcTables {
    uint32_t t[4][256];
};

inline void crc_tables_fill(CrcTables& T) {
    for (int j = 0; j < 4; j++)
        for (uint32_t b = 0; b < 256; b++) T.t[j][b] = crc32c_bitwise(0u, b << (8 * j));
}

__device__ __forceinline__ uint32_t crc32c_tab(const uint32_t* __restrict__ tab /* [kCrcSmemWords] in smem */,
                                               uint32_t seed, uint32_t key) {
    const uint32_t x = seed ^ key;
    return tab[x & 255u] ^ tab[256 + ((x >> 8) & 255u)] ^ tab[512 + ((x >> 16) & 255u)] ^ tab[768 + (x >> 24)];
}

// ---- CrapWow (hash.c:27-47) ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t hash_crapwow(uint32_t seed, uint32_t key) {
    const uint32_t n = 0x5052acdbu;
    uint32_t lo = 4u;             // h = sizeof(intkey_t)
    uint32_t hi = 4u + seed + n;  // k = h + seed + n
    uint64_t p = (uint64_t)key * n;
    lo ^= (uint32_t)p;
    hi ^= (uint32_t)(p >> 32);
    p = (uint64_t)(lo ^ (hi + n)) * n;
    lo ^= (uint32_t)p;
    hi ^= (uint32_t)(p >> 32);
    return hi ^ lo;
}

// the byte-wise hashes mix a SIGNED char (hash.c:19,61,75,89,113,125,137)
__host__ __device__ __forceinline__ uint32_t sbyte(uint32_t key, int i) {
    return (uint32_t)(int32_t)(int8_t)((key >> (8 * i)) & 0xFFu);
}

__host__ __device__ inline uint32_t hash_fnv(uint32_t seed, uint32_t key) {  // hash.c:12-25
    uint32_t h = seed ^ 2166136261u;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        h ^= sbyte(key, i);
        h *= 16777619u;
    }
    return h;
}

__host__ __device__ inline uint32_t hash_coffin(uint32_t, uint32_t key) {  // hash.c:55-66
    uint32_t r = 0x55555555u;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        r ^= sbyte(key, i);
        r = (r << 5) | (r >> 27);
    }
    return r;
}

__host__ __device__ inline uint32_t hash_murmur_oaat(uint32_t seed, uint32_t key) {  // hash.c:68-81
    uint32_t h = seed;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        h ^= sbyte(key, i);
        h *= 0x5bd1e995u;
        h ^= h >> 15;
    }
    return h;
}

__host__ __device__ inline uint32_t hash_jenkins_oaat(uint32_t seed, uint32_t key) {  // hash.c:83-99
    uint32_t h = seed;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        h += sbyte(key, i);
        h += h << 10;
        h ^= h >> 6;
    }
    h += h << 3;
    h ^= h >> 11;
    h += h << 15;
    return h;
}

__host__ __device__ __forceinline__ uint64_t rol64(uint64_t x, int b) { return (x << b) | (x >> (64 - b)); }

// hash_Spooky(seed,key) = hash_spooky32(key,seed) (hash.c:101-105, spooky.c:37-43): SpookyHash
// "Short" with a 4-byte message reduces to ShortEnd (spooky.h:110-146) on
// (h0,h1,h2,h3) = (seed, seed, sc_const + sext(key), 4<<56); the result is the low word of h0.
__host__ __device__ inline uint32_t hash_spooky(uint32_t seed, uint32_t key) {
    uint64_t h0 = seed, h1 = seed;
    uint64_t h2 = 0xdeadbeefdeadbeefULL + (uint64_t)(int64_t)(int32_t)key;
    uint64_t h3 = (uint64_t)4 << 56;
#define HWBRJ_SE(a, b, r) a ^= b; b = rol64(b, r); a += b;
    HWBRJ_SE(h3, h2, 15) HWBRJ_SE(h0, h3, 52) HWBRJ_SE(h1, h0, 26) HWBRJ_SE(h2, h1, 51)
    HWBRJ_SE(h3, h2, 28) HWBRJ_SE(h0, h3, 9)  HWBRJ_SE(h1, h0, 47) HWBRJ_SE(h2, h1, 54)
    HWBRJ_SE(h3, h2, 32) HWBRJ_SE(h0, h3, 25) HWBRJ_SE(h1, h0, 63)
#undef HWBRJ_SE
    return (uint32_t)h0;
}

__host__ __device__ inline uint32_t hash_kr_v2(uint32_t seed, uint32_t key) {  // hash.c:107-117
    uint32_t h = seed;
#pragma unroll
    for (int i = 0; i < 4; i++) h = sbyte(key, i) + 31u * h;
    return h;
}

__host__ __device__ inline uint32_t hash_djb2(uint32_t, uint32_t key) {  // hash.c:119-129
    uint32_t h = 5381u;
#pragma unroll
    for (int i = 0; i < 4; i++) h = ((h << 5) + h) + sbyte(key, i);
    return h;
}

__host__ __device__ inline uint32_t hash_x17(uint32_t seed, uint32_t key) {  // hash.c:131-140
    uint32_t h = seed;
#pragma unroll
    for (int i = 0; i < 4; i++) h = 17u * h + (sbyte(key, i) - (uint32_t)' ');
    return h ^ (h >> 16);
}

// order of hash.h:12-40
__host__ __device__ inline uint32_t hash_dispatch(int which, uint32_t seed, uint32_t key) {
    switch (which) {
        case 0: return crc32c_bitwise(seed, key);
        case 1: return hash_fnv(seed, key);
        case 2: return hash_crapwow(seed, key);
        case 3: return hash_coffin(seed, key);
        case 4: return hash_murmur_oaat(seed, key);
        case 5: return hash_jenkins_oaat(seed, key);
        case 6: return hash_spooky(seed, key);
        case 7: return hash_kr_v2(seed, key);
        case 8: return hash_djb2(seed, key);
        default: return hash_x17(seed, key);
    }
}

}  // namespace hwbrj
