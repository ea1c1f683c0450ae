// Counter-based random permutations of [0, n): a keyed 8-round Feistel network on ceil(log2 n) bits with cycle walking
// (values >= n are encrypted again until they fall inside).  pi_draw(i) is a pure function of (seed, draw, i), so every
// thread evaluates the entries it needs - no permutation array, no sort, no host RNG.  Used for the permutation
// inference of esda.moran.Moran(..., permutations=P) (np.random.permutation(z) per draw in the reference's library).
// chrysalis_b200/perm.py holds the same function in numpy (tests pin the two to each other).
#pragma once
#include <stdint.h>

namespace chr {

constexpr int kPermRounds = 8;

__host__ __device__ __forceinline__ uint32_t perm_mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU;
    x ^= x >> 15; x *= 0x846ca68bU;
    x ^= x >> 16;
    return x;
}

struct PermKeys {
    uint32_t k[kPermRounds];
    uint32_t n;
    int a, b;                 // bit widths of the two halves
};

__host__ __device__ __forceinline__ PermKeys perm_keys(uint64_t seed, uint32_t draw, uint32_t n) {
    PermKeys pk;
    pk.n = n;
    int bits = 1;
    while (bits < 32 && (1u << bits) < n) ++bits;
    if (bits < 2) bits = 2;
    pk.a = bits / 2;
    pk.b = bits - pk.a;
    const uint32_t lo = (uint32_t)seed, hi = (uint32_t)(seed >> 32);
#pragma unroll
    for (int i = 0; i < kPermRounds; ++i)
        pk.k[i] = perm_mix(lo ^ perm_mix(hi + draw * 0x9E3779B9u + (uint32_t)(i + 1) * 0x85EBCA6Bu));
    return pk;
}

__host__ __device__ __forceinline__ uint32_t perm_eval(const PermKeys& pk, uint32_t x) {
    const uint32_t ma = (1u << pk.a) - 1u, mb = (1u << pk.b) - 1u;
    do {
        uint32_t l = x & ma, r = x >> pk.a;
#pragma unroll
        for (int i = 0; i < kPermRounds; ++i) {
            if (i & 1) r ^= perm_mix(l * 0x9E3779B1u + pk.k[i]) & mb;
            else l ^= perm_mix(r * 0x9E3779B1u + pk.k[i]) & ma;
        }
        x = (r << pk.a) | l;
    } while (x >= pk.n);
    return x;
}

}  // namespace chr
