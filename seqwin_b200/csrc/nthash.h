// nthash.h -- ntHash2 canonical k-mer hashing, host + device.
//
// Restates (does not copy) the arithmetic of the reference's vendored btllib:
//   cpp/vendor/btllib/hashing_internals.hpp:12-17   canonical = fwd + rev (mod 2^64)
//   cpp/vendor/btllib/hashing_internals.hpp:29-73   srol / sror split rotation (33 | 31 bits)
//   cpp/vendor/btllib/hashing_internals.hpp:89-103  extend_hashes (h1 = mix(h0))
//   cpp/vendor/btllib/nthash_kmer.hpp:65-75,145-155 next_forward_hash / next_reverse_hash
// Closed form used for seeding: fwd = XOR_i srol^{k-1-i}(S[s_i]), rev = XOR_i srol^i(S[3-s_i]).
// Feeding k bases through the rolling update with the "outgoing" term dropped yields exactly
// that closed form, which is how the kernels warm a run up (RollTable entries 16..19).
#pragma once

#include "common.h"

namespace sw {

constexpr uint64_t kSeed[4] = {0x3c8bfbb395c60474ULL, 0x3193c18562a02b4cULL,
                               0x20323ed082572324ULL, 0x295549f54be24456ULL};
constexpr uint64_t kMultiSeed = 0x90b45d39fb6da1faULL;
constexpr uint32_t kMultiShift = 27;

// One-step split rotations on the (hi32, lo32) halves: bit 32 belongs to the low 33-bit
// sub-word, bits 33..63 form the high 31-bit sub-word.
SW_HD uint64_t srol1(uint64_t x)
{
    uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    uint32_t nlo = (lo << 1) | (hi & 1u);                       // bit32 wraps to bit0
    uint32_t t = (hi << 1) | (lo >> 31);                        // plain 64-bit shift of the top half
    uint32_t nhi = (t & ~2u) | ((hi >> 30) & 2u);               // bit63 wraps to bit33
    return ((uint64_t)nhi << 32) | nlo;
}

SW_HD uint64_t sror1(uint64_t x)
{
    uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    uint32_t nlo = (lo >> 1) | (hi << 31);                      // bit32 moves down to bit31
    uint32_t t = hi >> 1;                                       // bits 33..63 move down
    uint32_t nhi = (t & 0x7FFFFFFEu) | (lo & 1u) | ((hi & 2u) << 30);  // bit0->bit32, bit33->bit63
    return ((uint64_t)nhi << 32) | nlo;
}

inline uint64_t srol_n(uint64_t x, unsigned d)
{
    const uint64_t lo_mask = 0x1FFFFFFFFULL, hi_mask = 0x7FFFFFFFULL;
    uint64_t lo = x & lo_mask, hi = x >> 33;
    unsigned a = d % 33, b = d % 31;
    if (a) lo = ((lo << a) | (lo >> (33 - a))) & lo_mask;
    if (b) hi = ((hi << b) | (hi >> (31 - b))) & hi_mask;
    return (hi << 33) | lo;
}

SW_HD uint64_t h1_multiplier(uint32_t k) { return 1ULL ^ ((uint64_t)k * 0x90b45d39fb6da1faULL); }
SW_HD uint64_t h1_of(uint64_t h0, uint64_t mult)
{
    uint64_t t = h0 * mult;
    return t ^ (t >> 27);
}

inline RollTable make_roll_table(uint32_t k)
{
    RollTable t;
    uint64_t sk[4];
    for (int b = 0; b < 4; ++b) sk[b] = srol_n(kSeed[b], k);
    for (int out = 0; out < 4; ++out)
        for (int in = 0; in < 4; ++in) {
            t.e[(out << 2) | in].f = kSeed[in] ^ sk[out];
            t.e[(out << 2) | in].r = sk[3 - in] ^ kSeed[3 - out];
        }
    for (int in = 0; in < 4; ++in) {
        t.e[16 + in].f = kSeed[in];
        t.e[16 + in].r = sk[3 - in];
    }
    return t;
}

// One rolling step: state (fwd, rev) absorbs table entry e.
SW_HD void roll_step(uint64_t& fwd, uint64_t& rev, const RollEntry& e)
{
    fwd = srol1(fwd) ^ e.f;
    rev = sror1(rev ^ e.r);
}

}  // namespace sw
