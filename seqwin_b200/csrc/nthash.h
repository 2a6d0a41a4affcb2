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
// sub-word, bits 33..63 form the high 31-bit sub-word.  The device versions spell the bit merges
// as single LOP3s (select-by-mask), which nvcc does not derive on its own.
SW_HD uint64_t srol1(uint64_t x)
{
    const uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
#if defined(__CUDA_ARCH__)
    const uint32_t a = lo + lo;                    // lo << 1 (bit 0 clear)
    const uint32_t t = __funnelshift_l(lo, hi, 1); // (hi << 1) | (lo >> 31)
    const uint32_t u = hi >> 30;                   // bit 1 = old bit 63
    uint32_t nlo, nhi;
    asm("lop3.b32 %0, %1, %2, 1, 0xF8;" : "=r"(nlo) : "r"(a), "r"(hi));    // a | (hi & 1): bit32 wraps to bit0
    asm("lop3.b32 %0, %1, 2, %2, 0xB8;" : "=r"(nhi) : "r"(t), "r"(u));     // (t & ~2) | (u & 2): bit63 wraps to bit33
#else
    const uint32_t nlo = (lo << 1) | (hi & 1u);
    const uint32_t t = (hi << 1) | (lo >> 31);
    const uint32_t nhi = (t & ~2u) | ((hi >> 30) & 2u);
#endif
    return ((uint64_t)nhi << 32) | nlo;
}

SW_HD uint64_t sror1(uint64_t x)
{
    const uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
#if defined(__CUDA_ARCH__)
    const uint32_t nlo = __funnelshift_r(lo, hi, 1);  // (lo >> 1) | (hi << 31): bit32 moves down to bit31
    const uint32_t t = hi >> 1;                       // bits 33..63 move down
    const uint32_t v = hi << 30;                      // bit 31 = old bit 33
    uint32_t y, nhi;
    asm("lop3.b32 %0, %1, 0x7FFFFFFE, %2, 0xE2;" : "=r"(y) : "r"(t), "r"(lo));     // bits 1..30 from t, bit 0 (old bit 0 -> bit 32) from lo
    asm("lop3.b32 %0, %1, 0x80000000, %2, 0xB8;" : "=r"(nhi) : "r"(y), "r"(v));    // bit 31 (old bit 33 -> bit 63) from v
#else
    const uint32_t nlo = (lo >> 1) | (hi << 31);
    const uint32_t t = hi >> 1;
    const uint32_t nhi = (t & 0x7FFFFFFEu) | (lo & 1u) | ((hi & 2u) << 30);
#endif
    return ((uint64_t)nhi << 32) | nlo;
}

// Split rotation by 4 (tetramer-table warm-up): both sub-words rotate left by 4.
SW_HD uint64_t srol4(uint64_t x)
{
    const uint64_t lo_mask = 0x1FFFFFFFFULL, hi_mask = 0x7FFFFFFFULL;
    const uint64_t lo = x & lo_mask, hi = x >> 33;
    const uint64_t nlo = ((lo << 4) | (lo >> 29)) & lo_mask;
    const uint64_t nhi = ((hi << 4) | (hi >> 27)) & hi_mask;
    return (nhi << 33) | nlo;
}

// Running rightmost minimum: if (h <= best) { best = h; best_i = i; } with ONE 64-bit compare
// feeding all three selects (nvcc otherwise emits a min plus a second compare for the index).
SW_HD void take_if_le(uint64_t h, uint32_t i, uint64_t& best, uint32_t& best_i)
{
#if defined(__CUDA_ARCH__)
    asm("{\n\t.reg .pred p;\n\tsetp.gt.u64 p, %2, %0;\n\tselp.b64 %0, %0, %2, p;\n\tselp.b32 %1, %1, %3, p;\n\t}"
        : "+l"(best), "+r"(best_i)
        : "l"(h), "r"(i));
#else
    if (h <= best) { best = h; best_i = i; }
#endif
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

// 4-base tables for the warm-up of a run (the role of btllib's TETRAMER_TAB in base_forward_hash /
// base_reverse_hash, nthash_kmer.hpp:22-54,104-133).  Index = 4 packed bases, base j at bits 2j:
//   fwd4[b] = srol^3(S[b0]) ^ srol^2(S[b1]) ^ srol(S[b2]) ^ S[b3]       fwd = srol4(fwd) ^ fwd4[group]
//   rev4[b] = S'[b0] ^ srol(S'[b1]) ^ srol^2(S'[b2]) ^ srol^3(S'[b3])    rev = srol4(rev) ^ rev4[group], groups last to first
// with S'[c] = S[3 - c].
struct TetraTable {
    uint64_t fwd4[256];
    uint64_t rev4[256];
};
inline void make_tetra_table(TetraTable& t)
{
    for (int b = 0; b < 256; ++b) {
        const int c[4] = {b & 3, (b >> 2) & 3, (b >> 4) & 3, (b >> 6) & 3};
        uint64_t f = 0, r = 0;
        for (int j = 0; j < 4; ++j) {
            f ^= srol_n(kSeed[c[j]], 3 - j);
            r ^= srol_n(kSeed[3 - c[j]], j);
        }
        t.fwd4[b] = f;
        t.rev4[b] = r;
    }
}

// One rolling step: state (fwd, rev) absorbs table entry e.
SW_HD void roll_step(uint64_t& fwd, uint64_t& rev, const RollEntry& e)
{
    fwd = srol1(fwd) ^ e.f;
    rev = sror1(rev ^ e.r);
}

}  // namespace sw
