// common.h -- shared host/device definitions for libseqwin_b200.
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/seqwin_b200.h"

#if defined(__CUDACC__)
#define SW_HD __host__ __device__ __forceinline__
// rarely executed device code: kept out of line, so that it does not sit inside the hot loops' instruction footprint
#define SW_COLD __host__ __device__ __noinline__
#else
#define SW_HD inline
#define SW_COLD inline
#endif

namespace sw {

// ---- errors (no exception crosses the C ABI; api.cu catches these) -----------------------
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
[[noreturn]] inline void fail_runtime(const std::string& m) { throw Error(SW_ERR_RUNTIME, m); }
[[noreturn]] inline void fail_value(const std::string& m) { throw Error(SW_ERR_VALUE, m); }

// ---- 2-bit packed layout -------------------------------------------------------------------
// 16 bases per little-endian 32-bit word, base i of a record at bits 2*(i&15) of word i>>4.
// Every record starts on a 16-byte boundary (64 bases) of the packed stream so tiles can be
// staged with 128-bit / bulk copies.  Bytes that the reference cannot hash (anything but
// ACGTU in either case, hashing_internals.hpp:136-169) are packed as 0 and listed as
// record-relative "invalid runs".
constexpr uint32_t kBasesPerWord = 16;
constexpr uint32_t kRecordAlignWords = 4;
constexpr uint32_t kTailPadWords = 64;  // readable slack after the last record

// A maximal run of >= k hashable bases gives `n` consecutive entries of the record's
// VALID k-mer stream (the index space minimizer windows slide over, minimizer.cpp:69-87).
struct Piece {
    uint32_t kidx;  // index in the record's valid-k-mer stream of the first k-mer of the run
    uint32_t pos;   // record position of that k-mer
    uint32_t n;     // number of k-mers in the run
};

// One CTA-sized unit of sketch work: windows [e0 .. e0+n_kmers-w] of one record's valid
// k-mer stream.  If `first` is 0 the first window only provides the previous argmin.
struct Tile {
    uint32_t rec;       // batch-local record index
    uint32_t e0;        // valid-k-mer index of the first k-mer held by the tile
    uint32_t n_kmers;   // k-mers held (<= TK)
    uint32_t piece_lo;  // first piece (global index) overlapping the tile
    uint32_t n_pieces;  // pieces overlapping the tile
    uint32_t first;     // 1: window 0 is the record's first window (always emits)
};

// ntHash2 rolling tables for one k (see nthash.h): entry (out<<2)|in for the steady state,
// entries 16..19 for the warm-up (no outgoing base).
struct alignas(16) RollEntry {
    uint64_t f;  // xor into srol(fwd)
    uint64_t r;  // xor into rev before sror
};
struct RollTable {
    RollEntry e[20];
};

}  // namespace sw
