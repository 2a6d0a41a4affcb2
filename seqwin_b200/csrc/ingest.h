// ingest.h -- host ingest (FASTA -> 2-bit packed batch) and the sketch tile planner.
#pragma once

#include <string>
#include <vector>

#include "common.h"

struct sw_batch {
    uint32_t* words = nullptr;  // packed stream (pinned when a CUDA device is usable)
    size_t n_words = 0;         // including kTailPadWords of zeroed slack
    bool pinned = false;
    std::vector<uint64_t> rec_word_off;  // [R] first word of each record
    std::vector<uint32_t> rec_len;       // [R] bases
    std::vector<uint32_t> rec_inv_off;   // [R+1] slice of inv_* owned by each record
    std::vector<uint32_t> inv_start;     // record-relative start of each unhashable run
    std::vector<uint32_t> inv_len;
    std::vector<uint32_t> record_offsets;  // [A+1] cumulative records per assembly
    std::vector<std::string> ids;          // [R]
    size_t n_bases = 0;
    ~sw_batch();
};

namespace sw {

sw_batch* batch_from_fasta(const char* const* paths, size_t n_paths, uint32_t n_threads);
sw_batch* batch_from_memory(const uint8_t* const* seqs, const uint32_t* lens, const uint32_t* asm_of,
                            const char* const* ids, size_t n_records, size_t n_assemblies,
                            uint32_t n_threads);
sw_batch* batch_concat(const sw_batch* const* parts, size_t n_parts);

struct Plan {
    std::vector<Piece> pieces;
    std::vector<Tile> tiles;
    uint64_t n_windows = 0;  // total windows over all records (upper bound on minimizers)
    uint64_t n_kmers = 0;    // total valid k-mers
};

// Cut every record's valid-k-mer stream into tiles of at most `tk` k-mers (tk > w).
Plan plan_tiles(const sw_batch& b, uint32_t k, uint32_t w, uint32_t tk);

}  // namespace sw
