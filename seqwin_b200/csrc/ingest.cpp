// ingest.cpp -- host side of the hot path: FASTA / .gz -> 2-bit packed pinned batch, and the
// tile planner that cuts each record's valid-k-mer stream into CTA-sized tiles.
//
// Replaces (behaviourally, not textually) the reference's reader
//   cpp/src/utils/fasta_reader.cpp:40-95   read_fasta_core  (line rules, id extraction)
//   cpp/src/utils/fasta_reader.cpp:97-213  plain / gzip front ends (".gz" by suffix only)
// and the per-assembly bookkeeping of build_worker (cpp/src/seqwin/build.cpp:129-147,192-193).
#include "ingest.h"

#include <sys/mman.h>
#include <sys/stat.h>
#include <zlib.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <memory>
#include <mutex>
#include <thread>

namespace sw {

void* alloc_host(size_t bytes, bool* pinned);  // api.cu (cudaHostAlloc) or emul (malloc)
void free_host(void* p, bool pinned);

namespace {

// byte class: 0..3 = A,C,G,T/U code; 4 = unhashable (packed as 0, recorded as invalid);
// 5 = ASCII whitespace (dropped, fasta_reader.cpp:82-86 uses std::isspace)
struct ByteClass {
    uint8_t t[256];
    ByteClass()
    {
        for (int i = 0; i < 256; ++i) t[i] = 4;
        t[(int)'A'] = t[(int)'a'] = 0;
        t[(int)'C'] = t[(int)'c'] = 1;
        t[(int)'G'] = t[(int)'g'] = 2;
        t[(int)'T'] = t[(int)'t'] = t[(int)'U'] = t[(int)'u'] = 3;
        t[(int)' '] = t[(int)'\t'] = t[(int)'\n'] = t[(int)'\r'] = t[(int)'\f'] = t[(int)'\v'] = 5;
    }
};
const ByteClass kClass;

// One assembly, parsed and packed, before it is placed into the batch.
struct AsmPacked {
    std::vector<uint32_t> words;        // records laid out back to back, 4-word aligned
    std::vector<uint64_t> rec_word_off; // local
    std::vector<uint32_t> rec_len;
    std::vector<uint32_t> rec_inv_cnt;
    std::vector<uint32_t> inv_start, inv_len;
    std::vector<std::string> ids;
    size_t n_bases = 0;
};

#if defined(__x86_64__)
// 16 ASCII bases -> one packed word, or false if any byte is not A/C/G/T/U (either case).
// code = ((c >> 1) ^ (c >> 2)) & 3 maps A,C,G,T/U (and lower case) to 0,1,2,3.
__attribute__((target("ssse3"))) inline bool pack16_ssse3(const unsigned char* p, uint32_t* word)
{
    const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i*>(p));
    const __m128i lc = _mm_or_si128(v, _mm_set1_epi8(0x20));
    __m128i ok = _mm_or_si128(_mm_cmpeq_epi8(lc, _mm_set1_epi8('a')), _mm_cmpeq_epi8(lc, _mm_set1_epi8('c')));
    ok = _mm_or_si128(ok, _mm_cmpeq_epi8(lc, _mm_set1_epi8('g')));
    ok = _mm_or_si128(ok, _mm_cmpeq_epi8(lc, _mm_set1_epi8('t')));
    ok = _mm_or_si128(ok, _mm_cmpeq_epi8(lc, _mm_set1_epi8('u')));
    if (_mm_movemask_epi8(ok) != 0xFFFF) return false;
    const __m128i code = _mm_and_si128(_mm_xor_si128(_mm_srli_epi16(v, 1), _mm_srli_epi16(v, 2)), _mm_set1_epi8(3));
    const __m128i t = _mm_maddubs_epi16(code, _mm_set1_epi16(0x0401));   // 2 bases per 16-bit lane
    const __m128i u = _mm_madd_epi16(t, _mm_set1_epi32(0x00100001));     // 4 bases per 32-bit lane
    const __m128i g = _mm_shuffle_epi8(u, _mm_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1));
    *word = (uint32_t)_mm_cvtsi128_si32(g);
    return true;
}
const bool kHaveSsse3 = __builtin_cpu_supports("ssse3");

// 32 ASCII bases -> two packed words (as one u64), or false if any byte is not A/C/G/T/U.
__attribute__((target("avx2"))) inline bool pack32_avx2(const unsigned char* p, uint64_t* out)
{
    const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p));
    const __m256i lc = _mm256_or_si256(v, _mm256_set1_epi8(0x20));
    __m256i ok = _mm256_or_si256(_mm256_cmpeq_epi8(lc, _mm256_set1_epi8('a')), _mm256_cmpeq_epi8(lc, _mm256_set1_epi8('c')));
    ok = _mm256_or_si256(ok, _mm256_cmpeq_epi8(lc, _mm256_set1_epi8('g')));
    ok = _mm256_or_si256(ok, _mm256_cmpeq_epi8(lc, _mm256_set1_epi8('t')));
    ok = _mm256_or_si256(ok, _mm256_cmpeq_epi8(lc, _mm256_set1_epi8('u')));
    if ((uint32_t)_mm256_movemask_epi8(ok) != 0xFFFFFFFFu) return false;
    const __m256i code = _mm256_and_si256(_mm256_xor_si256(_mm256_srli_epi16(v, 1), _mm256_srli_epi16(v, 2)), _mm256_set1_epi8(3));
    const __m256i t = _mm256_maddubs_epi16(code, _mm256_set1_epi16(0x0401));   // 2 bases per 16-bit lane
    const __m256i u = _mm256_madd_epi16(t, _mm256_set1_epi32(0x00100001));     // 4 bases per 32-bit lane
    const __m256i sel = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                         0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    const __m256i g = _mm256_shuffle_epi8(u, sel);                             // 16 bases in the low word of each half
    *out = (uint64_t)(uint32_t)_mm256_extract_epi32(g, 0) | ((uint64_t)(uint32_t)_mm256_extract_epi32(g, 4) << 32);
    return true;
}
const bool kHaveAvx2 = __builtin_cpu_supports("avx2");
#endif

// Streaming packer for one record.
struct RecordPacker {
    AsmPacked& a;
    uint64_t acc = 0;   // pending bases, `fill` of them (2 bits each)
    uint32_t fill = 0;
    uint64_t cnt = 0;  // bases so far (64-bit so that > u32 records are detected, build.cpp:136)
    uint32_t n_inv = 0;
    bool open = false;

    explicit RecordPacker(AsmPacked& a_) : a(a_) {}

    void begin(std::string id)
    {
        a.ids.push_back(std::move(id));
        a.rec_word_off.push_back(a.words.size());
        acc = 0; fill = 0; cnt = 0; n_inv = 0; open = true;
    }
    inline void push_code(uint32_t c)
    {
        acc |= (uint64_t)c << (2 * fill);
        ++cnt;
        if (++fill == 16) { a.words.push_back((uint32_t)acc); acc = 0; fill = 0; }
    }
    inline void push_word16(uint32_t w16)  // 16 hashable bases at once
    {
        acc |= (uint64_t)w16 << (2 * fill);
        a.words.push_back((uint32_t)acc);
        acc >>= 32;
        cnt += 16;
    }
    inline void push_word32(uint64_t w32)  // 32 hashable bases at once
    {
        const uint32_t s = 2 * fill;
        const uint64_t x = acc | (w32 << s);
        a.words.push_back((uint32_t)x);
        a.words.push_back((uint32_t)(x >> 32));
        acc = s ? (w32 >> (64 - s)) : 0;
        cnt += 32;
    }
    inline void push_invalid()
    {
        if (n_inv && (uint64_t)a.inv_start.back() + a.inv_len.back() == cnt) ++a.inv_len.back();
        else { a.inv_start.push_back((uint32_t)cnt); a.inv_len.push_back(1); ++n_inv; }
        push_code(0);
    }
    inline void append_bytes(const unsigned char* p, size_t n)
    {
        for (size_t i = 0; i < n; ++i) {
            const uint8_t c = kClass.t[p[i]];
            if (c < 4) push_code(c);
            else if (c == 4) push_invalid();
        }
    }
    // n bytes that are expected to be bases only (a full-width sequence line): packs them block by
    // block and returns n, or stops before the first block that holds anything else and returns how
    // many bytes it consumed (the caller goes on with the general path from there)
    size_t append_bases(const unsigned char* p, size_t n)
    {
        size_t i = 0;
#if defined(__x86_64__)
        if (kHaveAvx2) {
            for (; i + 32 <= n; i += 32) {
                uint64_t w32;
                if (!pack32_avx2(p + i, &w32)) return i;
                push_word32(w32);
            }
        }
        if (kHaveSsse3) {
            for (; i + 16 <= n; i += 16) {
                uint32_t w16;
                if (!pack16_ssse3(p + i, &w16)) return i;
                push_word16(w16);
            }
        }
#endif
        for (; i < n; ++i) {
            const uint8_t c = kClass.t[p[i]];
            if (c >= 4) return i;
            push_code(c);
        }
        return n;
    }
    void append(const unsigned char* p, size_t n)
    {
        size_t i = 0;
#if defined(__x86_64__)
        if (kHaveAvx2) {
            for (; i + 32 <= n; i += 32) {
                uint64_t w32;
                if (pack32_avx2(p + i, &w32)) push_word32(w32);
                else append_bytes(p + i, 32);
            }
        }
        if (kHaveSsse3) {
            for (; i + 16 <= n; i += 16) {
                uint32_t w16;
                if (pack16_ssse3(p + i, &w16)) push_word16(w16);
                else append_bytes(p + i, 16);
            }
        }
#endif
        append_bytes(p + i, n - i);
    }
    void end(const std::string& path)
    {
        if (!open) return;
        if (cnt > 0xFFFFFFFFull)
            fail_runtime("Sequence length exceeds uint32 range for record " + a.ids.back() +
                         " in assembly " + path);
        if (fill) a.words.push_back((uint32_t)acc);
        while (a.words.size() % kRecordAlignWords) a.words.push_back(0);
        a.rec_len.push_back((uint32_t)cnt);
        a.rec_inv_cnt.push_back(n_inv);
        a.n_bases += cnt;
        open = false;
    }
};

bool ends_with(const std::string& s, const char* suf)
{
    size_t n = strlen(suf);
    return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

// Whole file in memory: plain files are mapped (no copy through the page cache), gzip streams are
// inflated into an uninitialised heap buffer.
struct FileBuf {
    std::unique_ptr<char[]> p;
    const char* data = nullptr;
    size_t n = 0, cap = 0;
    void* map = nullptr;
    size_t map_len = 0;
    FileBuf() = default;
    FileBuf(FileBuf&& o) noexcept : p(std::move(o.p)), data(o.data), n(o.n), cap(o.cap), map(o.map), map_len(o.map_len)
    {
        o.map = nullptr;
        o.data = nullptr;
    }
    FileBuf(const FileBuf&) = delete;
    ~FileBuf() { if (map) munmap(map, map_len); }
    void reserve(size_t c)
    {
        if (c <= cap) return;
        std::unique_ptr<char[]> q(new char[c]);
        if (n) memcpy(q.get(), p.get(), n);
        p = std::move(q);
        data = p.get();
        cap = c;
    }
};

FileBuf slurp(const std::string& path)
{
    FileBuf buf;
    if (ends_with(path, ".gz")) {
        gzFile f = gzopen(path.c_str(), "rb");
        if (!f) fail_runtime("Unable to open gzip FASTA: " + path);
        gzbuffer(f, 1 << 20);
        buf.reserve(1 << 22);
        for (;;) {
            if (buf.cap - buf.n < (1 << 20)) buf.reserve(buf.cap * 2);
            int got = gzread(f, buf.p.get() + buf.n, 1 << 20);
            if (got < 0) {
                int errnum = 0;
                const char* msg = gzerror(f, &errnum);
                std::string m = std::string("gzip read error: ") + (msg ? msg : "unknown");
                gzclose(f);
                fail_runtime(m);
            }
            if (got == 0) break;
            buf.n += (size_t)got;
        }
        gzclose(f);
    } else {
        FILE* f = fopen(path.c_str(), "rb");
        if (!f) fail_runtime("Unable to open FASTA: " + path);
        {
            struct stat sb;
            if (fstat(fileno(f), &sb) == 0 && S_ISREG(sb.st_mode) && sb.st_size > 0) {
                void* m = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fileno(f), 0);
                if (m != MAP_FAILED) {
                    madvise(m, (size_t)sb.st_size, MADV_SEQUENTIAL);
                    buf.map = m;
                    buf.map_len = (size_t)sb.st_size;
                    buf.data = static_cast<const char*>(m);
                    buf.n = (size_t)sb.st_size;
                    fclose(f);
                    return buf;
                }
            }
        }
        fseek(f, 0, SEEK_END);
        const long sz = ftell(f);
        fseek(f, 0, SEEK_SET);
        buf.reserve(sz > 0 ? (size_t)sz + 1 : (size_t)1 << 16);
        for (;;) {  // also copes with special files whose size is unknown
            if (buf.cap == buf.n) buf.reserve(buf.cap * 2);
            const size_t got = fread(buf.p.get() + buf.n, 1, buf.cap - buf.n, f);
            if (got == 0) break;
            buf.n += got;
        }
        fclose(f);
    }
    return buf;
}

// Line rules of read_fasta_core (fasta_reader.cpp:50-87): split at '\n', strip one trailing
// '\r', skip empty / whitespace-only lines, '>' in column 0 opens a record whose id runs to the
// first whitespace, every other line is sequence with whitespace bytes dropped.
void parse_fasta(const std::string& path, AsmPacked& out)
{
    const FileBuf buf = slurp(path);
    const unsigned char* p = (const unsigned char*)buf.data;
    const size_t n = buf.n;
    out.words.reserve(n / 16 + 64);
    RecordPacker rp(out);
    bool have = false;
    size_t i = 0;
    size_t width = 0;   // length of the previous sequence line: FASTA files keep one width per record
    while (i < n) {
        // Fast path: the line is as long as the one before, ends where predicted and holds bases only
        // (so no newline hides inside it).  Anything else falls through to the general rules below,
        // from the first byte the fast path did not take.
        if (have && width && p[i] != '>' && i + width < n &&
            (p[i + width] == '\n' || (p[i + width] == '\r' && i + width + 1 < n && p[i + width + 1] == '\n'))) {
            const size_t took = rp.append_bases(p + i, width);
            if (took == width) {
                i += width + (p[i + width] == '\n' ? 1 : 2);
                continue;
            }
            if (took) {   // the rest of this line, by the general rules (it cannot be a header: not column 0)
                const size_t q = i + took;
                const void* nl = memchr(p + q, '\n', n - q);
                const size_t j = nl ? (size_t)((const unsigned char*)nl - p) : n;
                size_t end = j;
                if (end > q && p[end - 1] == '\r') --end;
                rp.append(p + q, end - q);
                i = j + 1;
                continue;
            }
        }
        const void* nl = memchr(p + i, '\n', n - i);
        const size_t j = nl ? (size_t)((const unsigned char*)nl - p) : n;
        size_t end = j;
        if (end > i && p[end - 1] == '\r') --end;
        size_t t = i;
        while (t < end && kClass.t[p[t]] == 5) ++t;
        if (t < end) {  // not blank
            if (p[i] == '>') {
                rp.end(path);
                size_t e = i + 1;
                while (e < end && kClass.t[p[e]] != 5) ++e;
                rp.begin(std::string((const char*)p + i + 1, e - (i + 1)));
                have = true;
                width = 0;
            } else {
                if (!have) fail_runtime("Invalid FASTA: sequence encountered before header");
                rp.append(p + i, end - i);
                width = end - i;
            }
        }
        i = j + 1;
    }
    rp.end(path);
}

template <typename F>
void parallel_for(size_t n, uint32_t n_threads, F&& fn)
{
    n_threads = (uint32_t)std::max<size_t>(1, std::min<size_t>(n_threads ? n_threads : 1, n));
    std::atomic<size_t> next{0};
    std::exception_ptr err;
    std::mutex mu;
    auto body = [&]() {
        for (;;) {
            size_t i = next.fetch_add(1);
            if (i >= n) return;
            try {
                fn(i);
            } catch (...) {
                std::lock_guard<std::mutex> g(mu);
                if (!err) err = std::current_exception();
                next.store(n);
                return;
            }
        }
    };
    if (n_threads == 1) {
        body();
    } else {
        std::vector<std::thread> th;
        for (uint32_t t = 0; t < n_threads; ++t) th.emplace_back(body);
        for (auto& t : th) t.join();
    }
    if (err) std::rethrow_exception(err);
}

// Lay the per-assembly pieces out in one (pinned) stream and build the global tables.
sw_batch* assemble(std::vector<AsmPacked>& parts, uint32_t n_threads)
{
    auto* b = new sw_batch();
    try {
        const size_t A = parts.size();
        if (A > 0xFFFFFFFFull) fail_runtime("Number of input assemblies exceeds uint32 range");
        std::vector<size_t> word_base(A + 1, 0), rec_base(A + 1, 0), inv_base(A + 1, 0);
        for (size_t a = 0; a < A; ++a) {
            word_base[a + 1] = word_base[a] + parts[a].words.size();
            rec_base[a + 1] = rec_base[a] + parts[a].rec_len.size();
            inv_base[a + 1] = inv_base[a] + parts[a].inv_start.size();
            if (rec_base[a + 1] > 0xFFFFFFFFull)
                fail_runtime("Total number of FASTA records exceeds uint32 range");
        }
        const size_t R = rec_base[A];
        b->n_words = word_base[A] + kTailPadWords;
        b->words = (uint32_t*)alloc_host(b->n_words * sizeof(uint32_t), &b->pinned);
        if (!b->words) fail_runtime("host allocation of the packed batch failed");
        memset(b->words + word_base[A], 0, kTailPadWords * sizeof(uint32_t));
        b->rec_word_off.resize(R);
        b->rec_len.resize(R);
        b->rec_inv_off.assign(R + 1, 0);
        b->inv_start.resize(inv_base[A]);
        b->inv_len.resize(inv_base[A]);
        b->ids.resize(R);
        b->record_offsets.resize(A + 1);
        for (size_t a = 0; a <= A; ++a) b->record_offsets[a] = (uint32_t)rec_base[a];
        parallel_for(A, n_threads, [&](size_t a) {
            AsmPacked& p = parts[a];
            if (!p.words.empty())
                memcpy(b->words + word_base[a], p.words.data(), p.words.size() * sizeof(uint32_t));
            size_t iv = inv_base[a];
            for (size_t r = 0; r < p.rec_len.size(); ++r) {
                const size_t g = rec_base[a] + r;
                b->rec_word_off[g] = word_base[a] + p.rec_word_off[r];
                b->rec_len[g] = p.rec_len[r];
                b->rec_inv_off[g] = (uint32_t)iv;
                iv += p.rec_inv_cnt[r];
                b->ids[g] = std::move(p.ids[r]);
            }
            if (!p.inv_start.empty()) {
                memcpy(b->inv_start.data() + inv_base[a], p.inv_start.data(), p.inv_start.size() * 4);
                memcpy(b->inv_len.data() + inv_base[a], p.inv_len.data(), p.inv_len.size() * 4);
            }
            std::vector<uint32_t>().swap(p.words);
        });
        if (inv_base[A] > 0xFFFFFFFFull) fail_runtime("too many unhashable runs");
        b->rec_inv_off[R] = (uint32_t)inv_base[A];
        for (size_t a = 0; a < A; ++a) b->n_bases += parts[a].n_bases;
    } catch (...) {
        delete b;
        throw;
    }
    return b;
}

}  // namespace

sw_batch* batch_from_fasta(const char* const* paths, size_t n_paths, uint32_t n_threads)
{
    std::vector<AsmPacked> parts(n_paths);
    const auto t0 = std::chrono::steady_clock::now();
    parallel_for(n_paths, n_threads, [&](size_t a) { parse_fasta(paths[a], parts[a]); });
    const auto t1 = std::chrono::steady_clock::now();
    sw_batch* b = assemble(parts, n_threads);
    if (getenv("SEQWIN_INGEST_PROFILE")) {   // where the host side of a build from FASTA spends its time
        const auto t2 = std::chrono::steady_clock::now();
        fprintf(stderr, "[ingest] %zu files, %u threads: parse + pack %.1f ms, assemble into the pinned batch %.1f ms\n",
                n_paths, n_threads, std::chrono::duration<double, std::milli>(t1 - t0).count(),
                std::chrono::duration<double, std::milli>(t2 - t1).count());
    }
    return b;
}

sw_batch* batch_from_memory(const uint8_t* const* seqs, const uint32_t* lens, const uint32_t* asm_of,
                            const char* const* ids, size_t n_records, size_t n_assemblies,
                            uint32_t n_threads)
{
    std::vector<AsmPacked> parts(n_assemblies);
    std::vector<size_t> first(n_assemblies + 1, n_records);
    for (size_t r = n_records; r-- > 0;) {
        if (asm_of[r] >= n_assemblies) fail_value("asm_of out of range");
        if (r + 1 < n_records && asm_of[r] > asm_of[r + 1]) fail_value("asm_of must be nondecreasing");
        first[asm_of[r]] = r;
    }
    for (size_t a = n_assemblies; a-- > 0;)
        if (first[a] == n_records) first[a] = first[a + 1];
    first[n_assemblies] = n_records;
    parallel_for(n_assemblies, n_threads, [&](size_t a) {
        RecordPacker rp(parts[a]);
        size_t total = 0;
        for (size_t r = first[a]; r < first[a + 1]; ++r) total += lens[r];
        parts[a].words.reserve(total / 16 + 8 * (first[a + 1] - first[a]));
        for (size_t r = first[a]; r < first[a + 1]; ++r) {
            rp.begin(ids ? std::string(ids[r]) : std::string());
            rp.append(seqs[r], lens[r]);
            rp.end("<memory>");
        }
    });
    return assemble(parts, n_threads);
}

// Concatenate packed batches (assemblies of part 0 first): the way a caller that produces its input
// piecewise (synthetic sets, streamed downloads) builds one large batch without ever holding the
// unpacked bases of more than one piece.
sw_batch* batch_concat(const sw_batch* const* parts, size_t n_parts)
{
    auto* b = new sw_batch();
    try {
        size_t words = 0, R = 0, A = 0, I = 0;
        for (size_t p = 0; p < n_parts; ++p) {
            words += parts[p]->n_words - kTailPadWords;
            R += parts[p]->rec_len.size();
            A += parts[p]->record_offsets.size() - 1;
            I += parts[p]->inv_start.size();
        }
        if (R > 0xFFFFFFFFull) fail_runtime("Total number of FASTA records exceeds uint32 range");
        if (A > 0xFFFFFFFFull) fail_runtime("Number of input assemblies exceeds uint32 range");
        if (I > 0xFFFFFFFFull) fail_runtime("too many unhashable runs");
        b->n_words = words + kTailPadWords;
        b->words = (uint32_t*)alloc_host(b->n_words * sizeof(uint32_t), &b->pinned);
        if (!b->words) fail_runtime("host allocation of the packed batch failed");
        memset(b->words + words, 0, kTailPadWords * sizeof(uint32_t));
        b->rec_word_off.reserve(R);
        b->rec_len.reserve(R);
        b->rec_inv_off.reserve(R + 1);
        b->inv_start.reserve(I);
        b->inv_len.reserve(I);
        b->ids.reserve(R);
        b->record_offsets.reserve(A + 1);
        b->record_offsets.push_back(0);
        size_t w0 = 0;
        for (size_t p = 0; p < n_parts; ++p) {
            const sw_batch& s = *parts[p];
            const size_t nw = s.n_words - kTailPadWords;
            if (nw) memcpy(b->words + w0, s.words, nw * sizeof(uint32_t));
            const uint32_t r0 = (uint32_t)b->rec_len.size(), i0 = (uint32_t)b->inv_start.size();
            for (size_t r = 0; r < s.rec_len.size(); ++r) {
                b->rec_word_off.push_back(w0 + s.rec_word_off[r]);
                b->rec_len.push_back(s.rec_len[r]);
                b->rec_inv_off.push_back(i0 + s.rec_inv_off[r]);
            }
            b->inv_start.insert(b->inv_start.end(), s.inv_start.begin(), s.inv_start.end());
            b->inv_len.insert(b->inv_len.end(), s.inv_len.begin(), s.inv_len.end());
            b->ids.insert(b->ids.end(), s.ids.begin(), s.ids.end());
            for (size_t a = 1; a < s.record_offsets.size(); ++a) b->record_offsets.push_back(r0 + s.record_offsets[a]);
            b->n_bases += s.n_bases;
            w0 += nw;
        }
        b->rec_inv_off.push_back((uint32_t)b->inv_start.size());
    } catch (...) {
        delete b;
        throw;
    }
    return b;
}

Plan plan_tiles(const sw_batch& b, uint32_t k, uint32_t w, uint32_t tk)
{
    Plan plan;
    if (tk <= w) fail_runtime("windowsize too large for the sketch tile");
    const uint32_t tw = tk - w;  // windows per tile (one k-mer of slack for the previous window)
    const size_t R = b.rec_len.size();
    for (size_t r = 0; r < R; ++r) {
        const uint32_t L = b.rec_len[r];
        const size_t piece_lo = plan.pieces.size();
        uint64_t n_valid = 0;
        // valid runs between the record's unhashable runs
        uint32_t s = 0;
        auto add_run = [&](uint32_t start, uint32_t stop) {
            if (stop - start >= k) {
                plan.pieces.push_back(Piece{(uint32_t)n_valid, start, stop - start - k + 1});
                n_valid += stop - start - k + 1;
            }
        };
        for (uint32_t iv = b.rec_inv_off[r]; iv < b.rec_inv_off[r + 1]; ++iv) {
            add_run(s, b.inv_start[iv]);
            s = b.inv_start[iv] + b.inv_len[iv];
        }
        add_run(s, L);
        plan.n_kmers += n_valid;
        if (plan.pieces.size() > 0xFFFFFFF0ull) fail_runtime("too many valid runs");
        if (n_valid < w) {  // minimizer.cpp:56-58 and "fewer than w valid k-mers": nothing to emit
            plan.pieces.resize(piece_lo);
            continue;
        }
        const uint64_t n_win = n_valid - w + 1;
        plan.n_windows += n_win;
        size_t pc = piece_lo;
        for (uint64_t a0 = 0; a0 < n_win; a0 += tw) {
            const uint64_t a1 = std::min<uint64_t>(a0 + tw, n_win);
            const uint64_t e0 = a0 ? a0 - 1 : 0;
            const uint64_t nk = a1 + w - 1 - e0;
            while ((uint64_t)plan.pieces[pc].kidx + plan.pieces[pc].n <= e0) ++pc;
            size_t pe = pc;
            while (pe + 1 < plan.pieces.size() && plan.pieces[pe + 1].kidx < e0 + nk) ++pe;
            plan.tiles.push_back(Tile{(uint32_t)r, (uint32_t)e0, (uint32_t)nk, (uint32_t)pc,
                                      (uint32_t)(pe - pc + 1), a0 == 0 ? 1u : 0u});
        }
    }
    if (plan.tiles.size() > 0x7FFFFFFFull) fail_runtime("too many sketch tiles");
    return plan;
}

}  // namespace sw

sw_batch::~sw_batch()
{
    if (words) sw::free_host(words, pinned);
}
