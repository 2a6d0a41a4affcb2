// api.cu -- the C ABI of libseqwin_b200.so (see include/seqwin_b200.h for the mapping onto the
// reference's pybind11 module, cpp/src/bindings/python_bindings.cpp:43-169).
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <unordered_map>

#include "device.h"
#include "ingest.h"
#include "nthash.h"

namespace sw {

namespace {
thread_local std::string g_last_error;
std::once_flag g_init_flag;
int g_sm_count = 0;
bool g_have_device = false;
std::string g_init_error;

void do_init()
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_init_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "0 devices");
        cudaGetLastError();
        return;
    }
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { g_init_error = "cudaGetDevice failed"; return; }
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) { g_init_error = "cudaGetDeviceProperties failed"; return; }
    if (p.major < 10) {
        g_init_error = "libseqwin_b200 is built for sm_100a only; device is sm_" + std::to_string(p.major) +
                       std::to_string(p.minor);
        return;
    }
    g_sm_count = p.multiProcessorCount;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;  // keep freed blocks cached in the pool
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    g_have_device = true;
}
}  // namespace

void init_device_once()
{
    std::call_once(g_init_flag, do_init);
    if (!g_have_device) fail_runtime(g_init_error);
}
int sm_count()
{
    init_device_once();
    return g_sm_count;
}

namespace {
void* raw_alloc_host(size_t bytes, bool* pinned)
{
    std::call_once(g_init_flag, do_init);
    void* p = nullptr;
    if (g_have_device && cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) == cudaSuccess) {
        *pinned = true;
        return p;
    }
    cudaGetLastError();
    *pinned = false;  // host-only tooling (no device): pageable memory, nothing will be copied
    return aligned_alloc(64, ((bytes ? bytes : 1) + 63) / 64 * 64);
}
void raw_free_host(void* p, bool pinned)
{
    if (pinned) cudaFreeHost(p);
    else free(p);
}
std::mutex g_batch_mu;
std::unordered_map<void*, HostBuf> g_batch_bufs;  // packed batches drawn from the pinned pool
}  // namespace

// Packed batches (ingest.cpp) live in recycled pinned buffers too: page-locking hundreds of MB per
// call would cost more than parsing the FASTA.
void* alloc_host(size_t bytes, bool* pinned)
{
    HostBuf b = host_pool_get(bytes ? bytes : 1);
    *pinned = b.pinned;
    std::lock_guard<std::mutex> lk(g_batch_mu);
    g_batch_bufs[b.p] = b;
    return b.p;
}
void free_host(void* p, bool pinned)
{
    (void)pinned;
    HostBuf b;
    {
        std::lock_guard<std::mutex> lk(g_batch_mu);
        auto it = g_batch_bufs.find(p);
        if (it == g_batch_bufs.end()) return;
        b = it->second;
        g_batch_bufs.erase(it);
    }
    host_pool_put(b);
}

namespace {
struct Arena {
    struct Slab { char* p; size_t bytes; };
    std::vector<Slab> slabs;
    size_t cur = 0, off = 0;
    ~Arena() { for (auto& sl : slabs) cudaFree(sl.p); }
    void* alloc(size_t bytes)
    {
        bytes = (bytes + 255) & ~(size_t)255;
        while (cur < slabs.size() && off + bytes > slabs[cur].bytes) { ++cur; off = 0; }
        if (cur == slabs.size()) {
            Slab sl;
            sl.bytes = std::max(bytes, (size_t)256 << 20);
            cudaError_t err = cudaMalloc((void**)&sl.p, sl.bytes);
            if (err == cudaErrorMemoryAllocation) {
                // the stream-ordered pool keeps what earlier (larger) builds freed: give that back and retry
                cudaGetLastError();
                cudaDeviceSynchronize();
                int dev = 0;
                cudaMemPool_t pool;
                if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess)
                    cudaMemPoolTrimTo(pool, 0);
                err = cudaMalloc((void**)&sl.p, sl.bytes);
            }
            SW_CUDA(err);
            slabs.push_back(sl);
            off = 0;
        }
        void* r = slabs[cur].p + off;
        off += bytes;
        return r;
    }
    void trim()
    {
        for (auto& sl : slabs) cudaFree(sl.p);
        slabs.clear();
        cur = 0;
        off = 0;
    }
    void reset()
    {
        size_t held = 0;
        for (auto& sl : slabs) held += sl.bytes;
        size_t free_b = 0, total_b = 0;
        if (held > ((size_t)64 << 30) && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && held > total_b / 5 * 3) {
            // a build that needed most of the device: the next one may be shaped differently (outputs come from the
            // stream-ordered pool), so the slabs go back to the driver instead of staying reserved.  (Re-allocating
            // them costs tens of milliseconds per build, which is why smaller arenas are kept.)
            cudaDeviceSynchronize();
            trim();
            return;
        }
        if (slabs.size() > 1) {  // grew during the last build: come back as one slab next time
            size_t total = 0;
            for (auto& sl : slabs) { total += sl.bytes; cudaFree(sl.p); }
            slabs.clear();
            Slab sl;
            sl.bytes = total;
            if (cudaMalloc((void**)&sl.p, sl.bytes) == cudaSuccess) slabs.push_back(sl);
            else cudaGetLastError();
        }
        cur = 0;
        off = 0;
    }
};
Arena& tls_arena()
{
    static thread_local Arena a;
    return a;
}
}  // namespace

namespace {
__global__ void readback_kernel(const unsigned long long* __restrict__ src, unsigned long long* __restrict__ dst, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}
struct ReadbackRing {
    unsigned long long* host = nullptr;
    size_t cap = 0, off = 0;
    std::vector<unsigned long long*> retired;  // outgrown buffers, kept until the next build starts
    ~ReadbackRing()
    {
        if (host) cudaFreeHost(host);
        for (auto* p : retired) cudaFreeHost(p);
    }
    void release_retired()
    {
        for (auto* p : retired) cudaFreeHost(p);
        retired.clear();
    }
    unsigned long long* take(size_t count)
    {
        if (!host || count > cap / 2) {  // first use, or a request the ring cannot hold comfortably
            size_t ncap = (size_t)1 << 20;  // 8 MB of u64 slots
            while (ncap / 4 < count) ncap *= 2;
            if (ncap > cap) {
                if (host) retired.push_back(host);  // earlier slots may still be read after the next sync
                host = nullptr;
                SW_CUDA(cudaHostAlloc((void**)&host, ncap * sizeof(unsigned long long), cudaHostAllocMapped));
                cap = ncap;
                off = 0;
            }
        }
        if (off + count > cap) off = 0;  // wrap: earlier slots were consumed long ago
        unsigned long long* r = host + off;
        off += count;
        return r;
    }
};
ReadbackRing& tls_readback()
{
    static thread_local ReadbackRing r;
    return r;
}
}  // namespace

size_t arena_bytes()
{
    size_t t = 0;
    for (auto& sl : tls_arena().slabs) t += sl.bytes;
    return t;
}
void* arena_alloc(size_t bytes) { return tls_arena().alloc(bytes); }
ArenaMark arena_mark() { return ArenaMark{tls_arena().cur, tls_arena().off}; }
void arena_release(const ArenaMark& m)
{
    tls_arena().cur = m.slab;
    tls_arena().off = m.off;
}
size_t arena_free_bytes()
{
    const Arena& a = tls_arena();
    size_t t = 0;
    for (size_t i = a.cur; i < a.slabs.size(); ++i) t += a.slabs[i].bytes - (i == a.cur ? a.off : 0);
    return t;
}
void arena_reset()
{
    tls_arena().reset();
    tls_readback().release_retired();
}
void arena_trim() { tls_arena().trim(); }
namespace {
thread_local bool g_low_memory = false;
}
void set_low_memory(bool on) { g_low_memory = on; }
bool low_memory() { return g_low_memory; }

const unsigned long long* readback_u64(const unsigned long long* d_src, size_t count, cudaStream_t s)
{
    unsigned long long* h = tls_readback().take(count ? count : 1);
    if (count) {
        unsigned long long* d_dst = nullptr;
        SW_CUDA(cudaHostGetDevicePointer((void**)&d_dst, h, 0));
        readback_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(d_src, d_dst, count);
        SW_CUDA(cudaGetLastError());
    }
    return h;
}

namespace {
std::mutex g_pool_mu;
std::vector<HostBuf> g_pool_free;
size_t g_pool_cached = 0;
constexpr size_t kPoolMaxCached = (size_t)16 << 30;
}  // namespace

// Pinned allocations are expensive (page locking), so graph export buffers are recycled.
HostBuf host_pool_get(size_t bytes)
{
    if (bytes == 0) return HostBuf{};
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        int best = -1;
        for (size_t i = 0; i < g_pool_free.size(); ++i)
            if (g_pool_free[i].bytes >= bytes && g_pool_free[i].bytes <= bytes * 2 + (1u << 20) &&
                (best < 0 || g_pool_free[i].bytes < g_pool_free[best].bytes))
                best = (int)i;
        if (best >= 0) {
            HostBuf b = g_pool_free[best];
            g_pool_free.erase(g_pool_free.begin() + best);
            g_pool_cached -= b.bytes;
            return b;
        }
    }
    HostBuf b;
    b.bytes = bytes + bytes / 8 + 4096;
    b.p = raw_alloc_host(b.bytes, &b.pinned);
    if (!b.p) fail_runtime("host allocation of graph export buffer failed");
    return b;
}
void host_pool_put(HostBuf& b)
{
    if (!b.p) return;
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (g_pool_cached + b.bytes <= kPoolMaxCached) {
        g_pool_cached += b.bytes;
        g_pool_free.push_back(b);
    } else {
        raw_free_host(b.p, b.pinned);
    }
    b = HostBuf{};
}

namespace {

struct StreamGuard {
    cudaStream_t s = nullptr;
    StreamGuard() { SW_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)); }
    ~StreamGuard() { if (s) cudaStreamDestroy(s); }
};

thread_local cudaStream_t g_stream_override = nullptr;
thread_local bool g_have_override = false;

cudaStream_t lib_stream()
{
    if (g_have_override) return g_stream_override;
    static thread_local std::unique_ptr<StreamGuard> g;
    if (!g) g.reset(new StreamGuard());
    return g->s;
}

cudaStream_t copy_stream()
{
    static thread_local std::unique_ptr<StreamGuard> g;
    if (!g) g.reset(new StreamGuard());
    return g->s;
}

void copy_meta(const sw_batch& src, sw_batch& dst)
{
    dst.words = nullptr;
    dst.n_words = src.n_words;
    dst.rec_word_off = src.rec_word_off;
    dst.rec_len = src.rec_len;
    dst.rec_inv_off = src.rec_inv_off;
    dst.inv_start = src.inv_start;
    dst.inv_len = src.inv_len;
    dst.record_offsets = src.record_offsets;
    dst.ids = src.ids;
    dst.n_bases = src.n_bases;
}

std::vector<uint32_t> record_assembly_map(const std::vector<uint32_t>& offsets)
{
    const size_t A = offsets.size() - 1;
    std::vector<uint32_t> m(offsets.back());
    for (size_t a = 0; a < A; ++a)
        for (uint32_t r = offsets[a]; r < offsets[a + 1]; ++r) m[r] = (uint32_t)a;
    return m;
}

void check_kw(uint32_t k, uint32_t w)
{
    // k < 3 is undefined behaviour in the reference (nthash_kmer.hpp:26, unsigned k-3) and k is
    // a uint16 there (hashing_internals.hpp:10); w = 0 never selects anything.
    if (k < 3 || k > 65535) fail_runtime("kmerlen must be in [3, 65535]");
    if (w < 1) fail_runtime("windowsize must be >= 1");
}

// small per-record tables (everything but the packed bases)
void upload_tables(const sw_batch& b, sw_dev_batch& d, cudaStream_t s)
{
    const size_t R = b.rec_len.size();
    d.rec_word_off.alloc(R, s);
    d.rec_asm.alloc(R, s);
    d.rec_len.alloc(R, s);
    d.rec_inv_off.alloc(R + 1, s);
    d.inv_start.alloc(b.inv_start.size(), s);
    d.inv_len.alloc(b.inv_len.size(), s);
    const std::vector<uint32_t> ra = record_assembly_map(b.record_offsets);
    SW_CUDA(cudaMemcpyAsync(d.rec_inv_off.p, b.rec_inv_off.data(), (R + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    if (!b.inv_start.empty()) {
        SW_CUDA(cudaMemcpyAsync(d.inv_start.p, b.inv_start.data(), b.inv_start.size() * 4, cudaMemcpyHostToDevice, s));
        SW_CUDA(cudaMemcpyAsync(d.inv_len.p, b.inv_len.data(), b.inv_len.size() * 4, cudaMemcpyHostToDevice, s));
    }
    if (R) {
        SW_CUDA(cudaMemcpyAsync(d.rec_len.p, b.rec_len.data(), R * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
        SW_CUDA(cudaMemcpyAsync(d.rec_word_off.p, b.rec_word_off.data(), R * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
        SW_CUDA(cudaMemcpyAsync(d.rec_asm.p, ra.data(), R * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    }
    SW_CUDA(cudaStreamSynchronize(s));  // `ra` is a pageable temporary
}

sw_dev_batch* dev_upload(const sw_batch& b)
{
    init_device_once();
    auto d = std::make_unique<sw_dev_batch>();
    cudaStream_t s = lib_stream();
    d->stream = s;
    copy_meta(b, d->meta);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, s);
    d->words.alloc(b.n_words, s);
    SW_CUDA(cudaMemcpyAsync(d->words.p, b.words, b.n_words * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    upload_tables(b, *d, s);
    cudaEventRecord(e1, s);
    SW_CUDA(cudaStreamSynchronize(s));
    cudaEventElapsedTime(&d->h2d_ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return d.release();
}

// Optional host callback fired once kmers + nodes are final on the device (multi-GPU builds start
// their node / k-mer exchange from it while the edge stage is still running).
// Thread-local like the stream override and the arena: the hook belongs to the builder thread that set it.
thread_local sw_nodes_ready_fn g_nodes_ready = nullptr;
thread_local void* g_nodes_ready_user = nullptr;

void fire_nodes_ready(sw_graph* g)
{
    if (!g_nodes_ready) return;
    g->on_device = true;
    g->n_kmers = g->dev.n_kmers;
    g->n_nodes = g->dev.n_nodes;
    g->n_edges = 0;
    g_nodes_ready(g_nodes_ready_user, g);
}

// Classes of the batch's assemblies for a build that scores its nodes on the way (ScoreArgs): the
// checks of filter.cpp:33-60 that concern is_targets, then the flags go to the device.
struct ScoreUpload {
    DevBuf<uint8_t> d_t;
    ScoreArgs args{};
    // shard: one rank's part of a multi-GPU build -- any mix of classes, counts only
    ScoreUpload(const uint8_t* is_targets, size_t n_assemblies, size_t n_batch_assemblies, cudaStream_t s,
                bool shard = false)
    {
        if (n_assemblies != n_batch_assemblies) fail_value("len(is_targets) must equal the number of assemblies");
        size_t n_t = 0;
        for (size_t i = 0; i < n_assemblies; ++i) n_t += is_targets[i] ? 1 : 0;
        if (!shard) {
            if (!n_t) fail_value("is_targets must contain at least one target assembly");
            if (n_t == n_assemblies) fail_value("is_targets must contain at least one non-target assembly");
        }
        d_t.alloc(n_assemblies, s, true);
        if (n_assemblies) SW_CUDA(cudaMemcpyAsync(d_t.p, is_targets, n_assemblies, cudaMemcpyHostToDevice, s));
        args = ScoreArgs{d_t.p, n_t ? 1.0 / (double)n_t : 0.0, n_assemblies > n_t ? 1.0 / (double)(n_assemblies - n_t) : 0.0,
                         shard};
    }
};

sw_graph* dev_build(const sw_dev_batch& d, uint32_t k, uint32_t w, sw_stage_times* t, uint32_t rec_base = 0,
                    const uint8_t* is_targets = nullptr, size_t n_assemblies = 0, bool shard = false)
{
    init_device_once();
    check_kw(k, w);
    cudaStream_t s = d.stream;
    arena_reset();
    std::unique_ptr<ScoreUpload> score;
    if (is_targets) score = std::make_unique<ScoreUpload>(is_targets, n_assemblies, d.meta.record_offsets.size() - 1, s, shard);
    auto g = std::make_unique<sw_graph>();
    g->stream = s;
    g->record_offsets = d.meta.record_offsets;
    g->ids = d.meta.ids;

    cudaEvent_t e0, e1, e2;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventCreate(&e2);
    cudaEventRecord(e0, s);
    const auto host_t0 = std::chrono::steady_clock::now();
    DevPlan plan = make_plan(d, k, w, s);
    const float plan_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - host_t0).count();
    SketchStream st;
    run_sketch(d.words.p, d.rec_word_off.p, plan, k, w, rec_base, s, st);
    cudaEventRecord(e1, s);
    GraphTimes gt;
    const std::function<void()> after_nodes = [&] { fire_nodes_ready(g.get()); };
    build_graph(st, d.rec_asm.p, rec_base, s, g->dev, &gt, &after_nodes, score ? &score->args : nullptr);
    cudaEventRecord(e2, s);
    SW_CUDA(cudaStreamSynchronize(s));
    g->on_device = true;
    g->n_kmers = g->dev.n_kmers;
    g->n_nodes = g->dev.n_nodes;
    g->n_edges = g->dev.n_edges;
    if (t) {
        memset(t, 0, sizeof(*t));
        cudaEventElapsedTime(&t->sketch_ms, e0, e1);
        cudaEventElapsedTime(&t->total_ms, e0, e2);
        t->plan_ms = plan_ms;
        t->sketch_kernel_ms = st.kernel_ms;
        t->reorder_ms = st.reorder_ms;
        t->sort_nodes_ms = gt.sort_nodes_ms;
        t->nodes_ms = gt.nodes_ms;
        t->edges_ms = gt.edges_ms;
        t->n_bases = d.meta.n_bases;
        t->n_kmers = g->dev.n_kmers;
        t->n_nodes = g->dev.n_nodes;
        t->n_edges = g->dev.n_edges;
        t->n_tiles = plan.n_tiles;
        t->sketch_launches = st.launches;
        t->total_launches = st.launches + gt.launches;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaEventDestroy(e2);
    return g.release();
}

void graph_to_host(sw_graph& g)
{
    if (g.on_host) return;
    cudaStream_t s = g.stream;
    g.h_kmers = host_pool_get(g.n_kmers * sizeof(sw_kmer));
    g.h_nodes = host_pool_get(g.n_nodes * sizeof(sw_node));
    g.h_edges = host_pool_get(g.n_edges * sizeof(sw_edge));
    if (g.n_kmers)
        SW_CUDA(cudaMemcpyAsync(g.h_kmers.p, g.dev.kmers.p, g.n_kmers * sizeof(sw_kmer), cudaMemcpyDeviceToHost, s));
    if (g.n_nodes)
        SW_CUDA(cudaMemcpyAsync(g.h_nodes.p, g.dev.nodes.p, g.n_nodes * sizeof(sw_node), cudaMemcpyDeviceToHost, s));
    if (g.n_edges)
        SW_CUDA(cudaMemcpyAsync(g.h_edges.p, g.dev.edges.p, g.n_edges * sizeof(sw_edge), cudaMemcpyDeviceToHost, s));
    SW_CUDA(cudaStreamSynchronize(s));
    g.on_host = true;
}

// n_tar / n_neg / penalty of a device-resident graph (validation of filter.cpp:33-60 included).
// Returns the kernel's CUDA-event time in ms.
float penalty_on_device(DevGraph& dg, const std::vector<uint32_t>& offsets, const uint8_t* is_targets,
                        size_t n_assemblies, cudaStream_t s)
{
    if (offsets.size() != n_assemblies + 1) fail_value("len(record_offsets) must equal len(is_targets) + 1");
    if (offsets.empty() || offsets[0] != 0) fail_value("record_offsets must start with 0");
    size_t n_t = 0, n_n = 0;
    for (size_t i = 0; i < n_assemblies; ++i) {
        if (offsets[i + 1] < offsets[i]) fail_value("record_offsets must be nondecreasing");
        if (is_targets[i]) ++n_t; else ++n_n;
    }
    if (!n_t) fail_value("is_targets must contain at least one target assembly");
    if (!n_n) fail_value("is_targets must contain at least one non-target assembly");
    if (dg.n_nodes == 0) return 0.f;
    const std::vector<uint32_t> ra = record_assembly_map(offsets);
    DevBuf<uint32_t> d_ra(ra.size(), s, true);
    DevBuf<uint8_t> d_t(n_assemblies, s, true);
    if (!ra.empty()) SW_CUDA(cudaMemcpyAsync(d_ra.p, ra.data(), ra.size() * 4, cudaMemcpyHostToDevice, s));
    SW_CUDA(cudaMemcpyAsync(d_t.p, is_targets, n_assemblies, cudaMemcpyHostToDevice, s));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, s);
    const uint32_t bad = run_penalty(dg.kmers.p, dg.n_kmers, dg.nodes.p, dg.n_nodes, d_ra.p, offsets.back(), d_t.p,
                                     1.0 / (double)n_t, 1.0 / (double)n_n, s);
    cudaEventRecord(e1, s);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (bad & 4u) fail_value("node [start, stop) range lies outside kmers");
    if (bad & 1u) fail_value("record_idx is outside record_offsets range");
    if (bad & 2u) fail_value("record_idx must be nondecreasing within each node range");
    return ms;
}

// End-to-end build from a pinned host batch with the copies overlapped with the kernels:
//   copy stream:    bases slice 0 | slice 1 | ... | slice C-1            kmers+nodes D2H
//   compute stream: tables, plan  | sketch(slice 0) | sketch(slice 1) ... sort, nodes | edges | edges D2H
sw_graph* build_pipelined(const sw_batch& b, uint32_t k, uint32_t w, sw_stage_times* t, bool to_host, uint32_t rec_base = 0,
                          const uint8_t* is_targets = nullptr, size_t n_assemblies = 0, bool shard = false)
{
    init_device_once();
    check_kw(k, w);
    cudaStream_t s = lib_stream(), cs = copy_stream();
    arena_reset();
    auto d = std::make_unique<sw_dev_batch>();
    d->stream = s;
    copy_meta(b, d->meta);
    auto g = std::make_unique<sw_graph>();
    g->stream = s;
    g->record_offsets = b.record_offsets;
    g->ids = b.ids;

    constexpr int kMaxChunks = 8;
    cudaEvent_t ev[kMaxChunks + 9];
    for (auto& e : ev) cudaEventCreateWithFlags(&e, cudaEventDefault);
    cudaEvent_t& e_begin = ev[kMaxChunks], &e_alloc = ev[kMaxChunks + 1], &e_h2d0 = ev[kMaxChunks + 2],
                 &e_h2d1 = ev[kMaxChunks + 3], &e_nodes = ev[kMaxChunks + 4], &e_d2h0 = ev[kMaxChunks + 5],
                 &e_d2h1 = ev[kMaxChunks + 6], &e_end = ev[kMaxChunks + 7], &e_kmers = ev[kMaxChunks + 8];
    struct EvGuard {
        cudaEvent_t* e; size_t n;
        ~EvGuard() { for (size_t i = 0; i < n; ++i) cudaEventDestroy(e[i]); }
    } ev_guard{ev, sizeof(ev) / sizeof(ev[0])};
    // An exception after copies were enqueued must not hand pinned destinations back to the pool (or
    // free device sources) while the copy engine still uses them: drain both streams before `g`, `d`
    // and the events unwind (declared last, destroyed first).
    struct DrainOnUnwind {
        cudaStream_t a, b;
        bool armed = true;
        ~DrainOnUnwind()
        {
            if (!armed) return;
            cudaStreamSynchronize(a);
            cudaStreamSynchronize(b);
        }
    } drain{cs, s};

    cudaEventRecord(e_begin, s);
    // the small tables go first: behind the bulk copies they would wait for the whole upload
    upload_tables(b, *d, s);
    d->words.alloc(b.n_words, s);
    cudaEventRecord(e_alloc, s);
    SW_CUDA(cudaStreamWaitEvent(cs, e_alloc, 0));

    // slices of whole records, roughly equal in words
    const size_t R = b.rec_len.size();
    std::vector<size_t> rec_cut{0};
    const size_t data_words = b.n_words - kTailPadWords;
    const int n_chunks = (int)std::max<size_t>(1, std::min<size_t>(kMaxChunks, data_words / (4u << 20)));  // >= 16 MB each
    for (int c = 1; c < n_chunks; ++c) {
        const uint64_t target = (uint64_t)data_words * c / n_chunks;
        size_t r = std::lower_bound(b.rec_word_off.begin(), b.rec_word_off.end(), target) - b.rec_word_off.begin();
        if (r > rec_cut.back() && r < R) rec_cut.push_back(r);
    }
    rec_cut.push_back(R);
    const size_t C = rec_cut.size() - 1;
    cudaEventRecord(e_h2d0, cs);
    for (size_t c = 0; c < C; ++c) {
        const size_t w0 = c == 0 ? 0 : b.rec_word_off[rec_cut[c]];
        const size_t w1 = c + 1 == C ? b.n_words : b.rec_word_off[rec_cut[c + 1]];
        if (w1 > w0)
            SW_CUDA(cudaMemcpyAsync(d->words.p + w0, b.words + w0, (w1 - w0) * sizeof(uint32_t), cudaMemcpyHostToDevice, cs));
        cudaEventRecord(ev[c], cs);
    }
    cudaEventRecord(e_h2d1, cs);

    const auto host_t0 = std::chrono::steady_clock::now();
    DevPlan plan = make_plan(*d, k, w, s, true);
    const float plan_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - host_t0).count();
    std::vector<SketchChunk> chunks(C);
    for (size_t c = 0; c < C; ++c) {
        const uint32_t lo = plan.rec_tile_off.empty() ? 0 : (uint32_t)plan.rec_tile_off[rec_cut[c]];
        const uint32_t hi = plan.rec_tile_off.empty() ? 0
                            : (rec_cut[c + 1] >= R ? plan.n_tiles : (uint32_t)plan.rec_tile_off[rec_cut[c + 1]]);
        chunks[c] = SketchChunk{lo, hi, ev[c]};
    }
    SketchStream st;
    run_sketch(d->words.p, d->rec_word_off.p, plan, k, w, rec_base, s, st, &chunks);

    bool d2h_started = false;
    const float penalty_ms = 0;   // scoring is part of the node stage
    std::unique_ptr<ScoreUpload> score;
    if (is_targets) score = std::make_unique<ScoreUpload>(is_targets, n_assemblies, b.record_offsets.size() - 1, s, shard);
    const std::function<void()> after_nodes = [&] {
        if (to_host) {
            // kmers + nodes are final (scored in the node stage when the classes were given)
            g->n_kmers = g->dev.n_kmers;
            g->n_nodes = g->dev.n_nodes;
            g->h_kmers = host_pool_get(g->n_kmers * sizeof(sw_kmer));
            g->h_nodes = host_pool_get(g->n_nodes * sizeof(sw_node));
            cudaEventRecord(e_kmers, s);
            SW_CUDA(cudaStreamWaitEvent(cs, e_kmers, 0));
            cudaEventRecord(e_d2h0, cs);
            if (g->n_kmers)
                SW_CUDA(cudaMemcpyAsync(g->h_kmers.p, g->dev.kmers.p, g->n_kmers * sizeof(sw_kmer), cudaMemcpyDeviceToHost, cs));
            d2h_started = true;
        }
        fire_nodes_ready(g.get());
        if (!to_host) return;
        cudaEventRecord(e_nodes, s);
        SW_CUDA(cudaStreamWaitEvent(cs, e_nodes, 0));
        if (g->n_nodes)
            SW_CUDA(cudaMemcpyAsync(g->h_nodes.p, g->dev.nodes.p, g->n_nodes * sizeof(sw_node), cudaMemcpyDeviceToHost, cs));
    };
    GraphTimes gt;
    build_graph(st, d->rec_asm.p, rec_base, s, g->dev, &gt, &after_nodes, score ? &score->args : nullptr);
    g->on_device = true;
    g->n_kmers = g->dev.n_kmers;
    g->n_nodes = g->dev.n_nodes;
    g->n_edges = g->dev.n_edges;
    if (to_host) {
        if (!d2h_started) {  // empty stream: nothing was copied yet
            g->h_kmers = host_pool_get(0);
            g->h_nodes = host_pool_get(0);
            cudaEventRecord(e_d2h0, cs);
        }
        g->h_edges = host_pool_get(g->n_edges * sizeof(sw_edge));
        // edges follow kmers + nodes on the copy stream (one D2H engine anyway)
        cudaEventRecord(e_nodes, s);
        SW_CUDA(cudaStreamWaitEvent(cs, e_nodes, 0));
        if (g->n_edges)
            SW_CUDA(cudaMemcpyAsync(g->h_edges.p, g->dev.edges.p, g->n_edges * sizeof(sw_edge), cudaMemcpyDeviceToHost, cs));
        cudaEventRecord(e_d2h1, cs);
        SW_CUDA(cudaStreamWaitEvent(s, e_d2h1, 0));
    }
    cudaEventRecord(e_end, s);
    SW_CUDA(cudaStreamSynchronize(s));
    SW_CUDA(cudaStreamSynchronize(cs));
    drain.armed = false;
    if (to_host) {
        g->on_host = true;
        g->dev = DevGraph();  // the device copies are no longer needed once the host arrays exist
        g->on_device = false;
    }
    if (t) {
        memset(t, 0, sizeof(*t));
        cudaEventElapsedTime(&t->h2d_ms, e_h2d0, e_h2d1);
        if (to_host) cudaEventElapsedTime(&t->d2h_ms, e_d2h0, e_d2h1);
        cudaEventElapsedTime(&t->total_ms, e_begin, e_end);
        t->plan_ms = plan_ms;
        t->penalty_ms = penalty_ms;
        t->sketch_kernel_ms = st.kernel_ms;
        t->reorder_ms = st.reorder_ms;
        t->sort_nodes_ms = gt.sort_nodes_ms;
        t->nodes_ms = gt.nodes_ms;
        t->edges_ms = gt.edges_ms;
        t->n_bases = b.n_bases;
        t->n_kmers = g->n_kmers;
        t->n_nodes = g->n_nodes;
        t->n_edges = g->n_edges;
        t->n_tiles = plan.n_tiles;
        t->sketch_launches = st.launches;
        t->total_launches = st.launches + gt.launches;
    }
    // d (device batch) is released here, stream-ordered after everything above
    return g.release();
}

template <typename F>
int guarded(F&& fn)
{
    try {
        fn();
        return SW_OK;
    } catch (const Error& e) {
        g_last_error = e.what();
        return e.code;
    } catch (const std::bad_alloc&) {
        g_last_error = "out of host memory";
        return SW_ERR_RUNTIME;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return SW_ERR_RUNTIME;
    } catch (...) {
        g_last_error = "unknown error";
        return SW_ERR_RUNTIME;
    }
}

}  // namespace
}  // namespace sw

using namespace sw;

extern "C" {

const char* sw_last_error(void) { return g_last_error.c_str(); }

int sw_device_info(int* sm, int* major, int* minor, size_t* hbm)
{
    return guarded([&] {
        init_device_once();
        int dev = 0;
        SW_CUDA(cudaGetDevice(&dev));
        cudaDeviceProp p;
        SW_CUDA(cudaGetDeviceProperties(&p, dev));
        if (sm) *sm = p.multiProcessorCount;
        if (major) *major = p.major;
        if (minor) *minor = p.minor;
        if (hbm) *hbm = p.totalGlobalMem;
    });
}

int sw_mem_stats(uint64_t out[4])
{
    return guarded([&] {
        init_device_once();
        int dev = 0;
        SW_CUDA(cudaGetDevice(&dev));
        out[0] = arena_bytes();
        out[1] = out[2] = 0;
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            uint64_t v = 0;
            if (cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemHigh, &v) == cudaSuccess) out[1] = v;
            if (cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemHigh, &v) == cudaSuccess) out[2] = v;
        }
        size_t fr = 0, tot = 0;
        SW_CUDA(cudaMemGetInfo(&fr, &tot));
        out[3] = tot - fr;
    });
}

int sw_measure_int_peak(double lane_ops_per_s[4])
{
    return guarded([&] {
        init_device_once();
        arena_reset();
        measure_int_peak(lane_ops_per_s, lib_stream());
    });
}

int sw_batch_from_fasta(const char* const* paths, size_t n_paths, uint32_t n_threads, sw_batch** out)
{
    return guarded([&] { *out = batch_from_fasta(paths, n_paths, n_threads); });
}

int sw_batch_from_memory(const uint8_t* const* seqs, const uint32_t* lens, const uint32_t* asm_of,
                         const char* const* ids, size_t n_records, size_t n_assemblies, uint32_t n_threads,
                         sw_batch** out)
{
    return guarded([&] { *out = batch_from_memory(seqs, lens, asm_of, ids, n_records, n_assemblies, n_threads); });
}

int sw_batch_concat(const sw_batch* const* parts, size_t n_parts, sw_batch** out)
{
    return guarded([&] { *out = batch_concat(parts, n_parts); });
}

size_t sw_batch_n_bases(const sw_batch* b) { return b->n_bases; }
size_t sw_batch_n_records(const sw_batch* b) { return b->rec_len.size(); }
size_t sw_batch_packed_bytes(const sw_batch* b) { return b->n_words * sizeof(uint32_t); }
int sw_batch_record_offsets(const sw_batch* b, int64_t* out, size_t n)
{
    return guarded([&] {
        if (n != b->record_offsets.size()) fail_value("record_offsets size mismatch");
        for (size_t i = 0; i < n; ++i) out[i] = b->record_offsets[i];
    });
}
void sw_batch_free(sw_batch* b) { delete b; }

int sw_dev_upload(const sw_batch* b, sw_dev_batch** out)
{
    return guarded([&] { *out = dev_upload(*b); });
}
void sw_dev_batch_free(sw_dev_batch* d)
{
    if (!d) return;
    cudaStream_t s = d->stream;
    delete d;
    if (s) cudaStreamSynchronize(s);
}

int sw_dev_build(const sw_dev_batch* d, uint32_t k, uint32_t w, sw_graph** out, sw_stage_times* t)
{
    return guarded([&] { *out = dev_build(*d, k, w, t); });
}

int sw_dev_build_scored(const sw_dev_batch* d, uint32_t k, uint32_t w, const uint8_t* is_targets, size_t n_assemblies,
                        sw_graph** out, sw_stage_times* t)
{
    return guarded([&] {
        if (!is_targets) fail_value("is_targets is required");
        *out = dev_build(*d, k, w, t, 0u, is_targets, n_assemblies);
    });
}

int sw_set_stream(void* stream)
{
    g_stream_override = static_cast<cudaStream_t>(stream);
    g_have_override = stream != nullptr;
    return SW_OK;
}

int sw_dev_build_ex(const sw_dev_batch* d, uint32_t k, uint32_t w, uint32_t rec_base, const uint8_t* is_targets,
                    size_t n_assemblies, sw_graph** out, sw_stage_times* t)
{
    return guarded([&] { *out = dev_build(*d, k, w, t, rec_base, is_targets, n_assemblies, /*shard=*/true); });
}

struct sw_routed {
    sw::RoutedStream rs;
    sw::SketchStream st;       // fused build: the stream waits here (scratch arena) between the histogram and the scatter
    int route_bits = 8;
    cudaStream_t stream = nullptr;
};

int sw_dev_sketch_route(const sw_dev_batch* d, uint32_t k, uint32_t w, uint32_t rec_base, sw_routed** out, sw_stage_times* t)
{
    return guarded([&] {
        init_device_once();
        check_kw(k, w);
        cudaStream_t s = d->stream;
        arena_reset();
        auto r = std::make_unique<sw_routed>();
        r->stream = s;
        cudaEvent_t e0, e1, e2;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventCreate(&e2);
        struct EvGuard {
            cudaEvent_t &a, &b, &c;
            ~EvGuard() { cudaEventDestroy(a); cudaEventDestroy(b); cudaEventDestroy(c); }
        } guard{e0, e1, e2};
        cudaEventRecord(e0, s);
        const auto host_t0 = std::chrono::steady_clock::now();
        DevPlan plan = make_plan(*d, k, w, s);
        const float plan_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - host_t0).count();
        SketchStream st;
        run_sketch(d->words.p, d->rec_word_off.p, plan, k, w, rec_base, s, st);
        cudaEventRecord(e1, s);
        route_stream(st, s, r->rs);
        cudaEventRecord(e2, s);
        SW_CUDA(cudaStreamSynchronize(s));
        if (r->rs.zero_key) fail_runtime("a minimizer hash is 0: the routed multi-GPU build cannot mark 'no neighbour' (use the merge-based one)");
        if (t) {
            memset(t, 0, sizeof(*t));
            cudaEventElapsedTime(&t->sketch_ms, e0, e1);
            cudaEventElapsedTime(&t->sort_nodes_ms, e1, e2);   // the routing pass
            cudaEventElapsedTime(&t->total_ms, e0, e2);
            t->plan_ms = plan_ms;
            t->sketch_kernel_ms = st.kernel_ms;
            t->reorder_ms = st.reorder_ms;
            t->n_bases = d->meta.n_bases;
            t->n_kmers = st.n;
            t->n_tiles = plan.n_tiles;
            t->sketch_launches = st.launches;
            t->total_launches = st.launches + r->rs.launches;
        }
        *out = r.release();
    });
}

// ---- fused routing: the scatter of the routing pass writes into the owners' memory (peer access over NVLink) ----------
int sw_peer_alloc(size_t bytes, void** dev_ptr, void* ipc_handle)
{
    return guarded([&] {
        init_device_once();
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the handle travels as 64 bytes");
        void* p = nullptr;
        SW_CUDA(cudaMalloc(&p, bytes ? bytes : 1));
        cudaIpcMemHandle_t h;
        const cudaError_t err = cudaIpcGetMemHandle(&h, p);
        if (err != cudaSuccess) {
            cudaFree(p);
            SW_CUDA(err);
        }
        memcpy(ipc_handle, &h, sizeof(h));
        *dev_ptr = p;
    });
}

int sw_peer_open(const void* ipc_handle, void** dev_ptr)
{
    return guarded([&] {
        init_device_once();
        cudaIpcMemHandle_t h;
        memcpy(&h, ipc_handle, sizeof(h));
        SW_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    });
}

int sw_peer_close(void* dev_ptr)
{
    return guarded([&] { SW_CUDA(cudaIpcCloseMemHandle(dev_ptr)); });
}

int sw_peer_free(void* dev_ptr)
{
    return guarded([&] {
        SW_CUDA(cudaDeviceSynchronize());
        SW_CUDA(cudaFree(dev_ptr));
    });
}

int sw_dev_sketch_hist(const sw_dev_batch* d, uint32_t k, uint32_t w, uint32_t rec_base, uint32_t route_bits, sw_routed** out,
                       uint64_t* byte_counts, sw_stage_times* t)
{
    return guarded([&] {
        init_device_once();
        check_kw(k, w);
        if (route_bits < 1 || route_bits > 8) fail_value("route_bits must be 1..8");
        cudaStream_t s = d->stream;
        arena_reset();
        auto r = std::make_unique<sw_routed>();
        r->stream = s;
        r->route_bits = (int)route_bits;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        struct EvGuard {
            cudaEvent_t &a, &b;
            ~EvGuard() { cudaEventDestroy(a); cudaEventDestroy(b); }
        } guard{e0, e1};
        cudaEventRecord(e0, s);
        const auto host_t0 = std::chrono::steady_clock::now();
        DevPlan plan = make_plan(*d, k, w, s);
        const float plan_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - host_t0).count();
        run_sketch(d->words.p, d->rec_word_off.p, plan, k, w, rec_base, s, r->st);
        if (r->st.n > 0xFFFFFFFFull) fail_runtime("more than 2^32-1 minimizers on one device");
        DevBuf<unsigned long long> counts(256, s, true);
        route_histogram(r->st.keys.p, r->st.n, (int)route_bits, counts.p, s);
        cudaEventRecord(e1, s);
        const unsigned long long* h = readback_u64(counts.p, 256, s);
        SW_CUDA(cudaStreamSynchronize(s));
        std::copy(h, h + 256, byte_counts);
        r->rs.n = r->st.n;
        r->rs.items_per_key = r->st.items_per_key;
        r->rs.pairs_per_edge = r->st.pairs_per_edge;
        if (t) {
            memset(t, 0, sizeof(*t));
            cudaEventElapsedTime(&t->sketch_ms, e0, e1);
            t->total_ms = t->sketch_ms;
            t->plan_ms = plan_ms;
            t->sketch_kernel_ms = r->st.kernel_ms;
            t->reorder_ms = r->st.reorder_ms;
            t->n_bases = d->meta.n_bases;
            t->n_kmers = r->st.n;
            t->n_tiles = plan.n_tiles;
            t->sketch_launches = r->st.launches;
            t->total_launches = r->st.launches + 1;
        }
        *out = r.release();
    });
}

int sw_routed_scatter(sw_routed* r, const void* const* route_ptrs, const uint64_t* byte_base, float* kernel_ms)
{
    return guarded([&] {
        cudaStream_t s = r->stream;
        DevBuf<unsigned long long> d_base(256, s, true);
        DevBuf<uint64_t*> d_route(4 * 256, s, true);
        DevBuf<unsigned long long> flag(1, s, true);
        SW_CUDA(cudaMemcpyAsync(d_base.p, byte_base, 256 * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
        SW_CUDA(cudaMemcpyAsync(d_route.p, route_ptrs, 4 * 256 * sizeof(void*), cudaMemcpyHostToDevice, s));
        SW_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(unsigned long long), s));
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, s);
        route_scatter(r->st.keys.p, r->st.vals.p, r->st.n, r->route_bits, d_base.p, d_route.p, reinterpret_cast<unsigned int*>(flag.p), s);
        cudaEventRecord(e1, s);
        const unsigned long long* hz = readback_u64(flag.p, 1, s);
        SW_CUDA(cudaStreamSynchronize(s));
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        if (kernel_ms) *kernel_ms = ms;
        if ((*hz & 0xFFFFFFFFull) != 0)
            fail_runtime("a minimizer hash is 0: the routed multi-GPU build cannot mark 'no neighbour' (use the merge-based one)");
    });
}

int sw_routed_info(const sw_routed* r, void** keys, void** vals, void** prev, void** next, uint64_t* n, uint64_t* byte_off,
                   double* stats)
{
    const uint64_t M = r->rs.n;
    if (keys) *keys = r->rs.set.p;
    if (vals) *vals = r->rs.set.p + M;
    if (prev) *prev = r->rs.set.p + 2 * M;
    if (next) *next = r->rs.set.p + 3 * M;
    if (n) *n = M;
    if (byte_off) std::copy(r->rs.byte_off, r->rs.byte_off + 257, byte_off);
    if (stats) {
        stats[0] = r->rs.items_per_key;
        stats[1] = r->rs.pairs_per_edge;
    }
    return SW_OK;
}

void sw_routed_free(sw_routed* r)
{
    if (!r) return;
    if (r->stream) cudaStreamSynchronize(r->stream);
    delete r;
}

int sw_dev_aggregate(const void* keys, const void* vals, const void* prev, const void* next, uint64_t n, uint32_t byte_lo,
                     uint32_t byte_hi, const uint32_t* record_offsets, size_t n_offsets, const uint8_t* is_targets,
                     size_t n_assemblies, double pairs_per_edge, const uint64_t* byte_off, uint32_t range_bits, sw_graph** out,
                     sw_stage_times* t)
{
    return guarded([&] {
        init_device_once();
        if (range_bits < 1 || range_bits > 8) fail_value("range_bits must be 1..8");
        if (byte_lo > byte_hi || byte_hi > (1u << range_bits)) fail_value("hash range must be 0 <= lo <= hi <= 2^range_bits");
        if (byte_off && (byte_off[0] != 0 || byte_off[byte_hi - byte_lo] != n)) fail_value("byte_off must run from 0 to n");
        if (n_offsets == 0 || record_offsets[0] != 0) fail_value("record_offsets must start with 0");
        for (size_t i = 0; i + 1 < n_offsets; ++i)
            if (record_offsets[i + 1] < record_offsets[i]) fail_value("record_offsets must be nondecreasing");
        cudaStream_t s = lib_stream();
        arena_reset();
        const std::vector<uint32_t> offsets(record_offsets, record_offsets + n_offsets);
        const std::vector<uint32_t> ra = record_assembly_map(offsets);
        DevBuf<uint32_t> d_ra(ra.size(), s, true);
        if (!ra.empty()) SW_CUDA(cudaMemcpyAsync(d_ra.p, ra.data(), ra.size() * 4, cudaMemcpyHostToDevice, s));
        std::unique_ptr<ScoreUpload> score;
        if (is_targets) score = std::make_unique<ScoreUpload>(is_targets, n_assemblies, n_offsets - 1, s, /*shard=*/false);
        auto g = std::make_unique<sw_graph>();
        g->stream = s;
        g->record_offsets = offsets;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, s);
        const NbrBuffers in{const_cast<uint64_t*>(static_cast<const uint64_t*>(keys)), const_cast<uint64_t*>(static_cast<const uint64_t*>(vals)),
                            const_cast<uint64_t*>(static_cast<const uint64_t*>(prev)), const_cast<uint64_t*>(static_cast<const uint64_t*>(next))};
        GraphTimes gt;
        aggregate_range(in, n, byte_lo, byte_hi, d_ra.p, s, g->dev, &gt, score ? &score->args : nullptr, pairs_per_edge, byte_off,
                        (int)range_bits);
        cudaEventRecord(e1, s);
        SW_CUDA(cudaStreamSynchronize(s));
        g->on_device = true;
        g->n_kmers = g->dev.n_kmers;
        g->n_nodes = g->dev.n_nodes;
        g->n_edges = g->dev.n_edges;
        if (t) {
            memset(t, 0, sizeof(*t));
            cudaEventElapsedTime(&t->total_ms, e0, e1);
            t->sort_nodes_ms = gt.sort_nodes_ms;
            t->nodes_ms = gt.nodes_ms;
            t->edges_ms = gt.edges_ms;
            t->n_kmers = g->dev.n_kmers;
            t->n_nodes = g->dev.n_nodes;
            t->n_edges = g->dev.n_edges;
            t->total_launches = gt.launches;
        }
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        *out = g.release();
    });
}

int sw_graph_finish_penalty(sw_graph* g, uint64_t n_targets, uint64_t n_non_targets)
{
    return guarded([&] {
        if (!g->on_device) fail_runtime("graph is not device resident");
        if (!n_targets) fail_value("is_targets must contain at least one target assembly");
        if (!n_non_targets) fail_value("is_targets must contain at least one non-target assembly");
        host_pool_put(g->h_kmers);
        host_pool_put(g->h_nodes);
        host_pool_put(g->h_edges);
        g->on_host = false;
        finish_penalty(g->dev.nodes.p, g->dev.n_nodes, 1.0 / (double)n_targets, 1.0 / (double)n_non_targets, g->stream);
        SW_CUDA(cudaStreamSynchronize(g->stream));
    });
}

int sw_graph_device_ptrs(sw_graph* g, void** kmers, void** nodes, void** edges)
{
    return guarded([&] {
        if (!g->on_device) fail_runtime("graph is not device resident");
        if (kmers) *kmers = g->dev.kmers.p;
        if (nodes) *nodes = g->dev.nodes.p;
        if (edges) *edges = g->dev.edges.p;
    });
}

int sw_graph_split(sw_graph* g, uint32_t n_parts, uint64_t* node_split, uint64_t* kmer_split, uint64_t* edge_split)
{
    return guarded([&] {
        if (!g->on_device) fail_runtime("graph is not device resident");
        if (n_parts == 0) fail_value("n_parts must be >= 1");
        std::vector<unsigned long long> h(3 * (size_t)(n_parts + 1));
        graph_split(g->dev, n_parts, h.data(), g->stream);
        for (uint32_t i = 0; i <= n_parts; ++i) {
            if (node_split) node_split[i] = h[i];
            if (kmer_split) kmer_split[i] = h[(n_parts + 1) + i];
            if (edge_split) edge_split[i] = h[2 * (size_t)(n_parts + 1) + i];
        }
    });
}

int sw_set_low_memory(int on)
{
    sw::set_low_memory(on != 0);
    return SW_OK;
}

int sw_trim_memory(void)
{
    return guarded([&] {
        init_device_once();
        SW_CUDA(cudaDeviceSynchronize());
        arena_trim();
        int dev = 0;
        cudaMemPool_t pool;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess)
            cudaMemPoolTrimTo(pool, 0);
    });
}

int sw_set_nodes_ready(sw_nodes_ready_fn fn, void* user)
{
    g_nodes_ready = fn;
    g_nodes_ready_user = user;
    return SW_OK;
}

int sw_dist_merge(const void* recv_nodes, const uint64_t* node_counts, const void* recv_kmers,
                  const uint64_t* kmer_counts, const uint64_t* kmer_base, const void* recv_edges,
                  const uint64_t* edge_counts, uint32_t n_src, sw_graph** out, uint32_t* launches)
{
    return guarded([&] {
        init_device_once();
        cudaStream_t s = lib_stream();
        arena_reset();
        auto g = std::make_unique<sw_graph>();
        g->stream = s;
        uint32_t nl = 0;
        dist_merge_nodes(static_cast<const sw_node*>(recv_nodes), node_counts, static_cast<const sw_kmer*>(recv_kmers),
                         kmer_counts, kmer_base, n_src, s, g->dev, &nl);
        if (recv_edges) {
            arena_reset();
            dist_merge_edges(static_cast<const sw_edge*>(recv_edges), edge_counts, n_src, s, g->dev, &nl);
        } else {   // no edges, or sw_dist_merge_edges follows
            g->dev.edges.alloc(0, s);
            g->dev.n_edges = 0;
        }
        if (launches) *launches = nl;
        g->on_device = true;
        g->n_kmers = g->dev.n_kmers;
        g->n_nodes = g->dev.n_nodes;
        g->n_edges = g->dev.n_edges;
        g->record_offsets.assign(1, 0);
        *out = g.release();
    });
}

int sw_dist_merge_edges(sw_graph* g, const void* recv_edges, const uint64_t* edge_counts, uint32_t n_src, uint32_t* launches)
{
    return guarded([&] {
        if (!g->on_device) fail_runtime("graph is not device resident");
        arena_reset();
        uint32_t nl = 0;
        dist_merge_edges(static_cast<const sw_edge*>(recv_edges), edge_counts, n_src, g->stream, g->dev, &nl);
        g->n_edges = g->dev.n_edges;
        if (launches) *launches = nl;
    });
}

int sw_build_from_batch(const sw_batch* b, uint32_t k, uint32_t w, sw_graph** out, sw_stage_times* t)
{
    return guarded([&] { *out = build_pipelined(*b, k, w, t, /*to_host=*/true); });
}

int sw_build_from_batch_ex(const sw_batch* b, uint32_t k, uint32_t w, uint32_t rec_base, int to_host,
                           const uint8_t* is_targets, size_t n_assemblies, sw_graph** out, sw_stage_times* t)
{
    return guarded([&] {
        *out = build_pipelined(*b, k, w, t, to_host != 0, rec_base, is_targets, n_assemblies, /*shard=*/true);
    });
}

int sw_build_from_batch_scored(const sw_batch* b, uint32_t k, uint32_t w, const uint8_t* is_targets, size_t n_assemblies,
                               sw_graph** out, sw_stage_times* t)
{
    return guarded([&] { *out = build_pipelined(*b, k, w, t, /*to_host=*/true, 0u, is_targets, n_assemblies); });
}

int sw_graph_penalty(sw_graph* g, const uint32_t* record_offsets, size_t n_offsets, const uint8_t* is_targets,
                     size_t n_assemblies, float* kernel_ms)
{
    return guarded([&] {
        if (!g->on_device) fail_runtime("graph is not device resident");
        const std::vector<uint32_t> offs = record_offsets ? std::vector<uint32_t>(record_offsets, record_offsets + n_offsets)
                                                          : g->record_offsets;
        const float ms = penalty_on_device(g->dev, offs, is_targets, n_assemblies, g->stream);
        if (kernel_ms) *kernel_ms = ms;
        // any earlier host copy of the nodes is stale now: back to the pool with its buffers
        host_pool_put(g->h_kmers);
        host_pool_put(g->h_nodes);
        host_pool_put(g->h_edges);
        g->on_host = false;
    });
}

int sw_graph_fetch(sw_graph* g)
{
    return guarded([&] {
        if (!g->on_host) {
            if (!g->on_device) fail_runtime("graph holds no data");
            graph_to_host(*g);
        }
    });
}

int sw_build(const char* const* paths, size_t n_paths, uint32_t k, uint32_t w, uint32_t n_host_threads,
             int low_memory, sw_graph** out)
{
    // low_memory: identical output by contract (build.cpp:380-391, test_graph.py:222-245); here it selects the
    // hash-sliced aggregation (graph.cu), whose scratch is a fraction of the single-pass plan's
    return guarded([&] {
        check_kw(k, w);
        init_device_once();
        struct LowMemoryScope {
            bool prev;
            explicit LowMemoryScope(bool on) : prev(sw::low_memory()) { sw::set_low_memory(on || prev); }
            ~LowMemoryScope() { sw::set_low_memory(prev); }
        } scope(low_memory != 0);
        std::unique_ptr<sw_batch> b(batch_from_fasta(paths, n_paths, n_host_threads));
        // the arrays land in pinned host memory while the edge stage still runs; sw_graph_export then
        // copies them into the caller's (numpy) arrays with several threads, because first-touch page
        // faults of a fresh destination, not bandwidth, bound a single-threaded copy
        *out = build_pipelined(*b, k, w, nullptr, /*to_host=*/true);
    });
}

size_t sw_graph_size(const sw_graph* g, int which)
{
    switch (which) {
    case SW_KMERS: return g->n_kmers;
    case SW_NODES: return g->n_nodes;
    case SW_EDGES: return g->n_edges;
    case SW_OFFSETS: return g->record_offsets.size();
    case SW_RECORDS: return g->ids.size();
    default: return 0;
    }
}

namespace {
// memcpy into memory that has probably never been touched (a fresh numpy array): split across threads
void parallel_memcpy(void* dst, const void* src, size_t bytes)
{
    constexpr size_t kChunk = (size_t)4 << 20;
    const size_t n_thr = std::min<size_t>({(size_t)8, std::max<size_t>(1, std::thread::hardware_concurrency()), bytes / kChunk});
    if (n_thr <= 1) {
        memcpy(dst, src, bytes);
        return;
    }
    std::vector<std::thread> th;
    const size_t per = ((bytes / n_thr) + 4095) & ~(size_t)4095;
    for (size_t t = 0; t < n_thr; ++t) {
        const size_t lo = t * per, hi = std::min(bytes, lo + per);
        if (lo >= hi) break;
        th.emplace_back([=] { memcpy(static_cast<char*>(dst) + lo, static_cast<const char*>(src) + lo, hi - lo); });
    }
    for (auto& t : th) t.join();
}
}  // namespace

int sw_graph_export(sw_graph* g, void* kmers, void* nodes, void* edges, uint32_t* record_offsets)
{
    return guarded([&] {
        struct Part { void* dst; const void* host; const void* dev; size_t bytes; };
        const Part parts[3] = {
            {kmers, g->h_kmers.p, g->dev.kmers.p, g->n_kmers * sizeof(sw_kmer)},
            {nodes, g->h_nodes.p, g->dev.nodes.p, g->n_nodes * sizeof(sw_node)},
            {edges, g->h_edges.p, g->dev.edges.p, g->n_edges * sizeof(sw_edge)}};
        for (const Part& pt : parts) {
            if (!pt.dst || !pt.bytes) continue;
            if (g->on_host) parallel_memcpy(pt.dst, pt.host, pt.bytes);
            else  // straight from HBM into the caller's (numpy) buffer, no intermediate copy
                SW_CUDA(cudaMemcpyAsync(pt.dst, pt.dev, pt.bytes, cudaMemcpyDeviceToHost, g->stream));
        }
        if (!g->on_host) SW_CUDA(cudaStreamSynchronize(g->stream));
        if (record_offsets)
            memcpy(record_offsets, g->record_offsets.data(), g->record_offsets.size() * sizeof(uint32_t));
    });
}

size_t sw_graph_n_records(const sw_graph* g, size_t a)
{
    if (a + 1 >= g->record_offsets.size()) return 0;
    return g->record_offsets[a + 1] - g->record_offsets[a];
}
const char* sw_graph_record_id(const sw_graph* g, size_t a, size_t i)
{
    if (a + 1 >= g->record_offsets.size()) return nullptr;
    const size_t r = (size_t)g->record_offsets[a] + i;
    return r < g->ids.size() ? g->ids[r].c_str() : nullptr;
}
}  // extern "C"

sw_graph::~sw_graph()
{
    sw::host_pool_put(h_kmers);
    sw::host_pool_put(h_nodes);
    sw::host_pool_put(h_edges);
}

extern "C" {

void sw_graph_free(sw_graph* g)
{
    if (!g) return;
    cudaStream_t s = g->stream;
    delete g;
    if (s) cudaStreamSynchronize(s);
}

int sw_get_penalty(const sw_kmer* kmers, size_t n_kmers, sw_node* nodes, size_t n_nodes,
                   const uint32_t* record_offsets, size_t n_offsets, const uint8_t* is_targets,
                   size_t n_assemblies, uint32_t n_threads)
{
    (void)n_threads;
    return guarded([&] {
        // argument validation of filter.cpp:33-60 (std::invalid_argument -> ValueError)
        if (n_offsets != n_assemblies + 1) fail_value("len(record_offsets) must equal len(is_targets) + 1");
        if (n_offsets == 0 || record_offsets[0] != 0) fail_value("record_offsets must start with 0");
        if (n_assemblies > 0xFFFFFFFFull) fail_value("Number of assemblies exceeds uint32 range");
        size_t n_t = 0, n_n = 0;
        for (size_t i = 0; i < n_assemblies; ++i) {
            if (record_offsets[i + 1] < record_offsets[i]) fail_value("record_offsets must be nondecreasing");
            if (is_targets[i]) ++n_t; else ++n_n;
        }
        if (!n_t) fail_value("is_targets must contain at least one target assembly");
        if (!n_n) fail_value("is_targets must contain at least one non-target assembly");
        if (n_nodes == 0) return;
        init_device_once();
        cudaStream_t s = lib_stream();
        arena_reset();
        const std::vector<uint32_t> offs(record_offsets, record_offsets + n_offsets);
        const std::vector<uint32_t> ra = record_assembly_map(offs);
        const uint32_t n_records = offs.back();
        DevBuf<sw_kmer> d_k(n_kmers, s, true);
        DevBuf<sw_node> d_n(n_nodes, s, true);
        DevBuf<uint32_t> d_ra(ra.size(), s, true);
        DevBuf<uint8_t> d_t(n_assemblies, s, true);
        if (n_kmers) SW_CUDA(cudaMemcpyAsync(d_k.p, kmers, n_kmers * sizeof(sw_kmer), cudaMemcpyHostToDevice, s));
        SW_CUDA(cudaMemcpyAsync(d_n.p, nodes, n_nodes * sizeof(sw_node), cudaMemcpyHostToDevice, s));
        if (!ra.empty()) SW_CUDA(cudaMemcpyAsync(d_ra.p, ra.data(), ra.size() * 4, cudaMemcpyHostToDevice, s));
        SW_CUDA(cudaMemcpyAsync(d_t.p, is_targets, n_assemblies, cudaMemcpyHostToDevice, s));
        const uint32_t bad = run_penalty(d_k.p, n_kmers, d_n.p, n_nodes, d_ra.p, n_records, d_t.p,
                                         1.0 / (double)n_t, 1.0 / (double)n_n, s);
        if (bad & 4u) fail_value("node [start, stop) range lies outside kmers");
        if (bad & 1u) fail_value("record_idx is outside record_offsets range");
        if (bad & 2u) fail_value("record_idx must be nondecreasing within each node range");
        SW_CUDA(cudaMemcpyAsync(nodes, d_n.p, n_nodes * sizeof(sw_node), cudaMemcpyDeviceToHost, s));
        SW_CUDA(cudaStreamSynchronize(s));
    });
}

int sw_filter_kmers(const sw_kmer* kmers, size_t n_kmers, const sw_node* nodes, size_t n_nodes,
                    const uint64_t* used_hashes, size_t n_used, sw_kmer* kmers_out, sw_node* nodes_out,
                    size_t* n_kmers_out, size_t* n_nodes_out)
{
    // Host-side by design: a sorted-merge of two small lists plus a
    // segment copy; see DESIGN.md "out of scope / next".  filter.cpp:139-201.
    return guarded([&] {
        std::vector<uint64_t> used(used_hashes, used_hashes + n_used);
        std::sort(used.begin(), used.end());
        size_t ni = 0, ui = 0, nk = 0, nn = 0;
        while (ni < n_nodes && ui < used.size()) {
            if (nodes[ni].hash < used[ui]) { ++ni; continue; }
            if (used[ui] < nodes[ni].hash) { ++ui; continue; }
            const sw_node& nd = nodes[ni];
            if (nd.start > nd.stop || nd.stop > n_kmers) fail_value("node [start, stop) range lies outside kmers");
            const size_t size = nd.stop - nd.start;
            if (kmers_out) {
                nodes_out[nn] = nd;
                nodes_out[nn].start = nk;
                nodes_out[nn].stop = nk + size;
                if (size) memcpy(kmers_out + nk, kmers + nd.start, size * sizeof(sw_kmer));
            }
            nk += size;
            ++nn; ++ni; ++ui;
        }
        *n_kmers_out = nk;
        *n_nodes_out = nn;
    });
}

namespace {
// a device-resident graph is about to change: drop any host copy of it
void invalidate_host_copy(sw_graph* g)
{
    if (!g->on_device) fail_runtime("graph is not device resident");
    host_pool_put(g->h_kmers);
    host_pool_put(g->h_nodes);
    host_pool_put(g->h_edges);
    g->on_host = false;
}
void sync_sizes(sw_graph* g)
{
    g->n_kmers = g->dev.n_kmers;
    g->n_nodes = g->dev.n_nodes;
    g->n_edges = g->dev.n_edges;
}
}  // namespace

int sw_graph_filter_edges(sw_graph* g, uint64_t weight_th)
{
    return guarded([&] {
        invalidate_host_copy(g);
        arena_reset();
        graph_filter_edges(g->dev, weight_th, g->stream);
        sync_sizes(g);
    });
}

int sw_graph_filter_kmers(sw_graph* g, const uint64_t* used_hashes, size_t n_used)
{
    return guarded([&] {
        invalidate_host_copy(g);
        arena_reset();
        graph_filter_kmers(g->dev, used_hashes, n_used, g->stream);
        sync_sizes(g);
    });
}

int sw_graph_count_sums(sw_graph* g, uint64_t sums[3])
{
    return guarded([&] {
        if (!g->on_device) fail_runtime("graph is not device resident");
        arena_reset();
        unsigned long long out[3];
        graph_count_sums(g->dev, out, g->stream);
        for (int i = 0; i < 3; ++i) sums[i] = out[i];
    });
}

int sw_filter_edges_and_nodes(const sw_node* nodes, size_t n_nodes, const sw_edge* edges, size_t n_edges,
                              uint64_t weight_th, sw_node* nodes_out, sw_edge* edges_out, size_t* n_nodes_out,
                              size_t* n_edges_out)
{
    return guarded([&] {
        init_device_once();
        cudaStream_t s = lib_stream();
        arena_reset();
        DevGraph g;
        g.nodes.alloc(n_nodes, s);
        g.edges.alloc(n_edges, s);
        g.kmers.alloc(0, s);
        g.n_nodes = n_nodes;
        g.n_edges = n_edges;
        if (n_nodes) SW_CUDA(cudaMemcpyAsync(g.nodes.p, nodes, n_nodes * sizeof(sw_node), cudaMemcpyHostToDevice, s));
        if (n_edges) SW_CUDA(cudaMemcpyAsync(g.edges.p, edges, n_edges * sizeof(sw_edge), cudaMemcpyHostToDevice, s));
        graph_filter_edges(g, weight_th, s);
        if (g.n_nodes) SW_CUDA(cudaMemcpyAsync(nodes_out, g.nodes.p, g.n_nodes * sizeof(sw_node), cudaMemcpyDeviceToHost, s));
        if (g.n_edges) SW_CUDA(cudaMemcpyAsync(edges_out, g.edges.p, g.n_edges * sizeof(sw_edge), cudaMemcpyDeviceToHost, s));
        SW_CUDA(cudaStreamSynchronize(s));
        *n_nodes_out = g.n_nodes;
        *n_edges_out = g.n_edges;
    });
}

int sw_dev_sketch(const sw_dev_batch* d, uint32_t k, uint32_t w, uint64_t* h1_out, uint32_t* pos_out,
                  uint32_t* record_out, size_t capacity, size_t* n_out)
{
    return guarded([&] {
        init_device_once();
        check_kw(k, w);
        cudaStream_t s = d->stream;
        arena_reset();
        DevPlan plan = make_plan(*d, k, w, s);
        SketchStream st;
        run_sketch(d->words.p, d->rec_word_off.p, plan, k, w, 0u, s, st);
        *n_out = st.n;
        if (!h1_out) return;
        const size_t n = std::min<size_t>(st.n, capacity);
        std::vector<uint64_t> vals(n);
        if (n) {
            SW_CUDA(cudaMemcpyAsync(h1_out, st.keys.p, n * 8, cudaMemcpyDeviceToHost, s));
            SW_CUDA(cudaMemcpyAsync(vals.data(), st.vals.p, n * 8, cudaMemcpyDeviceToHost, s));
        }
        SW_CUDA(cudaStreamSynchronize(s));
        for (size_t i = 0; i < n; ++i) {
            pos_out[i] = (uint32_t)vals[i];
            record_out[i] = (uint32_t)(vals[i] >> 32);
        }
    });
}

}  // extern "C"
