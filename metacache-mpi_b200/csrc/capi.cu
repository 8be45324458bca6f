// capi.cu -- context management, memory helpers and the sketching entry points of
// the C-ABI declared in include/mcgpu.h
#include "common.cuh"

#include <cstdarg>
#include <algorithm>
#include <mutex>

namespace mcg {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---- caching device allocator -------------------------------------------------------------
// cudaFree of multi-GB blocks costs up to hundreds of milliseconds (and synchronises the device);
// released blocks are therefore parked in a per-device cache and handed out again (best fit).
// The cache is flushed when an allocation fails.
namespace {
struct Block { void* p; size_t cap; int dev; };
std::vector<Block> g_cache;
std::mutex g_cache_mtx;

void* cache_take(size_t bytes, size_t* cap)
{
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_cache_mtx);
    int best = -1;
    for(size_t i = 0; i < g_cache.size(); ++i) {
        const Block& b = g_cache[i];
        if(b.dev != dev || b.cap < bytes || b.cap > 2 * bytes + (size_t(1) << 20)) continue;
        if(best < 0 || b.cap < g_cache[size_t(best)].cap) best = int(i);
    }
    if(best < 0) return nullptr;
    void* p = g_cache[size_t(best)].p;
    *cap = g_cache[size_t(best)].cap;
    g_cache.erase(g_cache.begin() + best);
    return p;
}

void cache_flush()
{
    std::lock_guard<std::mutex> lk(g_cache_mtx);
    int cur = 0;
    cudaGetDevice(&cur);
    for(const Block& b : g_cache) { cudaSetDevice(b.dev); cudaFree(b.p); }
    cudaSetDevice(cur);
    g_cache.clear();
}

cudaError_t cached_malloc(void** p, size_t bytes, size_t* cap)
{
    void* q = cache_take(bytes, cap);
    if(q != nullptr) { *p = q; return cudaSuccess; }
    cudaError_t e = cudaMalloc(p, bytes);
    if(e != cudaSuccess) {
        cudaGetLastError();
        cache_flush();
        e = cudaMalloc(p, bytes);
    }
    *cap = bytes;
    return e;
}
}  // namespace

void release_cached_device_memory() { cache_flush(); }

// frees parked blocks of the current device, largest first, until at most keep_bytes stay parked:
// multi-GB scratch (the sort buffers of finalize) must not sit in the cache while other allocators of
// the process (torch, NCCL) run out of memory
void trim_cached_device_memory(size_t keep_bytes)
{
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_cache_mtx);
    size_t total = 0;
    for(const Block& b : g_cache) if(b.dev == dev) total += b.cap;
    while(total > keep_bytes) {
        int big = -1;
        for(size_t i = 0; i < g_cache.size(); ++i)
            if(g_cache[i].dev == dev && (big < 0 || g_cache[i].cap > g_cache[size_t(big)].cap)) big = int(i);
        if(big < 0) break;
        total -= g_cache[size_t(big)].cap;
        cudaFree(g_cache[size_t(big)].p);
        g_cache.erase(g_cache.begin() + big);
    }
}

int DevBuf::reserve(size_t bytes)
{
    if(bytes <= cap && p != nullptr) return MCG_OK;
    release();
    size_t want = bytes + bytes / 8 + 256;       // slack against repeated growth
    want = (want + 255) & ~size_t(255);
    size_t got = 0;
    cudaError_t e = cached_malloc(&p, want, &got);
    if(e != cudaSuccess) {
        p = nullptr;
        set_error("cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
        cudaGetLastError();
        return MCG_ERR_NOMEM;
    }
    cap = got;
    return MCG_OK;
}

int DevBuf::reserve_keep(size_t bytes, size_t used, cudaStream_t st)
{
    if(bytes <= cap && p != nullptr) return MCG_OK;
    size_t want = bytes + bytes / 2 + 256;
    want = (want + 255) & ~size_t(255);
    void* np = nullptr;
    size_t got = 0;
    cudaError_t e = cached_malloc(&np, want, &got);
    if(e != cudaSuccess) {
        // retry without slack before giving up
        want = (bytes + 255) & ~size_t(255);
        cudaGetLastError();
        e = cached_malloc(&np, want, &got);
    }
    if(e != cudaSuccess) {
        set_error("cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
        cudaGetLastError();
        return MCG_ERR_NOMEM;
    }
    if(p != nullptr && used > 0) {
        MCG_CUDA(cudaMemcpyAsync(np, p, used, cudaMemcpyDeviceToDevice, st));
        MCG_CUDA(cudaStreamSynchronize(st));
    }
    release();
    p = np;
    cap = got;
    return MCG_OK;
}

void DevBuf::release()
{
    if(p == nullptr) return;
    // in-flight work may still use the block: wait (cudaFree did so implicitly), then park it
    cudaDeviceSynchronize();
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lk(g_cache_mtx);
        g_cache.push_back(Block{p, cap, dev});
    }
    p = nullptr;
    cap = 0;
    // parked memory is capped (MCG_CACHE_CAP_MB, default a quarter of the device memory): beyond that the largest
    // blocks go back to the driver, so that other allocators of the process (torch, NCCL) are not starved;
    // mcg_ctx_trim frees all
    static const size_t cap_bytes = [] {
        if(getenv("MCG_CACHE_CAP_MB")) return size_t(atoll(getenv("MCG_CACHE_CAP_MB"))) << 20;
        size_t fr = 0, tot = 0;
        if(cudaMemGetInfo(&fr, &tot) != cudaSuccess) { cudaGetLastError(); return size_t(8) << 30; }
        return tot / 4;
    }();
    trim_cached_device_memory(cap_bytes);
}

// host pointer -> device, asynchronous when the host memory is page-locked
int copy_in(mcg_ctx* c, void* d, const void* h, size_t bytes)
{
    if(bytes == 0) return MCG_OK;
    MCG_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, c->stream));
    return MCG_OK;
}

int copy_out(mcg_ctx* c, void* h, const void* d, size_t bytes)
{
    if(bytes == 0) return MCG_OK;
    MCG_CUDA(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, c->stream));
    return MCG_OK;
}

}  // namespace mcg

using namespace mcg;

extern "C" {

const char* mcg_last_error(void) { return g_err; }

int mcg_version(void) { return MCG_VERSION; }

int mcg_device_count(void)
{
    int n = 0;
    if(cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int mcg_ctx_create(int device, void* stream, mcg_ctx** out)
{
    MCG_REQUIRE(out != nullptr, "mcg_ctx_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if(e != cudaSuccess || n == 0) {
        cudaGetLastError();
        set_error("no CUDA device available (%s); libmcgpu has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        return MCG_ERR_CUDA;
    }
    MCG_REQUIRE(device >= 0 && device < n, "device %d out of range (have %d)", device, n);
    MCG_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    MCG_CUDA(cudaGetDeviceProperties(&prop, device));
    if(prop.major != 10) {
        set_error("device %d is sm_%d%d; libmcgpu is built for sm_100a only", device, prop.major, prop.minor);
        return MCG_ERR_CUDA;
    }
    mcg_ctx* c = new mcg_ctx;
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    if(stream != nullptr) {
        c->stream = static_cast<cudaStream_t>(stream);
        c->own_stream = false;
    } else {
        MCG_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    *out = c;
    return MCG_OK;
}

int mcg_ctx_destroy(mcg_ctx* c)
{
    if(c == nullptr) return MCG_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    DevBuf* bufs[] = {&c->lineage, &c->rank, &c->lift, &c->seq, &c->seqoff, &c->winfirst, &c->outfirst, &c->tmp0, &c->tmp1,
                      &c->tmp2, &c->tmp3, &c->tmp4, &c->tmp5, &c->scan_tmp, &c->q_feat, &c->q_flen, &c->q_foff,
                      &c->q_fpos, &c->q_loc, &c->q_rinfo, &c->q_cls, &c->q_out_tax, &c->q_out_hits, &c->q_out_n, &c->q_done,
                      &c->m_tax, &c->m_hits, &c->m_n, &c->m_o_tax, &c->m_o_hits, &c->m_o_n, &c->m_best,
                      &c->m_brank, &c->m_stats, &c->m_bad, &c->tilectr};
    for(DevBuf* b : bufs) b->release();
    c->qp[0].release_all(); c->qp[1].release_all();
    c->s_wc.release(); c->s_featc.release(); c->plan.release(); c->q_spill.release();
    if(c->h_left != nullptr) cudaFreeHost(c->h_left);
    if(c->pinned != nullptr) cudaFreeHost(c->pinned);
    if(c->s_h2d) cudaStreamDestroy(c->s_h2d);
    if(c->s_d2h) cudaStreamDestroy(c->s_d2h);
    for(int i = 0; i < 4; ++i) { if(c->ev_h2d[i]) cudaEventDestroy(c->ev_h2d[i]); if(c->ev_k[i]) cudaEventDestroy(c->ev_k[i]); }
    if(c->ev_misc) cudaEventDestroy(c->ev_misc);
    if(c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    release_cached_device_memory();
    return MCG_OK;
}

int mcg_ctx_sync(mcg_ctx* c)
{
    MCG_REQUIRE(c != nullptr, "ctx is NULL");
    MCG_CUDA(cudaStreamSynchronize(c->stream));
    return MCG_OK;
}

void* mcg_ctx_stream(mcg_ctx* c) { return c ? static_cast<void*>(c->stream) : nullptr; }

int64_t mcg_ctx_launches(mcg_ctx* c) { return c ? c->launches : 0; }

int mcg_ctx_trim(mcg_ctx* c)
{
    MCG_REQUIRE(c != nullptr, "ctx is NULL");
    MCG_CUDA(cudaSetDevice(c->device));
    MCG_CUDA(cudaStreamSynchronize(c->stream));
    trim_cached_device_memory(0);
    return MCG_OK;
}

int mcg_ctx_profile(mcg_ctx* c, int enable)
{
    MCG_REQUIRE(c != nullptr, "ctx is NULL");
    c->profile = enable != 0;
    return MCG_OK;
}

int mcg_ctx_kernel_times(mcg_ctx* c, int which, double* ms, int64_t* launches)
{
    MCG_REQUIRE(c != nullptr && which >= 0 && which < 8, "mcg_ctx_kernel_times: bad argument");
    MCG_CUDA(cudaSetDevice(c->device));
    MCG_CUDA(cudaStreamSynchronize(c->stream));
    double tot = 0.0;
    std::vector<cudaEvent_t>& ev = c->prof_ev[which];
    for(size_t i = 0; i + 1 < ev.size(); i += 2) {
        float t = 0.f;
        if(cudaEventSynchronize(ev[i + 1]) == cudaSuccess && cudaEventElapsedTime(&t, ev[i], ev[i + 1]) == cudaSuccess) tot += double(t);
        cudaEventDestroy(ev[i]); cudaEventDestroy(ev[i + 1]);
    }
    cudaGetLastError();
    if(ms) *ms = tot;
    if(launches) *launches = int64_t(ev.size() / 2);
    ev.clear();
    return MCG_OK;
}

int mcg_dev_alloc(mcg_ctx* c, int64_t bytes, void** dptr)
{
    MCG_REQUIRE(c != nullptr && dptr != nullptr && bytes >= 0, "mcg_dev_alloc: bad argument");
    MCG_CUDA(cudaSetDevice(c->device));
    *dptr = nullptr;
    const size_t want = (size_t(bytes) + 271) & ~size_t(255);   // >= 16 bytes of slack for vector loads
    cudaError_t e = cudaMalloc(dptr, want);
    if(e != cudaSuccess) {
        cudaGetLastError();
        set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
        return MCG_ERR_NOMEM;
    }
    return MCG_OK;
}

int mcg_dev_free(mcg_ctx* c, void* dptr)
{
    MCG_REQUIRE(c != nullptr, "ctx is NULL");
    if(dptr != nullptr) MCG_CUDA(cudaFree(dptr));
    return MCG_OK;
}

int mcg_host_alloc(int64_t bytes, void** hptr)
{
    MCG_REQUIRE(hptr != nullptr && bytes >= 0, "mcg_host_alloc: bad argument");
    *hptr = nullptr;
    cudaError_t e = cudaHostAlloc(hptr, size_t(bytes) + 64, cudaHostAllocPortable);
    if(e != cudaSuccess) {
        cudaGetLastError();
        set_error("cudaHostAlloc(%lld) failed: %s", (long long)bytes, cudaGetErrorString(e));
        return MCG_ERR_NOMEM;
    }
    return MCG_OK;
}

int mcg_host_free(void* hptr)
{
    if(hptr != nullptr) MCG_CUDA(cudaFreeHost(hptr));
    return MCG_OK;
}

int mcg_memcpy_h2d(mcg_ctx* c, void* dptr, const void* hptr, int64_t bytes)
{
    MCG_REQUIRE(c != nullptr && bytes >= 0, "mcg_memcpy_h2d: bad argument");
    MCG_TRY(copy_in(c, dptr, hptr, size_t(bytes)));
    MCG_CUDA(cudaStreamSynchronize(c->stream));
    return MCG_OK;
}

int mcg_memcpy_d2h(mcg_ctx* c, void* hptr, const void* dptr, int64_t bytes)
{
    MCG_REQUIRE(c != nullptr && bytes >= 0, "mcg_memcpy_d2h: bad argument");
    MCG_TRY(copy_out(c, hptr, dptr, size_t(bytes)));
    MCG_CUDA(cudaStreamSynchronize(c->stream));
    return MCG_OK;
}

// ------------------------------------------------------------------ sketching
int64_t mcg_count_windows(const mcg_sketch_cfg* cfg, const int64_t* seq_offsets, int64_t nseq)
{
    if(cfg == nullptr || seq_offsets == nullptr || nseq < 0 || cfg->winlen < 1 || cfg->winstride < 1) return -1;
    int64_t n = 0;
    for(int64_t i = 0; i < nseq; ++i) n += num_windows(seq_offsets[i + 1] - seq_offsets[i], cfg->winlen, cfg->winstride);
    return n;
}

int mcg_sketch_windows_dev(mcg_ctx* c, const mcg_sketch_cfg* cfg,
                           const char* d_seq, int64_t total_bytes,
                           const int64_t* d_seq_offsets, int64_t nseq, int64_t nwin,
                           uint32_t* d_out_features, uint32_t* d_out_counts)
{
    MCG_REQUIRE(c != nullptr && cfg != nullptr, "mcg_sketch_windows_dev: NULL argument");
    MCG_REQUIRE(nseq >= 0 && nwin >= 0 && nwin < (1ll << 32), "window count out of range");
    if(nseq == 0 || nwin == 0) return MCG_OK;
    MCG_CUDA(cudaSetDevice(c->device));
    MCG_TRY(c->winfirst.reserve(size_t(nseq + 2) * sizeof(uint32_t)));
    uint64_t total = 0;
    MCG_TRY(window_prefix(c, cfg->winlen, cfg->winstride, d_seq_offsets, nseq, c->winfirst.as<uint32_t>(), nullptr));
    (void)total;
    return sketch_launch(c, *cfg, cfg->sketchlen, cfg->winlen, cfg->winstride, d_seq, total_bytes, d_seq_offsets, nseq,
                         c->winfirst.as<uint32_t>(), c->winfirst.as<uint32_t>(), uint64_t(nwin),
                         d_out_features, d_out_counts);
}

}  // extern "C"
namespace mcg {
int ensure_pipeline(mcg_ctx* c)
{
    if(c->s_h2d != nullptr) return MCG_OK;
    MCG_CUDA(cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
    MCG_CUDA(cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
    for(int i = 0; i < 4; ++i) {
        MCG_CUDA(cudaEventCreateWithFlags(&c->ev_h2d[i], cudaEventDisableTiming));
        MCG_CUDA(cudaEventCreateWithFlags(&c->ev_k[i], cudaEventDisableTiming));
    }
    MCG_CUDA(cudaEventCreateWithFlags(&c->ev_misc, cudaEventDisableTiming));
    for(int i = 0; i < 2; ++i) {
        MCG_CUDA(cudaEventCreateWithFlags(&c->qp[i].evA, cudaEventDisableTiming));
        MCG_CUDA(cudaEventCreateWithFlags(&c->qp[i].evB, cudaEventDisableTiming));
    }
    MCG_CUDA(cudaMallocHost(reinterpret_cast<void**>(&c->h_left), 64));
    return MCG_OK;
}
}  // namespace mcg
extern "C" {

// Host path: the input is cut into chunks of windows; chunk i+1 is copied to the device
// (H2D stream) while chunk i is sketched (ctx stream) and the sketches of chunk i-1 travel
// back (D2H stream).  PCIe is full duplex, so the call costs about max(H2D, D2H) instead
// of H2D + K1 + D2H.
int mcg_sketch_windows(mcg_ctx* c, const mcg_sketch_cfg* cfg,
                       const char* seq_bytes, const int64_t* seq_offsets, int64_t nseq,
                       uint32_t* out_features, uint32_t* out_counts, int64_t* out_nwin)
{
    MCG_REQUIRE(c != nullptr && cfg != nullptr && seq_offsets != nullptr, "mcg_sketch_windows: NULL argument");
    MCG_REQUIRE(nseq >= 0, "nseq < 0");
    const int64_t nwin = mcg_count_windows(cfg, seq_offsets, nseq);
    MCG_REQUIRE(nwin >= 0, "bad window parameters");
    MCG_REQUIRE(nwin < (1ll << 32), "more than 2^32-1 windows in one call");
    if(out_nwin != nullptr) *out_nwin = nwin;
    if(nseq == 0 || nwin == 0) return MCG_OK;
    MCG_REQUIRE(out_features != nullptr && seq_bytes != nullptr, "NULL buffer");
    MCG_REQUIRE(seq_offsets[0] == 0, "seq_offsets[0] must be 0");
    MCG_CUDA(cudaSetDevice(c->device));
    MCG_TRY(ensure_pipeline(c));
    const int64_t total = seq_offsets[nseq];
    const int s = cfg->sketchlen;
    int wpt = 0;
    MCG_TRY(sketch_config(*cfg, s, cfg->winlen, cfg->winstride, &wpt, nullptr, nullptr, nullptr));
    // chunking by windows: ~32 MB of sequence per chunk
    const int64_t adv = cfg->winstride;
    int64_t wchunk = (int64_t(32) << 20) / (adv > 0 ? adv : 1);
    wchunk = (wchunk / wpt + 1) * wpt;
    const int64_t nchunks = (nwin + wchunk - 1) / wchunk;
    MCG_TRY(c->seq.reserve(size_t(total) + 128));
    MCG_TRY(c->seqoff.reserve(size_t(nseq + 1) * sizeof(int64_t)));
    MCG_TRY(c->winfirst.reserve(size_t(nseq + 2) * sizeof(uint32_t)));
    MCG_TRY(c->q_feat.reserve(size_t(nwin) * s * sizeof(uint32_t)));
    if(out_counts != nullptr) MCG_TRY(c->tmp0.reserve(size_t(nwin) * sizeof(uint32_t)));
    MCG_TRY(c->outfirst.reserve(size_t(nwin / wpt + 8 * nchunks + 64) * sizeof(uint32_t)));
    // window prefix per sequence on the host (chunk boundaries) and on the device (kernel)
    std::vector<uint64_t> wf(size_t(nseq) + 1, 0);
    for(int64_t i = 0; i < nseq; ++i)
        wf[i + 1] = wf[i] + uint64_t(num_windows(seq_offsets[i + 1] - seq_offsets[i], cfg->winlen, cfg->winstride));
    MCG_TRY(copy_in(c, c->seqoff.p, seq_offsets, size_t(nseq + 1) * sizeof(int64_t)));
    MCG_TRY(window_prefix(c, cfg->winlen, cfg->winstride, c->seqoff.as<int64_t>(), nseq, c->winfirst.as<uint32_t>(), nullptr));
    MCG_CUDA(cudaEventRecord(c->ev_misc, c->stream));
    MCG_CUDA(cudaStreamWaitEvent(c->s_h2d, c->ev_misc, 0));     // scratch buffers are ours from here on
    MCG_CUDA(cudaStreamWaitEvent(c->s_d2h, c->ev_misc, 0));

    // end byte (exclusive) of window g
    auto window_end = [&](uint64_t g) -> int64_t {
        size_t lo = 0, hi = size_t(nseq);
        while(hi - lo > 1) { const size_t mid = (lo + hi) / 2; if(wf[mid] <= g) lo = mid; else hi = mid; }
        const int64_t len = seq_offsets[lo + 1] - seq_offsets[lo];
        if(len <= cfg->winlen) return seq_offsets[lo + 1];
        const int64_t b = int64_t(g - wf[lo]) * cfg->winstride;
        return seq_offsets[lo] + (b + cfg->winlen < len ? b + cfg->winlen : len);
    };

    int64_t copied = 0;
    uint64_t tile_off = 0;
    for(int64_t ch = 0; ch < nchunks; ++ch) {
        const uint64_t ga = uint64_t(ch) * uint64_t(wchunk);
        const uint64_t gb = std::min<uint64_t>(uint64_t(nwin), ga + uint64_t(wchunk));
        const int slot = int(ch & 3);
        // (1) bytes up to the end of the chunk's last window
        const int64_t hi = (ch + 1 == nchunks) ? total : window_end(gb - 1);
        if(hi > copied) {
            MCG_CUDA(cudaMemcpyAsync(c->seq.as<char>() + copied, seq_bytes + copied, size_t(hi - copied),
                                     cudaMemcpyHostToDevice, c->s_h2d));
            copied = hi;
        }
        MCG_CUDA(cudaEventRecord(c->ev_h2d[slot], c->s_h2d));
        // (2) sketch the chunk
        MCG_CUDA(cudaStreamWaitEvent(c->stream, c->ev_h2d[slot], 0));
        MCG_TRY(sketch_launch(c, *cfg, s, cfg->winlen, cfg->winstride, c->seq.as<char>(), total, c->seqoff.as<int64_t>(), nseq,
                              c->winfirst.as<uint32_t>(), c->winfirst.as<uint32_t>(), gb, c->q_feat.as<uint32_t>(),
                              out_counts ? c->tmp0.as<uint32_t>() : nullptr, ga, c->stream,
                              c->outfirst.as<uint32_t>() + tile_off));
        tile_off += (gb - ga) / uint64_t(wpt) + 4;
        MCG_CUDA(cudaEventRecord(c->ev_k[slot], c->stream));
        // (3) results of the chunk back to the host
        MCG_CUDA(cudaStreamWaitEvent(c->s_d2h, c->ev_k[slot], 0));
        MCG_CUDA(cudaMemcpyAsync(out_features + ga * uint64_t(s), c->q_feat.as<uint32_t>() + ga * uint64_t(s),
                                 size_t(gb - ga) * s * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->s_d2h));
        if(out_counts != nullptr)
            MCG_CUDA(cudaMemcpyAsync(out_counts + ga, c->tmp0.as<uint32_t>() + ga, size_t(gb - ga) * sizeof(uint32_t),
                                     cudaMemcpyDeviceToHost, c->s_d2h));
    }
    MCG_CUDA(cudaStreamSynchronize(c->s_d2h));
    MCG_CUDA(cudaStreamSynchronize(c->stream));
    return MCG_OK;
}

}  // extern "C"
