// capi.cu -- context management, memory helpers and the sketching entry points of
// the C-ABI declared in include/mcgpu.h
#include "common.cuh"

#include <cstdarg>

namespace mcg {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int DevBuf::reserve(size_t bytes)
{
    if(bytes <= cap && p != nullptr) return MCG_OK;
    if(p != nullptr) { cudaFree(p); p = nullptr; cap = 0; }
    size_t want = bytes + bytes / 8 + 256;       // slack against repeated growth
    want = (want + 255) & ~size_t(255);
    cudaError_t e = cudaMalloc(&p, want);
    if(e != cudaSuccess) {
        p = nullptr;
        set_error("cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
        cudaGetLastError();
        return MCG_ERR_NOMEM;
    }
    cap = want;
    return MCG_OK;
}

int DevBuf::reserve_keep(size_t bytes, size_t used, cudaStream_t st)
{
    if(bytes <= cap && p != nullptr) return MCG_OK;
    size_t want = bytes + bytes / 2 + 256;
    want = (want + 255) & ~size_t(255);
    void* np = nullptr;
    cudaError_t e = cudaMalloc(&np, want);
    if(e != cudaSuccess) {
        // retry without slack before giving up
        want = (bytes + 255) & ~size_t(255);
        cudaGetLastError();
        e = cudaMalloc(&np, want);
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
    if(p != nullptr) cudaFree(p);
    p = np;
    cap = want;
    return MCG_OK;
}

void DevBuf::release()
{
    if(p != nullptr) cudaFree(p);
    p = nullptr;
    cap = 0;
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
    DevBuf* bufs[] = {&c->lineage, &c->rank, &c->seq, &c->seqoff, &c->winfirst, &c->outfirst, &c->tmp0, &c->tmp1,
                      &c->tmp2, &c->tmp3, &c->tmp4, &c->tmp5, &c->scan_tmp, &c->q_feat, &c->q_flen, &c->q_foff,
                      &c->q_fpos, &c->q_loc, &c->q_rinfo, &c->q_cls, &c->q_out_tax, &c->q_out_hits, &c->q_out_n,
                      &c->m_tax, &c->m_hits, &c->m_n, &c->m_o_tax, &c->m_o_hits, &c->m_o_n, &c->m_best,
                      &c->m_brank, &c->m_stats};
    for(DevBuf* b : bufs) b->release();
    if(c->pinned != nullptr) cudaFreeHost(c->pinned);
    if(c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
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

int mcg_sketch_windows(mcg_ctx* c, const mcg_sketch_cfg* cfg,
                       const char* seq_bytes, const int64_t* seq_offsets, int64_t nseq,
                       uint32_t* out_features, uint32_t* out_counts, int64_t* out_nwin)
{
    MCG_REQUIRE(c != nullptr && cfg != nullptr && seq_offsets != nullptr, "mcg_sketch_windows: NULL argument");
    MCG_REQUIRE(nseq >= 0, "nseq < 0");
    const int64_t nwin = mcg_count_windows(cfg, seq_offsets, nseq);
    MCG_REQUIRE(nwin >= 0, "bad window parameters");
    if(out_nwin != nullptr) *out_nwin = nwin;
    if(nseq == 0 || nwin == 0) return MCG_OK;
    MCG_REQUIRE(out_features != nullptr, "out_features is NULL");
    MCG_CUDA(cudaSetDevice(c->device));
    const int64_t base = seq_offsets[0];
    const int64_t total = seq_offsets[nseq] - base;
    MCG_REQUIRE(base == 0, "seq_offsets[0] must be 0");
    const int s = cfg->sketchlen;
    MCG_TRY(c->seq.reserve(size_t(total) + 64));
    MCG_TRY(c->seqoff.reserve(size_t(nseq + 1) * sizeof(int64_t)));
    MCG_TRY(c->q_feat.reserve(size_t(nwin) * s * sizeof(uint32_t)));
    if(out_counts != nullptr) MCG_TRY(c->tmp0.reserve(size_t(nwin) * sizeof(uint32_t)));
    MCG_TRY(copy_in(c, c->seq.p, seq_bytes, size_t(total)));
    MCG_TRY(copy_in(c, c->seqoff.p, seq_offsets, size_t(nseq + 1) * sizeof(int64_t)));
    MCG_TRY(mcg_sketch_windows_dev(c, cfg, c->seq.as<char>(), total, c->seqoff.as<int64_t>(), nseq, nwin,
                                   c->q_feat.as<uint32_t>(), out_counts ? c->tmp0.as<uint32_t>() : nullptr));
    MCG_TRY(copy_out(c, out_features, c->q_feat.p, size_t(nwin) * s * sizeof(uint32_t)));
    if(out_counts != nullptr) MCG_TRY(copy_out(c, out_counts, c->tmp0.p, size_t(nwin) * sizeof(uint32_t)));
    MCG_CUDA(cudaStreamSynchronize(c->stream));
    return MCG_OK;
}

}  // extern "C"
