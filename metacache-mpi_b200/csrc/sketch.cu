// sketch.cu -- K1: 2-bit encode + canonical k-mer + thomas_mueller_hash +
// per-window bottom-s unique selection, fused.  Must equal
//   dna_encoding.h:261-276 (windows), :305-348 (2-bit k-mers + ambiguity),
//   :114-121,189-197 (reverse complement, canonical), hash_int.h:40-45 (hash),
//   hash_dna.h:115-152 (bottom-s unique, ascending, 0xFFFFFFFF never a feature).
//
// Work decomposition: a CTA owns SK_WPC consecutive windows.  Their bytes form
// one contiguous span of the sequence buffer (sequences are stored back to
// back), which is
//   (1) loaded once with aligned 16-byte loads and packed to 2 bits/base (+1
//       ambiguity bit/base) in shared memory,
//   (2) hashed position by position (every k-mer start of the span exactly
//       once; with stride = winlen-k+1 the k-mers of consecutive windows tile
//       the span) into a transposed shared-memory array, and
//   (3) reduced per window by one warp: threshold pre-filter, ballot
//       compaction, 32-wide bitonic network over shuffles, duplicate removal,
//       with an exact unfiltered multi-chunk fallback.
#include "common.cuh"

namespace mcg {

constexpr int SK_THREADS = 256;
constexpr int SK_WARPS = SK_THREADS / 32;
constexpr int SK_CBUF = 128;              // per-warp compaction buffer (hashes)
constexpr uint32_t NOF = 0xFFFFFFFFu;

struct SketchParams {
    const uint8_t* seq;
    int64_t total_bytes;
    const int64_t* seq_off;
    int64_t nseq;
    const uint32_t* win_first;   // [nseq+1] window prefix in processing order
    const uint32_t* out_first;   // [nseq]   output index of each sequence's first window
    uint64_t nwin;
    uint32_t* out_feat;
    uint32_t* out_cnt;
    int k, s;
    int64_t winlen, stride;
    int wpc;                     // windows per CTA
    int ngroups;                 // 16-base groups of span capacity
    int pitch;                   // row pitch of the transposed hash array (words)
};

// ---- 4 ASCII bases -> 4 x 2-bit codes (first base in the high bits) + validity diff
// code: A/a=0 C/c=1 G/g=2 T/t=3 (dna_encoding.h:330-336)
__device__ __forceinline__ void pack4(uint32_t x, uint32_t& code8, uint32_t& diff)
{
    uint32_t c = (x >> 1) & 0x03030303u;
    c ^= (c >> 1) & 0x01010101u;
    code8 = (c * 0x40100401u) >> 24;
    // the letter each code stands for, compared with the case-folded input
    const uint32_t ce = c & 0x00FF00FFu, co = c & 0xFF00FF00u;
    const uint32_t sel = (ce | (ce >> 8) | (co >> 4) | (co >> 12)) & 0xFFFFu;
    const uint32_t e = __byte_perm(0x54474341u, 0u, sel);
    diff = (x & 0xDFDFDFDFu) ^ e;
}

// non-zero bytes of d -> 4 bits, first byte in bit 3
__device__ __forceinline__ uint32_t nzmask4(uint32_t d)
{
    const uint32_t nz = ((((d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | d) & 0x80808080u) >> 7;
    return ((nz * 0x08040201u) >> 24) & 0xFu;
}

// reverse complement of all 16 letters of a packed word
__device__ __forceinline__ uint32_t rc16(uint32_t w)
{
    uint32_t r = __brev(w);
    r = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
    return ~r;
}

__device__ __forceinline__ uint32_t sort32(uint32_t v, int lane)
{
#pragma unroll
    for(int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for(int j = k >> 1; j > 0; j >>= 1) {
            const uint32_t o = __shfl_xor_sync(0xFFFFFFFFu, v, j);
            const bool up = (lane & k) == 0;
            const bool lower = (lane & j) == 0;
            v = (up == lower) ? min(v, o) : max(v, o);
        }
    }
    return v;
}

// sorts a bitonic sequence ascending
__device__ __forceinline__ uint32_t bmerge32(uint32_t v, int lane)
{
#pragma unroll
    for(int j = 16; j > 0; j >>= 1) {
        const uint32_t o = __shfl_xor_sync(0xFFFFFFFFu, v, j);
        v = ((lane & j) == 0) ? min(v, o) : max(v, o);
    }
    return v;
}

// v ascending with NOF padding; removes duplicates, keeps ascending + padding
__device__ __forceinline__ uint32_t dedup_sorted(uint32_t v, int lane, uint32_t* tmp)
{
    const uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, v, 1);
    const bool valid = v != NOF;
    const bool keep = valid && (lane == 0 || v != prev);
    const uint32_t bk = __ballot_sync(0xFFFFFFFFu, keep);
    const uint32_t bv = __ballot_sync(0xFFFFFFFFu, valid);
    if(bk == bv) return v;
    tmp[lane] = NOF;
    __syncwarp();
    if(keep) tmp[__popc(bk & ((1u << lane) - 1u))] = v;
    __syncwarp();
    v = tmp[lane];
    __syncwarp();
    return v;
}

// kept: <= 32 smallest distinct values so far (ascending, NOF padded)
// v   : 32 new values in any order (NOF = none)
// wide: keep all 32 smallest distinct (needed when sketchlen > 16); otherwise
//       at least the 16 smallest distinct survive, which is cheaper
__device__ __forceinline__ uint32_t merge_chunk(uint32_t kept, uint32_t v, bool first, bool wide,
                                                int lane, uint32_t* tmp)
{
    v = sort32(v, lane);
    v = dedup_sorted(v, lane, tmp);
    if(first) return v;
    const uint32_t t = __shfl_sync(0xFFFFFFFFu, v, 31 - lane);
    uint32_t lo = bmerge32(min(kept, t), lane);
    if(!wide) return dedup_sorted(lo, lane, tmp);
    uint32_t hi = bmerge32(max(kept, t), lane);
    // lo ++ hi is the sorted union (64 values, each value at most twice, twins adjacent)
    const uint32_t plo = __shfl_up_sync(0xFFFFFFFFu, lo, 1);
    const uint32_t lo_last = __shfl_sync(0xFFFFFFFFu, lo, 31);
    const uint32_t phi0 = __shfl_up_sync(0xFFFFFFFFu, hi, 1);
    const uint32_t phi = lane == 0 ? lo_last : phi0;
    const bool klo = lo != NOF && (lane == 0 || lo != plo);
    const bool khi = hi != NOF && hi != phi;
    const uint32_t blo = __ballot_sync(0xFFFFFFFFu, klo);
    const uint32_t bhi = __ballot_sync(0xFFFFFFFFu, khi);
    const int nlo = __popc(blo);
    tmp[lane] = NOF;
    __syncwarp();
    if(klo) tmp[__popc(blo & ((1u << lane) - 1u))] = lo;
    const int ph = nlo + __popc(bhi & ((1u << lane) - 1u));
    if(khi && ph < 32) tmp[ph] = hi;
    __syncwarp();
    const uint32_t r = tmp[lane];
    __syncwarp();
    return r;
}

__global__ void __launch_bounds__(SK_THREADS)
sketch_kernel(const SketchParams p)
{
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t* pk   = smem;                                 // [ngroups + 2]
    uint32_t* amb  = pk + p.ngroups + 2;                   // [ngroups + 2]
    uint32_t* H    = amb + p.ngroups + 2;                  // [16 * pitch]
    uint32_t* cbuf = H + 16 * p.pitch;                     // [SK_WARPS * SK_CBUF]
    uint32_t* tmpb = cbuf + SK_WARPS * SK_CBUF;            // [SK_WARPS * 32]
    int32_t*  wpos = reinterpret_cast<int32_t*>(tmpb + SK_WARPS * 32);   // [wpc]
    int32_t*  wlen = wpos + p.wpc;                         // [wpc]
    uint32_t* wout = reinterpret_cast<uint32_t*>(wlen + p.wpc);          // [wpc]
    __shared__ int64_t s_abs[64];
    __shared__ int64_t s_span_begin;
    __shared__ int32_t s_span_len;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const uint64_t g0 = uint64_t(blockIdx.x) * p.wpc;
    const int nw = int(min(uint64_t(p.wpc), p.nwin - g0));

    // ---- resolve the CTA's windows: global window -> (sequence, window) -> bytes
    if(tid < nw) {
        const uint32_t g = uint32_t(g0 + tid);
        int64_t lo = 0, hi = p.nseq;               // largest q with win_first[q] <= g
        while(hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if(p.win_first[mid] <= g) lo = mid; else hi = mid;
        }
        const int64_t q = lo;
        const int64_t w = g - p.win_first[q];
        const int64_t so = p.seq_off[q];
        const int64_t len = p.seq_off[q + 1] - so;
        int64_t b = 0, e = len;
        if(len > p.winlen) { b = w * p.stride; e = min(b + p.winlen, len); }
        s_abs[tid] = so + b;
        wlen[tid] = int32_t(e - b);
        wout[tid] = p.out_first[q] + uint32_t(w);
    }
    __syncthreads();
    if(tid == 0) {
        const int64_t a0 = s_abs[0];
        const int64_t sb = a0 - int64_t((reinterpret_cast<uintptr_t>(p.seq) + uint64_t(a0)) & 15u);
        s_span_begin = sb;
        s_span_len = int32_t(s_abs[nw - 1] + wlen[nw - 1] - sb);
    }
    __syncthreads();
    const int64_t span_begin = s_span_begin;
    const int span_len = s_span_len;
    if(tid < nw) wpos[tid] = int32_t(s_abs[tid] - span_begin);
    const int ng = (span_len + 15) >> 4;           // groups holding data

    // ---- (1) load + pack
    for(int j = tid; j < ng + 2; j += SK_THREADS) {
        uint32_t w = 0, a = 0xFFFFu;
        if(j < ng) {
            const int64_t off = span_begin + int64_t(j) * 16;
            uint4 v;
            if(off >= 0 && off + 16 <= p.total_bytes) {
                v = __ldg(reinterpret_cast<const uint4*>(p.seq + off));
            } else {
                uint32_t t[4] = {0x4E4E4E4Eu, 0x4E4E4E4Eu, 0x4E4E4E4Eu, 0x4E4E4E4Eu};   // 'N'
                for(int i = 0; i < 16; ++i) {
                    const int64_t o = off + i;
                    if(o >= 0 && o < p.total_bytes) {
                        t[i >> 2] = (t[i >> 2] & ~(0xFFu << ((i & 3) * 8))) | (uint32_t(p.seq[o]) << ((i & 3) * 8));
                    }
                }
                v = make_uint4(t[0], t[1], t[2], t[3]);
            }
            uint32_t c0, c1, c2, c3, d0, d1, d2, d3;
            pack4(v.x, c0, d0);
            pack4(v.y, c1, d1);
            pack4(v.z, c2, d2);
            pack4(v.w, c3, d3);
            w = (c0 << 24) | (c1 << 16) | (c2 << 8) | c3;
            a = 0;
            if((d0 | d1 | d2 | d3) != 0u) {
                a = (nzmask4(d0) << 12) | (nzmask4(d1) << 8) | (nzmask4(d2) << 4) | nzmask4(d3);
            }
        }
        pk[j] = w;
        amb[j] = a;
    }
    __syncthreads();

    // ---- (2) hash every k-mer start of the span
    {
        const int k = p.k;
        const int fsh = 32 - 2 * k;
        const uint32_t kmask = 0xFFFFFFFFu >> fsh;
        const int ash = 32 - k;
        for(int j = tid; j < ng; j += SK_THREADS) {
            const uint32_t w0 = pk[j], w1 = pk[j + 1];
            const uint32_t r0 = rc16(w0), r1 = rc16(w1);
            const uint32_t a = (amb[j] << 16) | amb[j + 1];
            uint32_t* hrow = H + j;
#pragma unroll
            for(int r = 0; r < 16; ++r) {
                const uint32_t f = __funnelshift_l(w1, w0, 2 * r) >> fsh;
                const uint32_t c = __funnelshift_r(r0, r1, 2 * r) & kmask;
                uint32_t h = tmh(min(f, c));
                if(((a << r) >> ash) != 0u) h = NOF;
                hrow[r * p.pitch] = h;
            }
        }
    }
    __syncthreads();

    // ---- (3) per-window bottom-s unique
    uint32_t* mycb = cbuf + warp * SK_CBUF;
    uint32_t* mytmp = tmpb + warp * 32;
    const uint32_t ltmask = (1u << lane) - 1u;
    for(int wi = warp; wi < nw; wi += SK_WARPS) {
        const int pos = wpos[wi];
        const int nk = wlen[wi] - p.k + 1;
        const uint32_t oidx = wout[wi];
        const int sp = nk < p.s ? (nk < 0 ? 0 : nk) : p.s;
        uint32_t kept = NOF;
        if(sp > 0) {
            uint32_t T = NOF;
            if(nk > 32) {
                const float pr = (1.45f * float(sp) + 2.0f) / float(nk);
                if(pr < 1.0f) T = uint32_t(pr * 4294967296.0f);
            }
            const bool wide = sp > 16;
            for(;;) {
                kept = NOF;
                bool first = true;
                if(nk <= SK_CBUF) {
                    int cnt = 0;
                    for(int c0 = 0; c0 < nk; c0 += 32) {
                        const int q = pos + c0 + lane;
                        const uint32_t h = (c0 + lane < nk) ? H[(q & 15) * p.pitch + (q >> 4)] : NOF;
                        const bool pass = h < T;
                        const uint32_t b = __ballot_sync(0xFFFFFFFFu, pass);
                        if(pass) mycb[cnt + __popc(b & ltmask)] = h;
                        cnt += __popc(b);
                    }
                    __syncwarp();
                    for(int c0 = 0; c0 < cnt; c0 += 32) {
                        const uint32_t v = (c0 + lane < cnt) ? mycb[c0 + lane] : NOF;
                        kept = merge_chunk(kept, v, first, wide, lane, mytmp);
                        first = false;
                    }
                    __syncwarp();
                } else {
                    for(int c0 = 0; c0 < nk; c0 += 32) {
                        const int q = pos + c0 + lane;
                        uint32_t h = (c0 + lane < nk) ? H[(q & 15) * p.pitch + (q >> 4)] : NOF;
                        if(h >= T) h = NOF;
                        if(__ballot_sync(0xFFFFFFFFu, h != NOF) == 0u) continue;
                        kept = merge_chunk(kept, h, first, wide, lane, mytmp);
                        first = false;
                    }
                }
                const int nuniq = __popc(__ballot_sync(0xFFFFFFFFu, kept != NOF));
                if(nuniq >= sp || T == NOF) break;
                T = NOF;        // too few survivors below the threshold: exact pass
            }
        }
        if(lane < p.s) p.out_feat[uint64_t(oidx) * p.s + lane] = (lane < sp) ? kept : NOF;
        if(p.out_cnt != nullptr) {
            const int c = __popc(__ballot_sync(0xFFFFFFFFu, lane < sp && kept != NOF));
            if(lane == 0) p.out_cnt[oidx] = uint32_t(c);
        }
    }
}

// ---- window counts per sequence -----------------------------------------------------
__global__ void window_count_kernel(const int64_t* seq_off, int64_t nseq, int64_t winlen, int64_t stride,
                                    uint32_t* out)
{
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if(i < nseq) out[i] = uint32_t(num_windows(seq_off[i + 1] - seq_off[i], winlen, stride));
}

int exclusive_scan_u32_inplace(mcg_ctx* c, uint32_t* d, uint64_t n, uint64_t* h_total);

int window_prefix(mcg_ctx* c, int64_t winlen, int64_t stride, const int64_t* d_seq_off, int64_t nseq,
                  uint32_t* d_win_first, uint64_t* h_total)
{
    if(nseq > 0) {
        window_count_kernel<<<unsigned(div_up(nseq, 256)), 256, 0, c->stream>>>(d_seq_off, nseq, winlen, stride, d_win_first);
        c->launches++;
        MCG_CUDA(cudaGetLastError());
    }
    return exclusive_scan_u32_inplace(c, d_win_first, uint64_t(nseq), h_total);
}

int sketch_config(const mcg_sketch_cfg& cfg, int sketchlen, int64_t winlen, int64_t stride,
                  int* wpc, int* ngroups, int* pitch, size_t* smem)
{
    MCG_REQUIRE(cfg.k >= 1 && cfg.k <= MCG_MAX_K, "kmerlen %d outside 1..%d", cfg.k, MCG_MAX_K);
    MCG_REQUIRE(sketchlen >= 1 && sketchlen <= MCG_MAX_SKETCH, "sketchlen %d outside 1..%d", sketchlen, MCG_MAX_SKETCH);
    MCG_REQUIRE(winlen >= 1 && stride >= 1, "window length and stride must be >= 1");
    const int64_t adv = winlen > stride ? winlen : stride;
    MCG_REQUIRE(adv <= 32768, "window length / stride above 32768 are not supported on the device");
    int w = 32;
    for(;;) {
        const int64_t span = (int64_t(w) * adv + 15 + 15) / 16 * 16 + 16;
        const int ng = int(span / 16);
        const int pt = ((ng + 31) / 32) * 32 + 2;
        const size_t bytes = (size_t(2) * (ng + 2) + size_t(16) * pt + SK_WARPS * SK_CBUF + SK_WARPS * 32 + 3 * size_t(w)) * 4;
        if(bytes <= 200 * 1024 || w == 1) {
            MCG_REQUIRE(bytes <= 200 * 1024, "window too long for the device sketcher");
            *wpc = w; *ngroups = ng; *pitch = pt; *smem = bytes;
            return MCG_OK;
        }
        w >>= 1;
    }
}

int sketch_launch(mcg_ctx* c, const mcg_sketch_cfg& cfg, int sketchlen, int64_t winlen, int64_t stride,
                  const char* d_seq, int64_t total_bytes, const int64_t* d_seq_off, int64_t nseq,
                  const uint32_t* d_win_first, const uint32_t* d_out_first, uint64_t nwin,
                  uint32_t* d_out_feat, uint32_t* d_out_cnt)
{
    if(nwin == 0) return MCG_OK;
    SketchParams p;
    size_t smem = 0;
    MCG_TRY(sketch_config(cfg, sketchlen, winlen, stride, &p.wpc, &p.ngroups, &p.pitch, &smem));
    p.seq = reinterpret_cast<const uint8_t*>(d_seq);
    p.total_bytes = total_bytes;
    p.seq_off = d_seq_off;
    p.nseq = nseq;
    p.win_first = d_win_first;
    p.out_first = d_out_first;
    p.nwin = nwin;
    p.out_feat = d_out_feat;
    p.out_cnt = d_out_cnt;
    p.k = cfg.k;
    p.s = sketchlen;
    p.winlen = winlen;
    p.stride = stride;
    static size_t configured = 0;
    if(smem > 48 * 1024 && smem > configured) {
        MCG_CUDA(cudaFuncSetAttribute(sketch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        configured = smem;
    }
    const uint64_t nblocks = (nwin + p.wpc - 1) / p.wpc;
    MCG_REQUIRE(nblocks < (1ull << 31), "too many windows in one launch");
    sketch_kernel<<<unsigned(nblocks), SK_THREADS, smem, c->stream>>>(p);
    c->launches++;
    MCG_CUDA(cudaGetLastError());
    return MCG_OK;
}

}  // namespace mcg
