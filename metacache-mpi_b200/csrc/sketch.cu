// sketch.cu -- K1, warp-per-window formulation (the path for sketchlen > 16 or windows > 512
// bases; everything else runs sketch_tpw.cu) + the host side of both: window prefix, tile table,
// launch.  2-bit encode + canonical k-mer + thomas_mueller_hash + per-window bottom-s unique
// selection, fused.  Must equal
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
#include "sketch_dev.cuh"

#include <algorithm>

namespace mcg {


// hashes the 16 k-mer starts of group j (K16: k == 16 known at compile time; PITCH > 0:
// row pitch of the transposed hash array known at compile time)
template<bool K16, bool CHECK, int PITCH>
__device__ __forceinline__ void hash_group(uint32_t w0, uint32_t w1, uint32_t a, int k, uint32_t* hrow, int pitch_rt)
{
    const int pitch = PITCH > 0 ? PITCH : pitch_rt;
    const uint32_t r0 = rc16(w0), r1 = rc16(w1);
    const int fsh = 32 - 2 * k;
    const uint32_t kmask = 0xFFFFFFFFu >> fsh;
    const int ash = 32 - k;
#pragma unroll
    for(int r = 0; r < 16; ++r) {
        uint32_t f = __funnelshift_l(w1, w0, 2 * r);
        uint32_t c = __funnelshift_r(r0, r1, 2 * r);
        if(!K16) { f >>= fsh; c &= kmask; }
        uint32_t y = tmh_head(min(f, c));
        if(CHECK) { if(((a << r) >> ash) != 0u) y = NOH; }
        hrow[r * pitch] = y;
    }
}

// fast path of the selection: threshold filter + ballot compaction of the (<= 128) k-mer
// hashes of one window into cb; returns the number of survivors
__device__ __forceinline__ int filter_compact(const uint32_t* hp, int rel0, int nk, uint32_t T, uint32_t* cb, uint32_t ltmask)
{
    int cnt = 0;
    if(T != 0u) {
#pragma unroll
        for(int i = 0; i < 4; ++i) {
            const uint32_t y = hp[2 * i];
            const bool pass = (uint32_t(rel0 + 32 * i) < uint32_t(nk)) && (y < T);
            const uint32_t b = __ballot_sync(0xFFFFFFFFu, pass);
            if(pass) cb[cnt + __popc(b & ltmask)] = y;
            cnt += __popc(b);
        }
    } else {
#pragma unroll
        for(int i = 0; i < 4; ++i) {
            const uint32_t y = hp[2 * i];
            const bool pass = (uint32_t(rel0 + 32 * i) < uint32_t(nk)) && (y != NOH);
            const uint32_t b = __ballot_sync(0xFFFFFFFFu, pass);
            if(pass) cb[cnt + __popc(b & ltmask)] = y;
            cnt += __popc(b);
        }
    }
    return cnt;
}

struct SelOut { uint32_t kept; int nout; };

__device__ __noinline__ SelOut chunked_select(const uint32_t* cb, int cnt, int sp, int lane, uint32_t* tmp)
{
    const bool wide = sp > 16;
    uint32_t kept = NOF;
    bool first = true;
    for(int c0 = 0; c0 < cnt; c0 += 32) {
        const uint32_t v = (c0 + lane < cnt) ? tmh_tail(cb[c0 + lane]) : NOF;
        kept = merge_chunk(kept, v, first, wide, lane, tmp);
        first = false;
    }
    SelOut o;
    o.kept = kept;
    o.nout = __popc(__ballot_sync(0xFFFFFFFFu, kept != NOF));
    return o;
}

// exact selection over all k-mer hashes of a window, 32 at a time (any window length)
__device__ __noinline__ SelOut exact_select(const uint32_t* hp, int rel0, int nk, int span, int sp, int lane,
                                            uint32_t* tmp)
{
    const bool wide = sp > 16;
    uint32_t kept = NOF;
    bool first = true;
    for(int i = 0; i * 32 < span; ++i) {
        const uint32_t y = hp[2 * i];
        const bool in = uint32_t(rel0 + 32 * i) < uint32_t(nk);
        const uint32_t h = (in && y != NOH) ? tmh_tail(y) : NOF;
        if(__ballot_sync(0xFFFFFFFFu, h != NOF) == 0u) continue;
        kept = merge_chunk(kept, h, first, wide, lane, tmp);
        first = false;
    }
    SelOut o;
    o.kept = kept;
    o.nout = __popc(__ballot_sync(0xFFFFFFFFu, kept != NOF));
    return o;
}

// threshold with ~1.5 s expected survivors among nk hashes; a multiple of 65536 (see NOH);
// 0 = accept everything
__device__ __forceinline__ uint32_t select_threshold(int nk, int sp)
{
    if(nk <= 32) return 0u;
    const float pr = 1.5f * float(sp) / float(nk);
    return pr < 0.999f ? (uint32_t(pr * 65536.0f) << 16) : 0u;
}


// ---- per-warp selection state of one window
struct Win {
    const uint32_t* hp;     // lane's first hash slot
    int rel0;               // lane's k-mer index of that slot (may be negative)
    int nk, sp, span;
    uint32_t T, oidx;
};

// Threshold filter (+ one refinement when more than 32 hashes survive).  Returns the number of
// survivors in cb, or -1 when the window needs the exact pass.
__device__ __forceinline__ int prefilter(const Win& w, uint32_t* cb, int lane, uint32_t ltmask)
{
    int cnt = filter_compact(w.hp, w.rel0, w.nk, w.T, cb, ltmask);
    __syncwarp();
    if(cnt <= 32) return (cnt >= w.sp || w.T == 0u) ? cnt : -1;
    if(w.T == 0u) return cnt;                       // no threshold: chunked merge of everything
    // too many survivors: filter them again with 3/4 of the threshold
    const uint32_t T2 = ((w.T >> 16) * 3u / 4u) << 16;
    uint32_t ys[4];
#pragma unroll
    for(int i = 0; i < 4; ++i) ys[i] = (lane + 32 * i < cnt) ? cb[lane + 32 * i] : NOH;
    __syncwarp();
    int c2 = 0;
#pragma unroll
    for(int i = 0; i < 4; ++i) {
        const bool pass = ys[i] < T2;
        const uint32_t b = __ballot_sync(0xFFFFFFFFu, pass);
        if(pass) cb[c2 + __popc(b & ltmask)] = ys[i];
        c2 += __popc(b);
    }
    __syncwarp();
    return (c2 >= w.sp && c2 <= 32) ? c2 : -1;
}

// two independent 32-lane sorts, interleaved stage by stage (two shuffles in flight)
__device__ __forceinline__ void sort32x2(uint32_t& a, uint32_t& b, int lane)
{
#pragma unroll
    for(int k = 2; k <= 32; k <<= 1) {
        {
            const uint32_t oa = __shfl_xor_sync(0xFFFFFFFFu, a, k - 1);
            const uint32_t ob = __shfl_xor_sync(0xFFFFFFFFu, b, k - 1);
            const bool up = (lane & (k >> 1)) != 0;
            a = up ? max(a, oa) : min(a, oa);
            b = up ? max(b, ob) : min(b, ob);
        }
#pragma unroll
        for(int j = k >> 2; j > 0; j >>= 1) {
            const uint32_t oa = __shfl_xor_sync(0xFFFFFFFFu, a, j);
            const uint32_t ob = __shfl_xor_sync(0xFFFFFFFFu, b, j);
            const bool up = (lane & j) != 0;
            a = up ? max(a, oa) : min(a, oa);
            b = up ? max(b, ob) : min(b, ob);
        }
    }
}

// finishes one window whose (<= 32 or unfiltered) survivors sit in cb: sorted/unique/cut
__device__ __forceinline__ uint32_t finish_window(const Win& w, int cnt, uint32_t v, bool sorted, const uint32_t* cb,
                                                  int lane, uint32_t* tmp, int& nout)
{
    uint32_t kept = NOF;
    bool done = false;
    nout = 0;
    if(w.sp <= 0) return NOF;
    if(sorted) {
        const uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, v, 1);
        const bool dup = (lane > 0) && (v == prev) && (v != NOF);
        if(__any_sync(0xFFFFFFFFu, dup)) {
            kept = dedup_sorted(v, lane, tmp);
            nout = __popc(__ballot_sync(0xFFFFFFFFu, kept != NOF));
        } else {
            kept = v;
            nout = cnt;
        }
        done = (nout >= w.sp) || (w.T == 0u);
    } else if(cnt > 32) {                            // only without threshold
        const SelOut o = chunked_select(cb, cnt, w.sp, lane, tmp);
        kept = o.kept;
        nout = o.nout;
        done = true;
    }
    if(!done) {
        const SelOut o = exact_select(w.hp, w.rel0, w.nk, w.span, w.sp, lane, tmp);
        kept = o.kept;
        nout = o.nout;
    }
    if(nout > w.sp) nout = w.sp;
    return kept;
}

// Persistent, warp-autonomous kernel.  A warp owns tiles of `wpt` consecutive windows (4 at the
// default shape: one 16-base group per lane).  No block-wide barrier exists: all staging lives in
// the warp's private slice of shared memory and is ordered by __syncwarp / the warp's own
// mbarriers.  While a tile is processed, the warp has already resolved the next tile and the TMA
// unit (cp.async.bulk) is copying its bytes into the warp's other raw buffer.
template<bool K16, int PITCH>
__global__ void __launch_bounds__(SK_THREADS, 4)
sketch_kernel(const SketchParams p)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int pitch = PITCH > 0 ? PITCH : p.pitch;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int ngc = p.ngroups;                               // group capacity of a tile span
    const int rawcap = (ngc + 2) * 16;
    // per-warp slice
    uint8_t*  base = smem_raw + size_t(warp) * p.warp_smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(base);      // [2]
    uint8_t*  raw0 = base + 16;
    uint8_t*  raw1 = raw0 + rawcap;
    uint32_t* pk   = reinterpret_cast<uint32_t*>(raw1 + rawcap);   // [ngc + 2]
    uint32_t* amb  = pk + ngc + 2;                           // [ngc + 2]
    uint32_t* H    = amb + ngc + 2;                          // [16 * pitch]
    uint32_t* cb0  = H + 16 * pitch;                         // [SK_CBUF]
    uint32_t* cb1  = cb0 + SK_CBUF;                          // [SK_CBUF]
    uint32_t* tmp  = cb1 + SK_CBUF;                          // [32]
    int32_t*  tabs = reinterpret_cast<int32_t*>(tmp + 32);   // 2 x (4 * wpt + 2)
    const int wpt = p.wpc;
    const int tabsz = 4 * wpt + 2;
    const uint64_t nwin_end = sk_nwin(p);
    const uint64_t ntiles = nwin_end > p.win_lo ? (nwin_end - p.win_lo + wpt - 1) / wpt : 0;
    const uint64_t nwarps = uint64_t(gridDim.x) * SK_WARPS;
    const uint64_t gw = uint64_t(blockIdx.x) * SK_WARPS + warp;

    // resolve tile `tile` into table `stage` and start the TMA copy of its bytes
    auto prefetch = [&](uint64_t tile, int stage) {
        int32_t* tb = tabs + stage * tabsz;
        uint32_t* tbu = reinterpret_cast<uint32_t*>(tb);
        const uint64_t g0 = p.win_lo + tile * uint64_t(wpt);
        const int nw = int(min(uint64_t(wpt), nwin_end - g0));
        int64_t first_abs = 0, last_end = 0;
        const uint32_t q0 = p.cta_seq[tile];
        const uint32_t wf0 = p.win_first[q0], wf1 = p.win_first[q0 + 1];
        if(g0 + uint64_t(nw) <= uint64_t(wf1) && nw <= 32) {
            // common case: the whole tile lies in one sequence
            const int64_t so = p.seq_off[q0];
            const int64_t len = p.seq_off[q0 + 1] - so;
            const uint32_t of = p.out_first[q0];
            const uint32_t w = uint32_t(g0 - wf0) + uint32_t(lane);
            int64_t b = 0, e = len;
            if(len > p.winlen) { b = int64_t(w) * p.stride; e = min(b + p.winlen, len); }
            const int nk = int(e - b) - p.k + 1;
            first_abs = __shfl_sync(0xFFFFFFFFu, so + b, 0);
            last_end = __shfl_sync(0xFFFFFFFFu, so + e, nw - 1);
            const int sb_lo = int((reinterpret_cast<uintptr_t>(p.seq) + uint64_t(first_abs)) & 15u);
            if(lane < nw) {
                tb[lane] = int32_t(so + b - first_abs) + sb_lo;
                tb[wpt + lane] = nk;
                tbu[2 * wpt + lane] = of + w;
                tbu[3 * wpt + lane] = (nk == p.nk_full) ? p.thr_full : select_threshold(nk, min(p.s, max(nk, 0)));
            }
        } else {
            for(int i0 = 0; i0 < nw; i0 += 32) {
                const int i = i0 + lane;
                int64_t abs_i = 0, end_i = 0;
                if(i < nw) {
                    const uint32_t g = uint32_t(g0 + i);
                    int64_t q = q0;
                    while(p.win_first[q + 1] <= g) ++q;          // few steps: tiles are short
                    const int64_t w = g - p.win_first[q];
                    const int64_t so = p.seq_off[q];
                    const int64_t len = p.seq_off[q + 1] - so;
                    int64_t b = 0, e = len;
                    if(len > p.winlen) { b = w * p.stride; e = min(b + p.winlen, len); }
                    abs_i = so + b;
                    end_i = so + e;
                    const int nk = int(e - b) - p.k + 1;
                    tb[wpt + i] = nk;
                    tbu[2 * wpt + i] = p.out_first[q] + uint32_t(w);
                    tbu[3 * wpt + i] = select_threshold(nk, min(p.s, max(nk, 0)));
                }
                if(i0 == 0) first_abs = __shfl_sync(0xFFFFFFFFu, abs_i, 0);
                const int sb_lo = int((reinterpret_cast<uintptr_t>(p.seq) + uint64_t(first_abs)) & 15u);
                if(i < nw) tb[i] = int32_t(abs_i - first_abs) + sb_lo;
                if(i0 + 32 >= nw) last_end = __shfl_sync(0xFFFFFFFFu, end_i, (nw - 1) & 31);
            }
        }
        const int64_t sb = first_abs - int64_t((reinterpret_cast<uintptr_t>(p.seq) + uint64_t(first_abs)) & 15u);
        const int span_len = int(last_end - sb);
        if(lane == 0) {
            tb[4 * wpt] = nw;
            tb[4 * wpt + 1] = span_len;
            const uint32_t bytes = uint32_t(((span_len + 15) >> 4) + 1) * 16u;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&bars[stage], bytes);
            tma_load_1d(stage ? raw1 : raw0, p.seq + sb, bytes, &bars[stage]);
        }
        __syncwarp();
    };

    if(lane == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if(gw < ntiles) prefetch(gw, 0);

    const uint32_t ltmask = (1u << lane) - 1u;
    const uint32_t* hlane = H + (lane & 15) * pitch + (lane >> 4);
    uint32_t* olane = p.out_feat + lane;

    int it = 0;
    for(uint64_t tile = gw; tile < ntiles; tile += nwarps, ++it) {
        const int stage = it & 1;
        const uint32_t parity = (it >> 1) & 1;
        if(tile + nwarps < ntiles) prefetch(tile + nwarps, stage ^ 1);
        const int32_t* tb = tabs + stage * tabsz;
        const int nw = tb[4 * wpt];
        const int ng = (tb[4 * wpt + 1] + 15) >> 4;
        mbar_wait(&bars[stage], parity);

        // ---- (1) pack: raw ASCII (shared) -> 2-bit words + ambiguity bits
        const uint4* rawv = reinterpret_cast<const uint4*>(stage ? raw1 : raw0);
        for(int j = lane; j < ng + 1; j += 32) {
            const uint4 v = rawv[j];
            uint32_t c0, c1, c2, c3, d0, d1, d2, d3;
            pack4(v.x, c0, d0);
            pack4(v.y, c1, d1);
            pack4(v.z, c2, d2);
            pack4(v.w, c3, d3);
            uint32_t a = 0;
            if((d0 | d1 | d2 | d3) != 0u) {
                a = (nzmask4(d0) << 12) | (nzmask4(d1) << 8) | (nzmask4(d2) << 4) | nzmask4(d3);
            }
            pk[j] = (c0 << 24) | (c1 << 16) | (c2 << 8) | c3;
            amb[j] = a;
        }
        __syncwarp();

        // ---- (2) hash every k-mer start of the span
        for(int j = lane; j < ng; j += 32) {
            const uint32_t w0 = pk[j], w1 = pk[j + 1];
            const uint32_t a = (amb[j] << 16) | amb[j + 1];
            if(a == 0u) hash_group<K16, false, PITCH>(w0, w1, a, p.k, H + j, pitch);
            else        hash_group<K16, true, PITCH>(w0, w1, a, p.k, H + j, pitch);
        }
        __syncwarp();

        // ---- (3) bottom-s unique, two windows at a time
        for(int wi = 0; wi < nw; wi += 2) {
            const bool two = wi + 1 < nw;
            const int wj = two ? wi + 1 : wi;
            Win A, B;
            {
                const int pos = tb[wi];
                A.nk = tb[wpt + wi];
                A.oidx = reinterpret_cast<const uint32_t*>(tb)[2 * wpt + wi];
                A.T = reinterpret_cast<const uint32_t*>(tb)[3 * wpt + wi];
                A.sp = min(p.s, max(A.nk, 0));
                A.hp = hlane + (pos >> 4);
                A.rel0 = lane - (pos & 15);
                A.span = A.nk + (pos & 15);
            }
            {
                const int pos = tb[wj];
                B.nk = tb[wpt + wj];
                B.oidx = reinterpret_cast<const uint32_t*>(tb)[2 * wpt + wj];
                B.T = reinterpret_cast<const uint32_t*>(tb)[3 * wpt + wj];
                B.sp = two ? min(p.s, max(B.nk, 0)) : 0;
                B.hp = hlane + (pos >> 4);
                B.rel0 = lane - (pos & 15);
                B.span = B.nk + (pos & 15);
            }
            int cntA = -1, cntB = -1;
            if(A.sp > 0 && A.span <= 128) cntA = prefilter(A, cb0, lane, ltmask);
            if(B.sp > 0 && B.span <= 128) cntB = prefilter(B, cb1, lane, ltmask);
            const bool sA = cntA >= 0 && cntA <= 32, sB = cntB >= 0 && cntB <= 32;
            uint32_t va = (sA && lane < cntA) ? tmh_tail(cb0[lane]) : NOF;
            uint32_t vb = (sB && lane < cntB) ? tmh_tail(cb1[lane]) : NOF;
            if(sA || sB) sort32x2(va, vb, lane);
            int noutA, noutB;
            const uint32_t keptA = finish_window(A, cntA, va, sA, cb0, lane, tmp, noutA);
            if(lane < p.s) olane[uint64_t(A.oidx) * uint32_t(p.s)] = (lane < A.sp) ? keptA : NOF;
            if(p.out_cnt != nullptr && lane == 0) p.out_cnt[A.oidx] = uint32_t(noutA);
            if(two) {
                const uint32_t keptB = finish_window(B, cntB, vb, sB, cb1, lane, tmp, noutB);
                if(lane < p.s) olane[uint64_t(B.oidx) * uint32_t(p.s)] = (lane < B.sp) ? keptB : NOF;
                if(p.out_cnt != nullptr && lane == 0) p.out_cnt[B.oidx] = uint32_t(noutB);
            }
            __syncwarp();
        }
    }
}

// sequence that holds the first window of every CTA (one binary search per CTA, done up front
// so that the main kernel starts with one load instead of a chain of dependent ones)
__global__ void cta_seq_kernel(const uint32_t* win_first, int64_t nseq, int wpc, uint32_t nblocks, uint64_t win_lo, uint32_t* cta_seq,
                               uint32_t* tile_ctr)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if(b == 0 && tile_ctr != nullptr) *tile_ctr = 0u;
    if(b > nblocks) return;
    if(b == nblocks) { cta_seq[b] = uint32_t(nseq - 1); return; }
    const uint32_t g = uint32_t(win_lo + uint64_t(b) * uint32_t(wpc));
    int64_t lo = 0, hi = nseq;
    while(hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if(win_first[mid] <= g) lo = mid; else hi = mid;
    }
    cta_seq[b] = uint32_t(lo);
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

// sketch_tpw.cu
bool tpw_eligible(int sketchlen, int64_t winlen, int64_t stride);
void tpw_config(int64_t winlen, int64_t stride, int rounds, int* ngroups, size_t* warp_smem);
int tpw_launch(mcg_ctx* c, SketchParams& p, const mcg_sketch_cfg& cfg, uint64_t nblocks, cudaStream_t stream);

int sketch_config(const mcg_sketch_cfg& cfg, int sketchlen, int64_t winlen, int64_t stride,
                  int* wpt, int* ngroups, int* pitch, size_t* warp_smem)
{
    int dummy_i; size_t dummy_s;
    if(!ngroups) ngroups = &dummy_i;
    if(!pitch) pitch = &dummy_i;
    if(!warp_smem) warp_smem = &dummy_s;
    MCG_REQUIRE(cfg.k >= 1 && cfg.k <= MCG_MAX_K, "kmerlen %d outside 1..%d", cfg.k, MCG_MAX_K);
    MCG_REQUIRE(sketchlen >= 1 && sketchlen <= MCG_MAX_SKETCH, "sketchlen %d outside 1..%d", sketchlen, MCG_MAX_SKETCH);
    MCG_REQUIRE(winlen >= 1 && stride >= 1, "window length and stride must be >= 1");
    const int64_t adv = winlen > stride ? winlen : stride;
    MCG_REQUIRE(adv <= 8192, "window length / stride above 8192 are not supported on the device");
    if(tpw_eligible(sketchlen, winlen, stride)) {
        // thread-per-window kernel: 32 windows per warp tile
        *wpt = 32;
        *pitch = 0;
        tpw_config(winlen, stride, 1, ngroups, warp_smem);
        return MCG_OK;
    }
    // windows per warp tile: a typical span (same sequence) fits one 16-base group per lane
    static int span_target = -1;
    if(span_target < 0) { const char* e = getenv("MCG_SK_SPAN"); span_target = e ? atoi(e) : 1024; }
    int64_t w = (span_target - 30 - winlen) / stride + 1;
    if(winlen > span_target - 30) w = 1;
    if(w > 16) w = 16;
    if(w < 1) w = 1;
    // capacity for the worst case (every window in a different sequence)
    const int ng = int((w * adv + 30 + 15) / 16) + 1;
    int pt = ng + 4;
    while((pt & 3) != 2) ++pt;                  // pitch/2 odd: conflict-free transposed reads
    size_t bytes = 16 + size_t(2) * (ng + 2) * 16 + (size_t(2) * (ng + 2) + size_t(16) * pt + 2 * SK_CBUF + 32 + 2 * (4 * size_t(w) + 2)) * 4;
    bytes = (bytes + 127) & ~size_t(127);
    MCG_REQUIRE(bytes * SK_WARPS <= 220 * 1024, "window too long for the device sketcher");
    *wpt = int(w); *ngroups = ng; *pitch = pt; *warp_smem = bytes;
    return MCG_OK;
}

int sketch_launch(mcg_ctx* c, const mcg_sketch_cfg& cfg, int sketchlen, int64_t winlen, int64_t stride,
                  const char* d_seq, int64_t total_bytes, const int64_t* d_seq_off, int64_t nseq,
                  const uint32_t* d_win_first, const uint32_t* d_out_first, uint64_t nwin,
                  uint32_t* d_out_feat, uint32_t* d_out_cnt, uint64_t win_lo, cudaStream_t stream, uint32_t* d_tile_scratch,
                  const uint32_t* d_nwin)
{
    if(nwin <= win_lo) return MCG_OK;
    if(stream == nullptr) stream = c->stream;
    SketchParams p;
    size_t wsmem = 0;
    MCG_TRY(sketch_config(cfg, sketchlen, winlen, stride, &p.wpc, &p.ngroups, &p.pitch, &wsmem));
    p.warp_smem = int(wsmem);
    const size_t smem = wsmem * SK_WARPS;
    p.seq = reinterpret_cast<const uint8_t*>(d_seq);
    p.total_bytes = total_bytes;
    p.seq_off = d_seq_off;
    p.nseq = nseq;
    p.win_first = d_win_first;
    p.out_first = d_out_first;
    p.nwin = nwin;
    p.nwin_dev = d_nwin;
    p.win_lo = win_lo;
    {
        const int nk = int(winlen) - cfg.k + 1;
        const int sp = std::min(sketchlen, std::max(nk, 0));
        p.nk_full = nk;
        p.thr_full = 0u;
        if(nk > 32) {
            const float pr = 1.5f * float(sp) / float(nk);
            if(pr < 0.999f) p.thr_full = uint32_t(pr * 65536.0f) << 16;
        }
    }
    p.out_feat = d_out_feat;
    p.out_cnt = d_out_cnt;
    p.k = cfg.k;
    p.s = sketchlen;
    p.winlen = winlen;
    p.stride = stride;
    const bool tpw = tpw_eligible(sketchlen, winlen, stride);
    if(tpw && nseq > 0 && total_bytes / nseq < 4 * winlen) {
        // read-like input (long and short windows alternate): 64-window tiles handled in two
        // interleaved rounds, so that the lanes of a round have windows of one kind
        int ng2 = 0;
        size_t ws2 = 0;
        tpw_config(winlen, stride, 2, &ng2, &ws2);
        static const bool no_rounds = getenv("MCG_SK_NO_ROUNDS") != nullptr;
        if(ws2 * 4 <= 56 * 1024 && !no_rounds) { p.wpc = 64; p.ngroups = ng2; p.warp_smem = int(ws2); }
    }
    const uint64_t nblocks = (nwin - win_lo + p.wpc - 1) / p.wpc;
    MCG_REQUIRE(nblocks < (1ull << 31), "too many windows in one launch");
    MCG_REQUIRE(nseq < (1ll << 32), "too many sequences in one launch");
    // per-launch tile table (pipelined callers pass disjoint slices of a buffer they reserved)
    uint32_t* tile_seq = d_tile_scratch;
    if(tile_seq == nullptr) {
        MCG_TRY(c->outfirst.reserve((nblocks + 4) * sizeof(uint32_t)));
        tile_seq = c->outfirst.as<uint32_t>();
    }
    MCG_TRY(c->tilectr.reserve(64));
    p.tile_ctr = c->tilectr.as<uint32_t>();
    cta_seq_kernel<<<unsigned((nblocks + 256) / 256), 256, 0, stream>>>(d_win_first, nseq, p.wpc, uint32_t(nblocks), win_lo, tile_seq,
                                                                         p.tile_ctr);
    c->launches++;
    p.cta_seq = tile_seq;
    if(tpw) {
        c->prof_begin(0, stream);
        const int rc = tpw_launch(c, p, cfg, nblocks, stream);
        c->prof_end(0, stream);
        return rc;
    }
    c->prof_begin(0, stream);
#define SK_LAUNCH(K16, PT)                                                                                        \
    do {                                                                                                          \
        auto kern = sketch_kernel<K16, PT>;                                                                       \
        if(smem > 48 * 1024) MCG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem))); \
        int per_sm = 1;                                                                                           \
        MCG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SK_THREADS, smem));                 \
        if(per_sm < 1) per_sm = 1;                                                                                \
        const uint64_t grid = std::min<uint64_t>((nblocks + SK_WARPS - 1) / SK_WARPS, uint64_t(c->num_sms) * uint64_t(per_sm));                   \
        kern<<<unsigned(grid), SK_THREADS, smem, stream>>>(p);                                                 \
    } while(0)
    if(cfg.k == 16 && p.pitch == SK_PITCH_DEFAULT) SK_LAUNCH(true, SK_PITCH_DEFAULT);
    else if(cfg.k == 16)                           SK_LAUNCH(true, 0);
    else                                           SK_LAUNCH(false, 0);
#undef SK_LAUNCH
    c->prof_end(0, stream);
    c->launches++;
    MCG_CUDA(cudaGetLastError());
    return MCG_OK;
}

}  // namespace mcg
