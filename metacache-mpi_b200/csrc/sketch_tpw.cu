// sketch_tpw.cu -- K1, thread-per-window formulation (sketchlen <= 16, windows <= 512 bases).
// Same contract as sketch.cu (bit-exact with dna_encoding.h:261-348,114-121,189-197,
// hash_int.h:40-45, hash_dna.h:115-152).
//
// A warp owns tiles of 32 consecutive windows, one window per lane:
//   (1) the tile's contiguous byte span arrives through TMA (cp.async.bulk + mbarrier) and is
//       packed by the whole warp to 2 bits/base + 1 ambiguity bit/base (shared memory);
//       the TMA copy of the warp's next tile is issued as soon as the raw bytes are consumed;
//   (2) every lane walks the k-mers of ITS window: 16 k-mers per step out of two aligned
//       packed words (compile-time funnel shifts), canonical k-mer, hash, and keeps the hashes
//       below a threshold with ~sketchlen+15 expected survivors in its own shared-memory column
//       (slot i of lane l lives at word i*32+l: conflict-free, no ballots, no shuffles);
//   (3) every lane sorts its <= 48 survivors in REGISTERS with a fixed compare-exchange
//       network (three 16-input networks + two bottom-half bitonic merges) and stores its
//       16 smallest.  A window whose survivors are too few / too many / not distinct is redone
//       exactly by the whole warp (rare).
// Against the warp-per-window kernel this removes the 15-stage shuffle network per window:
// ~75 instead of ~325 warp instructions per 128-base window.
#include "sketch_dev.cuh"

#include <algorithm>

namespace mcg {

constexpr int TW_THREADS = 128;
constexpr int TW_WARPS = TW_THREADS / 32;
constexpr int TW_WPT = 32;               // windows per warp tile (one per lane)
constexpr int TW_LCAP = 64;              // candidate slots per window in shared memory
constexpr int TW_SEL = 48;               // candidates the register network takes
constexpr int TW_TABSZ = 4 * TW_WPT + 2;

// threshold on the hash head (multiple of 65536, see NOH) with sp+15 expected survivors among
// nk k-mers; 0 = keep everything (short windows)
__host__ __device__ __forceinline__ uint32_t tw_threshold(int nk, int sp)
{
    if(nk <= TW_SEL) return 0u;
    const float pr = float(sp + 15) / float(nk);
    return uint32_t(pr * 65536.0f) << 16;
}

// 4 ASCII bases -> product whose top byte holds four 2-bit codes in the order A0 C1 T2 G3
// (first base in the high bits), and a per-byte validity difference (0 <=> one of ACGTacgt)
__device__ __forceinline__ void tw_pack4(uint32_t x, uint32_t& prod, uint32_t& diff)
{
    const uint32_t t = (x & 0xDFDFDFDFu) ^ 0x41414141u;      // A 00, C 02, G 06, T 15
    const uint32_t s1 = t >> 1, s2 = t >> 2;
    const uint32_t b = s1 & 0x03030303u;
    const uint32_t g = s2 & ~s1 & 0x01010101u;               // T
    diff = (t & 0xF9F9F9F9u) ^ (g * 0x11u);
    prod = b * 0x40100401u;
}

// 16 k-mers of one lane's window: starts 0..15 of the aligned word pair (A0,A1); R = rc16(A)
template<bool K16, bool CHECK>
__device__ __forceinline__ void tw_group(uint32_t A0, uint32_t A1, uint32_t R0, uint32_t R1, uint32_t a32, int k,
                                         uint32_t T, uint32_t*& lp)
{
    const int fsh = 32 - 2 * k;
    const uint32_t kmask = 0xFFFFFFFFu >> fsh;
    const int ash = 32 - k;
#pragma unroll
    for(int r = 0; r < 16; ++r) {
        uint32_t f = r ? __funnelshift_l(A1, A0, 2 * r) : A0;
        uint32_t c = r ? __funnelshift_r(R0, R1, 2 * r) : R0;
        if(!K16) { f >>= fsh; c &= kmask; }
        uint32_t y = tmh_head(min(f, c));
        if(CHECK) { if(((a32 << r) >> ash) != 0u) y = NOH; }
        if(y < T) { *lp = y; lp += 32; }
    }
}

#define SN_CE(i, j)                                       \
    {                                                     \
        const uint32_t lo__ = min(v[SN_B + (i)], v[SN_B + (j)]); \
        v[SN_B + (j)] = max(v[SN_B + (i)], v[SN_B + (j)]);       \
        v[SN_B + (i)] = lo__;                             \
    }

// v[0..15] <- the 16 smallest of v[0..15] (sorted) and v[OFF..OFF+15] (sorted), ascending
template<int OFF>
__device__ __forceinline__ void tw_merge_low(uint32_t (&v)[TW_SEL])
{
#pragma unroll
    for(int i = 0; i < 16; ++i) v[i] = min(v[i], v[OFF + 15 - i]);      // bitonic
#pragma unroll
    for(int d = 8; d > 0; d >>= 1) {
#pragma unroll
        for(int i = 0; i < 16; ++i) {
            if((i & d) == 0) {
                const uint32_t lo = min(v[i], v[i + d]);
                v[i + d] = max(v[i], v[i + d]);
                v[i] = lo;
            }
        }
    }
}

// exact bottom-sp of one window by the whole warp (any content): every k-mer re-hashed from
// the packed words, 32 at a time, merged with the chunk network of sketch_dev.cuh
__device__ __noinline__ void tw_slow_window(const uint2* pk2, int pos, int nk, int sp, int k, int s, uint32_t* out,
                                            uint32_t* out_cnt, int lane, uint32_t* tmp)
{
    const int fsh = 32 - 2 * k;
    uint32_t kept = NOF;
    bool first = true;
    for(int c0 = 0; c0 < nk; c0 += 32) {
        const int i = c0 + lane;
        uint32_t h = NOF;
        if(i < nk) {
            const int b = pos + i, j = b >> 4, r = b & 15;
            const uint2 w0 = pk2[j], w1 = pk2[j + 1];
            const uint32_t f = __funnelshift_l(w1.x, w0.x, 2 * r) >> fsh;
            const uint32_t c = rc16(f) >> fsh;
            const uint32_t a = (w0.y << 16) | w1.y;
            if(((a << r) >> (32 - k)) == 0u) h = tmh(min(f, c));
        }
        if(__ballot_sync(0xFFFFFFFFu, h != NOF) == 0u) continue;
        kept = merge_chunk(kept, h, first, sp > 16, lane, tmp);
        first = false;
    }
    int nout = __popc(__ballot_sync(0xFFFFFFFFu, kept != NOF));
    if(nout > sp) nout = sp;
    if(lane < s) out[lane] = (lane < sp) ? kept : NOF;
    if(out_cnt != nullptr && lane == 0) *out_cnt = uint32_t(nout);
}

template<bool K16>
__global__ void __launch_bounds__(TW_THREADS, 3)
sketch_tpw_kernel(const SketchParams p)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int ngc = p.ngroups;
    const int rawcap = (ngc + 2) * 16;
    uint8_t*  base = smem_raw + size_t(warp) * p.warp_smem;
    uint64_t* bar  = reinterpret_cast<uint64_t*>(base);
    uint8_t*  raw  = base + 16;
    uint2*    pk2  = reinterpret_cast<uint2*>(raw + rawcap);                 // [ngc + 2] {packed, ambiguity}
    uint32_t* list = reinterpret_cast<uint32_t*>(pk2 + ngc + 2);             // [TW_LCAP][32]
    uint32_t* tmp  = list + TW_LCAP * 32;                                    // [32]
    int32_t*  tabs = reinterpret_cast<int32_t*>(tmp + 32);                   // 2 x TW_TABSZ
    constexpr int wpt = TW_WPT;
    const uint64_t ntiles = (p.nwin - p.win_lo + wpt - 1) / wpt;
    const uint64_t nwarps = uint64_t(gridDim.x) * TW_WARPS;
    const uint64_t gw = uint64_t(blockIdx.x) * TW_WARPS + warp;

    // resolve tile `tile` into table `stage` and start the TMA copy of its bytes into raw
    auto prefetch = [&](uint64_t tile, int stage) {
        int32_t* tb = tabs + stage * TW_TABSZ;
        uint32_t* tbu = reinterpret_cast<uint32_t*>(tb);
        const uint64_t g0 = p.win_lo + tile * uint64_t(wpt);
        const int nw = int(min(uint64_t(wpt), p.nwin - g0));
        int64_t first_abs = 0, last_end = 0;
        const uint32_t q0 = p.cta_seq[tile];
        const uint32_t wf0 = p.win_first[q0], wf1 = p.win_first[q0 + 1];
        if(g0 + uint64_t(nw) <= uint64_t(wf1)) {
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
                tbu[3 * wpt + lane] = (nk == p.nk_full) ? p.thr_full : tw_threshold(nk, min(p.s, max(nk, 0)));
            }
        } else {
            int64_t abs_i = 0, end_i = 0;
            if(lane < nw) {
                const uint32_t g = uint32_t(g0 + lane);
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
                tb[wpt + lane] = nk;
                tbu[2 * wpt + lane] = p.out_first[q] + uint32_t(w);
                tbu[3 * wpt + lane] = tw_threshold(nk, min(p.s, max(nk, 0)));
            }
            first_abs = __shfl_sync(0xFFFFFFFFu, abs_i, 0);
            const int sb_lo = int((reinterpret_cast<uintptr_t>(p.seq) + uint64_t(first_abs)) & 15u);
            if(lane < nw) tb[lane] = int32_t(abs_i - first_abs) + sb_lo;
            last_end = __shfl_sync(0xFFFFFFFFu, end_i, nw - 1);
        }
        const int64_t sb = first_abs - int64_t((reinterpret_cast<uintptr_t>(p.seq) + uint64_t(first_abs)) & 15u);
        const int span_len = int(last_end - sb);
        if(lane == 0) {
            tb[4 * wpt] = nw;
            tb[4 * wpt + 1] = span_len;
            const uint32_t bytes = uint32_t(((span_len + 15) >> 4) + 1) * 16u;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(bar, bytes);
            tma_load_1d(raw, p.seq + sb, bytes, bar);
        }
        __syncwarp();
    };

    if(lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if(gw < ntiles) prefetch(gw, 0);

    uint32_t* const lbase = list + lane;
    const int k = p.k;
    const int fsh = 32 - 2 * k;

    int it = 0;
    for(uint64_t tile = gw; tile < ntiles; tile += nwarps, ++it) {
        const int stage = it & 1;
        const int32_t* tb = tabs + stage * TW_TABSZ;
        const int nw = tb[4 * wpt];
        const int ng = (tb[4 * wpt + 1] + 15) >> 4;
        mbar_wait(bar, uint32_t(it & 1));

        // ---- (1) pack: raw ASCII -> {2-bit word, ambiguity bits} per 16-base group
        const uint4* rawv = reinterpret_cast<const uint4*>(raw);
        for(int j = lane; j < ng + 1; j += 32) {
            const uint4 x = rawv[j];
            uint32_t p0, p1, p2, p3, d0, d1, d2, d3;
            tw_pack4(x.x, p0, d0);
            tw_pack4(x.y, p1, d1);
            tw_pack4(x.z, p2, d2);
            tw_pack4(x.w, p3, d3);
            const uint32_t hi = __byte_perm(p0, p1, 0x0037);        // [.. .. p0.b3 p1.b3]
            const uint32_t lo = __byte_perm(p2, p3, 0x0037);        // [.. .. p2.b3 p3.b3]
            uint32_t w = __byte_perm(lo, hi, 0x5410);
            w ^= (w >> 1) & 0x55555555u;                            // A0 C1 T2 G3 -> A0 C1 G2 T3
            uint32_t a = 0;
            if((d0 | d1 | d2 | d3) != 0u)
                a = (nzmask4(d0) << 12) | (nzmask4(d1) << 8) | (nzmask4(d2) << 4) | nzmask4(d3);
            pk2[j] = make_uint2(w, a);
        }
        __syncwarp();
        // raw is consumed: the copy of the next tile overlaps everything below
        if(tile + nwarps < ntiles) prefetch(tile + nwarps, stage ^ 1);

        // ---- (2) this lane's window: hash + threshold filter into the lane's column
        const bool have = lane < nw;
        const int pos = have ? tb[lane] : 0;
        const int nk = have ? tb[wpt + lane] : 0;
        const uint32_t oidx = reinterpret_cast<const uint32_t*>(tb)[2 * wpt + lane];
        uint32_t T = reinterpret_cast<const uint32_t*>(tb)[3 * wpt + lane];
        const int sp = min(p.s, max(nk, 0));
        const bool thr = T != 0u;
#pragma unroll
        for(int i = 0; i < TW_SEL; ++i) lbase[i * 32] = NOH;
        uint32_t* lp = lbase;
        uint32_t* const lguard = lbase + (TW_LCAP - 16) * 32;
        uint32_t* const lend = lbase + TW_LCAP * 32;
        bool ovf = false;
        int done = 0;
        if(thr && nk >= 16) {
            const int j0 = pos >> 4, sh = 2 * (pos & 15), sa = pos & 15;
            uint2 w0 = pk2[j0], w1 = pk2[j0 + 1];
            uint32_t A0 = __funnelshift_l(w1.x, w0.x, sh);
            uint32_t B0 = (((w0.y << 16) | w1.y) << sa) >> 16;
            uint32_t R0 = rc16(A0);
            const int nfull = nk >> 4;
            for(int g = 0; g < nfull; ++g) {
                const uint2 w2 = pk2[j0 + g + 2];
                const uint32_t A1 = __funnelshift_l(w2.x, w1.x, sh);
                const uint32_t B1 = (((w1.y << 16) | w2.y) << sa) >> 16;
                const uint32_t R1 = rc16(A1);
                if(lp > lguard) { ovf = true; T = 0u; }
                const uint32_t a32 = (B0 << 16) | B1;
                if(a32 == 0u) tw_group<K16, false>(A0, A1, R0, R1, a32, k, T, lp);
                else          tw_group<K16, true>(A0, A1, R0, R1, a32, k, T, lp);
                A0 = A1; B0 = B1; R0 = R1; w1 = w2;
            }
            done = nfull << 4;
        }
        for(int i = done; i < nk; ++i) {
            const int b = pos + i, j = b >> 4, r = b & 15;
            const uint2 w0 = pk2[j], w1 = pk2[j + 1];
            const uint32_t f = __funnelshift_l(w1.x, w0.x, 2 * r) >> fsh;
            const uint32_t c = rc16(f) >> fsh;
            const uint32_t a = (w0.y << 16) | w1.y;
            uint32_t y = tmh_head(min(f, c));
            if(((a << r) >> (32 - k)) != 0u) y = NOH;
            const bool pass = thr ? (y < T) : (y != NOH);
            if(pass) {
                if(lp < lend) { *lp = y; lp += 32; }
                else ovf = true;
            }
        }
        const int cnt = int(lp - lbase) >> 5;

        // ---- (3) register network: 16 smallest of the <= 48 survivors
        uint32_t v[TW_SEL];
#pragma unroll
        for(int i = 0; i < TW_SEL; ++i) v[i] = tmh_tail(lbase[i * 32]);
#define SN_B 0
#include "sortnet16.inc"
#undef SN_B
#define SN_B 16
#include "sortnet16.inc"
#undef SN_B
#define SN_B 32
#include "sortnet16.inc"
#undef SN_B
        tw_merge_low<16>(v);
        tw_merge_low<32>(v);
        bool dup = false;
#pragma unroll
        for(int i = 0; i < 15; ++i) dup |= (v[i] == v[i + 1]);
        if(dup) {                                   // equal padding does not count
            dup = false;
#pragma unroll
            for(int i = 0; i < 15; ++i) dup |= (v[i] == v[i + 1]) && (v[i] != NOF);
        }
        const bool bad = (sp > 0) && (ovf || cnt > TW_SEL || dup || (thr && cnt < sp));
        if(have && !bad) {
            uint32_t* o = p.out_feat + uint64_t(oidx) * uint32_t(p.s);
            if(p.s == 16) {
                uint4* o4 = reinterpret_cast<uint4*>(o);
                o4[0] = make_uint4(v[0], v[1], v[2], v[3]);
                o4[1] = make_uint4(v[4], v[5], v[6], v[7]);
                o4[2] = make_uint4(v[8], v[9], v[10], v[11]);
                o4[3] = make_uint4(v[12], v[13], v[14], v[15]);
            } else {
#pragma unroll
                for(int i = 0; i < 16; ++i)
                    if(i < p.s) o[i] = v[i];
            }
            if(p.out_cnt != nullptr) {
                int n = 0;
#pragma unroll
                for(int i = 0; i < 16; ++i) n += (i < p.s && v[i] != NOF) ? 1 : 0;
                p.out_cnt[oidx] = uint32_t(min(n, sp));
            }
        }
        // ---- (4) rare: exact redo of flagged windows by the whole warp
        uint32_t bm = __ballot_sync(0xFFFFFFFFu, have && bad);
        while(bm != 0u) {
            const int src = __ffs(bm) - 1;
            bm &= bm - 1u;
            const int pos_s = __shfl_sync(0xFFFFFFFFu, pos, src);
            const int nk_s = __shfl_sync(0xFFFFFFFFu, nk, src);
            const uint32_t o_s = __shfl_sync(0xFFFFFFFFu, oidx, src);
            tw_slow_window(pk2, pos_s, nk_s, min(p.s, max(nk_s, 0)), k, p.s, p.out_feat + uint64_t(o_s) * uint32_t(p.s),
                           p.out_cnt ? p.out_cnt + o_s : nullptr, lane, tmp);
        }
        __syncwarp();
    }
}
#undef SN_CE

bool tpw_eligible(int sketchlen, int64_t winlen, int64_t stride)
{
    static int mode = -1;
    if(mode < 0) { const char* e = getenv("MCG_SK_IMPL"); mode = (e && strcmp(e, "warp") == 0) ? 0 : 1; }
    const int64_t adv = winlen > stride ? winlen : stride;
    return mode == 1 && sketchlen <= 16 && adv <= 512;
}

// shared-memory geometry of the thread-per-window kernel
void tpw_config(int64_t winlen, int64_t stride, int* ngroups, size_t* warp_smem)
{
    const int64_t adv = winlen > stride ? winlen : stride;
    const int ng = int((TW_WPT * adv + 30 + 15) / 16) + 1;      // worst case: every window in another sequence
    size_t bytes = 16 + size_t(ng + 2) * 16 + size_t(ng + 2) * 8 + size_t(TW_LCAP) * 32 * 4 + 32 * 4 + 2 * TW_TABSZ * 4;
    bytes = (bytes + 127) & ~size_t(127);
    *ngroups = ng;
    *warp_smem = bytes;
}

int tpw_launch(mcg_ctx* c, SketchParams& p, const mcg_sketch_cfg& cfg, uint64_t nblocks, cudaStream_t stream)
{
    const size_t smem = size_t(p.warp_smem) * TW_WARPS;
    MCG_REQUIRE(smem <= 220 * 1024, "window too long for the device sketcher");
    {
        const int nk = int(p.winlen) - cfg.k + 1;
        const int sp = std::min(p.s, std::max(nk, 0));
        p.nk_full = nk;
        p.thr_full = tw_threshold(nk, sp);
    }
#define TW_LAUNCH(K16)                                                                                              \
    do {                                                                                                            \
        auto kern = sketch_tpw_kernel<K16>;                                                                         \
        if(smem > 48 * 1024) MCG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem))); \
        int per_sm = 1;                                                                                             \
        MCG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TW_THREADS, smem));                   \
        if(per_sm < 1) per_sm = 1;                                                                                  \
        const uint64_t grid = std::min<uint64_t>((nblocks + TW_WARPS - 1) / TW_WARPS, uint64_t(kNumSMs) * uint64_t(per_sm)); \
        kern<<<unsigned(grid), TW_THREADS, smem, stream>>>(p);                                                      \
    } while(0)
    if(cfg.k == 16) TW_LAUNCH(true);
    else            TW_LAUNCH(false);
#undef TW_LAUNCH
    c->launches++;
    MCG_CUDA(cudaGetLastError());
    return MCG_OK;
}

}  // namespace mcg
