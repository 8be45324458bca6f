// sketch_tpw.cu -- K1, thread-per-window formulation (sketchlen <= 16, windows <= 512 bases).
// Same contract as sketch.cu (bit-exact with dna_encoding.h:261-348,114-121,189-197,
// hash_int.h:40-45, hash_dna.h:115-152).
//
// A warp owns tiles of 32 consecutive windows, one window per lane.  Tiles are handed out by
// an atomic counter (warps the scheduler favours take more tiles).  Per tile:
//   (1) the tile's contiguous byte span arrives through TMA (cp.async.bulk + mbarrier) and is
//       packed by the whole warp to 2 bits/base + 1 ambiguity bit/base (shared memory);
//   (2) every lane walks the k-mers of ITS window: 16 k-mers per step out of two aligned
//       packed words (compile-time funnel shifts), canonical k-mer, hash, and keeps the hashes
//       below a threshold with ~sketchlen+15 expected survivors in its own shared-memory column
//       (slot i of lane l lives at word i*32+l: conflict-free, no ballots, no shuffles; the
//       columns reuse the raw-byte buffer, which is dead after packing);
//   (3) every lane reduces its column in REGISTERS, 16 survivors at a time: a 60-comparator
//       16-input sorting network, then "16 smallest of two sorted 16s" (one min layer + a
//       bitonic merge).  The max of each compare-exchange is a + b - min(a, b) as two IMADs:
//       the FMA pipe idles while the ALU pipe (shifts, logic, min, compares) is the limit;
//   (4) the warp's next tile was resolved in three steps spread over (1)-(3) so that no
//       global-memory latency is waited for; its TMA copy starts as soon as the columns are
//       consumed and overlaps the stores and the other warps.
// A window whose survivors are too few / too many / not distinct is redone exactly by the whole
// warp (rare).  ~105 warp instructions per 128-base window (warp-per-window kernel: ~325).
#include "sketch_dev.cuh"

#include <algorithm>

namespace mcg {

constexpr int TW_THREADS = 128;
constexpr int TW_WARPS = TW_THREADS / 32;
constexpr int TW_WPT = 32;               // windows per warp tile (one per lane)
constexpr int TW_LCAP = 64;              // candidate slots per window in shared memory 
constexpr int TW_SEL = 48;               // candidates the register network takes

// threshold on the hash head (multiple of 65536, see NOH) with sp+15 expected survivors among
// nk k-mers; 0 = keep everything (short windows)
__host__ __device__ __forceinline__ uint32_t tw_threshold(int nk, int sp, int extra)
{
    if(nk <= TW_SEL) return 0u;
    const float pr = float(sp + extra) / float(nk);
    return pr < 0.999f ? uint32_t(pr * 65536.0f) << 16 : 0u;
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

// compare-exchange.  FMA form: hi = a + b - min(a, b) as two multiply-adds by registers holding
// 1 and -1 (IMAD, FMA pipe) instead of a second min/max on the ALU pipe that limits this kernel
template<bool FMA>
__device__ __forceinline__ void tw_ce(uint32_t& a, uint32_t& b, uint32_t one, uint32_t m1)
{
    const uint32_t lo = min(a, b);
    if(FMA) {
        uint32_t s, hi;
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(s) : "r"(a), "r"(one), "r"(b));
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(hi) : "r"(lo), "r"(m1), "r"(s));
        b = hi;
    } else {
        b = max(a, b);
    }
    a = lo;
}
#define SN_CE(i, j) tw_ce<CEF>(v[SN_B + (i)], v[SN_B + (j)], one, m1);

__device__ __forceinline__ void st_shared_u32(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// 16 k-mers of one lane's window: starts 0..15 of the aligned word pair (A0,A1); R = rc16(A);
// a32 = ambiguity bits of the 32 bases (CHECK).  Survivors of the threshold go to the lane's
// column at shared address la.
template<bool K16, bool CHECK>
__device__ __forceinline__ void tw_group_fast(uint32_t A0, uint32_t A1, uint32_t R0, uint32_t R1, uint32_t a32, int k, uint32_t T,
                                              uint32_t& la)
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
        if(CHECK) { if(((a32 << r) >> ash) != 0u) y = NOH; }     // NOH >= every threshold
        if(y < T) { st_shared_u32(la, y); la += 128u; }
    }
}

// the first nr (<= 16) k-mers of a group, any content: ambiguity check, threshold or keep-all
// mode, capacity check (rolled: rare and small)
template<bool K16>
__device__ __forceinline__ void tw_group_any(uint32_t A0, uint32_t A1, uint32_t R0, uint32_t R1, uint32_t a32, int nr, int k,
                                             bool thr, uint32_t T, uint32_t& la, uint32_t lend, bool& ovf)
{
    const int fsh = 32 - 2 * k;
    const uint32_t kmask = 0xFFFFFFFFu >> fsh;
    const int ash = 32 - k;
#pragma unroll 1
    for(int r = 0; r < nr; ++r) {
        uint32_t f = __funnelshift_l(A1, A0, 2 * r);
        uint32_t c = __funnelshift_r(R0, R1, 2 * r);
        if(!K16) { f >>= fsh; c &= kmask; }
        uint32_t y = tmh_head(min(f, c));
        if(((a32 << r) >> ash) != 0u) y = NOH;
        const bool pass = thr ? (y < T) : (y != NOH);
        if(pass) {
            if(la < lend) { st_shared_u32(la, y); la += 128u; }
            else ovf = true;
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

// what a lane knows about one of its windows
struct TwWin {
    int pos, nk;            // first base of the window inside the packed span, k-mer count
    uint32_t oidx, T;       // output row, threshold
    bool have;              // the window exists (tiles at the end of a launch are short)
};
// a tile = 32 * R consecutive windows; lane l owns windows R*l .. R*l + R-1 and handles them in R
// rounds.  R = 2 is for read-like input (long and short windows alternate: interleaving gives
// every round windows of one kind); + warp-uniform span data
template<int R>
struct TwTile {
    TwWin w[R];
    int nw, span_len;       // uniform: windows in the tile, bytes from the aligned span start
    int64_t sb;             // uniform: 16-byte aligned offset of the span in the sequence buffer
};
struct TwSeq { uint32_t q0, wf0, wf1, of; int64_t so, so1; };

__device__ __forceinline__ TwSeq tw_load_seq(const SketchParams& p, uint32_t q0)
{
    TwSeq q;
    q.q0 = q0;
    q.wf0 = __ldg(p.win_first + q0);
    q.wf1 = __ldg(p.win_first + q0 + 1);
    q.of = __ldg(p.out_first + q0);
    q.so = __ldg(p.seq_off + q0);
    q.so1 = __ldg(p.seq_off + q0 + 1);
    return q;
}

template<int R>
__device__ __forceinline__ TwTile<R> tw_resolve(const SketchParams& p, uint64_t tile, const TwSeq& q, int lane)
{
    TwTile<R> t;
    constexpr int WPT = TW_WPT * R;
    const uint64_t g0 = p.win_lo + tile * uint64_t(WPT);
    const int nw = int(min(uint64_t(WPT), sk_nwin(p) - g0));
    const int winlen = int(p.winlen), stride = int(p.stride);
    const int last_lane = (nw - 1) / R, last_round = (nw - 1) % R;
    int64_t first_abs, last_end;
    if(g0 + uint64_t(nw) <= uint64_t(q.wf1)) {
        // common case: the whole tile lies in one sequence
        const int64_t len = q.so1 - q.so;
        const uint32_t w0 = uint32_t(g0 - q.wf0);
        const bool multi = len > winlen;
        const int64_t b0 = multi ? int64_t(w0) * stride : 0;
        const int rem0 = int(min(len - b0, int64_t(1) << 30));       // bases from the tile's first window to the end
        int wl_sel = 0;
#pragma unroll
        for(int r = 0; r < R; ++r) {
            const int j = R * lane + r;
            int off = 0, wl = int(min(len, int64_t(winlen)));
            if(multi) { off = j * stride; wl = min(winlen, rem0 - off); }
            t.w[r].pos = off;
            t.w[r].nk = wl - p.k + 1;
            t.w[r].oidx = q.of + w0 + uint32_t(j);
            if(r == last_round) wl_sel = wl;
        }
        first_abs = q.so + b0;
        last_end = first_abs + (multi ? (nw - 1) * stride : 0) + __shfl_sync(0xFFFFFFFFu, wl_sel, last_lane);
    } else {
        int64_t abs0 = 0, end_sel = 0;
#pragma unroll
        for(int r = 0; r < R; ++r) {
            const int j = R * lane + r;
            int64_t abs_i = 0;
            int wl = 0;
            t.w[r].nk = 0;
            t.w[r].oidx = 0;
            if(j < nw) {
                const uint32_t g = uint32_t(g0) + uint32_t(j);
                // every sequence has >= 1 window: the sequence of window g0+j is in [q0, q0+j]
                uint32_t lo = q.q0, hi = uint32_t(min(int64_t(q.q0) + j + 1, p.nseq));
                while(hi - lo > 1) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if(__ldg(p.win_first + mid) <= g) lo = mid; else hi = mid;
                }
                const uint32_t w = g - __ldg(p.win_first + lo);
                const int64_t so = __ldg(p.seq_off + lo);
                const int64_t len = __ldg(p.seq_off + lo + 1) - so;
                int64_t b = 0, e = len;
                if(len > winlen) { b = int64_t(w) * stride; e = min(b + winlen, len); }
                abs_i = so + b;
                wl = int(e - b);
                t.w[r].nk = wl - p.k + 1;
                t.w[r].oidx = __ldg(p.out_first + lo) + w;
            }
            if(r == 0) abs0 = abs_i;
            if(r == last_round) end_sel = abs_i + wl;
            if(r == 0) first_abs = __shfl_sync(0xFFFFFFFFu, abs0, 0);
            t.w[r].pos = int(abs_i - first_abs);
        }
        last_end = __shfl_sync(0xFFFFFFFFu, end_sel, last_lane);
    }
    const int sb_lo = int((reinterpret_cast<uintptr_t>(p.seq) + uint64_t(first_abs)) & 15u);
    t.sb = first_abs - sb_lo;
    t.span_len = int(last_end - t.sb);
    t.nw = nw;
#pragma unroll
    for(int r = 0; r < R; ++r) {
        t.w[r].have = R * lane + r < nw;
        t.w[r].pos += sb_lo;
        if(!t.w[r].have) t.w[r].nk = 0;
        t.w[r].T = (t.w[r].nk == p.nk_full) ? p.thr_full : tw_threshold(t.w[r].nk, min(p.s, max(t.w[r].nk, 0)), p.thr_extra);
    }
    return t;
}

template<bool K16, bool CEF, int R>
__global__ void __launch_bounds__(TW_THREADS, R == 1 ? 5 : 4)
sketch_tpw_kernel(const SketchParams p)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int ngc = p.ngroups;
    const int rawcap = max((ngc + 2) * 16, TW_LCAP * 128);
    uint8_t*  base = smem_raw + size_t(warp) * p.warp_smem;
    uint64_t* bar  = reinterpret_cast<uint64_t*>(base);
    uint8_t*  raw  = base + 16;                                              // TMA destination ...
    uint32_t* list = reinterpret_cast<uint32_t*>(raw);                       // ... reused as [TW_LCAP][32] candidate columns
    uint2*    pk2  = reinterpret_cast<uint2*>(raw + rawcap);                 // [ngc + 2] {packed, ambiguity}
    uint32_t* tmp  = reinterpret_cast<uint32_t*>(pk2 + ngc + 2);             // [32]
    const uint64_t nwin_end = sk_nwin(p);
    const uint32_t ntiles = nwin_end > p.win_lo ? uint32_t((nwin_end - p.win_lo + TW_WPT * R - 1) / (TW_WPT * R)) : 0u;
    const int k = p.k;
    const uint32_t one = p.c65536 >> 16, m1 = 0u - one;      // 1 and -1, opaque to the compiler

    // starts the TMA copy of a tile's bytes (the candidate columns of the previous tile must be consumed)
    auto tma_issue = [&](const TwTile<R>& t) {
        if(lane == 0) {
            const uint32_t bytes = uint32_t(((t.span_len + 15) >> 4) + 1) * 16u;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(bar, bytes);
            tma_load_1d(raw, p.seq + t.sb, bytes, bar);
        }
    };

    if(lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    // tiles are handed out by an atomic counter: warps that the scheduler favours simply take
    // more tiles, and every warp stays busy until the last tile is gone
    auto next_tile = [&]() -> uint32_t {
        uint32_t t = 0;
        if(lane == 0) t = atomicAdd(p.tile_ctr, 1u);
        return __shfl_sync(0xFFFFFFFFu, t, 0);
    };
    uint32_t tile = next_tile();
    if(tile >= ntiles) return;
    TwTile<R> cur = tw_resolve<R>(p, tile, tw_load_seq(p, __ldg(p.cta_seq + tile)), lane);
    tma_issue(cur);

    const uint32_t la0 = smem_u32(list) + uint32_t(lane) * 4u;
    const uint32_t lguard = la0 + uint32_t(TW_LCAP - 16) * 128u;
    const uint32_t lend = la0 + uint32_t(TW_LCAP) * 128u;

    for(int it = 0;; ++it) {
        const uint32_t nt = next_tile();
        const bool more = nt < ntiles;
        // the descriptor of the warp's next tile is resolved in three steps spread over this
        // tile's work, so that no global-memory latency is ever waited for
        const uint32_t q0n = more ? __ldg(p.cta_seq + nt) : 0u;             // step A
        const int ng = (cur.span_len + 15) >> 4;
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
        TwSeq sn;
        if(more) sn = tw_load_seq(p, q0n);                                   // step B

        TwTile<R> nxt = cur;
        TwWin d = cur.w[0];
#pragma unroll 1
        for(int rd = 0; rd < R; ++rd) {
        // ---- (2) this lane's window of the round: hash + threshold filter into the lane's column
        const int nk = d.nk;
        const int sp = min(p.s, max(nk, 0));
        uint32_t T = d.T;
        const bool thr = T != 0u;
        uint32_t la = la0;
        bool ovf = false;
        if(nk > 0) {
            const int j0 = d.pos >> 4, sa = d.pos & 15, sh = 2 * sa;
            const uint2 w0 = pk2[j0];
            uint2 w1 = pk2[j0 + 1];
            uint32_t A0 = __funnelshift_l(w1.x, w0.x, sh);
            uint32_t B0 = (((w0.y << 16) | w1.y) << sa) >> 16;
            uint32_t R0 = rc16(A0);
            const uint2* pw = pk2 + j0 + 2;
            const int nfull = thr ? (nk >> 4) : 0;              // full groups of the threshold mode
            int left = nk;
            for(int g = 0; g < nfull; ++g) {
                const uint2 w2 = *pw++;
                const uint32_t A1 = __funnelshift_l(w2.x, w1.x, sh);
                const uint32_t B1 = (((w1.y << 16) | w2.y) << sa) >> 16;
                const uint32_t R1 = rc16(A1);
                const uint32_t a32 = (B0 << 16) | B1;
                if(la > lguard) { ovf = true; T = 0u; }
                if(a32 == 0u) tw_group_fast<K16, false>(A0, A1, R0, R1, a32, k, T, la);
                else          tw_group_fast<K16, true>(A0, A1, R0, R1, a32, k, T, la);
                A0 = A1; B0 = B1; R0 = R1; w1 = w2;
            }
            left -= nfull << 4;
            while(left > 0) {                                   // last partial group / keep-everything mode
                const uint2 w2 = *pw++;
                const uint32_t A1 = __funnelshift_l(w2.x, w1.x, sh);
                const uint32_t B1 = (((w1.y << 16) | w2.y) << sa) >> 16;
                const uint32_t R1 = rc16(A1);
                tw_group_any<K16>(A0, A1, R0, R1, (B0 << 16) | B1, min(16, left), k, thr, T, la, lend, ovf);
                A0 = A1; B0 = B1; R0 = R1; w1 = w2;
                left -= 16;
            }
        }
        const int cnt = int(la - la0) >> 7;
        // pad the column with the sentinel up to the next multiple of 16 (the sorter reads blocks of 16)
        if(la <= lguard) {
#pragma unroll
            for(int i = 0; i < 16; ++i) st_shared_u32(la + uint32_t(i) * 128u, NOH);
        } else {
            for(uint32_t a = la; a < lend; a += 128u) st_shared_u32(a, NOH);
        }

        // ---- (3) register network, 16 survivors at a time: 16-input sort, then the 16 smallest
        //      of (best, block) by one min layer + a bitonic merge
        uint32_t best[16];
        {
            uint32_t (&v)[16] = best;
            const uint32_t* src = list + lane;
#pragma unroll
            for(int i = 0; i < 16; ++i) v[i] = tmh_tail(src[i * 32]);
#define SN_B 0
#include "sortnet16.inc"
#undef SN_B
        }
        const int nb = (cnt + 15) >> 4;
#pragma unroll 1
        for(int b = 1; b < nb; ++b) {
            uint32_t v[16];
            const uint32_t* src = list + lane + b * 512;
#pragma unroll
            for(int i = 0; i < 16; ++i) v[i] = tmh_tail(src[i * 32]);
#define SN_B 0
#include "sortnet16.inc"
#undef SN_B
#pragma unroll
            for(int i = 0; i < 16; ++i) best[i] = min(best[i], v[15 - i]);
#pragma unroll
            for(int d = 8; d > 0; d >>= 1) {
#pragma unroll
                for(int i = 0; i < 16; ++i) {
                    if((i & d) == 0) tw_ce<CEF>(best[i], best[i + d], one, m1);
                }
            }
        }
        __syncwarp();
        // the columns of the last round are consumed: resolve the next tile and start its copy   (step C)
        if(rd == R - 1 && more) {
            nxt = tw_resolve<R>(p, nt, sn, lane);
            tma_issue(nxt);
        }

        bool dup = false;
#pragma unroll
        for(int i = 0; i < 15; ++i) dup |= (best[i] == best[i + 1]);
        if(dup) {                                   // equal padding does not count
            dup = false;
#pragma unroll
            for(int i = 0; i < 15; ++i) dup |= (best[i] == best[i + 1]) && (best[i] != NOF);
        }
        const bool have = d.have;
        const bool bad = (sp > 0) && (ovf || dup || (thr && cnt < sp));
        if(have && !bad) {
            uint32_t* o = p.out_feat + uint64_t(d.oidx) * uint32_t(p.s);
            if(p.s == 16) {
                uint4* o4 = reinterpret_cast<uint4*>(o);
                o4[0] = make_uint4(best[0], best[1], best[2], best[3]);
                o4[1] = make_uint4(best[4], best[5], best[6], best[7]);
                o4[2] = make_uint4(best[8], best[9], best[10], best[11]);
                o4[3] = make_uint4(best[12], best[13], best[14], best[15]);
            } else {
#pragma unroll
                for(int i = 0; i < 16; ++i)
                    if(i < p.s) o[i] = best[i];
            }
            if(p.out_cnt != nullptr) {
                int n = 0;
#pragma unroll
                for(int i = 0; i < 16; ++i) n += (i < p.s && best[i] != NOF) ? 1 : 0;
                p.out_cnt[d.oidx] = uint32_t(min(n, sp));
            }
        }
        // ---- (4) rare: exact redo of flagged windows by the whole warp
        uint32_t bm = __ballot_sync(0xFFFFFFFFu, have && bad);
        while(bm != 0u) {
            const int src = __ffs(bm) - 1;
            bm &= bm - 1u;
            const int pos_s = __shfl_sync(0xFFFFFFFFu, d.pos, src);
            const int nk_s = __shfl_sync(0xFFFFFFFFu, nk, src);
            const uint32_t o_s = __shfl_sync(0xFFFFFFFFu, d.oidx, src);
            tw_slow_window(pk2, pos_s, nk_s, min(p.s, max(nk_s, 0)), k, p.s, p.out_feat + uint64_t(o_s) * uint32_t(p.s),
                           p.out_cnt ? p.out_cnt + o_s : nullptr, lane, tmp);
        }
        __syncwarp();
        if(R > 1) d = cur.w[R - 1];                 // R <= 2
        }   // rounds
        if(!more) break;
        cur = nxt;
        tile = nt;
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
void tpw_config(int64_t winlen, int64_t stride, int rounds, int* ngroups, size_t* warp_smem)
{
    const int64_t adv = winlen > stride ? winlen : stride;
    const int ng = int((TW_WPT * rounds * adv + 30 + 15) / 16) + 1;      // worst case: every window in another sequence
    const size_t region = std::max(size_t(ng + 2) * 16, size_t(TW_LCAP) * 32 * 4);     // raw bytes, then candidate columns
    size_t bytes = 16 + region + size_t(ng + 2) * 8 + 32 * 4;
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
        // expected survivors = sketchlen + 15; MCG_SK_EXTRA (tests) moves the filter to force the rare paths
        static int extra = -1000;
        if(extra == -1000) { const char* e = getenv("MCG_SK_EXTRA"); extra = e ? atoi(e) : 15; }
        p.thr_extra = extra;
        p.thr_full = tw_threshold(nk, sp, extra);
    }
    static int cef = -1;
    if(cef < 0) { const char* e = getenv("MCG_SK_CEF"); cef = e ? atoi(e) : 1; }
    p.c65536 = 65536u;
    const int rounds = p.wpc / TW_WPT;
#define TW_LAUNCH(K16, CEF, R)                                                                                      \
    do {                                                                                                            \
        auto kern = sketch_tpw_kernel<K16, CEF, R>;                                                                 \
        if(smem > 48 * 1024) MCG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem))); \
        int per_sm = 1;                                                                                             \
        MCG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TW_THREADS, smem));                   \
        if(per_sm < 1) per_sm = 1;                                                                                  \
        const uint64_t grid = std::min<uint64_t>((nblocks + TW_WARPS - 1) / TW_WARPS, uint64_t(c->num_sms) * uint64_t(per_sm)); \
        kern<<<unsigned(grid), TW_THREADS, smem, stream>>>(p);                                                      \
    } while(0)
    if(rounds == 2) {
        if(cfg.k != 16)   TW_LAUNCH(false, true, 2);
        else              TW_LAUNCH(true, true, 2);
    } else {
        if(cfg.k != 16)   TW_LAUNCH(false, true, 1);
        else if(cef == 0) TW_LAUNCH(true, false, 1);
        else              TW_LAUNCH(true, true, 1);
    }
#undef TW_LAUNCH
    c->launches++;
    MCG_CUDA(cudaGetLastError());
    return MCG_OK;
}

}  // namespace mcg
