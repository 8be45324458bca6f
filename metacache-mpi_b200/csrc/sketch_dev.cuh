// sketch_dev.cuh -- device helpers shared by the sketch kernels (sketch.cu, sketch_tpw.cu)
#pragma once
#include "common.cuh"

namespace mcg {

constexpr int SK_THREADS = 256;
constexpr int SK_WARPS = SK_THREADS / 32;
constexpr int SK_CBUF = 128;              // per-warp compaction buffer (hashes)
constexpr uint32_t NOF = 0xFFFFFFFFu;
constexpr int SK_PITCH_DEFAULT = 74;        // pitch of the default shape (k=16, winlen=128, stride=113)

struct SketchParams {
    const uint8_t* seq;
    int64_t total_bytes;
    const int64_t* seq_off;
    int64_t nseq;
    const uint32_t* win_first;   // [nseq+1] window prefix in processing order
    const uint32_t* out_first;   // [nseq]   output index of each sequence's first window
    const uint32_t* cta_seq;     // [ntiles+1] sequence holding the first window of each tile
    uint64_t nwin;               // end of the window range of this launch (global window index) -- or an upper bound of it
    const uint32_t* nwin_dev;    // when non-null: the exact end on the device (callers that only know a bound on the host)
    uint64_t win_lo;             // first window of this launch
    uint32_t* out_feat;
    uint32_t* out_cnt;
    int k, s;
    int64_t winlen, stride;
    int wpc;                     // windows per warp tile
    int nk_full;                 // k-mer starts of a full window and its precomputed threshold
    uint32_t thr_full;
    int thr_extra;               // thread-per-window kernel: expected filter survivors = sketchlen + thr_extra
    int warp_smem;               // bytes of shared memory per warp
    int ngroups;                 // 16-base groups of span capacity
    int pitch;                   // row pitch of the transposed hash array (words)
    uint32_t* tile_ctr;          // dynamic tile counter of the thread-per-window kernel (zeroed by cta_seq_kernel)
    uint32_t c65536;             // the constant 65536, opaque to the compiler (keeps x >> 16 as a multiply-high)
};

__device__ __forceinline__ uint64_t sk_nwin(const SketchParams& p)
{
    return p.nwin_dev ? min(p.nwin, uint64_t(__ldg(p.nwin_dev))) : p.nwin;
}

// The hash array holds the hash BEFORE its last xor-shift: y = ((x1>>16)^x1)*C, h = y ^ (y>>16).
// h and y share their upper 16 bits, so "h < T" == "y < T" for thresholds that are multiples of
// 65536; the last step is applied to the few survivors only.  h == 0xFFFFFFFF <=> y == NOH.
constexpr uint32_t NOH = 0xFFFF0000u;

__host__ __device__ __forceinline__ uint32_t tmh_head(uint32_t x)
{
    x = ((x >> 16) ^ x) * 0x45d9f3bu;
    x = ((x >> 16) ^ x) * 0x45d9f3bu;
    return x;
}
__device__ __forceinline__ uint32_t tmh_tail(uint32_t y) { return y ^ (y >> 16); }

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

// ascending 32-lane bitonic sort, "flip" formulation: the first step of every merge pairs lane
// with its mirror inside the block (xor k-1), later steps pair lane ^ j; every comparator is
// ascending, so the direction of a step depends on ONE lane bit (5 loop-invariant predicates)
__device__ __forceinline__ uint32_t sort32(uint32_t v, int lane)
{
#pragma unroll
    for(int k = 2; k <= 32; k <<= 1) {
        {
            const uint32_t o = __shfl_xor_sync(0xFFFFFFFFu, v, k - 1);
            v = (lane & (k >> 1)) ? max(v, o) : min(v, o);
        }
#pragma unroll
        for(int j = k >> 2; j > 0; j >>= 1) {
            const uint32_t o = __shfl_xor_sync(0xFFFFFFFFu, v, j);
            v = (lane & j) ? max(v, o) : min(v, o);
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

// ---- mbarrier + TMA (cp.async.bulk) helpers ---------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared through the TMA unit; completion is signalled on the mbarrier
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

}  // namespace mcg
