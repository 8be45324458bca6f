// radix_sort.cu -- stable LSD radix sort of (uint32 key, uint32 value) pairs, 8 bits
// per pass.  Used by the database build (sort feature slots by feature; the
// stability keeps each feature's locations in ascending (target, window) order,
// which is what makes the reference's "first maxlocs win" capping reproducible:
// sketch_database.h:1089-1092, hash_multimap.h:706-713).
//
// Per pass: (1) per-tile digit histogram, (2) exclusive scan over
// [digit][tile], (3) tile-local stable counting sort in shared memory followed
// by coalesced run-wise stores.  HBM-bound: 20 B per element per pass.
#include "common.cuh"

namespace mcg {

constexpr int RS_ROUNDS = 16;                       // elements per thread

int exclusive_scan_u32_to_u32(mcg_ctx* c, const uint32_t* d_in, uint32_t* d_out, uint64_t n);

// digit histogram of every tile (order-free: 16 consecutive keys per thread, vector loads)
template<int THREADS>
__global__ void __launch_bounds__(THREADS)
rs_hist_kernel(const uint32_t* __restrict__ keys, uint64_t n, int shift, uint32_t* __restrict__ hist, uint32_t ntiles)
{
    constexpr int TILE = THREADS * RS_ROUNDS;
    __shared__ uint32_t h[256];
    if(threadIdx.x < 256) h[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t e0 = uint64_t(blockIdx.x) * TILE + uint64_t(threadIdx.x) * RS_ROUNDS;
    if(e0 + RS_ROUNDS <= n) {
        uint4 v[RS_ROUNDS / 4];
#pragma unroll
        for(int q = 0; q < RS_ROUNDS / 4; ++q) v[q] = __ldg(reinterpret_cast<const uint4*>(keys + e0) + q);
#pragma unroll
        for(int q = 0; q < RS_ROUNDS / 4; ++q) {
            atomicAdd(&h[(v[q].x >> shift) & 255u], 1u);
            atomicAdd(&h[(v[q].y >> shift) & 255u], 1u);
            atomicAdd(&h[(v[q].z >> shift) & 255u], 1u);
            atomicAdd(&h[(v[q].w >> shift) & 255u], 1u);
        }
    } else {
        for(int r = 0; r < RS_ROUNDS; ++r)
            if(e0 + r < n) atomicAdd(&h[(keys[e0 + r] >> shift) & 255u], 1u);
    }
    __syncthreads();
    if(threadIdx.x < 256) hist[uint64_t(threadIdx.x) * ntiles + blockIdx.x] = h[threadIdx.x];
}

// Ranking as in the "match" scheme of one-sweep sorts: a warp owns RS_ROUNDS x 32 CONSECUTIVE
// elements of the tile and ranks them round by round against its own 256 digit counters
// (peer masks from ballots, no block barrier inside the loop); one pass over [warp][digit] turns the
// counters into tile positions.  Stable: tile order = (warp, round, lane).  The larger the tile,
// the longer the runs of one digit it stores (tile / 256 elements on average): 8192 elements
// give 128-byte runs.
template<int THREADS>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS)
rs_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                  uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                  uint64_t n, int shift, const uint32_t* __restrict__ offsets, uint32_t ntiles, uint32_t iota)
{
    constexpr int TILE = THREADS * RS_ROUNDS, WARPS = THREADS / 32;
    extern __shared__ __align__(16) uint32_t rs_sm[];
    uint32_t* s_key = rs_sm;                        // [TILE]
    uint32_t* s_val = rs_sm + TILE;                 // [TILE]
    uint32_t* s_wcnt = rs_sm + 2 * TILE;            // [WARPS][256] per warp: running count, then exclusive prefix over warps
    __shared__ uint32_t s_start[256];               // tile-local start of each digit
    __shared__ uint32_t s_gbase[256];               // global base of each digit for this tile - s_start
    __shared__ uint32_t s_wsum[8];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t base = uint64_t(blockIdx.x) * TILE;
    const int tile_n = int(min(uint64_t(TILE), n - base));
    const int wbase = warp * (RS_ROUNDS * 32);
    const uint32_t lt = (1u << lane) - 1u;

    for(int i = tid; i < WARPS * 256; i += THREADS) s_wcnt[i] = 0;
    // the values of the tile travel to shared memory (the s_key region, in tile order) behind the
    // ranking: cp.async, no registers, no wait for HBM in the scatter phase
    if(!iota) {
        for(int ch = tid; ch < TILE / 4; ch += THREADS) {
            const int li = ch * 4;
            if(li + 4 <= tile_n) {
                const uint32_t dst = uint32_t(__cvta_generic_to_shared(s_key + li));
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(vals_in + base + li) : "memory");
            } else {
                for(int q = li; q < tile_n; ++q) s_key[q] = vals_in[base + q];
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    uint32_t mykey[RS_ROUNDS];
#pragma unroll
    for(int r = 0; r < RS_ROUNDS; ++r) {
        const int li = wbase + r * 32 + lane;
        mykey[r] = li < tile_n ? keys_in[base + li] : 0xFFFFFFFFu;
    }
    __syncthreads();
    uint32_t* wc = s_wcnt + warp * 256;
    uint32_t myrank[RS_ROUNDS / 2];                         // two 16-bit ranks per register
#pragma unroll
    for(int r = 0; r < RS_ROUNDS; ++r) {
        const uint32_t d = (mykey[r] >> shift) & 255u;      // padding elements rank as digit 255 behind all real ones
        // lanes with the same digit: 8 ballots (match.any occupies the ADU pipe for ~60 cycles per warp
        // instruction on sm_100 -- ncu: 65 % ADU utilisation was the limit of this kernel)
        uint32_t peers = 0xFFFFFFFFu;
#pragma unroll
        for(int b = 0; b < 8; ++b) {
            const bool bit = (d >> b) & 1u;
            const uint32_t bal = __ballot_sync(0xFFFFFFFFu, bit);
            peers &= bit ? bal : ~bal;
        }
        const uint32_t before = wc[d];
        __syncwarp();
        if((peers & lt) == 0) wc[d] = before + __popc(peers);
        __syncwarp();
        const uint32_t rk = before + __popc(peers & lt);
        if(r & 1) myrank[r >> 1] |= rk << 16; else myrank[r >> 1] = rk;
    }
    if(!iota) asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    uint32_t run = 0, inc = 0;
    if(tid < 256) {
        // thread tid owns digit tid: exclusive prefix over the warps, then over the digits
#pragma unroll
        for(int w = 0; w < WARPS; ++w) {
            const uint32_t cnt = s_wcnt[w * 256 + tid];
            s_wcnt[w * 256 + tid] = run;
            run += cnt;
        }
        inc = run;
#pragma unroll
        for(int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if(lane >= d) inc += o;
        }
        if(lane == 31) s_wsum[warp] = inc;
    }
    __syncthreads();
    if(tid < 256) {
        uint32_t wb = 0;
        for(int w = 0; w < warp; ++w) wb += s_wsum[w];
        const uint32_t start = wb + inc - run;
        s_start[tid] = start;
        s_gbase[tid] = offsets[uint64_t(tid) * ntiles + blockIdx.x] - start;
    }
    __syncthreads();
    // values first (staged copy in the s_key region -> tile position in s_val), then the keys
#pragma unroll
    for(int r = 0; r < RS_ROUNDS; ++r) {
        const int li = wbase + r * 32 + lane;
        const uint32_t d = (mykey[r] >> shift) & 255u;
        const uint32_t pos = s_start[d] + wc[d] + ((r & 1) ? (myrank[r >> 1] >> 16) : (myrank[r >> 1] & 0xFFFFu));
        if(r & 1) myrank[r >> 1] = (myrank[r >> 1] & 0xFFFFu) | (pos << 16); else myrank[r >> 1] = (myrank[r >> 1] & 0xFFFF0000u) | pos;
        if(li < tile_n) s_val[pos] = iota ? uint32_t(base + li) / iota : s_key[li];      // iota = divisor of the element index (0: values given)
    }
    __syncthreads();
#pragma unroll
    for(int r = 0; r < RS_ROUNDS; ++r) {
        const int li = wbase + r * 32 + lane;
        const uint32_t pos = (r & 1) ? (myrank[r >> 1] >> 16) : (myrank[r >> 1] & 0xFFFFu);
        if(li < tile_n) s_key[pos] = mykey[r];
    }
    __syncthreads();
    // coalesced stores: consecutive tile positions of one digit are consecutive globally
#pragma unroll 4
    for(int i = tid; i < tile_n; i += THREADS) {
        const uint32_t key = s_key[i];
        const uint32_t d = (key >> shift) & 255u;
        const uint64_t g = uint64_t(s_gbase[d] + uint32_t(i));
        keys_out[g] = key;
        vals_out[g] = s_val[i];
    }
}

__global__ void iota_kernel(uint32_t* v, uint64_t n, uint32_t div)
{
    const uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if(i < n) v[i] = uint32_t(i) / div;
}

int radix_sort_pairs(mcg_ctx* c, uint32_t* keys, uint32_t* vals, uint32_t* keys_alt, uint32_t* vals_alt,
                     uint64_t n, int key_bits, bool iota_vals, uint32_t iota_div)
{
    if(iota_div == 0) iota_div = 1;
    if(n == 0) return MCG_OK;
    MCG_REQUIRE(n < (1ull << 32), "radix_sort_pairs: more than 2^32-1 elements");
    static const int threads = [] { const char* e = getenv("MCG_RS_THREADS"); return (e && atoi(e) == 512) ? 512 : 256; }();
    // function attributes are per device: set on every call
    if(threads == 256) MCG_CUDA(cudaFuncSetAttribute(rs_scatter_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (2 * 256 * RS_ROUNDS + 8 * 256) * 4));
    else               MCG_CUDA(cudaFuncSetAttribute(rs_scatter_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (2 * 512 * RS_ROUNDS + 16 * 256) * 4));
    const int tile = threads * RS_ROUNDS;
    const size_t smem = size_t(2 * tile + (threads / 32) * 256) * 4;
    const uint32_t ntiles = uint32_t((n + tile - 1) / tile);
    const uint64_t nh = uint64_t(256) * ntiles;
    MCG_TRY(c->tmp5.reserve((nh + 2) * sizeof(uint32_t)));
    uint32_t* hist = c->tmp5.as<uint32_t>();
    const int passes = (key_bits + 7) / 8;
    uint32_t* ki = keys; uint32_t* vi = vals; uint32_t* ko = keys_alt; uint32_t* vo = vals_alt;
    for(int p = 0; p < passes; ++p) {
        const int shift = 8 * p;
        const uint32_t iota = (iota_vals && p == 0) ? iota_div : 0u;
        if(threads == 256) rs_hist_kernel<256><<<ntiles, 256, 0, c->stream>>>(ki, n, shift, hist, ntiles);
        else               rs_hist_kernel<512><<<ntiles, 512, 0, c->stream>>>(ki, n, shift, hist, ntiles);
        c->launches++;
        MCG_TRY(exclusive_scan_u32_to_u32(c, hist, hist, nh));
        c->prof_begin(2, c->stream);
        if(threads == 256) rs_scatter_kernel<256><<<ntiles, 256, smem, c->stream>>>(ki, vi, ko, vo, n, shift, hist, ntiles, iota);
        else               rs_scatter_kernel<512><<<ntiles, 512, smem, c->stream>>>(ki, vi, ko, vo, n, shift, hist, ntiles, iota);
        c->prof_end(2, c->stream);
        c->launches++;
        MCG_CUDA(cudaGetLastError());
        uint32_t* t = ki; ki = ko; ko = t;
        t = vi; vi = vo; vo = t;
    }
    if(passes == 0 && iota_vals) {
        iota_kernel<<<unsigned(div_up(int64_t(n), 256)), 256, 0, c->stream>>>(vals, n, iota_div);
        c->launches++;
    }
    if(ki != keys) {        // odd number of passes: result sits in the alt buffers
        MCG_CUDA(cudaMemcpyAsync(keys, ki, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream));
        MCG_CUDA(cudaMemcpyAsync(vals, vi, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream));
    }
    return MCG_OK;
}

}  // namespace mcg
