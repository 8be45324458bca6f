// radix_sort.cu -- stable LSD radix sort of (uint32 key, uint32 value) pairs, 8 bits
// per pass.  Used by the database build (sort feature slots by feature; the
// stability keeps each feature's locations in ascending (target, window) order,
// which is what makes the reference's "first maxlocs win" capping reproducible:
// sketch_database.h:1089-1092, hash_multimap.h:706-713).
//
// Per pass: (1) per-tile digit histogram, (2) exclusive scan over
// [digit][tile], (3) tile-local stable counting sort in shared memory followed
// by coalesced run-wise stores.  HBM-bound: ~24 B per element per pass.
#include "common.cuh"

namespace mcg {

constexpr int RS_THREADS = 256;
constexpr int RS_ROUNDS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ROUNDS;     // 4096 elements per tile
constexpr int RS_WARPS = RS_THREADS / 32;

int exclusive_scan_u32_to_u32(mcg_ctx* c, const uint32_t* d_in, uint32_t* d_out, uint64_t n);

__global__ void __launch_bounds__(RS_THREADS)
rs_hist_kernel(const uint32_t* __restrict__ keys, uint64_t n, int shift, uint32_t* __restrict__ hist, uint32_t ntiles)
{
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = uint64_t(blockIdx.x) * RS_TILE;
#pragma unroll 4
    for(int r = 0; r < RS_ROUNDS; ++r) {
        const uint64_t i = base + uint64_t(r) * RS_THREADS + threadIdx.x;
        if(i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[uint64_t(threadIdx.x) * ntiles + blockIdx.x] = h[threadIdx.x];
}

// Ranking as in the "match" scheme of one-sweep sorts: a warp owns RS_ROUNDS x 32 CONSECUTIVE
// elements of the tile and ranks them round by round against its own 256 digit counters
// (__match_any_sync, no block barrier inside the loop); one pass over [warp][digit] turns the
// counters into tile positions.  Stable: tile order = (warp, round, lane).
__global__ void __launch_bounds__(RS_THREADS, 4)
rs_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                  uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                  uint64_t n, int shift, const uint32_t* __restrict__ offsets, uint32_t ntiles, int iota)
{
    __shared__ uint32_t s_key[RS_TILE];
    __shared__ uint32_t s_val[RS_TILE];
    __shared__ uint32_t s_start[256];            // tile-local start of each digit
    __shared__ uint32_t s_gbase[256];            // global base of each digit for this tile - s_start
    __shared__ uint32_t s_wcnt[RS_WARPS][256];   // per warp: running count, then exclusive prefix over warps
    __shared__ uint32_t s_wsum[RS_WARPS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t base = uint64_t(blockIdx.x) * RS_TILE;
    const int tile_n = int(min(uint64_t(RS_TILE), n - base));
    const int wbase = warp * (RS_ROUNDS * 32);
    const uint32_t lt = (1u << lane) - 1u;

#pragma unroll
    for(int w = 0; w < RS_WARPS; ++w) s_wcnt[w][tid] = 0;
    uint32_t mykey[RS_ROUNDS];
#pragma unroll
    for(int r = 0; r < RS_ROUNDS; ++r) {
        const int li = wbase + r * 32 + lane;
        mykey[r] = li < tile_n ? keys_in[base + li] : 0xFFFFFFFFu;
    }
    __syncthreads();
    uint32_t* wc = s_wcnt[warp];
    uint32_t myrank[RS_ROUNDS / 2];                         // two 16-bit ranks per register
#pragma unroll
    for(int r = 0; r < RS_ROUNDS; ++r) {
        const uint32_t d = (mykey[r] >> shift) & 255u;      // padding elements rank as digit 255 behind all real ones
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, d);
        const uint32_t before = wc[d];
        __syncwarp();
        if((peers & lt) == 0) wc[d] = before + __popc(peers);
        __syncwarp();
        const uint32_t rk = before + __popc(peers & lt);
        if(r & 1) myrank[r >> 1] |= rk << 16; else myrank[r >> 1] = rk;
    }
    __syncthreads();
    {
        // thread tid owns digit tid: exclusive prefix over the warps, then over the digits
        uint32_t run = 0;
#pragma unroll
        for(int w = 0; w < RS_WARPS; ++w) {
            const uint32_t cnt = s_wcnt[w][tid];
            s_wcnt[w][tid] = run;
            run += cnt;
        }
        uint32_t inc = run;
#pragma unroll
        for(int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if(lane >= d) inc += o;
        }
        if(lane == 31) s_wsum[warp] = inc;
        __syncthreads();
        uint32_t wb = 0;
        for(int w = 0; w < warp; ++w) wb += s_wsum[w];
        const uint32_t start = wb + inc - run;
        s_start[tid] = start;
        s_gbase[tid] = offsets[uint64_t(tid) * ntiles + blockIdx.x] - start;
    }
    __syncthreads();
#pragma unroll
    for(int r = 0; r < RS_ROUNDS; ++r) {
        const int li = wbase + r * 32 + lane;
        const uint32_t d = (mykey[r] >> shift) & 255u;
        if(li < tile_n) {
            const uint32_t pos = s_start[d] + wc[d] + ((r & 1) ? (myrank[r >> 1] >> 16) : (myrank[r >> 1] & 0xFFFFu));
            s_key[pos] = mykey[r];
            s_val[pos] = iota ? uint32_t(base + li) : vals_in[base + li];
        }
    }
    __syncthreads();
    // coalesced stores: consecutive tile positions of one digit are consecutive globally
#pragma unroll 4
    for(int i = tid; i < tile_n; i += RS_THREADS) {
        const uint32_t key = s_key[i];
        const uint32_t d = (key >> shift) & 255u;
        const uint64_t g = uint64_t(s_gbase[d] + uint32_t(i));
        keys_out[g] = key;
        vals_out[g] = s_val[i];
    }
}

__global__ void iota_kernel(uint32_t* v, uint64_t n)
{
    const uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if(i < n) v[i] = uint32_t(i);
}

int radix_sort_pairs(mcg_ctx* c, uint32_t* keys, uint32_t* vals, uint32_t* keys_alt, uint32_t* vals_alt,
                     uint64_t n, int key_bits, bool iota_vals)
{
    if(n == 0) return MCG_OK;
    MCG_REQUIRE(n < (1ull << 32), "radix_sort_pairs: more than 2^32-1 elements");
    const uint32_t ntiles = uint32_t((n + RS_TILE - 1) / RS_TILE);
    const uint64_t nh = uint64_t(256) * ntiles;
    MCG_TRY(c->tmp5.reserve((nh + 2) * sizeof(uint32_t)));
    uint32_t* hist = c->tmp5.as<uint32_t>();
    const int passes = (key_bits + 7) / 8;
    uint32_t* ki = keys; uint32_t* vi = vals; uint32_t* ko = keys_alt; uint32_t* vo = vals_alt;
    for(int p = 0; p < passes; ++p) {
        const int shift = 8 * p;
        rs_hist_kernel<<<ntiles, RS_THREADS, 0, c->stream>>>(ki, n, shift, hist, ntiles);
        c->launches++;
        MCG_TRY(exclusive_scan_u32_to_u32(c, hist, hist, nh));
        rs_scatter_kernel<<<ntiles, RS_THREADS, 0, c->stream>>>(ki, vi, ko, vo, n, shift, hist, ntiles,
                                                                (iota_vals && p == 0) ? 1 : 0);
        c->launches++;
        MCG_CUDA(cudaGetLastError());
        uint32_t* t = ki; ki = ko; ko = t;
        t = vi; vi = vo; vo = t;
    }
    if(passes == 0 && iota_vals) {
        iota_kernel<<<unsigned(div_up(int64_t(n), 256)), 256, 0, c->stream>>>(vals, n);
        c->launches++;
    }
    if(ki != keys) {        // odd number of passes: result sits in the alt buffers
        MCG_CUDA(cudaMemcpyAsync(keys, ki, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream));
        MCG_CUDA(cudaMemcpyAsync(vals, vi, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream));
    }
    return MCG_OK;
}

}  // namespace mcg
