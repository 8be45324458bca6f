// scan.cu -- device-wide exclusive prefix sums (three-phase: tile sums, scan of
// tile sums, tile scans with their base).  HBM-bound helper of the build and query paths.
#include "common.cuh"

namespace mcg {

constexpr int SC_THREADS = 256;
constexpr int SC_ITEMS = 16;
constexpr int SC_TILE = SC_THREADS * SC_ITEMS;

template<class T>
__device__ __forceinline__ T block_exclusive_scan(T v, T* total, T* sh /* [SC_THREADS/32] */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T inc = v;
#pragma unroll
    for(int d = 1; d < 32; d <<= 1) {
        const T o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if(lane >= d) inc += o;
    }
    if(lane == 31) sh[warp] = inc;
    __syncthreads();
    T wbase = 0, tot = 0;
#pragma unroll
    for(int w = 0; w < SC_THREADS / 32; ++w) {
        const T s = sh[w];
        if(w < warp) wbase += s;
        tot += s;
    }
    __syncthreads();
    *total = tot;
    return wbase + inc - v;
}

// out[i] = exclusive prefix within the tile; tile_sum[block] = tile total.
// Items beyond n count as 0.  When add != nullptr the tile base add[block] is added.
template<class TIn, class TOut>
__global__ void __launch_bounds__(SC_THREADS)
tile_scan_kernel(const TIn* in, TOut* out, uint64_t n, TOut* tile_sum, const TOut* add)
{
    __shared__ TOut sh[SC_THREADS / 32];
    const uint64_t base = uint64_t(blockIdx.x) * SC_TILE + uint64_t(threadIdx.x) * SC_ITEMS;
    TOut v[SC_ITEMS];
    TOut tsum = 0;
#pragma unroll
    for(int i = 0; i < SC_ITEMS; ++i) {
        const uint64_t idx = base + i;
        v[i] = idx < n ? TOut(in[idx]) : TOut(0);
        tsum += v[i];
    }
    TOut total;
    TOut ex = block_exclusive_scan<TOut>(tsum, &total, sh);
    if(add != nullptr) ex += add[blockIdx.x];
#pragma unroll
    for(int i = 0; i < SC_ITEMS; ++i) {
        const uint64_t idx = base + i;
        if(idx < n) out[idx] = ex;
        ex += v[i];
    }
    if(tile_sum != nullptr && threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}

// tile_sum[block] = sum of the tile (items beyond n count as 0): the first phase of a scan that
// reads its input twice instead of writing, re-reading and re-writing its output
template<class TIn, class TOut>
__global__ void __launch_bounds__(SC_THREADS)
tile_reduce_kernel(const TIn* in, uint64_t n, TOut* tile_sum)
{
    __shared__ TOut sh[SC_THREADS / 32];
    const uint64_t base = uint64_t(blockIdx.x) * SC_TILE;
    TOut s = 0;
#pragma unroll 4
    for(int i = 0; i < SC_ITEMS; ++i) {
        const uint64_t idx = base + uint64_t(i) * SC_THREADS + threadIdx.x;      // coalesced
        if(idx < n) s += TOut(in[idx]);
    }
#pragma unroll
    for(int d = 16; d > 0; d >>= 1) s += __shfl_down_sync(0xFFFFFFFFu, s, d);
    if((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if(threadIdx.x == 0) {
        TOut t = 0;
        for(int w = 0; w < SC_THREADS / 32; ++w) t += sh[w];
        tile_sum[blockIdx.x] = t;
    }
}

template<class T>
__global__ void set_value_kernel(T* p, T v) { *p = v; }

// exclusive scan of `in` (n items) -> out (n items); returns grand total in *d_total (device)
template<class TIn, class TOut>
static int scan_rec(mcg_ctx* c, const TIn* in, TOut* out, uint64_t n, TOut* lvl_mem, TOut* d_total)
{
    const uint64_t nb = (n + SC_TILE - 1) / SC_TILE;
    if(nb == 0) {
        set_value_kernel<TOut><<<1, 1, 0, c->stream>>>(d_total, TOut(0));
        c->launches++;
        return MCG_OK;
    }
    if(nb == 1) {
        tile_scan_kernel<TIn, TOut><<<1, SC_THREADS, 0, c->stream>>>(in, out, n, d_total, nullptr);
        c->launches++;
        return MCG_OK;
    }
    // reduce -> scan of the tile sums -> tile scans with their base: the input is read twice, the output
    // written once (tile scan + add-base pass: output written, re-read and re-written)
    TOut* sums = lvl_mem;                  // [nb]
    tile_reduce_kernel<TIn, TOut><<<unsigned(nb), SC_THREADS, 0, c->stream>>>(in, n, sums);
    c->launches++;
    MCG_TRY((scan_rec<TOut, TOut>(c, sums, sums, nb, lvl_mem + nb, d_total)));
    tile_scan_kernel<TIn, TOut><<<unsigned(nb), SC_THREADS, 0, c->stream>>>(in, out, n, nullptr, sums);
    c->launches++;
    return MCG_OK;
}

static size_t scan_scratch_items(uint64_t n)
{
    size_t tot = 1;
    uint64_t m = n;
    while(m > SC_TILE) { m = (m + SC_TILE - 1) / SC_TILE; tot += m; }
    return tot + 8;
}

int exclusive_scan_u32_to_u64(mcg_ctx* c, const uint32_t* d_in, uint64_t* d_out, uint64_t n)
{
    MCG_TRY(c->scan_tmp.reserve(scan_scratch_items(n) * sizeof(uint64_t)));
    MCG_TRY((scan_rec<uint32_t, uint64_t>(c, d_in, d_out, n, c->scan_tmp.as<uint64_t>(), d_out + n)));
    MCG_CUDA(cudaGetLastError());
    return MCG_OK;
}

int exclusive_scan_u8_to_u64(mcg_ctx* c, const uint8_t* d_in, uint64_t* d_out, uint64_t n)
{
    MCG_TRY(c->scan_tmp.reserve(scan_scratch_items(n) * sizeof(uint64_t)));
    MCG_TRY((scan_rec<uint8_t, uint64_t>(c, d_in, d_out, n, c->scan_tmp.as<uint64_t>(), d_out + n)));
    MCG_CUDA(cudaGetLastError());
    return MCG_OK;
}

int exclusive_scan_u32_to_u32(mcg_ctx* c, const uint32_t* d_in, uint32_t* d_out, uint64_t n)
{
    MCG_TRY(c->scan_tmp.reserve(scan_scratch_items(n) * sizeof(uint64_t)));
    MCG_TRY((scan_rec<uint32_t, uint32_t>(c, d_in, d_out, n, c->scan_tmp.as<uint32_t>(), d_out + n)));
    MCG_CUDA(cudaGetLastError());
    return MCG_OK;
}

// d[0..n) counts -> d[0..n] exclusive prefix (d[n] = total); optionally copies the total to the host
int exclusive_scan_u32_inplace(mcg_ctx* c, uint32_t* d, uint64_t n, uint64_t* h_total)
{
    MCG_TRY(exclusive_scan_u32_to_u32(c, d, d, n));
    if(h_total != nullptr) {
        uint32_t t = 0;
        MCG_CUDA(cudaMemcpyAsync(&t, d + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
        MCG_CUDA(cudaStreamSynchronize(c->stream));
        *h_total = t;
    }
    return MCG_OK;
}

}  // namespace mcg
