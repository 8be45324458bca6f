// build.cu -- K2: database shard construction and maintenance.
//
// Reference semantics reproduced (not its layout):
//   sketch_database.h:519-563  add_target_distributed (shard = target_id % nshards, :540)
//   sketch_database.h:1079-1097 add_all_window_sketches: every feature of every window
//                               is inserted with location {tgt,win}; lists are cut to
//                               maxlocs (first come = smallest (tgt,win), quirk Q11)
//   hash_multimap.h:653,706,720,779  insert / shrink / clear(bucket) / find
//   sketch_database.h:398-417,1060-1075 + mode_build.cpp:847-1074 overpopulated-feature removal
//
// HBM layout of a finalized shard:
//   keys [nkeys]  uint32 ascending        sizes[nkeys] uint8 (0 = emptied "dead" key)
//   offs [nkeys+1] uint64 (CSR)           locs [nlocs] uint64 = tgt << 32 | win
//   table[nbuckets] 128-byte buckets of 3 slots {key | size, off, loc0, loc1} (layout and probe
//                 rule: common.cuh, "feature table"); the first two locations of a list sit in
//                 the bucket line itself, so the (dominant) lists of length <= 2 need no second
//                 memory access at query time.
#include "common.cuh"

#include <algorithm>
#include <numeric>
#include <ctime>

namespace mcg {

int copy_in(mcg_ctx* c, void* d, const void* h, size_t bytes);
int copy_out(mcg_ctx* c, void* h, const void* d, size_t bytes);
int exclusive_scan_u32_to_u32(mcg_ctx* c, const uint32_t* d_in, uint32_t* d_out, uint64_t n);

constexpr uint32_t EMPTY_KEY = 0xFFFFFFFFu;
constexpr double TABLE_LOAD = 0.45;      // keys per slot

// ---- runs of equal feature in the sorted slot list -> keys, sizes, offsets, locations
//
// Everything a slot needs is a prefix quantity of slot-local values: its key index is the number
// of run heads up to it, its position in the run is its index minus the last head index (a
// max-scan), and -- because exactly the first maxlocs slots of a run are kept
// (sketch_database.h:1089-1092, hash_multimap.h:706-713) -- its location index is the number of
// kept slots before it.  Two passes over tiles of the sorted keys: COUNT (heads and kept slots
// per tile), a scan of the tile counts, EMIT (keys, sizes, offsets and locations written in
// place).  The run a tile starts in is found by galloping backwards from the tile boundary.
// Padding slots (0xFFFFFFFF) sort to the end and are neither heads nor kept.
constexpr int RN_THREADS = 256, RN_ITEMS = 8, RN_TILE = RN_THREADS * RN_ITEMS;

__device__ __forceinline__ void rn_load8(const uint32_t* a, uint64_t e0, uint64_t n, uint32_t fill, uint32_t (&v)[RN_ITEMS])
{
    if(e0 + RN_ITEMS <= n) {
        const uint4 x = __ldg(reinterpret_cast<const uint4*>(a + e0)), y = __ldg(reinterpret_cast<const uint4*>(a + e0) + 1);
        v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
    } else {
#pragma unroll
        for(int i = 0; i < RN_ITEMS; ++i) v[i] = e0 + i < n ? a[e0 + i] : fill;
    }
}

// target of a global window: coarse[c] = largest t with firstwin[t] <= c << TGT_COARSE_SHIFT
constexpr int TGT_COARSE_SHIFT = 10;

__global__ void tgt_coarse_kernel(const uint64_t* __restrict__ firstwin, uint32_t ntgt, uint32_t ncells, uint32_t* __restrict__ coarse)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if(c > ncells) return;
    const uint64_t w = uint64_t(c) << TGT_COARSE_SHIFT;
    uint32_t lo = 0, hi = ntgt;
    while(hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if(firstwin[mid] <= w) lo = mid; else hi = mid;
    }
    coarse[c] = lo;
}

template<bool EMIT>
__global__ void __launch_bounds__(RN_THREADS)
runs_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ slots, uint64_t n, uint32_t maxlocs,
            uint32_t* __restrict__ tile_heads, uint32_t* __restrict__ tile_kept,      // COUNT: per-tile counts out; EMIT: their exclusive prefixes
            uint32_t* __restrict__ ukeys, uint8_t* __restrict__ sizes, uint64_t* __restrict__ offs, uint64_t* __restrict__ locs,
            const uint32_t* __restrict__ tgt_coarse, const uint64_t* __restrict__ tgt_firstwin, const uint32_t* __restrict__ tgt_id, uint32_t ntgt)
{
    __shared__ long long s_max[RN_THREADS / 32];
    __shared__ unsigned long long s_sum[RN_THREADS / 32];
    __shared__ long long s_first;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t tbase = uint64_t(blockIdx.x) * RN_TILE;
    const uint64_t e0 = tbase + uint64_t(tid) * RN_ITEMS;
    uint32_t k[RN_ITEMS];
    rn_load8(keys, e0, n, EMPTY_KEY, k);
    const uint32_t prev = (e0 > 0 && e0 <= n) ? keys[e0 - 1] : ~k[0];
    if(tid == 0) {
        long long st = (long long)tbase;
        if(tbase > 0 && k[0] != EMPTY_KEY && prev == k[0]) {
            // keys[hi] == k[0]; step back 1, 2, 4, ... until a different key (or the front), then bisect
            long long hi = (long long)tbase - 1, lo = -1, step = 1;
            while(hi - step >= 0) {
                if(keys[hi - step] != k[0]) { lo = hi - step; break; }
                hi -= step;
                step <<= 1;
            }
            while(hi - lo > 1) {
                const long long mid = (lo + hi) >> 1;
                if(keys[mid] == k[0]) hi = mid; else lo = mid;
            }
            st = hi;
        }
        s_first = st;
    }
    // heads of this thread; index of its last head
    uint32_t hmask = 0;
    long long last = -1;
#pragma unroll
    for(int i = 0; i < RN_ITEMS; ++i) {
        const uint32_t p = i == 0 ? prev : k[i - 1];
        if(k[i] != EMPTY_KEY && p != k[i]) { hmask |= 1u << i; last = (long long)(e0 + i); }
    }
    // exclusive max-scan of `last` over the block: start of the run this thread begins in
    long long inc = last;
#pragma unroll
    for(int d = 1; d < 32; d <<= 1) {
        const long long o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if(lane >= d) inc = max(inc, o);
    }
    if(lane == 31) s_max[warp] = inc;
    long long start = __shfl_up_sync(0xFFFFFFFFu, inc, 1);
    if(lane == 0) start = -1;
    __syncthreads();
    for(int w = 0; w < warp; ++w) start = max(start, s_max[w]);
    if(start < 0) start = s_first;
    // positions in the run, kept slots
    uint32_t kmask = 0;
    uint32_t pos[RN_ITEMS];
#pragma unroll
    for(int i = 0; i < RN_ITEMS; ++i) {
        if(hmask >> i & 1u) start = (long long)(e0 + i);
        pos[i] = uint32_t((long long)(e0 + i) - start);
        if(k[i] != EMPTY_KEY && pos[i] < maxlocs) kmask |= 1u << i;
    }
    // block sums / exclusive prefixes of (heads, kept), packed
    const unsigned long long mine = (unsigned long long)(__popc(hmask)) << 32 | (unsigned long long)(__popc(kmask));
    unsigned long long sinc = mine;
#pragma unroll
    for(int d = 1; d < 32; d <<= 1) {
        const unsigned long long o = __shfl_up_sync(0xFFFFFFFFu, sinc, d);
        if(lane >= d) sinc += o;
    }
    if(lane == 31) s_sum[warp] = sinc;
    __syncthreads();
    if(!EMIT) {
        if(tid == 0) {
            unsigned long long tot = 0;
            for(int w = 0; w < RN_THREADS / 32; ++w) tot += s_sum[w];
            tile_heads[blockIdx.x] = uint32_t(tot >> 32);
            tile_kept[blockIdx.x] = uint32_t(tot);
        }
        return;
    }
    unsigned long long ex = sinc - mine;
    for(int w = 0; w < warp; ++w) ex += s_sum[w];
    uint32_t kidx = tile_heads[blockIdx.x] + uint32_t(ex >> 32) - 1u;      // run this thread begins in (may have started in an earlier tile)
    uint64_t lidx = uint64_t(tile_kept[blockIdx.x]) + uint32_t(ex);
    uint32_t sl[RN_ITEMS];
    if(kmask) rn_load8(slots, e0, n, 0u, sl);
    const uint32_t next = e0 + RN_ITEMS < n ? keys[e0 + RN_ITEMS] : EMPTY_KEY;
#pragma unroll
    for(int i = 0; i < RN_ITEMS; ++i) {
        if(hmask >> i & 1u) {
            ++kidx;
            ukeys[kidx] = k[i];
            offs[kidx] = lidx;
        }
        const uint32_t nk = i + 1 < RN_ITEMS ? k[i + 1] : next;
        if(k[i] != EMPTY_KEY && nk != k[i]) sizes[kidx] = uint8_t(min(pos[i] + 1u, maxlocs));
        if(kmask >> i & 1u) {
            // the sort carried the GLOBAL WINDOW of every slot; its target = largest t with firstwin[t] <= gw, looked
            // up between the targets that hold the first window of this and of the next coarse cell (usually the same)
            const uint64_t gw = sl[i];
            const uint32_t cell = uint32_t(gw >> TGT_COARSE_SHIFT);
            uint32_t lo = __ldg(tgt_coarse + cell), hi = min(__ldg(tgt_coarse + cell + 1) + 1u, ntgt);
            while(hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if(__ldg(tgt_firstwin + mid) <= gw) lo = mid; else hi = mid;
            }
            locs[lidx++] = (uint64_t(__ldg(tgt_id + lo)) << 32) | (gw - __ldg(tgt_firstwin + lo));
        }
    }
}

__global__ void set_u64_kernel(uint64_t* p, uint64_t v) { *p = v; }

// ---- ordered table construction (layout and rule: common.cuh, "feature table")
__global__ void table_clear_kernel(uint4* table, uint64_t nbuckets)
{
    // the 32-byte key sector of every bucket, as a whole sector (no fill from DRAM); payload
    // sectors are only ever read after a key match
    const uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if(i < 2 * nbuckets) table[8 * (i >> 1) + (i & 1)] = (i & 1) ? make_uint4(0u, 0u, 0u, 0u) : make_uint4(EMPTY_KEY, EMPTY_KEY, EMPTY_KEY, 0u);
}

// cdf[c] = number of keys below c << (32 - TABLE_CDF_BITS)
__global__ void table_cdf_kernel(const uint32_t* ukeys, uint32_t nkeys, uint32_t* cdf)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if(c > (1u << TABLE_CDF_BITS)) return;
    if(c == (1u << TABLE_CDF_BITS)) { cdf[c] = nkeys; return; }
    const uint32_t bound = c << (32 - TABLE_CDF_BITS);
    uint32_t lo = 0, hi = nkeys;
    while(lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if(ukeys[mid] < bound) lo = mid + 1; else hi = mid;
    }
    cdf[c] = lo;
}

constexpr int TB_THREADS = 256, TB_ITEMS = 8, TB_TILE = TB_THREADS * TB_ITEMS;

// d_j = 3 * home_j - j; slot_j = j + max_{i <= j} d_i.  Pass 1: maximum of d per tile.
__global__ void __launch_bounds__(TB_THREADS)
table_tilemax_kernel(const uint32_t* ukeys, uint32_t nkeys, TableGeom g, int64_t* tmax)
{
    __shared__ int64_t s[TB_THREADS / 32];
    const uint64_t base = uint64_t(blockIdx.x) * TB_TILE + uint64_t(threadIdx.x) * TB_ITEMS;
    int64_t m = INT64_MIN;
    for(int t = 0; t < TB_ITEMS; ++t) {
        const uint64_t j = base + t;
        if(j < nkeys) m = max(m, int64_t(TABLE_BUCKET_SLOTS) * table_bucket(ukeys[j], g) - int64_t(j));
    }
    for(int d = 16; d > 0; d >>= 1) m = max(m, __shfl_down_sync(0xFFFFFFFFu, m, d));
    if((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
    __syncthreads();
    if(threadIdx.x == 0) {
        for(int w = 1; w < TB_THREADS / 32; ++w) m = max(m, s[w]);
        tmax[blockIdx.x] = m;
    }
}

// pass 2 (one block): exclusive running maximum over the tiles; out[ntiles] = overall maximum
__global__ void __launch_bounds__(1024)
table_tilescan_kernel(int64_t* tmax, uint32_t ntiles)
{
    __shared__ int64_t s[32];
    __shared__ int64_t s_carry;
    if(threadIdx.x == 0) s_carry = INT64_MIN;
    __syncthreads();
    for(uint32_t b0 = 0; b0 < ntiles; b0 += 1024) {
        const uint32_t i = b0 + threadIdx.x;
        const int64_t own = i < ntiles ? tmax[i] : INT64_MIN;
        int64_t v = own;
        for(int d = 1; d < 32; d <<= 1) {
            const int64_t o = __shfl_up_sync(0xFFFFFFFFu, v, d);
            if(int(threadIdx.x & 31) >= d) v = max(v, o);
        }
        if((threadIdx.x & 31) == 31) s[threadIdx.x >> 5] = v;
        __syncthreads();
        int64_t pre = s_carry;
        for(uint32_t w = 0; w < (threadIdx.x >> 5); ++w) pre = max(pre, s[w]);
        const int64_t incl = max(pre, v);
        // exclusive value of element i = max(carry, everything before i)
        int64_t excl = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
        if((threadIdx.x & 31) == 0) {
            excl = s_carry;
            for(uint32_t w = 0; w < (threadIdx.x >> 5); ++w) excl = max(excl, s[w]);
        }
        __syncthreads();
        if(i < ntiles) tmax[i] = excl;
        if(threadIdx.x == 1023) s_carry = incl;
        __syncthreads();
    }
    if(threadIdx.x == 0) tmax[ntiles] = s_carry;
}

// pass 3: running maximum inside the tile (+ carry of pass 2) -> slot of every key.  The table is
// written BUCKET-centric: tile t owns the buckets (E(t-1), E(t)], E(t) = bucket of its last key (the
// last tile: up to the end of the table), assembles them -- empty ones included -- in shared memory,
// TBL_CH buckets at a time, and streams them out as whole 128-byte lines: every byte of the table is
// written exactly once, front to back, no clearing pass, no partial-sector writes.  Bucket E(t) can
// also hold the first one or two keys of tile t+1: tile t places them (their slots follow from its
// own running maximum), tile t+1 skips them.
constexpr int TBL_CH = 256;

__global__ void __launch_bounds__(TB_THREADS)
table_build_kernel(uint4* __restrict__ table4, const uint32_t* __restrict__ ukeys, const uint8_t* __restrict__ sizes,
                   const uint64_t* __restrict__ offs, const uint64_t* __restrict__ locs,
                   uint32_t nkeys, TableGeom g, const int64_t* __restrict__ tcarry, uint64_t cap, uint32_t ntiles, int key32, int wbits)
{
    extern __shared__ __align__(16) uint4 lines[];              // [TBL_CH][8]
    __shared__ int64_t s[TB_THREADS / 32];
    __shared__ int32_t s_slot[TB_TILE + 2];                     // slot - 3 * B0 of the tile's keys (+ <= 2 keys of the next tile)
    __shared__ int64_t s_end[2];                                // running maximum / slot at the last key of the tile
    __shared__ int s_extra;
    const int tid = threadIdx.x;
    const uint64_t j0 = uint64_t(blockIdx.x) * TB_TILE;
    const uint64_t j1 = min(j0 + uint64_t(TB_TILE), uint64_t(nkeys));
    const uint64_t base = j0 + uint64_t(tid) * TB_ITEMS;
    int64_t d[TB_ITEMS];
    int64_t m = INT64_MIN;
#pragma unroll
    for(int t = 0; t < TB_ITEMS; ++t) {
        const uint64_t j = base + t;
        d[t] = j < nkeys ? int64_t(TABLE_BUCKET_SLOTS) * table_bucket(ukeys[j], g) - int64_t(j) : INT64_MIN;
        m = max(m, d[t]);
        d[t] = m;                                   // inclusive running maximum inside the thread
    }
    const int lane = tid & 31, warp = tid >> 5;
    int64_t v = m;
    for(int dd = 1; dd < 32; dd <<= 1) {
        const int64_t o = __shfl_up_sync(0xFFFFFFFFu, v, dd);
        if(lane >= dd) v = max(v, o);
    }
    if(lane == 31) s[warp] = v;
    if(tid == 0) s_extra = 0;
    __syncthreads();
    const int64_t mprev = tcarry[blockIdx.x];
    int64_t pre = mprev;
    for(int w = 0; w < warp; ++w) pre = max(pre, s[w]);
    int64_t before = __shfl_up_sync(0xFFFFFFFFu, v, 1);     // maximum of the lanes before this one
    if(lane == 0) before = INT64_MIN;
    pre = max(pre, before);
    const uint64_t B0 = blockIdx.x == 0 ? 0ull : uint64_t(int64_t(j0) - 1 + mprev) / TABLE_BUCKET_SLOTS + 1ull;
#pragma unroll
    for(int t = 0; t < TB_ITEMS; ++t) {
        const uint64_t j = base + t;
        if(j < j1) {
            const int64_t mx = max(pre, d[t]);
            const int64_t slot = int64_t(j) + mx;
            s_slot[j - j0] = int32_t(slot - int64_t(TABLE_BUCKET_SLOTS * B0));      // negative: placed by the previous tile
            if(j + 1 == j1) { s_end[0] = mx; s_end[1] = slot; }
        }
    }
    __syncthreads();
    const bool last_tile = blockIdx.x + 1 == ntiles;
    const uint64_t B1 = last_tile ? cap - 1 : uint64_t(s_end[1]) / TABLE_BUCKET_SLOTS;
    if(tid == 0 && !last_tile) {
        int64_t mx = s_end[0];
        int extra = 0;
        for(int q = 0; q < 2; ++q) {
            const uint64_t jn = j1 + q;
            if(jn >= nkeys) break;
            mx = max(mx, int64_t(TABLE_BUCKET_SLOTS) * table_bucket(ukeys[jn], g) - int64_t(jn));
            const int64_t slot = int64_t(jn) + mx;
            if(uint64_t(slot) / TABLE_BUCKET_SLOTS != B1) break;
            s_slot[TB_TILE + q] = int32_t(slot - int64_t(TABLE_BUCKET_SLOTS * B0));
            ++extra;
        }
        s_extra = extra;
    }
    __syncthreads();
    // items 0 .. nitems-1 are the keys j0 + i (a full tile whenever there is a next one)
    const int nitems = int(j1 - j0) + s_extra;
    int ia = 0;
    while(ia < nitems && s_slot[ia] < 0) ++ia;              // <= 2 keys
    uint32_t* words = reinterpret_cast<uint32_t*>(lines);
    for(uint64_t cb = B0; cb <= B1; cb += TBL_CH) {
        const int nb = int(min(uint64_t(TBL_CH), B1 - cb + 1));
        for(int i = tid; i < nb * 8; i += TB_THREADS)
            lines[i] = (i & 7) == 0 ? make_uint4(TABLE_EMPTY, TABLE_EMPTY, TABLE_EMPTY, 0u) : make_uint4(0u, 0u, 0u, 0u);
        const int64_t lo_rel = int64_t(TABLE_BUCKET_SLOTS) * int64_t(cb - B0), hi_rel = lo_rel + int64_t(TABLE_BUCKET_SLOTS) * nb;
        int lo = ia, hi = nitems;                           // first item at or beyond the end of the chunk
        while(lo < hi) {
            const int mid = (lo + hi) >> 1;
            if(int64_t(s_slot[mid]) < hi_rel) lo = mid + 1; else hi = mid;
        }
        const int ib = lo;
        __syncthreads();
        for(int i = ia + tid; i < ib; i += TB_THREADS) {
            const uint64_t j = j0 + uint64_t(i);
            const uint32_t rel = uint32_t(int64_t(s_slot[i]) - lo_rel);
            const uint32_t b = rel / TABLE_BUCKET_SLOTS, st = rel - b * TABLE_BUCKET_SLOTS;
            const uint32_t size = sizes[j];
            const uint64_t o = offs[j];
            uint32_t w[6] = {0u, 0u, 0u, 0u, 0u, 0u};
            if(key32) {
#pragma unroll
                for(uint32_t t = 0; t < TABLE_INLINE_32; ++t)
                    if(t < size) { const uint64_t l = locs[o + t]; w[t] = (uint32_t(l >> 32) << wbits) | uint32_t(l); }
            } else {
#pragma unroll
                for(uint32_t t = 0; t < TABLE_INLINE_64; ++t)
                    if(t < size) { const uint64_t l = locs[o + t]; w[2 * t] = uint32_t(l); w[2 * t + 1] = uint32_t(l >> 32); }
            }
            words[b * TABLE_BUCKET_WORDS + st] = ukeys[j];
            lines[b * 8 + 2 + 2 * st] = make_uint4(size, uint32_t(o), w[0], w[1]);
            lines[b * 8 + 3 + 2 * st] = make_uint4(w[2], w[3], w[4], w[5]);
        }
        __syncthreads();
        uint4* out = table4 + cb * 8;
        for(int i = tid; i < nb * 8; i += TB_THREADS) out[i] = lines[i];
        __syncthreads();
        ia = ib;
    }
}

// clear(bucket) for a list of keys: hash_multimap.h:720 -- the key stays, its list is emptied
__global__ void remove_keys_kernel(uint32_t* table, TableGeom g, const uint32_t* ukeys, uint8_t* sizes, uint32_t nkeys,
                                   const uint32_t* rem, uint64_t nrem, unsigned long long* removed)
{
    const uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if(i >= nrem) return;
    const uint32_t key = rem[i];
    uint32_t lo = 0, hi = nkeys;
    while(lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if(ukeys[mid] < key) lo = mid + 1; else hi = mid;
    }
    if(lo >= nkeys || ukeys[lo] != key) return;
    atomicAdd(removed, 1ull);                       // found (reference counts found keys, :410-413)
    sizes[lo] = 0;
    uint32_t b = table_bucket(key, g);
    for(;;) {
        uint32_t* line = table + uint64_t(b) * TABLE_BUCKET_WORDS;
        const uint4 k4 = *reinterpret_cast<const uint4*>(line);
        const int t = table_match(k4, key);
        if(t >= 0) { line[8 + 8 * t] = 0u; return; }
        if(t == -1) return;
        ++b;
    }
}

__global__ void count_sizes_kernel(const uint8_t* sizes, uint32_t nkeys, unsigned long long* out /* [2] */)
{
    __shared__ unsigned long long s_ne, s_nl;
    if(threadIdx.x == 0) { s_ne = 0; s_nl = 0; }
    __syncthreads();
    unsigned long long ne = 0, nl = 0;
    for(uint64_t j = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; j < nkeys; j += uint64_t(gridDim.x) * blockDim.x) {
        const uint32_t s = sizes[j];
        ne += s != 0;
        nl += s;
    }
    for(int d = 16; d > 0; d >>= 1) {
        ne += __shfl_down_sync(0xFFFFFFFFu, ne, d);
        nl += __shfl_down_sync(0xFFFFFFFFu, nl, d);
    }
    if((threadIdx.x & 31) == 0) { atomicAdd(&s_ne, ne); atomicAdd(&s_nl, nl); }
    __syncthreads();
    if(threadIdx.x == 0) { atomicAdd(&out[0], s_ne); atomicAdd(&out[1], s_nl); }
}

static inline unsigned grid_for(uint64_t n, int threads = 256) { return unsigned((n + threads - 1) / threads); }

static int build_table(mcg_db* db)
{
    mcg_ctx* c = db->ctx;
    const uint64_t nkeys = db->nkeys;
    MCG_REQUIRE(nkeys < (1ull << 32), "too many keys in one shard");
    MCG_REQUIRE(db->nlocs < (1ull << 32), "too many locations in one shard");       // slots hold 32-bit offsets into locs
    static double load = -1.0;
    if(load < 0) { const char* e = getenv("MCG_TABLE_LOAD"); load = e ? atof(e) : TABLE_LOAD; if(load < 0.05 || load > 0.95) load = TABLE_LOAD; }
    const uint64_t home = uint64_t(double(nkeys) / load / double(TABLE_BUCKET_SLOTS)) + 1;      // home buckets
    MCG_REQUIRE(home < (1ull << 31), "feature table too large");
    db->table_scale = nkeys > 0 ? uint32_t(std::min<uint64_t>((home << 32) / nkeys, 0xFFFFFFFFull)) : 0u;
    MCG_TRY(db->cdf.reserve(((1u << TABLE_CDF_BITS) + 2) * sizeof(uint32_t)));
    table_cdf_kernel<<<((1u << TABLE_CDF_BITS) + 256) / 256, 256, 0, c->stream>>>(db->keys.as<uint32_t>(), uint32_t(nkeys), db->cdf.as<uint32_t>());
    c->launches++;
    const TableGeom g = db->geom();
    // slot of the last key = nkeys - 1 + max_j (3 * home_j - j): the table ends right after it
    uint64_t last_slot = 0;
    const uint32_t ntiles = uint32_t((nkeys + TB_TILE - 1) / TB_TILE);
    if(nkeys > 0) {
        MCG_TRY(c->tmp5.reserve((size_t(ntiles) + 2) * sizeof(int64_t)));
        table_tilemax_kernel<<<ntiles, TB_THREADS, 0, c->stream>>>(db->keys.as<uint32_t>(), uint32_t(nkeys), g, c->tmp5.as<int64_t>());
        table_tilescan_kernel<<<1, 1024, 0, c->stream>>>(c->tmp5.as<int64_t>(), ntiles);
        c->launches += 2;
        int64_t dmax = 0;
        MCG_TRY(copy_out(c, &dmax, c->tmp5.as<int64_t>() + ntiles, sizeof(int64_t)));
        last_slot = uint64_t(int64_t(nkeys) - 1 + dmax);
    }
    const uint64_t cap = last_slot / TABLE_BUCKET_SLOTS + 2;        // + an empty bucket that ends every walk
    MCG_REQUIRE(cap < (1ull << 32), "feature table too large");
    MCG_TRY(db->table.reserve(cap * 128));
    db->table_cap = cap;
    if(nkeys > 0) {
        // (function attributes are per device: set on every call, not once per process)
        MCG_CUDA(cudaFuncSetAttribute(table_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TBL_CH * 128));
        table_build_kernel<<<ntiles, TB_THREADS, TBL_CH * 128, c->stream>>>(db->table.as<uint4>(), db->keys.as<uint32_t>(), db->sizes.as<uint8_t>(),
                                                                          db->offs.as<uint64_t>(), db->locs.as<uint64_t>(), uint32_t(nkeys), g,
                                                                          c->tmp5.as<int64_t>(), cap, ntiles, db->key32 ? 1 : 0, db->wbits);
    } else {
        table_clear_kernel<<<grid_for(2 * cap), 256, 0, c->stream>>>(db->table.as<uint4>(), cap);
    }
    c->launches++;
    MCG_CUDA(cudaGetLastError());
    return MCG_OK;
}

static double wall() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
#define TRACE(msg) do { if(trace) { cudaStreamSynchronize(c->stream); const double t = wall(); fprintf(stderr, "[mcg finalize] %-22s %8.2f ms\n", msg, 1e3 * (t - t_prev)); t_prev = t; } } while(0)

int db_finalize(mcg_db* db)
{
    mcg_ctx* c = db->ctx;
    const bool trace = getenv("MCG_TRACE") != nullptr;
    double t_prev = wall();
    MCG_CUDA(cudaSetDevice(c->device));
    const uint32_t s = uint32_t(db->cfg.sketchlen);
    const uint64_t nslots = db->nwin * s;
    MCG_REQUIRE(nslots < (1ull << 32), "shard holds %llu feature slots; the device build supports < 2^32 per shard",
                (unsigned long long)nslots);
    db->nkeys = 0;
    db->nlocs = 0;
    if(nslots == 0) {
        MCG_TRY(db->keys.reserve(16)); MCG_TRY(db->sizes.reserve(16)); MCG_TRY(db->offs.reserve(16)); MCG_TRY(db->locs.reserve(16));
        MCG_CUDA(cudaMemsetAsync(db->offs.p, 0, 16, c->stream));
        MCG_TRY(build_table(db));
        db->finalized = true;
        return MCG_OK;
    }
    // (1) sort slots by feature, stable; value = slot index (window-major => (tgt,win) order)
    MCG_TRY(c->tmp0.reserve(nslots * 4));     // slot ids
    MCG_TRY(c->tmp1.reserve(nslots * 4));     // keys alt
    MCG_TRY(c->tmp2.reserve(nslots * 4));     // vals alt
    uint32_t* skeys = db->feat.as<uint32_t>();
    uint32_t* svals = c->tmp0.as<uint32_t>();
    TRACE("alloc sort buffers");
    MCG_TRY(radix_sort_pairs(c, skeys, svals, c->tmp1.as<uint32_t>(), c->tmp2.as<uint32_t>(), nslots, 32, true, s));   // value = slot / s = window
    TRACE("radix sort");
    // (2) runs of equal feature: counts per tile, scan, then keys / sizes / offsets / locations in one pass
    const uint32_t ntiles = uint32_t((nslots + RN_TILE - 1) / RN_TILE);
    MCG_TRY(c->tmp1.reserve((size_t(ntiles) + 2) * 4));
    MCG_TRY(c->tmp2.reserve((size_t(ntiles) + 2) * 4));
    uint32_t* th = c->tmp1.as<uint32_t>();
    uint32_t* tk = c->tmp2.as<uint32_t>();
    const uint32_t ntgt = uint32_t(db->h_tgt_id.size());
    MCG_TRY(c->tmp3.reserve((size_t(ntgt) + 2) * 8 + 64));
    MCG_TRY(c->tmp4.reserve((size_t(ntgt) + 2) * 4 + 64));
    MCG_CUDA(cudaMemcpyAsync(c->tmp3.p, db->h_tgt_firstwin.data(), (size_t(ntgt) + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    MCG_CUDA(cudaMemcpyAsync(c->tmp4.p, db->h_tgt_id.data(), size_t(ntgt) * 4, cudaMemcpyHostToDevice, c->stream));
    const uint32_t ncells = uint32_t(db->nwin >> TGT_COARSE_SHIFT) + 1;
    MCG_TRY(c->tmp5.reserve((size_t(ncells) + 2) * 4));
    tgt_coarse_kernel<<<grid_for(ncells + 1), 256, 0, c->stream>>>(c->tmp3.as<uint64_t>(), ntgt, ncells, c->tmp5.as<uint32_t>());
    runs_kernel<false><<<ntiles, RN_THREADS, 0, c->stream>>>(skeys, svals, nslots, uint32_t(db->maxlocs), th, tk, nullptr, nullptr, nullptr, nullptr,
                                                             nullptr, nullptr, nullptr, 0);
    c->launches++;
    c->launches++;
    MCG_TRY(exclusive_scan_u32_to_u32(c, th, th, ntiles));
    MCG_TRY(exclusive_scan_u32_to_u32(c, tk, tk, ntiles));
    uint32_t nkeys = 0, nlocs32 = 0;
    MCG_CUDA(cudaMemcpyAsync(&nkeys, th + ntiles, 4, cudaMemcpyDeviceToHost, c->stream));
    MCG_CUDA(cudaMemcpyAsync(&nlocs32, tk + ntiles, 4, cudaMemcpyDeviceToHost, c->stream));
    MCG_CUDA(cudaStreamSynchronize(c->stream));       // (also: the host vectors were the copy source)
    TRACE("runs: count");
    const uint64_t nlocs = nlocs32;
    db->nkeys = nkeys;
    db->nlocs = nlocs;
    MCG_TRY(db->keys.reserve((uint64_t(nkeys) + 1) * 4));
    MCG_TRY(db->sizes.reserve(uint64_t(nkeys) + 16));
    MCG_TRY(db->offs.reserve((uint64_t(nkeys) + 2) * 8));
    MCG_TRY(db->locs.reserve((nlocs + 2) * 8));
    runs_kernel<true><<<ntiles, RN_THREADS, 0, c->stream>>>(skeys, svals, nslots, uint32_t(db->maxlocs), th, tk, db->keys.as<uint32_t>(),
                                                            db->sizes.as<uint8_t>(), db->offs.as<uint64_t>(), db->locs.as<uint64_t>(),
                                                            c->tmp5.as<uint32_t>(), c->tmp3.as<uint64_t>(), c->tmp4.as<uint32_t>(), ntgt);
    set_u64_kernel<<<1, 1, 0, c->stream>>>(db->offs.as<uint64_t>() + nkeys, nlocs);
    c->launches += 2;
    MCG_CUDA(cudaGetLastError());
    TRACE("runs: emit");
    // the sort scratch goes back before the table is allocated (peak memory)
    MCG_CUDA(cudaStreamSynchronize(c->stream));
    db->feat.release();
    c->tmp0.release(); c->tmp1.release(); c->tmp2.release();
    TRACE("free scratch");
    // (5) open-addressing table over the unique keys (its slots hold match keys in the shard's key layout)
    {
        uint64_t maxwin = 0;
        for(size_t t = 0; t + 1 < db->h_tgt_firstwin.size(); ++t) maxwin = std::max<uint64_t>(maxwin, db->h_tgt_firstwin[t + 1] - db->h_tgt_firstwin[t]);
        db->set_key_layout(db->h_tgt_id.empty() ? 0 : db->h_tgt_id.back(), maxwin);
        if(getenv("MCG_FORCE_KEY64")) { db->wbits = 32; db->key32 = false; }
    }
    MCG_TRY(build_table(db));
    MCG_CUDA(cudaStreamSynchronize(c->stream));
    TRACE("table");
    db->finalized = true;
    return MCG_OK;
}

int db_counts(mcg_db* db, uint64_t* nkeys, uint64_t* nonempty, uint64_t* nlocs)
{
    mcg_ctx* c = db->ctx;
    MCG_CUDA(cudaSetDevice(c->device));
    unsigned long long h[2] = {0, 0};
    if(db->nkeys > 0) {
        MCG_TRY(c->tmp3.reserve(64));
        MCG_CUDA(cudaMemsetAsync(c->tmp3.p, 0, 16, c->stream));
        count_sizes_kernel<<<c->num_sms * 4, 256, 0, c->stream>>>(db->sizes.as<uint8_t>(), uint32_t(db->nkeys),
                                                              c->tmp3.as<unsigned long long>());
        c->launches++;
        MCG_CUDA(cudaMemcpyAsync(h, c->tmp3.p, 16, cudaMemcpyDeviceToHost, c->stream));
        MCG_CUDA(cudaStreamSynchronize(c->stream));
    }
    if(nkeys) *nkeys = db->nkeys;
    if(nonempty) *nonempty = h[0];
    if(nlocs) *nlocs = h[1];
    return MCG_OK;
}

int db_remove_keys_dev(mcg_db* db, const uint32_t* d_keys, uint64_t n, uint64_t* removed)
{
    mcg_ctx* c = db->ctx;
    unsigned long long h = 0;
    if(n > 0 && db->nkeys > 0) {
        MCG_TRY(c->tmp3.reserve(64));
        MCG_CUDA(cudaMemsetAsync(c->tmp3.p, 0, 8, c->stream));
        remove_keys_kernel<<<grid_for(n), 256, 0, c->stream>>>(db->table.as<uint32_t>(), db->geom(), db->keys.as<uint32_t>(),
                                                              db->sizes.as<uint8_t>(), uint32_t(db->nkeys), d_keys, n,
                                                              c->tmp3.as<unsigned long long>());
        c->launches++;
        MCG_CUDA(cudaMemcpyAsync(&h, c->tmp3.p, 8, cudaMemcpyDeviceToHost, c->stream));
        MCG_CUDA(cudaStreamSynchronize(c->stream));
    }
    if(removed) *removed = h;
    return MCG_OK;
}

}  // namespace mcg

using namespace mcg;

extern "C" {

int mcg_db_create(mcg_ctx* c, const mcg_sketch_cfg* cfg, int32_t maxlocs, int32_t shard, int32_t nshards, mcg_db** out)
{
    MCG_REQUIRE(c != nullptr && cfg != nullptr && out != nullptr, "mcg_db_create: NULL argument");
    *out = nullptr;
    MCG_REQUIRE(cfg->k >= 1 && cfg->k <= MCG_MAX_K, "kmerlen %d outside 1..%d", cfg->k, MCG_MAX_K);
    MCG_REQUIRE(cfg->sketchlen >= 1 && cfg->sketchlen <= MCG_MAX_SKETCH, "sketchlen %d outside 1..%d", cfg->sketchlen, MCG_MAX_SKETCH);
    MCG_REQUIRE(cfg->winlen >= 1 && cfg->winstride >= 1, "window length / stride must be >= 1");
    MCG_REQUIRE(nshards >= 1 && shard >= 0 && shard < nshards, "shard %d / %d out of range", shard, nshards);
    mcg_db* db = new mcg_db;
    db->ctx = c;
    db->cfg = *cfg;
    // sketch_database.h:356-368 on a uint8 argument (mode_build.cpp:659; quirk Q7)
    int m = int(uint8_t(maxlocs));
    if(m < 1) m = 1;
    if(m >= MCG_MAX_LOCS) m = MCG_MAX_LOCS;
    db->maxlocs = m;
    db->shard = shard;
    db->nshards = nshards;
    db->h_tgt_firstwin.push_back(0);
    *out = db;
    return MCG_OK;
}

int mcg_db_destroy(mcg_db* db)
{
    if(db == nullptr) return MCG_OK;
    cudaSetDevice(db->ctx->device);
    cudaStreamSynchronize(db->ctx->stream);
    db->feat.release(); db->keys.release(); db->offs.release(); db->sizes.release(); db->locs.release(); db->table.release(); db->cdf.release();
    delete db;
    return MCG_OK;
}

// shared body of mcg_db_add_targets / mcg_db_add_targets_dev: `dev` = the sequences already are on the device
static int add_targets(mcg_db* db, const char* seq_bytes, const int64_t* seq_offsets,
                       const uint32_t* target_ids, int64_t n, uint32_t* out_windows, bool dev)
{
    MCG_REQUIRE(db != nullptr && seq_offsets != nullptr && target_ids != nullptr, "mcg_db_add_targets: NULL argument");
    MCG_REQUIRE(n >= 0, "n < 0");
    if(db->finalized) { set_error("database already finalized"); return MCG_ERR_STATE; }
    mcg_ctx* c = db->ctx;
    MCG_CUDA(cudaSetDevice(c->device));
    const int s = db->cfg.sketchlen;
    // owned targets of this call, packed back to back
    std::vector<int64_t> offs{0};
    std::vector<int64_t> src;           // index into the caller's arrays
    for(int64_t i = 0; i < n; ++i) {
        if(out_windows) out_windows[i] = 0;
        const uint32_t t = target_ids[i];
        if(i > 0) MCG_REQUIRE(t > target_ids[i - 1], "target ids must be strictly ascending");
        if(!db->h_tgt_id.empty()) MCG_REQUIRE(t > db->h_tgt_id.back() || int(t % uint32_t(db->nshards)) != db->shard,
                                              "target ids must ascend across calls");
        if(int(t % uint32_t(db->nshards)) != db->shard) continue;      // sketch_database.h:540
        const int64_t len = seq_offsets[i + 1] - seq_offsets[i];
        MCG_REQUIRE(len >= 0, "negative sequence length");
        src.push_back(i);
        offs.push_back(offs.back() + len);
    }
    const int64_t m = int64_t(src.size());
    if(m == 0) return MCG_OK;
    const int64_t total = offs.back();
    int64_t nwin = 0;
    std::vector<uint64_t> wins(m);
    for(int64_t j = 0; j < m; ++j) {
        wins[j] = uint64_t(num_windows(offs[j + 1] - offs[j], db->cfg.winlen, db->cfg.winstride));
        MCG_REQUIRE(wins[j] < (1ull << 32), "more than 2^32-1 windows in one target");     // window_id = uint32
        nwin += int64_t(wins[j]);
    }
    MCG_REQUIRE((db->nwin + uint64_t(nwin)) * uint64_t(s) < (1ull << 32),
                "shard would exceed 2^32-1 feature slots");
    // the owned sequences contiguously on the device: staged from the host, or -- device input whose
    // targets all belong to this shard -- used where they lie
    const bool in_place = dev && m == n;
    const char* d_seq = seq_bytes + seq_offsets[0];
    if(!in_place) {
        MCG_TRY(c->seq.reserve(size_t(total) + 64));
        int64_t j = 0;
        while(j < m) {                  // coalesce runs of adjacent source targets into one copy
            int64_t e = j;
            while(e + 1 < m && src[e + 1] == src[e] + 1) ++e;
            const int64_t b0 = seq_offsets[src[j]], b1 = seq_offsets[src[e] + 1];
            MCG_CUDA(cudaMemcpyAsync(c->seq.as<char>() + offs[j], seq_bytes + b0, size_t(b1 - b0),
                                     dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
            j = e + 1;
        }
        d_seq = c->seq.as<char>();
    }
    MCG_TRY(c->seqoff.reserve(size_t(m + 1) * 8));
    MCG_TRY(copy_in(c, c->seqoff.p, offs.data(), size_t(m + 1) * 8));
    // append the sketches of the new windows to the slot array
    MCG_TRY(db->feat.reserve_keep((db->nwin + uint64_t(nwin)) * s * 4, db->nwin * s * 4, c->stream));
    MCG_TRY(c->winfirst.reserve(size_t(m + 2) * 4));
    MCG_TRY(window_prefix(c, db->cfg.winlen, db->cfg.winstride, c->seqoff.as<int64_t>(), m, c->winfirst.as<uint32_t>(), nullptr));
    MCG_TRY(sketch_launch(c, db->cfg, s, db->cfg.winlen, db->cfg.winstride, d_seq, total,
                          c->seqoff.as<int64_t>(), m, c->winfirst.as<uint32_t>(), c->winfirst.as<uint32_t>(),
                          uint64_t(nwin), db->feat.as<uint32_t>() + db->nwin * s, nullptr));
    MCG_CUDA(cudaStreamSynchronize(c->stream));       // offs (host vector) was a copy source
    for(int64_t j = 0; j < m; ++j) {
        db->h_tgt_id.push_back(target_ids[src[j]]);
        db->h_tgt_firstwin.push_back(db->h_tgt_firstwin.back() + wins[j]);
        if(out_windows) out_windows[src[j]] = uint32_t(wins[j]);
    }
    db->nwin += uint64_t(nwin);
    return MCG_OK;
}

int mcg_db_add_targets(mcg_db* db, const char* seq_bytes, const int64_t* seq_offsets,
                       const uint32_t* target_ids, int64_t n, uint32_t* out_windows)
{
    return add_targets(db, seq_bytes, seq_offsets, target_ids, n, out_windows, false);
}

int mcg_db_add_targets_dev(mcg_db* db, const char* d_seq_bytes, const int64_t* seq_offsets,
                           const uint32_t* target_ids, int64_t n, uint32_t* out_windows)
{
    MCG_REQUIRE(n == 0 || d_seq_bytes != nullptr, "mcg_db_add_targets_dev: NULL argument");
    MCG_REQUIRE(n == 0 || ((reinterpret_cast<uintptr_t>(d_seq_bytes) + uintptr_t(seq_offsets ? seq_offsets[0] : 0)) & 15u) == 0,
                "device sequences must be 16-byte aligned");
    return add_targets(db, d_seq_bytes, seq_offsets, target_ids, n, out_windows, true);
}

int mcg_db_finalize(mcg_db* db)
{
    MCG_REQUIRE(db != nullptr, "db is NULL");
    if(db->finalized) return MCG_OK;
    return db_finalize(db);
}

int mcg_db_counts(mcg_db* db, uint64_t* nkeys, uint64_t* nkeys_nonempty, uint64_t* nlocs)
{
    MCG_REQUIRE(db != nullptr, "db is NULL");
    if(!db->finalized) { set_error("database not finalized"); return MCG_ERR_STATE; }
    return db_counts(db, nkeys, nkeys_nonempty, nlocs);
}

static int fetch_csr(mcg_db* db, std::vector<uint32_t>& keys, std::vector<uint8_t>& sizes, std::vector<uint64_t>& offs)
{
    mcg_ctx* c = db->ctx;
    keys.resize(db->nkeys); sizes.resize(db->nkeys); offs.resize(db->nkeys + 1);
    MCG_TRY(copy_out(c, keys.data(), db->keys.p, db->nkeys * 4));
    MCG_TRY(copy_out(c, sizes.data(), db->sizes.p, db->nkeys));
    MCG_TRY(copy_out(c, offs.data(), db->offs.p, (db->nkeys + 1) * 8));
    MCG_CUDA(cudaStreamSynchronize(c->stream));
    return MCG_OK;
}

// ---- export / import on the device (the host loops over tens of millions of buckets were the
// longest phases of the command-line driver)
struct TmpBuf : DevBuf {                         // scratch that goes back to the allocator on every return path
    ~TmpBuf() { if(p != nullptr) release(); }
};

__global__ void exp_flags_kernel(const uint8_t* sizes, uint32_t nkeys, uint32_t* live)
{
    const uint64_t j = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if(j < nkeys) live[j] = sizes[j] != 0;
}

// non-empty keys, compacted: key, size, and the list split into targets / windows
__global__ void exp_emit_kernel(const uint32_t* ukeys, const uint8_t* sizes, const uint64_t* offs, const uint64_t* locs, uint32_t nkeys,
                                const uint32_t* kpos, const uint64_t* lpos, uint32_t* okeys, uint32_t* osizes, uint32_t* otg, uint32_t* own)
{
    const uint64_t j = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if(j >= nkeys) return;
    const uint32_t sz = sizes[j];
    if(sz == 0) return;
    if(okeys) okeys[kpos[j]] = ukeys[j];
    if(osizes) osizes[kpos[j]] = sz;
    if(otg == nullptr) return;
    const uint64_t src = offs[j], dst = lpos[j];
    for(uint32_t t = 0; t < sz; ++t) {
        const uint64_t l = locs[src + t];
        otg[dst + t] = uint32_t(l >> 32);
        own[dst + t] = uint32_t(l);
    }
}

int mcg_db_export(mcg_db* db, uint32_t* keys, uint32_t* sizes, uint32_t* tgts, uint32_t* wins)
{
    MCG_REQUIRE(db != nullptr, "db is NULL");
    if(!db->finalized) { set_error("database not finalized"); return MCG_ERR_STATE; }
    mcg_ctx* c = db->ctx;
    MCG_CUDA(cudaSetDevice(c->device));
    const uint32_t nkeys = uint32_t(db->nkeys);
    if(nkeys == 0) return MCG_OK;
    const bool want_locs = tgts != nullptr && wins != nullptr;
    uint32_t* kpos = nullptr; uint64_t* lpos = nullptr;
    MCG_TRY(c->tmp0.reserve((size_t(nkeys) + 2) * 4));
    MCG_TRY(c->tmp4.reserve((size_t(nkeys) + 2) * 8));
    kpos = c->tmp0.as<uint32_t>();
    lpos = c->tmp4.as<uint64_t>();
    exp_flags_kernel<<<grid_for(nkeys), 256, 0, c->stream>>>(db->sizes.as<uint8_t>(), nkeys, kpos);
    c->launches++;
    MCG_TRY(exclusive_scan_u32_to_u32(c, kpos, kpos, nkeys));
    MCG_TRY(exclusive_scan_u8_to_u64(c, db->sizes.as<uint8_t>(), lpos, nkeys));
    uint32_t nlive = 0; uint64_t nl = 0;
    MCG_TRY(copy_out(c, &nlive, kpos + nkeys, 4));
    MCG_TRY(copy_out(c, &nl, lpos + nkeys, 8));
    MCG_CUDA(cudaStreamSynchronize(c->stream));
    MCG_TRY(c->tmp1.reserve((size_t(nlive) + 1) * 4));
    MCG_TRY(c->tmp2.reserve((size_t(nlive) + 1) * 4));
    TmpBuf otg, own;
    if(want_locs) { MCG_TRY(otg.reserve((nl + 1) * 4)); MCG_TRY(own.reserve((nl + 1) * 4)); }
    exp_emit_kernel<<<grid_for(nkeys), 256, 0, c->stream>>>(db->keys.as<uint32_t>(), db->sizes.as<uint8_t>(), db->offs.as<uint64_t>(),
                                                           db->locs.as<uint64_t>(), nkeys, kpos, lpos, c->tmp1.as<uint32_t>(),
                                                           c->tmp2.as<uint32_t>(), want_locs ? otg.as<uint32_t>() : nullptr,
                                                           want_locs ? own.as<uint32_t>() : nullptr);
    c->launches++;
    MCG_CUDA(cudaGetLastError());
    if(keys) MCG_TRY(copy_out(c, keys, c->tmp1.p, size_t(nlive) * 4));
    if(sizes) MCG_TRY(copy_out(c, sizes, c->tmp2.p, size_t(nlive) * 4));
    if(want_locs) { MCG_TRY(copy_out(c, tgts, otg.p, nl * 4)); MCG_TRY(copy_out(c, wins, own.p, nl * 4)); }
    MCG_CUDA(cudaStreamSynchronize(c->stream));
    return MCG_OK;
}

int mcg_db_key_sizes(mcg_db* db, uint32_t* keys, uint32_t* sizes)
{
    return mcg_db_export(db, keys, sizes, nullptr, nullptr);
}

int mcg_db_key_sizes_dev(mcg_db* db, uint32_t* d_keys, uint32_t* d_sizes)
{
    MCG_REQUIRE(db != nullptr, "db is NULL");
    if(!db->finalized) { set_error("database not finalized"); return MCG_ERR_STATE; }
    mcg_ctx* c = db->ctx;
    MCG_CUDA(cudaSetDevice(c->device));
    const uint32_t nkeys = uint32_t(db->nkeys);
    if(nkeys == 0) return MCG_OK;
    MCG_TRY(c->tmp0.reserve((size_t(nkeys) + 2) * 4));
    uint32_t* kpos = c->tmp0.as<uint32_t>();
    exp_flags_kernel<<<grid_for(nkeys), 256, 0, c->stream>>>(db->sizes.as<uint8_t>(), nkeys, kpos);
    c->launches++;
    MCG_TRY(exclusive_scan_u32_to_u32(c, kpos, kpos, nkeys));
    exp_emit_kernel<<<grid_for(nkeys), 256, 0, c->stream>>>(db->keys.as<uint32_t>(), db->sizes.as<uint8_t>(), db->offs.as<uint64_t>(),
                                                           db->locs.as<uint64_t>(), nkeys, kpos, nullptr, d_keys, d_sizes, nullptr, nullptr);
    c->launches++;
    MCG_CUDA(cudaGetLastError());
    return MCG_OK;
}

// err: bit 0 size > 255, bit 1 key 0xFFFFFFFF; mx: [0] largest target, [1] largest window
__global__ void imp_check_kernel(const uint32_t* keys, const uint32_t* sizes, uint64_t nkeys, const uint32_t* tgts, const uint32_t* wins,
                                 uint64_t nlocs, uint32_t* err, uint32_t* mx)
{
    const uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    uint32_t e = 0, mt = 0, mw = 0;
    if(i < nkeys) { if(sizes[i] > 255u) e |= 1u; if(keys[i] == EMPTY_KEY) e |= 2u; }
    if(i < nlocs) { mt = tgts[i]; mw = wins[i]; }
    for(int d = 16; d > 0; d >>= 1) {
        e |= __shfl_down_sync(0xFFFFFFFFu, e, d);
        mt = max(mt, __shfl_down_sync(0xFFFFFFFFu, mt, d));
        mw = max(mw, __shfl_down_sync(0xFFFFFFFFu, mw, d));
    }
    if((threadIdx.x & 31) == 0) {
        if(e) atomicOr(err, e);
        if(mt) atomicMax(mx, mt);
        if(mw) atomicMax(mx + 1, mw);
    }
}

// sorted position j <- bucket order[j] of the file: size cut to maxlocs (sketch_database.h:361-364), duplicate keys flagged
__global__ void imp_sizes_kernel(const uint32_t* skeys, const uint32_t* order, const uint32_t* sizes, uint32_t nkeys, uint32_t maxlocs,
                                 uint8_t* osz, uint32_t* err)
{
    const uint64_t j = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if(j >= nkeys) return;
    if(j > 0 && skeys[j] == skeys[j - 1]) atomicOr(err, 4u);
    osz[j] = uint8_t(min(sizes[order[j]], maxlocs));
}

__global__ void imp_locs_kernel(const uint32_t* order, const uint64_t* start, const uint8_t* osz, const uint64_t* offs, uint32_t nkeys,
                                const uint32_t* tgts, const uint32_t* wins, uint64_t* locs)
{
    const uint64_t j = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if(j >= nkeys) return;
    const uint64_t src = start[order[j]], dst = offs[j];
    const uint32_t sz = osz[j];
    for(uint32_t t = 0; t < sz; ++t) locs[dst + t] = (uint64_t(tgts[src + t]) << 32) | wins[src + t];
}

int mcg_db_import(mcg_db* db, const uint32_t* keys, const uint32_t* sizes,
                  const uint32_t* tgts, const uint32_t* wins, uint64_t nkeys, uint64_t nlocs)
{
    MCG_REQUIRE(db != nullptr, "db is NULL");
    MCG_REQUIRE(nkeys == 0 || (keys && sizes), "keys/sizes NULL");
    MCG_REQUIRE(nlocs == 0 || (tgts && wins), "tgts/wins NULL");
    MCG_REQUIRE(nkeys < (1ull << 32), "too many keys");
    if(db->finalized || db->nwin > 0) { set_error("import needs an empty database"); return MCG_ERR_STATE; }
    mcg_ctx* c = db->ctx;
    MCG_CUDA(cudaSetDevice(c->device));
    // the buckets of a reference-written file come in table order: everything is sorted by key on the device
    TmpBuf d_sizes, d_tg, d_wn, d_start;
    MCG_TRY(db->keys.reserve((nkeys + 1) * 4)); MCG_TRY(db->sizes.reserve(nkeys + 16)); MCG_TRY(db->offs.reserve((nkeys + 2) * 8));
    MCG_TRY(d_sizes.reserve((nkeys + 1) * 4)); MCG_TRY(d_start.reserve((nkeys + 2) * 8));
    MCG_TRY(d_tg.reserve((nlocs + 1) * 4)); MCG_TRY(d_wn.reserve((nlocs + 1) * 4));
    MCG_TRY(c->tmp0.reserve((nkeys + 1) * 4)); MCG_TRY(c->tmp1.reserve((nkeys + 1) * 4)); MCG_TRY(c->tmp2.reserve((nkeys + 1) * 4));
    MCG_TRY(c->tmp3.reserve(64));
    uint32_t* skeys = db->keys.as<uint32_t>();
    uint32_t* order = c->tmp0.as<uint32_t>();
    uint32_t* flags = c->tmp3.as<uint32_t>();                  // [0] error bits, [1] max target, [2] max window
    MCG_CUDA(cudaMemsetAsync(flags, 0, 16, c->stream));
    MCG_TRY(copy_in(c, skeys, keys, nkeys * 4));
    MCG_TRY(copy_in(c, d_sizes.p, sizes, nkeys * 4));
    MCG_TRY(copy_in(c, d_tg.p, tgts, nlocs * 4));
    MCG_TRY(copy_in(c, d_wn.p, wins, nlocs * 4));
    uint64_t total = 0, lo = 0;
    uint32_t h_flags[3] = {0, 0, 0};
    if(nkeys > 0) {
        imp_check_kernel<<<grid_for(std::max(nkeys, nlocs)), 256, 0, c->stream>>>(skeys, d_sizes.as<uint32_t>(), nkeys, d_tg.as<uint32_t>(),
                                                                                 d_wn.as<uint32_t>(), nlocs, flags, flags + 1);
        c->launches++;
        MCG_TRY(exclusive_scan_u32_to_u64(c, d_sizes.as<uint32_t>(), d_start.as<uint64_t>(), nkeys));
        MCG_TRY(copy_out(c, &total, d_start.as<uint64_t>() + nkeys, 8));
        MCG_TRY(copy_out(c, h_flags, flags, 12));
        MCG_CUDA(cudaStreamSynchronize(c->stream));
        MCG_REQUIRE((h_flags[0] & 1u) == 0, "bucket larger than 255");
        MCG_REQUIRE((h_flags[0] & 2u) == 0, "0xFFFFFFFF is not a valid feature");
        MCG_REQUIRE(total == nlocs, "sizes do not add up to nlocs");
        MCG_TRY(radix_sort_pairs(c, skeys, order, c->tmp1.as<uint32_t>(), c->tmp2.as<uint32_t>(), nkeys, 32, true));
        imp_sizes_kernel<<<grid_for(nkeys), 256, 0, c->stream>>>(skeys, order, d_sizes.as<uint32_t>(), uint32_t(nkeys), uint32_t(db->maxlocs),
                                                                db->sizes.as<uint8_t>(), flags);
        c->launches++;
        MCG_TRY(exclusive_scan_u8_to_u64(c, db->sizes.as<uint8_t>(), db->offs.as<uint64_t>(), nkeys));
        MCG_TRY(copy_out(c, &lo, db->offs.as<uint64_t>() + nkeys, 8));
        MCG_TRY(copy_out(c, h_flags, flags, 4));
        MCG_CUDA(cudaStreamSynchronize(c->stream));
        MCG_REQUIRE((h_flags[0] & 4u) == 0, "duplicate key in the imported buckets");
    } else {
        MCG_REQUIRE(nlocs == 0, "sizes do not add up to nlocs");
        MCG_CUDA(cudaMemsetAsync(db->offs.p, 0, 16, c->stream));
    }
    db->set_key_layout(h_flags[1], h_flags[2]);
    if(getenv("MCG_FORCE_KEY64")) { db->wbits = 32; db->key32 = false; }
    db->nkeys = nkeys;
    db->nlocs = lo;
    MCG_TRY(db->locs.reserve((lo + 2) * 8));
    if(nkeys > 0) {
        imp_locs_kernel<<<grid_for(nkeys), 256, 0, c->stream>>>(order, d_start.as<uint64_t>(), db->sizes.as<uint8_t>(), db->offs.as<uint64_t>(),
                                                               uint32_t(nkeys), d_tg.as<uint32_t>(), d_wn.as<uint32_t>(), db->locs.as<uint64_t>());
        c->launches++;
        MCG_CUDA(cudaGetLastError());
    }
    MCG_TRY(build_table(db));
    MCG_CUDA(cudaStreamSynchronize(c->stream));
    db->finalized = true;
    return MCG_OK;
}

int mcg_db_remove_keys(mcg_db* db, const uint32_t* keys, uint64_t n, uint64_t* removed)
{
    MCG_REQUIRE(db != nullptr, "db is NULL");
    if(!db->finalized) { set_error("database not finalized"); return MCG_ERR_STATE; }
    if(removed) *removed = 0;
    if(n == 0) return MCG_OK;
    mcg_ctx* c = db->ctx;
    MCG_CUDA(cudaSetDevice(c->device));
    MCG_TRY(c->tmp4.reserve(n * 4));
    MCG_TRY(copy_in(c, c->tmp4.p, keys, n * 4));
    return db_remove_keys_dev(db, c->tmp4.as<uint32_t>(), n, removed);
}

int mcg_remove_overpopulated(mcg_db** shards, int32_t P, uint64_t* removed)
{
    MCG_REQUIRE(shards != nullptr && P >= 1, "bad arguments");
    // mode_build.cpp:847-1074: sum of the per-rank (capped) bucket sizes per key; keys whose
    // sum exceeds maxlocs-1 are cleared on every rank
    std::vector<std::vector<uint32_t>> keys(P), sizes(P);
    for(int r = 0; r < P; ++r) {
        uint64_t nk = 0, ne = 0, nl = 0;
        MCG_TRY(mcg_db_counts(shards[r], &nk, &ne, &nl));
        keys[r].resize(ne); sizes[r].resize(ne);
        MCG_TRY(mcg_db_key_sizes(shards[r], keys[r].data(), sizes[r].data()));
    }
    const uint32_t limit = uint32_t(shards[0]->maxlocs) - 1u;
    std::vector<size_t> idx(P, 0);
    std::vector<uint32_t> rem;
    for(;;) {
        bool have = false;
        uint32_t key = 0;
        for(int r = 0; r < P; ++r)
            if(idx[r] < keys[r].size() && (!have || keys[r][idx[r]] < key)) { key = keys[r][idx[r]]; have = true; }
        if(!have) break;
        uint32_t total = 0;
        for(int r = 0; r < P; ++r)
            if(idx[r] < keys[r].size() && keys[r][idx[r]] == key) total += sizes[r][idx[r]++];
        if(total > limit) rem.push_back(key);
    }
    uint64_t tot = 0;
    for(int r = 0; r < P; ++r) {
        uint64_t x = 0;
        MCG_TRY(mcg_db_remove_keys(shards[r], rem.data(), rem.size(), &x));
        tot += x;
    }
    if(removed) *removed = tot;
    return MCG_OK;
}

}  // extern "C"
