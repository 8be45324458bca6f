// query.cu -- K3..K6: per-shard query, cross-shard merge and decision.
//
// Reference semantics reproduced:
//   sketch_database.h:806-823  accumulate_matches: for every query window, every
//                              feature found in the table contributes its whole list
//   querying.h:90-106          merge_sort: total order by (tgt,win), sketch_database.h:172-177
//   candidates.h:119-180       for_all_contiguous_window_ranges: per target the largest
//                              number of hits inside < numWindows consecutive windows
//   classification.cpp:210-226 numWindows = 2 + max(len1+len2, insertSizeMax)/target stride
//   candidates.h:236-285       sorted top list, stable, optional lift to -lowest + max-merge
//   querying.h:905-971         cross-rank merge = insert in rank order
//   classification.cpp:236-265 threshold + ranked LCA (taxonomy.h:531-537)
//   classification_statistics.h:70-78 assign
//
// Pipeline of one part of a call (all on the ctx stream, 8 launches, no host round trip):
//   q_prep_kernel (window counts, prefixes, per-read meta: one single-pass chained scan) -> K1 of mate 1 and mate 2
//   -> [sharded: exchange of sketches + meta, comm.cu] -> q_fused_warp_kernel: one warp per read (pair) probes the
//   bucket table (one 128-byte line per probe, common.cuh "feature table"), takes short lists out of the slot itself,
//   gathers longer ones, sorts, derives range hits and the top list -- all in registers / shared memory
//   -> q_spill_warp_kernel: the few short reads with more matches than a warp holds in shared memory (global scratch)
//   -> [sharded: all-to-all of the per-read records] -> q_merge_classify_kernel.
//   Long reads (> 128 feature slots) take the staged path: find -> scan of list lengths -> read binning -> gather ->
//   size-binned sorters (warp / CTA / big CTA / global-memory fallback) with the same window-range + top-list logic;
//   when only a few reads are left over (< 1/8 of the part) their windows and features are first compacted into a
//   sub-batch of their own (q_left_*).  The host learns ONE number per part (how many long reads are left) and reads it
//   after the next part's stage A is enqueued.  Host entry points cut a call into parts and overlap H2D, kernels and D2H;
//   mcg_query_batch_sharded runs the whole of query_batched_parallel2 (querying.h:721-1173) across ranks.
#include "common.cuh"

#include <ctime>

namespace mcg {

int copy_in(mcg_ctx* c, void* d, const void* h, size_t bytes);
int copy_out(mcg_ctx* c, void* h, const void* d, size_t bytes);
int exclusive_scan_u32_to_u32(mcg_ctx* c, const uint32_t* d_in, uint32_t* d_out, uint64_t n);
int ensure_pipeline(mcg_ctx* c);
// comm.cu
int comm_rank(const mcg_comm* m);
int comm_size(const mcg_comm* m);
int comm_allgather_dev_u64(mcg_comm* m, const uint64_t* d_mine, int count, uint64_t* out);
int comm_allgatherv_inplace(mcg_comm* m, void* buf, const uint64_t* off, const uint64_t* len, cudaStream_t st);
int comm_alltoallv(mcg_comm* m, const void* send, const uint64_t* soff, const uint64_t* slen,
                   void* recv, const uint64_t* roff, const uint64_t* rlen, cudaStream_t st);
int merge_classify_rows(mcg_ctx* c, const mcg_query_opts* o, const int32_t* d_cand_tax, const uint32_t* d_cand_hits, const int32_t* d_cand_n,
                        uint32_t rs, uint32_t ns, int32_t nshards, int64_t n,
                        int32_t* d_out_tax, uint32_t* d_out_hits, int32_t* d_out_ncand,
                        int32_t* d_out_best, int32_t* d_out_best_rank, uint64_t* d_stats22);
int comm_group_begin(mcg_comm* m);
int comm_group_end(mcg_comm* m);

constexpr uint32_t EMPTY_KEY = 0xFFFFFFFFu;
constexpr int Q_SMALL = 512;       // matches per read handled by one warp
constexpr int Q_MED = 4096;        // by one 256-thread CTA
constexpr int Q_LARGE = 16384;     // by one 512-thread CTA with 192 KB of shared memory

static inline unsigned grid_for(uint64_t n, int threads = 256) { return unsigned((n + threads - 1) / threads); }

// ------------------------------------------------------------------ window bookkeeping
// ------------------------------------------------------------------ window bookkeeping, one pass
// Per read: windows of mate 1 / mate 2 (dna_encoding.h:261-276) and meta = {windows of the read (pair),
// window-range width numWindows = 2 + max(len1 + len2, insertsize) / target stride (classification.cpp:217-219)};
// over all reads: the exclusive prefixes wf1 / wf2 (processing order of the two K1 launches), W = wf1 + wf2
// (read-major output order) and of2 = W + windows of mate 1 (where mate 2's sketches go) -- one kernel, a
// single-pass chained scan (decoupled look-back): tiles are taken in order from a counter, a tile publishes its
// aggregate, then walks back over its predecessors until it meets an inclusive prefix.  The two running sums
// (< 2^31 each) and the 2 status bits share one 64-bit word, so a status is read and written atomically.
constexpr int QP_THREADS = 256, QP_ITEMS = 4, QP_TILE = QP_THREADS * QP_ITEMS;
constexpr unsigned long long QP_AGG = 1ull << 62, QP_INC = 2ull << 62, QP_VAL = (1ull << 62) - 1;

struct PrepParams {
    const int64_t* offs1;
    const int64_t* offs2;
    int64_t n;
    int64_t winlen, stride, insertsize, tstride;
    uint32_t *wf1, *wf2, *W, *of2;         // [n + 1] each (of2: [n]); wf2 / of2 may be NULL (single-end)
    uint2* meta;                           // [n]
    unsigned long long* status;            // [tiles], zeroed; followed by the tile counter
    uint32_t* tile_ctr;
    uint32_t* clear8;                      // 16 bytes zeroed on the way (the counters of the fused / spill kernels); may be NULL
};

__global__ void __launch_bounds__(QP_THREADS)
q_prep_kernel(const PrepParams p)
{
    __shared__ unsigned long long s_warp[QP_THREADS / 32];
    __shared__ unsigned long long s_base;
    __shared__ uint32_t s_tile;
    if(threadIdx.x == 0) s_tile = atomicAdd(p.tile_ctr, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    if(tile == 0 && threadIdx.x == 0 && p.clear8 != nullptr) { p.clear8[0] = 0u; p.clear8[1] = 0u; p.clear8[2] = 0u; p.clear8[3] = 0u; }
    const int64_t r0 = int64_t(tile) * QP_TILE + int64_t(threadIdx.x) * QP_ITEMS;
    uint32_t a[QP_ITEMS], b[QP_ITEMS];
    unsigned long long mine = 0;
#pragma unroll
    for(int i = 0; i < QP_ITEMS; ++i) {
        const int64_t r = r0 + i;
        a[i] = 0; b[i] = 0;
        if(r < p.n) {
            const int64_t l1 = p.offs1[r + 1] - p.offs1[r], l2 = p.offs2 ? p.offs2[r + 1] - p.offs2[r] : 0;
            a[i] = uint32_t(num_windows(l1, p.winlen, p.stride));
            b[i] = p.offs2 ? uint32_t(num_windows(l2, p.winlen, p.stride)) : 0u;
            int64_t span = l1 + l2;
            if(p.insertsize > span) span = p.insertsize;
            // 32-bit division when possible (the 64-bit one is emulated with ~150 instructions)
            const uint64_t nwr = (span < (int64_t(1) << 31) && p.tstride < (int64_t(1) << 31)) ? 2u + uint32_t(span) / uint32_t(p.tstride)
                                                                                              : uint64_t(2 + span / p.tstride);
            p.meta[r] = make_uint2(a[i] + b[i], uint32_t(nwr < 0xFFFFFFFFull ? nwr : 0xFFFFFFFFull));
        }
        mine += (unsigned long long)a[i] | ((unsigned long long)b[i] << 31);
    }
    // block-exclusive scan of the packed pair
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long inc = mine;
#pragma unroll
    for(int d = 1; d < 32; d <<= 1) {
        const unsigned long long o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if(lane >= d) inc += o;
    }
    if(lane == 31) s_warp[warp] = inc;
    __syncthreads();
    unsigned long long wbase = 0, agg = 0;
#pragma unroll
    for(int w = 0; w < QP_THREADS / 32; ++w) { const unsigned long long x = s_warp[w]; if(w < warp) wbase += x; agg += x; }
    if(threadIdx.x == 0) {
        volatile unsigned long long* st = p.status;
        unsigned long long base = 0;
        if(tile == 0) {
            st[0] = QP_INC | agg;
        } else {
            st[tile] = QP_AGG | agg;
            __threadfence();
            for(int64_t q = int64_t(tile) - 1; q >= 0; --q) {
                unsigned long long v;
                do { v = st[q]; } while((v >> 62) == 0ull);
                base += v & QP_VAL;
                if((v >> 62) == 2ull) break;
            }
            st[tile] = QP_INC | ((base + agg) & QP_VAL);
        }
        s_base = base;
    }
    __syncthreads();
    unsigned long long ex = s_base + wbase + inc - mine;
#pragma unroll
    for(int i = 0; i < QP_ITEMS; ++i) {
        const int64_t r = r0 + i;
        const uint32_t e1 = uint32_t(ex & 0x7FFFFFFFull), e2 = uint32_t((ex >> 31) & 0x7FFFFFFFull);
        if(r <= p.n) {                                           // index n receives the totals
            p.wf1[r] = e1;
            if(p.wf2) p.wf2[r] = e2;
            p.W[r] = e1 + e2;
            if(p.of2 && r < p.n) p.of2[r] = e1 + e2 + a[i];
        }
        ex += (unsigned long long)a[i] | ((unsigned long long)b[i] << 31);
    }
}

__global__ void q_meta_windows_kernel(const uint2* meta, int64_t n, uint32_t* nws)
{
    const int64_t r = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if(r < n) nws[r] = meta[r].x;
}

// ------------------------------------------------------------------ K3: find
// hash_multimap.h:779-789 / 1034-1047 find, on the bucketized table (common.cuh): walk the buckets
// from the key's home bucket until the key or a free slot shows up
__global__ void __launch_bounds__(256)
q_find_kernel(const uint32_t* __restrict__ feat, uint64_t nslots, const uint32_t* __restrict__ table, const TableGeom g,
              uint8_t* __restrict__ flen, uint64_t* __restrict__ foff)
{
    const uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if(i >= nslots) return;
    const uint32_t f = feat[i];
    uint32_t len = 0;
    uint64_t off = 0;
    if(f != EMPTY_KEY) {
        uint32_t b = table_bucket(f, g);
        for(;;) {
            const uint32_t* line = table + uint64_t(b) * TABLE_BUCKET_WORDS;
            const uint4 k4 = __ldg(reinterpret_cast<const uint4*>(line));
            const int t = table_match(k4, f);
            if(t >= 0) {
                const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(line + 8 + 8 * t));
                len = q0.x;
                off = q0.y;
                break;
            }
            if(t == -1) break;
            ++b;
        }
    }
    flen[i] = uint8_t(len);
    foff[i] = off;
}

// ------------------------------------------------------------------ read binning
// counters: [0] small [1] medium [2] large [3] xl
__global__ void q_bin_reads_kernel(const uint64_t* fpos, const uint32_t* W, int s, int64_t n,
                                   uint32_t* lists /* [4][n] */, uint32_t* counters,
                                   int32_t* out_n, int32_t* out_tax, uint32_t* out_hits, int K, uint32_t rs, uint32_t ns,
                                   const uint8_t* done, const uint32_t* rmap)
{
    const int64_t r = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    int cls = 4;                                          // 4 = nothing to do
    if(r < n && !(done != nullptr && done[r])) {
        const uint64_t m = fpos[uint64_t(W[r + 1]) * s] - fpos[uint64_t(W[r]) * s];
        if(m == 0) {
            const uint64_t o = rmap ? rmap[r] : uint32_t(r);
            out_n[o * ns] = 0;
            for(int c = 0; c < K; ++c) { out_tax[o * rs + c] = -1; out_hits[o * rs + c] = 0u; }
        } else cls = m <= Q_SMALL ? 0 : (m <= Q_MED ? 1 : (m <= Q_LARGE ? 2 : 3));
    }
    // one atomic per class per warp
    const int lane = threadIdx.x & 31;
    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, cls);
    if(cls < 4) {
        const int leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if(lane == leader) base = atomicAdd(&counters[cls], uint32_t(__popc(peers)));
        base = __shfl_sync(peers, base, leader);
        lists[uint64_t(cls) * n + base + __popc(peers & ((1u << lane) - 1u))] = uint32_t(r);
    }
}

// ------------------------------------------------------------------ gather
template<class KeyT>
__device__ __forceinline__ KeyT make_key(uint64_t loc, int wbits)
{
    return KeyT((KeyT(loc >> 32) << wbits) | KeyT(uint32_t(loc)));
}

template<class KeyT>
__global__ void __launch_bounds__(256)
q_gather_kernel(const uint8_t* __restrict__ flen, const uint64_t* __restrict__ foff, const uint64_t* __restrict__ fpos,
                uint64_t nslots, const uint64_t* __restrict__ locs, KeyT* __restrict__ out, int wbits)
{
    const int lane = threadIdx.x & 31;
    const uint64_t slot = (blockIdx.x * uint64_t(blockDim.x) + threadIdx.x);
    uint32_t len = 0;
    uint64_t off = 0, pos = 0;
    if(slot < nslots) {
        len = flen[slot];
        if(len) { off = foff[slot]; pos = fpos[slot]; }
    }
    if(len > 0 && len <= 2) {
        out[pos] = make_key<KeyT>(locs[off], wbits);
        if(len == 2) out[pos + 1] = make_key<KeyT>(locs[off + 1], wbits);
    }
    uint32_t big = __ballot_sync(0xFFFFFFFFu, len > 2);
    while(big) {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        const uint32_t L = __shfl_sync(0xFFFFFFFFu, len, src);
        const uint64_t O = __shfl_sync(0xFFFFFFFFu, off, src);
        const uint64_t P = __shfl_sync(0xFFFFFFFFu, pos, src);
        for(uint32_t t = lane; t < L; t += 32) out[P + t] = make_key<KeyT>(locs[O + t], wbits);
    }
}

// ------------------------------------------------------------------ top list (candidates.h:236-285)
struct TopK {
    int32_t tax[MCG_MAX_CAND];
    uint32_t hits[MCG_MAX_CAND];
    int n;
};

__device__ __forceinline__ void topk_insert(TopK& L, int K, int32_t tax, uint32_t hits, int lowest,
                                            const int32_t* __restrict__ lineage, const uint8_t* __restrict__ rank)
{
    if(tax < 0) return;
    if(lowest > MCG_RANK_SEQUENCE) {
        const int32_t anc = lineage[int64_t(tax) * MCG_NUM_RANKS + lowest];
        if(anc >= 0) tax = anc;
    }
    int pos = -1;
    if(rank[tax] != MCG_RANK_SEQUENCE) {
        for(int i = 0; i < L.n; ++i) if(L.tax[i] == tax) { pos = i; break; }
    }
    if(pos >= 0) {
        // taxon already listed: keep the larger hit count, restore order (stable)
        uint32_t h = L.hits[pos];
        if(hits > h) h = hits;
        int j = pos;
        while(j > 0 && L.hits[j - 1] < h) { L.tax[j] = L.tax[j - 1]; L.hits[j] = L.hits[j - 1]; --j; }
        L.tax[j] = tax;
        L.hits[j] = h;
        return;
    }
    int i = 0;
    while(i < L.n && L.hits[i] >= hits) ++i;         // upper_bound in a descending list
    if(i < L.n || L.n < K) {
        const int last = L.n < K ? L.n : K - 1;
        for(int j = last; j > i; --j) { L.tax[j] = L.tax[j - 1]; L.hits[j] = L.hits[j - 1]; }
        L.tax[i] = tax;
        L.hits[i] = hits;
        if(L.n < K) L.n++;
    }
}

// the same list held one entry per lane (entry i in lane i): insertion is a handful of ballots and
// one shuffle instead of loops over a local-memory array (warp_topk_insert_lifted below).
struct WarpTopK {
    int32_t tax;        // lane i: entry i (valid for i < n)
    uint32_t hits;
    int n;              // uniform
};

// ------------------------------------------------------------------ K4 + K5
// Match keys: (tgt << wbits) | win.  When target and window ids of the shard fit 32 bits together
// (e.g. 5000 targets x 35k windows) the whole per-read stage runs on 32-bit keys, which halves the
// sorting work and the gathered bytes; otherwise wbits = 32 and keys are the stored 64-bit locations.
struct CandParams {
    const uint64_t* fpos;
    const uint32_t* W;
    int s;
    const uint2* meta;           // per read {windows, numWindows of the candidate ranges}
    void* loc;
    int wbits;
    int kbits;                   // significant bits of a match key
    uint32_t* hits_scratch;      // only for the global-memory class: range hits, candidate scratch
    uint32_t* hits_scratch2;
    const uint32_t* list;
    const uint32_t* list_count;
    const uint32_t* rmap;        // staged path over a compacted sub-batch: position -> read of the batch (nullptr: identity)
    int K;
    int lowest;
    const int32_t* lineage;
    const uint8_t* rank;
    const int32_t* lift;         // [ntargets] taxon a target's candidates are listed under at -lowest (candidates.h:242-245:
                                 // its ancestor at that rank if it has one, else its own sequence-level taxon), bit 31 = that
                                 // taxon is of sequence rank: 4 bytes per target instead of two dependent loads from the lineage table
    int32_t* out_tax;
    uint32_t* out_hits;
    int32_t* out_n;
    uint32_t rs, ns;             // row strides in words of out_tax / out_hits (K when the arrays are separate) and of out_n (1);
                                 // a packed record per read {tax[K], hits[K], n} has rs = ns = 2K + 1 (one exchange instead of three)
    int write_sorted;
};

template<bool WARP>
__device__ __forceinline__ void gsync()
{
    if(WARP) __syncwarp(); else __syncthreads();
}

template<class KeyT> __device__ __forceinline__ uint32_t tgt_of(KeyT k, int wbits) { return uint32_t(k >> wbits); }

// window-range width of a read (classification.cpp:217-219), precomputed by q_prep_kernel
__device__ __forceinline__ uint32_t num_windows_in_range(const CandParams& p, uint32_t r)
{
    return __ldg(&p.meta[r].y);
}

// hits of the best window range ENDING at element i (candidates.h:141-152): number of
// elements of the same target with win > win_i - numwin, found by binary search in [0, i]
template<class KeyT>
__device__ __forceinline__ uint32_t hits_ending_at(const KeyT* A, uint32_t i, uint32_t numwin, int wbits)
{
    const KeyT key = A[i];
    const KeyT wmask = (KeyT(1) << wbits) - 1;
    const KeyT win = key & wmask;
    const KeyT lowkey = (key & ~wmask) | (win >= KeyT(numwin - 1) ? win - KeyT(numwin - 1) : KeyT(0));
    // most matches of a read against a large database are chance hits on targets of their own (runs of one element):
    // look at the neighbour first, then gallop backwards, then bisect what is left
    if(i == 0 || A[i - 1] < lowkey) return 1u;
    uint32_t lo = 0, hi = i - 1, step = 1;                 // A[hi] >= lowkey
    while(hi >= step) {
        if(A[hi - step] < lowkey) { lo = hi - step + 1; break; }
        hi -= step;
        step <<= 1;
    }
    while(lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if(A[mid] < lowkey) lo = mid + 1; else hi = mid;
    }
    return i - lo + 1;
}

// ascending bitonic network in the "flip" formulation: every comparator is (lo < hi,
// ascending), so elements beyond m behave as +infinity without being stored
template<bool WARP, class KeyT>
__device__ __forceinline__ void bitonic_sort(KeyT* A, uint32_t m, int tid, int nt)
{
    if(m < 2) return;
    const uint32_t lgP = 32u - uint32_t(__clz(int(m - 1)));      // P = 2^lgP >= m
    const uint32_t half = 1u << (lgP - 1);
    for(uint32_t lk = 1; lk <= lgP; ++lk) {                       // merge size k = 2^lk
        const uint32_t k = 1u << lk, hk = k >> 1;
        for(uint32_t t = tid; t < half; t += nt) {
            const uint32_t blk = t >> (lk - 1), o = t & (hk - 1);
            const uint32_t lo = (blk << lk) + o, hi = (blk << lk) + k - 1 - o;
            if(hi < m) {
                const KeyT a = A[lo], b = A[hi];
                if(a > b) { A[lo] = b; A[hi] = a; }
            }
        }
        gsync<WARP>();
        for(uint32_t lj = lk - 1; lj-- > 0;) {                    // j = 2^lj
            const uint32_t j = 1u << lj;
            for(uint32_t t = tid; t < half; t += nt) {
                const uint32_t lo = ((t >> lj) << (lj + 1)) + (t & (j - 1)), hi = lo + j;
                if(hi < m) {
                    const KeyT a = A[lo], b = A[hi];
                    if(a > b) { A[lo] = b; A[hi] = a; }
                }
            }
            gsync<WARP>();
        }
    }
}

// ---- register sorting networks of one warp (flip formulation, ascending)
template<class KeyT>
__device__ __forceinline__ KeyT wsort32(KeyT v, int lane)
{
#pragma unroll
    for(int k = 2; k <= 32; k <<= 1) {
        {
            const KeyT o = __shfl_xor_sync(0xFFFFFFFFu, v, k - 1);
            v = (lane & (k >> 1)) ? max(v, o) : min(v, o);
        }
#pragma unroll
        for(int j = k >> 2; j > 0; j >>= 1) {
            const KeyT o = __shfl_xor_sync(0xFFFFFFFFu, v, j);
            v = (lane & j) ? max(v, o) : min(v, o);
        }
    }
    return v;
}

// 64 keys: element e = lane (a) and lane + 32 (b)
template<class KeyT>
__device__ __forceinline__ void wsort64(KeyT& a, KeyT& b, int lane)
{
    a = wsort32(a, lane);
    b = wsort32(b, lane);
    // merge of size 64: flip step pairs e with 63 - e, i.e. a of lane L with b of lane 31 - L
    {
        const KeyT ob = __shfl_sync(0xFFFFFFFFu, b, 31 - lane);
        const KeyT oa = __shfl_sync(0xFFFFFFFFu, a, 31 - lane);
        a = min(a, ob);
        b = max(b, oa);
    }
    // distance 32 is inside the flip step's halves already sorted apart; remaining distances < 32
#pragma unroll
    for(int j = 16; j > 0; j >>= 1) {
        const KeyT oa = __shfl_xor_sync(0xFFFFFFFFu, a, j);
        const KeyT ob = __shfl_xor_sync(0xFFFFFFFFu, b, j);
        a = (lane & j) ? max(a, oa) : min(a, oa);
        b = (lane & j) ? max(b, ob) : min(b, ob);
    }
}

// 96 keys (element e = 32 * i + lane in a, b, c): the 128-key network with its last quarter known to be +infinity --
// the merge's flip step pairs the first quarter with infinity (no-op), the 32-distance step of the upper half too
template<class KeyT>
__device__ __forceinline__ void wsort96(KeyT& a, KeyT& b, KeyT& c, int lane)
{
    wsort64(a, b, lane);
    c = wsort32(c, lane);
    {
        const KeyT oc = __shfl_sync(0xFFFFFFFFu, c, 31 - lane), ob = __shfl_sync(0xFFFFFFFFu, b, 31 - lane);
        b = min(b, oc);
        c = max(c, ob);
    }
    {
        const KeyT lo = min(a, b), hi = max(a, b);
        a = lo; b = hi;
    }
#pragma unroll
    for(int j = 16; j > 0; j >>= 1) {
        const KeyT oa = __shfl_xor_sync(0xFFFFFFFFu, a, j), ob = __shfl_xor_sync(0xFFFFFFFFu, b, j), oc = __shfl_xor_sync(0xFFFFFFFFu, c, j);
        a = (lane & j) ? max(a, oa) : min(a, oa);
        b = (lane & j) ? max(b, ob) : min(b, ob);
        c = (lane & j) ? max(c, oc) : min(c, oc);
    }
}

// 128 keys: element e = 32 * i + lane lives in v[i]
template<class KeyT>
__device__ __forceinline__ void wsort128(KeyT (&v)[4], int lane)
{
    wsort64(v[0], v[1], lane);
    wsort64(v[2], v[3], lane);
    {   // flip step of merge size 128: e with 127 - e, i.e. v[i] of lane L with v[3 - i] of lane 31 - L
        const KeyT o0 = __shfl_sync(0xFFFFFFFFu, v[0], 31 - lane), o1 = __shfl_sync(0xFFFFFFFFu, v[1], 31 - lane);
        const KeyT o2 = __shfl_sync(0xFFFFFFFFu, v[2], 31 - lane), o3 = __shfl_sync(0xFFFFFFFFu, v[3], 31 - lane);
        v[0] = min(v[0], o3); v[1] = min(v[1], o2); v[2] = max(v[2], o1); v[3] = max(v[3], o0);
    }
    {   // distance 32: inside the lane
        const KeyT a = min(v[0], v[1]), b = max(v[0], v[1]), c = min(v[2], v[3]), d = max(v[2], v[3]);
        v[0] = a; v[1] = b; v[2] = c; v[3] = d;
    }
#pragma unroll
    for(int j = 16; j > 0; j >>= 1) {
#pragma unroll
        for(int i = 0; i < 4; ++i) {
            const KeyT o = __shfl_xor_sync(0xFFFFFFFFu, v[i], j);
            v[i] = (lane & j) ? max(v[i], o) : min(v[i], o);
        }
    }
}

template<class KeyT>
__device__ __forceinline__ void toplist_from_sorted(const CandParams& p, uint32_t r, const KeyT* A, const uint32_t* Hh,
                                                    uint32_t m, int32_t* Ct, uint32_t* Ch, int lane, uint32_t numwin = 0);

// one read handled by one warp (<= Q_SMALL matches); Ct, Ch: candidate scratch of Q_SMALL entries
template<class KeyT>
__device__ __forceinline__ void process_read_warp(const CandParams& p, uint32_t r, KeyT* A, uint32_t* Hh, uint32_t m,
                                                  KeyT* gsrc, int lane, int32_t* Ct, uint32_t* Ch)
{
    const KeyT KMAX = ~KeyT(0);
    if(m <= 32) {
        KeyT a = lane < int(m) ? gsrc[lane] : KMAX;
        a = wsort32(a, lane);
        A[lane] = a;
    } else if(m <= 64) {
        KeyT a = gsrc[lane];
        KeyT b = lane + 32 < int(m) ? gsrc[lane + 32] : KMAX;
        wsort64(a, b, lane);
        A[lane] = a;
        A[lane + 32] = b;
    } else {
        for(uint32_t i = lane; i < m; i += 32) A[i] = gsrc[i];
        __syncwarp();
        bitonic_sort<true>(A, m, lane, 32);
    }
    __syncwarp();
    if(p.write_sorted) for(uint32_t i = lane; i < m; i += 32) gsrc[i] = A[i];
    const uint32_t numwin = num_windows_in_range(p, r);
    for(uint32_t i = lane; i < m; i += 32) Hh[i] = hits_ending_at(A, i, numwin, p.wbits);
    __syncwarp();
    toplist_from_sorted<KeyT>(p, r, A, Hh, m, Ct, Ch, lane);
}

__device__ __forceinline__ void warp_topk_insert_lifted(WarpTopK& L, int K, int32_t tax, uint32_t hits, bool seq, int lane);

// All candidates of a read (Ct: lifted taxon | sequence-rank flag in bit 31, Ch: hits; ascending
// target order) -> top list.  The insertion itself is order-dependent (candidates.h:236-285) and
// runs one candidate at a time, but a candidate with hits <= the hits of entry K-1 of a FULL list
// changes nothing in any branch of insert() (a new taxon would land behind the list, a known one
// only replaces on MORE hits), and that bound never decreases: such candidates are skipped, 32 at a
// time, by a ballot.  In a large database most candidates are chance hits of 1..2 and never get in.
__device__ __forceinline__ void warp_toplist_insert_all(WarpTopK& L, int K, const int32_t* Ct, const uint32_t* Ch, uint32_t ncand, int lane)
{
    for(uint32_t base = 0; base < ncand; base += 32) {
        const uint32_t j = base + uint32_t(lane);
        const int32_t myt = j < ncand ? Ct[j] : 0;
        const uint32_t myh = j < ncand ? Ch[j] : 0u;          // hits of a real candidate are >= 1
        uint32_t rest = 0xFFFFFFFFu;
        for(;;) {
            const uint32_t bound = L.n == K ? __shfl_sync(0xFFFFFFFFu, L.hits, K - 1) : 0u;
            const uint32_t pass = __ballot_sync(0xFFFFFFFFu, myh > bound) & rest;
            if(pass == 0u) break;
            const int src = __ffs(int(pass)) - 1;
            const int32_t t = __shfl_sync(0xFFFFFFFFu, myt, src);
            const uint32_t h = __shfl_sync(0xFFFFFFFFu, myh, src);
            warp_topk_insert_lifted(L, K, t & 0x7FFFFFFF, h, t < 0, lane);
            rest = src == 31 ? 0u : ~((2u << src) - 1u);
        }
    }
}

// The same list for at most 32 candidates (lane j holds candidate j: t = lifted taxon | sequence-rank flag, h = hits, 0 =
// none), without replaying the insertions: insert() (candidates.h:236-285) keeps per taxon the MAXIMUM of its candidates'
// hits, places an entry behind every entry with at least as many hits at the moment it reaches its value, and an entry
// pushed out of a full list can never return with the same value -- so the final list is the first K taxa ordered by
// (hits descending, position of the first candidate that carries the taxon's maximum ascending).  That candidate is the
// warp-wide maximum of (hits, 31 - lane); it is taken K times, each time retiring all candidates of its taxon
// (sequence-rank candidates carry distinct values, so they retire alone).  ~7 instructions per list entry.
__device__ __forceinline__ void warp_toplist_select32(WarpTopK& L, int K, int32_t t, uint32_t h, int lane)
{
    uint32_t key = h ? ((h << 5) | uint32_t(31 - lane)) : 0u;          // h < 2^27: a read has far fewer features
    for(int k = 0; k < K; ++k) {
        const uint32_t best = __reduce_max_sync(0xFFFFFFFFu, key);
        if(best == 0u) break;
        const int32_t bt = __shfl_sync(0xFFFFFFFFu, t, 31 - int(best & 31u));
        if(lane == k) { L.tax = bt & 0x7FFFFFFF; L.hits = best >> 5; }
        if(t == bt) key = 0u;
        L.n = k + 1;
    }
}

// Stable LSD radix sort of m 32-bit keys in shared memory by a whole CTA, 8 bits per pass over the
// `bits` significant bits: A -> B -> A ...; the result ends in A.  Every warp owns a contiguous chunk;
// pass = count (shared atomics on warp-private digit counters) -> prefix over (digit, warp) -> rank
// (peer masks from ballots, as in radix_sort.cu) and move.  ~3.4x fewer instructions than the bitonic
// network for the 1-2 k matches of a 10 kbp read (ncu: the network was 66 % of q_cand_cta_kernel).
// wcnt: [warps][256] counters; blockDim.x >= 256.
__device__ __forceinline__ void cta_radix_sort(uint32_t* A, uint32_t* B, uint32_t m, int bits, int tid, int nt, uint32_t* wcnt)
{
    __shared__ uint32_t s_dsum[8];
    const int lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t chunk = ((m + uint32_t(nwarp) - 1u) / uint32_t(nwarp) + 31u) & ~31u;
    const uint32_t w0 = min(m, uint32_t(warp) * chunk), w1 = min(m, w0 + chunk);
    uint32_t* wc = wcnt + warp * 256;
    uint32_t* src = A;
    uint32_t* dst = B;
    for(int shift = 0; shift < bits; shift += 8) {
        for(int i = tid; i < nwarp * 256; i += nt) wcnt[i] = 0;
        __syncthreads();
        for(uint32_t i = w0 + lane; i < w1; i += 32) atomicAdd(&wc[(src[i] >> shift) & 255u], 1u);
        __syncthreads();
        uint32_t run = 0, inc = 0;
        if(tid < 256) {                                       // thread d owns digit d
            for(int w = 0; w < nwarp; ++w) { const uint32_t cnt = wcnt[w * 256 + tid]; wcnt[w * 256 + tid] = run; run += cnt; }
            inc = run;
#pragma unroll
            for(int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
                if(lane >= d) inc += o;
            }
            if(lane == 31) s_dsum[warp] = inc;
        }
        __syncthreads();
        if(tid < 256) {
            uint32_t start = inc - run;
            for(int w = 0; w < warp; ++w) start += s_dsum[w];
            for(int w = 0; w < nwarp; ++w) wcnt[w * 256 + tid] += start;   // running position of (warp, digit)
        }
        __syncthreads();
        for(uint32_t i0 = w0; i0 < w1; i0 += 32) {
            const uint32_t i = i0 + lane;
            const bool valid = i < w1;
            const uint32_t key = valid ? src[i] : 0u;
            const uint32_t d = (key >> shift) & 255u;
            uint32_t peers = __ballot_sync(0xFFFFFFFFu, valid);
#pragma unroll
            for(int b = 0; b < 8; ++b) {
                const bool bit = (d >> b) & 1u;
                const uint32_t bal = __ballot_sync(0xFFFFFFFFu, bit);
                peers &= bit ? bal : ~bal;
            }
            const uint32_t before = wc[d];
            __syncwarp();
            if(valid && (peers & lt) == 0u) wc[d] = before + __popc(peers);
            __syncwarp();
            if(valid) dst[before + __popc(peers & lt)] = key;
        }
        __syncthreads();
        uint32_t* t = src; src = dst; dst = t;
    }
    if(src != A) {
        for(uint32_t i = tid; i < m; i += nt) A[i] = src[i];
        __syncthreads();
    }
}

// one read handled by a whole CTA: A = its matches (shared or global memory), Hh = m uint32,
// Ct = m int32 of candidate scratch (may alias A: candidate j is written after element j was read)
template<class KeyT>
__device__ __forceinline__ void process_read_cta(const CandParams& p, uint32_t r, KeyT* A, uint32_t* Hh, uint32_t m,
                                                 KeyT* gsrc, bool staged, int tid, int nt, int32_t* Ct, uint32_t* wcnt = nullptr)
{
    if(staged) {
        for(uint32_t i = tid; i < m; i += nt) A[i] = gsrc[i];
        __syncthreads();
    }
    if(sizeof(KeyT) == 4 && wcnt != nullptr && m > 256u)
        cta_radix_sort(reinterpret_cast<uint32_t*>(A), Hh, m, p.kbits, tid, nt, wcnt);     // Hh is free until the range hits are written
    else
        bitonic_sort<false>(A, m, tid, nt);
    if(staged && p.write_sorted) for(uint32_t i = tid; i < m; i += nt) gsrc[i] = A[i];
    const uint32_t numwin = num_windows_in_range(p, r);
    for(uint32_t i = tid; i < m; i += nt) Hh[i] = hits_ending_at(A, i, numwin, p.wbits);
    __syncthreads();
    // the last element of each target run collects the run's maximum
    for(uint32_t i = tid; i < m; i += nt) {
        const uint32_t tgt = tgt_of(A[i], p.wbits);
        uint32_t lo = i, hi = m;                     // first index whose target is larger
        while(lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if(tgt_of(A[mid], p.wbits) <= tgt) lo = mid + 1; else hi = mid;
        }
        const uint32_t tail = lo - 1;
        if(tail != i) atomicMax(&Hh[tail], Hh[i]);
    }
    __syncthreads();
    // the run tails are the candidates, in ascending target order (candidates.h:207-219).  Their
    // taxa are lifted by all threads at once (the lineage loads are in flight together) and they
    // are compacted in order, nt elements at a time; only the order-dependent list insertion is
    // sequential (warp 0, out of the compacted arrays, no memory latency).
    __shared__ uint32_t s_wcount[32];
    const int lane = tid & 31, warp = tid >> 5, nwarp = (nt + 31) >> 5;
    uint32_t* Ch = Hh;
    uint32_t ncand = 0;
    for(uint32_t base = 0; base < m; base += nt) {
        const uint32_t i = base + tid;
        const bool valid = i < m;
        const uint32_t tgt = valid ? tgt_of(A[i], p.wbits) : 0xFFFFFFFFu;
        const bool tail = valid && (i + 1 == m || tgt_of(A[i + 1], p.wbits) != tgt);
        uint32_t v = 0;
        int32_t tax = 0;
        if(tail) {
            v = Hh[i];
            tax = __ldg(p.lift + tgt);            // the target's taxon at -lowest, sequence-rank flag in bit 31
        }
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, tail);
        if(lane == 0) s_wcount[warp] = __popc(bal);
        __syncthreads();
        uint32_t pre = 0, tot = 0;
        for(int w = 0; w < nwarp; ++w) { const uint32_t cw = s_wcount[w]; if(w < warp) pre += cw; tot += cw; }
        if(tail) {
            const uint32_t j = ncand + pre + __popc(bal & ((1u << lane) - 1u));
            Ct[j] = tax;
            Ch[j] = v;
        }
        ncand += tot;
        __syncthreads();
    }
    if(warp == 0) {
        WarpTopK L;
        L.tax = -1; L.hits = 0; L.n = 0;
        warp_toplist_insert_all(L, p.K, Ct, Ch, ncand, lane);
        if(lane < p.K) {
            p.out_tax[uint64_t(r) * p.rs + lane] = lane < L.n ? L.tax : -1;
            p.out_hits[uint64_t(r) * p.rs + lane] = lane < L.n ? L.hits : 0u;
        }
        if(lane == 0) p.out_n[uint64_t(r) * p.ns] = L.n;
    }
    __syncthreads();
}

// small reads: one warp per read
template<class KeyT>
__global__ void __launch_bounds__(256)
q_cand_warp_kernel(const CandParams p)
{
    extern __shared__ __align__(16) unsigned char sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* wbase = sm + size_t(warp) * Q_SMALL * (sizeof(KeyT) + 12);
    KeyT* A = reinterpret_cast<KeyT*>(wbase);
    uint32_t* Hh = reinterpret_cast<uint32_t*>(wbase + size_t(Q_SMALL) * sizeof(KeyT));
    int32_t* Ct = reinterpret_cast<int32_t*>(Hh + Q_SMALL);
    uint32_t* Ch = reinterpret_cast<uint32_t*>(Ct + Q_SMALL);
    const uint32_t cnt = *p.list_count;
    for(uint32_t it = blockIdx.x * 8 + warp; it < cnt; it += gridDim.x * 8) {
        const uint32_t r = p.list[it];
        const uint64_t beg = p.fpos[uint64_t(p.W[r]) * p.s];
        const uint32_t m = uint32_t(p.fpos[uint64_t(p.W[r + 1]) * p.s] - beg);
        process_read_warp<KeyT>(p, p.rmap ? p.rmap[r] : r, A, Hh, m, static_cast<KeyT*>(p.loc) + beg, lane, Ct, Ch);
    }
}

// Short reads, fully fused: one warp probes the read's features, takes the locations out of the
// 32-byte table slots themselves (lists of length <= 2) or gathers longer lists, sorts them in
// registers / shared memory and derives the top list -- the match list never touches HBM.  Reads
// with more than 32 * NC feature slots or more than CAP matches are left to the staged path:
// they are listed in `left_list` (few left: compacted sub-batch) and flagged in `done` (many left:
// the features of the finished reads are overwritten with NO_FEATURE so that the staged path skips them).
struct FusedParams {
    uint32_t* feat;
    const uint32_t* table;      // 128-byte buckets of 3 slots (common.cuh)
    TableGeom geom;
    const uint64_t* locs;
    int64_t n;
    uint8_t* done;
    uint32_t* left_over;        // [0] number of reads left to the staged path, [1] read counter (dynamic scheduling)
    uint32_t* left_list;        // the reads left over, in no particular order
};

constexpr int Q_FUSED = 128;      // default; databases with long lists (nlocs > 2.1 nkeys) use 256
constexpr uint32_t Q_LANE_GATHER = 12;   // lists up to this length are copied by their own lane    // matches per read of the fused path (small: shared memory is taken from L1, which the probes need)

// candidates.h:236-285 with the taxon already lifted to -lowest and its rank known (seq = the
// taxon is of sequence rank: such taxa are never merged); entry i of the list lives in lane i
__device__ __forceinline__ void warp_topk_insert_lifted(WarpTopK& L, int K, int32_t tax, uint32_t hits, bool seq, int lane)
{
    const bool live = lane < L.n;
    int pos = -1;
    if(!seq) {
        const uint32_t hit = __ballot_sync(0xFFFFFFFFu, live && L.tax == tax);
        if(hit) pos = __ffs(int(hit)) - 1;
    }
    const int32_t up_tax = __shfl_up_sync(0xFFFFFFFFu, L.tax, 1);
    const uint32_t up_hits = __shfl_up_sync(0xFFFFFFFFu, L.hits, 1);
    if(pos >= 0) {
        const uint32_t old = __shfl_sync(0xFFFFFFFFu, L.hits, pos);
        const uint32_t h = hits > old ? hits : old;
        const int j = __popc(__ballot_sync(0xFFFFFFFFu, lane < pos && L.hits >= h));
        if(lane > j && lane <= pos) { L.tax = up_tax; L.hits = up_hits; }
        if(lane == j) { L.tax = tax; L.hits = h; }
        return;
    }
    const int i0 = __popc(__ballot_sync(0xFFFFFFFFu, live && L.hits >= hits));      // upper bound, descending list
    if(i0 < L.n || L.n < K) {
        const int last = L.n < K ? L.n : K - 1;
        if(lane > i0 && lane <= last) { L.tax = up_tax; L.hits = up_hits; }
        if(lane == i0) { L.tax = tax; L.hits = hits; }
        if(L.n < K) L.n++;
    }
}

// sorted matches A[0..m) + per-element range hits Hh -> top list.  The target runs are found, their
// maxima taken (segmented warp scan) and their taxa lifted by all lanes at once (the lift-table
// loads of all candidates are in flight together).  Up to 32 candidates the list is SELECTED
// (warp_toplist_select32); beyond that the order-dependent insertion (candidates.h:236-285) is
// replayed one candidate after the other, out of shared memory.
// Hh == nullptr: the range hits (hits_ending_at) are computed here instead of by a pass of their own -- from the start of
// an element's target run (the nearest run head at or before it, out of the ballot of heads): as long as the run spans
// fewer than numwin windows up to the element, ALL its elements so far are in range (the usual case for a short read) and
// the value rises along the run, so the run's maximum is simply its last value; only otherwise the range is searched and
// the maximum taken by a segmented scan.
template<class KeyT>
__device__ __forceinline__ void toplist_from_sorted(const CandParams& p, uint32_t r, const KeyT* A, const uint32_t* Hh,
                                                    uint32_t m, int32_t* Ct, uint32_t* Ch, int lane, uint32_t numwin)
{
    const uint32_t ltmask = (1u << lane) - 1u;
    uint32_t ncand = 0, carry = 0, last_tgt = 0xFFFFFFFFu, run_start = 0;
    for(uint32_t b = 0; b < m; b += 32) {
        const uint32_t i = b + lane;
        const bool valid = i < m;
        const KeyT key = valid ? A[i] : KeyT(0);
        const uint32_t tgt = valid ? tgt_of(key, p.wbits) : 0xFFFFFFFFu;
        uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, tgt, 1);
        if(lane == 0) prev = last_tgt;
        last_tgt = __shfl_sync(0xFFFFFFFFu, tgt, 31);
        const bool head = valid && (i == 0 || prev != tgt);
        const bool tail = valid && (i + 1 == m || tgt_of(A[i + 1], p.wbits) != tgt);
        const uint32_t hm = __ballot_sync(0xFFFFFFFFu, head);
        uint32_t v = 0u;
        bool rising = false;            // every element's range reaches back to the start of its run: v rises along each run
        if(Hh != nullptr) {
            if(valid) v = Hh[i];
        } else {
            const uint32_t below = hm & (ltmask | (1u << lane));             // run heads at or before this lane
            const uint32_t st = below ? b + 31u - uint32_t(__clz(int(below))) : run_start;
            if(hm) run_start = b + 31u - uint32_t(__clz(int(hm)));
            bool whole = true;
            if(valid) {
                const KeyT wmask = (KeyT(1) << p.wbits) - 1;
                const uint32_t win = uint32_t(key & wmask), win0 = uint32_t(A[st] & wmask);
                whole = win - win0 < numwin;
                v = whole ? i - st + 1u : hits_ending_at(A, i, numwin, p.wbits);
            }
            rising = __all_sync(0xFFFFFFFFu, whole);
        }
        if(!rising) {
            if(lane == 0 && !head) v = max(v, carry);
            // inclusive segmented max scan: take the value d lanes down unless a run starts in between
#pragma unroll
            for(int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, v, d);
                const uint32_t between = (hm >> (lane >= d ? lane - d + 1 : 0)) & ((1u << d) - 1u);    // heads in (lane-d, lane]
                if(lane >= d && between == 0u) v = max(v, o);
            }
        }
        carry = __shfl_sync(0xFFFFFFFFu, v, 31);
        const uint32_t tm = __ballot_sync(0xFFFFFFFFu, tail);
        if(tail) {
            const uint32_t j = ncand + __popc(tm & ltmask);
            Ct[j] = __ldg(p.lift + tgt);          // the target's taxon at -lowest, sequence-rank flag in bit 31
            Ch[j] = v;
        }
        ncand += __popc(tm);
    }
    __syncwarp();
    WarpTopK L;
    L.tax = -1; L.hits = 0; L.n = 0;
    if(ncand <= 32u) warp_toplist_select32(L, p.K, uint32_t(lane) < ncand ? Ct[lane] : 0, uint32_t(lane) < ncand ? Ch[lane] : 0u, lane);
    else warp_toplist_insert_all(L, p.K, Ct, Ch, ncand, lane);
    if(lane < p.K) {
        p.out_tax[uint64_t(r) * p.rs + lane] = lane < L.n ? L.tax : -1;
        p.out_hits[uint64_t(r) * p.rs + lane] = lane < L.n ? L.hits : 0u;
    }
    if(lane == 0) p.out_n[uint64_t(r) * p.ns] = L.n;
    __syncwarp();
}

// few reads left over: their windows / features are compacted into a sub-batch of their own
__global__ void q_left_counts_kernel(const uint32_t* W, const uint32_t* left_list, uint32_t nleft, uint32_t* Wc)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < nleft) { const uint32_t r = left_list[i]; Wc[i] = W[r + 1] - W[r]; }
}

__global__ void q_left_copy_kernel(const uint32_t* feat, const uint32_t* W, const uint32_t* left_list, uint32_t nleft, const uint32_t* Wc,
                                   int s, uint32_t* featc)
{
    const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if(i >= nleft) return;
    const uint32_t r = left_list[i];
    const uint64_t src = uint64_t(W[r]) * s, dst = uint64_t(Wc[i]) * s, cnt = uint64_t(Wc[i + 1] - Wc[i]) * s;
    for(uint64_t t = lane; t < cnt; t += 32) featc[dst + t] = feat[src + t];
}

// features of the reads the fused kernel finished -> NO_FEATURE: the staged path skips them
__global__ void q_hide_done_kernel(uint32_t* feat, const uint32_t* W, int s, const uint8_t* done, int64_t n)
{
    const int64_t r = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if(r >= n || !done[r]) return;
    const uint64_t s0 = uint64_t(W[r]) * s, s1 = uint64_t(W[r + 1]) * s;
    for(uint64_t i = s0 + lane; i < s1; i += 32) feat[i] = EMPTY_KEY;
}

// (launch bound: the 2-slot, 128-match variant fits 40 registers without spills = 6 resident CTAs, 48 warps per SM)
template<class KeyT, int NC, int CAP>  // NC: feature slots per lane (reads with more than 32*NC slots are left over); CAP: matches per read
__global__ void __launch_bounds__(256, NC == 2 ? ((CAP <= 128 && sizeof(KeyT) == 4) ? 6 : 5) : 3)
q_fused_warp_kernel(const CandParams p, const FusedParams f)
{
    extern __shared__ __align__(16) unsigned char sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // per warp: A[CAP] keys, Ct[CAP], Ch[CAP] (the range hits are computed on the fly: no array of their own)
    unsigned char* wbase = sm + size_t(warp) * CAP * (sizeof(KeyT) + 8);
    KeyT* A = reinterpret_cast<KeyT*>(wbase);
    int32_t* Ct = reinterpret_cast<int32_t*>(wbase + size_t(CAP) * sizeof(KeyT));
    uint32_t* Ch = reinterpret_cast<uint32_t*>(Ct + CAP);
    const KeyT KMAX = ~KeyT(0);
    // reads are handed out 8 at a time by an atomic counter (their cost varies with their matches)
    for(int64_t rr = 0;; ++rr) {
        if((rr & 7) == 0) {
            uint32_t b0 = 0;
            if(lane == 0) b0 = atomicAdd(f.left_over + 1, 8u);
            rr = int64_t(__shfl_sync(0xFFFFFFFFu, b0, 0));
        }
        const int64_t r = rr;
        if((r & ~int64_t(7)) >= f.n) break;
        if(r >= f.n) continue;
        const uint64_t s0 = uint64_t(p.W[r]) * p.s, s1 = uint64_t(p.W[r + 1]) * p.s;
        const uint32_t ns = uint32_t(s1 - s0);
        if(ns > 32u * NC) {
            if(lane == 0) { f.left_list[atomicAdd(f.left_over, 1u)] = uint32_t(r); f.done[r] = 0; }
            continue;
        }
        const int nc = int(ns + 31) >> 5;                 // feature slots per lane (uniform)
        // the features of the next read (they follow in memory: reads are taken 8 in a row) towards L1 meanwhile
        if(lane < 2) asm volatile("prefetch.global.L1 [%0];" :: "l"(f.feat + s1 + 32 * lane));
        // (1) probe.  A pass loads the bucket line (keys + sizes) of every unresolved feature of the
        //     lane at once; a feature leaves when its key or a free slot is found.  Passes repeat
        //     while any lane still has a feature whose bucket was full (few: <= 2 in most reads).
        // pl[c][0..5]: the slot's inline match keys (common.cuh, "feature table": 6 packed 32-bit keys or 3 64-bit
        // locations); a list longer than that has its offset into locs in pl[c][5] instead
        constexpr uint32_t INL = sizeof(KeyT) == 4 ? TABLE_INLINE_32 : TABLE_INLINE_64;
        uint32_t key[NC], len[NC], bkt[NC];
        uint32_t pl[NC][6];
        bool pend[NC];
#pragma unroll
        for(int c = 0; c < NC; ++c) {
            key[c] = EMPTY_KEY;
            if(c < nc) {
                const uint32_t sl = uint32_t(c) * 32u + uint32_t(lane);
                if(sl < ns) key[c] = f.feat[s0 + sl];
            }
            pend[c] = key[c] != EMPTY_KEY;
            bkt[c] = pend[c] ? table_bucket(key[c], f.geom) : 0u;
            len[c] = 0;
#pragma unroll
            for(int w = 0; w < 6; ++w) pl[c][w] = 0;
        }
        for(;;) {
            uint4 k4[NC];
#pragma unroll
            for(int c = 0; c < NC; ++c) {
                if(c < nc && pend[c]) k4[c] = __ldg(reinterpret_cast<const uint4*>(f.table + uint64_t(bkt[c]) * TABLE_BUCKET_WORDS));
            }
            bool again = false;
#pragma unroll
            for(int c = 0; c < NC; ++c) {
                if(c < nc && pend[c]) {
                    const int t = table_match(k4[c], key[c]);
                    if(t >= 0) {
                        // payload of the slot: one sector of the same line (L1 hits)
                        const uint4* pay = reinterpret_cast<const uint4*>(f.table + uint64_t(bkt[c]) * TABLE_BUCKET_WORDS + 8 + 8 * t);
                        const uint4 q0 = __ldg(pay), q1 = __ldg(pay + 1);
                        len[c] = q0.x;
                        pl[c][0] = q0.z; pl[c][1] = q0.w; pl[c][2] = q1.x; pl[c][3] = q1.y; pl[c][4] = q1.z;
                        pl[c][5] = q0.x > INL ? q0.y : q1.w;
                        pend[c] = false;
                    } else if(t == -1) {
                        pend[c] = false;
                    } else {
                        ++bkt[c];
                        again = true;
                    }
                }
            }
            if(!__any_sync(0xFFFFFFFFu, again)) break;
        }
        uint32_t tot = 0;
#pragma unroll
        for(int c = 0; c < NC; ++c) tot += len[c];
        // (2) positions: exclusive scan of the per-lane totals
        uint32_t inc = tot;
#pragma unroll
        for(int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if(lane >= d) inc += o;
        }
        const uint32_t m = __shfl_sync(0xFFFFFFFFu, inc, 31);
        if(m > uint32_t(CAP)) {
            if(lane == 0) { f.left_list[atomicAdd(f.left_over, 1u)] = uint32_t(r); f.done[r] = 0; }
            continue;
        }
        // this read is finished here (q_hide_done_kernel hides its features from the staged path,
        // should any read of the batch be left over)
        if(lane == 0) f.done[r] = 1;
        if(m == 0) {
            if(lane == 0) p.out_n[uint64_t(r) * p.ns] = 0;
            if(lane < p.K) { p.out_tax[uint64_t(r) * p.rs + lane] = -1; p.out_hits[uint64_t(r) * p.rs + lane] = 0u; }
            continue;
        }
        // (3) matches into shared memory: short lists straight from the slot, longer ones gathered
        uint32_t pos = inc - tot;
#pragma unroll
        for(int c = 0; c < NC; ++c) {
            if(c < nc) {
                const uint32_t L = len[c];
                if(L >= 1 && L <= INL) {
                    if(sizeof(KeyT) == 4) {
#pragma unroll
                        for(uint32_t t = 0; t < TABLE_INLINE_32; ++t) if(t < L) A[pos + t] = KeyT(pl[c][t]);
                    } else {
#pragma unroll
                        for(uint32_t t = 0; t < TABLE_INLINE_64; ++t)
                            if(t < L) A[pos + t] = KeyT(uint64_t(pl[c][2 * t]) | (uint64_t(pl[c][2 * t + 1]) << 32));
                    }
                }
                if(L > INL && L <= Q_LANE_GATHER) {
                    // short lists: every lane copies its own (a whole-warp copy per list costs more
                    // instructions than the list has elements)
                    const uint64_t* src = f.locs + pl[c][5];
                    for(uint32_t t = 0; t < L; t += 4) {
                        uint64_t x0 = __ldg(src + t), x1 = 0, x2 = 0, x3 = 0;
                        if(t + 1 < L) x1 = __ldg(src + t + 1);
                        if(t + 2 < L) x2 = __ldg(src + t + 2);
                        if(t + 3 < L) x3 = __ldg(src + t + 3);
                        A[pos + t] = make_key<KeyT>(x0, p.wbits);
                        if(t + 1 < L) A[pos + t + 1] = make_key<KeyT>(x1, p.wbits);
                        if(t + 2 < L) A[pos + t + 2] = make_key<KeyT>(x2, p.wbits);
                        if(t + 3 < L) A[pos + t + 3] = make_key<KeyT>(x3, p.wbits);
                    }
                }
                uint32_t big = __ballot_sync(0xFFFFFFFFu, L > Q_LANE_GATHER);
                while(big) {
                    const int src = __ffs(int(big)) - 1;
                    big &= big - 1;
                    const uint32_t Ls = __shfl_sync(0xFFFFFFFFu, L, src);
                    const uint32_t Os = __shfl_sync(0xFFFFFFFFu, pl[c][5], src);
                    const uint32_t Ps = __shfl_sync(0xFFFFFFFFu, pos, src);
                    for(uint32_t t = lane; t < Ls; t += 32) A[Ps + t] = make_key<KeyT>(f.locs[Os + t], p.wbits);
                }
                pos += L;
            }
        }
        __syncwarp();
        // (4) sort by (tgt, win)
        if(m <= 32) {
            KeyT a = lane < int(m) ? A[lane] : KMAX;
            __syncwarp();
            a = wsort32(a, lane);
            A[lane] = a;
        } else if(m <= 64) {
            KeyT a = A[lane];
            KeyT b2 = lane + 32 < int(m) ? A[lane + 32] : KMAX;
            __syncwarp();
            wsort64(a, b2, lane);
            A[lane] = a;
            A[lane + 32] = b2;
        } else if(m <= 96) {
            KeyT a = A[lane], b2 = A[lane + 32];
            KeyT c2 = lane + 64 < int(m) ? A[lane + 64] : KMAX;
            __syncwarp();
            wsort96(a, b2, c2, lane);
            A[lane] = a;
            A[lane + 32] = b2;
            A[lane + 64] = c2;
        } else if(m <= 128) {
            KeyT v[4];
#pragma unroll
            for(int i = 0; i < 4; ++i) v[i] = 32 * i + lane < int(m) ? A[32 * i + lane] : KMAX;
            __syncwarp();
            wsort128(v, lane);
#pragma unroll
            for(int i = 0; i < 4; ++i) A[32 * i + lane] = v[i];
        } else {
            bitonic_sort<true>(A, m, lane, 32);
        }
        __syncwarp();
        // (5) window ranges + top list
        const uint32_t numwin = num_windows_in_range(p, uint32_t(r));
        toplist_from_sorted<KeyT>(p, uint32_t(r), A, nullptr, m, Ct, Ch, lane, numwin);
    }
}

// Reads the fused kernel left over although they are short (more matches than it holds per warp, or more
// feature slots than its variant holds per lane -- up to Q_SPILL_SLOTS): the same per-read pipeline, one warp per
// read, with the match list in a per-warp region of GLOBAL memory sized for the worst case (every slot a full
// list).  They are rare (about 1 in 1000 read pairs against a 2.5 Gbp shard), so this kernel's speed does not
// matter; what matters is that it needs no decision of the host: it is launched unconditionally behind the fused
// kernel and finds its work in the left-over list.  Long reads (more slots) are passed on in left_list2 for the
// staged path.
constexpr int Q_SPILL_SLOTS = 128;
constexpr int Q_SPILL_WARPS = 4;

struct SpillParams {
    const uint32_t* feat;
    const uint32_t* table;
    TableGeom geom;
    const uint64_t* locs;
    uint8_t* done;
    const uint32_t* left_count;   // reads left over by the fused kernel
    const uint32_t* left_list;
    uint32_t* left_count2;        // reads passed on to the staged path
    uint32_t* left_list2;
    unsigned char* scratch;       // [warps of the grid][cap * (sizeof(KeyT) + 12)]
    uint32_t cap;                 // Q_SPILL_SLOTS * maxlocs
};

template<class KeyT>
__global__ void __launch_bounds__(Q_SPILL_WARPS * 32)
q_spill_warp_kernel(const CandParams p, const SpillParams f)
{
    const int lane = threadIdx.x & 31;
    const uint32_t gw = blockIdx.x * Q_SPILL_WARPS + (threadIdx.x >> 5), nw = gridDim.x * Q_SPILL_WARPS;
    unsigned char* base = f.scratch + size_t(gw) * f.cap * (sizeof(KeyT) + 12);
    KeyT* A = reinterpret_cast<KeyT*>(base);
    uint32_t* Hh = reinterpret_cast<uint32_t*>(base + size_t(f.cap) * sizeof(KeyT));
    int32_t* Ct = reinterpret_cast<int32_t*>(Hh + f.cap);
    uint32_t* Ch = reinterpret_cast<uint32_t*>(Ct + f.cap);
    const uint32_t nleft = *f.left_count;
    constexpr int NCS = Q_SPILL_SLOTS / 32;
    for(uint32_t it = gw; it < nleft; it += nw) {
        const uint32_t r = f.left_list[it];
        const uint64_t s0 = uint64_t(p.W[r]) * p.s, s1 = uint64_t(p.W[r + 1]) * p.s;
        if(s1 - s0 > uint64_t(Q_SPILL_SLOTS)) {
            if(lane == 0) f.left_list2[atomicAdd(f.left_count2, 1u)] = r;
            continue;
        }
        const uint32_t ns = uint32_t(s1 - s0);
        constexpr uint32_t INL = sizeof(KeyT) == 4 ? TABLE_INLINE_32 : TABLE_INLINE_64;
        uint32_t len[NCS], off[NCS];
        uint32_t pl[NCS][6];
        uint32_t tot = 0;
#pragma unroll
        for(int c = 0; c < NCS; ++c) {
            len[c] = 0; off[c] = 0;
#pragma unroll
            for(int w = 0; w < 6; ++w) pl[c][w] = 0;
            const uint32_t sl = uint32_t(c) * 32u + uint32_t(lane);
            const uint32_t key = sl < ns ? f.feat[s0 + sl] : EMPTY_KEY;
            if(key != EMPTY_KEY) {
                uint32_t b = table_bucket(key, f.geom);
                for(;;) {
                    const uint32_t* line = f.table + uint64_t(b) * TABLE_BUCKET_WORDS;
                    const uint4 k4 = __ldg(reinterpret_cast<const uint4*>(line));
                    const int t = table_match(k4, key);
                    if(t >= 0) {
                        const uint4* pay = reinterpret_cast<const uint4*>(line + 8 + 8 * t);
                        const uint4 q0 = __ldg(pay), q1 = __ldg(pay + 1);
                        len[c] = q0.x;
                        off[c] = q0.y;
                        pl[c][0] = q0.z; pl[c][1] = q0.w; pl[c][2] = q1.x; pl[c][3] = q1.y; pl[c][4] = q1.z; pl[c][5] = q1.w;
                        break;
                    }
                    if(t == -1) break;
                    ++b;
                }
            }
            tot += len[c];
        }
        uint32_t inc = tot;
#pragma unroll
        for(int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if(lane >= d) inc += o;
        }
        const uint32_t m = __shfl_sync(0xFFFFFFFFu, inc, 31);
        if(lane == 0) f.done[r] = 1;
        if(m == 0) {
            if(lane == 0) p.out_n[uint64_t(r) * p.ns] = 0;
            if(lane < p.K) { p.out_tax[uint64_t(r) * p.rs + lane] = -1; p.out_hits[uint64_t(r) * p.rs + lane] = 0u; }
            continue;
        }
        // matches in feature-slot order (any order would do: they are sorted next)
        uint32_t pos = inc - tot;
#pragma unroll
        for(int c = 0; c < NCS; ++c) {
            const uint32_t L = len[c];
            if(L <= INL) {
#pragma unroll
                for(uint32_t t = 0; t < INL; ++t)
                    if(t < L) A[pos + t] = sizeof(KeyT) == 4 ? KeyT(pl[c][t]) : KeyT(uint64_t(pl[c][2 * t]) | (uint64_t(pl[c][2 * t + 1]) << 32));
            } else {
                for(uint32_t t = 0; t < L; ++t) A[pos + t] = make_key<KeyT>(f.locs[off[c] + t], p.wbits);
            }
            pos += L;
        }
        __syncwarp();
        bitonic_sort<true>(A, m, lane, 32);
        __syncwarp();
        const uint32_t numwin = num_windows_in_range(p, r);
        for(uint32_t i = lane; i < m; i += 32) Hh[i] = hits_ending_at(A, i, numwin, p.wbits);
        __syncwarp();
        toplist_from_sorted<KeyT>(p, r, A, Hh, m, Ct, Ch, lane);
    }
}

// medium / large reads: one CTA per read, capacity fixed by the launch
template<class KeyT>
__global__ void q_cand_cta_kernel(const CandParams p, uint32_t capacity)
{
    extern __shared__ __align__(16) unsigned char sm[];
    KeyT* A = reinterpret_cast<KeyT*>(sm);
    uint32_t* Hh = reinterpret_cast<uint32_t*>(sm + size_t(capacity) * sizeof(KeyT));
    const uint32_t cnt = *p.list_count;
    for(uint32_t it = blockIdx.x; it < cnt; it += gridDim.x) {
        const uint32_t r = p.list[it];
        const uint64_t beg = p.fpos[uint64_t(p.W[r]) * p.s];
        const uint32_t m = uint32_t(p.fpos[uint64_t(p.W[r + 1]) * p.s] - beg);
        process_read_cta<KeyT>(p, p.rmap ? p.rmap[r] : r, A, Hh, m, static_cast<KeyT*>(p.loc) + beg, true, threadIdx.x, blockDim.x, reinterpret_cast<int32_t*>(A),
                               sizeof(KeyT) == 4 ? Hh + capacity : nullptr);
    }
}

// oversized reads: sorted in place in global memory by one CTA (correctness fallback)
template<class KeyT>
__global__ void __launch_bounds__(1024)
q_cand_global_kernel(const CandParams p)
{
    const uint32_t cnt = *p.list_count;
    for(uint32_t it = blockIdx.x; it < cnt; it += gridDim.x) {
        const uint32_t r = p.list[it];
        const uint64_t beg = p.fpos[uint64_t(p.W[r]) * p.s];
        const uint64_t m64 = p.fpos[uint64_t(p.W[r + 1]) * p.s] - beg;
        KeyT* A = static_cast<KeyT*>(p.loc) + beg;
        process_read_cta<KeyT>(p, p.rmap ? p.rmap[r] : r, A, p.hits_scratch + beg, uint32_t(m64), A, false, threadIdx.x, blockDim.x,
                               reinterpret_cast<int32_t*>(p.hits_scratch2 + beg));
    }
}

template<class KeyT>
__global__ void q_split_locs_kernel(const KeyT* loc, uint64_t n, int wbits, uint32_t* tg, uint32_t* wn)
{
    const uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if(i < n) { tg[i] = uint32_t(loc[i] >> wbits); wn[i] = uint32_t(loc[i] & ((KeyT(1) << wbits) - 1)); }
}

__global__ void q_read_counts_kernel(const uint64_t* fpos, const uint32_t* W, int s, int64_t n, int64_t* out)
{
    const int64_t r = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if(r < n) out[r] = int64_t(fpos[uint64_t(W[r + 1]) * s] - fpos[uint64_t(W[r]) * s]);
}

__global__ void q_lift_kernel(const int32_t* lineage, const uint8_t* rank, int64_t ntargets, int lowest, int32_t* lift)
{
    const int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if(t >= ntargets) return;
    int32_t tax = int32_t(t);
    if(lowest > MCG_RANK_SEQUENCE) {
        const int32_t anc = lineage[t * MCG_NUM_RANKS + lowest];
        if(anc >= 0) tax = anc;
    }
    if(rank[tax] == MCG_RANK_SEQUENCE) tax |= int32_t(0x80000000u);
    lift[t] = tax;
}

// ------------------------------------------------------------------ K6: merge + classify + statistics
struct MergeParams {
    const int32_t* cand_tax;
    const uint32_t* cand_hits;
    const int32_t* cand_n;
    uint32_t rs, ns;             // row strides in words of cand_tax / cand_hits and of cand_n (K and 1, or 2K + 1 for packed records)
    int P;
    int64_t n;
    int K;
    int lowest, highest, hitmin;
    float hitdiff;
    int compat;                  // reproduce quirk Q1 along the reference's reduction tree
    int32_t ntaxa, ntargets;
    const int32_t* lineage;
    const uint8_t* rank;
    int32_t* out_tax;
    uint32_t* out_hits;
    int32_t* out_n;
    int32_t* out_best;
    int32_t* out_brank;
    unsigned long long* stats;
    uint32_t* bad;               // set when a candidate taxon lies outside [-1, ntaxa)
};

// inserts the entries of `src` into `dst` in list order (querying.h:944-971); with `drop_seq` the
// entries whose taxon is a target's own sequence-level taxon are skipped: their ids are negative
// (sketch_database.h:150), travel as uint32 (querying.h:984,1001) and no longer name a taxon once
// widened on the receiving rank (:958) -- quirk Q1
__device__ __forceinline__ void merge_lists(TopK& dst, const TopK& src, const MergeParams& p, bool drop_seq)
{
    for(int c = 0; c < src.n; ++c) {
        if(drop_seq && src.tax[c] < p.ntargets) continue;
        topk_insert(dst, p.K, src.tax[c], src.hits[c], p.lowest, p.lineage, p.rank);
    }
}

__device__ __forceinline__ void load_list(TopK& L, const MergeParams& p, int q, int64_t r)
{
    const int64_t row = int64_t(q) * p.n + r;
    int cnt = p.cand_n[row * p.ns];
    if(cnt < 0 || cnt > p.K) { atomicOr(p.bad, 2u); cnt = cnt < 0 ? 0 : p.K; }
    L.n = 0;
    for(int c = 0; c < cnt; ++c) {
        const int32_t t = p.cand_tax[row * p.rs + c];
        if(t < 0 || t >= p.ntaxa) { if(t != -1) atomicOr(p.bad, 1u); continue; }
        // a per-shard list is a valid top list already: re-inserting it (querying.h:910-940) reproduces it
        topk_insert(L, p.K, t, p.cand_hits[row * p.rs + c], p.lowest, p.lineage, p.rank);
    }
}

__global__ void __launch_bounds__(128)
q_merge_classify_kernel(const MergeParams p)
{
    __shared__ unsigned int s_stats[22];
    if(threadIdx.x < 22) s_stats[threadIdx.x] = 0;
    __syncthreads();
    const int64_t r = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if(r < p.n) {
        TopK L;
        L.n = 0;
        if(!p.compat) {
            for(int q = 0; q < p.P; ++q) {                   // rank order (querying.h:905-971)
                const int64_t row = int64_t(q) * p.n + r;
                int cnt = p.cand_n[row * p.ns];
                if(cnt < 0 || cnt > p.K) { atomicOr(p.bad, 2u); cnt = cnt < 0 ? 0 : p.K; }
                for(int c = 0; c < cnt; ++c) {
                    const int32_t t = p.cand_tax[row * p.rs + c];
                    if(t < 0 || t >= p.ntaxa) { if(t != -1) atomicOr(p.bad, 1u); continue; }
                    topk_insert(L, p.K, t, p.cand_hits[row * p.rs + c], p.lowest, p.lineage, p.rank);
                }
            }
        } else if((p.P & (p.P - 1)) == 0) {
            // the reduction tree of querying.h:892-1066 for P = 2^d: round 1 pairs (1->0), (3->2), ...,
            // round 2 (2->0), (6->4), ...: a receiver keeps its own list whole and takes from the
            // sender's (already merged and truncated) list only what survives the uint32 round trip.
            // Evaluated like a binary counter over the ranks in order; the stack never holds more
            // than log2(P) + 1 lists.
            TopK st[7];
            int lvl[7];
            int sp = 0;
            for(int q = 0; q < p.P; ++q) {
                load_list(st[sp], p, q, r);
                lvl[sp] = 0;
                ++sp;
                while(sp >= 2 && lvl[sp - 1] == lvl[sp - 2]) {
                    merge_lists(st[sp - 2], st[sp - 1], p, true);
                    ++lvl[sp - 2];
                    --sp;
                }
            }
            L = st[0];
        } else {
            // no reduction tree exists for other P in the reference (quirk Q12): rank order, rank 0 whole
            TopK S;
            for(int q = 0; q < p.P; ++q) {
                load_list(S, p, q, r);
                merge_lists(L, S, p, q > 0);
            }
        }
        for(int c = 0; c < p.K; ++c) {
            p.out_tax[r * p.K + c] = c < L.n ? L.tax[c] : -1;
            p.out_hits[r * p.K + c] = c < L.n ? L.hits[c] : 0u;
        }
        p.out_n[r] = L.n;
        // classification.cpp:236-265
        int32_t lca = -1;
        if(L.n > 0 && L.hits[0] >= uint32_t(p.hitmin)) {
            lca = L.tax[0];
            const float thr = L.hits[0] > uint32_t(p.hitmin)
                            ? __fmul_rn(__uint2float_rn(L.hits[0] - uint32_t(p.hitmin)), p.hitdiff) : 0.0f;
            for(int i = 1; i < L.n; ++i) {
                if(!(__uint2float_rn(L.hits[i]) > thr)) break;
                // taxonomy.h:531-537 ranked_lca
                const int32_t* la = p.lineage + int64_t(lca) * MCG_NUM_RANKS;
                const int32_t* lb = p.lineage + int64_t(L.tax[i]) * MCG_NUM_RANKS;
                int32_t x = -1;
                for(int k = 0; k < MCG_NUM_RANKS; ++k) {
                    const int32_t a = la[k];
                    if(a >= 0 && a == lb[k]) { x = a; break; }
                }
                lca = x;
                if(lca < 0 || int(p.rank[lca]) > p.highest) { lca = -1; break; }
            }
            if(lca >= 0 && int(p.rank[lca]) > p.highest) lca = -1;
        }
        const int br = lca >= 0 ? int(p.rank[lca]) : MCG_RANK_NONE;
        p.out_best[r] = lca;
        p.out_brank[r] = br;
        // classification_statistics.h:70-78
        if(br >= MCG_RANK_NONE) atomicAdd(&s_stats[MCG_RANK_NONE], 1u);
        else for(int k = br; k < MCG_NUM_RANKS; ++k) atomicAdd(&s_stats[k], 1u);
    }
    __syncthreads();
    if(p.stats != nullptr && threadIdx.x < 22 && s_stats[threadIdx.x] != 0)
        atomicAdd(&p.stats[threadIdx.x], (unsigned long long)s_stats[threadIdx.x]);
}

// ------------------------------------------------------------------ host orchestration
//
// A query call is cut into PARTS.  Stage A of a part is enqueued without any host round trip:
//   window counts + per-read meta -> 3 scans -> K1 on both mates (launches sized from upper bounds, the
//   kernels read the exact window counts from the device) -> [sharded: all-gather of features + meta over
//   NVLink] -> q_fused_warp_kernel over all reads of the part -> left-over count into pinned memory.
// Before the host looks at that count (the only decision it has to take: reads the fused kernel could
// not hold take the staged path), stage A of the NEXT part is already enqueued, so the GPU never waits
// for the host.  Stage B: left-over reads (rare), [sharded: all-to-all of the per-read top lists, merge +
// decision of the own block], copy-out.
struct QueryRun {
    uint64_t nwin = 0, nslots = 0, total = 0;
    uint32_t cls[4] = {0, 0, 0, 0};
    // what the staged path works on: the whole batch, or the compacted reads the fused kernel left over
    const uint32_t* feat = nullptr;     // [nslots]
    const uint32_t* W = nullptr;        // [n + 1] window prefix
    const uint32_t* rmap = nullptr;     // position -> read of the batch (nullptr: identity)
};

static int check_opts(const mcg_db* db, const mcg_query_opts* o)
{
    MCG_REQUIRE(o->sketchlen >= 1 && o->sketchlen <= MCG_MAX_SKETCH, "query sketchlen %d outside 1..%d", o->sketchlen, MCG_MAX_SKETCH);
    MCG_REQUIRE(o->winlen >= 1 && o->winstride >= 1, "query window length / stride must be >= 1");
    MCG_REQUIRE(o->maxcand >= 1 && o->maxcand <= MCG_MAX_CAND, "maxcand %d outside 1..%d (unlimited lists are not supported on the device)", o->maxcand, MCG_MAX_CAND);
    MCG_REQUIRE(o->lowest >= 0 && o->lowest < MCG_NUM_RANKS && o->highest >= 0 && o->highest < MCG_NUM_RANKS, "rank out of range");
    MCG_REQUIRE(o->hitmin >= 0, "hitmin %d is negative", o->hitmin);
    MCG_REQUIRE(o->insertsize >= 0, "insertsize %lld is negative", (long long)o->insertsize);
    // every target id stored in the shard indexes the taxonomy tables (lineage[tgt], rank[tgt])
    if(db != nullptr && db->nlocs > 0)
        MCG_REQUIRE(int64_t(db->max_tgt) < db->ctx->ntargets, "shard holds target id %llu but the uploaded taxonomy has %lld targets",
                    (unsigned long long)db->max_tgt, (long long)db->ctx->ntargets);
    return MCG_OK;
}

static int check_ready(mcg_db* db, const mcg_query_opts* o)
{
    MCG_TRY(check_opts(db, o));
    if(!db->finalized) { set_error("database not finalized"); return MCG_ERR_STATE; }
    mcg_ctx* c = db->ctx;
    if(c->lineage.p == nullptr) { set_error("taxonomy not uploaded (mcg_tax_upload)"); return MCG_ERR_STATE; }
    if(c->lift_lowest != o->lowest) {                    // per-target lifting table of this -lowest (rebuilt when it changes)
        MCG_TRY(c->lift.reserve(size_t(c->ntargets + 1) * 4));
        if(c->ntargets > 0) {
            q_lift_kernel<<<grid_for(uint64_t(c->ntargets)), 256, 0, c->stream>>>(c->lineage.as<int32_t>(), c->rank.as<uint8_t>(), c->ntargets,
                                                                                  o->lowest, c->lift.as<int32_t>());
            c->launches++;
            MCG_CUDA(cudaGetLastError());
        }
        c->lift_lowest = o->lowest;
    }
    return MCG_OK;
}

// upper bound of the windows of n sequences with `bytes` bases in total: a sequence has at most len / stride + 2
static uint64_t windows_bound(int64_t bytes, int64_t n, int64_t stride, int64_t winlen, int64_t max_len)
{
    if(n <= 0) return 0;
    uint64_t b = uint64_t(bytes / stride) + 2ull * uint64_t(n);
    if(max_len > 0) b = std::min<uint64_t>(b, uint64_t(n) * uint64_t(num_windows(max_len, winlen, stride)));
    return b;
}

// window bookkeeping of n reads in one launch (+ one memset of its status words)
static int launch_prep(mcg_db* db, const mcg_query_opts* o, const int64_t* doffs1, const int64_t* doffs2, int64_t n,
                       uint32_t* wf1, uint32_t* wf2, uint32_t* W, uint32_t* of2, uint2* meta, uint32_t* clear8)
{
    mcg_ctx* c = db->ctx;
    const uint32_t ntiles = uint32_t((n + 1 + QP_TILE - 1) / QP_TILE);
    MCG_TRY(c->scan_tmp.reserve((size_t(ntiles) + 2) * 8));
    MCG_CUDA(cudaMemsetAsync(c->scan_tmp.p, 0, (size_t(ntiles) + 1) * 8, c->stream));
    PrepParams p;
    p.offs1 = doffs1; p.offs2 = doffs2; p.n = n;
    p.winlen = o->winlen; p.stride = o->winstride; p.insertsize = o->insertsize; p.tstride = db->cfg.winstride;
    p.wf1 = wf1; p.wf2 = wf2; p.W = W; p.of2 = of2; p.meta = meta;
    p.status = c->scan_tmp.as<unsigned long long>();
    p.tile_ctr = reinterpret_cast<uint32_t*>(p.status + ntiles);
    p.clear8 = clear8;
    q_prep_kernel<<<ntiles, QP_THREADS, 0, c->stream>>>(p);
    c->launches++;
    MCG_CUDA(cudaGetLastError());
    return MCG_OK;
}

// stage A, local part: window bookkeeping + K1 of both mates of n device-resident reads.  Sketches land
// read-major at feat_out[W[r] * s ...], W (exclusive prefix, total at [n]) and meta are written; nothing
// returns to the host.  feat_out must hold *wbound * s features.
static int qa_sketch(mcg_db* db, const mcg_query_opts* o,
                     const char* d1, const int64_t* doffs1, int64_t bytes1,
                     const char* d2, const int64_t* doffs2, int64_t bytes2, int64_t n, int64_t max_len,
                     uint32_t* W, uint2* meta, uint32_t* feat_out, uint64_t wb1, uint64_t wb2, uint32_t* clear8 = nullptr)
{
    mcg_ctx* c = db->ctx;
    const int s = o->sketchlen;
    MCG_TRY(c->tmp0.reserve(size_t(n + 2) * 4));
    MCG_TRY(c->tmp1.reserve(size_t(n + 2) * 4));
    MCG_TRY(c->tmp3.reserve(size_t(n + 2) * 4));
    uint32_t* wf1 = c->tmp0.as<uint32_t>();
    uint32_t* wf2 = c->tmp1.as<uint32_t>();
    uint32_t* of2 = c->tmp3.as<uint32_t>();
    MCG_TRY(launch_prep(db, o, doffs1, d2 ? doffs2 : nullptr, n, wf1, d2 ? wf2 : nullptr, W, d2 ? of2 : nullptr, meta, clear8));
    mcg_sketch_cfg cfg = db->cfg;
    (void)max_len;
    MCG_TRY(sketch_launch(c, cfg, s, o->winlen, o->winstride, d1, bytes1, doffs1, n, wf1, W, wb1, feat_out, nullptr, 0, nullptr, nullptr, wf1 + n));
    if(d2) MCG_TRY(sketch_launch(c, cfg, s, o->winlen, o->winstride, d2, bytes2, doffs2, n, wf2, of2, wb2, feat_out, nullptr, 0, nullptr, nullptr, wf2 + n));
    MCG_CUDA(cudaGetLastError());
    return MCG_OK;
}

// staged path: find -> scan -> binning -> gather (reads marked done are skipped)
static int query_find_gather(mcg_db* db, const mcg_query_opts* o, int64_t n, int32_t* d_out_tax, uint32_t* d_out_hits, int32_t* d_out_n,
                             uint32_t rs, uint32_t ns, const uint8_t* d_done, QueryRun* run)
{
    mcg_ctx* c = db->ctx;
    const int s = o->sketchlen;
    const uint64_t nslots = run->nslots;
    const uint32_t* W = run->W;
    // K3: find
    MCG_TRY(c->q_flen.reserve(nslots + 16));
    MCG_TRY(c->q_foff.reserve((nslots + 1) * 8));
    MCG_TRY(c->q_fpos.reserve((nslots + 2) * 8));
    if(nslots > 0) {
        q_find_kernel<<<grid_for(nslots), 256, 0, c->stream>>>(run->feat, nslots, db->table.as<uint32_t>(), db->geom(),
                                                              c->q_flen.as<uint8_t>(), c->q_foff.as<uint64_t>());
        c->launches++;
    }
    MCG_TRY(exclusive_scan_u8_to_u64(c, c->q_flen.as<uint8_t>(), c->q_fpos.as<uint64_t>(), nslots));
    // binning (also writes out_n = 0 for reads without matches)
    MCG_TRY(c->q_cls.reserve(size_t(4) * size_t(n) * 4 + 64));
    MCG_TRY(c->q_rinfo.reserve(64));
    uint32_t* counters = c->q_rinfo.as<uint32_t>();
    MCG_CUDA(cudaMemsetAsync(counters, 0, 16, c->stream));
    q_bin_reads_kernel<<<grid_for(n), 256, 0, c->stream>>>(c->q_fpos.as<uint64_t>(), W, s, n, c->q_cls.as<uint32_t>(), counters, d_out_n, d_out_tax, d_out_hits,
                                                          o->maxcand, rs, ns, d_done, run->rmap);
    c->launches++;
    uint64_t total = 0;
    MCG_CUDA(cudaMemcpyAsync(&total, c->q_fpos.as<uint64_t>() + nslots, 8, cudaMemcpyDeviceToHost, c->stream));
    MCG_CUDA(cudaMemcpyAsync(run->cls, counters, 16, cudaMemcpyDeviceToHost, c->stream));
    MCG_CUDA(cudaStreamSynchronize(c->stream));
    run->total = total;
    // gather
    MCG_TRY(c->q_loc.reserve((total + 2) * 8));
    if(nslots > 0 && total > 0) {
        if(db->key32)
            q_gather_kernel<uint32_t><<<grid_for(nslots), 256, 0, c->stream>>>(c->q_flen.as<uint8_t>(), c->q_foff.as<uint64_t>(), c->q_fpos.as<uint64_t>(),
                                                                              nslots, db->locs.as<uint64_t>(), c->q_loc.as<uint32_t>(), db->wbits);
        else
            q_gather_kernel<uint64_t><<<grid_for(nslots), 256, 0, c->stream>>>(c->q_flen.as<uint8_t>(), c->q_foff.as<uint64_t>(), c->q_fpos.as<uint64_t>(),
                                                                              nslots, db->locs.as<uint64_t>(), c->q_loc.as<uint64_t>(), db->wbits);
        c->launches++;
    }
    MCG_CUDA(cudaGetLastError());
    return MCG_OK;
}

// where the per-read lists go: separate arrays (rs = K, ns = 1) or one packed record per read
struct OutRows {
    int32_t* tax; uint32_t* hits; int32_t* n; uint32_t rs, ns;
    static OutRows separate(int32_t* t, uint32_t* h, int32_t* n_, int K) { return OutRows{t, h, n_, uint32_t(K), 1u}; }
    static OutRows packed(uint32_t* base, int K)
    {
        return OutRows{reinterpret_cast<int32_t*>(base), base + K, reinterpret_cast<int32_t*>(base + 2 * K), uint32_t(2 * K + 1), uint32_t(2 * K + 1)};
    }
};

static void fill_cand_params(mcg_db* db, const mcg_query_opts* o, const uint32_t* W, const uint2* meta,
                             const OutRows& out, int write_sorted, CandParams* pp)
{
    int32_t* d_out_tax = out.tax; uint32_t* d_out_hits = out.hits; int32_t* d_out_n = out.n;
    mcg_ctx* c = db->ctx;
    CandParams& p = *pp;
    p.fpos = c->q_fpos.as<uint64_t>();
    p.W = W;
    p.rmap = nullptr;
    p.s = o->sketchlen;
    p.meta = meta;
    p.loc = c->q_loc.p;
    p.wbits = db->wbits;
    p.kbits = db->kbits;
    p.hits_scratch = nullptr;
    p.hits_scratch2 = nullptr;
    p.list = nullptr;
    p.list_count = nullptr;
    p.K = o->maxcand;
    p.lowest = o->lowest;
    p.lineage = c->lineage.as<int32_t>();
    p.rank = c->rank.as<uint8_t>();
    p.lift = c->lift.as<int32_t>();
    p.out_tax = d_out_tax;
    p.out_hits = d_out_hits;
    p.out_n = d_out_n;
    p.rs = out.rs; p.ns = out.ns;
    p.write_sorted = write_sorted;
}

static int query_back(mcg_db* db, const mcg_query_opts* o, const uint2* meta, int64_t n,
                      const QueryRun& run, const OutRows& out, int write_sorted)
{
    mcg_ctx* c = db->ctx;
    CandParams p;
    fill_cand_params(db, o, run.W, meta, out, write_sorted, &p);
    const uint32_t* lists = c->q_cls.as<uint32_t>();
    const uint32_t* counters = c->q_rinfo.as<uint32_t>();
    p.rmap = run.rmap;
    const bool k32 = db->key32;
    const int ksz = k32 ? 4 : 8;
    // function attributes are per DEVICE (one process may drive several GPUs): set before every launch
#define QC_SMEM(kern, bytes) MCG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes)))
    if(run.cls[0] > 0) {
        p.list = lists; p.list_count = counters;
        const unsigned g = unsigned(std::min<uint64_t>((run.cls[0] + 7) / 8, uint64_t(c->num_sms) * 16));
        if(k32) { QC_SMEM(q_cand_warp_kernel<uint32_t>, 8 * Q_SMALL * 16); q_cand_warp_kernel<uint32_t><<<g, 256, 8 * Q_SMALL * (ksz + 12), c->stream>>>(p); }
        else    { QC_SMEM(q_cand_warp_kernel<uint64_t>, 8 * Q_SMALL * 20); q_cand_warp_kernel<uint64_t><<<g, 256, 8 * Q_SMALL * (ksz + 12), c->stream>>>(p); }
        c->launches++;
    }
    if(run.cls[1] > 0) {
        p.list = lists + uint64_t(n); p.list_count = counters + 1;
        const unsigned g = unsigned(std::min<uint64_t>(run.cls[1], uint64_t(c->num_sms) * 8));
        if(k32) { QC_SMEM(q_cand_cta_kernel<uint32_t>, Q_LARGE * 8 + 16 * 1024); q_cand_cta_kernel<uint32_t><<<g, 256, Q_MED * (ksz + 4) + 8 * 1024, c->stream>>>(p, Q_MED); }        // + digit counters of the radix sort
        else    { QC_SMEM(q_cand_cta_kernel<uint64_t>, Q_LARGE * 12); q_cand_cta_kernel<uint64_t><<<g, 256, Q_MED * (ksz + 4), c->stream>>>(p, Q_MED); }
        c->launches++;
    }
    if(run.cls[2] > 0) {
        p.list = lists + uint64_t(2) * n; p.list_count = counters + 2;
        const unsigned g = unsigned(std::min<uint64_t>(run.cls[2], uint64_t(c->num_sms) * 2));
        if(k32) { QC_SMEM(q_cand_cta_kernel<uint32_t>, Q_LARGE * 8 + 16 * 1024); q_cand_cta_kernel<uint32_t><<<g, 512, Q_LARGE * (ksz + 4) + 16 * 1024, c->stream>>>(p, Q_LARGE); }
        else    { QC_SMEM(q_cand_cta_kernel<uint64_t>, Q_LARGE * 12); q_cand_cta_kernel<uint64_t><<<g, 512, Q_LARGE * (ksz + 4), c->stream>>>(p, Q_LARGE); }
        c->launches++;
    }
    if(run.cls[3] > 0) {
        MCG_TRY(c->tmp4.reserve((run.total + 2) * 8));
        p.hits_scratch = c->tmp4.as<uint32_t>();
        p.hits_scratch2 = c->tmp4.as<uint32_t>() + run.total + 2;
        p.list = lists + uint64_t(3) * n; p.list_count = counters + 3;
        const unsigned g = unsigned(std::min<uint64_t>(run.cls[3], uint64_t(c->num_sms) * 2));
        if(k32) q_cand_global_kernel<uint32_t><<<g, 1024, 0, c->stream>>>(p);
        else    q_cand_global_kernel<uint64_t><<<g, 1024, 0, c->stream>>>(p);
        c->launches++;
    }
#undef QC_SMEM
    MCG_CUDA(cudaGetLastError());
    return MCG_OK;
}

// stage A, fused kernel: all reads of the part (features / W / meta in `part`) -> per-read top lists; the
// number of reads it could not hold goes to pinned memory (h_left), followed by part.evA
static int qa_fused(mcg_db* db, const mcg_query_opts* o, QPart& part, int slot, int64_t n, double slots_per_read,
                    const OutRows& out, bool ctr_cleared)
{
    mcg_ctx* c = db->ctx;
    // (every read gets its whole row -- K slots, unused ones taxon -1 / hits 0 -- from the kernel that finishes it)
    MCG_TRY(part.done.reserve(size_t(n) + 16));
    MCG_TRY(part.left_list.reserve(size_t(n + 2) * 4));
    uint32_t* left = part.ctr.as<uint32_t>();
    if(!ctr_cleared) MCG_CUDA(cudaMemsetAsync(left, 0, 16, c->stream));
    CandParams p;
    fill_cand_params(db, o, part.W.as<uint32_t>(), part.meta.as<uint2>(), out, 0, &p);
    FusedParams f;
    f.feat = part.feat.as<uint32_t>(); f.table = db->table.as<uint32_t>(); f.geom = db->geom();
    f.locs = db->locs.as<uint64_t>(); f.n = n; f.done = part.done.as<uint8_t>(); f.left_over = left;
    f.left_list = part.left_list.as<uint32_t>();
    // feature slots per lane: 2 (reads of <= 64 slots, e.g. 2 x 150 bp at sketchlen 16) keeps the
    // registers at the level of 6 resident CTAs; longer reads take the 4-slot variant (a read with more
    // slots than its variant holds is left to the staged path: the choice only affects speed)
    const bool nc2 = slots_per_read <= 64.0;
    unsigned g = 1;
    int per_sm = 1;
    // matches per read: 128 unless the lists of this database are long (a read pair of 64 features then gathers
    // more than 128 locations as a rule and the staged path would take over the batch)
    static const int cap_env = getenv("MCG_FUSED_CAP") ? atoi(getenv("MCG_FUSED_CAP")) : 0;
    const bool cap256 = cap_env ? cap_env >= 256 : (double(db->nlocs) > 2.1 * double(db->nkeys));
    // shared memory comes out of L1, which the probes live on: ask for no more than the 5 CTAs need
    static const int carve_env = getenv("MCG_FUSED_CARVE") ? atoi(getenv("MCG_FUSED_CARVE")) : -1;
    const int carve = carve_env >= 0 ? carve_env : (cap256 ? 56 : 36);
#define QF_LAUNCH(KT, NC) do { if(cap256) QF_LAUNCH_CAP(KT, NC, 256); else QF_LAUNCH_CAP(KT, NC, Q_FUSED); } while(0)
#define QF_LAUNCH_CAP(KT, NC, CAP)                                                                                   \
    do {                                                                                                             \
        auto kern = q_fused_warp_kernel<KT, NC, CAP>;                                                                \
        const int smem = 8 * CAP * (int(sizeof(KT)) + 8);                                                            \
        MCG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));                     \
        if(carve >= 0) MCG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve));    \
        MCG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem));                           \
        g = unsigned(std::min<uint64_t>(uint64_t(n + 63) / 64, uint64_t(c->num_sms) * uint64_t(std::max(1, per_sm)))); \
        kern<<<g, 256, smem, c->stream>>>(p, f);                                                                     \
    } while(0)
    c->prof_begin(1, c->stream);
    if(db->key32) { if(nc2) QF_LAUNCH(uint32_t, 2); else QF_LAUNCH(uint32_t, 4); }
    else          { if(nc2) QF_LAUNCH(uint64_t, 2); else QF_LAUNCH(uint64_t, 4); }
    c->prof_end(1, c->stream);
#undef QF_LAUNCH
#undef QF_LAUNCH_CAP
    c->launches++;
    MCG_CUDA(cudaGetLastError());
    // short reads it left over: finished by the spill kernel, without the host
    {
        const unsigned g = unsigned(c->num_sms);
        const size_t cap = size_t(Q_SPILL_SLOTS) * size_t(db->maxlocs);
        const size_t per_warp = cap * ((db->key32 ? 4 : 8) + 12);
        MCG_TRY(c->q_spill.reserve(size_t(g) * Q_SPILL_WARPS * per_warp));
        MCG_TRY(part.left_list2.reserve(size_t(n + 2) * 4));
        SpillParams sp;
        sp.feat = part.feat.as<uint32_t>(); sp.table = db->table.as<uint32_t>(); sp.geom = db->geom(); sp.locs = db->locs.as<uint64_t>();
        sp.done = part.done.as<uint8_t>(); sp.left_count = left; sp.left_list = part.left_list.as<uint32_t>();
        sp.left_count2 = left + 2; sp.left_list2 = part.left_list2.as<uint32_t>();
        sp.scratch = c->q_spill.as<unsigned char>(); sp.cap = uint32_t(cap);
        if(db->key32) q_spill_warp_kernel<uint32_t><<<g, Q_SPILL_WARPS * 32, 0, c->stream>>>(p, sp);
        else          q_spill_warp_kernel<uint64_t><<<g, Q_SPILL_WARPS * 32, 0, c->stream>>>(p, sp);
        c->launches++;
        MCG_CUDA(cudaGetLastError());
    }
    MCG_CUDA(cudaMemcpyAsync(c->h_left + slot, left + 2, 4, cudaMemcpyDeviceToHost, c->stream));
    MCG_CUDA(cudaEventRecord(part.evA, c->stream));
    return MCG_OK;
}

// stage B, first half: waits for the part's fused kernel and sends the reads it left over through the
// staged path (find -> scan -> binning -> gather -> size-binned candidate kernels).  Synchronises.
static int qb_leftovers(mcg_db* db, const mcg_query_opts* o, QPart& part, int slot, int64_t n, const OutRows& out)
{
    mcg_ctx* c = db->ctx;
    MCG_CUDA(cudaEventSynchronize(part.evA));
    const uint32_t h_left = c->h_left[slot];
    static const bool trace = getenv("MCG_TRACE") != nullptr;
    if(trace) fprintf(stderr, "[mcg query] %lld reads, left over by the fused kernel: %u\n", (long long)n, h_left);
    if(h_left == 0) return MCG_OK;
    QueryRun run;
    const uint8_t* d_done = nullptr;
    int64_t n_staged = n;
    const uint32_t* W = part.W.as<uint32_t>();
    static const bool compact_off = getenv("MCG_QUERY_NO_COMPACT") != nullptr;
    if(uint64_t(h_left) * 8 <= uint64_t(n) && !compact_off) {
        // few reads left: the staged path runs over a sub-batch of just these reads (its kernels
        // are per feature slot of the batch they are given)
        MCG_TRY(c->s_wc.reserve(size_t(h_left + 2) * 4));
        uint32_t* Wc = c->s_wc.as<uint32_t>();
        const uint32_t* left_list = part.left_list2.as<uint32_t>();
        q_left_counts_kernel<<<grid_for(h_left), 256, 0, c->stream>>>(W, left_list, h_left, Wc);
        c->launches++;
        MCG_TRY(exclusive_scan_u32_to_u32(c, Wc, Wc, h_left));
        uint32_t wtot = 0;
        MCG_CUDA(cudaMemcpyAsync(&wtot, Wc + h_left, 4, cudaMemcpyDeviceToHost, c->stream));
        MCG_CUDA(cudaStreamSynchronize(c->stream));
        const uint64_t nsl = uint64_t(wtot) * uint64_t(o->sketchlen);
        MCG_TRY(c->s_featc.reserve((nsl + 1) * 4));
        q_left_copy_kernel<<<grid_for(uint64_t(h_left) * 32), 256, 0, c->stream>>>(part.feat.as<uint32_t>(), W, left_list,
                                                                                 h_left, Wc, o->sketchlen, c->s_featc.as<uint32_t>());
        c->launches++;
        run.feat = c->s_featc.as<uint32_t>();
        run.W = Wc;
        run.rmap = left_list;
        run.nwin = wtot;
        run.nslots = nsl;
        n_staged = int64_t(h_left);
    } else {
        uint32_t wtot = 0;
        MCG_CUDA(cudaMemcpyAsync(&wtot, W + n, 4, cudaMemcpyDeviceToHost, c->stream));
        MCG_CUDA(cudaStreamSynchronize(c->stream));
        d_done = part.done.as<uint8_t>();
        q_hide_done_kernel<<<grid_for(uint64_t(n) * 32), 256, 0, c->stream>>>(part.feat.as<uint32_t>(), W, o->sketchlen, d_done, n);
        c->launches++;
        run.feat = part.feat.as<uint32_t>();
        run.W = W;
        run.rmap = nullptr;
        run.nwin = wtot;
        run.nslots = uint64_t(wtot) * uint64_t(o->sketchlen);
    }
    MCG_TRY(query_find_gather(db, o, n_staged, out.tax, out.hits, out.n, out.rs, out.ns, d_done, &run));
    MCG_TRY(query_back(db, o, part.meta.as<uint2>(), n_staged, run, out, 0));
    return MCG_OK;
}

static double est_slots_per_read(const mcg_query_opts* o, int64_t bytes1, int64_t bytes2, int64_t n, bool paired)
{
    if(n <= 0) return 0.0;
    auto est = [&](int64_t bytes) { const int64_t avg = (bytes + n - 1) / n; return double(num_windows(avg, o->winlen, o->winstride)); };
    return double(o->sketchlen) * (est(bytes1) + (paired ? est(bytes2) : 0.0));
}

// one part, device in / device out, per-shard lists (the body of mcg_query_batch_dev)
int query_batch_dev(mcg_db* db, const mcg_query_opts* o,
                    const char* d1, const int64_t* doffs1, int64_t bytes1,
                    const char* d2, const int64_t* doffs2, int64_t bytes2, int64_t n, int64_t max_len,
                    int32_t* d_out_tax, uint32_t* d_out_hits, int32_t* d_out_n, int slot, bool finish)
{
    mcg_ctx* c = db->ctx;
    QPart& part = c->qp[slot];
    const int s = o->sketchlen;
    const uint64_t wb1 = windows_bound(bytes1, n, o->winstride, o->winlen, max_len);
    const uint64_t wb2 = d2 ? windows_bound(bytes2, n, o->winstride, o->winlen, max_len) : 0;
    MCG_REQUIRE((wb1 + wb2) * uint64_t(s) < (1ull << 32), "too many feature slots in one query part");
    MCG_TRY(part.W.reserve(size_t(n + 2) * 4));
    MCG_TRY(part.meta.reserve(size_t(n + 1) * 8));
    MCG_TRY(part.feat.reserve(((wb1 + wb2) * s + 65) * 4));       // (+ 64: the fused kernel prefetches one read ahead)
    MCG_TRY(part.ctr.reserve(64));
    MCG_TRY(qa_sketch(db, o, d1, doffs1, bytes1, d2, doffs2, bytes2, n, max_len, part.W.as<uint32_t>(), part.meta.as<uint2>(),
                      part.feat.as<uint32_t>(), wb1, wb2, part.ctr.as<uint32_t>()));
    const OutRows out = OutRows::separate(d_out_tax, d_out_hits, d_out_n, o->maxcand);
    MCG_TRY(qa_fused(db, o, part, slot, n, est_slots_per_read(o, bytes1, bytes2, n, d2 != nullptr), out, true));
    if(finish) MCG_TRY(qb_leftovers(db, o, part, slot, n, out));
    return MCG_OK;
}

// the debugging seam of mcg_query_matches: everything staged, sorted match lists written back
static int query_staged_sorted(mcg_db* db, const mcg_query_opts* o,
                               const char* d1, const int64_t* doffs1, int64_t bytes1,
                               const char* d2, const int64_t* doffs2, int64_t bytes2, int64_t n,
                               int32_t* d_out_tax, uint32_t* d_out_hits, int32_t* d_out_n, QueryRun* run_out)
{
    mcg_ctx* c = db->ctx;
    QPart& part = c->qp[0];
    const int s = o->sketchlen, K = o->maxcand;
    const uint64_t wb1 = windows_bound(bytes1, n, o->winstride, o->winlen, 0);
    const uint64_t wb2 = d2 ? windows_bound(bytes2, n, o->winstride, o->winlen, 0) : 0;
    MCG_REQUIRE((wb1 + wb2) * uint64_t(s) < (1ull << 32), "too many feature slots in one query part");
    MCG_TRY(part.W.reserve(size_t(n + 2) * 4));
    MCG_TRY(part.meta.reserve(size_t(n + 1) * 8));
    MCG_TRY(part.feat.reserve(((wb1 + wb2) * s + 65) * 4));       // (+ 64: the fused kernel prefetches one read ahead)
    MCG_CUDA(cudaMemsetAsync(d_out_tax, 0xFF, size_t(n) * K * 4, c->stream));
    MCG_CUDA(cudaMemsetAsync(d_out_hits, 0, size_t(n) * K * 4, c->stream));
    MCG_TRY(qa_sketch(db, o, d1, doffs1, bytes1, d2, doffs2, bytes2, n, 0, part.W.as<uint32_t>(), part.meta.as<uint2>(),
                      part.feat.as<uint32_t>(), wb1, wb2));
    uint32_t wtot = 0;
    MCG_CUDA(cudaMemcpyAsync(&wtot, part.W.as<uint32_t>() + n, 4, cudaMemcpyDeviceToHost, c->stream));
    MCG_CUDA(cudaStreamSynchronize(c->stream));
    QueryRun run;
    run.feat = part.feat.as<uint32_t>();
    run.W = part.W.as<uint32_t>();
    run.rmap = nullptr;
    run.nwin = wtot;
    run.nslots = uint64_t(wtot) * uint64_t(s);
    const OutRows out = OutRows::separate(d_out_tax, d_out_hits, d_out_n, K);
    MCG_TRY(query_find_gather(db, o, n, out.tax, out.hits, out.n, out.rs, out.ns, nullptr, &run));
    MCG_TRY(query_back(db, o, part.meta.as<uint2>(), n, run, out, 1));
    if(run_out) *run_out = run;
    return MCG_OK;
}

// host reads -> device staging (reads1 | reads2 share one buffer)
static int stage_reads(mcg_ctx* c, const char* reads1, const int64_t* offs1, const char* reads2, const int64_t* offs2, int64_t n,
                       const char** d1, const int64_t** do1, int64_t* b1, const char** d2, const int64_t** do2, int64_t* b2)
{
    MCG_REQUIRE(offs1 != nullptr && offs1[0] == 0, "offs1[0] must be 0");
    *b1 = offs1[n];
    *b2 = 0;
    if(reads2 != nullptr) {
        MCG_REQUIRE(offs2 != nullptr && offs2[0] == 0, "offs2[0] must be 0");
        *b2 = offs2[n];
    }
    const size_t pad1 = (size_t(*b1) + 63) & ~size_t(63);
    MCG_TRY(c->seq.reserve(pad1 + size_t(*b2) + 64));
    MCG_TRY(c->seqoff.reserve(size_t(2) * size_t(n + 1) * 8));
    MCG_TRY(copy_in(c, c->seq.p, reads1, size_t(*b1)));
    MCG_TRY(copy_in(c, c->seqoff.p, offs1, size_t(n + 1) * 8));
    *d1 = c->seq.as<char>();
    *do1 = c->seqoff.as<int64_t>();
    *d2 = nullptr;
    *do2 = nullptr;
    if(reads2 != nullptr) {
        MCG_TRY(copy_in(c, c->seq.as<char>() + pad1, reads2, size_t(*b2)));
        MCG_TRY(copy_in(c, c->seqoff.as<int64_t>() + (n + 1), offs2, size_t(n + 1) * 8));
        *d2 = c->seq.as<char>() + pad1;
        *do2 = c->seqoff.as<int64_t>() + (n + 1);
    }
    return MCG_OK;
}

// ------------------------------------------------------------------ sharded query (all ranks, one call)
// A slice of n reads is cut into np parts at part_cut(n, np, 0..np).  A call costs the PCIe copy of its first part
// (nothing to overlap it with) plus the kernels of its last part: with three or more parts the first and the last
// are half as long as the ones in between.
__host__ __device__ __forceinline__ int64_t part_cut(int64_t n, int np, int j)
{
    if(j <= 0) return 0;
    if(j >= np) return n;
    return np >= 3 ? n * (2 * j - 1) / (2 * (np - 1)) : n * j / np;
}

// window totals of the parts of a slice
__global__ void q_plan_kernel(const uint32_t* W, int64_t n, int np, uint64_t* out)
{
    const int j = threadIdx.x;
    if(j < np) out[1 + j] = uint64_t(W[part_cut(n, np, j + 1)] - W[part_cut(n, np, j)]);
    if(j == 0) out[0] = uint64_t(n);
}

constexpr int QS_MAX_PARTS = 16;

struct ShardPlan {
    int P = 1, me = 0, np = 1;                    // ranks, own rank, parts
    std::vector<uint64_t> n, w;                   // [np][P] reads / windows of part j on rank r
    uint64_t rn(int j, int r) const { return n[size_t(j) * size_t(P) + size_t(r)]; }
    uint64_t rw(int j, int r) const { return w[size_t(j) * size_t(P) + size_t(r)]; }
};

// sizes of every part on every rank (the only thing the ranks have to agree on before they enqueue).
// P == 1: nothing is exchanged and nothing returns to the host (window counts are bounded instead).
static int qs_plan(mcg_db* db, mcg_comm* comm, const mcg_query_opts* o, const int64_t* doffs1, const int64_t* doffs2,
                   int64_t n_loc, int np_want, ShardPlan* pl)
{
    mcg_ctx* c = db->ctx;
    const int P = comm_size(comm), me = comm_rank(comm);
    pl->P = P; pl->me = me;
    if(P == 1) {
        pl->np = np_want;
        pl->n.assign(size_t(np_want), 0); pl->w.assign(size_t(np_want), 0);
        for(int j = 0; j < np_want; ++j) pl->n[size_t(j)] = uint64_t(part_cut(n_loc, np_want, j + 1) - part_cut(n_loc, np_want, j));
        return MCG_OK;                                        // w stays unknown (0): callers use bounds
    }
    const int NP = np_want;
    MCG_TRY(c->tmp0.reserve(size_t(n_loc + 2) * 4)); MCG_TRY(c->tmp1.reserve(size_t(n_loc + 2) * 4));
    MCG_TRY(c->tmp5.reserve(size_t(n_loc + 2) * 4)); MCG_TRY(c->tmp4.reserve(size_t(n_loc + 2) * 8));
    MCG_TRY(c->plan.reserve(size_t(QS_MAX_PARTS + 2) * 8));
    uint32_t* Ww = c->tmp5.as<uint32_t>();
    if(n_loc > 0) {
        MCG_TRY(launch_prep(db, o, doffs1, doffs2, n_loc, c->tmp0.as<uint32_t>(), doffs2 ? c->tmp1.as<uint32_t>() : nullptr, Ww, nullptr,
                            c->tmp4.as<uint2>(), nullptr));
    } else {
        MCG_CUDA(cudaMemsetAsync(Ww, 0, 8, c->stream));
    }
    q_plan_kernel<<<1, 32, 0, c->stream>>>(Ww, n_loc, NP, c->plan.as<uint64_t>());
    c->launches++;
    std::vector<uint64_t> all(size_t(P) * size_t(NP + 1), 0);
    MCG_TRY(comm_allgather_dev_u64(comm, c->plan.as<uint64_t>(), NP + 1, all.data()));
    uint64_t nmax = 0;
    for(int r = 0; r < P; ++r) nmax = std::max(nmax, all[size_t(r) * size_t(NP + 1)]);
    static const uint64_t min_part = getenv("MCG_SHARD_MIN_PART") ? uint64_t(atoll(getenv("MCG_SHARD_MIN_PART"))) : 16384;
    const bool collapse = nmax < min_part * uint64_t(NP);        // small calls: one part (same decision on every rank)
    pl->np = collapse ? 1 : NP;
    pl->n.assign(size_t(pl->np) * size_t(P), 0); pl->w.assign(size_t(pl->np) * size_t(P), 0);
    for(int r = 0; r < P; ++r) {
        const uint64_t* a = all.data() + size_t(r) * size_t(NP + 1);
        const int64_t nr = int64_t(a[0]);
        for(int j = 0; j < NP; ++j) {
            const int jj = collapse ? 0 : j;
            pl->n[size_t(jj) * size_t(P) + size_t(r)] += uint64_t(part_cut(nr, NP, j + 1) - part_cut(nr, NP, j));
            pl->w[size_t(jj) * size_t(P) + size_t(r)] += a[1 + j];
        }
    }
    return MCG_OK;
}

// stage A of part j: local K1, feature / meta exchange, fused kernel over all reads of the part
static int qs_stage_a(mcg_db* db, mcg_comm* comm, const mcg_query_opts* o, const ShardPlan& pl, int j, int slot,
                      const char* d1, const int64_t* do1, int64_t b1, const char* d2, const int64_t* do2, int64_t b2,
                      int64_t max_len)
{
    mcg_ctx* c = db->ctx;
    QPart& part = c->qp[slot];
    const int P = pl.P, me = pl.me, s = o->sketchlen, K = o->maxcand;
    const int64_t n_loc = int64_t(pl.rn(j, me));
    uint64_t N = 0, Wt = 0, nb = 0, wb = 0;
    for(int r = 0; r < P; ++r) { if(r == me) { nb = N; wb = Wt; } N += pl.rn(j, r); Wt += pl.rw(j, r); }
    const uint64_t wb1 = windows_bound(b1, n_loc, o->winstride, o->winlen, max_len);
    const uint64_t wb2 = d2 ? windows_bound(b2, n_loc, o->winstride, o->winlen, max_len) : 0;
    if(P == 1) Wt = wb1 + wb2;                                 // bound instead of an exchanged exact count
    MCG_REQUIRE((Wt + 1) * uint64_t(s) < (1ull << 32) && N < (1ull << 31), "too many reads / feature slots in one query part");
    MCG_TRY(part.W.reserve(size_t(N + 2) * 4));
    MCG_TRY(part.meta.reserve(size_t(N + 1) * 8));
    MCG_TRY(part.feat.reserve((Wt * s + 65) * 4));
    MCG_TRY(part.lt.reserve(size_t(N + 1) * size_t(2 * K + 1) * 4));       // one packed record per read: {tax[K], hits[K], n}
    uint32_t* Wl = part.W.as<uint32_t>();
    if(P > 1) { MCG_TRY(c->tmp2.reserve(size_t(n_loc + 2) * 4)); Wl = c->tmp2.as<uint32_t>(); }
    c->prof_begin(3, c->stream);
    MCG_TRY(part.ctr.reserve(64));
    if(n_loc > 0)
        MCG_TRY(qa_sketch(db, o, d1, do1, b1, d2, do2, b2, n_loc, max_len, Wl, part.meta.as<uint2>() + nb,
                          part.feat.as<uint32_t>() + wb * uint64_t(s), wb1, wb2, part.ctr.as<uint32_t>()));
    c->prof_end(3, c->stream);
    if(P > 1) {
        c->prof_begin(4, c->stream);
        std::vector<uint64_t> off(static_cast<size_t>(P)), len(static_cast<size_t>(P));
        uint64_t a = 0;
        MCG_TRY(comm_group_begin(comm));                       // sketches + per-read meta: one NCCL group
        for(int r = 0; r < P; ++r) { off[size_t(r)] = a * uint64_t(s) * 4; len[size_t(r)] = pl.rw(j, r) * uint64_t(s) * 4; a += pl.rw(j, r); }
        MCG_TRY(comm_allgatherv_inplace(comm, part.feat.p, off.data(), len.data(), c->stream));
        a = 0;
        for(int r = 0; r < P; ++r) { off[size_t(r)] = a * 8; len[size_t(r)] = pl.rn(j, r) * 8; a += pl.rn(j, r); }
        MCG_TRY(comm_allgatherv_inplace(comm, part.meta.p, off.data(), len.data(), c->stream));
        MCG_TRY(comm_group_end(comm));
        if(N > 0) {
            q_meta_windows_kernel<<<grid_for(N), 256, 0, c->stream>>>(part.meta.as<uint2>(), int64_t(N), part.W.as<uint32_t>());
            c->launches++;
        }
        MCG_TRY(exclusive_scan_u32_to_u32(c, part.W.as<uint32_t>(), part.W.as<uint32_t>(), N));
        c->prof_end(4, c->stream);
    }
    if(N == 0) { c->h_left[slot] = 0; MCG_CUDA(cudaEventRecord(part.evA, c->stream)); return MCG_OK; }
    const double spr = est_slots_per_read(o, b1, b2, std::max<int64_t>(n_loc, 1), d2 != nullptr);
    return qa_fused(db, o, part, slot, int64_t(N), n_loc > 0 ? spr : 64.0, OutRows::packed(part.lt.as<uint32_t>(), K), n_loc > 0);
}

// stage B of part j: left-over reads, exchange of the per-read lists, merge + decision of the own block
static int qs_stage_b(mcg_db* db, mcg_comm* comm, const mcg_query_opts* o, const ShardPlan& pl, int j, int slot,
                      int32_t* d_out_tax, uint32_t* d_out_hits, int32_t* d_out_n, int32_t* d_out_best, int32_t* d_out_brank,
                      uint64_t* d_stats)
{
    mcg_ctx* c = db->ctx;
    QPart& part = c->qp[slot];
    const int P = pl.P, me = pl.me, K = o->maxcand;
    uint64_t N = 0, nb = 0;
    for(int r = 0; r < P; ++r) { if(r == me) nb = N; N += pl.rn(j, r); }
    const uint64_t n_me = pl.rn(j, me);
    if(N > 0) {
        MCG_CUDA(cudaEventSynchronize(part.evA));
        const bool left = c->h_left[slot] != 0;
        if(left) c->prof_begin(5, c->stream);
        MCG_TRY(qb_leftovers(db, o, part, slot, int64_t(N), OutRows::packed(part.lt.as<uint32_t>(), K)));
        if(left) c->prof_end(5, c->stream);
    }
    const uint64_t row = uint64_t(2 * K + 1) * 4;               // bytes of a packed record
    const uint32_t* recs = part.lt.as<uint32_t>();              // P == 1: the own records, in place
    if(P > 1) {
        c->prof_begin(6, c->stream);
        MCG_TRY(part.ct.reserve(size_t(P) * (n_me + 1) * row));
        std::vector<uint64_t> so(static_cast<size_t>(P)), sl(static_cast<size_t>(P)), ro(static_cast<size_t>(P)), rl(static_cast<size_t>(P));
        uint64_t a = 0;
        for(int r = 0; r < P; ++r) {
            so[size_t(r)] = a * row; sl[size_t(r)] = pl.rn(j, r) * row; a += pl.rn(j, r);
            ro[size_t(r)] = uint64_t(r) * n_me * row; rl[size_t(r)] = n_me * row;
        }
        MCG_TRY(comm_alltoallv(comm, part.lt.p, so.data(), sl.data(), part.ct.p, ro.data(), rl.data(), c->stream));
        recs = part.ct.as<uint32_t>();
        c->prof_end(6, c->stream);
    }
    (void)nb;
    c->prof_begin(7, c->stream);
    if(n_me > 0)
        MCG_TRY(merge_classify_rows(c, o, reinterpret_cast<const int32_t*>(recs), recs + K, reinterpret_cast<const int32_t*>(recs + 2 * K),
                                    uint32_t(2 * K + 1), uint32_t(2 * K + 1), P, int64_t(n_me), d_out_tax, d_out_hits, d_out_n, d_out_best,
                                    d_out_brank, d_stats));
    c->prof_end(7, c->stream);
    MCG_CUDA(cudaEventRecord(part.evB, c->stream));
    return MCG_OK;
}

}  // namespace mcg

using namespace mcg;

extern "C" {

int mcg_tax_upload(mcg_ctx* c, const int32_t* lineage, const uint8_t* rank, int64_t ntaxa, int64_t ntargets)
{
    MCG_REQUIRE(c != nullptr && lineage != nullptr && rank != nullptr, "mcg_tax_upload: NULL argument");
    MCG_REQUIRE(ntaxa >= 1 && ntargets >= 0 && ntargets <= ntaxa && ntaxa < (1ll << 31), "bad taxon counts");
    for(int64_t t = 0; t < ntaxa; ++t) {
        MCG_REQUIRE(rank[t] <= MCG_RANK_NONE, "taxon %lld has rank %d", (long long)t, int(rank[t]));
        for(int r = 0; r < MCG_NUM_RANKS; ++r) {
            const int32_t x = lineage[t * MCG_NUM_RANKS + r];
            MCG_REQUIRE(x >= -1 && x < ntaxa, "lineage entry out of range");
        }
    }
    MCG_CUDA(cudaSetDevice(c->device));
    MCG_TRY(c->lineage.reserve(size_t(ntaxa) * MCG_NUM_RANKS * 4));
    MCG_TRY(c->rank.reserve(size_t(ntaxa)));
    MCG_TRY(copy_in(c, c->lineage.p, lineage, size_t(ntaxa) * MCG_NUM_RANKS * 4));
    MCG_TRY(copy_in(c, c->rank.p, rank, size_t(ntaxa)));
    MCG_CUDA(cudaStreamSynchronize(c->stream));
    c->ntaxa = ntaxa;
    c->ntargets = ntargets;
    c->lift_lowest = -1;
    return MCG_OK;
}

int mcg_query_batch_dev(mcg_db* db, const mcg_query_opts* o,
                        const char* d_reads1, const int64_t* d_offs1, int64_t bytes1,
                        const char* d_reads2, const int64_t* d_offs2, int64_t bytes2,
                        int64_t n, int64_t max_read_len,
                        int32_t* d_out_tax, uint32_t* d_out_hits, int32_t* d_out_ncand)
{
    MCG_REQUIRE(db != nullptr && o != nullptr && d_reads1 != nullptr && d_offs1 != nullptr, "mcg_query_batch_dev: NULL argument");
    MCG_REQUIRE(n >= 0 && n < (1ll << 31), "read count out of range");
    MCG_CUDA(cudaSetDevice(db->ctx->device));
    MCG_TRY(check_ready(db, o));
    if(n == 0) return MCG_OK;
    MCG_TRY(ensure_pipeline(db->ctx));
    return query_batch_dev(db, o, d_reads1, d_offs1, bytes1, d_reads2, d_offs2, d_reads2 ? bytes2 : 0, n, max_read_len,
                           d_out_tax, d_out_hits, d_out_ncand, 0, true);
}

// Host path.  The batch is cut into parts: while part i is queried (ctx stream), the reads of part
// i+1 are copied in (H2D stream, second staging buffer) and the top lists of part i-1 travel back
// (D2H stream) -- PCIe is full duplex, so the call costs about the H2D copy of the reads instead of
// H2D + query + D2H.  The host looks at the left-over count of part i only after part i+1 is enqueued.
int mcg_query_batch(mcg_db* db, const mcg_query_opts* o,
                    const char* reads1, const int64_t* offs1,
                    const char* reads2, const int64_t* offs2, int64_t n,
                    int32_t* out_tax, uint32_t* out_hits, int32_t* out_ncand)
{
    MCG_REQUIRE(db != nullptr && o != nullptr && reads1 != nullptr && offs1 != nullptr, "mcg_query_batch: NULL argument");
    MCG_REQUIRE(n >= 0, "n < 0");
    if(n == 0) return MCG_OK;
    MCG_REQUIRE(out_tax && out_hits && out_ncand, "output buffers are NULL");
    MCG_REQUIRE(offs1[0] == 0, "offs1[0] must be 0");
    if(reads2 != nullptr) MCG_REQUIRE(offs2 != nullptr && offs2[0] == 0, "offs2[0] must be 0");
    mcg_ctx* c = db->ctx;
    MCG_CUDA(cudaSetDevice(c->device));
    MCG_TRY(check_ready(db, o));
    MCG_TRY(ensure_pipeline(c));
    const int K = o->maxcand;
    static int64_t sub_reads = -1;
    if(sub_reads < 0) { const char* e = getenv("MCG_QUERY_SUB"); sub_reads = e ? atoll(e) : 262144; if(sub_reads < 1) sub_reads = 1; }
    const int64_t nsub = (n + sub_reads - 1) / sub_reads;
    // staging: two buffers, each [reads1 | pad | reads2 | slack]; offsets [2 buffers][2 mates][sub+1]
    int64_t maxb1 = 0, maxb2 = 0;
    for(int64_t i = 0; i < nsub; ++i) {
        const int64_t a = i * sub_reads, b = std::min(n, a + sub_reads);
        maxb1 = std::max(maxb1, offs1[b] - offs1[a]);
        if(reads2) maxb2 = std::max(maxb2, offs2[b] - offs2[a]);
    }
    const size_t pad1 = (size_t(maxb1) + 127) & ~size_t(63);
    const size_t bufsz = pad1 + ((size_t(maxb2) + 127) & ~size_t(63));
    const size_t offsz = size_t(std::min(n, sub_reads) + 1);
    MCG_TRY(c->seq.reserve(2 * bufsz + 64));
    MCG_TRY(c->seqoff.reserve(4 * offsz * 8));
    MCG_TRY(c->q_out_tax.reserve(size_t(n) * K * 4));
    MCG_TRY(c->q_out_hits.reserve(size_t(n) * K * 4));
    MCG_TRY(c->q_out_n.reserve(size_t(n) * 4));
    MCG_CUDA(cudaEventRecord(c->ev_misc, c->stream));
    MCG_CUDA(cudaStreamWaitEvent(c->s_h2d, c->ev_misc, 0));     // staging buffers are ours from here on
    MCG_CUDA(cudaStreamWaitEvent(c->s_d2h, c->ev_misc, 0));
    auto stage = [&](int64_t i) -> int {
        const int slot = int(i & 1);
        const int64_t a = i * sub_reads, b = std::min(n, a + sub_reads);
        char* dbuf = c->seq.as<char>() + size_t(slot) * bufsz;
        int64_t* doff = c->seqoff.as<int64_t>() + size_t(slot) * 2 * offsz;
        if(i >= 2) MCG_CUDA(cudaStreamWaitEvent(c->s_h2d, c->qp[slot].evA, 0));     // stage A of part i-2 has consumed the buffer
        // the (small) offset arrays first: they usually come from pageable memory, whose copy blocks the
        // host until everything queued before it on the stream is done
        MCG_CUDA(cudaMemcpyAsync(doff, offs1 + a, size_t(b - a + 1) * 8, cudaMemcpyHostToDevice, c->s_h2d));
        if(reads2) MCG_CUDA(cudaMemcpyAsync(doff + offsz, offs2 + a, size_t(b - a + 1) * 8, cudaMemcpyHostToDevice, c->s_h2d));
        MCG_CUDA(cudaMemcpyAsync(dbuf, reads1 + offs1[a], size_t(offs1[b] - offs1[a]), cudaMemcpyHostToDevice, c->s_h2d));
        if(reads2) MCG_CUDA(cudaMemcpyAsync(dbuf + pad1, reads2 + offs2[a], size_t(offs2[b] - offs2[a]), cudaMemcpyHostToDevice, c->s_h2d));
        MCG_CUDA(cudaEventRecord(c->ev_h2d[slot], c->s_h2d));
        return MCG_OK;
    };
    // stage B of part i: left-over reads, then the lists travel back
    auto finish = [&](int64_t i) -> int {
        const int slot = int(i & 1);
        const int64_t a = i * sub_reads, b = std::min(n, a + sub_reads);
        MCG_TRY(qb_leftovers(db, o, c->qp[slot], slot, b - a, OutRows::separate(c->q_out_tax.as<int32_t>() + a * K, c->q_out_hits.as<uint32_t>() + a * K,
                                                                                   c->q_out_n.as<int32_t>() + a, K)));
        MCG_CUDA(cudaEventRecord(c->ev_k[slot], c->stream));
        MCG_CUDA(cudaStreamWaitEvent(c->s_d2h, c->ev_k[slot], 0));
        MCG_CUDA(cudaMemcpyAsync(out_tax + a * K, c->q_out_tax.as<int32_t>() + a * K, size_t(b - a) * K * 4, cudaMemcpyDeviceToHost, c->s_d2h));
        MCG_CUDA(cudaMemcpyAsync(out_hits + a * K, c->q_out_hits.as<uint32_t>() + a * K, size_t(b - a) * K * 4, cudaMemcpyDeviceToHost, c->s_d2h));
        MCG_CUDA(cudaMemcpyAsync(out_ncand + a, c->q_out_n.as<int32_t>() + a, size_t(b - a) * 4, cudaMemcpyDeviceToHost, c->s_d2h));
        return MCG_OK;
    };
    MCG_TRY(stage(0));
    for(int64_t i = 0; i < nsub; ++i) {
        const int slot = int(i & 1);
        const int64_t a = i * sub_reads, b = std::min(n, a + sub_reads);
        // stage A of part i is enqueued BEFORE the host waits for part i-1's left-over count (finish): the GPU
        // never idles on the host.  (Stage B reads no staging buffer: the per-read meta carries what it needs.)
        if(i + 1 < nsub) MCG_TRY(stage(i + 1));
        MCG_CUDA(cudaStreamWaitEvent(c->stream, c->ev_h2d[slot], 0));
        // the offsets keep their values of the whole batch: the base pointers are moved instead
        const char* d1 = c->seq.as<char>() + size_t(slot) * bufsz - offs1[a];
        const int64_t* do1 = c->seqoff.as<int64_t>() + size_t(slot) * 2 * offsz;
        const char* d2 = reads2 ? c->seq.as<char>() + size_t(slot) * bufsz + pad1 - offs2[a] : nullptr;
        const int64_t* do2 = reads2 ? do1 + offsz : nullptr;
        MCG_TRY(query_batch_dev(db, o, d1, do1, offs1[b] - offs1[a], d2, do2, reads2 ? offs2[b] - offs2[a] : 0, b - a, 0,
                                c->q_out_tax.as<int32_t>() + a * K, c->q_out_hits.as<uint32_t>() + a * K,
                                c->q_out_n.as<int32_t>() + a, slot, false));
        if(i >= 1) MCG_TRY(finish(i - 1));
    }
    MCG_TRY(finish(nsub - 1));
    MCG_CUDA(cudaStreamSynchronize(c->s_d2h));
    MCG_CUDA(cudaStreamSynchronize(c->stream));
    return MCG_OK;
}

int mcg_query_matches(mcg_db* db, const mcg_query_opts* o,
                      const char* reads1, const int64_t* offs1,
                      const char* reads2, const int64_t* offs2, int64_t n,
                      int64_t* out_counts, uint32_t* out_tgts, uint32_t* out_wins)
{
    MCG_REQUIRE(db != nullptr && o != nullptr && reads1 != nullptr && offs1 != nullptr && out_counts != nullptr,
                "mcg_query_matches: NULL argument");
    if(n == 0) return MCG_OK;
    mcg_ctx* c = db->ctx;
    MCG_CUDA(cudaSetDevice(c->device));
    MCG_TRY(check_ready(db, o));
    const char *d1, *d2; const int64_t *do1, *do2; int64_t b1, b2;
    MCG_TRY(stage_reads(c, reads1, offs1, reads2, offs2, n, &d1, &do1, &b1, &d2, &do2, &b2));
    const int K = o->maxcand;
    MCG_TRY(c->q_out_tax.reserve(size_t(n) * K * 4));
    MCG_TRY(c->q_out_hits.reserve(size_t(n) * K * 4));
    MCG_TRY(c->q_out_n.reserve(size_t(n) * 4));
    QueryRun run;
    MCG_TRY(query_staged_sorted(db, o, d1, do1, b1, d2, do2, b2, n, c->q_out_tax.as<int32_t>(), c->q_out_hits.as<uint32_t>(),
                                c->q_out_n.as<int32_t>(), &run));
    MCG_TRY(c->m_tax.reserve(size_t(n) * 8));
    q_read_counts_kernel<<<grid_for(n), 256, 0, c->stream>>>(c->q_fpos.as<uint64_t>(), run.W, o->sketchlen, n,
                                                            c->m_tax.as<int64_t>());
    c->launches++;
    MCG_TRY(copy_out(c, out_counts, c->m_tax.p, size_t(n) * 8));
    if(out_tgts != nullptr && out_wins != nullptr && run.total > 0) {
        MCG_TRY(c->m_hits.reserve(run.total * 4));
        MCG_TRY(c->m_n.reserve(run.total * 4));
        if(db->key32)
            q_split_locs_kernel<uint32_t><<<grid_for(run.total), 256, 0, c->stream>>>(c->q_loc.as<uint32_t>(), run.total, db->wbits,
                                                                                     c->m_hits.as<uint32_t>(), c->m_n.as<uint32_t>());
        else
            q_split_locs_kernel<uint64_t><<<grid_for(run.total), 256, 0, c->stream>>>(c->q_loc.as<uint64_t>(), run.total, db->wbits,
                                                                                     c->m_hits.as<uint32_t>(), c->m_n.as<uint32_t>());
        c->launches++;
        MCG_TRY(copy_out(c, out_tgts, c->m_hits.p, run.total * 4));
        MCG_TRY(copy_out(c, out_wins, c->m_n.p, run.total * 4));
    }
    MCG_CUDA(cudaStreamSynchronize(c->stream));
    return MCG_OK;
}

int mcg_merge_classify_dev(mcg_ctx* c, const mcg_query_opts* o,
                           const int32_t* d_cand_tax, const uint32_t* d_cand_hits,
                           const int32_t* d_cand_n, int32_t nshards, int64_t n,
                           int32_t* d_out_tax, uint32_t* d_out_hits, int32_t* d_out_ncand,
                           int32_t* d_out_best, int32_t* d_out_best_rank, uint64_t* d_stats22)
{
    MCG_REQUIRE(c != nullptr && o != nullptr, "mcg_merge_classify_dev: NULL argument");
    return mcg::merge_classify_rows(c, o, d_cand_tax, d_cand_hits, d_cand_n, uint32_t(o->maxcand), 1u, nshards, n, d_out_tax, d_out_hits,
                                    d_out_ncand, d_out_best, d_out_best_rank, d_stats22);
}

}  // extern "C"

namespace mcg {
int merge_classify_rows(mcg_ctx* c, const mcg_query_opts* o, const int32_t* d_cand_tax, const uint32_t* d_cand_hits, const int32_t* d_cand_n,
                        uint32_t rs, uint32_t ns, int32_t nshards, int64_t n,
                        int32_t* d_out_tax, uint32_t* d_out_hits, int32_t* d_out_ncand,
                        int32_t* d_out_best, int32_t* d_out_best_rank, uint64_t* d_stats22)
{
    MCG_REQUIRE(nshards >= 1 && n >= 0, "bad shard / read count");
    MCG_REQUIRE(o->maxcand >= 1 && o->maxcand <= MCG_MAX_CAND, "maxcand %d outside 1..%d", o->maxcand, MCG_MAX_CAND);
    MCG_REQUIRE(o->lowest >= 0 && o->lowest < MCG_NUM_RANKS && o->highest >= 0 && o->highest < MCG_NUM_RANKS, "rank out of range");
    MCG_REQUIRE(o->hitmin >= 0, "hitmin %d is negative", o->hitmin);
    MCG_REQUIRE(!o->reference_compat || nshards <= 64, "reference-compatible merge supports up to 64 shards");
    if(c->lineage.p == nullptr) { set_error("taxonomy not uploaded (mcg_tax_upload)"); return MCG_ERR_STATE; }
    if(n == 0) return MCG_OK;
    MCG_CUDA(cudaSetDevice(c->device));
    MCG_TRY(c->m_bad.reserve(64));
    if(!c->m_bad_clear) { MCG_CUDA(cudaMemsetAsync(c->m_bad.p, 0, 4, c->stream)); c->m_bad_clear = true; }
    MergeParams p;
    p.compat = o->reference_compat; p.ntaxa = int32_t(c->ntaxa); p.ntargets = int32_t(c->ntargets); p.bad = c->m_bad.as<uint32_t>();
    p.cand_tax = d_cand_tax; p.cand_hits = d_cand_hits; p.cand_n = d_cand_n;
    p.rs = rs; p.ns = ns;
    p.P = nshards; p.n = n; p.K = o->maxcand;
    p.lowest = o->lowest; p.highest = o->highest; p.hitmin = o->hitmin; p.hitdiff = o->hitdiff;
    p.lineage = c->lineage.as<int32_t>(); p.rank = c->rank.as<uint8_t>();
    p.out_tax = d_out_tax; p.out_hits = d_out_hits; p.out_n = d_out_ncand;
    p.out_best = d_out_best; p.out_brank = d_out_best_rank;
    p.stats = reinterpret_cast<unsigned long long*>(d_stats22);
    q_merge_classify_kernel<<<grid_for(n, 128), 128, 0, c->stream>>>(p);
    c->launches++;
    MCG_CUDA(cudaGetLastError());
    return MCG_OK;
}
}  // namespace mcg

extern "C" {

int mcg_merge_classify(mcg_ctx* c, const mcg_query_opts* o,
                       const int32_t* cand_tax, const uint32_t* cand_hits,
                       const int32_t* cand_n, int32_t nshards, int64_t n,
                       int32_t* out_tax, uint32_t* out_hits, int32_t* out_ncand,
                       int32_t* out_best, int32_t* out_best_rank, uint64_t* out_stats22)
{
    MCG_REQUIRE(c != nullptr && o != nullptr, "mcg_merge_classify: NULL argument");
    MCG_REQUIRE(nshards >= 1 && n >= 0, "bad shard / read count");
    if(n == 0) return MCG_OK;
    MCG_REQUIRE(o->maxcand >= 1 && o->maxcand <= MCG_MAX_CAND, "maxcand %d outside 1..%d", o->maxcand, MCG_MAX_CAND);
    MCG_CUDA(cudaSetDevice(c->device));
    const size_t K = size_t(o->maxcand), rows = size_t(nshards) * size_t(n);
    MCG_TRY(c->m_tax.reserve(rows * K * 4)); MCG_TRY(c->m_hits.reserve(rows * K * 4)); MCG_TRY(c->m_n.reserve(rows * 4));
    MCG_TRY(c->m_o_tax.reserve(size_t(n) * K * 4)); MCG_TRY(c->m_o_hits.reserve(size_t(n) * K * 4)); MCG_TRY(c->m_o_n.reserve(size_t(n) * 4));
    MCG_TRY(c->m_best.reserve(size_t(n) * 4)); MCG_TRY(c->m_brank.reserve(size_t(n) * 4)); MCG_TRY(c->m_stats.reserve(22 * 8));
    MCG_TRY(copy_in(c, c->m_tax.p, cand_tax, rows * K * 4));
    MCG_TRY(copy_in(c, c->m_hits.p, cand_hits, rows * K * 4));
    MCG_TRY(copy_in(c, c->m_n.p, cand_n, rows * 4));
    MCG_CUDA(cudaMemsetAsync(c->m_stats.p, 0, 22 * 8, c->stream));
    MCG_TRY(mcg_merge_classify_dev(c, o, c->m_tax.as<int32_t>(), c->m_hits.as<uint32_t>(), c->m_n.as<int32_t>(), nshards, n,
                                   c->m_o_tax.as<int32_t>(), c->m_o_hits.as<uint32_t>(), c->m_o_n.as<int32_t>(),
                                   c->m_best.as<int32_t>(), c->m_brank.as<int32_t>(), c->m_stats.as<uint64_t>()));
    if(out_tax) MCG_TRY(copy_out(c, out_tax, c->m_o_tax.p, size_t(n) * K * 4));
    if(out_hits) MCG_TRY(copy_out(c, out_hits, c->m_o_hits.p, size_t(n) * K * 4));
    if(out_ncand) MCG_TRY(copy_out(c, out_ncand, c->m_o_n.p, size_t(n) * 4));
    if(out_best) MCG_TRY(copy_out(c, out_best, c->m_best.p, size_t(n) * 4));
    if(out_best_rank) MCG_TRY(copy_out(c, out_best_rank, c->m_brank.p, size_t(n) * 4));
    uint64_t st[22];
    uint32_t bad = 0;
    MCG_TRY(copy_out(c, st, c->m_stats.p, 22 * 8));
    MCG_TRY(copy_out(c, &bad, c->m_bad.p, 4));
    MCG_CUDA(cudaStreamSynchronize(c->stream));
    if(bad != 0) {
        c->m_bad_clear = false;
        set_error("mcg_merge_classify: %s", (bad & 1u) ? "candidate taxon outside [-1, ntaxa)" : "candidate count outside [0, maxcand]");
        return MCG_ERR_ARG;
    }
    if(out_stats22) for(int i = 0; i < 22; ++i) out_stats22[i] += st[i];
    return MCG_OK;
}


int mcg_query_batch_sharded_dev(mcg_db* db, mcg_comm* comm, const mcg_query_opts* o,
                                const char* d_reads1, const int64_t* d_offs1, int64_t bytes1,
                                const char* d_reads2, const int64_t* d_offs2, int64_t bytes2,
                                int64_t n_local, int64_t max_read_len,
                                int32_t* d_out_tax, uint32_t* d_out_hits, int32_t* d_out_ncand,
                                int32_t* d_out_best, int32_t* d_out_best_rank, uint64_t* d_stats22)
{
    MCG_REQUIRE(db != nullptr && o != nullptr, "mcg_query_batch_sharded_dev: NULL argument");
    MCG_REQUIRE(n_local >= 0 && n_local < (1ll << 31), "read count out of range");
    MCG_REQUIRE(n_local == 0 || (d_reads1 != nullptr && d_offs1 != nullptr), "reads are NULL");
    MCG_REQUIRE(d_out_tax && d_out_hits && d_out_ncand && d_out_best && d_out_best_rank, "device output arrays are NULL");
    mcg_ctx* c = db->ctx;
    MCG_CUDA(cudaSetDevice(c->device));
    MCG_TRY(check_ready(db, o));
    MCG_TRY(ensure_pipeline(c));
    if(comm_size(comm) == 1 && n_local == 0) return MCG_OK;
    ShardPlan pl;
    MCG_TRY(qs_plan(db, comm, o, d_offs1, d_reads2 ? d_offs2 : nullptr, n_local, 1, &pl));
    MCG_TRY(qs_stage_a(db, comm, o, pl, 0, 0, d_reads1, d_offs1, bytes1, d_reads2, d_reads2 ? d_offs2 : nullptr, d_reads2 ? bytes2 : 0, max_read_len));
    return qs_stage_b(db, comm, o, pl, 0, 0, d_out_tax, d_out_hits, d_out_ncand, d_out_best, d_out_best_rank, d_stats22);
}

int mcg_query_batch_sharded(mcg_db* db, mcg_comm* comm, const mcg_query_opts* o,
                            const char* reads1, const int64_t* offs1,
                            const char* reads2, const int64_t* offs2, int64_t n,
                            int32_t* out_tax, uint32_t* out_hits, int32_t* out_ncand,
                            int32_t* out_best, int32_t* out_best_rank, uint64_t* out_stats22)
{
    MCG_REQUIRE(db != nullptr && o != nullptr, "mcg_query_batch_sharded: NULL argument");
    MCG_REQUIRE(n >= 0 && n < (1ll << 31), "read count out of range");
    MCG_REQUIRE(n == 0 || (reads1 != nullptr && offs1 != nullptr), "reads1 / offs1 are NULL");
    if(reads2 != nullptr) MCG_REQUIRE(offs2 != nullptr, "offs2 is NULL");
    mcg_ctx* c = db->ctx;
    MCG_CUDA(cudaSetDevice(c->device));
    MCG_TRY(check_ready(db, o));
    MCG_TRY(ensure_pipeline(c));
    const int P = comm_size(comm);
    if(P == 1 && n == 0) return MCG_OK;
    const int K = o->maxcand;
    static const int64_t sub_reads = [] { const char* e = getenv("MCG_QUERY_SUB"); const int64_t v = e ? atoll(e) : 262144; return v < 1 ? int64_t(1) : v; }();
    // parts per call with P > 1: the slice of a rank is 1/P of the batch, its PCIe copy is short; every part costs
    // launch tails and one NCCL group per exchange: three parts (1/4, 1/2, 1/4 of the slice)
    static const int np_env = getenv("MCG_SHARD_PARTS") ? atoi(getenv("MCG_SHARD_PARTS")) : 3;
    const bool paired = reads2 != nullptr;
    // the offsets of the whole slice go to the device first (the plan needs them), the bases follow part by part
    static const int64_t zero_off[1] = {0};
    if(n == 0) { offs1 = zero_off; offs2 = zero_off; }
    MCG_TRY(c->seqoff.reserve(2 * size_t(n + 1) * 8));
    int64_t* do1 = c->seqoff.as<int64_t>();
    int64_t* do2 = do1 + (n + 1);
    MCG_CUDA(cudaEventRecord(c->ev_misc, c->stream));
    MCG_CUDA(cudaStreamWaitEvent(c->s_h2d, c->ev_misc, 0));     // staging buffers are ours from here on
    MCG_CUDA(cudaStreamWaitEvent(c->s_d2h, c->ev_misc, 0));
    MCG_CUDA(cudaMemcpyAsync(do1, offs1, size_t(n + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    if(paired) MCG_CUDA(cudaMemcpyAsync(do2, offs2, size_t(n + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    int np_want = P == 1 ? int(std::min<int64_t>(QS_MAX_PARTS, n >= 2 * sub_reads ? (n + sub_reads - 1) / sub_reads + 1 : std::max<int64_t>(1, (n + sub_reads - 1) / sub_reads)))
                         : std::max(1, std::min(np_env, QS_MAX_PARTS));
    ShardPlan pl;
    MCG_TRY(qs_plan(db, comm, o, do1, paired ? do2 : nullptr, n, np_want, &pl));
    const int np = pl.np, me = pl.me;
    // part j of the own slice = reads [lo(j), lo(j + 1))
    std::vector<int64_t> lo(size_t(np) + 1, 0);
    for(int j = 0; j < np; ++j) lo[size_t(j) + 1] = lo[size_t(j)] + int64_t(pl.rn(j, me));
    int64_t maxb1 = 0, maxb2 = 0, maxn = 0;
    for(int j = 0; j < np; ++j) {
        maxb1 = std::max(maxb1, offs1[lo[size_t(j) + 1]] - offs1[lo[size_t(j)]]);
        if(paired) maxb2 = std::max(maxb2, offs2[lo[size_t(j) + 1]] - offs2[lo[size_t(j)]]);
        maxn = std::max(maxn, lo[size_t(j) + 1] - lo[size_t(j)]);
    }
    const size_t pad1 = (size_t(maxb1) + 127) & ~size_t(63);
    const size_t bufsz = pad1 + ((size_t(maxb2) + 127) & ~size_t(63));
    MCG_TRY(c->seq.reserve(2 * bufsz + 64));
    MCG_TRY(c->m_stats.reserve(22 * 8));
    MCG_CUDA(cudaMemsetAsync(c->m_stats.p, 0, 22 * 8, c->stream));
    for(int sl = 0; sl < 2; ++sl) {
        QPart& q = c->qp[sl];
        MCG_TRY(q.mt.reserve(size_t(maxn + 1) * K * 4)); MCG_TRY(q.mh.reserve(size_t(maxn + 1) * K * 4)); MCG_TRY(q.mn.reserve(size_t(maxn + 1) * 4));
        MCG_TRY(q.mbest.reserve(size_t(maxn + 1) * 4)); MCG_TRY(q.mbrank.reserve(size_t(maxn + 1) * 4));
    }
    MCG_CUDA(cudaEventRecord(c->ev_misc, c->stream));
    MCG_CUDA(cudaStreamWaitEvent(c->s_h2d, c->ev_misc, 0));     // (allocations above may have recycled blocks)
    auto stage = [&](int j) -> int {
        const int slot = j & 1;
        const int64_t a = lo[size_t(j)], b = lo[size_t(j) + 1];
        char* dbuf = c->seq.as<char>() + size_t(slot) * bufsz;
        if(j >= 2) MCG_CUDA(cudaStreamWaitEvent(c->s_h2d, c->qp[slot].evA, 0));     // stage A of part j-2 has consumed the buffer
        if(b > a) {
            MCG_CUDA(cudaMemcpyAsync(dbuf, reads1 + offs1[a], size_t(offs1[b] - offs1[a]), cudaMemcpyHostToDevice, c->s_h2d));
            if(paired) MCG_CUDA(cudaMemcpyAsync(dbuf + pad1, reads2 + offs2[a], size_t(offs2[b] - offs2[a]), cudaMemcpyHostToDevice, c->s_h2d));
        }
        MCG_CUDA(cudaEventRecord(c->ev_h2d[slot], c->s_h2d));
        return MCG_OK;
    };
    auto finish = [&](int j) -> int {
        const int slot = j & 1;
        QPart& q = c->qp[slot];
        const int64_t a = lo[size_t(j)], b = lo[size_t(j) + 1];
        MCG_TRY(qs_stage_b(db, comm, o, pl, j, slot, q.mt.as<int32_t>(), q.mh.as<uint32_t>(), q.mn.as<int32_t>(), q.mbest.as<int32_t>(),
                           q.mbrank.as<int32_t>(), c->m_stats.as<uint64_t>()));
        MCG_CUDA(cudaEventRecord(c->ev_k[slot], c->stream));
        MCG_CUDA(cudaStreamWaitEvent(c->s_d2h, c->ev_k[slot], 0));
        const size_t m = size_t(b - a);
        if(m > 0) {
            if(out_tax) MCG_CUDA(cudaMemcpyAsync(out_tax + a * K, q.mt.p, m * K * 4, cudaMemcpyDeviceToHost, c->s_d2h));
            if(out_hits) MCG_CUDA(cudaMemcpyAsync(out_hits + a * K, q.mh.p, m * K * 4, cudaMemcpyDeviceToHost, c->s_d2h));
            if(out_ncand) MCG_CUDA(cudaMemcpyAsync(out_ncand + a, q.mn.p, m * 4, cudaMemcpyDeviceToHost, c->s_d2h));
            if(out_best) MCG_CUDA(cudaMemcpyAsync(out_best + a, q.mbest.p, m * 4, cudaMemcpyDeviceToHost, c->s_d2h));
            if(out_best_rank) MCG_CUDA(cudaMemcpyAsync(out_best_rank + a, q.mbrank.p, m * 4, cudaMemcpyDeviceToHost, c->s_d2h));
        }
        MCG_CUDA(cudaEventRecord(c->ev_h2d[2 + slot], c->s_d2h));      // the merged results of this slot have left the device
        return MCG_OK;
    };
    static const bool trace = getenv("MCG_TRACE") != nullptr;
    auto now = [] { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; };
    double t_stage = 0, t_a = 0, t_fin = 0, t0 = now(), tt = 0;
    MCG_TRY(stage(0));
    for(int j = 0; j < np; ++j) {
        const int slot = j & 1;
        const int64_t a = lo[size_t(j)], b = lo[size_t(j) + 1];
        // stage A of part j is enqueued BEFORE the host waits for part j-1's left-over count (finish)
        tt = now();
        if(j + 1 < np) MCG_TRY(stage(j + 1));
        t_stage += now() - tt; tt = now();
        MCG_CUDA(cudaStreamWaitEvent(c->stream, c->ev_h2d[slot], 0));
        if(j >= 2) MCG_CUDA(cudaStreamWaitEvent(c->stream, c->ev_h2d[2 + slot], 0));   // q.m* of this slot are free again
        // the offsets keep their values of the whole slice: the base pointers are moved instead
        const char* d1 = c->seq.as<char>() + size_t(slot) * bufsz - offs1[a];
        const char* d2 = paired ? c->seq.as<char>() + size_t(slot) * bufsz + pad1 - offs2[a] : nullptr;
        MCG_TRY(qs_stage_a(db, comm, o, pl, j, slot, d1, do1 + a, offs1[b] - offs1[a], d2, paired ? do2 + a : nullptr,
                           paired ? offs2[b] - offs2[a] : 0, 0));
        t_a += now() - tt; tt = now();
        if(j >= 1) MCG_TRY(finish(j - 1));
        t_fin += now() - tt;
    }
    tt = now();
    MCG_TRY(finish(np - 1));
    t_fin += now() - tt;
    uint64_t st[22];
    MCG_CUDA(cudaMemcpyAsync(st, c->m_stats.p, 22 * 8, cudaMemcpyDeviceToHost, c->stream));
    tt = now();
    MCG_CUDA(cudaStreamSynchronize(c->s_d2h));
    MCG_CUDA(cudaStreamSynchronize(c->stream));
    if(trace) fprintf(stderr, "[mcg sharded host] %d parts: stage %.3f ms, stage A %.3f ms, finish %.3f ms, final wait %.3f ms, call %.3f ms\n",
                      np, 1e3 * t_stage, 1e3 * t_a, 1e3 * t_fin, 1e3 * (now() - tt), 1e3 * (now() - t0));
    if(out_stats22) for(int i = 0; i < 22; ++i) out_stats22[i] += st[i];
    return MCG_OK;
}

}  // extern "C"
