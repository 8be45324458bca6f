// common.cuh -- shared host/device helpers of libmcgpu (sm_100a only)
#pragma once

#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mcgpu.h"

namespace mcg {

// ------------------------------------------------------------------ errors
void set_error(const char* fmt, ...);

#define MCG_CUDA(call)                                                              \
    do {                                                                            \
        cudaError_t e__ = (call);                                                   \
        if(e__ != cudaSuccess) {                                                    \
            mcg::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,            \
                           cudaGetErrorString(e__));                                \
            return MCG_ERR_CUDA;                                                    \
        }                                                                           \
    } while(0)

#define MCG_TRY(call)                                                               \
    do {                                                                            \
        int s__ = (call);                                                           \
        if(s__ != MCG_OK) return s__;                                               \
    } while(0)

#define MCG_REQUIRE(cond, ...)                                                      \
    do {                                                                            \
        if(!(cond)) {                                                               \
            mcg::set_error(__VA_ARGS__);                                            \
            return MCG_ERR_ARG;                                                     \
        }                                                                           \
    } while(0)


// ------------------------------------------------------------------ device memory
// grow-only device buffer (cudaMalloc is synchronising; buffers are reused across calls)
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);        // contents are NOT preserved
    int reserve_keep(size_t bytes, size_t used, cudaStream_t st);   // contents preserved
    void release();
    template<class T> T* as() const { return static_cast<T*>(p); }
};

void release_cached_device_memory();
void trim_cached_device_memory(size_t keep_bytes);

// parameters of the feature table's bucket function (see "feature table" below)
struct TableGeom {
    const uint32_t* cdf;      // [2^TABLE_CDF_BITS + 1] keys below each cell boundary
    uint32_t scale;           // floor(home buckets * 2^32 / nkeys)
};

}  // namespace mcg

// ------------------------------------------------------------------ context / db
struct mcg_comm;

namespace mcg {
// Buffers of one query part in flight.  Stage A of a part (sketch, feature exchange, fused kernel) is
// enqueued before the host looks at the previous part's left-over count, so what stage A writes and the
// later stages (left-over reads, list exchange, merge, copy-out) read exists twice.
struct QPart {
    DevBuf W, feat, meta, left_list, left_list2, done, ctr;   // window prefix, features, per-read meta; fused / spill kernel bookkeeping
    DevBuf lt;                                         // per-read records {tax[K], hits[K], n} of every read of the part
    DevBuf ct;                                         // records received from the other shards for the own block [P][n]
    DevBuf mt, mh, mn, mbest, mbrank;                  // merged lists + decision of the own block
    cudaEvent_t evA = nullptr, evB = nullptr;
    void release_all()
    {
        DevBuf* b[] = {&W, &feat, &meta, &left_list, &left_list2, &done, &ctr, &lt, &ct, &mt, &mh, &mn, &mbest, &mbrank};
        for(DevBuf* x : b) x->release();
        if(evA) cudaEventDestroy(evA);
        if(evB) cudaEventDestroy(evB);
        evA = evB = nullptr;
    }
};
}  // namespace mcg

struct mcg_ctx {
    int device = 0;
    int num_sms = 148;                // multiProcessorCount of the device (B200: 2 dies x 74 SMs): persistent grids are sized from it
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int64_t launches = 0;
    // measurement hook (mcg_ctx_profile): event pairs around the launches of the dominant kernels
    bool profile = false;
    std::vector<cudaEvent_t> prof_ev[8];      // start, stop, start, stop, ... per kernel / phase class
    void prof_begin(int which, cudaStream_t st)
    {
        if(!profile) return;
        cudaEvent_t a = nullptr, b = nullptr;
        cudaEventCreate(&a); cudaEventCreate(&b);
        prof_ev[which].push_back(a); prof_ev[which].push_back(b);
        cudaEventRecord(a, st);
    }
    void prof_end(int which, cudaStream_t st)
    {
        if(!profile || prof_ev[which].empty()) return;
        cudaEventRecord(prof_ev[which].back(), st);
    }
    // taxonomy tables (replicated per GPU)
    mcg::DevBuf lineage, rank, lift;  // lift: per target, the taxon its candidates are listed under at `lift_lowest`
    int64_t ntaxa = 0, ntargets = 0;
    int lift_lowest = -1;
    // scratch
    mcg::DevBuf seq, seqoff, winfirst, outfirst, tmp0, tmp1, tmp2, tmp3, tmp4, tmp5, scan_tmp, tilectr;
    mcg::DevBuf q_feat, q_flen, q_foff, q_fpos, q_loc, q_rinfo, q_cls, q_out_tax, q_out_hits, q_out_n, q_done;
    mcg::DevBuf m_tax, m_hits, m_n, m_o_tax, m_o_hits, m_o_n, m_best, m_brank, m_stats, m_bad;
    bool m_bad_clear = false;         // m_bad (flags of the merge kernel's input checks) is zero
    void* pinned = nullptr;           // small pinned staging area for scalars
    size_t pinned_cap = 0;
    mcg::QPart qp[2];                 // query parts in flight (see QPart)
    uint32_t* h_left = nullptr;       // pinned: left-over counts of the two parts
    mcg::DevBuf s_wc, s_featc, plan;  // left-over path: compacted window prefix / features; sharded query: part sizes
    mcg::DevBuf q_spill;              // match lists of the spill kernel (global memory, per warp)
    // copy streams + events of the pipelined host entry points
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_h2d[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_k[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_misc = nullptr;
};

struct mcg_db {
    mcg_ctx* ctx = nullptr;
    mcg_sketch_cfg cfg{};
    int maxlocs = 254;
    int shard = 0, nshards = 1;
    bool finalized = false;
    // insert phase: feature slots [nslots] in window-major order + per-window (tgt,win)
    mcg::DevBuf feat;                 // uint32[nwin_total * sketchlen]
    uint64_t nwin = 0;
    std::vector<uint32_t> h_tgt_id;   // owned targets in insertion order
    std::vector<uint64_t> h_tgt_firstwin;   // first global window of each owned target (+ sentinel)
    // finalized: CSR + open-addressing table
    mcg::DevBuf keys, offs, sizes;    // unique keys ascending, uint64 offsets, uint8 sizes
    mcg::DevBuf locs;                 // uint64 (tgt << 32 | win)
    mcg::DevBuf table;                // 128-byte buckets of 3 slots (layout: table_bucket below); table_cap = buckets
    mcg::DevBuf cdf;                  // cumulative key histogram of the bucket function
    uint32_t table_scale = 0;
    mcg::TableGeom geom() const { mcg::TableGeom g; g.cdf = cdf.as<uint32_t>(); g.scale = table_scale; return g; }
    uint64_t nkeys = 0, nlocs = 0, table_cap = 0;
    uint64_t max_tgt = 0;             // largest target id stored in the shard (checked against the taxonomy at query time)
    // per-read match keys: (tgt << wbits) | win; 32-bit when target and window ids fit together
    int wbits = 32;
    bool key32 = false;
    int kbits = 64;                   // significant bits of a match key (target bits + window bits)
    void set_key_layout(uint64_t max_tgt_, uint64_t max_win)
    {
        const uint64_t max_tgt = max_tgt_;
        this->max_tgt = max_tgt_;
        int wb = 1, tb = 1;
        while((max_win >> wb) != 0) ++wb;
        while((max_tgt >> tb) != 0) ++tb;
        if(wb + tb <= 32) { wbits = wb; key32 = true; kbits = wb + tb; } else { wbits = 32; key32 = false; kbits = 64; }
    }
};

namespace mcg {

// ------------------------------------------------------------------ device helpers
__host__ __device__ __forceinline__ uint32_t tmh(uint32_t x)
{
    x = ((x >> 16) ^ x) * 0x45d9f3bu;
    x = ((x >> 16) ^ x) * 0x45d9f3bu;
    return (x >> 16) ^ x;
}

// slot hash of the feature -> list table (features are bottom-s hash values and
// therefore crowd the low end of the range: re-mix before reducing)
__host__ __device__ __forceinline__ uint32_t slot_mix(uint32_t x)
{
    x *= 0x9E3779B1u;
    x ^= x >> 15;
    x *= 0x85EBCA77u;
    x ^= x >> 13;
    return x;
}

// ------------------------------------------------------------------ feature table (build.cu / query.cu)
// Buckets of one 128-byte line = 3 slots, one 32-byte DRAM sector each for the keys and for
// every slot's payload (uint32 words of the line):
//   [0..2] keys   [3..7] unused
//   [8 + 8t ..] payload of slot t: size (u32), off (u32 index into locs: < 2^32 locations per shard), then 24 bytes of
//               inline matches: when target and window ids of the shard fit 32 bits together (db->key32) the first SIX
//               locations as packed match keys (tgt << wbits) | win -- the form the per-read sort works on --
//               otherwise the first THREE locations as stored (tgt << 32 | win)
// The home bucket of a key is MONOTONE in the key: bucket = rank_estimate(key) * nbuckets / nkeys,
// where rank_estimate interpolates a 4096-cell cumulative histogram of the shard's keys (16 KB,
// stays in L1).  Keys are bottom-s hashes, i.e. far from uniform, but locally uniform inside a
// cell, so bucket loads are as even as with a random hash -- and because the unique keys leave
// the radix sort ascending, the whole table is written front to back by a max-scan
// (slot_j = max(3 * home_j, slot_{j-1} + 1)): no atomics, no random writes.  A key lives in the
// first bucket at or after its home bucket that had a free slot when it was placed; lookups walk
// forward until the key or a free slot shows up (the table ends with an empty bucket: no wrap).
// A probe is ONE line -- what an L1 miss fetches anyway -- with the first six (three) locations of the list.
constexpr uint32_t TABLE_EMPTY = 0xFFFFFFFFu;
constexpr int TABLE_BUCKET_WORDS = 32;
constexpr int TABLE_BUCKET_SLOTS = 3;
constexpr int TABLE_CDF_BITS = 12;
constexpr uint32_t TABLE_INLINE_32 = 6, TABLE_INLINE_64 = 3;      // inline matches per slot (packed 32-bit keys / 64-bit locations)

__device__ __forceinline__ uint32_t table_bucket(uint32_t key, const TableGeom& g)
{
    constexpr int SH = 32 - TABLE_CDF_BITS;
    const uint32_t c = key >> SH;
    const uint32_t lo = __ldg(g.cdf + c), hi = __ldg(g.cdf + c + 1);
    const uint32_t rank = lo + uint32_t((uint64_t(hi - lo) * (key & ((1u << SH) - 1u))) >> SH);
    return uint32_t((uint64_t(rank) * g.scale) >> 32);
}

// slot of `key` in a bucket whose key sector starts with k4: 0..2, -1 = the key is not in the table, -2 = look
// in the next bucket.  Slots are handed out in key order, so the table as a whole is sorted: a walk that starts
// at the key's home bucket goes on only past a FULL bucket whose largest key (its last slot) is smaller than the
// key; a free slot, or a larger key in the bucket, proves absence -- most probes of absent keys that land in a full
// bucket end there instead of costing another dependent line fill.
__device__ __forceinline__ int table_match(const uint4& k4, uint32_t key)
{
    if(k4.x == key) return 0;
    if(k4.y == key) return 1;
    if(k4.z == key) return 2;
    return (k4.z != TABLE_EMPTY && k4.z < key) ? -2 : -1;
}

__host__ __device__ __forceinline__ int64_t num_windows(int64_t len, int64_t winlen, int64_t stride)
{
    if(len <= winlen) return 1;
    const int64_t nfull = (len - winlen) / stride + 1;
    return nfull + ((nfull * stride < len) ? 1 : 0);
}

inline int64_t div_up(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ------------------------------------------------------------------ primitives (scan.cu, radix_sort.cu)
// exclusive prefix sum of n uint32 -> uint64 out[n+1] (out[n] = total)
int exclusive_scan_u32_to_u64(mcg_ctx* c, const uint32_t* d_in, uint64_t* d_out, uint64_t n);
// exclusive prefix sum of n uint8 -> uint64 out[n+1]
int exclusive_scan_u8_to_u64(mcg_ctx* c, const uint8_t* d_in, uint64_t* d_out, uint64_t n);
// stable LSD radix sort of (key,val) uint32 pairs on bits [0, key_bits); result in keys/vals
// (alt buffers are scratch of the same size); iota_vals: the values are not read but generated, value of
// element i = i / iota_div
int radix_sort_pairs(mcg_ctx* c, uint32_t* keys, uint32_t* vals, uint32_t* keys_alt, uint32_t* vals_alt,
                     uint64_t n, int key_bits, bool iota_vals, uint32_t iota_div = 1);

// ------------------------------------------------------------------ stages
int sketch_launch(mcg_ctx* c, const mcg_sketch_cfg& cfg, int sketchlen, int64_t winlen, int64_t stride,
                  const char* d_seq, int64_t total_bytes, const int64_t* d_seq_off, int64_t nseq,
                  const uint32_t* d_win_first, const uint32_t* d_out_first, uint64_t nwin,
                  uint32_t* d_out_feat, uint32_t* d_out_cnt, uint64_t win_lo = 0, cudaStream_t stream = nullptr,
                  uint32_t* d_tile_scratch = nullptr, const uint32_t* d_nwin = nullptr);   // d_nwin: exact window count on the device when `nwin` is only a bound
int sketch_config(const mcg_sketch_cfg& cfg, int sketchlen, int64_t winlen, int64_t stride,
                  int* wpt, int* ngroups, int* pitch, size_t* warp_smem);
int window_prefix(mcg_ctx* c, int64_t winlen, int64_t stride, const int64_t* d_seq_off, int64_t nseq,
                  uint32_t* d_win_first /* nseq+1 */, uint64_t* h_total);

}  // namespace mcg
