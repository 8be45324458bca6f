// comm.cu -- the multi-GPU data plane of libmcgpu: an NCCL communicator behind the C-ABI and the two
// exchange patterns the path needs (variable-size all-gather, variable-size all-to-all), plus the
// distributed -remove-overpopulated-features reduction.
//
// Replaces the reference's MPI traffic on this path:
//   querying.h:867-1071      blocking tree reduce of (query id, taxon id, hits) triples to rank 0
//   mode_build.cpp:847-1074  tree reduce of (key, bucket size) pairs + MPI_Bcast of the removal list
// One rank = one GPU = one mcg_ctx; ranks may be processes (torchrun, mpirun) or threads of one
// process (metacache_b200 -gpus P).  NCCL is loaded at first use (dlopen "libnccl.so.2"): a process
// that already carries an NCCL (torch) shares it, the stand-alone driver takes the system library,
// and single-GPU users of libmcgpu do not need NCCL at all.
#include "common.cuh"

#include <dlfcn.h>
#include <mutex>
#include <nccl.h>

namespace mcg {

namespace {
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
};
NcclApi g_nccl;
std::once_flag g_nccl_once;
std::string g_nccl_err;

void load_nccl()
{
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for(const char* n : names) {
        g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if(g_nccl.lib) break;
    }
    if(!g_nccl.lib) { g_nccl_err = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
#define MCG_SYM(field, name)                                                                  \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(g_nccl.lib, name));          \
    if(!g_nccl.field) { g_nccl_err = std::string("libnccl lacks ") + name; g_nccl.lib = nullptr; return; }
    MCG_SYM(GetUniqueId, "ncclGetUniqueId")
    MCG_SYM(CommInitRank, "ncclCommInitRank")
    MCG_SYM(CommDestroy, "ncclCommDestroy")
    MCG_SYM(GetErrorString, "ncclGetErrorString")
    MCG_SYM(GroupStart, "ncclGroupStart")
    MCG_SYM(GroupEnd, "ncclGroupEnd")
    MCG_SYM(Send, "ncclSend")
    MCG_SYM(Recv, "ncclRecv")
    MCG_SYM(AllGather, "ncclAllGather")
    MCG_SYM(AllReduce, "ncclAllReduce")
    MCG_SYM(GetVersion, "ncclGetVersion")
#undef MCG_SYM
}

int need_nccl()
{
    std::call_once(g_nccl_once, load_nccl);
    if(g_nccl.lib == nullptr) { set_error("%s", g_nccl_err.c_str()); return MCG_ERR_STATE; }
    return MCG_OK;
}
}  // namespace

#define MCG_NCCL(call)                                                                  \
    do {                                                                                \
        ncclResult_t r__ = (call);                                                      \
        if(r__ != ncclSuccess) {                                                        \
            mcg::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r__)); \
            return MCG_ERR_CUDA;                                                        \
        }                                                                               \
    } while(0)

}  // namespace mcg

struct mcg_comm {
    mcg_ctx* ctx = nullptr;
    ncclComm_t nccl = nullptr;
    int rank = 0, nranks = 1;
    mcg::DevBuf small;            // exchange of counts
    uint64_t* h_small = nullptr;  // pinned mirror
};

namespace mcg {

int comm_rank(const mcg_comm* m) { return m ? m->rank : 0; }
int comm_size(const mcg_comm* m) { return m ? m->nranks : 1; }

// every rank contributes `count` uint64 values; out[nranks * count] on the HOST of every rank (synchronises)
int comm_allgather_host_u64(mcg_comm* m, const uint64_t* mine, int count, uint64_t* out)
{
    if(m == nullptr || m->nranks == 1) { for(int i = 0; i < count; ++i) out[i] = mine[i]; return MCG_OK; }
    mcg_ctx* c = m->ctx;
    MCG_CUDA(cudaSetDevice(c->device));
    const size_t n = size_t(count);
    MCG_TRY(m->small.reserve((n + n * size_t(m->nranks)) * 8));
    uint64_t* d_in = m->small.as<uint64_t>();
    uint64_t* d_out = d_in + n;
    MCG_CUDA(cudaMemcpyAsync(d_in, mine, n * 8, cudaMemcpyHostToDevice, c->stream));
    MCG_NCCL(g_nccl.AllGather(d_in, d_out, n, ncclUint64, m->nccl, c->stream));
    MCG_CUDA(cudaMemcpyAsync(out, d_out, n * size_t(m->nranks) * 8, cudaMemcpyDeviceToHost, c->stream));
    MCG_CUDA(cudaStreamSynchronize(c->stream));
    return MCG_OK;
}

// the same with the contribution already on the device (d_mine[count]); enqueues + synchronises
int comm_allgather_dev_u64(mcg_comm* m, const uint64_t* d_mine, int count, uint64_t* out)
{
    mcg_ctx* c = m->ctx;
    const size_t n = size_t(count);
    if(m->nranks == 1) {
        MCG_CUDA(cudaMemcpyAsync(out, d_mine, n * 8, cudaMemcpyDeviceToHost, c->stream));
        MCG_CUDA(cudaStreamSynchronize(c->stream));
        return MCG_OK;
    }
    MCG_TRY(m->small.reserve(n * size_t(m->nranks) * 8));
    MCG_NCCL(g_nccl.AllGather(d_mine, m->small.p, n, ncclUint64, m->nccl, c->stream));
    MCG_CUDA(cudaMemcpyAsync(out, m->small.p, n * size_t(m->nranks) * 8, cudaMemcpyDeviceToHost, c->stream));
    MCG_CUDA(cudaStreamSynchronize(c->stream));
    return MCG_OK;
}

// variable-size all-gather in place: `buf` holds room for all blocks, block r = bytes [off[r], off[r] + len[r]);
// this rank's block is already there.  Grouped point-to-point copies: through NVSwitch every GPU reaches
// every peer at full link bandwidth, so P-1 sends of the own block cost what a ring all-gather costs.
int comm_allgatherv_inplace(mcg_comm* m, void* buf, const uint64_t* off, const uint64_t* len, cudaStream_t st)
{
    if(m == nullptr || m->nranks == 1) return MCG_OK;
    char* b = static_cast<char*>(buf);
    // equal blocks back to back (equal slices of equal-length reads): NCCL's own in-place all-gather, which on an
    // NVSwitch system multicasts every block once (NVLS) instead of sending it P-1 times
    static const bool no_ag = getenv("MCG_NO_NCCL_ALLGATHER") != nullptr;
    bool uniform = !no_ag && len[0] > 0;
    for(int q = 0; q < m->nranks && uniform; ++q) uniform = len[q] == len[0] && off[q] == uint64_t(q) * len[0];
    if(uniform) {
        MCG_NCCL(g_nccl.AllGather(b + off[m->rank], b, len[0], ncclUint8, m->nccl, st));
        return MCG_OK;
    }
    MCG_NCCL(g_nccl.GroupStart());
    for(int q = 0; q < m->nranks; ++q) {
        if(q == m->rank) continue;
        if(len[m->rank] > 0) MCG_NCCL(g_nccl.Send(b + off[m->rank], len[m->rank], ncclUint8, q, m->nccl, st));
        if(len[q] > 0) MCG_NCCL(g_nccl.Recv(b + off[q], len[q], ncclUint8, q, m->nccl, st));
    }
    MCG_NCCL(g_nccl.GroupEnd());
    return MCG_OK;
}

// all-to-all: block q of `send` (bytes [soff[q], soff[q] + slen[q])) goes to rank q, which stores it as block
// `rank` of its `recv` (bytes [roff[r], ...)); the own block is copied locally
int comm_alltoallv(mcg_comm* m, const void* send, const uint64_t* soff, const uint64_t* slen,
                   void* recv, const uint64_t* roff, const uint64_t* rlen, cudaStream_t st)
{
    const char* s = static_cast<const char*>(send);
    char* r = static_cast<char*>(recv);
    const int me = m ? m->rank : 0, P = m ? m->nranks : 1;
    if(slen[me] > 0 && s + soff[me] != r + roff[me])
        MCG_CUDA(cudaMemcpyAsync(r + roff[me], s + soff[me], slen[me], cudaMemcpyDeviceToDevice, st));
    if(P == 1) return MCG_OK;
    MCG_NCCL(g_nccl.GroupStart());
    for(int q = 0; q < P; ++q) {
        if(q == me) continue;
        if(slen[q] > 0) MCG_NCCL(g_nccl.Send(s + soff[q], slen[q], ncclUint8, q, m->nccl, st));
        if(rlen[q] > 0) MCG_NCCL(g_nccl.Recv(r + roff[q], rlen[q], ncclUint8, q, m->nccl, st));
    }
    MCG_NCCL(g_nccl.GroupEnd());
    return MCG_OK;
}

// several exchanges as ONE NCCL group (one kernel launch): used for the pieces of a logical record that live in
// separate arrays (features + per-read meta; taxa + hits + counts of the top lists)
int comm_group_begin(mcg_comm* m)
{
    if(m == nullptr || m->nranks == 1) return MCG_OK;
    MCG_NCCL(g_nccl.GroupStart());
    return MCG_OK;
}

int comm_group_end(mcg_comm* m)
{
    if(m == nullptr || m->nranks == 1) return MCG_OK;
    MCG_NCCL(g_nccl.GroupEnd());
    return MCG_OK;
}

int comm_allreduce_sum_u64(mcg_comm* m, uint64_t* d_buf, int count, cudaStream_t st)
{
    if(m == nullptr || m->nranks == 1) return MCG_OK;
    MCG_NCCL(g_nccl.AllReduce(d_buf, d_buf, size_t(count), ncclUint64, ncclSum, m->nccl, st));
    return MCG_OK;
}

int copy_in(mcg_ctx* c, void* d, const void* h, size_t bytes);
int copy_out(mcg_ctx* c, void* h, const void* d, size_t bytes);
int db_remove_keys_dev(mcg_db* db, const uint32_t* d_keys, uint64_t n, uint64_t* removed);
int exclusive_scan_u32_to_u32(mcg_ctx* c, const uint32_t* d_in, uint32_t* d_out, uint64_t n);

// ---- -remove-overpopulated-features across ranks (mode_build.cpp:847-1074, sketch_database.h:398-417)
// Every rank holds (key, size) of its non-empty buckets, keys ascending.  The owner of a key is rank
// key % P: pairs are binned by owner (stable, so every bin stays ascending), exchanged all-to-all, the
// owner merges the P ascending runs it received by sorting them once and sums the sizes of equal keys;
// keys whose sum exceeds maxlocs - 1 form the owner's part of the removal list, the parts are all-gathered
// and every rank clears those buckets.  Traffic per rank: 8 B per key out and in, then the removal list.
__global__ void ovp_owner_count_kernel(const uint32_t* keys, uint64_t n, int P, uint32_t* cnt /* [P] */)
{
    __shared__ uint32_t s[64];
    if(threadIdx.x < 64) s[threadIdx.x] = 0;
    __syncthreads();
    for(uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x)
        atomicAdd(&s[keys[i] % uint32_t(P)], 1u);
    __syncthreads();
    if(threadIdx.x < P && s[threadIdx.x]) atomicAdd(&cnt[threadIdx.x], s[threadIdx.x]);
}

// stable binning by owner: rank of element i inside its bin = number of earlier elements of the same owner;
// one pass per owner keeps it simple (P <= 64 small, keys stream from L2 after the first pass)
__global__ void ovp_flag_kernel(const uint32_t* keys, uint64_t n, int P, int owner, uint32_t* flag)
{
    const uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if(i < n) flag[i] = (keys[i] % uint32_t(P)) == uint32_t(owner);
}

__global__ void ovp_scatter_kernel(const uint32_t* keys, const uint32_t* sizes, uint64_t n, int P, int owner,
                                   const uint32_t* pos, uint64_t base, uint2* out)
{
    const uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if(i < n && (keys[i] % uint32_t(P)) == uint32_t(owner)) out[base + pos[i]] = make_uint2(keys[i], sizes[i]);
}

__global__ void ovp_split_kernel(const uint2* pairs, uint64_t n, uint32_t* keys, uint32_t* sizes)
{
    const uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if(i < n) { keys[i] = pairs[i].x; sizes[i] = pairs[i].y; }
}

// sorted keys + the permutation that sorted them: head of every run of equal keys sums the run (runs are <= P long)
__global__ void ovp_sum_kernel(const uint32_t* skeys, const uint32_t* order, const uint32_t* sizes, uint64_t n, uint32_t limit,
                               uint32_t* flag)
{
    const uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if(i >= n) return;
    uint32_t f = 0;
    if(i == 0 || skeys[i - 1] != skeys[i]) {
        uint64_t tot = 0;
        for(uint64_t j = i; j < n && skeys[j] == skeys[i]; ++j) tot += sizes[order[j]];
        f = tot > limit;
    }
    flag[i] = f;
}

__global__ void ovp_emit_kernel(const uint32_t* skeys, const uint32_t* flag, const uint32_t* pos, uint64_t n, uint32_t* out)
{
    const uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if(i < n && flag[i]) out[pos[i]] = skeys[i];
}

struct Scratch : DevBuf { ~Scratch() { if(p != nullptr) release(); } };

static inline unsigned grid_for(uint64_t n, int threads = 256) { return unsigned((n + threads - 1) / threads); }

}  // namespace mcg

using namespace mcg;

extern "C" {

int mcg_comm_get_id(uint8_t* id)
{
    MCG_REQUIRE(id != nullptr, "mcg_comm_get_id: NULL argument");
    MCG_TRY(need_nccl());
    static_assert(MCG_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "id size");
    ncclUniqueId u;
    MCG_NCCL(g_nccl.GetUniqueId(&u));
    std::memcpy(id, u.internal, NCCL_UNIQUE_ID_BYTES);
    return MCG_OK;
}

int mcg_comm_create(mcg_ctx* c, const uint8_t* id, int32_t rank, int32_t nranks, mcg_comm** out)
{
    MCG_REQUIRE(c != nullptr && out != nullptr, "mcg_comm_create: NULL argument");
    *out = nullptr;
    MCG_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "rank %d / %d out of range", rank, nranks);
    MCG_REQUIRE(nranks <= 64, "at most 64 ranks");
    mcg_comm* m = new mcg_comm;
    m->ctx = c;
    m->rank = rank;
    m->nranks = nranks;
    if(nranks > 1) {
        MCG_REQUIRE(id != nullptr, "mcg_comm_create: id is NULL");
        MCG_TRY(need_nccl());
        MCG_CUDA(cudaSetDevice(c->device));
        ncclUniqueId u;
        std::memcpy(u.internal, id, NCCL_UNIQUE_ID_BYTES);
        MCG_NCCL(g_nccl.CommInitRank(&m->nccl, nranks, u, rank));
    }
    *out = m;
    return MCG_OK;
}

int mcg_comm_destroy(mcg_comm* m)
{
    if(m == nullptr) return MCG_OK;
    cudaSetDevice(m->ctx->device);
    cudaStreamSynchronize(m->ctx->stream);
    if(m->nccl != nullptr) g_nccl.CommDestroy(m->nccl);
    m->small.release();
    delete m;
    return MCG_OK;
}

int mcg_comm_rank(const mcg_comm* m) { return comm_rank(m); }
int mcg_comm_size(const mcg_comm* m) { return comm_size(m); }

int mcg_db_remove_overpopulated_dist(mcg_db* db, mcg_comm* m, uint64_t* removed_here, uint64_t* removal_list_len)
{
    MCG_REQUIRE(db != nullptr && m != nullptr, "mcg_db_remove_overpopulated_dist: NULL argument");
    if(!db->finalized) { set_error("database not finalized"); return MCG_ERR_STATE; }
    MCG_REQUIRE(m->ctx == db->ctx, "communicator and database live on different contexts");
    mcg_ctx* c = db->ctx;
    MCG_CUDA(cudaSetDevice(c->device));
    const int P = m->nranks, me = m->rank;
    const uint32_t limit = uint32_t(db->maxlocs) - 1u;
    // (key, size) of the non-empty buckets, ascending keys (device-side export)
    uint64_t nk_all = 0, ne = 0, nl = 0;
    MCG_TRY(mcg_db_counts(db, &nk_all, &ne, &nl));
    Scratch d_keys, d_sizes, d_flag, d_pos, d_send, d_recv, d_cnt;
    MCG_TRY(d_keys.reserve((ne + 1) * 4)); MCG_TRY(d_sizes.reserve((ne + 1) * 4));
    MCG_TRY(mcg_db_key_sizes_dev(db, d_keys.as<uint32_t>(), d_sizes.as<uint32_t>()));
    // bin by owner
    MCG_TRY(d_cnt.reserve(64 * 4));
    MCG_CUDA(cudaMemsetAsync(d_cnt.p, 0, 64 * 4, c->stream));
    if(ne > 0) { ovp_owner_count_kernel<<<std::min<unsigned>(grid_for(ne), unsigned(c->num_sms) * 8u), 256, 0, c->stream>>>(d_keys.as<uint32_t>(), ne, P, d_cnt.as<uint32_t>()); c->launches++; }
    uint32_t h_cnt[64];
    MCG_TRY(copy_out(c, h_cnt, d_cnt.p, 64 * 4));
    MCG_CUDA(cudaStreamSynchronize(c->stream));
    const size_t PP = size_t(P);
    std::vector<uint64_t> mine(PP, 0), all(PP * PP, 0);
    for(int q = 0; q < P; ++q) mine[size_t(q)] = h_cnt[q];
    MCG_TRY(comm_allgather_host_u64(m, mine.data(), P, all.data()));          // all[r * P + q] = pairs rank r sends to rank q
    std::vector<uint64_t> soff(PP, 0), slen(PP, 0), roff(PP, 0), rlen(PP, 0);
    uint64_t stot = 0, rtot = 0;
    for(int q = 0; q < P; ++q) {
        soff[size_t(q)] = stot * 8; slen[size_t(q)] = mine[size_t(q)] * 8; stot += mine[size_t(q)];
        roff[size_t(q)] = rtot * 8; rlen[size_t(q)] = all[size_t(q) * size_t(P) + size_t(me)] * 8; rtot += all[size_t(q) * size_t(P) + size_t(me)];
    }
    MCG_TRY(d_send.reserve((stot + 1) * 8)); MCG_TRY(d_recv.reserve((rtot + 1) * 8));
    MCG_TRY(d_flag.reserve((std::max(ne, rtot) + 2) * 4)); MCG_TRY(d_pos.reserve((std::max(ne, rtot) + 2) * 4));
    for(int q = 0; q < P && ne > 0; ++q) {
        if(mine[size_t(q)] == 0) continue;
        ovp_flag_kernel<<<grid_for(ne), 256, 0, c->stream>>>(d_keys.as<uint32_t>(), ne, P, q, d_flag.as<uint32_t>());
        MCG_TRY(exclusive_scan_u32_to_u32(c, d_flag.as<uint32_t>(), d_pos.as<uint32_t>(), ne));
        ovp_scatter_kernel<<<grid_for(ne), 256, 0, c->stream>>>(d_keys.as<uint32_t>(), d_sizes.as<uint32_t>(), ne, P, q, d_pos.as<uint32_t>(),
                                                                soff[size_t(q)] / 8, d_send.as<uint2>());
        c->launches += 2;
    }
    MCG_TRY(comm_alltoallv(m, d_send.p, soff.data(), slen.data(), d_recv.p, roff.data(), rlen.data(), c->stream));
    // owner: sort the received pairs by key, sum the runs of equal keys, keep the overpopulated ones
    uint64_t nrem_mine = 0;
    Scratch d_rk, d_rs, d_ord, d_alt1, d_alt2, d_rem;
    if(rtot > 0) {
        MCG_REQUIRE(rtot < (1ull << 32), "too many keys for one owner");
        MCG_TRY(d_rk.reserve((rtot + 1) * 4)); MCG_TRY(d_rs.reserve((rtot + 1) * 4)); MCG_TRY(d_ord.reserve((rtot + 1) * 4));
        MCG_TRY(d_alt1.reserve((rtot + 1) * 4)); MCG_TRY(d_alt2.reserve((rtot + 1) * 4));
        ovp_split_kernel<<<grid_for(rtot), 256, 0, c->stream>>>(d_recv.as<uint2>(), rtot, d_rk.as<uint32_t>(), d_rs.as<uint32_t>());
        c->launches++;
        MCG_TRY(radix_sort_pairs(c, d_rk.as<uint32_t>(), d_ord.as<uint32_t>(), d_alt1.as<uint32_t>(), d_alt2.as<uint32_t>(), rtot, 32, true));
        ovp_sum_kernel<<<grid_for(rtot), 256, 0, c->stream>>>(d_rk.as<uint32_t>(), d_ord.as<uint32_t>(), d_rs.as<uint32_t>(), rtot, limit, d_flag.as<uint32_t>());
        c->launches++;
        MCG_TRY(exclusive_scan_u32_to_u32(c, d_flag.as<uint32_t>(), d_pos.as<uint32_t>(), rtot));
        uint32_t h = 0;
        MCG_TRY(copy_out(c, &h, d_pos.as<uint32_t>() + rtot, 4));
        MCG_CUDA(cudaStreamSynchronize(c->stream));
        nrem_mine = h;
    }
    // all-gather of the removal list parts
    std::vector<uint64_t> one(1, nrem_mine), cnts(PP, 0);
    MCG_TRY(comm_allgather_host_u64(m, one.data(), 1, cnts.data()));
    std::vector<uint64_t> goff(PP, 0), glen(PP, 0);
    uint64_t gtot = 0;
    for(int q = 0; q < P; ++q) { goff[size_t(q)] = gtot * 4; glen[size_t(q)] = cnts[size_t(q)] * 4; gtot += cnts[size_t(q)]; }
    MCG_TRY(d_rem.reserve((gtot + 1) * 4));
    if(nrem_mine > 0) {
        ovp_emit_kernel<<<grid_for(rtot), 256, 0, c->stream>>>(d_rk.as<uint32_t>(), d_flag.as<uint32_t>(), d_pos.as<uint32_t>(), rtot,
                                                             d_rem.as<uint32_t>() + goff[size_t(me)] / 4);
        c->launches++;
    }
    MCG_TRY(comm_allgatherv_inplace(m, d_rem.p, goff.data(), glen.data(), c->stream));
    MCG_CUDA(cudaGetLastError());
    uint64_t rem = 0;
    if(gtot > 0) MCG_TRY(db_remove_keys_dev(db, d_rem.as<uint32_t>(), gtot, &rem));
    MCG_CUDA(cudaStreamSynchronize(c->stream));
    if(removed_here) *removed_here = rem;
    if(removal_list_len) *removal_list_len = gtot;
    return MCG_OK;
}

}  // extern "C"
