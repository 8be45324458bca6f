/* oracle/mpi_shim/mpi.h -- TEST INFRASTRUCTURE, not product code.
 *
 * Minimal single-node stand-in for the handful of MPI entry points that the
 * reference (jmabuin/metacache-mpi, /root/reference/src) calls, so that the
 * UNMODIFIED reference sources can be compiled into oracle/_ref/ and run as
 * P cooperating processes without an MPI installation (none exists in this
 * image).  Used call sites: main.cpp:48-53,103; mode_build.cpp:842-1099;
 * querying.h:744-1073,1360; mode_query.cpp:129.
 *
 * Transport: the launcher (oracle/mpirun_shim.py) creates one named FIFO per
 * ordered pair of ranks in $MPISHIM_DIR and starts P copies of the binary with
 * MPISHIM_RANK / MPISHIM_SIZE set.  Every FIFO is opened O_RDWR so that open()
 * never blocks.  A message is {int64 tag, int64 nbytes} followed by the
 * payload.  The reference only uses blocking, pairwise, tag-ordered traffic,
 * so per-pair FIFO order is sufficient.  With MPISHIM_SIZE unset the shim
 * degenerates to a single rank.
 */
#ifndef MPISHIM_MPI_H
#define MPISHIM_MPI_H

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <fcntl.h>
#include <unistd.h>

typedef int MPI_Comm;
typedef int MPI_Datatype;           /* value = size of one element in bytes */
struct MPI_Status { int MPI_SOURCE; int MPI_TAG; int MPI_ERROR; };

#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_UINT32_T 4
#define MPI_INT 4
#define MPI_UINT64_T 8
#define MPI_BYTE 1
#define MPI_CHAR 1

namespace mpishim {

struct state {
    int rank = 0;
    int size = 1;
    std::vector<int> out;   /* out[p]: fifo this rank -> p */
    std::vector<int> in;    /* in[p] : fifo p -> this rank */
    bool up = false;
};

inline state& st() { static state s; return s; }

inline void die(const char* what) {
    std::fprintf(stderr, "[mpishim rank %d] fatal: %s\n", st().rank, what);
    std::_Exit(97);
}

inline void write_all(int fd, const void* buf, size_t n) {
    const char* p = static_cast<const char*>(buf);
    while(n > 0) {
        ssize_t w = ::write(fd, p, n);
        if(w <= 0) die("write failed");
        p += w; n -= size_t(w);
    }
}

inline void read_all(int fd, void* buf, size_t n) {
    char* p = static_cast<char*>(buf);
    while(n > 0) {
        ssize_t r = ::read(fd, p, n);
        if(r <= 0) die("read failed");
        p += r; n -= size_t(r);
    }
}

inline void send_raw(int dst, std::int64_t tag, const void* buf, std::int64_t nbytes) {
    std::int64_t hdr[2] = {tag, nbytes};
    write_all(st().out[dst], hdr, sizeof(hdr));
    if(nbytes > 0) write_all(st().out[dst], buf, size_t(nbytes));
}

inline void recv_raw(int src, std::int64_t tag, void* buf, std::int64_t maxbytes) {
    std::int64_t hdr[2];
    read_all(st().in[src], hdr, sizeof(hdr));
    if(hdr[0] != tag) die("tag mismatch (messages of a pair must arrive in send order)");
    if(hdr[1] > maxbytes) die("message longer than receive buffer");
    if(hdr[1] > 0) read_all(st().in[src], buf, size_t(hdr[1]));
}

constexpr std::int64_t tag_barrier = -1001;
constexpr std::int64_t tag_bcast   = -1002;
constexpr std::int64_t tag_gather  = -1003;

} // namespace mpishim

inline int MPI_Init(int*, char***) {
    auto& s = mpishim::st();
    const char* r = std::getenv("MPISHIM_RANK");
    const char* n = std::getenv("MPISHIM_SIZE");
    const char* d = std::getenv("MPISHIM_DIR");
    if(!r || !n || !d) { s.rank = 0; s.size = 1; s.up = true; return MPI_SUCCESS; }
    s.rank = std::atoi(r);
    s.size = std::atoi(n);
    s.out.assign(s.size, -1);
    s.in.assign(s.size, -1);
    for(int p = 0; p < s.size; ++p) {
        if(p == s.rank) continue;
        std::string fo = std::string(d) + "/f_" + std::to_string(s.rank) + "_" + std::to_string(p);
        std::string fi = std::string(d) + "/f_" + std::to_string(p) + "_" + std::to_string(s.rank);
        s.out[p] = ::open(fo.c_str(), O_RDWR);
        s.in[p]  = ::open(fi.c_str(), O_RDWR);
        if(s.out[p] < 0 || s.in[p] < 0) mpishim::die("cannot open fifo");
    }
    s.up = true;
    return MPI_SUCCESS;
}

inline int MPI_Finalize() {
    auto& s = mpishim::st();
    for(int fd : s.out) if(fd >= 0) ::close(fd);
    for(int fd : s.in)  if(fd >= 0) ::close(fd);
    s.out.clear(); s.in.clear();
    return MPI_SUCCESS;
}

inline int MPI_Comm_size(MPI_Comm, int* n) { *n = mpishim::st().size; return MPI_SUCCESS; }
inline int MPI_Comm_rank(MPI_Comm, int* r) { *r = mpishim::st().rank; return MPI_SUCCESS; }

inline int MPI_Send(const void* buf, int count, MPI_Datatype dt, int dst, int tag, MPI_Comm) {
    mpishim::send_raw(dst, tag, buf, std::int64_t(count) * dt);
    return MPI_SUCCESS;
}

inline int MPI_Recv(void* buf, int count, MPI_Datatype dt, int src, int tag, MPI_Comm, MPI_Status* stt) {
    mpishim::recv_raw(src, tag, buf, std::int64_t(count) * dt);
    if(stt) { stt->MPI_SOURCE = src; stt->MPI_TAG = tag; stt->MPI_ERROR = 0; }
    return MPI_SUCCESS;
}

inline int MPI_Barrier(MPI_Comm) {
    auto& s = mpishim::st();
    if(s.size < 2) return MPI_SUCCESS;
    char c = 0;
    if(s.rank == 0) {
        for(int p = 1; p < s.size; ++p) mpishim::recv_raw(p, mpishim::tag_barrier, &c, 1);
        for(int p = 1; p < s.size; ++p) mpishim::send_raw(p, mpishim::tag_barrier, &c, 1);
    } else {
        mpishim::send_raw(0, mpishim::tag_barrier, &c, 1);
        mpishim::recv_raw(0, mpishim::tag_barrier, &c, 1);
    }
    return MPI_SUCCESS;
}

inline int MPI_Bcast(void* buf, int count, MPI_Datatype dt, int root, MPI_Comm) {
    auto& s = mpishim::st();
    if(s.size < 2) return MPI_SUCCESS;
    const std::int64_t nb = std::int64_t(count) * dt;
    if(s.rank == root) {
        for(int p = 0; p < s.size; ++p) if(p != root) mpishim::send_raw(p, mpishim::tag_bcast, buf, nb);
    } else {
        mpishim::recv_raw(root, mpishim::tag_bcast, buf, nb);
    }
    return MPI_SUCCESS;
}

inline int MPI_Gather(const void* sbuf, int scount, MPI_Datatype sdt,
                      void* rbuf, int rcount, MPI_Datatype rdt, int root, MPI_Comm) {
    auto& s = mpishim::st();
    const std::int64_t nb = std::int64_t(scount) * sdt;
    if(s.rank == root) {
        char* out = static_cast<char*>(rbuf);
        const std::int64_t rb = std::int64_t(rcount) * rdt;
        for(int p = 0; p < s.size; ++p) {
            if(p == root) std::memcpy(out + p*rb, sbuf, size_t(nb));
            else mpishim::recv_raw(p, mpishim::tag_gather, out + p*rb, rb);
        }
    } else {
        mpishim::send_raw(root, mpishim::tag_gather, sbuf, nb);
    }
    return MPI_SUCCESS;
}

#endif
