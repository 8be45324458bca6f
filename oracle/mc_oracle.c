/* oracle/mc_oracle.c -- TEST INFRASTRUCTURE, not product code (see mc_oracle.h).
 *
 * Plain-C restatement of the MetaCache-MPI build/query hot path.  Every
 * function cites the reference file:line whose behaviour it restates
 * (paths relative to /root/reference/src).  Written from the behavioural
 * specification in SURVEY.md Appendix A; shares no code with the reference.
 */
#define _POSIX_C_SOURCE 200809L
#include "mc_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ======================================================================== */
/* L1                                                                        */
/* ======================================================================== */

/* hash_int.h:40-45 thomas_mueller_hash */
uint32_t mco_tmh(uint32_t x)
{
    x = ((x >> 16) ^ x) * 0x45d9f3bu;
    x = ((x >> 16) ^ x) * 0x45d9f3bu;
    return (x >> 16) ^ x;
}

/* dna_encoding.h:114-121 make_reverse_complement(uint32_t,k): reverse the
 * sixteen 2-bit letters, complement (3 - letter), keep the k letters that
 * were the original low 2k bits */
uint32_t mco_revcomp(uint32_t x, int k)
{
    uint32_t r = 0;
    for(int i = 0; i < 16; ++i) {
        r = (r << 2) | (x & 3u);
        x >>= 2;
    }
    r = ~r;
    return r >> (32 - 2 * k);
}

/* dna_encoding.h:189-197 make_canonical */
uint32_t mco_canonical(uint32_t x, int k)
{
    const uint32_t rc = mco_revcomp(x, k);
    return x < rc ? x : rc;
}

/* dna_encoding.h:330-336: letter code, -1 = ambiguous */
static inline int base_code(char c)
{
    switch(c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return -1;
    }
}

/* dna_encoding.h:305-348 for_each_kmer_2bit<uint32_t>: every position i >= k-1
 * yields (kmer of letters i-k+1..i, bitfield of ambiguous letters among them) */
int64_t mco_kmers(const char* s, int64_t n, int k, uint32_t* kmers, uint16_t* ambig)
{
    if(n < k) return 0;
    const uint32_t kmask = 0xFFFFFFFFu >> (32 - 2 * k);
    const uint16_t amask = (uint16_t)(0xFFFFu >> (16 - k));
    uint32_t km = 0;
    uint16_t am = 0;
    int64_t c = 0;
    for(int64_t i = 0; i < n; ++i) {
        const int b = base_code(s[i]);
        km <<= 2;
        am = (uint16_t)(am << 1);
        if(b < 0) am |= 1; else km |= (uint32_t)b;
        if(i >= k - 1) {
            km &= kmask;
            am &= amask;
            kmers[c] = km;
            ambig[c] = am;
            ++c;
        }
    }
    return c;
}

/* dna_encoding.h:261-276 for_each_window */
int64_t mco_num_windows(int64_t len, int64_t winlen, int64_t stride)
{
    if(len <= winlen) return 1;
    /* full windows: first = 0, stride, ... while first + winlen <= len */
    const int64_t nfull = (len - winlen) / stride + 1;
    const int64_t next = nfull * stride;
    return nfull + (next < len ? 1 : 0);
}

void mco_window(int64_t len, int64_t winlen, int64_t stride, int64_t w,
                int64_t* beg, int64_t* end)
{
    if(len <= winlen) { *beg = 0; *end = len; return; }
    const int64_t b = w * stride;
    *beg = b;
    *end = (b + winlen <= len) ? b + winlen : len;
}

/* hash_dna.h:115-152 single_function_unique_min_hasher::operator():
 * the s' = min(sketchlen, n-k+1) smallest DISTINCT hashes of the canonical
 * k-mers without ambiguous letters, ascending; 0xFFFFFFFF is never a feature */
int mco_sketch_window(const char* s, int64_t n, int k, int sketchlen, uint32_t* out)
{
    if(n < k) return 0;
    int64_t cap = n - k + 1;
    if(cap > sketchlen) cap = sketchlen;
    if(cap < 1) return 0;
    const int sk = (int)cap;
    for(int i = 0; i < sk; ++i) out[i] = MCO_NO_FEATURE;

    const uint32_t kmask = 0xFFFFFFFFu >> (32 - 2 * k);
    uint32_t km = 0;
    int64_t lastbad = -1;               /* index of the latest ambiguous letter */
    for(int64_t i = 0; i < n; ++i) {
        const int b = base_code(s[i]);
        km = (km << 2) & kmask;
        if(b < 0) lastbad = i; else km |= (uint32_t)b;
        if(i < k - 1) continue;
        if(lastbad > i - k) continue;   /* ambiguity inside the k-mer: dna_encoding.h:462-465 */
        const uint32_t h = mco_tmh(mco_canonical(km, k));
        if(h >= out[sk - 1]) continue;
        /* position of first element >= h; skip if equal (unique) */
        int p = 0;
        while(out[p] < h) ++p;
        if(out[p] == h) continue;
        for(int j = sk - 1; j > p; --j) out[j] = out[j - 1];
        out[p] = h;
    }
    int cnt = sk;
    while(cnt > 0 && out[cnt - 1] == MCO_NO_FEATURE) --cnt;   /* hash_dna.h:142-149 */
    return cnt;
}

int64_t mco_sketch_sequences(const char* buf, const int64_t* offs, int64_t nseq,
                             int k, int sketchlen, int64_t winlen, int64_t stride,
                             uint32_t* out_feat, uint32_t* out_cnt)
{
    int64_t g = 0;
    uint32_t tmp[256];
    for(int64_t q = 0; q < nseq; ++q) {
        const char* s = buf + offs[q];
        const int64_t len = offs[q + 1] - offs[q];
        const int64_t nw = mco_num_windows(len, winlen, stride);
        for(int64_t w = 0; w < nw; ++w, ++g) {
            int64_t b, e;
            mco_window(len, winlen, stride, w, &b, &e);
            const int c = mco_sketch_window(s + b, e - b, k, sketchlen, tmp);
            uint32_t* o = out_feat + g * sketchlen;
            for(int i = 0; i < sketchlen; ++i) o[i] = i < c ? tmp[i] : MCO_NO_FEATURE;
            out_cnt[g] = (uint32_t)c;
        }
    }
    return g;
}

/* ---- timing leg (bench.py cpu_baseline, kind "port") -------------------- */
typedef struct {
    const char* buf; const int64_t* offs; int64_t nseq;
    int k, sketchlen; int64_t winlen, stride;
    int64_t* next; pthread_mutex_t* mtx;
    uint64_t sum, wins;
} tp_arg;

static void* tp_worker(void* p)
{
    tp_arg* a = (tp_arg*)p;
    uint32_t tmp[256];
    for(;;) {
        pthread_mutex_lock(a->mtx);
        const int64_t q = (*a->next)++;
        pthread_mutex_unlock(a->mtx);
        if(q >= a->nseq) break;
        const char* s = a->buf + a->offs[q];
        const int64_t len = a->offs[q + 1] - a->offs[q];
        const int64_t nw = mco_num_windows(len, a->winlen, a->stride);
        for(int64_t w = 0; w < nw; ++w) {
            int64_t b, e;
            mco_window(len, a->winlen, a->stride, w, &b, &e);
            const int c = mco_sketch_window(s + b, e - b, a->k, a->sketchlen, tmp);
            for(int i = 0; i < c; ++i) a->sum += tmp[i];
            a->wins++;
        }
    }
    return NULL;
}

double mco_sketch_throughput(const char* buf, const int64_t* offs, int64_t nseq,
                             int k, int sketchlen, int64_t winlen, int64_t stride,
                             int threads, uint64_t* checksum, uint64_t* nwindows)
{
    if(threads < 1) threads = 1;
    if(threads > 256) threads = 256;
    pthread_t th[256];
    tp_arg args[256];
    int64_t next = 0;
    pthread_mutex_t mtx = PTHREAD_MUTEX_INITIALIZER;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for(int t = 0; t < threads; ++t) {
        args[t] = (tp_arg){buf, offs, nseq, k, sketchlen, winlen, stride, &next, &mtx, 0, 0};
        pthread_create(&th[t], NULL, tp_worker, &args[t]);
    }
    uint64_t sum = 0, wins = 0;
    for(int t = 0; t < threads; ++t) {
        pthread_join(th[t], NULL);
        sum += args[t].sum;
        wins += args[t].wins;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if(checksum) *checksum = sum;
    if(nwindows) *nwindows = wins;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* ======================================================================== */
/* L2: database shard                                                        */
/* ======================================================================== */

typedef struct { uint32_t key, tgt, win; } tuple3;

struct mco_db {
    int k, sketchlen, maxlocs, shard, nshards;
    int64_t winlen, stride;
    /* insertion log, in (tgt,win) order */
    tuple3* tup; uint64_t ntup, captup;
    /* finalized: keys ascending (including emptied ones), CSR lists */
    uint32_t* keys; uint64_t* off; uint8_t* size; uint64_t nkeys;
    uint32_t* ltgt; uint32_t* lwin; uint64_t nloc;
    int finalized;
};

/* sketch_database.h:356-368 max_locations_per_feature (bucket_size_type is
 * uint8_t: the caller's value is narrowed first, mode_build.cpp:659; quirk Q7) */
mco_db* mco_db_create(int k, int sketchlen, int64_t winlen, int64_t stride,
                      int maxlocs, int shard, int nshards)
{
    mco_db* d = (mco_db*)calloc(1, sizeof(mco_db));
    d->k = k; d->sketchlen = sketchlen; d->winlen = winlen; d->stride = stride;
    int m = (int)(uint8_t)maxlocs;
    if(m < 1) m = 1;
    if(m >= 254) m = 254;
    d->maxlocs = m;
    d->shard = shard; d->nshards = nshards;
    return d;
}

void mco_db_destroy(mco_db* d)
{
    if(!d) return;
    free(d->tup); free(d->keys); free(d->off); free(d->size); free(d->ltgt); free(d->lwin);
    free(d);
}

int mco_db_maxlocs(const mco_db* d) { return d->maxlocs; }

/* sketch_database.h:519-563 add_target_distributed (partition rule :540) and
 * :1079-1097 add_all_window_sketches */
int64_t mco_db_add_target(mco_db* d, const char* seq, int64_t len, uint32_t target_id)
{
    if((int)(target_id % (uint32_t)d->nshards) != d->shard) return 0;
    uint32_t tmp[256];
    const int64_t nw = mco_num_windows(len, d->winlen, d->stride);
    for(int64_t w = 0; w < nw; ++w) {
        int64_t b, e;
        mco_window(len, d->winlen, d->stride, w, &b, &e);
        const int c = mco_sketch_window(seq + b, e - b, d->k, d->sketchlen, tmp);
        if(d->ntup + (uint64_t)c > d->captup) {
            d->captup = d->captup ? d->captup * 2 : (1u << 16);
            while(d->captup < d->ntup + (uint64_t)c) d->captup *= 2;
            d->tup = (tuple3*)realloc(d->tup, d->captup * sizeof(tuple3));
        }
        for(int i = 0; i < c; ++i) d->tup[d->ntup++] = (tuple3){tmp[i], target_id, (uint32_t)w};
    }
    d->finalized = 0;
    return nw;
}

/* bench.py's parity sample at full database size: the lists of a GIVEN set of features only.  A database
 * restricted to the keys in allowed[] (ascending) answers lookups of those keys exactly like the whole
 * database (every list is built from the same insertions in the same order), and sketching -- the part that
 * cannot be skipped -- runs on `threads` pthreads over the targets. */
typedef struct {
    const mco_db* d; const char* buf; const int64_t* offs; const uint32_t* ids; int64_t n;
    const uint32_t* allowed; int64_t nallowed;
    int64_t* next; pthread_mutex_t* mtx;
    tuple3* tup; uint64_t ntup, captup; int64_t wins;
} rs_arg;

static int key_allowed(const uint32_t* a, int64_t n, uint32_t key)
{
    int64_t lo = 0, hi = n;
    while(lo < hi) { const int64_t mid = (lo + hi) / 2; if(a[mid] < key) lo = mid + 1; else hi = mid; }
    return lo < n && a[lo] == key;
}

static void* rs_worker(void* p)
{
    rs_arg* a = (rs_arg*)p;
    uint32_t tmp[256];
    for(;;) {
        pthread_mutex_lock(a->mtx);
        const int64_t q = (*a->next)++;
        pthread_mutex_unlock(a->mtx);
        if(q >= a->n) break;
        if((int)(a->ids[q] % (uint32_t)a->d->nshards) != a->d->shard) continue;        /* sketch_database.h:540 */
        const char* s = a->buf + a->offs[q];
        const int64_t len = a->offs[q + 1] - a->offs[q];
        const int64_t nw = mco_num_windows(len, a->d->winlen, a->d->stride);
        for(int64_t w = 0; w < nw; ++w) {
            int64_t b, e;
            mco_window(len, a->d->winlen, a->d->stride, w, &b, &e);
            const int c = mco_sketch_window(s + b, e - b, a->d->k, a->d->sketchlen, tmp);
            for(int i = 0; i < c; ++i) {
                if(!key_allowed(a->allowed, a->nallowed, tmp[i])) continue;
                if(a->ntup + 1 > a->captup) { a->captup = a->captup ? a->captup * 2 : 4096; a->tup = (tuple3*)realloc(a->tup, a->captup * sizeof(tuple3)); }
                a->tup[a->ntup++] = (tuple3){tmp[i], a->ids[q], (uint32_t)w};
            }
        }
        a->wins += nw;
    }
    return NULL;
}

int64_t mco_db_add_targets_restricted(mco_db* d, const char* buf, const int64_t* offs, const uint32_t* ids, int64_t n,
                                      const uint32_t* allowed, int64_t nallowed, int threads)
{
    if(threads < 1) threads = 1;
    if(threads > 256) threads = 256;
    pthread_t th[256];
    rs_arg args[256];
    int64_t next = 0, wins = 0;
    pthread_mutex_t mtx = PTHREAD_MUTEX_INITIALIZER;
    for(int t = 0; t < threads; ++t) {
        args[t] = (rs_arg){d, buf, offs, ids, n, allowed, nallowed, &next, &mtx, NULL, 0, 0, 0};
        pthread_create(&th[t], NULL, rs_worker, &args[t]);
    }
    for(int t = 0; t < threads; ++t) {
        pthread_join(th[t], NULL);
        if(d->ntup + args[t].ntup > d->captup) {
            d->captup = d->captup ? d->captup : (1u << 16);
            while(d->captup < d->ntup + args[t].ntup) d->captup *= 2;
            d->tup = (tuple3*)realloc(d->tup, d->captup * sizeof(tuple3));
        }
        if(args[t].ntup) memcpy(d->tup + d->ntup, args[t].tup, args[t].ntup * sizeof(tuple3));
        d->ntup += args[t].ntup;
        wins += args[t].wins;
        free(args[t].tup);
    }
    d->finalized = 0;          /* finalize sorts by (key, tgt, win): the order the threads finished in is irrelevant */
    return wins;
}

static int cmp_tuple(const void* a, const void* b)
{
    const tuple3* x = (const tuple3*)a;
    const tuple3* y = (const tuple3*)b;
    if(x->key != y->key) return x->key < y->key ? -1 : 1;
    if(x->tgt != y->tgt) return x->tgt < y->tgt ? -1 : 1;
    if(x->win != y->win) return x->win < y->win ? -1 : 1;
    return 0;
}

/* hash_multimap.h:653 insert + sketch_database.h:1089-1092 shrink: each key
 * keeps its FIRST maxlocs locations in insertion order = ascending (tgt,win)
 * because targets arrive in ascending id (quirk Q11) */
void mco_db_finalize(mco_db* d)
{
    free(d->keys); free(d->off); free(d->size); free(d->ltgt); free(d->lwin);
    qsort(d->tup, d->ntup, sizeof(tuple3), cmp_tuple);
    uint64_t nk = 0;
    for(uint64_t i = 0; i < d->ntup; ++i) if(i == 0 || d->tup[i].key != d->tup[i - 1].key) ++nk;
    d->keys = (uint32_t*)malloc((nk + 1) * sizeof(uint32_t));
    d->off  = (uint64_t*)malloc((nk + 1) * sizeof(uint64_t));
    d->size = (uint8_t*)malloc(nk + 1);
    d->ltgt = (uint32_t*)malloc((d->ntup + 1) * sizeof(uint32_t));
    d->lwin = (uint32_t*)malloc((d->ntup + 1) * sizeof(uint32_t));
    uint64_t ki = 0, li = 0, i = 0;
    while(i < d->ntup) {
        uint64_t j = i;
        while(j < d->ntup && d->tup[j].key == d->tup[i].key) ++j;
        uint64_t n = j - i;
        if(n > (uint64_t)d->maxlocs) n = (uint64_t)d->maxlocs;
        d->keys[ki] = d->tup[i].key;
        d->off[ki] = li;
        d->size[ki] = (uint8_t)n;
        for(uint64_t t = 0; t < n; ++t) { d->ltgt[li] = d->tup[i + t].tgt; d->lwin[li] = d->tup[i + t].win; ++li; }
        ++ki;
        i = j;
    }
    d->nkeys = nk;
    d->nloc = li;
    d->finalized = 1;
}

void mco_db_counts(const mco_db* d, uint64_t* nkeys, uint64_t* nkeys_nonempty, uint64_t* nlocs)
{
    uint64_t ne = 0, nl = 0;
    for(uint64_t i = 0; i < d->nkeys; ++i) if(d->size[i]) { ++ne; nl += d->size[i]; }
    if(nkeys) *nkeys = d->nkeys;
    if(nkeys_nonempty) *nkeys_nonempty = ne;
    if(nlocs) *nlocs = nl;
}

void mco_db_export(const mco_db* d, uint32_t* keys, uint32_t* sizes, uint32_t* tgts, uint32_t* wins)
{
    uint64_t ko = 0, lo = 0;
    for(uint64_t i = 0; i < d->nkeys; ++i) {
        if(!d->size[i]) continue;
        keys[ko] = d->keys[i];
        sizes[ko] = d->size[i];
        ++ko;
        for(unsigned t = 0; t < d->size[i]; ++t) {
            tgts[lo] = d->ltgt[d->off[i] + t];
            wins[lo] = d->lwin[d->off[i] + t];
            ++lo;
        }
    }
}

static int64_t db_key_index(const mco_db* d, uint32_t key)
{
    int64_t lo = 0, hi = (int64_t)d->nkeys - 1;
    while(lo <= hi) {
        const int64_t mid = (lo + hi) / 2;
        if(d->keys[mid] == key) return mid;
        if(d->keys[mid] < key) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
}

/* hash_multimap.h:779-789 find */
int mco_db_find(const mco_db* d, uint32_t key, uint32_t* tgts, uint32_t* wins)
{
    const int64_t i = db_key_index(d, key);
    if(i < 0) return 0;
    for(unsigned t = 0; t < d->size[i]; ++t) {
        if(tgts) tgts[t] = d->ltgt[d->off[i] + t];
        if(wins) wins[t] = d->lwin[d->off[i] + t];
    }
    return d->size[i];
}

/* mode_build.cpp:847-1074 (-remove-overpopulated-features): sum over ranks of
 * the per-rank (already capped) bucket sizes of every non-empty bucket
 * (sketch_database.h:1060-1075); keys whose sum is > maxlocs-1 (:1032-1040)
 * are cleared on every rank (sketch_database.h:399-417, hash_multimap.h:720) */
uint64_t mco_remove_overpopulated(mco_db** sh, int P)
{
    uint64_t removed = 0;
    uint64_t idx[64] = {0};
    if(P > 64) return 0;
    const uint32_t limit = (uint32_t)sh[0]->maxlocs - 1u;
    for(;;) {
        /* smallest current key over the shards (k-way merge of sorted key arrays) */
        int have = 0;
        uint32_t key = 0;
        for(int r = 0; r < P; ++r) {
            if(idx[r] < sh[r]->nkeys && (!have || sh[r]->keys[idx[r]] < key)) { key = sh[r]->keys[idx[r]]; have = 1; }
        }
        if(!have) break;
        uint32_t total = 0;
        for(int r = 0; r < P; ++r)
            if(idx[r] < sh[r]->nkeys && sh[r]->keys[idx[r]] == key) total += sh[r]->size[idx[r]];
        for(int r = 0; r < P; ++r) {
            if(idx[r] < sh[r]->nkeys && sh[r]->keys[idx[r]] == key) {
                if(total > limit && sh[r]->size[idx[r]] > 0) { sh[r]->size[idx[r]] = 0; ++removed; }
                ++idx[r];
            }
        }
    }
    return removed;
}

/* ======================================================================== */
/* L3/L4: query                                                              */
/* ======================================================================== */

typedef struct { uint32_t tgt, win; } loc2;

static int cmp_loc(const void* a, const void* b)
{
    const loc2* x = (const loc2*)a;
    const loc2* y = (const loc2*)b;
    if(x->tgt != y->tgt) return x->tgt < y->tgt ? -1 : 1;        /* sketch_database.h:172-177 */
    if(x->win != y->win) return x->win < y->win ? -1 : 1;
    return 0;
}

/* sketch_database.h:806-823 accumulate_matches for both mates (querying.h:
 * 798-799) followed by the (tgt,win) sort of querying.h:90-106.  Equal keys
 * are indistinguishable, so any correct sort reproduces merge_sort's result. */
static loc2* gather_sorted(const mco_db* d, const mco_qopt* o, const char* s1, int64_t n1,
                           const char* s2, int64_t n2, int64_t* count)
{
    loc2* m = NULL;
    int64_t nm = 0, cap = 0;
    uint32_t sk[256], lt[256], lw[256];
    for(int mate = 0; mate < 2; ++mate) {
        const char* s = mate ? s2 : s1;
        const int64_t len = mate ? n2 : n1;
        /* an empty mate still yields one (empty) window: dna_encoding.h:269 */
        const int64_t nw = mco_num_windows(len, o->winlen, o->stride);
        for(int64_t w = 0; w < nw; ++w) {
            int64_t b, e;
            mco_window(len, o->winlen, o->stride, w, &b, &e);
            const int c = mco_sketch_window(s + b, e - b, d->k, o->sketchlen, sk);
            for(int i = 0; i < c; ++i) {
                const int n = mco_db_find(d, sk[i], lt, lw);
                if(nm + n > cap) { cap = cap ? cap * 2 : 1024; while(cap < nm + n) cap *= 2; m = (loc2*)realloc(m, (size_t)cap * sizeof(loc2)); }
                for(int t = 0; t < n; ++t) m[nm++] = (loc2){lt[t], lw[t]};
            }
        }
    }
    if(nm > 1) qsort(m, (size_t)nm, sizeof(loc2), cmp_loc);
    *count = nm;
    return m;
}

int64_t mco_matches(const mco_db* d, const mco_qopt* o, const char* s1, int64_t n1,
                    const char* s2, int64_t n2, uint32_t* tgts, uint32_t* wins, int64_t cap)
{
    int64_t nm;
    loc2* m = gather_sorted(d, o, s1, n1, s2, n2, &nm);
    for(int64_t i = 0; i < nm && i < cap; ++i) { tgts[i] = m[i].tgt; wins[i] = m[i].win; }
    free(m);
    return nm;
}

typedef struct { int32_t tax; uint32_t hits; } cand_t;
typedef struct { cand_t* c; int n, cap; } candlist;

/* candidates.h:236-285 best_distinct_matches_in_contiguous_window_ranges::insert.
 * The std::sort of :270 runs on <= 16 elements in the supported configurations,
 * where libstdc++ uses insertion sort: an element whose hits grew moves left
 * past strictly smaller hits only (stable). */
static void cand_insert(candlist* L, const mco_tax* tx, int32_t tax, uint32_t hits,
                        int64_t maxcand, int lowest)
{
    if(tax < 0) return;
    if(lowest > MCO_RANK_SEQUENCE) {
        const int32_t anc = tx->lineage[(int64_t)tax * MCO_NUM_RANKS + lowest];
        if(anc >= 0) tax = anc;
    }
    int pos = -1;
    if(tx->rank[tax] != MCO_RANK_SEQUENCE) {
        for(int i = 0; i < L->n; ++i) if(L->c[i].tax == tax) { pos = i; break; }
    }
    if(pos >= 0) {
        if(hits > L->c[pos].hits) L->c[pos].hits = hits;
        const cand_t v = L->c[pos];
        int j = pos;
        while(j > 0 && L->c[j - 1].hits < v.hits) { L->c[j] = L->c[j - 1]; --j; }
        L->c[j] = v;
        return;
    }
    int i = 0;
    while(i < L->n && L->c[i].hits >= hits) ++i;           /* upper_bound, descending */
    if(i < L->n || (int64_t)L->n < maxcand) {
        if(L->n + 1 > L->cap) { L->cap = L->cap ? L->cap * 2 : 8; L->c = (cand_t*)realloc(L->c, (size_t)L->cap * sizeof(cand_t)); }
        for(int j = L->n; j > i; --j) L->c[j] = L->c[j - 1];
        L->c[i] = (cand_t){tax, hits};
        L->n++;
        if((int64_t)L->n > maxcand) L->n = (int)maxcand;
    }
}

static int64_t eff_maxcand(int maxcand) { return maxcand < 1 ? INT64_MAX : maxcand; }   /* query_options.cpp:176-178 */

/* candidates.h:119-180 for_all_contiguous_window_ranges +
 * classification.cpp:210-226 make_classification_candidates */
int mco_query_local(const mco_db* d, const mco_tax* tx, const mco_qopt* o,
                    const char* s1, int64_t n1, const char* s2, int64_t n2,
                    int32_t* out_tax, uint32_t* out_hits, int cap)
{
    int64_t nm;
    loc2* m = gather_sorted(d, o, s1, n1, s2, n2, &nm);
    int64_t span = n1 + n2;
    if(o->insertsize > span) span = o->insertsize;
    const uint32_t numwin = (uint32_t)(2 + span / d->stride);
    candlist L = {0};
    int64_t i = 0;
    while(i < nm) {
        int64_t j = i;
        while(j < nm && m[j].tgt == m[i].tgt) ++j;
        uint32_t best = 1;
        int64_t fst = i;
        for(int64_t lst = i + 1; lst < j; ++lst) {
            while(fst != lst && (uint32_t)(m[lst].win - m[fst].win) >= numwin) ++fst;
            const uint32_t h = (uint32_t)(lst - fst + 1);
            if(h > best) best = h;
        }
        /* taxon of target t is index t (sequence level) */
        cand_insert(&L, tx, (int32_t)m[i].tgt, best, eff_maxcand(o->maxcand), o->lowest);
        i = j;
    }
    free(m);
    for(int c = 0; c < L.n && c < cap; ++c) { out_tax[c] = L.c[c].tax; out_hits[c] = L.c[c].hits; }
    const int n = L.n;
    free(L.c);
    return n;
}

/* taxonomy.h:531-537 ranked_lca */
static int32_t ranked_lca(const mco_tax* tx, int32_t a, int32_t b)
{
    if(a < 0 || b < 0) return -1;
    const int32_t* la = tx->lineage + (int64_t)a * MCO_NUM_RANKS;
    const int32_t* lb = tx->lineage + (int64_t)b * MCO_NUM_RANKS;
    for(int r = 0; r <= MCO_RANK_ROOT; ++r) if(la[r] >= 0 && la[r] == lb[r]) return la[r];
    return -1;
}

/* querying.h:905-971 (merge = insert in rank order) and
 * classification.cpp:236-265 classify */
int mco_merge_classify(const mco_tax* tx, const mco_qopt* o,
                       const int32_t* in_tax, const uint32_t* in_hits, int nin,
                       int32_t* out_tax, uint32_t* out_hits, int cap,
                       int32_t* best_tax, int* best_rank)
{
    candlist L = {0};
    for(int i = 0; i < nin; ++i) cand_insert(&L, tx, in_tax[i], in_hits[i], eff_maxcand(o->maxcand), o->lowest);
    for(int c = 0; c < L.n && c < cap; ++c) { out_tax[c] = L.c[c].tax; out_hits[c] = L.c[c].hits; }

    int32_t lca = -1;
    if(L.n > 0 && L.c[0].hits >= (uint32_t)o->hitmin) {
        lca = L.c[0].tax;
        const float thr = L.c[0].hits > (uint32_t)o->hitmin
                        ? (float)(L.c[0].hits - (uint32_t)o->hitmin) * o->hitdiff : 0.0f;
        for(int i = 1; i < L.n; ++i) {
            if((float)L.c[i].hits > thr) {
                lca = ranked_lca(tx, lca, L.c[i].tax);
                if(lca < 0 || tx->rank[lca] > o->highest) { lca = -1; break; }
            } else break;
        }
        if(lca >= 0 && tx->rank[lca] > o->highest) lca = -1;
    }
    *best_tax = lca;
    *best_rank = lca >= 0 ? tx->rank[lca] : MCO_RANK_NONE;
    const int n = L.n;
    free(L.c);
    return n;
}

/* The same with the reference's wire format reproduced: taxon ids travel as
 * uint32 (querying.h:984,1001) and are widened unsigned on receipt (:958), so a sequence-level taxon
 * (id -(tgt)-1, dense index < ntargets here) received from another rank names no taxon and is dropped
 * (candidates.h:240).  Lists move along the reduction tree of querying.h:892-1066 (P a power of two):
 * round 1 (1->0), (3->2), ...; round 2 (2->0), (6->4), ...; a receiver keeps its own list whole
 * (:910-940), a sender ships its merged, truncated list (:994-1009).  in_n[r] = length of rank r's list. */
static void compat_reduce(const mco_tax* tx, const mco_qopt* o, candlist* lists, int lo, int hi)
{
    /* result of ranks [lo, hi) ends in lists[lo] */
    if(hi - lo <= 1) return;
    const int mid = lo + (hi - lo) / 2;
    compat_reduce(tx, o, lists, lo, mid);
    compat_reduce(tx, o, lists, mid, hi);
    for(int c = 0; c < lists[mid].n; ++c) {
        if((int64_t)lists[mid].c[c].tax < tx->ntargets) continue;          /* negative id: lost on the wire */
        cand_insert(&lists[lo], tx, lists[mid].c[c].tax, lists[mid].c[c].hits, eff_maxcand(o->maxcand), o->lowest);
    }
}

int mco_merge_classify_compat(const mco_tax* tx, const mco_qopt* o, int P,
                              const int32_t* in_tax, const uint32_t* in_hits, const int* in_n,
                              int32_t* out_tax, uint32_t* out_hits, int cap,
                              int32_t* best_tax, int* best_rank)
{
    candlist* lists = (candlist*)calloc((size_t)P, sizeof(candlist));
    int at = 0;
    for(int r = 0; r < P; ++r) {
        for(int c = 0; c < in_n[r]; ++c) cand_insert(&lists[r], tx, in_tax[at + c], in_hits[at + c], eff_maxcand(o->maxcand), o->lowest);
        at += in_n[r];
    }
    if((P & (P - 1)) == 0) compat_reduce(tx, o, lists, 0, P);
    else                                                    /* no tree exists in the reference (Q12): rank order */
        for(int r = 1; r < P; ++r)
            for(int c = 0; c < lists[r].n; ++c) {
                if((int64_t)lists[r].c[c].tax < tx->ntargets) continue;
                cand_insert(&lists[0], tx, lists[r].c[c].tax, lists[r].c[c].hits, eff_maxcand(o->maxcand), o->lowest);
            }
    /* decision on the merged list: same code as mco_merge_classify (re-inserting a valid list reproduces it) */
    int nin = lists[0].n;
    int32_t* t = (int32_t*)malloc((size_t)(nin > 0 ? nin : 1) * sizeof(int32_t));
    uint32_t* h = (uint32_t*)malloc((size_t)(nin > 0 ? nin : 1) * sizeof(uint32_t));
    for(int c = 0; c < nin; ++c) { t[c] = lists[0].c[c].tax; h[c] = lists[0].c[c].hits; }
    const int n = mco_merge_classify(tx, o, t, h, nin, out_tax, out_hits, cap, best_tax, best_rank);
    free(t); free(h);
    for(int r = 0; r < P; ++r) free(lists[r].c);
    free(lists);
    return n;
}

/* classification_statistics.h:70-78 assign */
void mco_statistics(const int* ranks, int64_t n, uint64_t* out22)
{
    for(int r = 0; r < 22; ++r) out22[r] = 0;
    for(int64_t i = 0; i < n; ++i) {
        if(ranks[i] >= MCO_RANK_NONE) out22[MCO_RANK_NONE]++;
        else for(int r = ranks[i]; r <= MCO_RANK_ROOT; ++r) out22[r]++;
    }
}

void mco_query_batch(mco_db* const* shards, int P, const mco_tax* tx, const mco_qopt* o,
                     const char* buf1, const int64_t* offs1,
                     const char* buf2, const int64_t* offs2, int64_t nreads,
                     int32_t* out_tax, uint32_t* out_hits, int32_t* out_ncand,
                     int32_t* out_best, int32_t* out_best_rank, uint64_t* out_stats22)
{
    const int K = o->maxcand;
    int32_t* lt = (int32_t*)malloc((size_t)P * K * sizeof(int32_t));
    uint32_t* lh = (uint32_t*)malloc((size_t)P * K * sizeof(uint32_t));
    int* ranks = (int*)malloc((size_t)(nreads > 0 ? nreads : 1) * sizeof(int));
    for(int64_t q = 0; q < nreads; ++q) {
        const char* s1 = buf1 + offs1[q];
        const int64_t n1 = offs1[q + 1] - offs1[q];
        const char* s2 = buf2 ? buf2 + offs2[q] : "";
        const int64_t n2 = buf2 ? offs2[q + 1] - offs2[q] : 0;
        int nin = 0;
        int per_rank[256];
        for(int r = 0; r < P; ++r) {
            const int c = mco_query_local(shards[r], tx, o, s1, n1, s2, n2, lt + nin, lh + nin, K);
            if(r < 256) per_rank[r] = c;
            nin += c;
        }
        int32_t best; int brank;
        const int n = (o->reference_compat && P <= 256)
                    ? mco_merge_classify_compat(tx, o, P, lt, lh, per_rank, out_tax + q * K, out_hits + q * K, K, &best, &brank)
                    : mco_merge_classify(tx, o, lt, lh, nin, out_tax + q * K, out_hits + q * K, K, &best, &brank);
        for(int c = n; c < K; ++c) { out_tax[q * K + c] = -1; out_hits[q * K + c] = 0; }
        out_ncand[q] = n;
        out_best[q] = best;
        out_best_rank[q] = brank;
        ranks[q] = brank;
    }
    if(out_stats22) mco_statistics(ranks, nreads, out_stats22);
    free(lt); free(lh); free(ranks);
}
