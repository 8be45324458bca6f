/* oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, not product code.
 *
 * Thin extern "C" wrapper around the UNMODIFIED reference sources, included /
 * linked from where they lie under /root/reference/src (nothing is copied).
 * It exposes the reference's own classes at every seam of the hot path so that
 * the plain-C restatement (mc_oracle.c) and the CUDA path can be pinned
 * stage by stage against the real thing:
 *
 *   sketch        -> mc::sketcher::operator()            hash_dna.h:115
 *   windows       -> mc::for_each_window                 dna_encoding.h:261
 *   build         -> database::add_target_distributed    sketch_database.h:519
 *   shard content -> database::print_feature_map         sketch_database.h:1036
 *   find + gather -> database::accumulate_matches        sketch_database.h:806
 *   sort          -> mc::merge_sort                      querying.h:90
 *   candidates    -> classification_candidates{...}      candidates.h:207
 *   merge         -> classification_candidates::insert   candidates.h:236
 *   decision      -> mc::classify                        classification.cpp:236
 *   statistics    -> classification_statistics::assign   classification_statistics.h:70
 *
 * Built by oracle/Makefile into oracle/_ref/libref_l1.so together with all
 * reference .cpp files except main.cpp.  Only tests/, bench.py's cpu_baseline
 * / --impl reference legs and the golden-vector generator may load it.
 */
#include <cstdint>
#include <cstring>
#include <sstream>
#include <string>
#include <thread>
#include <vector>
#include <atomic>
#include <chrono>

#include "timer.h"
#include "config.h"
#include "dna_encoding.h"
#include "hash_int.h"
#include "hash_dna.h"
#include "sketch_database.h"
#include "candidates.h"
#include "querying.h"
#include "query_options.h"
#include "classification.h"
#include "classification_statistics.h"

namespace mc {
/* defined in /root/reference/src/classification.cpp:236 (external linkage) */
const taxon* classify(const database& db, const classification_options& opt,
                      const classification_candidates& cand);
}

using namespace mc;

extern "C" {

/* ---------------------------------------------------------------- L1 ---- */
uint32_t ref_tmh(uint32_t x) { return thomas_mueller_hash(x); }

uint32_t ref_revcomp(uint32_t x, int k) { return make_reverse_complement(x, numk_t(k)); }

uint32_t ref_canonical(uint32_t x, int k) { return make_canonical(x, numk_t(k)); }

/* all (kmer, ambig) pairs of a string; returns count */
int64_t ref_kmers(const char* s, int64_t n, int k, uint32_t* kmers, uint16_t* ambig)
{
    int64_t c = 0;
    if(n < k) return 0;
    for_each_kmer_2bit<uint32_t>(numk_t(k), s, s + n,
        [&](uint32_t km, half_size_t<uint32_t> a) { kmers[c] = km; ambig[c] = a; ++c; });
    return c;
}

/* window decomposition; returns number of windows, begin/end may be NULL */
int64_t ref_windows(int64_t len, int64_t winlen, int64_t stride, int64_t* beg, int64_t* end)
{
    std::vector<char> dummy(size_t(len > 0 ? len : 1));
    const char* b0 = dummy.data();
    int64_t n = 0;
    for_each_window(b0, b0 + len, size_t(winlen), size_t(stride),
        [&](const char* b, const char* e) {
            if(beg) beg[n] = b - b0;
            if(end) end[n] = e - b0;
            ++n;
        });
    return n;
}

/* sketch of one window; out must hold sketchlen entries; returns count */
int ref_sketch_window(const char* s, int64_t n, int k, int sketchlen, uint32_t* out)
{
    sketcher sk;
    sk.kmer_size(k);
    sk.sketch_size(sketchlen);
    auto r = sk(s, s + n);
    for(size_t i = 0; i < r.size(); ++i) out[i] = r[i];
    return int(r.size());
}

/* sketches of all windows of one sequence: out_feat[nwin][sketchlen] padded
 * with 0xFFFFFFFF, out_cnt[nwin]; returns nwin */
int64_t ref_sketch_sequence(const char* s, int64_t n, int k, int sketchlen,
                            int64_t winlen, int64_t stride,
                            uint32_t* out_feat, uint32_t* out_cnt)
{
    sketcher sk;
    sk.kmer_size(k);
    sk.sketch_size(sketchlen);
    int64_t w = 0;
    for_each_window(s, s + n, size_t(winlen), size_t(stride),
        [&](const char* b, const char* e) {
            auto r = sk(b, e);
            uint32_t* o = out_feat + w * sketchlen;
            for(int i = 0; i < sketchlen; ++i) o[i] = (size_t(i) < r.size()) ? r[i] : 0xFFFFFFFFu;
            out_cnt[w] = uint32_t(r.size());
            ++w;
        });
    return w;
}

/* CPU baseline: sketch nseq sequences with `threads` threads (dynamic work
 * queue over sequences, the analogue of target_id % num_procs partitioning);
 * returns seconds; *checksum = sum of all features (so the work is not elided) */
double ref_sketch_throughput(const char* buf, const int64_t* offs, int64_t nseq,
                             int k, int sketchlen, int64_t winlen, int64_t stride,
                             int threads, uint64_t* checksum, uint64_t* nwindows)
{
    std::atomic<int64_t> next{0};
    std::atomic<uint64_t> sum{0}, wins{0};
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for(int t = 0; t < threads; ++t) {
        th.emplace_back([&] {
            sketcher sk;
            sk.kmer_size(k);
            sk.sketch_size(sketchlen);
            uint64_t ls = 0, lw = 0;
            for(;;) {
                int64_t i = next.fetch_add(1);
                if(i >= nseq) break;
                const char* s = buf + offs[i];
                const char* e = buf + offs[i + 1];
                for_each_window(s, e, size_t(winlen), size_t(stride),
                    [&](const char* b, const char* we) {
                        auto r = sk(b, we);
                        for(auto f : r) ls += f;
                        ++lw;
                    });
            }
            sum += ls;
            wins += lw;
        });
    }
    for(auto& x : th) x.join();
    auto t1 = std::chrono::steady_clock::now();
    if(checksum) *checksum = sum.load();
    if(nwindows) *nwindows = wins.load();
    return std::chrono::duration<double>(t1 - t0).count();
}

/* ---------------------------------------------------------- database ---- */
struct refdb {
    database db;
    explicit refdb(const sketcher& sk): db{sk} {}
};

void* refdb_create(int k, int sketchlen, int64_t winlen, int64_t stride, int maxlocs)
{
    sketcher sk;
    sk.kmer_size(k);
    sk.sketch_size(sketchlen);
    auto* h = new refdb{sk};
    h->db.target_window_size(size_t(winlen));
    h->db.target_window_stride(size_t(stride));
    h->db.query_window_size(size_t(winlen));
    h->db.query_window_stride(size_t(stride));
    /* narrowing to the bucket size type happens here exactly as in
     * mode_build.cpp:659 / sketch_database.h:356 (quirk Q7) */
    h->db.max_locations_per_feature(database::bucket_size_type(maxlocs));
    return h;
}

void refdb_destroy(void* p) { delete static_cast<refdb*>(p); }

int refdb_max_locations(void* p) { return int(static_cast<refdb*>(p)->db.max_locations_per_feature()); }

void refdb_set_query_params(void* p, int sketchlen, int64_t winlen, int64_t stride)
{
    auto& db = static_cast<refdb*>(p)->db;
    sketcher sk = db.query_sketcher();
    sk.sketch_size(sketchlen);
    db.query_sketcher(sk);
    db.query_window_size(size_t(winlen));
    db.query_window_stride(size_t(stride));
}

/* taxa above sequence level; ranks use taxonomy::rank numbering (taxonomy.h:61) */
void refdb_set_taxonomy(void* p, const int64_t* ids, const int64_t* parents,
                        const uint8_t* ranks, int64_t n)
{
    auto& db = static_cast<refdb*>(p)->db;
    taxonomy tax;
    for(int64_t i = 0; i < n; ++i) {
        tax.emplace(ids[i], parents[i], "taxon" + std::to_string(ids[i]),
                    taxonomy::rank(ranks[i]));
    }
    db.reset_taxa_above_sequence_level(std::move(tax));
}

/* returns 1 if the target was accepted (counted), 0 otherwise */
int refdb_add_target(void* p, const char* seq, int64_t len, const char* name,
                     int64_t parent_taxid, int num_procs, int my_id)
{
    auto& db = static_cast<refdb*>(p)->db;
    std::string s(seq, size_t(len));
    return db.add_target_distributed(s, name, parent_taxid,
                                     database::file_source{}, num_procs, my_id) ? 1 : 0;
}

uint64_t refdb_target_count(void* p) { return static_cast<refdb*>(p)->db.target_count(); }
uint64_t refdb_feature_count(void* p) { return static_cast<refdb*>(p)->db.feature_count(); }
uint64_t refdb_dead_feature_count(void* p) { return static_cast<refdb*>(p)->db.dead_feature_count(); }
uint64_t refdb_location_count(void* p) { return static_cast<refdb*>(p)->db.location_count(); }

/* shard contents through the reference's own print_feature_map
 * (sketch_database.h:1036): non-empty buckets in table order.
 * Call with NULL arrays to obtain counts. */
void refdb_dump(void* p, uint64_t* nkeys, uint64_t* nlocs,
                uint32_t* keys, uint32_t* sizes, uint32_t* tgts, uint32_t* wins)
{
    auto& db = static_cast<refdb*>(p)->db;
    std::ostringstream os;
    db.print_feature_map(os);
    const std::string txt = os.str();
    uint64_t nk = 0, nl = 0;
    const char* c = txt.c_str();
    const char* e = c + txt.size();
    while(c < e) {
        /* "<key> -> (t,w)(t,w)...\n" */
        char* q = nullptr;
        long long key = std::strtoll(c, &q, 10);
        c = q;
        while(c < e && *c != '(' && *c != '\n') ++c;
        uint32_t cnt = 0;
        while(c < e && *c == '(') {
            ++c;
            long long t = std::strtoll(c, &q, 10); c = q + 1;   /* skip ',' */
            long long w = std::strtoll(c, &q, 10); c = q + 1;   /* skip ')' */
            if(tgts) { tgts[nl] = uint32_t(t); wins[nl] = uint32_t(w); }
            ++nl; ++cnt;
        }
        if(keys) { keys[nk] = uint32_t(key); sizes[nk] = cnt; }
        ++nk;
        while(c < e && *c != '\n') ++c;
        ++c;
    }
    if(nkeys) *nkeys = nk;
    if(nlocs) *nlocs = nl;
}

/* (key,size) pairs as the reduce of mode_build.cpp:847 sees them */
uint64_t refdb_remove_keys(void* p, const uint32_t* keys, uint32_t n)
{
    auto& db = static_cast<refdb*>(p)->db;
    return db.remove_features_with_more_locations_than_distributed(
        const_cast<uint32_t*>(keys), n);
}

/* ------------------------------------------------------------- query ---- */
/* find + gather + sort for one read pair (querying.h:792-811).
 * matches returned as (tgt,win) in sorted order; returns number of matches
 * (call with cap=0 to get the count only) */
int64_t refdb_matches(void* p, const char* s1, int64_t n1, const char* s2, int64_t n2,
                      uint32_t* tgts, uint32_t* wins, int64_t cap)
{
    auto& db = static_cast<refdb*>(p)->db;
    match_target_locations buf, buf2;
    std::vector<size_t> offsets{0};
    db.accumulate_matches(std::string(s1, size_t(n1)), buf, offsets);
    db.accumulate_matches(std::string(s2, size_t(n2)), buf, offsets);
    merge_sort(buf, offsets, buf2);
    const int64_t m = int64_t(buf.size());
    for(int64_t i = 0; i < m && i < cap; ++i) { tgts[i] = buf[i].tgt; wins[i] = buf[i].win; }
    return m;
}

static classification_options make_copt(int maxcand, int lowest, int highest,
                                        int hitmin, float hitdiff, int64_t insertsize)
{
    classification_options o;
    o.lowestRank  = taxon_rank(lowest);
    o.highestRank = taxon_rank(highest);
    o.hitsMin = uint16_t(hitmin);
    o.hitsDiffFraction = hitdiff;
    o.insertSizeMax = size_t(insertsize);
    /* query_options.cpp:176-178: < 1 means unlimited */
    o.maxNumCandidatesPerQuery = maxcand < 1 ? std::numeric_limits<size_t>::max() : size_t(maxcand);
    return o;
}

/* per-shard candidate list of one read pair (querying.h:792-814 +
 * classification.cpp:210-226); taxon ids as int64; returns ncand */
int refdb_query_local(void* p, const char* s1, int64_t n1, const char* s2, int64_t n2,
                      int maxcand, int lowest, int64_t insertsize,
                      int64_t* out_tax, uint64_t* out_hits, int cap)
{
    auto& db = static_cast<refdb*>(p)->db;
    match_target_locations buf, buf2;
    std::vector<size_t> offsets{0};
    const std::string q1(s1, size_t(n1)), q2(s2, size_t(n2));
    db.accumulate_matches(q1, buf, offsets);
    db.accumulate_matches(q2, buf, offsets);
    merge_sort(buf, offsets, buf2);
    match_locations matches;
    for(auto& m : buf) matches.emplace_back(db.taxon_of_target(m.tgt), m.win);

    candidate_generation_rules rules;
    rules.maxWindowsInRange = window_id(2 + (std::max(q1.size() + q2.size(), size_t(insertsize))
                                             / db.target_window_stride()));
    rules.mergeBelow = taxon_rank(lowest);
    rules.maxCandidates = maxcand < 1 ? std::numeric_limits<size_t>::max() : size_t(maxcand);
    classification_candidates cand{db, matches, rules};
    int n = 0;
    for(auto& c : cand) {
        if(n < cap) { out_tax[n] = c.tax->id(); out_hits[n] = c.hits; }
        ++n;
    }
    return n;
}

/* cross-rank merge = repeated insert (querying.h:905-971) followed by the
 * decision (classification.cpp:236-265).  in_tax/in_hits hold the per-rank
 * lists concatenated in rank order; returns the merged count, fills the merged
 * list, *best_tax = taxon id of the classification (0 = none), *best_rank */
int refdb_merge_classify(void* p, const int64_t* in_tax, const uint64_t* in_hits, int nin,
                         int maxcand, int lowest, int highest, int hitmin, float hitdiff,
                         int64_t* out_tax, uint64_t* out_hits, int cap,
                         int64_t* best_tax, int* best_rank)
{
    auto& db = static_cast<refdb*>(p)->db;
    auto opt = make_copt(maxcand, lowest, highest, hitmin, hitdiff, 0);
    candidate_generation_rules rules;
    rules.mergeBelow = opt.lowestRank;
    rules.maxCandidates = opt.maxNumCandidatesPerQuery;
    classification_candidates merged;
    for(int i = 0; i < nin; ++i) {
        match_candidate c{db.taxon_with_id(in_tax[i]), in_hits[i]};
        merged.insert(c, db, rules);
    }
    int n = 0;
    for(auto& c : merged) {
        if(n < cap) { out_tax[n] = c.tax->id(); out_hits[n] = c.hits; }
        ++n;
    }
    const taxon* best = classify(db, opt, merged);
    *best_tax = best ? best->id() : 0;
    *best_rank = best ? int(best->rank()) : int(taxon_rank::none);
    return n;
}

/* ranked lineage of a taxon id (taxonomy.h:576): 21 ids, 0 = null */
void refdb_lineage(void* p, int64_t taxid, int64_t* out21)
{
    auto& db = static_cast<refdb*>(p)->db;
    const taxon* t = db.taxon_with_id(taxid);
    for(int i = 0; i < taxonomy::num_ranks; ++i) out21[i] = 0;
    if(!t) return;
    const auto& lin = db.ranks(t);
    for(int i = 0; i < taxonomy::num_ranks; ++i) out21[i] = lin[i] ? lin[i]->id() : 0;
}

/* classification_statistics::assign for a list of ranks -> 22 counters
 * (index = rank, 21 = none) via the reference's own accessors */
void ref_statistics(const int* ranks, int64_t n, uint64_t* out22)
{
    classification_statistics st;
    for(int64_t i = 0; i < n; ++i) st.assign(taxon_rank(ranks[i]));
    for(int r = 0; r < 21; ++r) out22[r] = st.assigned(taxon_rank(r));
    out22[21] = st.unassigned();
}

int ref_num_ranks() { return taxonomy::num_ranks; }

} /* extern "C" */
