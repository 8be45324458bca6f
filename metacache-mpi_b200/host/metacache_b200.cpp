// metacache_b200 -- C++ host driver of the B200 build/query path (links libmcgpu.so through the
// C-ABI of include/mcgpu.h only).  It keeps the reference's command line, shard files and
// output format for that path:
//
//   metacache_b200 build <db> <fasta files / directories ...> [-gpus P] [-kmerlen 16]
//       [-sketchlen 16] [-winlen 128] [-winstride winlen-k+1] [-max-locations-per-feature 254]
//       [-remove-overpopulated-features] [-taxonomy dir] [-taxpostmap file]  (mode_build.cpp:93-138; an
//       `assembly_summary.txt` next to the sequence files names their taxa, taxonomy_io.cpp:190-311; targets still without a
//       taxon are looked up in <dir>/*.accession2taxid, mode_build.cpp:493-522)
//   metacache_b200 query <db> <r1.fq> <r2.fq> -pairfiles [-gpus P] [-lowest rank] [-highest rank]
//       [-hitmin n] [-hitdiff f] [-maxcand n] [-insertsize n] [-winlen n] [-winstride n]
//       [-sketchlen n] [-out file] [-tophits] [-queryids] [-taxids|-taxids-only] [-lineage]
//       [-mapped-only|-nomap] [-no-summary] [-no-query-params] [-abundances [file]] [-abundance-per rank]
//       [-hits-per-seq [file]] [-reference-compat]                       (query_options.cpp)
//   metacache_b200 merge <result files ...> -taxonomy dir [-lowest rank] [-highest rank] [-hitmin n] [-hitdiff f]
//       [-maxcand n] [-out file] [-tophits] [-queryids] ...              (mode_merge.cpp, docs/merge.txt)
//
// `-threads T` and `-query-limit n` are accepted and have no effect: in the reference they only set the size of the
// blocks its workers take (querying.h:764-765, 1321); every read is classified either way (checked with the reference:
// `-threads 2 -query-limit 100` on 300 pairs prints 300 lines).
//
// `-gpus P` replaces `mpirun -np P`: the database has P shards (shard = target_id % P,
// sketch_database.h:540), shard r is written to / read from `<db>.db_<r>` in the reference's own
// binary format (sketch_database.h:858-999, hash_multimap.h:923-1029) and lives on GPU
// r % device_count.  One host process drives all shards.
//
// What is host glue here and OUT OF SCOPE of the GPU path: option parsing,
// FASTA/FASTQ reading, NCBI taxonomy files, text output.  They are written from the formats the
// reference reads and writes, not from its code.
#include <algorithm>
#include <array>
#include <atomic>
#include <fcntl.h>
#include <memory>
#include <sys/mman.h>
#include <unistd.h>
#include <charconv>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include <dirent.h>
#include <unistd.h>
#include <sys/stat.h>

#include "../../include/mcgpu.h"

namespace {

// ------------------------------------------------------------------------------- utilities
[[noreturn]] void die(const std::string& msg)
{
    std::cerr << "metacache_b200: " << msg << std::endl;
    std::fflush(nullptr);
    _exit(1);                 // no static destructors: a background thread may still be running
}

void check(int rc, const char* what)
{
    if(rc != MCG_OK) die(std::string(what) + ": " + mcg_last_error());
}

const char* const RANK_NAMES[22] = {"sequence", "form", "variety", "subspecies", "species", "subgenus", "genus",
                                    "subtribe", "tribe", "subfamily", "family", "suborder", "order", "subclass",
                                    "class", "subphylum", "phylum", "subkingdom", "kingdom", "domain", "root", "none"};

// taxonomy.h:174-214 rank_from_name
int rank_from_name(std::string n)
{
    std::transform(n.begin(), n.end(), n.begin(), ::tolower);
    static const std::map<std::string, int> m = {
        {"sequence", 0}, {"genome", 0}, {"form", 1}, {"forma", 1}, {"variety", 2}, {"varietas", 2}, {"subspecies", 3},
        {"species", 4}, {"species group", 5}, {"species subgroup", 5}, {"subgenus", 5}, {"genus", 6}, {"subtribe", 7},
        {"tribe", 8}, {"subfamily", 9}, {"family", 10}, {"superfamily", 11}, {"parvorder", 11}, {"infraorder", 11},
        {"suborder", 11}, {"order", 12}, {"superorder", 13}, {"infraclass", 13}, {"subclass", 13}, {"class", 14},
        {"superclass", 15}, {"subphylum", 15}, {"phylum", 16}, {"division", 16}, {"superphylum", 17}, {"subkingdom", 17},
        {"kingdom", 18}, {"subdomain", 18}, {"superkingdom", 19}, {"domain", 19}, {"root", 20}};
    auto it = m.find(n);
    return it == m.end() ? 21 : it->second;
}

struct Args {
    std::vector<std::string> pos;                 // positional arguments after the mode
    std::map<std::string, std::string> opt;       // -name [value]
    bool has(std::initializer_list<const char*> names) const
    {
        for(auto n : names) if(opt.count(n)) return true;
        return false;
    }
    std::string get(std::initializer_list<const char*> names, const std::string& def) const
    {
        for(auto n : names) { auto it = opt.find(n); if(it != opt.end() && !it->second.empty()) return it->second; }
        return def;
    }
    long long geti(std::initializer_list<const char*> names, long long def) const
    {
        const std::string s = get(names, "");
        return s.empty() ? def : std::atoll(s.c_str());
    }
};

Args parse_args(int argc, char** argv, int first)
{
    Args a;
    for(int i = first; i < argc; ++i) {
        std::string s = argv[i];
        if(s.size() > 1 && s[0] == '-' && !(s[1] >= '0' && s[1] <= '9')) {
            std::string name = s.substr(1);
            std::string val;
            if(i + 1 < argc && !(argv[i + 1][0] == '-' && !(argv[i + 1][1] >= '0' && argv[i + 1][1] <= '9') && argv[i + 1][1] != 0)) {
                val = argv[i + 1];
                // flags without value must not swallow the next positional argument
                static const std::set<std::string> flags = {
                    "pairfiles", "pair-files", "pairseq", "tophits", "top-hits", "queryids", "query-ids", "taxids",
                    "taxid", "taxids-only", "taxidsonly", "lineage", "mapped-only", "mappedonly", "nomap", "no-map",
                    "no-summary", "nosummary", "no-query-params", "noqueryparams", "remove-overpopulated-features",
                    "silent", "verbose", "omit-ranks", "separate-cols", "reference-compat", "reference_compat", "refcompat"};
                if(flags.count(name)) val.clear(); else ++i;
            }
            a.opt[name] = val;
        } else {
            a.pos.push_back(s);
        }
    }
    return a;
}

std::string db_name(std::string n)            // args_handling.cpp:34-45
{
    if(n.find(".db") == std::string::npos) n += ".db";
    return n;
}

// minimal ostream-like sink over a std::string: the per-read output is formatted by several threads
// into their own buffers and written in order (decimal integers print like operator<< of ostream)
struct StrOut {
    std::string s;
    StrOut& operator<<(const std::string& x) { s += x; return *this; }
    StrOut& operator<<(const char* x) { s += x; return *this; }
    StrOut& operator<<(char x) { s += x; return *this; }
    template<class I, class = typename std::enable_if<std::is_integral<I>::value>::type>
    StrOut& operator<<(I x)
    {
        char b[24];
        auto r = std::to_chars(b, b + sizeof(b), x);
        s.append(b, size_t(r.ptr - b));
        return *this;
    }
};

unsigned host_threads()
{
    unsigned t = std::thread::hardware_concurrency();
    return t < 1 ? 1 : (t > 32 ? 32 : t);
}

template<class T> void wr(std::ostream& os, T x) { os.write(reinterpret_cast<const char*>(&x), sizeof(x)); }
void wr_str(std::ostream& os, const std::string& s) { wr<uint64_t>(os, s.size()); os.write(s.data(), std::streamsize(s.size())); }

// ------------------------------------------------------------------------------- taxonomy
struct Taxon {
    int64_t id = 0, parent = 0;
    uint8_t rank = 21;
    std::string name = "--", file;
    uint64_t index = 0, windows = 0;
};

struct Taxonomy {
    std::map<int64_t, Taxon> taxa;                // ascending id, like the reference's std::set

    // nodes.dmp / names.dmp / merged.dmp: taxonomy_io.cpp:57-180
    void load(const std::string& dir)
    {
        std::map<int64_t, std::string> names;
        {
            std::ifstream is(dir + "/names.dmp");
            std::string line;
            while(std::getline(is, line)) {
                std::vector<std::string> f = split_dmp(line);
                if(f.size() >= 4 && f[3].find("scientific") != std::string::npos) names.emplace(std::atoll(f[0].c_str()), f[1]);
            }
        }
        std::map<int64_t, int64_t> merged;
        {
            std::ifstream is(dir + "/merged.dmp");
            std::string line;
            while(std::getline(is, line)) {
                std::vector<std::string> f = split_dmp(line);
                if(f.size() >= 2) {
                    const int64_t o = std::atoll(f[0].c_str()), n = std::atoll(f[1].c_str());
                    merged[o] = n;
                    Taxon t; t.id = o; t.parent = n; t.name = ""; t.rank = 21;
                    taxa[o] = t;
                }
            }
        }
        std::ifstream is(dir + "/nodes.dmp");
        if(!is.good()) die("could not read taxonomic nodes file " + dir + "/nodes.dmp");
        std::string line;
        while(std::getline(is, line)) {
            std::vector<std::string> f = split_dmp(line);
            if(f.size() < 3) continue;
            int64_t id = std::atoll(f[0].c_str()), par = std::atoll(f[1].c_str());
            auto nm = names.find(id);
            std::string name = nm != names.end() ? nm->second : "--";
            if(name.empty()) name = "<" + std::to_string(id) + ">";
            auto mi = merged.find(id); if(mi != merged.end()) id = mi->second;
            mi = merged.find(par); if(mi != merged.end()) par = mi->second;
            if(id == 0) continue;
            Taxon t; t.id = id; t.parent = par; t.name = name; t.rank = uint8_t(rank_from_name(f[2]));
            if(!taxa.count(id)) taxa[id] = t;      // std::set::emplace keeps the first
        }
        auto r = taxa.find(1);
        if(r != taxa.end()) r->second.rank = 20;   // root
    }

    static std::vector<std::string> split_dmp(const std::string& line)
    {
        std::vector<std::string> f;
        size_t p = 0;
        while(p <= line.size()) {
            size_t q = line.find('|', p);
            if(q == std::string::npos) q = line.size();
            std::string x = line.substr(p, q - p);
            size_t a = x.find_first_not_of(" \t"), b = x.find_last_not_of(" \t");
            f.push_back(a == std::string::npos ? "" : x.substr(a, b - a + 1));
            p = q + 1;
        }
        return f;
    }
};

// flattened tables for mcg_tax_upload: index t < ntargets = target t, then the higher taxa
struct TaxTables {
    std::vector<int64_t> taxid;
    std::vector<const Taxon*> tax;
    std::vector<int32_t> lineage;
    std::vector<uint8_t> rank;
    int64_t ntargets = 0;
    std::unordered_map<int64_t, int32_t> index_of;

    void build(const Taxonomy& t, int64_t ntgt)
    {
        ntargets = ntgt;
        taxid.clear(); tax.clear(); index_of.clear();
        for(int64_t i = 0; i < ntgt; ++i) {
            auto it = t.taxa.find(-i - 1);
            if(it == t.taxa.end()) die("database lacks the sequence-level taxon of target " + std::to_string(i));
            taxid.push_back(-i - 1); tax.push_back(&it->second);
        }
        for(auto& kv : t.taxa) if(kv.first >= 0) { taxid.push_back(kv.first); tax.push_back(&kv.second); }
        for(size_t i = 0; i < taxid.size(); ++i) index_of[taxid[i]] = int32_t(i);
        const size_t n = taxid.size();
        rank.resize(n);
        lineage.assign(n * 21, -1);
        for(size_t i = 0; i < n; ++i) {
            rank[i] = tax[i]->rank;
            // taxonomy.h:576-597 ranks(): walk to the root, ranked taxa fill their slot, later overwrite
            int64_t id = taxid[i];
            size_t guard = 0;
            while(id != 0 && guard++ <= n) {
                auto it = index_of.find(id);
                if(it == index_of.end()) break;
                const Taxon* x = tax[it->second];
                if(x->rank != 21) lineage[i * 21 + x->rank] = it->second;
                if(x->parent == id) break;
                id = x->parent;
            }
        }
    }
};

// ------------------------------------------------------------------------------- sequence files
struct Record { std::string header, data; };

bool read_text_file(const std::string& path, std::string& out)
{
    std::ifstream is(path, std::ios::binary);
    if(!is.good()) return false;
    is.seekg(0, std::ios::end);
    const std::streamoff n = is.tellg();
    is.seekg(0);
    out.resize(size_t(n));
    if(n) is.read(&out[0], n);
    return true;
}

// FASTA ('>') or FASTQ ('@'); sequence_io.cpp:121-170, 253-284.  Zero-length FASTA records and
// everything after them in the file are dropped like in the reference (its reader throws there).
void parse_sequences(const std::string& txt, std::vector<Record>& out)
{
    size_t p = 0;
    const size_t n = txt.size();
    auto line_end = [&](size_t s) { size_t e = txt.find('\n', s); return e == std::string::npos ? n : e; };
    while(p < n && (txt[p] == '\n' || txt[p] == '\r')) ++p;
    if(p >= n) return;
    if(txt[p] == '>') {
        while(p < n && txt[p] == '>') {
            size_t e = line_end(p);
            Record r;
            r.header = txt.substr(p + 1, e - p - 1);
            if(!r.header.empty() && r.header.back() == '\r') r.header.pop_back();
            p = e + 1;
            while(p < n && txt[p] != '>') {
                e = line_end(p);
                size_t len = e - p;
                if(len && txt[p + len - 1] == '\r') --len;
                r.data.append(txt, p, len);
                p = e + 1;
            }
            if(r.data.empty()) return;
            out.push_back(std::move(r));
        }
    } else if(txt[p] == '@') {
        while(p < n && txt[p] == '@') {
            size_t e = line_end(p);
            Record r;
            r.header = txt.substr(p + 1, e - p - 1);
            if(!r.header.empty() && r.header.back() == '\r') r.header.pop_back();
            p = e + 1;
            if(p > n) break;
            e = line_end(p);
            r.data = txt.substr(p, e - p);
            if(!r.data.empty() && r.data.back() == '\r') r.data.pop_back();
            p = e + 1;
            if(p >= n || txt[p] != '+') break;
            p = line_end(p) + 1;               // quality header
            p = line_end(p) + 1;               // qualities
            out.push_back(std::move(r));
        }
    }
}

// query reads: positions inside the file text instead of copies.  FASTQ records are 4 lines
// (sequence_io.cpp:253-284); FASTA sequences may span lines and are then copied into `extra`.
struct ReadRec { size_t hpos, hlen, spos, slen; bool in_extra; };

void index_reads(const std::string& txt, std::vector<ReadRec>& out, std::string& extra)
{
    size_t p = 0;
    const size_t n = txt.size();
    const char* base = txt.data();
    auto line_end = [&](size_t s) { const void* e = std::memchr(base + s, '\n', n - s); return e ? size_t(static_cast<const char*>(e) - base) : n; };
    while(p < n && (txt[p] == '\n' || txt[p] == '\r')) ++p;
    if(p >= n) return;
    if(txt[p] == '@') {
        while(p < n && txt[p] == '@') {
            size_t e = line_end(p);
            ReadRec r;
            r.hpos = p + 1; r.hlen = e - p - 1; r.in_extra = false;
            if(r.hlen && txt[r.hpos + r.hlen - 1] == '\r') --r.hlen;
            p = e + 1;
            if(p > n) break;
            e = line_end(p);
            r.spos = p; r.slen = e - p;
            if(r.slen && txt[r.spos + r.slen - 1] == '\r') --r.slen;
            p = e + 1;
            if(p >= n || txt[p] != '+') break;
            p = line_end(p) + 1;               // quality header
            p = line_end(p) + 1;               // qualities
            out.push_back(r);
        }
    } else if(txt[p] == '>') {
        std::vector<Record> recs;
        parse_sequences(txt, recs);
        size_t total = 0;
        for(auto& r : recs) total += r.header.size() + r.data.size();
        extra.reserve(total);
        for(auto& r : recs) {
            ReadRec x;
            x.in_extra = true;
            x.hpos = extra.size(); x.hlen = r.header.size(); extra += r.header;
            x.spos = extra.size(); x.slen = r.data.size(); extra += r.data;
            out.push_back(x);
        }
    }
}

std::string trimmed(const std::string& s)
{
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? "" : s.substr(a, b - a + 1);
}

// ---- sequence id of a FASTA header.  Behaviour to match (what the reference's database stores as target name,
// observed through `info` dumps and the query output; its extractor lives in sequence_io.cpp:576-716):
//   1. the first of a fixed list of NCBI accession prefixes (in list order, not in text order) that occurs in the
//      header and is followed by a '.' at most 25 characters after its start: prefix ... '.' version;
//   2. otherwise a leading "name.version" whose '.' comes within the first 25 characters;
//   3. otherwise the first listed prefix that occurs at all, without a version;
//   4. otherwise the number after "gi|";
//   an id ends before the first of '|', ' ', '-', '_', ',' -- tried in THIS order -- found after its start.
// Written as a small rule table over string views of the header, not as a port of that code.
struct IdSpan { size_t from = 0, to = 0; bool ok = false; };

size_t id_end(const std::string& t, size_t from)
{
    if(from >= t.size()) return t.size();
    static const char stops[] = {'|', ' ', '-', '_', ','};
    for(char c : stops) { const size_t k = t.find(c, from); if(k != std::string::npos) return k; }
    return t.size();
}

const std::vector<std::string>& accession_prefixes()
{
    static const std::vector<std::string> v = {"GCF_", "AC_", "NC_", "NG_", "NS_", "NT_", "NW_", "NZ_", "MKHE", "AE", "AJ", "AL", "AM",
                                               "AP", "AY", "BA", "BK", "BX", "CC", "CM", "CP", "CR", "CT", "CU", "FM", "FN", "FO", "FP", "FQ", "FR"};
    return v;
}

std::string extract_accession(const std::string& text)
{
    if(text.size() < 2) return "";
    using Rule = IdSpan (*)(const std::string&);
    static const Rule rules[] = {
        [](const std::string& t) {                       // 1. listed prefix with version
            IdSpan r;
            for(const std::string& p : accession_prefixes()) {
                const size_t i = t.find(p);
                if(i == std::string::npos) continue;
                const size_t dot = t.find('.', i + p.size());
                if(dot == std::string::npos || dot - i > 25) continue;
                r.from = i; r.to = id_end(t, dot + 1); r.ok = true;
                break;
            }
            return r;
        },
        [](const std::string& t) {                       // 2. leading name.version
            IdSpan r;
            const size_t dot = t.find('.', 1);
            if(dot < 25) { r.from = 0; r.to = id_end(t, dot + 1); r.ok = true; }
            return r;
        },
        [](const std::string& t) {                       // 3. listed prefix without version
            IdSpan r;
            for(const std::string& p : accession_prefixes()) {
                const size_t i = t.find(p);
                if(i == std::string::npos) continue;
                const size_t body = i + p.size();
                r.from = i; r.to = std::min(id_end(t, body), t.find('.', body)); r.ok = true;
                break;
            }
            return r;
        },
        [](const std::string& t) {                       // 4. gi number
            IdSpan r;
            const size_t i = t.find("gi|");
            if(i == std::string::npos) return r;
            r.from = i + 3;
            r.to = t.find('|', r.from);
            if(r.to == std::string::npos) r.to = t.find(' ', r.from);
            if(r.to == std::string::npos) r.to = t.size();
            r.ok = true;
            return r;
        }};
    for(Rule rule : rules) {
        const IdSpan r = rule(text);
        if(r.ok) return trimmed(text.substr(r.from, r.to - r.from));
    }
    return "";
}

int64_t extract_taxid(const std::string& text)          // sequence_io.cpp:724-748
{
    size_t i = text.find("taxid");
    if(i == std::string::npos) return 0;
    i += 6;
    if(i >= text.size()) return 0;
    size_t j = text.find('|', i);
    if(j == std::string::npos) { j = text.find(' ', i); if(j == std::string::npos) j = text.size(); }
    const std::string num = text.substr(i, j - i);
    if(num.empty() || !(num[0] >= '0' && num[0] <= '9')) return 0;
    return std::atoll(num.c_str());
}

// Sequence -> taxon tables that lie next to the sequence files ("assembly_summary.txt" of an NCBI RefSeq download).
// Behaviour to match (args_handling.cpp:120, taxonomy_io.cpp:190-311, mode_build.cpp:531-604): every directory that holds an
// input file is searched for the table; of its first ten lines the last leading '#' line is the column header (the word
// before the first column name, "#", is not a column); the columns named "taxid" and "assembly_accession" /
// "accession.version" give the taxon and the key, the taxon column counted from the key column on; a table without a
// "taxid" column is read from its first line as key <tab> taxon; the first row of a key wins, a row that does not hold a
// key and a number ends the table.  A sequence is named by the accession of its header, then by the accession in its file
// name; a table key that merely starts with that name counts when no key equals it.
struct SequenceTaxa {
    std::map<std::string, int64_t> by_key;

    void read_table(const std::string& path)
    {
        std::string txt;
        if(!read_text_file(path, txt)) return;
        std::cout << "Reading sequence to taxon mappings from " << path << std::endl;
        std::vector<std::pair<size_t, size_t>> lines;                 // [from, to) without the line feed
        for(size_t i = 0; i < txt.size();) {
            size_t e = txt.find('\n', i);
            if(e == std::string::npos) e = txt.size();
            lines.emplace_back(i, e);
            i = e + 1;
        }
        size_t lead = 0;
        while(lead < lines.size() && lead < 10 && lines[lead].second > lines[lead].first && txt[lines[lead].first] == '#') ++lead;
        const size_t head = lead ? lead - 1 : 0;
        size_t keycol = 0, taxcol = 0, first_row = head + 1;
        if(head < lines.size()) {
            std::istringstream words(txt.substr(lines[head].first, lines[head].second - lines[head].first));
            std::string w;
            words >> w;
            for(size_t col = 0; words >> w; ++col) {
                if(w == "taxid") taxcol = col;
                else if(w == "accession.version" || w == "assembly_accession") keycol = col;
            }
        }
        if(taxcol < 1) { taxcol = 1; first_row = 0; }
        for(size_t l = first_row; l < lines.size(); ++l) {
            const char* b = txt.data() + lines[l].first;
            const char* e = txt.data() + lines[l].second;
            if(std::all_of(b, e, [](char c) { return c == ' ' || c == '\t' || c == '\r'; })) continue;
            std::vector<const char*> field{b};
            for(const char* q = b; q < e; ++q) if(*q == '\t') field.push_back(q + 1);
            field.push_back(e + 1);
            if(keycol + taxcol + 1 >= field.size()) break;
            const char* k0 = field[keycol];
            const char* k1 = field[keycol + 1] - 1;
            while(k0 < k1 && *k0 == ' ') ++k0;
            const char* kend = std::find_if(k0, k1, [](char c) { return c == ' ' || c == '\r'; });
            if(kend == k0) break;
            const char* t0 = field[keycol + taxcol];
            const char* t1 = field[keycol + taxcol + 1] - 1;
            while(t0 < t1 && *t0 == ' ') ++t0;
            const bool neg = t0 < t1 && *t0 == '-';
            const char* dg = t0 + ((t0 < t1 && (*t0 == '-' || *t0 == '+')) ? 1 : 0);
            int64_t v = 0;
            const char* q = dg;
            for(; q < t1 && *q >= '0' && *q <= '9'; ++q) v = v * 10 + (*q - '0');
            by_key.emplace(std::string(k0, kend), q == dg ? 0 : (neg ? -v : v));
            if(q == dg) break;
        }
    }

    void read_tables_next_to(const std::vector<std::string>& files)
    {
        std::set<std::string> dirs;
        for(const auto& f : files) dirs.insert(f.substr(0, f.find_last_of("/\\")));
        for(const auto& d : dirs) read_table(d + "/assembly_summary.txt");
    }

    int64_t taxon_of(const std::string& name) const
    {
        if(by_key.empty() || name.empty()) return 0;
        auto it = by_key.find(name);
        if(it != by_key.end()) return it->second;
        it = by_key.upper_bound(name);
        return it != by_key.end() && it->first.compare(0, name.size(), name) == 0 ? it->second : 0;
    }
};

void list_files(const std::string& path, std::vector<std::string>& out)
{
    struct stat st;
    if(stat(path.c_str(), &st) != 0) return;
    if(S_ISDIR(st.st_mode)) {
        DIR* d = opendir(path.c_str());
        if(!d) return;
        std::vector<std::string> names;
        while(dirent* e = readdir(d)) { std::string n = e->d_name; if(n != "." && n != "..") names.push_back(n); }
        closedir(d);
        for(auto& n : names) list_files(path + "/" + n, out);
    } else {
        out.push_back(path);
    }
}

// Targets whose header and mapping files gave no parent taxon are looked up in NCBI accession -> taxon tables after all
// sequences are in.  Behaviour to match (mode_build.cpp:493-522 with :249-311, args_handling.cpp:121-145,
// sketch_database.h:632-639): files in this order: -taxpostmap <file>, then nucl_gb / nucl_wgs / nucl_est / nucl_gss
// .accession2taxid of the taxonomy directory, then every other *.accession2taxid found there; one header line, then rows
// "accession accession.version taxid gi"; a row names a target by (1) accession.version == target name, (2) the first
// target name greater than the accession that starts with it, (3) gi == target name; the first row that names a still
// unranked target sets its parent.
void rank_unranked_targets(Taxonomy& tax, uint32_t ntargets, std::string taxdir, const std::string& postmap)
{
    std::map<std::string, int64_t> by_name;            // target name -> taxon id -(t)-1
    std::set<int64_t> unranked;
    for(uint32_t t = 0; t < ntargets; ++t) {
        const Taxon& x = tax.taxa[-int64_t(t) - 1];
        by_name.emplace(x.name, x.id);
        if(x.parent == 0) unranked.insert(x.id);
    }
    if(!unranked.empty()) {
        std::cout << unranked.size() << " targets are unranked." << std::endl;
        if(!taxdir.empty() && taxdir.back() != '/') taxdir += '/';
        std::vector<std::string> files;
        if(!postmap.empty()) files.push_back(postmap);
        for(const char* f : {"nucl_gb", "nucl_wgs", "nucl_est", "nucl_gss"}) files.push_back(taxdir + f + ".accession2taxid");
        if(!taxdir.empty()) {
            std::vector<std::string> more;
            list_files(taxdir.substr(0, taxdir.size() - 1), more);
            std::sort(more.begin(), more.end());
            for(const auto& f : more)
                if(f.find(".accession2taxid") != std::string::npos && std::find(files.begin(), files.end(), f) == files.end()) files.push_back(f);
        }
        for(const auto& file : files) {
            std::ifstream is(file);
            if(!is.good()) continue;
            std::cout << "Try to map sequences to taxa using '" << file << "'" << std::endl;
            std::string acc, accver, gi;
            uint64_t taxid = 0;
            std::getline(is, acc);                                // header line
            while(is >> acc >> accver >> taxid >> gi) {
                auto hit = by_name.find(accver);
                if(hit == by_name.end() && !acc.empty()) {
                    auto near = by_name.upper_bound(acc);
                    if(near != by_name.end() && near->first.compare(0, acc.size(), acc) == 0) hit = near;
                    else hit = by_name.find(gi);
                }
                if(hit == by_name.end()) continue;
                auto u = unranked.find(hit->second);
                if(u == unranked.end()) continue;
                tax.taxa[hit->second].parent = int64_t(taxid);
                unranked.erase(u);
                if(unranked.empty()) break;
            }
            if(unranked.empty()) break;
        }
    }
    if(unranked.empty()) std::cout << "All targets are ranked." << std::endl;
    else std::cout << unranked.size() << " targets remain unranked." << std::endl;
}

// ------------------------------------------------------------------------------- shard files
struct DbMeta {
    mcg_sketch_cfg tcfg{16, 16, 128, 113};
    uint64_t qsketch = 16, qwinlen = 128, qstride = 113, maxlocs = 254;
    Taxonomy tax;
    uint32_t ntargets = 0;
};

// device -> host bucket arrays of one shard
struct ShardData { std::vector<uint32_t> keys, sizes, tg, wn; uint64_t ne = 0, nl = 0; };

void export_shard(mcg_db* db, ShardData& d)
{
    uint64_t nk = 0;
    check(mcg_db_counts(db, &nk, &d.ne, &d.nl), "mcg_db_counts");
    d.keys.resize(d.ne ? d.ne : 1); d.sizes.resize(d.ne ? d.ne : 1); d.tg.resize(d.nl ? d.nl : 1); d.wn.resize(d.nl ? d.nl : 1);
    check(mcg_db_export(db, d.keys.data(), d.sizes.data(), d.tg.data(), d.wn.data()), "mcg_db_export");
}

// host only (runs in one thread per shard): the reference's file layout, sketch_database.h:959-999
void write_shard(const std::string& file, const DbMeta& m, const ShardData& d)
{
    const uint64_t ne = d.ne, nl = d.nl;
    std::ofstream os(file, std::ios::binary);
    if(!os.good()) die("can't open file " + file);
    wr<uint64_t>(os, 20181001);                                   // MC_DB_VERSION
    wr<uint8_t>(os, 4); wr<uint8_t>(os, 4); wr<uint8_t>(os, 4); wr<uint8_t>(os, 1); wr<uint8_t>(os, 8); wr<uint8_t>(os, 21);
    wr<uint64_t>(os, m.tcfg.k); wr<uint64_t>(os, m.tcfg.sketchlen); wr<uint64_t>(os, m.tcfg.winlen); wr<uint64_t>(os, m.tcfg.winstride);
    wr<uint64_t>(os, m.tcfg.k); wr<uint64_t>(os, m.qsketch); wr<uint64_t>(os, m.qwinlen); wr<uint64_t>(os, m.qstride);
    wr<uint64_t>(os, m.maxlocs);
    wr<uint64_t>(os, m.tax.taxa.size());
    for(auto& kv : m.tax.taxa) {
        const Taxon& t = kv.second;
        wr<int64_t>(os, t.id); wr<int64_t>(os, t.parent); wr<uint8_t>(os, t.rank);
        wr_str(os, t.name); wr_str(os, t.file); wr<uint64_t>(os, t.index); wr<uint64_t>(os, t.windows);
    }
    wr<uint32_t>(os, m.ntargets);
    wr<uint64_t>(os, ne); wr<uint64_t>(os, nl);
    // per bucket: key u32, n u8, (n u64, n x tgt u32), (n u64, n x win u32)
    // bucket j starts at byte 21 j + 8 start[j] of the bucket stream: the stream is filled by all host threads
    std::vector<uint64_t> start(size_t(ne) + 1, 0);
    for(uint64_t j = 0; j < ne; ++j) start[j + 1] = start[j] + d.sizes[j];
    if(start[ne] != nl) die("internal error: bucket sizes of " + file + " do not add up");
    std::unique_ptr<char[]> buf(new char[size_t(ne) * 21 + size_t(nl) * 8 + 1]);
    const unsigned NT = host_threads();
    {
        std::vector<std::thread> th;
        for(unsigned t = 0; t < NT; ++t) th.emplace_back([&, t] {
            const uint64_t ja = ne * t / NT, jb = ne * (t + 1) / NT;
            char* q = buf.get() + 21 * ja + 8 * start[ja];
            for(uint64_t j = ja; j < jb; ++j) {
                const uint32_t key = d.keys[j];
                const uint64_t n = d.sizes[j], lo = start[j];
                const uint8_t n8 = uint8_t(n);
                std::memcpy(q, &key, 4); q += 4;
                *q++ = char(n8);
                std::memcpy(q, &n, 8); q += 8;
                std::memcpy(q, &d.tg[lo], size_t(n) * 4); q += n * 4;
                std::memcpy(q, &n, 8); q += 8;
                std::memcpy(q, &d.wn[lo], size_t(n) * 4); q += n * 4;
            }
        });
        for(auto& x : th) x.join();
    }
    os.write(buf.get(), std::streamsize(size_t(ne) * 21 + size_t(nl) * 8));
    if(!os.good()) die("could not write " + file);
}

// returns the bucket arrays of a shard file written by this program or by the reference
void read_shard(const std::string& file, DbMeta& m, bool meta_only,
                std::vector<uint32_t>& keys, std::vector<uint32_t>& sizes, std::vector<uint32_t>& tg, std::vector<uint32_t>& wn)
{
    // the file is mapped, not copied; the bucket stream is indexed by one thread and copied out by all
    const int fd = ::open(file.c_str(), O_RDONLY);
    if(fd < 0) die("can't open file " + file);
    struct stat fst;
    if(::fstat(fd, &fst) != 0 || fst.st_size <= 0) die("can't open file " + file);
    const size_t fsize = size_t(fst.st_size);
    void* mapping = ::mmap(nullptr, fsize, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);
    if(mapping == MAP_FAILED) die("can't map file " + file);
    struct Unmap { void* p; size_t n; ~Unmap() { ::munmap(p, n); } } unmap{mapping, fsize};
    ::madvise(mapping, fsize, MADV_WILLNEED);
    const char* p = static_cast<const char*>(mapping);
    const char* const end = p + fsize;
    auto need = [&](size_t n) { if(size_t(end - p) < n) die("database " + file + " is truncated or corrupt"); };
    auto get = [&](auto& x) { need(sizeof(x)); std::memcpy(&x, p, sizeof(x)); p += sizeof(x); };
    auto get_str = [&](std::string& x) { uint64_t n = 0; get(n); need(size_t(n)); x.assign(p, size_t(n)); p += n; };
    uint64_t ver = 0; get(ver);
    if(ver != 20181001) die("database " + file + " has incompatible version " + std::to_string(ver));
    uint8_t w[6]; for(auto& x : w) get(x);
    if(w[0] != 4 || w[1] != 4 || w[2] != 4 || w[3] != 1 || w[4] != 8 || w[5] != 21) die("database " + file + " uses different data type sizes");
    uint64_t k, sk, wl, ws, qk;
    get(k); get(sk); get(wl); get(ws);
    m.tcfg.k = int32_t(k); m.tcfg.sketchlen = int32_t(sk); m.tcfg.winlen = int64_t(wl); m.tcfg.winstride = int64_t(ws);
    get(qk); get(m.qsketch); get(m.qwinlen); get(m.qstride);
    get(m.maxlocs);
    uint64_t nt = 0; get(nt);
    m.tax.taxa.clear();
    for(uint64_t i = 0; i < nt; ++i) {
        Taxon t;
        get(t.id); get(t.parent); get(t.rank); get_str(t.name); get_str(t.file); get(t.index); get(t.windows);
        m.tax.taxa[t.id] = t;
    }
    get(m.ntargets);
    if(meta_only) return;
    uint64_t nk = 0, nv = 0; get(nk); get(nv);
    keys.resize(nk); sizes.resize(nk); tg.resize(nv); wn.resize(nv);
    // pass 1 (one thread): only the size byte of every bucket is read; every CHUNK buckets the byte position
    // and the location count so far are noted.  Pass 2 (all threads): chunks are parsed independently.
    const uint64_t CHUNK = 1u << 16;
    const uint64_t nchunks = (nk + CHUNK - 1) / CHUNK;
    std::vector<const char*> cpos(size_t(nchunks) + 1);
    std::vector<uint64_t> cloc(size_t(nchunks) + 1);
    {
        uint64_t lo = 0;
        for(uint64_t j = 0; j < nk; ++j) {
            if(j % CHUNK == 0) { cpos[j / CHUNK] = p; cloc[j / CHUNK] = lo; }
            need(5);
            const uint8_t n8 = uint8_t(p[4]);
            p += 5;
            if(n8 > 0) {
                // both counts repeat the size byte (hash_multimap.h:975-1029); checked in pass 2
                const size_t len = 16 + size_t(n8) * 8;
                need(len);
                p += len;
                lo += n8;
            }
        }
        cpos[nchunks] = p; cloc[nchunks] = lo;
        if(lo != nv) die("database " + file + " is truncated or corrupt");
    }
    {
        const unsigned NT = host_threads();
        std::atomic<bool> bad{false};
        std::vector<std::thread> th;
        for(unsigned t = 0; t < NT; ++t) th.emplace_back([&, t] {
            for(uint64_t c = t; c < nchunks; c += NT) {
                const char* q = cpos[c];
                uint64_t lo = cloc[c];
                const uint64_t jb = std::min(nk, (c + 1) * CHUNK);
                for(uint64_t j = c * CHUNK; j < jb; ++j) {
                    std::memcpy(&keys[j], q, 4);
                    const uint8_t n8 = uint8_t(q[4]);
                    q += 5;
                    sizes[j] = n8;
                    if(n8 > 0) {
                        uint64_t n1 = 0, n2 = 0;
                        std::memcpy(&n1, q, 8); q += 8;
                        std::memcpy(&tg[lo], q, size_t(n8) * 4); q += size_t(n8) * 4;
                        std::memcpy(&n2, q, 8); q += 8;
                        std::memcpy(&wn[lo], q, size_t(n8) * 4); q += size_t(n8) * 4;
                        if(n1 != n8 || n2 != n8) bad = true;
                        lo += n8;
                    }
                }
                if(q != cpos[c + 1] || lo != cloc[c + 1]) bad = true;
            }
        });
        for(auto& x : th) x.join();
        if(bad) die("database " + file + " is truncated or corrupt");
    }
}

std::vector<mcg_ctx*> g_ctx;        // one per device

mcg_ctx* ctx_for_shard(int shard)
{
    const int ndev = mcg_device_count();
    if(ndev < 1) die("no CUDA device available; there is no CPU fallback");
    if(g_ctx.empty()) g_ctx.assign(size_t(ndev), nullptr);
    const int d = shard % ndev;
    if(!g_ctx[d]) check(mcg_ctx_create(d, nullptr, &g_ctx[d]), "mcg_ctx_create");
    return g_ctx[d];
}

// CUDA start-up (driver initialisation + context creation: 1-2 s on an 8-GPU box) runs in the
// background while the main thread reads and parses files
std::thread g_gpu_init;
void start_gpu_init(int nshards)
{
    g_gpu_init = std::thread([nshards] {
        const int ndev = mcg_device_count();                      // driver initialisation
        if(ndev < 1) die("no CUDA device available; there is no CPU fallback");
        g_ctx.assign(size_t(ndev), nullptr);
        std::vector<std::thread> th;                              // one context per GPU in use, created side by side
        for(int d = 0; d < ndev && d < nshards; ++d)
            th.emplace_back([d] { check(mcg_ctx_create(d, nullptr, &g_ctx[size_t(d)]), "mcg_ctx_create"); });
        for(auto& t : th) t.join();
    });
}
void wait_gpu_init() { if(g_gpu_init.joinable()) g_gpu_init.join(); }

// f(shard) for every shard: one thread per GPU, the shards of a GPU one after the other (a context
// is not re-entrant)
template<class F>
void for_each_shard_parallel(int P, F f)
{
    const int ndev = std::max(1, mcg_device_count());
    if(ndev == 1 || P == 1) { for(int r = 0; r < P; ++r) f(r); return; }
    std::vector<std::thread> th;
    for(int d = 0; d < ndev && d < P; ++d) th.emplace_back([=] { for(int r = d; r < P; r += ndev) f(r); });
    for(auto& t : th) t.join();
}

// one NCCL rank per shard / GPU inside this process: the id is drawn once, the ranks join side by side
std::vector<mcg_comm*> make_comms(int P)
{
    std::vector<mcg_comm*> comms(static_cast<size_t>(P), nullptr);
    uint8_t id[MCG_COMM_ID_BYTES];
    check(mcg_comm_get_id(id), "mcg_comm_get_id");
    std::vector<std::thread> th;
    for(int r = 0; r < P; ++r)
        th.emplace_back([&, r] { check(mcg_comm_create(ctx_for_shard(r), id, r, P, &comms[size_t(r)]), "mcg_comm_create"); });
    for(auto& t : th) t.join();
    return comms;
}

// page-locked buffer (mcg_host_alloc) that grows and is freed with the process
template<class T>
struct PinnedBuf {
    T* p = nullptr;
    size_t cap = 0;
    void reserve(size_t n)
    {
        if(n <= cap) return;
        if(p) mcg_host_free(p);
        cap = n + n / 4 + 64;
        void* q = nullptr;
        check(mcg_host_alloc(int64_t(cap * sizeof(T)), &q), "mcg_host_alloc");
        p = static_cast<T*>(q);
    }
    T* data() { return p; }
    T& operator[](size_t i) { return p[i]; }
};

double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// phase timer on stderr (MCG_TRACE=1)
struct Lap {
    bool on = getenv("MCG_TRACE") != nullptr;
    double t = now();
    void operator()(const char* what) { if(on) { const double n = now(); std::fprintf(stderr, "[metacache_b200] %-28s %8.1f ms\n", what, 1e3 * (n - t)); t = n; } }
};

// ------------------------------------------------------------------------------- build
int mode_build(const Args& a)
{
    if(a.pos.size() < 2) die("usage: metacache_b200 build <database> <sequence files / directories ...> [options]");
    Lap lap;
    const std::string dbfile = db_name(a.pos[0]);
    int P = int(a.geti({"gpus", "np"}, 0));
    if(P < 1) P = std::max(1, mcg_device_count());          // one shard per GPU unless told otherwise
    start_gpu_init(P);
    DbMeta m;
    m.tcfg.k = int32_t(a.geti({"kmerlen"}, 16));
    m.tcfg.sketchlen = int32_t(a.geti({"sketchlen"}, 16));
    m.tcfg.winlen = a.geti({"winlen"}, 128);
    m.tcfg.winstride = a.geti({"winstride"}, m.tcfg.winlen - m.tcfg.k + 1);       // mode_build.cpp:108-112
    if(m.tcfg.winstride < 1) m.tcfg.winstride = 1;
    m.qsketch = uint64_t(m.tcfg.sketchlen); m.qwinlen = uint64_t(m.tcfg.winlen); m.qstride = uint64_t(m.tcfg.winstride);
    int maxlocs_arg = int(a.geti({"max-locations-per-feature", "max_locations_per_feature"}, 254));
    {   // mode_build.cpp:659 narrows to the uint8 bucket size type
        int v = int(uint8_t(maxlocs_arg));
        if(v < 1) v = 1;
        if(v > 254) v = 254;
        m.maxlocs = uint64_t(v);
    }
    const bool remove_over = a.has({"remove-overpopulated-features"});
    const std::string taxdir = a.get({"taxonomy"}, "");
    if(!taxdir.empty()) m.tax.load(taxdir);

    std::vector<std::string> files;
    for(size_t i = 1; i < a.pos.size(); ++i) list_files(a.pos[i], files);
    std::sort(files.begin(), files.end());                                     // mode_build.cpp:575
    if(files.empty()) die("no sequence files found");

    lap("taxonomy + file list");
    std::vector<mcg_db*> dbs(static_cast<size_t>(P), nullptr);
    auto ensure_dbs = [&]() {          // first use of the GPUs: their start-up ran beside the file parsing
        if(dbs[0] != nullptr) return;
        wait_gpu_init();
        for(int r = 0; r < P; ++r) check(mcg_db_create(ctx_for_shard(r), &m.tcfg, maxlocs_arg, r, P, &dbs[size_t(r)]), "mcg_db_create");
    };

    const double t0 = now();
    // per-shard batches of owned targets
    struct Batch { std::string bytes; std::vector<int64_t> offs{0}; std::vector<uint32_t> ids; };
    std::vector<Batch> batch(static_cast<size_t>(P));
    std::vector<uint64_t> target_windows;
    auto flush = [&](int r) {
        Batch& b = batch[size_t(r)];
        if(b.ids.empty()) return;
        ensure_dbs();
        std::vector<uint32_t> wins(b.ids.size());
        check(mcg_db_add_targets(dbs[size_t(r)], b.bytes.data(), b.offs.data(), b.ids.data(), int64_t(b.ids.size()), wins.data()),
              "mcg_db_add_targets");
        for(size_t i = 0; i < b.ids.size(); ++i) target_windows[b.ids[i]] = wins[i];
        b.bytes.clear(); b.offs.assign(1, 0); b.ids.clear();
    };
    std::set<std::string> names;
    uint32_t ntargets = 0;
    uint64_t total_bp = 0;
    std::string txt;
    SequenceTaxa mapped;
    mapped.read_tables_next_to(files);
    for(const auto& f : files) {
        if(!read_text_file(f, txt)) continue;
        std::vector<Record> recs;
        parse_sequences(txt, recs);
        const std::string file_id = extract_accession(f);
        uint64_t index = 0;
        for(auto& rec : recs) {
            ++index;
            if(rec.data.empty()) continue;
            std::string sid = extract_accession(rec.header);
            if(sid.empty()) sid = rec.header;
            if(names.count(sid)) continue;                                      // sketch_database.h:534
            names.insert(sid);
            int64_t parent = mapped.taxon_of(sid);                              // mode_build.cpp:598-604
            if(parent == 0) parent = mapped.taxon_of(file_id);
            if(parent == 0) parent = extract_taxid(rec.header);
            if(parent < 1) parent = 0;
            const uint32_t tgt = ntargets++;
            const int r = int(tgt % uint32_t(P));
            Taxon t;
            t.id = -int64_t(tgt) - 1; t.parent = parent; t.rank = 0; t.name = sid; t.file = f; t.index = index; t.windows = 0;
            m.tax.taxa[t.id] = t;
            target_windows.push_back(0);
            Batch& b = batch[size_t(r)];
            b.bytes += rec.data;
            b.offs.push_back(int64_t(b.bytes.size()));
            b.ids.push_back(tgt);
            total_bp += rec.data.size();
            if(b.bytes.size() > (size_t(512) << 20)) flush(r);
        }
    }
    lap("parse sequence files");
    rank_unranked_targets(m.tax, ntargets, taxdir, a.get({"taxpostmap", "taxonomy-postmap"}, ""));
    lap("rank unranked targets");
    ensure_dbs();
    lap("wait for CUDA start-up");
    for_each_shard_parallel(P, [&](int r) { flush(r); });
    lap("add_targets");
    m.ntargets = ntargets;
    for_each_shard_parallel(P, [&](int r) { check(mcg_db_finalize(dbs[size_t(r)]), "mcg_db_finalize"); });
    lap("finalize");
    const double t1 = now();
    std::cout << "Added " << ntargets << " reference sequences (" << total_bp << " bp) in " << (t1 - t0) << " s" << std::endl;
    if(remove_over) {
        uint64_t rem = 0;
        if(P > 1 && P <= mcg_device_count()) {
            // one shard per GPU: the reduction runs on the devices (pairs to their key owner over NVLink, removal list back)
            std::vector<mcg_comm*> comms = make_comms(P);
            std::vector<uint64_t> part(static_cast<size_t>(P), 0);
            for_each_shard_parallel(P, [&](int r) {
                check(mcg_db_remove_overpopulated_dist(dbs[size_t(r)], comms[size_t(r)], &part[size_t(r)], nullptr), "mcg_db_remove_overpopulated_dist");
            });
            for(uint64_t x : part) rem += x;
            for(mcg_comm* m : comms) mcg_comm_destroy(m);
        } else {
            check(mcg_remove_overpopulated(dbs.data(), P, &rem), "mcg_remove_overpopulated");
        }
        std::cout << rem << " overpopulated feature lists removed" << std::endl;
    }
    {
        // device -> host per shard, then one writer thread per shard
        std::vector<ShardData> data(static_cast<size_t>(P));
        std::vector<DbMeta> metas(static_cast<size_t>(P), m);
        std::vector<std::thread> writers;
        for(int r = 0; r < P; ++r) {
            // source.windows is only known on the owning rank (sketch_database.h:541), 0 elsewhere
            DbMeta& mr = metas[size_t(r)];
            for(uint32_t t = 0; t < ntargets; ++t) mr.tax.taxa[-int64_t(t) - 1].windows = (int(t % uint32_t(P)) == r) ? target_windows[t] : 0;
            export_shard(dbs[size_t(r)], data[size_t(r)]);
            const std::string out = dbfile + "_" + std::to_string(r);
            uint64_t nk = 0, ne = 0, nl = 0;
            mcg_db_counts(dbs[size_t(r)], &nk, &ne, &nl);
            std::cout << "shard " << r << ": features " << ne << " dead " << (nk - ne) << " locations " << nl << " -> " << out << std::endl;
            mcg_db_destroy(dbs[size_t(r)]);
            writers.emplace_back([&metas, &data, out, r] { write_shard(out, metas[size_t(r)], data[size_t(r)]); });
        }
        for(auto& w : writers) w.join();
    }
    lap("write shards");
    std::cout << "Total build time: " << (now() - t0) << " s" << std::endl;
    return 0;
}

// ------------------------------------------------------------------------------- query
struct OutOpt {
    bool queryids = false, tophits = false, lineage = false, show_ranks = true, summary = true, params = true;
    int mode = 0;               // 0 rank_name, 1 rank_id (taxids-only), 2 rank_name_id (taxids)
    int map = 0;                // 0 all, 1 mapped only, 2 none
    int lowest = 0, highest = 19;
    std::string comment = "# ", none = "--", col = "\t|\t", taxsep = ",", ranksuf = ":", idpre = "(", idsuf = ")";
};

template<class OS>
void show_one_taxon(OS& os, const OutOpt& o, const Taxon* t)
{
    if(o.show_ranks) os << RANK_NAMES[t->rank] << o.ranksuf;
    if(o.mode == 1) os << t->id;
    else if(o.mode == 2) os << t->name << o.idpre << t->id << o.idsuf;
    else os << t->name;
}

template<class OS>
void show_no_taxon(OS& os, const OutOpt& o, int r)
{
    if(o.show_ranks) os << RANK_NAMES[r] << o.ranksuf;
    if(o.mode == 1) os << 0;
    else if(o.mode == 2) os << o.none << o.idpre << 0 << o.idsuf;
    else os << o.none;
}

// printing.cpp:305-330 show_taxon (collapseUnclassified = true)
template<class OS>
void show_classification(OS& os, const OutOpt& o, const TaxTables& tt, int32_t best)
{
    if(best < 0 || tt.rank[size_t(best)] > o.highest) {
        if(o.mode == 1) os << 0; else os << o.none;
        return;
    }
    const int rmin = o.lowest < tt.rank[size_t(best)] ? tt.rank[size_t(best)] : o.lowest;
    const int rmax = o.lineage ? o.highest : rmin;
    for(int r = rmin; r <= rmax; ++r) {
        const int32_t x = tt.lineage[size_t(best) * 21 + size_t(r)];
        if(x >= 0) show_one_taxon(os, o, tt.tax[size_t(x)]); else show_no_taxon(os, o, r);
        if(r < rmax) os << o.taxsep;
    }
}

// ---- per-taxon read counts and abundance estimates (classification.h:107-115, classification.cpp:362-428,
// printing.cpp:474-516).  Counts are floats in the reference; its arithmetic is kept expression by expression.
struct TaxaByRank {                                   // root first, then by id (classification.h:107-114)
    bool operator()(const Taxon* l, const Taxon* r) const
    {
        if(l->rank > r->rank) return true;
        if(l->rank < r->rank) return false;
        return l->id < r->id;
    }
};
using TaxCountMap = std::map<const Taxon*, float, TaxaByRank>;

// Abundance estimate at `rank`.  Behaviour to match (the table the reference prints for -abundance-per; its code is
// classification.cpp:362-428; pinned by tests/test_host_driver_cpu.py against tables the reference printed):
//   a. counts of taxa below the estimation rank move to their closest ancestor at or above that rank;
//   b. among the taxa that hold counts, every taxon hangs below its closest counted ancestor; a taxon's weight is the
//      total count below it;
//   c. top down, a counted taxon with counted descendants hands its own count to its children in proportion to
//      (child count + child weight) / own weight, and leaves the table.
// Counts are single-precision and weights 64-bit integers in the reference, and rows are visited in table order
// (higher ranks first, then taxon id): the same types and visiting order are kept here, because they decide the
// rounding of the printed numbers.  Implemented over one flat, ordered node array with parent / child indices.
void estimate_abundance(const TaxTables& tt, TaxCountMap& all, int rank)
{
    auto ancestor_at = [&](const Taxon* t, unsigned r) -> const Taxon* {
        const int32_t x = tt.lineage[size_t(tt.index_of.at(t->id)) * 21 + r];
        return x >= 0 ? tt.tax[size_t(x)] : nullptr;
    };
    // (a) fold what lies below the estimation rank, in table order
    if(rank > 0) {
        std::vector<const Taxon*> low;
        for(const auto& kv : all) if(int(kv.first->rank) < rank) low.push_back(kv.first);
        for(const Taxon* t : low) {
            const Taxon* up = nullptr;
            for(unsigned r = unsigned(rank); r < 21 && up == nullptr; ++r) up = ancestor_at(t, r);
            if(up == nullptr) continue;                      // nothing above it at that rank: the row stays
            const float moved = all[t];
            all[up] += moved;
            all.erase(t);
        }
    }
    // (b) flat node array in table order; parent = closest counted ancestor
    struct Node { const Taxon* tax; float count; uint64_t weight; int parent; std::vector<int> kids; bool gone; };
    std::vector<Node> nodes;
    nodes.reserve(all.size());
    std::unordered_map<const Taxon*, int> where;
    for(const auto& kv : all) { where[kv.first] = int(nodes.size()); nodes.push_back(Node{kv.first, kv.second, 0, -1, {}, false}); }
    for(int i = int(nodes.size()) - 1; i >= 0; --i) {         // leaves first: a node's weight is final before it is handed up
        Node& n = nodes[size_t(i)];
        for(unsigned r = unsigned(n.tax->rank) + 1; r < 21; ++r) {
            const Taxon* up = ancestor_at(n.tax, r);
            if(up == nullptr) continue;
            auto it = where.find(up);
            if(it == where.end()) continue;
            Node& par = nodes[size_t(it->second)];
            n.parent = it->second;
            par.weight += n.weight + n.count;                // integer += (integer + float): evaluated in float, stored as integer
            par.kids.push_back(i);
            break;
        }
    }
    // (c) top down
    for(Node& n : nodes) {
        if(n.kids.empty()) continue;
        const uint64_t below = n.weight;
        for(int k : n.kids) {
            Node& ch = nodes[size_t(k)];
            ch.count += n.count * (ch.count + ch.weight) / below;
        }
        n.gone = true;
    }
    all.clear();
    for(const Node& n : nodes) if(!n.gone) all[n.tax] = n.count;
}

void show_abundance_table(std::ostream& os, const TaxCountMap& all, uint64_t total, const std::string& ranksuf, const std::string& col)
{
    for(const auto& tc : all)
        os << RANK_NAMES[tc.first->rank] << ranksuf << tc.first->name << col << tc.second << col
           << (tc.second / double(total) * 100) << "%\n";
}

int mode_query(const Args& a)
{
    if(a.pos.size() < 2) die("usage: metacache_b200 query <database> <reads1> [<reads2> -pairfiles] [options]");
    const std::string dbfile = db_name(a.pos[0]);
    int P = int(a.geti({"gpus", "np"}, 0));
    if(P < 1) { P = 0; struct stat st; while(stat((dbfile + "_" + std::to_string(P)).c_str(), &st) == 0) ++P; }
    if(P < 1) die("no shard files " + dbfile + "_<r> found");
    const bool paired = a.has({"pairfiles", "pair-files"}) && a.pos.size() >= 3;

    // shards: CUDA starts up in the background while one thread per shard reads and parses its file
    Lap lap;
    start_gpu_init(P);
    std::vector<DbMeta> metas(static_cast<size_t>(P));
    std::vector<mcg_db*> dbs(static_cast<size_t>(P), nullptr);
    {
        std::vector<ShardData> data(static_cast<size_t>(P));
        std::vector<std::thread> readers;
        for(int r = 0; r < P; ++r)
            readers.emplace_back([&, r] { ShardData& d = data[size_t(r)]; read_shard(dbfile + "_" + std::to_string(r), metas[size_t(r)], false, d.keys, d.sizes, d.tg, d.wn); });
        for(auto& t : readers) t.join();
        lap("read shard files");
        wait_gpu_init();
        lap("wait for CUDA start-up");
        for(int r = 0; r < P; ++r) {
            const DbMeta& m = metas[size_t(r)];
            const ShardData& d = data[size_t(r)];
            check(mcg_db_create(ctx_for_shard(r), &m.tcfg, int(m.maxlocs), r, P, &dbs[size_t(r)]), "mcg_db_create");
            check(mcg_db_import(dbs[size_t(r)], d.keys.data(), d.sizes.data(), d.tg.data(), d.wn.data(), d.keys.size(), d.tg.size()), "mcg_db_import");
        }
        lap("import shards to device");
    }
    const DbMeta& m0 = metas[0];
    TaxTables tt;
    tt.build(m0.tax, m0.ntargets);
    for(mcg_ctx* c : g_ctx) if(c) check(mcg_tax_upload(c, tt.lineage.data(), tt.rank.data(), int64_t(tt.rank.size()), tt.ntargets), "mcg_tax_upload");

    // options (query_options.cpp)
    mcg_query_opts q{};
    q.sketchlen = int32_t(a.geti({"sketchlen"}, int64_t(m0.qsketch)));
    q.winlen = a.geti({"winlen"}, int64_t(m0.qwinlen));
    q.winstride = a.has({"winstride"}) ? a.geti({"winstride"}, 0) : (a.has({"winlen"}) ? q.winlen : int64_t(m0.qstride));   // quirk Q9
    q.lowest = 0; q.highest = 19;
    { const std::string s = a.get({"lowest"}, ""); if(!s.empty()) { int r = rank_from_name(s); if(r < 21) q.lowest = r; } }
    { const std::string s = a.get({"highest"}, ""); if(!s.empty()) { int r = rank_from_name(s); if(r <= 20) q.highest = r; } }
    if(q.lowest > q.highest) q.lowest = q.highest;
    q.insertsize = a.geti({"insertsize", "insert-size"}, 0);
    q.hitmin = int32_t(a.geti({"hitmin"}, 0));
    {
        float f = float(std::atof(a.get({"hitdiff"}, "1").c_str()));
        if(f > 1) f = float(double(f) * 0.01);                                  // query_options.cpp:167-169
        q.hitdiff = f;
    }
    q.maxcand = int32_t(a.geti({"max-cand", "maxcand", "max_cand", "max-candidates", "maxcandidates"}, 2));
    const int maxcand_shown = q.maxcand;
    if(q.maxcand < 1 || q.maxcand > MCG_MAX_CAND) {
        // query_options.cpp:176-178: < 1 means "unlimited"; the device keeps MCG_MAX_CAND (one list entry per lane)
        std::cerr << "metacache_b200: -maxcand " << q.maxcand << (q.maxcand < 1 ? " (unlimited)" : "") << " is limited to "
                  << MCG_MAX_CAND << " candidates per query on the device\n";
        q.maxcand = MCG_MAX_CAND;
    }
    // -reference-compat: the cross-shard merge drops what the reference's uint32 wire format loses (quirk Q1,
    // querying.h:958,984,1001): with the default -lowest sequence only shard 0's candidates survive
    q.reference_compat = a.has({"reference-compat", "reference_compat", "refcompat"}) ? 1 : 0;
    if(q.hitmin < 1) {                                                          // mode_query.cpp:247-260
        const int sks = m0.tcfg.sketchlen;
        q.hitmin = sks >= 6 ? int(sks / 3.0) : (sks >= 4 ? 2 : 1);
    }
    OutOpt o;
    o.queryids = a.has({"queryids", "query-ids", "queryid", "query-id"});
    o.tophits = a.has({"tophits", "top-hits"});
    o.lineage = a.has({"lineage"});
    o.show_ranks = !a.has({"omit-ranks", "omitranks"});
    if(a.has({"taxidsonly", "taxids-only", "taxidonly", "taxid-only"})) o.mode = 1;
    else if(a.has({"taxids", "taxid"})) o.mode = 2;
    if(a.has({"nomap", "no-map"})) o.map = 2; else if(a.has({"mapped-only", "mappedonly"})) o.map = 1;
    if(o.map == 2 && o.tophits) o.map = 1;
    o.summary = !a.has({"no-summary", "nosummary"});
    o.params = !a.has({"no-query-params", "noqueryparams"});
    o.lowest = q.lowest; o.highest = q.highest;
    const bool show_abund = a.has({"abundances"});                              // query_options.cpp:310-320
    std::string abund_name = a.get({"abundances"}, "");
    int est_rank = 21;
    {
        const std::string s = a.get({"abundance-per", "abundances-per", "abundance_per", "abundances_per"}, "");
        if(!s.empty()) { const int r = rank_from_name(s); if(r < 20) est_rank = r; }
    }
    const bool tax_counts = show_abund || est_rank != 21;
    // -hits-per-seq: the per-target match lists are dead code in this fork (classification.cpp:694-699 is commented
    // out): the reference prints the table header only, and so does this driver
    const std::initializer_list<const char*> HPS = {"hits-per-seq", "hitsperseq", "hits_per_seq", "hits-per-sequence",
                                                    "hitspersequence", "hits_per_sequence"};
    const bool hits_per_seq = a.has(HPS);
    std::string targets_name = a.get(HPS, "");
    if(hits_per_seq) o.queryids = true;                                         // query_options.cpp:308

    std::ofstream fout;
    const std::string outname = a.get({"out"}, "");
    if(!outname.empty()) {
        fout.open(outname);
        if(!fout.good()) die("Could not write to file " + outname);
        std::cout << "Per-Read mappings will be written to file: " << outname << std::endl;
    }
    std::ostream& os = outname.empty() ? std::cout : fout;
    if(targets_name == outname) targets_name.clear();                           // query_options.cpp:353
    std::ofstream ftargets;
    if(!targets_name.empty()) {                                                 // mode_query.cpp:91-102
        ftargets.open(targets_name);
        if(!ftargets.good()) die("Could not write to file " + targets_name);
        std::cout << "Per-Target mappings will be written to file: " << targets_name << std::endl;
    }
    std::ostream& os_tgt = targets_name.empty() ? os : ftargets;
    if(abund_name == outname) abund_name.clear();                               // query_options.cpp:355
    std::ofstream fabund;
    if(!abund_name.empty()) {                                                   // mode_query.cpp:104-115
        fabund.open(abund_name);
        if(!fabund.good()) die("Could not write to file " + abund_name);
        std::cout << "Per-Taxon mappings will be written to file: " << abund_name << std::endl;
    }
    std::ostream& os_tax = abund_name.empty() ? os : fabund;
    const std::string& cm = o.comment;
    if(o.params) {                                                              // printing.cpp:40-112
        if(o.map != 2) {
            os << cm << "Reporting per-read mappings (non-mapping lines start with '" << cm << "').\n";
            os << cm << (o.lineage ? "The complete lineage will be reported starting with the lowest match.\n"
                                   : "Only the lowest matching rank will be reported.\n");
        } else {
            os << cm << "Per-Read mappings will not be shown.\n";
        }
        os << cm << "Classification will be constrained to ranks from '" << RANK_NAMES[q.lowest] << "' to '" << RANK_NAMES[q.highest] << "'.\n";
        os << cm << "Classification hit threshold is " << q.hitmin << " per query\n";
        os << cm << "At maximum ";
        if(maxcand_shown < 1) os << std::numeric_limits<std::size_t>::max(); else os << maxcand_shown;      // query_options.cpp:176-178
        os << " classification candidates will be considered per query.\n";
        if(paired) {
            os << cm << "File based paired-end mode:\n" << cm << "  Reads from two consecutive files will be interleaved.\n"
               << cm << "  Max insert size considered " << q.insertsize << ".\n";
        }
        if(hits_per_seq)                                                        // printing.cpp:95-103 (both lines hang on this switch)
            os << cm << "A list of hits per reference sequence will be generated after the read mapping.\n"
               << cm << "A list of absolute and relative abundances per taxon will be generated after the read mapping.\n";
        if(est_rank != 21)
            os << cm << "A list of absolute and relative abundances for each '" << RANK_NAMES[est_rank]
               << "' will be generated after the read mapping.\n";
        os << cm << "Using " << a.geti({"threads"}, 1) << " threads\n";
    }
    if(o.map != 2) {                                                            // classification.cpp:486-511
        os << cm << "TABLE_LAYOUT: ";
        if(o.queryids) os << "query_id" << o.col;
        os << "query_header" << o.col;
        if(o.tophits) os << "top_hits" << o.col;
        const int rmax = o.lineage ? o.highest : o.lowest;
        for(int r = o.lowest; r <= rmax; ++r) {
            if(o.show_ranks) os << (o.lowest == rmax ? "rank" : RANK_NAMES[r]) << o.ranksuf;
            if(o.mode == 1) os << "taxid"; else if(o.mode == 2) os << "taxname" << o.idpre << "taxid" << o.idsuf; else os << "taxname";
            if(r < rmax) os << o.taxsep;
        }
        os << '\n';
    }
    os << cm << a.pos[1];
    if(paired) os << " + " << a.pos[2];
    os << '\n';

    // reads
    const double t0 = now();
    std::string txt1, txt2, extra1, extra2;
    std::vector<ReadRec> r1, r2;
    {
        // the two read files are read and indexed by two threads
        std::thread th2;
        if(paired) th2 = std::thread([&] { if(!read_text_file(a.pos[2], txt2)) die("can't open file " + a.pos[2]); index_reads(txt2, r2, extra2); });
        if(!read_text_file(a.pos[1], txt1)) die("can't open file " + a.pos[1]);
        index_reads(txt1, r1, extra1);
        if(th2.joinable()) th2.join();
    }
    auto seq_of = [](const ReadRec& r, const std::string& txt, const std::string& extra) { return (r.in_extra ? extra.data() : txt.data()) + r.spos; };
    const size_t nreads = paired ? std::min(r1.size(), r2.size()) : r1.size();
    lap("read + index read files");
    const int K = q.maxcand;
    uint64_t stats[22] = {0};
    std::vector<uint64_t> tax_reads(tax_counts ? tt.tax.size() : 0, 0);      // reads per assigned taxon
    const size_t BATCH = size_t(1) << 20;
    const unsigned NT = host_threads();
    // One shard per GPU (or a single shard): the whole cross-shard pipeline is ONE library call per GPU and batch --
    // GPU r ingests slice r of the batch over its own PCIe link, the sketches travel by NCCL all-gather, the top
    // lists by all-to-all, every GPU merges and classifies its slice.  More shards than GPUs (NCCL has one rank per
    // device): every shard is queried with all reads, the lists are merged on GPU 0 from host arrays.
    // Small inputs (< 256 k reads) are not worth the NCCL start-up (1-2 s for the first communicator of a process).
    const size_t NCCL_MIN_READS = getenv("MCG_NCCL_MIN_READS") ? size_t(atoll(getenv("MCG_NCCL_MIN_READS"))) : (size_t(1) << 18);
    const bool sharded = P == 1 || (P <= mcg_device_count() && nreads >= NCCL_MIN_READS);
    std::vector<mcg_comm*> comms;
    if(sharded && P > 1) comms = make_comms(P);
    lap("communicator");
    // two sets of batch buffers: batch i+1 is assembled and batch i-1 is formatted while batch i is on the GPUs
    struct BatchBuf {
        PinnedBuf<char> b1, b2;
        PinnedBuf<int64_t> o1, o2;
        PinnedBuf<int32_t> mtax, mn, best, brank;
        PinnedBuf<uint32_t> mhits;
        std::vector<int32_t> ctax, cn;          // unsharded path: per-shard lists
        std::vector<uint32_t> chits;
        size_t lo = 0, n = 0;
        uint64_t stats[22] = {0};
    } bufs[2];
    auto assemble = [&](BatchBuf& B, size_t lo) {
        const size_t n = std::min(BATCH, nreads - lo);
        B.lo = lo; B.n = n;
        B.o1.reserve(n + 1); B.o2.reserve(paired ? n + 1 : 1);
        B.o1[0] = 0; B.o2[0] = 0;
        for(size_t i = 0; i < n; ++i) {
            B.o1[i + 1] = B.o1[i] + int64_t(r1[lo + i].slen);
            if(paired) B.o2[i + 1] = B.o2[i] + int64_t(r2[lo + i].slen);
        }
        B.b1.reserve(size_t(B.o1[n]) + 64); B.b2.reserve(paired ? size_t(B.o2[n]) + 64 : 64);
        std::vector<std::thread> th;
        for(unsigned t = 0; t < NT; ++t) th.emplace_back([&, t] {
            const size_t ia = n * t / NT, ib = n * (t + 1) / NT;
            for(size_t i = ia; i < ib; ++i) {
                std::memcpy(B.b1.data() + B.o1[i], seq_of(r1[lo + i], txt1, extra1), r1[lo + i].slen);
                if(paired) std::memcpy(B.b2.data() + B.o2[i], seq_of(r2[lo + i], txt2, extra2), r2[lo + i].slen);
            }
        });
        for(auto& x : th) x.join();
        B.mtax.reserve(n * K); B.mhits.reserve(n * K); B.mn.reserve(n); B.best.reserve(n); B.brank.reserve(n);
    };
    auto run_gpu = [&](BatchBuf& B) {
        const size_t n = B.n;
        for(auto& x : B.stats) x = 0;
        if(sharded) {
            std::vector<std::array<uint64_t, 22>> part(static_cast<size_t>(P));
            for_each_shard_parallel(P, [&](int r) {
                const size_t a = n * size_t(r) / size_t(P), b = n * size_t(r + 1) / size_t(P);
                part[size_t(r)].fill(0);
                // slice r: reads [a, b) of the batch -- the library takes a window of the (pinned) offset arrays
                check(mcg_query_batch_sharded(dbs[size_t(r)], P > 1 ? comms[size_t(r)] : nullptr, &q,
                                              B.b1.data(), B.o1.data() + a, paired ? B.b2.data() : nullptr,
                                              paired ? B.o2.data() + a : nullptr, int64_t(b - a),
                                              B.mtax.data() + a * K, B.mhits.data() + a * K, B.mn.data() + a, B.best.data() + a,
                                              B.brank.data() + a, part[size_t(r)].data()), "mcg_query_batch_sharded");
            });
            for(auto& pr : part) for(int i = 0; i < 22; ++i) B.stats[i] += pr[size_t(i)];
        } else {
            B.ctax.resize(size_t(P) * n * K); B.chits.resize(size_t(P) * n * K); B.cn.resize(size_t(P) * n);
            for_each_shard_parallel(P, [&](int r) {          // every GPU queries its shard(s) with all reads of the batch
                check(mcg_query_batch(dbs[size_t(r)], &q, B.b1.data(), B.o1.data(), paired ? B.b2.data() : nullptr, paired ? B.o2.data() : nullptr,
                                      int64_t(n), B.ctax.data() + size_t(r) * n * K, B.chits.data() + size_t(r) * n * K, B.cn.data() + size_t(r) * n),
                      "mcg_query_batch");
            });
            check(mcg_merge_classify(ctx_for_shard(0), &q, B.ctax.data(), B.chits.data(), B.cn.data(), P, int64_t(n), B.mtax.data(), B.mhits.data(),
                                     B.mn.data(), B.best.data(), B.brank.data(), B.stats), "mcg_merge_classify");
        }
    };
    auto emit = [&](BatchBuf& B) {
        const size_t n = B.n, lo = B.lo;
        for(int i = 0; i < 22; ++i) stats[i] += B.stats[i];
        if(tax_counts) for(size_t i = 0; i < n; ++i) if(B.best[i] >= 0) ++tax_reads[size_t(B.best[i])];
        if(o.map == 2) return;
        // per-read lines: formatted by all threads into their own buffers, written in order
        std::vector<StrOut> parts(NT);
        {
            std::vector<std::thread> th;
            for(unsigned t = 0; t < NT; ++t) th.emplace_back([&, t] {
                StrOut& ls = parts[t];
                const size_t ia = n * t / NT, ib = n * (t + 1) / NT;
                ls.s.reserve((ib - ia) * 96);
                for(size_t i = ia; i < ib; ++i) {
                    if(o.map == 1 && B.best[i] < 0) continue;
                    if(o.queryids) ls << (lo + i + 1) << o.col;
                    const ReadRec& rr = r1[lo + i];
                    const char* h = (rr.in_extra ? extra1.data() : txt1.data()) + rr.hpos;
                    const void* sp = std::memchr(h, ' ', rr.hlen);
                    ls.s.append(h, sp ? size_t(static_cast<const char*>(sp) - h) : rr.hlen);
                    ls << o.col;
                    if(o.tophits) {                                                     // printing.cpp:333-360
                        for(int c = 0; c < B.mn[i] && B.mhits[i * K + c] > 0; ++c) {
                            if(c > 0) ls << ',';
                            const int32_t tx = B.mtax[i * K + c];
                            const Taxon* t2 = tt.tax[size_t(tx)];
                            if(q.lowest == 0) ls << t2->name;
                            else {
                                const int32_t an = (t2->rank < q.lowest) ? tt.lineage[size_t(tx) * 21 + size_t(q.lowest)] : tx;
                                if(an >= 0) ls << tt.taxid[size_t(an)]; else ls << t2->name;
                            }
                            ls << ':' << B.mhits[i * K + c];
                        }
                        ls << o.col;
                    }
                    show_classification(ls, o, tt, B.best[i]);
                    ls << '\n';
                }
            });
            for(auto& x : th) x.join();
        }
        for(auto& part : parts) os.write(part.s.data(), std::streamsize(part.s.size()));
    };
    {
        const size_t nbatches = (nreads + BATCH - 1) / BATCH;
        if(nbatches > 0) assemble(bufs[0], 0);
        lap("assemble batch");
        for(size_t bi = 0; bi < nbatches; ++bi) {
            BatchBuf& B = bufs[bi & 1];
            std::thread gpu([&] { run_gpu(B); });
            // meanwhile on the host: the previous batch is written out, the next one assembled
            if(bi >= 1) emit(bufs[(bi - 1) & 1]);
            if(bi + 1 < nbatches) assemble(bufs[(bi + 1) & 1], (bi + 1) * BATCH);
            gpu.join();
            lap("query batch (device) || format previous / assemble next");
        }
        if(nbatches > 0) emit(bufs[(nbatches - 1) & 1]);
        lap("format + write output");
    }
    if(hits_per_seq)                                                            // printing.cpp:437-451, no rows (see above)
        os_tgt << cm << "--- list of hits for each reference sequence ---\n"
               << cm << "window start position within sequence = window_index * window_stride(=" << q.winstride << ")\n"
               << cm << "TABLE_LAYOUT: " << " sequence " << o.col << " windows_in_sequence " << o.col
               << "queryid/window_index:hits/window_index:hits/...,queryid/...\n";
    if(tax_counts) {                                                            // classification.cpp:753-764
        TaxCountMap all;
        for(size_t t = 0; t < tax_reads.size(); ++t) if(tax_reads[t] > 0) all[tt.tax[t]] = float(tax_reads[t]);
        const uint64_t total = stats[20] + stats[21];
        if(show_abund) {
            os_tax << cm << "query summary: number of queries mapped per taxon\n";
            show_abundance_table(os_tax, all, total, o.ranksuf, o.col);
        }
        if(est_rank != 21) {
            estimate_abundance(tt, all, est_rank);
            os_tax << cm << "estimated abundance (number of queries) per " << RANK_NAMES[est_rank] << "\n";
            show_abundance_table(os_tax, all, total, o.ranksuf, o.col);
        }
    }
    const double t1 = now();
    if(o.summary) {                                                             // printing.cpp:522-641
        const uint64_t total = stats[20] + stats[21];
        const uint64_t nq = paired ? 2 * total : total;
        const double minutes = (t1 - t0) / 60.0;
        os << cm << "queries: " << nq << '\n' << cm << "time:    " << int64_t((t1 - t0) * 1000) << " ms\n"
           << cm << "speed:   " << (double(nq) / minutes) << " queries/min\n";
        if(total > 0) {
            if(stats[20] < 1) os << "None of the input sequences could be classified.\n";
            else {
                if(stats[21] > 0) os << cm << "unclassified: " << (100 * (double(stats[21]) / double(total))) << "% (" << stats[21] << ")\n";
                os << cm << "classified:\n";
                const int ranks[] = {0, 3, 4, 6, 10, 12, 14, 16, 18, 19, 20};
                for(int r : ranks) {
                    if(stats[r] > 0) {
                        std::string rn = RANK_NAMES[r];
                        rn.resize(11, ' ');
                        os << cm << "  " << rn << (100 * (double(stats[r]) / double(total))) << "% (" << stats[r] << ")\n";
                    }
                }
            }
        } else {
            std::cerr << cm << "No valid query sequences found.\n";
        }
    }
    for(mcg_comm* m : comms) mcg_comm_destroy(m);
    for(mcg_db* d : dbs) mcg_db_destroy(d);
    return 0;
}

// ------------------------------------------------------------------------------- merge
// `merge <result files ...> -taxonomy <dir>`: classifies reads from the top-hit lists of several independent query
// runs (docs/merge.txt, mode_merge.cpp).  Behaviour: result files in sorted order; each must have been written with
// -tophits -queryids at -lowest species or above; per query the `taxid:hits` entries of every file are inserted, file
// after file, into one candidate list (candidates.h:236-285, merge below -lowest >= species), then classify + statistics
// as in a query.  Here the files play the role of shards: their lists go to mcg_merge_classify in file order.
int mode_merge(const Args& a)
{
    if(a.pos.size() < 1) die("usage: metacache_b200 merge <result files / directories ...> -taxonomy <dir> [options]");
    std::vector<std::string> files;
    for(const auto& p : a.pos) list_files(p, files);
    std::sort(files.begin(), files.end());
    const std::string taxdir = a.get({"taxonomy"}, "");
    if(taxdir.empty()) { std::cout << "No taxonomy specified. Unable to perform merge." << std::endl; return 0; }
    if(files.size() < 2) { std::cout << "Please provide at least two files to be merged!" << std::endl; return 0; }
    start_gpu_init(1);
    Taxonomy tax;
    tax.load(taxdir);
    TaxTables tt;
    tt.build(tax, 0);
    std::cout << "Taxonomy applied to database." << std::endl;
    mcg_query_opts q{};
    q.sketchlen = 16; q.winlen = 128; q.winstride = 113;
    q.lowest = 4; q.highest = 19;                                               // mode_merge.cpp:83-86: species at least
    { const std::string s = a.get({"lowest"}, ""); if(!s.empty()) { int r = rank_from_name(s); if(r < 21 && r > 4) q.lowest = r; } }
    { const std::string s = a.get({"highest"}, ""); if(!s.empty()) { int r = rank_from_name(s); if(r <= 20) q.highest = r; } }
    if(q.lowest > q.highest) q.lowest = q.highest;
    q.hitmin = int32_t(a.geti({"hitmin"}, 0));
    if(q.hitmin < 1) q.hitmin = 5;                                              // mode_merge.cpp:81-82
    {
        float f = float(std::atof(a.get({"hitdiff"}, "1").c_str()));
        if(f > 1) f = float(double(f) * 0.01);
        q.hitdiff = f;
    }
    q.maxcand = int32_t(a.geti({"max-cand", "maxcand", "max_cand", "max-candidates", "maxcandidates"}, 2));
    if(q.maxcand < 1) q.maxcand = 2;                                            // mode_merge.cpp:284-286: "else default to 2"
    if(q.maxcand > MCG_MAX_CAND) q.maxcand = MCG_MAX_CAND;
    const int K = q.maxcand;
    OutOpt o;
    o.queryids = a.has({"queryids", "query-ids", "queryid", "query-id"});
    o.tophits = a.has({"tophits", "top-hits"});
    o.lineage = a.has({"lineage"});
    o.show_ranks = !a.has({"omit-ranks", "omitranks"});
    if(a.has({"taxidsonly", "taxids-only", "taxidonly", "taxid-only"})) o.mode = 1;
    else if(a.has({"taxids", "taxid"})) o.mode = 2;
    if(a.has({"nomap", "no-map"})) o.map = 2; else if(a.has({"mapped-only", "mappedonly"})) o.map = 1;
    o.summary = !a.has({"no-summary", "nosummary"});
    o.params = !a.has({"no-query-params", "noqueryparams"});
    o.lowest = q.lowest; o.highest = q.highest;

    // ---- parse the result files
    struct Source { std::string text; size_t begin = 0, nq = 0; int tophits_col = 0; };
    std::vector<Source> src(files.size());
    size_t n = 0;
    for(size_t f = 0; f < files.size(); ++f) {
        Source& S = src[f];
        if(!read_text_file(files[f], S.text)) die("could not open file " + files[f]);
        size_t p = 0;
        bool have_rank = false, have_layout = false;
        while(p < S.text.size() && S.text[p] == '#') {
            size_t e = S.text.find('\n', p);
            if(e == std::string::npos) e = S.text.size();
            const std::string line = S.text.substr(p, e - p);
            if(!have_rank && line.compare(0, 16, "# Classification") == 0) {
                if(line.find("sequence") != std::string::npos) die("cannot merge results on sequence level (" + files[f] + ")");
                have_rank = true;
            }
            if(line.compare(0, 15, "# TABLE_LAYOUT:") == 0) {
                std::vector<std::string> cols;
                size_t c0 = 15;
                while(c0 <= line.size()) {
                    size_t c1 = line.find('|', c0);
                    if(c1 == std::string::npos) c1 = line.size();
                    cols.push_back(trimmed(line.substr(c0, c1 - c0)));
                    c0 = c1 + 1;
                }
                if(cols.empty() || cols[0] != "query_id") die("no query_id in file " + files[f]);
                for(size_t c = 1; c < cols.size(); ++c) if(cols[c] == "top_hits") { S.tophits_col = int(c); break; }
                have_layout = true;
            }
            p = e + 1;
        }
        if(!have_rank) die("classificaion ranks not found in file " + files[f]);
        if(!have_layout) die("TABLE_LAYOUT not found in file " + files[f]);
        if(S.tophits_col < 1) die("no top_hits in file " + files[f]);
        S.begin = p;
        for(size_t i = p; i < S.text.size();) {
            if(S.text[i] != '#') ++S.nq;
            const size_t e = S.text.find('\n', i);
            if(e == std::string::npos) break;
            i = e + 1;
        }
        n = std::max(n, S.nq);
    }
    const size_t F = files.size();
    std::vector<std::string> headers(n);
    std::vector<int32_t> ctax(F * n * size_t(K), -1), cn(F * n, 0);
    std::vector<uint32_t> chits(F * n * size_t(K), 0);
    for(size_t f = 0; f < F; ++f) {
        const Source& S = src[f];
        size_t p = S.begin;
        while(p < S.text.size()) {
            size_t e = S.text.find('\n', p);
            if(e == std::string::npos) e = S.text.size();
            if(S.text[p] != '#' && e > p) {
                // columns are separated by "\t|\t"
                std::vector<std::pair<size_t, size_t>> col;
                size_t c0 = p;
                while(c0 <= e) {
                    size_t c1 = S.text.find("\t|\t", c0);
                    if(c1 == std::string::npos || c1 > e) c1 = e;
                    col.emplace_back(c0, c1);
                    c0 = c1 + 3;
                }
                size_t qid = size_t(std::atoll(S.text.c_str() + col[0].first));
                if(qid > 0) --qid;
                if(qid < n && int(col.size()) > S.tophits_col) {
                    if(headers[qid].empty() && col.size() > 1) {
                        std::string h = S.text.substr(col[1].first, col[1].second - col[1].first);
                        const size_t sp = h.find_first_of(" \t");
                        headers[qid] = sp == std::string::npos ? h : h.substr(0, sp);
                    }
                    size_t t = col[size_t(S.tophits_col)].first;
                    const size_t te = col[size_t(S.tophits_col)].second;
                    int32_t& cnt = cn[f * n + qid];
                    while(t < te && cnt < K) {
                        const int64_t taxid = std::atoll(S.text.c_str() + t);
                        const size_t colon = S.text.find(':', t);
                        if(colon == std::string::npos || colon >= te) break;
                        const uint32_t hits = uint32_t(std::atoll(S.text.c_str() + colon + 1));
                        auto it = tt.index_of.find(taxid);
                        if(it != tt.index_of.end()) {
                            ctax[(f * n + qid) * size_t(K) + size_t(cnt)] = it->second;
                            chits[(f * n + qid) * size_t(K) + size_t(cnt)] = hits;
                            ++cnt;
                        } else {
                            std::cerr << "taxid not found" << std::endl;
                        }
                        size_t comma = S.text.find(',', colon);
                        if(comma == std::string::npos || comma >= te) break;
                        t = comma + 1;
                    }
                }
            }
            p = e + 1;
        }
    }
    // ---- merge + decision on the device
    wait_gpu_init();
    mcg_ctx* c = ctx_for_shard(0);
    check(mcg_tax_upload(c, tt.lineage.data(), tt.rank.data(), int64_t(tt.rank.size()), 0), "mcg_tax_upload");
    std::vector<int32_t> mtax(n * size_t(K)), mn(n), best(n), brank(n);
    std::vector<uint32_t> mhits(n * size_t(K));
    uint64_t stats[22] = {0};
    const double t0 = now();
    check(mcg_merge_classify(c, &q, ctax.data(), chits.data(), cn.data(), int32_t(F), int64_t(n), mtax.data(), mhits.data(), mn.data(),
                             best.data(), brank.data(), stats), "mcg_merge_classify");
    const double t1 = now();
    // ---- output (mode_merge.cpp:330-375, classification.cpp:889-925)
    std::ofstream fout;
    const std::string outname = a.get({"out"}, "");
    if(!outname.empty()) {
        fout.open(outname);
        if(!fout.good()) die("Could not write to file " + outname);
        std::cout << "Per-Read mappings will be written to file: " << outname << std::endl;
    }
    std::ostream& os = outname.empty() ? std::cout : fout;
    const std::string& cm = o.comment;
    if(o.params) {
        if(o.map != 2) {
            os << cm << "Reporting per-read mappings (non-mapping lines start with '" << cm << "').\n";
            os << cm << (o.lineage ? "The complete lineage will be reported starting with the lowest match.\n"
                                   : "Only the lowest matching rank will be reported.\n");
        } else {
            os << cm << "Per-Read mappings will not be shown.\n";
        }
        os << cm << "Classification will be constrained to ranks from '" << RANK_NAMES[q.lowest] << "' to '" << RANK_NAMES[q.highest] << "'.\n";
        os << cm << "Classification hit threshold is " << q.hitmin << " per query\n";
        os << cm << "At maximum " << q.maxcand << " classification candidates will be considered per query.\n";
        os << cm << "Using 1 threads\n";
    }
    os << cm << "Merging " << F << " files:\n";
    for(const auto& f : files) os << cm << f << '\n';
    if(o.map != 2) {
        StrOut ls;
        for(size_t i = 0; i < n; ++i) {
            if(o.map == 1 && best[i] < 0) continue;
            if(o.queryids) ls << (i + 1) << o.col;
            ls << headers[i] << o.col;
            if(o.tophits) {
                for(int k = 0; k < mn[i] && mhits[i * size_t(K) + size_t(k)] > 0; ++k) {
                    if(k > 0) ls << ',';
                    ls << tt.taxid[size_t(mtax[i * size_t(K) + size_t(k)])] << ':' << mhits[i * size_t(K) + size_t(k)];
                }
                ls << o.col;
            }
            show_classification(ls, o, tt, best[i]);
            ls << '\n';
        }
        os.write(ls.s.data(), std::streamsize(ls.s.size()));
    }
    if(o.summary) {
        const uint64_t total = stats[20] + stats[21];
        const double minutes = (t1 - t0) / 60.0;
        os << cm << "queries: " << total << '\n' << cm << "time:    " << int64_t((t1 - t0) * 1000) << " ms\n"
           << cm << "speed:   " << (double(total) / minutes) << " queries/min\n";
        if(total > 0) {
            if(stats[20] < 1) os << "None of the input sequences could be classified.\n";
            else {
                if(stats[21] > 0) os << cm << "unclassified: " << (100 * (double(stats[21]) / double(total))) << "% (" << stats[21] << ")\n";
                os << cm << "classified:\n";
                const int ranks[] = {0, 3, 4, 6, 10, 12, 14, 16, 18, 19, 20};
                for(int r : ranks) {
                    if(stats[r] > 0) {
                        std::string rn = RANK_NAMES[r];
                        rn.resize(11, ' ');
                        os << cm << "  " << rn << (100 * (double(stats[r]) / double(total))) << "% (" << stats[r] << ")\n";
                    }
                }
            }
        }
    }
    return 0;
}

// ------------------------------------------------------------------------------- host-only self tests
// (no CUDA call: they run in the CPU test suite)

// selftest-shard <file> <nkeys>: pseudo-random buckets -> write_shard -> read_shard -> compare
int mode_selftest_shard(const Args& a)
{
    if(a.pos.size() < 2) die("usage: metacache_b200 selftest-shard <file> <nkeys>");
    const uint64_t nk = uint64_t(std::atoll(a.pos[1].c_str()));
    DbMeta m;
    m.ntargets = 3;
    for(int t = 0; t < 3; ++t) {
        Taxon x; x.id = -t - 1; x.parent = 100 + t; x.rank = 0; x.name = "SEQ" + std::to_string(t); x.file = "dir/g" + std::to_string(t) + ".fna";
        x.index = 0; x.windows = 1000 + uint64_t(t);
        m.tax.taxa[x.id] = x;
    }
    { Taxon x; x.id = 1; x.parent = 1; x.rank = 20; x.name = "root"; m.tax.taxa[1] = x; }
    ShardData d;
    uint64_t st = 88172645463325252ull;
    auto rnd = [&] { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; };
    d.keys.resize(nk); d.sizes.resize(nk);
    for(uint64_t j = 0; j < nk; ++j) {
        d.keys[j] = uint32_t(j * 7919u + 13u);
        const uint64_t r = rnd() % 100;
        d.sizes[j] = r < 70 ? 1 : (r < 90 ? 2 : (r < 99 ? uint32_t(3 + rnd() % 20) : 254));
        d.nl += d.sizes[j];
    }
    d.ne = nk;
    d.tg.resize(d.nl); d.wn.resize(d.nl);
    for(uint64_t i = 0; i < d.nl; ++i) { d.tg[i] = uint32_t(rnd() % 3); d.wn[i] = uint32_t(rnd() % 1000); }
    write_shard(a.pos[0], m, d);
    DbMeta m2;
    std::vector<uint32_t> k2, s2, t2, w2;
    read_shard(a.pos[0], m2, false, k2, s2, t2, w2);
    bool ok = k2 == d.keys && s2 == d.sizes && t2 == d.tg && w2 == d.wn && m2.ntargets == 3 && m2.tax.taxa.size() == 4 &&
              m2.tax.taxa[-2].file == "dir/g1.fna" && m2.tax.taxa[-3].windows == 1002 && m2.maxlocs == m.maxlocs &&
              m2.tcfg.k == 16 && m2.qstride == 113;
    std::cout << (ok ? "selftest-shard ok" : "selftest-shard MISMATCH") << ": " << nk << " keys, " << d.nl << " locations" << std::endl;
    return ok ? 0 : 1;
}

// shards-equal <file a> <file b>: same header parameters, taxa, target count and -- order-insensitively, the
// reference writes buckets in hash-table order -- the same buckets with the same location lists
int mode_shards_equal(const Args& a)
{
    if(a.pos.size() < 2) die("usage: metacache_b200 shards-equal <shard file> <shard file>");
    DbMeta m[2];
    std::vector<uint32_t> keys[2], sizes[2], tg[2], wn[2];
    for(int i = 0; i < 2; ++i) read_shard(a.pos[size_t(i)], m[i], false, keys[i], sizes[i], tg[i], wn[i]);
    bool ok = m[0].tcfg.k == m[1].tcfg.k && m[0].tcfg.sketchlen == m[1].tcfg.sketchlen && m[0].tcfg.winlen == m[1].tcfg.winlen &&
              m[0].tcfg.winstride == m[1].tcfg.winstride && m[0].qsketch == m[1].qsketch && m[0].qwinlen == m[1].qwinlen &&
              m[0].qstride == m[1].qstride && m[0].maxlocs == m[1].maxlocs && m[0].ntargets == m[1].ntargets &&
              m[0].tax.taxa.size() == m[1].tax.taxa.size() && keys[0].size() == keys[1].size() && tg[0].size() == tg[1].size();
    if(ok) {
        auto ia = m[0].tax.taxa.begin(), ib = m[1].tax.taxa.begin();
        for(; ia != m[0].tax.taxa.end(); ++ia, ++ib) {
            const Taxon &x = ia->second, &y = ib->second;
            auto base = [](const std::string& f) { const size_t p = f.rfind('/'); return p == std::string::npos ? f : f.substr(p + 1); };
            if(x.id != y.id || x.parent != y.parent || x.rank != y.rank || x.name != y.name || base(x.file) != base(y.file) ||
               x.index != y.index || x.windows != y.windows) { ok = false; break; }
        }
    }
    if(ok) {
        // bucket order by key on both sides
        std::vector<uint64_t> start[2];
        std::vector<uint32_t> ord[2];
        for(int i = 0; i < 2; ++i) {
            start[i].assign(keys[i].size() + 1, 0);
            for(size_t j = 0; j < keys[i].size(); ++j) start[i][j + 1] = start[i][j] + sizes[i][j];
            ord[i].resize(keys[i].size());
            for(size_t j = 0; j < ord[i].size(); ++j) ord[i][j] = uint32_t(j);
            std::sort(ord[i].begin(), ord[i].end(), [&](uint32_t x, uint32_t y) { return keys[i][x] < keys[i][y]; });
        }
        for(size_t j = 0; j < ord[0].size() && ok; ++j) {
            const uint32_t x = ord[0][j], y = ord[1][j];
            if(keys[0][x] != keys[1][y] || sizes[0][x] != sizes[1][y]) { ok = false; break; }
            const uint64_t n = sizes[0][x];
            if(std::memcmp(&tg[0][start[0][x]], &tg[1][start[1][y]], n * 4) != 0 || std::memcmp(&wn[0][start[0][x]], &wn[1][start[1][y]], n * 4) != 0) ok = false;
        }
    }
    std::cout << (ok ? "shards equal" : "shards DIFFER") << ": " << keys[0].size() << " buckets, " << tg[0].size() << " locations" << std::endl;
    return ok ? 0 : 1;
}

// selftest-taxmap <sequence files / directories ...>: the parent taxon `build` gives every sequence (tables next to the
// files, then the header), one "file <tab> sequence id <tab> taxon" line per sequence
int mode_selftest_taxmap(const Args& a)
{
    std::vector<std::string> files;
    for(const auto& p : a.pos) list_files(p, files);
    std::sort(files.begin(), files.end());
    SequenceTaxa mapped;
    mapped.read_tables_next_to(files);
    std::string txt;
    for(const auto& f : files) {
        if(!read_text_file(f, txt)) continue;
        std::vector<Record> recs;
        parse_sequences(txt, recs);
        const std::string file_id = extract_accession(f);
        for(const auto& rec : recs) {
            if(rec.data.empty()) continue;
            std::string sid = extract_accession(rec.header);
            if(sid.empty()) sid = rec.header;
            int64_t parent = mapped.taxon_of(sid);
            if(parent == 0) parent = mapped.taxon_of(file_id);
            if(parent == 0) parent = extract_taxid(rec.header);
            std::cout << f.substr(f.find_last_of('/') + 1) << "\t" << sid << "\t" << (parent < 1 ? 0 : parent) << "\n";
        }
    }
    return 0;
}

// selftest-abundance <taxonomy dir> <rank> <taxid:count> ...: the two tables of -abundances / -abundance-per
int mode_selftest_abundance(const Args& a)
{
    if(a.pos.size() < 3) die("usage: metacache_b200 selftest-abundance <taxonomy dir> <rank> <taxid:count> ...");
    Taxonomy tax;
    tax.load(a.pos[0]);
    TaxTables tt;
    tt.build(tax, 0);
    const int rank = rank_from_name(a.pos[1]);
    TaxCountMap all;
    uint64_t total = 0;
    for(size_t i = 2; i < a.pos.size(); ++i) {
        const size_t c = a.pos[i].find(':');
        if(c == std::string::npos) die("expected taxid:count");
        const int64_t id = std::atoll(a.pos[i].substr(0, c).c_str());
        const uint64_t n = uint64_t(std::atoll(a.pos[i].substr(c + 1).c_str()));
        auto it = tt.index_of.find(id);
        if(it == tt.index_of.end()) die("unknown taxon " + std::to_string(id));
        all[tt.tax[size_t(it->second)]] = float(n);
        total += n;
    }
    if(a.has({"total"})) total = uint64_t(a.geti({"total"}, 0));
    std::cout << "# query summary: number of queries mapped per taxon\n";
    show_abundance_table(std::cout, all, total, ":", "\t|\t");
    estimate_abundance(tt, all, rank);
    std::cout << "# estimated abundance (number of queries) per " << RANK_NAMES[rank] << "\n";
    show_abundance_table(std::cout, all, total, ":", "\t|\t");
    return 0;
}

}  // namespace

int main(int argc, char** argv)
{
    Lap total;
    if(argc < 2) {
        std::cerr << "usage: metacache_b200 build|query ... (see the header of host/metacache_b200.cpp)" << std::endl;
        return 1;
    }
    const std::string mode = argv[1];
    const Args a = parse_args(argc, argv, 2);
    int rc = 1;
    if(mode == "build") rc = mode_build(a);
    else if(mode == "query") rc = mode_query(a);
    else if(mode == "merge") rc = mode_merge(a);
    else if(mode == "selftest-shard") return mode_selftest_shard(a);
    else if(mode == "selftest-abundance") return mode_selftest_abundance(a);
    else if(mode == "selftest-taxmap") return mode_selftest_taxmap(a);
    else if(mode == "shards-equal") return mode_shards_equal(a);
    else die("unknown mode '" + mode + "' (build, query and merge are implemented)");
    wait_gpu_init();
    total("whole run before teardown");
    std::cout.flush();
    std::fflush(nullptr);
    if(getenv("MCG_SLOW_EXIT") == nullptr) _exit(rc);        // the driver reclaims the contexts faster than an orderly teardown
    for(mcg_ctx* c : g_ctx) if(c) mcg_ctx_destroy(c);
    total("context teardown");
    return rc;
}
