// host.cpp -- host orchestration behind the unchanged `pregraph -s cfg -K k -p P [-a G] [-d D] [-R] -o prefix` CLI,
// plus the extern "C" veneer declared in include/pregraph_b200.h.
//
// Mirrors, in new code, the host-side behaviour of (standardPregraph/):
//   call_pregraph / initenv        pregraph.c:62-220   (getopt string, K fix-ups, phase order, stderr lines)
//   scan_libInfo / splitColumn     lib.c:70-506        (key=value config, [LIB] sections, sort by avg_ins)
//   openNextFile / nextValidIndex  prlHashReads.c:903-951, readseq1by1.c:595-674 (library + file-type iteration order)
//   file writers                   prlHashReads.c:1104-1132 (.kmerFreq), node2edge.c:61-70 (.edge.gz via zlib),
//                                  prlRead2path.c:426-476 (.preArc/.markOnEdge), output_pregraph.c:50-86 (.vertex, .preGraphBasic)
// All k-mer work happens on the GPU through IEngine; this file only moves bytes between files and the engine.
#include "../../include/pregraph_b200.h"
#include "engine.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>
#include <getopt.h>
#include <unistd.h>
#include <zlib.h>
#include <cuda_runtime_api.h>

using namespace pgb;

static thread_local std::string g_err;
struct pgb200_engine {
    IEngine* e;
    PgParams prm;
};

// every engine entry point binds the calling thread to its engine's GPU first (one process may drive several engines)
#define PG_TRY try { cudaSetDevice(e->prm.device);
#define PG_CATCH                                   \
    }                                              \
    catch (const std::exception& ex) {             \
        g_err = ex.what();                         \
        return -1;                                 \
    }                                              \
    catch (...) {                                  \
        g_err = "unknown error";                   \
        return -1;                                 \
    }                                              \
    return 0;

extern "C" const char* pgb200_last_error(void) { return g_err.c_str(); }

extern "C" void pgb200_default_params(pgb200_params* p) {
    memset(p, 0, sizeof *p);
    p->K = 23; p->P = 8; p->max_rd_len = 100; p->world = 1;
}

extern "C" pgb200_engine* pgb200_create(const pgb200_params* p) {
    try {
        PgParams q;
        q.K = p->K; q.P = p->P; q.initG = p->initG; q.D = p->D; q.repsTie = p->repsTie; q.flavour127 = p->flavour127;
        q.device = p->device; q.max_rd_len = p->max_rd_len > 0 ? p->max_rd_len : 100; q.table_slots = p->table_slots;
        q.verbose = p->verbose; q.world = p->world > 0 ? p->world : 1; q.rank = p->rank;
        if (q.K < 13 || q.K % 2 == 0 || q.K > (q.flavour127 ? 127 : 63)) throw std::runtime_error("pgb200: K must be odd, 13..63 (63-mer flavour) or 13..127 (127-mer flavour)");
        // first-occurrence rank = (read ordinal << 16) | k-mer position: positions must fit 16 bits
        if (q.max_rd_len - q.K + 1 > 65536) throw std::runtime_error("pgb200: max_rd_len - K + 1 must not exceed 65536 (k-mer positions are 16-bit)");
        if (q.world > 16 || q.rank < 0 || q.rank >= q.world) throw std::runtime_error("pgb200: world must be 1..16 and 0 <= rank < world");
        pgb200_engine* h = new pgb200_engine;
        h->prm = q;
        h->e = make_engine(q);
        return h;
    } catch (const std::exception& ex) {
        g_err = ex.what();
        return nullptr;
    }
}
extern "C" void pgb200_destroy(pgb200_engine* e) {
    if (!e) return;
    delete e->e;
    delete e;
}
extern "C" void* pgb200_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { g_err = "cudaHostAlloc failed"; return nullptr; }
    return p;
}
extern "C" void pgb200_host_free(void* p) { if (p) cudaFreeHost(p); }

extern "C" int pgb200_feed_text(pgb200_engine* e, const char* text, size_t nbytes, int on_device, int fastq, uint64_t ord_base,
                                uint64_t ord_stride, int reverse_seq, int maxlen) {
    PG_TRY e->e->feed_text(text, nbytes, on_device != 0, fastq, ord_base, ord_stride, reverse_seq, maxlen); PG_CATCH
}
extern "C" uint64_t pgb200_last_chunk_records(pgb200_engine* e) { return e->e->last_chunk_records(); }
extern "C" int pgb200_xchg_setup(pgb200_engine* e, uint64_t cap_records) { PG_TRY e->e->xchg_setup(cap_records); PG_CATCH }
extern "C" int pgb200_xchg_export(pgb200_engine* e, void* handle64) { PG_TRY e->e->xchg_export(handle64); PG_CATCH }
extern "C" int pgb200_xchg_import(pgb200_engine* e, int peer, const void* handle64) { PG_TRY e->e->xchg_import(peer, handle64); PG_CATCH }
extern "C" void* pgb200_xchg_base(pgb200_engine* e) { try { return e->e->xchg_base(); } catch (const std::exception& ex) { g_err = ex.what(); return nullptr; } }
extern "C" int pgb200_xchg_import_ptr(pgb200_engine* e, int peer, int peer_device, void* base) { PG_TRY e->e->xchg_import_ptr(peer, peer_device, base); PG_CATCH }
extern "C" int pgb200_xchg_fence(pgb200_engine* e) { PG_TRY e->e->xchg_fence(); PG_CATCH }
extern "C" int pgb200_flush(pgb200_engine* e) { PG_TRY e->e->flush(); PG_CATCH }
extern "C" int pgb200_xchg_room(pgb200_engine* e, uint64_t n_rec) { try { return e->e->xchg_room(n_rec) ? 1 : 0; } catch (const std::exception& ex) { g_err = ex.what(); return -1; } }
extern "C" int pgb200_absorb(pgb200_engine* e, pgb200_engine* other) { PG_TRY e->e->absorb(other->e); PG_CATCH }
extern "C" int pgb200_finish_pass1(pgb200_engine* e, pgb200_pass1_stats* st) {
    PG_TRY
    Pass1Stats s;
    e->e->finish_pass1(&s);
    if (st) {
        st->records = s.records; st->reads_kept = s.reads_kept; st->instances = s.instances; st->distinct = s.distinct;
        st->table_slots = s.table_slots; st->launches = s.launches; st->ms_decode = s.ms_decode; st->ms_insert = s.ms_insert; st->ms_apply = s.ms_apply;
    }
    PG_CATCH
}
extern "C" int pgb200_reset_pass1(pgb200_engine* e) { PG_TRY e->e->reset_pass1(); PG_CATCH }
extern "C" int pgb200_sweeps(pgb200_engine* e, long long hist[256], uint64_t* linear_marked, uint64_t* removed) {
    PG_TRY
    SweepStats s;
    e->e->sweeps(&s);
    if (hist) memcpy(hist, s.hist, sizeof s.hist);
    if (linear_marked) *linear_marked = s.linear;
    if (removed) *removed = s.removed;
    PG_CATCH
}
extern "C" int pgb200_build_layout(pgb200_engine* e) { PG_TRY e->e->build_layout(); PG_CATCH }
extern "C" uint64_t pgb200_node_count(pgb200_engine* e) { return e->e->node_count(); }
extern "C" int pgb200_dump_nodes(pgb200_engine* e, void* out) { PG_TRY e->e->dump_nodes(out); PG_CATCH }

// ------------------------------------------------------------------------------------------------ file writers
static void write_file(const std::string& name, const void* data, size_t n) {
    FILE* f = fopen(name.c_str(), "wb");
    if (!f) throw std::runtime_error("Cannot open " + name + ". Now exit to system...");   // ckopen, check.c:30-34
    if (n && fwrite(data, 1, n, f) != n) { fclose(f); throw std::runtime_error("short write on " + name); }
    fclose(f);
}

static void tip_lines(const TipStats& t, int K, int D) {
    if (D == 0) {
        fprintf(stderr, "Start to remove frequency-one-kmer tips shorter than %d.\n", 2 * K);
        fprintf(stderr, "Total %llu tip(s) removed.\n", (unsigned long long)t.single_tips);
        fprintf(stderr, "%llu linear node(s) marked.\n", (unsigned long long)t.single_relinear);
    }
    fprintf(stderr, "Start to remove tips with minority links.\n");
    for (size_t i = 0; i < t.minor_cycles.size(); i++) fprintf(stderr, "%llu tip(s) removed in cycle %zu.\n", (unsigned long long)t.minor_cycles[i], i + 1);
    fprintf(stderr, "Total %llu tip(s) removed.\n", (unsigned long long)t.minor_tips);
    fprintf(stderr, "%llu linear node(s) marked.\n", (unsigned long long)t.minor_relinear);
}

extern "C" int pgb200_remove_tips(pgb200_engine* e, pgb200_graph_stats* st) {
    PG_TRY
    TipStats t;
    e->e->remove_tips(&t);
    tip_lines(t, e->prm.K, (int)(signed char)e->prm.D);
    if (st) { st->single_tips = t.single_tips; st->minor_tips = t.minor_tips; }
    PG_CATCH
}
// gzopen(name,"w") + gzwrite: same zlib, same default level => the same byte stream as the reference's gzprintf calls
static void write_edge_gz(const std::string& name, const std::string& text) {
    gzFile gz = gzopen(name.c_str(), "w");
    if (!gz) throw std::runtime_error("Cannot open " + name);
    size_t off = 0;
    while (off < text.size()) {
        size_t n = std::min<size_t>(text.size() - off, 1u << 30);
        if (gzwrite(gz, text.data() + off, (unsigned)n) != (int)n) { gzclose(gz); throw std::runtime_error("gzwrite failed on " + name); }
        off += n;
    }
    gzclose(gz);
}

// f2: the edges as a binary sidecar for a `contig` that links csrc/contig_sidecar.c (format: include/pregraph_b200.h).  Converts the
// edge TEXT the GPU emitted (">length L,<from>,<to>,cvg C, B" + bases, output_pregraph.c:88-110) -- host only, no GPU involved.
static bool parse_hex_words(const char*& p, const char* end, uint64_t* w, int n) {
    for (int i = 0; i < n; i++) {
        uint64_t v = 0;
        int digits = 0;
        while (p < end) {
            char c = *p;
            int d = c >= '0' && c <= '9' ? c - '0' : (c >= 'a' && c <= 'f' ? c - 'a' + 10 : -1);
            if (d < 0) break;
            v = (v << 4) | (uint64_t)d;
            p++; digits++;
        }
        if (!digits) return false;
        w[i] = v;
        if (i + 1 < n) { if (p >= end || *p != ' ') return false; p++; }
    }
    return true;
}
static bool parse_int(const char*& p, const char* end, long long* out) {
    long long v = 0;
    int digits = 0;
    while (p < end && *p >= '0' && *p <= '9') { v = v * 10 + (*p - '0'); p++; digits++; }
    *out = v;
    return digits > 0;
}
static bool expect(const char*& p, const char* end, const char* lit) {
    size_t n = strlen(lit);
    if ((size_t)(end - p) < n || memcmp(p, lit, n) != 0) return false;
    p += n;
    return true;
}
extern "C" int pgb200_edge_text_to_sidecar(const char* text, size_t nbytes, int K, int flavour127, uint64_t num_ed, const char* path) {
    try {
        const int kw = flavour127 ? 4 : 2;
        std::string out;
        out.reserve(nbytes / 3 + 4096);
        struct { char magic[8]; uint32_t version, K, kmer_words, r0; uint64_t n_records, num_ed, r1; } h;
        memset(&h, 0, sizeof h);
        memcpy(h.magic, "PGB2EDGE", 8);
        h.version = 1; h.K = (uint32_t)K; h.kmer_words = (uint32_t)kw; h.num_ed = num_ed;
        out.append(reinterpret_cast<const char*>(&h), sizeof h);
        const char* p = text;
        const char* end = text + nbytes;
        uint64_t n_rec = 0;
        while (p < end) {
            long long length, cvg, bal;
            uint64_t from[4], to[4];
            if (!expect(p, end, ">length ") || !parse_int(p, end, &length) || !expect(p, end, ",") || !parse_hex_words(p, end, from, kw) || !expect(p, end, ",") ||
                !parse_hex_words(p, end, to, kw) || !expect(p, end, ",cvg ") || !parse_int(p, end, &cvg) || !expect(p, end, ", ") || !parse_int(p, end, &bal) ||
                !expect(p, end, "\n"))
                throw std::runtime_error("pgb200: edge text does not parse (record " + std::to_string(n_rec) + ")");
            int32_t rec[4] = {(int32_t)length, (int32_t)cvg, (int32_t)bal, (int32_t)(length / 4 + 1)};
            out.append(reinterpret_cast<const char*>(rec), sizeof rec);
            out.append(reinterpret_cast<const char*>(from), kw * 8);
            out.append(reinterpret_cast<const char*>(to), kw * 8);
            const size_t seq0 = out.size();
            out.append((size_t)rec[3], '\0');
            long long pos = 0;
            while (pos < length) {
                if (p >= end) throw std::runtime_error("pgb200: edge text ends inside a sequence");
                const char c = *p++;
                if (c == '\n') continue;
                const unsigned code = ((unsigned)c & 6u) >> 1;                       // base2int, inc/def.h:39
                out[seq0 + (size_t)(pos >> 2)] |= (char)(code << (6 - 2 * (pos & 3)));   // writeChar2tightString, seq.c:81-107
                pos++;
            }
            if (p < end && *p == '\n') p++;
            n_rec++;
        }
        reinterpret_cast<decltype(h)*>(&out[0])->n_records = n_rec;
        write_file(path, out.data(), out.size());
    } catch (const std::exception& ex) {
        g_err = ex.what();
        return -1;
    }
    return 0;
}

// The way back: <prefix>.edge.b200 -> the byte-identical <prefix>.edge.gz (the sidecar holds every field of the text; the record
// layout is output_pregraph.c:88-110: header line, then the bases 100 per line).  Host only; a pipeline that ran the stage with
// PGB200_EDGE_SIDECAR=only can produce the .edge.gz later, or beside `contig`, with `pregraph-b200-<flavour> edgegz -g prefix`.
extern "C" int pgb200_sidecar_to_edge_gz(const char* prefix) {
    try {
        const std::string in = std::string(prefix) + ".edge.b200";
        FILE* f = fopen(in.c_str(), "rb");
        if (!f) throw std::runtime_error("pgb200: cannot open " + in);
        std::string raw;
        char buf[1 << 16];
        size_t got;
        while ((got = fread(buf, 1, sizeof buf, f)) > 0) raw.append(buf, got);
        fclose(f);
        struct Hdr { char magic[8]; uint32_t version, K, kmer_words, r0; uint64_t n_records, num_ed, r1; } h;
        if (raw.size() < sizeof h) throw std::runtime_error("pgb200: " + in + " is truncated");
        memcpy(&h, raw.data(), sizeof h);
        if (memcmp(h.magic, "PGB2EDGE", 8) != 0 || h.version != 1 || (h.kmer_words != 2 && h.kmer_words != 4)) throw std::runtime_error("pgb200: " + in + " is not an edge sidecar");
        const int kw = (int)h.kmer_words;
        std::string text;
        text.reserve(raw.size() * 4 + (1 << 20));
        size_t off = sizeof h;
        for (uint64_t r = 0; r < h.n_records; r++) {
            int32_t rec[4];
            uint64_t km[8];
            if (off + sizeof rec + (size_t)kw * 16 > raw.size()) throw std::runtime_error("pgb200: " + in + " is truncated");
            memcpy(rec, raw.data() + off, sizeof rec); off += sizeof rec;
            memcpy(km, raw.data() + off, (size_t)kw * 16); off += (size_t)kw * 16;
            if (rec[0] < 0 || rec[3] != rec[0] / 4 + 1 || off + (size_t)rec[3] > raw.size()) throw std::runtime_error("pgb200: " + in + " is corrupt");
            int n = snprintf(buf, sizeof buf, ">length %d,", rec[0]);
            for (int side = 0; side < 2; side++) {
                for (int w = 0; w < kw; w++) n += snprintf(buf + n, sizeof buf - n, w ? " %llx" : "%llx", (unsigned long long)km[side * kw + w]);
                buf[n++] = ',';
            }
            n += snprintf(buf + n, sizeof buf - n, "cvg %d, %d\n", rec[1], rec[2]);
            text.append(buf, (size_t)n);
            const unsigned char* seq = reinterpret_cast<const unsigned char*>(raw.data() + off);
            for (int i = 0; i < rec[0]; i++) {
                text.push_back("ACTG"[(seq[i >> 2] >> (6 - 2 * (i & 3))) & 3]);
                if ((i + 1) % 100 == 0) text.push_back('\n');
            }
            if (rec[0] % 100 != 0) text.push_back('\n');
            off += (size_t)rec[3];
        }
        write_edge_gz(std::string(prefix) + ".edge.gz", text);
    } catch (const std::exception& ex) {
        g_err = ex.what();
        return -1;
    }
    return 0;
}

// PGB200_EDGE_SIDECAR unset: <prefix>.edge.gz only (the reference's output).  Set: the sidecar first (a fraction of a second, so a
// contig that links contig_sidecar.c never waits for the deflate), then the .edge.gz.  "only": the sidecar alone -- the deflate of
// the edge text is sequential host work (its bytes must equal the reference's gz stream) and is the longest single item of a
// full-size stage run, so a pipeline whose contig reads the sidecar can leave it out.
static void write_edge_outputs(const std::string& prefix, const std::string& text, int K, int flavour127, uint64_t num_ed) {
    const char* sc = getenv("PGB200_EDGE_SIDECAR");
    if (sc && pgb200_edge_text_to_sidecar(text.data(), text.size(), K, flavour127, num_ed, (prefix + ".edge.b200").c_str())) throw std::runtime_error(g_err);
    if (!(sc && !strcmp(sc, "only"))) write_edge_gz(prefix + ".edge.gz", text);
}

extern "C" int pgb200_kmer2edges(pgb200_engine* e, const char* prefix, pgb200_graph_stats* st) {
    PG_TRY
    EdgeStats es;
    std::string text;
    e->e->build_edges(&es, &text);
    write_edge_outputs(prefix, text, e->prm.K, e->prm.flavour127, es.num_ed);
    fprintf(stderr, "%llu (%llu) edge(s) and %llu extra node(s) constructed.\n", (unsigned long long)es.num_ed, (unsigned long long)es.edges,
            (unsigned long long)es.extra_nodes);
    if (st) { st->num_ed = es.num_ed; st->edges = es.edges; st->extra_nodes = es.extra_nodes; }
    PG_CATCH
}
extern "C" int pgb200_read2edge(pgb200_engine* e, const char* prefix, pgb200_graph_stats* st) {
    PG_TRY
    Pass2Stats ps;
    std::string arcs, path, mark;
    e->e->pass2(&ps, &arcs, &path, &mark);
    write_file(std::string(prefix) + ".preArc", arcs.data(), arcs.size());
    if (e->prm.repsTie) {
        write_file(std::string(prefix) + ".path", path.data(), path.size());
        write_file(std::string(prefix) + ".markOnEdge", mark.data(), mark.size());
        fprintf(stderr, "%llu marker(s) output.\n", (unsigned long long)ps.markers);
    }
    fprintf(stderr, "Reads alignment done, %llu read(s) deleted, %llu pre-arc(s) added.\n", (unsigned long long)ps.deleted_reads,
            (unsigned long long)ps.arcs);
    if (st) { st->deleted_reads = ps.deleted_reads; st->arcs = ps.arcs; }
    PG_CATCH
}
extern "C" int pgb200_output_vertex(pgb200_engine* e, const char* prefix, pgb200_graph_stats* st) {
    PG_TRY
    std::string vt;
    uint64_t nv = 0;
    e->e->vertices(&vt, &nv);
    write_file(std::string(prefix) + ".vertex", vt.data(), vt.size());
    fprintf(stderr, "%llu vertex(es) output.\n", (unsigned long long)nv);
    char buf[512];
    const uint64_t num_ed = e->e->num_ed();   // the engine's own count (st is an output here)
    int n = snprintf(buf, sizeof buf, "VERTEX %llu K %d\n\nEDGEs %llu\n\nMaxReadLen %d MinReadLen %d MaxNameLen %d\n", (unsigned long long)nv,
                     e->prm.K, (unsigned long long)num_ed, e->prm.max_rd_len, 0, 256);
    write_file(std::string(prefix) + ".preGraphBasic", buf, n);
    if (st) { st->vertices = nv; st->num_ed = num_ed; }
    PG_CATCH
}

// ------------------------------------------------------------------------------------------------ config (lib.c)
struct Lib {
    int avg_ins = 0, asm_flag = 3, reverse = 0, rd_len_cutoff = 0;
    std::vector<std::string> f[7];   // [1]=f1 [0]=f2 [2]=q1 [4]=q2 [3]=p [5]=f [6]=q
};
static bool split_column(const char* line, std::string& a, std::string& b) {   // splitColumn lib.c:70-108
    int len = (int)strlen(line), i = 0, n = 0;
    std::string* t[2] = {&a, &b};
    a.clear(); b.clear();
    while (i < len) {
        if (line[i] >= 32 && line[i] <= 126 && line[i] != '=') {
            while (i < len && line[i] >= 32 && line[i] <= 126 && line[i] != '=') t[n]->push_back(line[i++]);
            if (++n == 2) return true;
        }
        i++;
    }
    return false;
}
static void scan_lib(const char* cfg, std::vector<Lib>& libs, int& max_rd_len) {
    FILE* fp = fopen(cfg, "r");
    if (!fp) { fprintf(stderr, "Cannot open %s. Now exit to system...\n", cfg); exit(-1); }
    char line[1024];
    std::string a, b;
    max_rd_len = 0;
    while (fgets(line, sizeof line, fp)) {
        if (strncmp(line, "[LIB]", 5) == 0) { libs.emplace_back(); continue; }
        if (!split_column(line, a, b)) continue;
        if (libs.empty()) { if (a == "max_rd_len") max_rd_len = atoi(b.c_str()); continue; }   // only before the first [LIB] (lib.c:152-165)
        Lib& L = libs.back();
        if (a == "f1") L.f[1].push_back(b); else if (a == "f2") L.f[0].push_back(b);
        else if (a == "q1") L.f[2].push_back(b); else if (a == "q2") L.f[4].push_back(b);
        else if (a == "p") L.f[3].push_back(b); else if (a == "f") L.f[5].push_back(b); else if (a == "q") L.f[6].push_back(b);
        else if (a == "b") { fprintf(stderr, "pgb200: BAM input (b=) is not supported by the GPU engine\n"); exit(-1); }
        else if (a == "avg_ins") L.avg_ins = atoi(b.c_str()); else if (a == "reverse_seq") L.reverse = atoi(b.c_str());
        else if (a == "asm_flags") L.asm_flag = atoi(b.c_str()); else if (a == "rd_len_cutoff") L.rd_len_cutoff = atoi(b.c_str());
    }
    fclose(fp);
    if (libs.empty()) { fprintf(stderr, "Config file error: no [LIB] in file\n"); exit(-1); }
    for (size_t i = 0; i < libs.size(); i++) {
        if (libs[i].f[1].size() != libs[i].f[0].size()) { fprintf(stderr, "Config file error: the number of mark \"f1\" is not the same as \"f2\"!\n"); exit(-1); }
        if (libs[i].f[2].size() != libs[i].f[4].size()) { fprintf(stderr, "Config file error: the number of mark \"q1\" is not the same as \"q2\"!\n"); exit(-1); }
        bool pe = !libs[i].f[1].empty() || !libs[i].f[2].empty() || !libs[i].f[3].empty();
        if (pe && libs[i].avg_ins == 0) { fprintf(stderr, "Config file error: PE reads need avg_ins in [LIB] %zu\n", i + 1); exit(-1); }
    }
    std::stable_sort(libs.begin(), libs.end(), [](const Lib& x, const Lib& y) { return x.avg_ins < y.avg_ins; });   // qsort by avg_ins, lib.c:505
}

// ------------------------------------------------------------------------------------------------ the read-stream plan
// One entry per file in the order the reference opens them (openNextFile / nextValidIndex): libraries sorted by avg_ins, only
// asm_flags 1|3; inside a library f1/f2 pairs, q1/q2 pairs, p, (b: unsupported), f, q.  Mates share an ordinal range with stride 2.
struct PlanEntry {
    std::string path;
    bool fastq;
    int mate;        // -1 single file; 0/1 = first/second file of an interleaved pair
    int reverse, cut;
};
static std::vector<PlanEntry> build_plan(const std::vector<Lib>& libs, int max_rd_len) {
    std::vector<PlanEntry> plan;
    for (const Lib& L : libs) {
        if (L.asm_flag != 1 && L.asm_flag != 3) continue;                           // nextValidIndex, readseq1by1.c:601
        int cut = (L.rd_len_cutoff > 0 && L.rd_len_cutoff < max_rd_len) ? L.rd_len_cutoff : max_rd_len;   // prlHashReads.c:921-928
        for (int type = 1; type <= 6; type++) {
            if (type == 4) continue;
            bool fq = (type == 2 || type == 6);
            for (size_t fi = 0; fi < L.f[type].size(); fi++) {
                if (type <= 2) {
                    plan.push_back({L.f[type][fi], fq, 0, L.reverse, cut});
                    plan.push_back({L.f[type == 1 ? 0 : 4][fi], fq, 1, L.reverse, cut});
                } else plan.push_back({L.f[type][fi], fq, -1, L.reverse, cut});
            }
        }
    }
    return plan;
}

// CPU-testable view of the host logic: "mate fastq reverse cut path" per line, in stream order
extern "C" int pgb200_plan_files(const char* cfg, char* out, size_t cap) {
    std::vector<Lib> libs;
    int max_rd_len = 0;
    scan_lib(cfg, libs, max_rd_len);
    if (!max_rd_len) max_rd_len = 100;
    std::string s = "max_rd_len " + std::to_string(max_rd_len) + "\n";
    for (const PlanEntry& e : build_plan(libs, max_rd_len))
        s += std::to_string(e.mate) + " " + std::to_string((int)e.fastq) + " " + std::to_string(e.reverse) + " " + std::to_string(e.cut) + " " + e.path + "\n";
    if (s.size() + 1 > cap) return -1;
    memcpy(out, s.c_str(), s.size() + 1);
    return 0;
}

// ------------------------------------------------------------------------------------------------ streaming a file into the engine
// Chunks are cut at record boundaries on the host (only the tail of each chunk is inspected); the GPU does the parsing.
static size_t last_record_start(const char* buf, size_t n, bool fastq) {
    // returns the offset of the last position that starts a record, such that buf[0..off) holds whole records
    if (n == 0) return 0;
    size_t p = n;
    for (;;) {
        // find previous line start
        if (p == 0) return 0;
        size_t q = p - 1;
        while (q > 0 && buf[q - 1] != '\n') q--;
        // q is a line start
        if (!fastq) { if (buf[q] == '>') return q; }
        else if (buf[q] == '@') {
            // a FASTQ header is followed two lines later by a '+' line; a quality line starting with '@' is followed
            // two lines later by a sequence line, which never starts with '+'
            const char* e1 = (const char*)memchr(buf + q, '\n', n - q);
            if (e1) {
                const char* e2 = (const char*)memchr(e1 + 1, '\n', n - (e1 + 1 - buf));
                if (e2 && (size_t)(e2 + 1 - buf) < n && e2[1] == '+') return q;
            }
        }
        p = q;
    }
}

// CPU-testable view of the chunk cutter (host logic only)
extern "C" size_t pgb200_cut_chunk(const char* buf, size_t n, int fastq) { return last_record_start(buf, n, fastq != 0); }

static double now_s() { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

// One pinned staging buffer; chunk i goes to engine i % G.  pgb200_feed_text returns when the chunk's H2D copy is done, its kernels
// keep running, so with several GPUs the copy of chunk i+1 (to the next GPU) overlaps the kernels of chunk i.
struct Feeder {
    std::vector<pgb200_engine*> engs;
    size_t next = 0;          // engine of the next chunk
    char* pin = nullptr;
    size_t cap = 0;
    ~Feeder() { if (pin) pgb200_host_free(pin); }
    void collective_flush() {
        for (auto* e : engs) if (pgb200_xchg_fence(e)) { fprintf(stderr, "pgb200: %s\n", pgb200_last_error()); exit(-1); }
        for (auto* e : engs) if (pgb200_flush(e)) { fprintf(stderr, "pgb200: %s\n", pgb200_last_error()); exit(-1); }
        fed_in_epoch = 0;
    }
    size_t fed_in_epoch = 0;
    double s_read = 0, s_feed = 0;   // wall seconds inside fread / inside pgb200_feed_text (PGB200_VERBOSE)
    // streams one file; returns number of records
    uint64_t run(const std::string& fn, bool fastq, uint64_t ord_base, uint64_t ord_stride, int reverse, int maxlen) {
        fprintf(stderr, "Import reads from file:\n %s\n", fn.c_str());
        FILE* f = fopen(fn.c_str(), "rb");
        if (!f) { fprintf(stderr, "Cannot open %s. Now exit to system...\n", fn.c_str()); exit(-1); }
        if (!pin) {
            const char* env = getenv("PGB200_CHUNK_MB");
            cap = (size_t)(env ? atoi(env) : 256) << 20;
            pin = (char*)pgb200_host_alloc(cap + 16);
            if (!pin) { fprintf(stderr, "pgb200: %s\n", pgb200_last_error()); exit(-1); }
        }
        uint64_t recs = 0;
        size_t have = 0;
        bool eof = false;
        while (!eof || have) {
            const double t_r = now_s();
            size_t got = eof ? 0 : fread(pin + have, 1, cap - have, f);
            s_read += now_s() - t_r;
            if (got == 0) eof = true;
            have += got;
            if (have == 0) break;
            size_t cut;
            if (eof) {
                while (have > 1 && pin[have - 1] == '\n' && pin[have - 2] == '\n') have--;   // trailing blank lines are harmless
                if (have == 1 && pin[0] == '\n') have = 0;
                if (have == 0) break;
                cut = have;
            } else {
                cut = last_record_start(pin, have, fastq);
                if (cut == 0) {
                    if (have == cap) { fprintf(stderr, "pgb200: a single record exceeds the %zu MB chunk\n", cap >> 20); exit(-1); }
                    continue;
                }
            }
            pgb200_engine* eng = engs[next];
            if (engs.size() > 1) {
                // several GPUs: room for this chunk's records in every arena region, and for its segment (128 per epoch over all GPUs)
                // reads in this chunk, estimated generously from the library's read length (a record is a header, the bases and,
                // for FASTQ, as many quality characters); an estimate that is too low is caught on the device (arena overflow error)
                const uint64_t L = (uint64_t)std::max(8, maxlen);
                const uint64_t upper = cut / (fastq ? L + 6 : L / 2 + 4) + 1;
                if (fed_in_epoch + engs.size() > 120 || pgb200_xchg_room(eng, upper) != 1) collective_flush();
            }
            const double t_f = now_s();
            if (pgb200_feed_text(eng, pin, cut, 0, fastq, ord_base + recs * ord_stride, ord_stride, reverse, maxlen)) {
                fprintf(stderr, "readseqInLib return error! please make sure input file is correct fastq/fasta file \n(%s)\n", pgb200_last_error());
                exit(-1);
            }
            s_feed += now_s() - t_f;
            recs += pgb200_last_chunk_records(eng);
            next = (next + 1) % engs.size();
            fed_in_epoch++;
            memmove(pin, pin + cut, have - cut);
            have -= cut;
        }
        fclose(f);
        return recs;
    }
};

// ------------------------------------------------------------------------------------------------ the stage
static void usage(int flavour127) {
    fprintf(stderr, "\npregraph -s configFile -o outputGraph [-R] [-K kmer -p n_cpu -a initMemoryAssumption -d KmerFreqCutoff]\n");
    fprintf(stderr, "  -s <string>      configFile: the config file of solexa reads\n");
    fprintf(stderr, "  -o <string>      outputGraph: prefix of output graph file name\n");
    fprintf(stderr, "  -K <int>         kmer(min 13, max %d): kmer size, [23]\n", flavour127 ? 127 : 63);
    fprintf(stderr, "  -p <int>         n_cpu: number of reference hash sets (layout parameter of the GPU engine), [8]\n");
    fprintf(stderr, "  -a <int>         initMemoryAssumption: memory assumption initialized to avoid further reallocation, unit GB, [0]\n");
    fprintf(stderr, "  -R (optional)    output extra information for resolving repeats in contig step, [NO]\n");
    fprintf(stderr, "  -d <int>         KmerFreqCutoff: kmers with frequency no larger than KmerFreqCutoff will be deleted, [0]\n");
}

// PGB200_VERBOSE: where the stage's wall time went, in milliseconds (the reference's own "Time spent" lines are whole seconds)
struct Timeline {
    double t_prev;
    std::string text;
    explicit Timeline(double t) : t_prev(t) {}
    void mark(const char* what) {
        const double t = now_s();
        char b[96];
        snprintf(b, sizeof b, "%s%s %.0f ms", text.empty() ? "" : ", ", what, (t - t_prev) * 1e3);
        text += b;
        t_prev = t;
    }
};

extern "C" int pgb200_pregraph_main(int argc, char** argv, int flavour127) {
    double t_all = now_s();
    Timeline tl(t_all);
    fprintf(stderr, "\n********************\nPregraph\n********************\n\n");
    // ---- initenv (pregraph.c:142-220)
    pgb200_params prm;
    pgb200_default_params(&prm);
    prm.flavour127 = flavour127;
    std::string cfg, prefix;
    int inp = 0, outp = 0, c;
    optind = 1;
    fprintf(stderr, "Parameters: pregraph ");
    while ((c = getopt(argc, argv, "a:s:o:K:p:d:R")) != EOF) {
        switch (c) {
            case 's': fprintf(stderr, "-s %s ", optarg); inp = 1; cfg = optarg; break;
            case 'o': fprintf(stderr, "-o %s ", optarg); outp = 1; prefix = optarg; break;
            case 'K': fprintf(stderr, "-K %s ", optarg); prm.K = atoi(optarg); break;
            case 'p': fprintf(stderr, "-p %s ", optarg); prm.P = atoi(optarg); break;
            case 'R': prm.repsTie = 1; fprintf(stderr, "-R "); break;
            case 'd': fprintf(stderr, "-d %s ", optarg); prm.D = atoi(optarg) >= 0 ? atoi(optarg) : 0; break;
            case 'a': fprintf(stderr, "-a %s ", optarg); prm.initG = atoi(optarg); break;
            default:
                if (!inp || !outp) { usage(flavour127); exit(-1); }
        }
    }
    fprintf(stderr, "\n\n");
    if (!inp || !outp) { usage(flavour127); exit(-1); }
    // ---- K fix-ups (pregraph.c:71-97)
    if (prm.K % 2 == 0) { prm.K++; fprintf(stderr, "K should be an odd number.\n"); }
    if (prm.K < 13) { prm.K = 13; fprintf(stderr, "K should not be less than 13.\n"); }
    else if (prm.K > (flavour127 ? 127 : 63)) { prm.K = flavour127 ? 127 : 63; fprintf(stderr, "K should not be greater than %d.\n", prm.K); }
    if (const char* v = getenv("PGB200_DEVICE")) prm.device = atoi(v);
    if (const char* v = getenv("PGB200_TABLE_SLOTS")) prm.table_slots = strtoull(v, nullptr, 10);
    if (const char* v = getenv("PGB200_VERBOSE")) prm.verbose = atoi(v);

    // ---- pass 1 (prlRead2HashTable)
    double t0 = now_s();
    std::vector<Lib> libs;
    int max_rd_len = 0;
    scan_lib(cfg.c_str(), libs, max_rd_len);
    if (!max_rd_len) max_rd_len = 100;   // prlHashReads.c:326-329
    prm.max_rd_len = max_rd_len;
    fprintf(stderr, "In %s, %d lib(s), maximum read length %d, maximum name length %d.\n\n", cfg.c_str(), (int)libs.size(), max_rd_len, 256);
    // ---- engines: one per GPU (PGB200_GPUS=n | all; default 1).  Pass 1 is sharded: GPU g owns bucket range g, every GPU decodes
    // and partitions the chunks dealt to it and stores the records straight into their owners' arenas (peer access).
    int n_gpus = 1;
    if (const char* v = getenv("PGB200_GPUS")) {
        int have = 0;
        cudaGetDeviceCount(&have);
        n_gpus = strcmp(v, "all") == 0 ? have : atoi(v);
        if (n_gpus < 1) n_gpus = 1;
        if (n_gpus > 16) n_gpus = 16;
        if (prm.device + n_gpus > have) { fprintf(stderr, "pgb200: PGB200_GPUS=%d but only %d GPU(s) visible\n", n_gpus, have); exit(-1); }
    }
    auto die = [&](const char* what) { fprintf(stderr, "pgb200: %s failed: %s\n", what, pgb200_last_error()); exit(-1); };
    std::vector<pgb200_engine*> engs;
    for (int g = 0; g < n_gpus; g++) {
        pgb200_params q = prm;
        q.device = prm.device + g;
        q.world = n_gpus;
        q.rank = g;
        pgb200_engine* e = pgb200_create(&q);
        if (!e) { fprintf(stderr, "pgb200: %s\n", pgb200_last_error()); exit(-1); }
        engs.push_back(e);
    }
    pgb200_engine* eng = engs[0];
    if (n_gpus > 1) {
        for (auto* e : engs) if (pgb200_xchg_setup(e, 0)) die("exchange arena");
        for (int g = 0; g < n_gpus; g++)
            for (int h = 0; h < n_gpus; h++)
                if (h != g && pgb200_xchg_import_ptr(engs[g], h, prm.device + h, pgb200_xchg_base(engs[h]))) die("peer access");
        fprintf(stderr, "[pgb200] pass 1 sharded over %d GPUs (minimizer-bucket ranges, records stored peer to peer)\n", n_gpus);
    }
    fprintf(stderr, "%d thread(s) initialized.\n", prm.P);
    tl.mark("engines");
    uint64_t ord_next = 0, n_reads = 0;
    {
        Feeder fd;
        fd.engs = engs;
        std::vector<PlanEntry> plan = build_plan(libs, max_rd_len);
        for (size_t i = 0; i < plan.size(); i++) {
            const PlanEntry& e = plan[i];
            if (e.mate == 0) {
                // mates interleave r1,r2,r1,r2 (prlHashReads.c:480-583): ordinal = base + 2*pair + mate
                uint64_t n1 = fd.run(e.path, e.fastq, ord_next, 2, e.reverse, e.cut);
                const PlanEntry& m = plan[++i];
                uint64_t n2 = fd.run(m.path, m.fastq, ord_next + 1, 2, m.reverse, m.cut);
                if (n1 != n2) { fprintf(stderr, "pgb200: mate files hold different numbers of reads (%llu vs %llu): unsupported\n", (unsigned long long)n1, (unsigned long long)n2); exit(-1); }
                ord_next += 2 * n1; n_reads += 2 * n1;
            } else {
                uint64_t n = fd.run(e.path, e.fastq, ord_next, 1, e.reverse, e.cut);
                ord_next += n; n_reads += n;
            }
        }
        if (n_gpus > 1) fd.collective_flush();
        if (prm.verbose) fprintf(stderr, "[pgb200] reading the files: %.0f ms in fread, %.0f ms in feed_text (H2D copy + launches)\n", fd.s_read * 1e3, fd.s_feed * 1e3);
    }
    pgb200_pass1_stats p1;
    memset(&p1, 0, sizeof p1);
    for (auto* e : engs) {
        pgb200_pass1_stats q;
        if (pgb200_finish_pass1(e, &q)) die("pass 1");
        p1.distinct += q.distinct; p1.instances += q.instances; p1.table_slots += q.table_slots;
        p1.ms_decode = std::max(p1.ms_decode, q.ms_decode); p1.ms_insert = std::max(p1.ms_insert, q.ms_insert);
    }
    double t1 = now_s();
    tl.mark("reads -> k-mer table");
    fprintf(stderr, "Time spent on hashing reads: %ds, %lld read(s) processed.\n", (int)(t1 - t0), (long long)n_reads);
    fprintf(stderr, "%lli node(s) allocated, %lli kmer(s) in reads, %lli kmer(s) processed.\n", (long long)p1.distinct, (long long)p1.instances, (long long)p1.instances);
    fprintf(stderr, "[pgb200] pass 1: %.3f s wall, decode %.1f ms + insert %.1f ms on the GPU%s, table %llu slots\n", t1 - t0, p1.ms_decode, p1.ms_insert,
            n_gpus > 1 ? " (slowest GPU)" : "", (unsigned long long)p1.table_slots);
    fprintf(stderr, "done hashing nodes\n");
    long long hist[256];
    memset(hist, 0, sizeof hist);
    uint64_t lin = 0, rem = 0;
    for (auto* e : engs) {
        long long h1[256];
        uint64_t l1 = 0, r1 = 0;
        if (pgb200_sweeps(e, h1, &l1, &r1)) die("sweeps");
        for (int i = 0; i < 256; i++) hist[i] += h1[i];
        lin += l1; rem += r1;
    }
    tl.mark("sweeps");
    if ((signed char)prm.D) fprintf(stderr, "%llu kmer(s) removed.\n", (unsigned long long)rem);
    fprintf(stderr, "%llu linear node(s) marked.\n", (unsigned long long)lin);
    // The graph phases walk across buckets: the shards (tables with their swept flags, packed reads) are folded into GPU 0, which
    // runs layout, tips, edges and pass 2 exactly as in the single-GPU case.
    for (int g = 1; g < n_gpus; g++) {
        if (pgb200_absorb(eng, engs[g])) die("gathering the table shards");
        pgb200_destroy(engs[g]);
        engs[g] = nullptr;
    }
    if (n_gpus > 1) tl.mark("gather shards");
    {
        std::string s;
        char b[32];
        for (int i = 1; i < 256; i++) { snprintf(b, sizeof b, "%lld\n", hist[i]); s += b; }   // freqStat, prlHashReads.c:1104-1132
        try { write_file(prefix + ".kmerFreq", s.data(), s.size()); } catch (const std::exception& ex) { fprintf(stderr, "%s\n", ex.what()); exit(-1); }
    }
    fprintf(stderr, "Time spent on pre-graph construction: %ds.\n\n", (int)(now_s() - t0));
    if (getenv("PGB200_PASS1_ONLY")) { for (auto* e : engs) pgb200_destroy(e); return 0; }

    // ---- layout + tips (removeSingleTips / removeMinorTips)
    t0 = now_s();
    if (pgb200_build_layout(eng)) die("layout");
    pgb200_graph_stats gs;
    memset(&gs, 0, sizeof gs);
    if (pgb200_remove_tips(eng, &gs)) die("tips");
    tl.mark("layout + tips");
    fprintf(stderr, "Time spent on removing tips: %ds.\n\n", (int)(now_s() - t0));
    // ---- edges (kmer2edges)
    t0 = now_s();
    // The deflate of the edge text is sequential host work (it has to be: the bytes must equal the reference's gz stream);
    // it runs on a host thread while the GPU does pass 2.
    std::string edge_text, gz_error;
    std::thread gz_thread;
    try {
        EdgeStats es;
        eng->e->build_edges(&es, &edge_text);
        fprintf(stderr, "%llu (%llu) edge(s) and %llu extra node(s) constructed.\n", (unsigned long long)es.num_ed, (unsigned long long)es.edges,
                (unsigned long long)es.extra_nodes);
        gs.num_ed = es.num_ed; gs.edges = es.edges; gs.extra_nodes = es.extra_nodes;
        const uint64_t ne = es.num_ed;
        gz_thread = std::thread([&, ne]() {
            try {
                write_edge_outputs(prefix, edge_text, prm.K, flavour127, ne);
            } catch (const std::exception& ex) { gz_error = ex.what(); }
        });
    } catch (const std::exception& ex) { fprintf(stderr, "pgb200: edges failed: %s\n", ex.what()); exit(-1); }
    tl.mark("edges");
    fprintf(stderr, "Time spent on constructing edges: %ds.\n\n", (int)(now_s() - t0));
    // ---- pass 2 (prlRead2edge)
    t0 = now_s();
    if (pgb200_read2edge(eng, prefix.c_str(), &gs)) { if (gz_thread.joinable()) gz_thread.join(); die("pass 2"); }
    fprintf(stderr, "Time spent on aligning reads: %ds.\n\n", (int)(now_s() - t0));
    tl.mark("pass 2 + its files");
    gz_thread.join();
    tl.mark("waiting for the edge file");
    if (!gz_error.empty()) { fprintf(stderr, "%s\n", gz_error.c_str()); exit(-1); }
    if (pgb200_output_vertex(eng, prefix.c_str(), &gs)) die("vertex output");
    pgb200_destroy(eng);
    tl.mark("vertex + teardown");
    if (prm.verbose) fprintf(stderr, "[pgb200] stage wall %.2f s: %s\n", now_s() - t_all, tl.text.c_str());
    fprintf(stderr, "Overall time spent on constructing pre-graph: %dm.\n\n", (int)(now_s() - t_all) / 60);
    return 0;
}

// The library's own call_pregraph has the 63-mer semantics (what dlopen / ctypes users get).  A SOAPdenovo-127mer build links
// pregraph_shim.c (-DPGB_FLAVOUR127=1) instead: the executable's definition takes precedence, so the flavour is fixed at link
// time exactly as the reference fixes it with -DMER63 / -DMER127 -- no environment variable is involved.
extern "C" int call_pregraph(int argc, char** argv) { return pgb200_pregraph_main(argc, argv, 0); }
