// Host-side check of the super-k-mer partition logic (soapdenovo2_b200/csrc/skm.cuh), compiled with nvcc and run on the CPU by
// tests/test_skm_host.py.  Against a naive per-position restatement of SURVEY.md A.2/A.3 it verifies, for random reads:
//   * the runs of skm_scan_read tile the k-mer positions of every read exactly once, in order, each run <= SKM_MAX_RUN;
//   * every k-mer of a run has the run's bucket, and a k-mer and its reverse complement have the same bucket;
//   * the self-contained record of every run (skm_make_rec) carries exactly what its k-mers need: skm_instance_rec(record, t) yields
//     the (canonical k-mer, left, right) instance and the first-occurrence rank of every position of every run, whatever garbage
//     follows the read in its packed words;
//   * skm_pick_record maps instance q of a tile of records to (record, position) for every q;
//   * bucket ownership ranges tile the buckets for every world size;
//   * aggregating per bucket in two halves and merging with payload_merge equals applying all instances in read order.
#include "../soapdenovo2_b200/csrc/skm.cuh"
#include <cstdio>
#include <cstdlib>
#include <map>
#include <vector>
#include <array>
using namespace pgb;

static u64 rng_state = 88172645463325252ull;
static u64 rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }

template <int NW>
struct Inst { Kmer<NW> k; unsigned left, right; int j; };
template <int NW>
struct KeyLess { bool operator()(const Kmer<NW>& a, const Kmer<NW>& b) const { return kless(a, b); } };
struct Agg { u64 payload = PAYLOAD_FRESH; u64 rank = ~0ull; };

template <int NW>
static std::vector<Inst<NW>> naive(const KParams<NW>& kp, const std::vector<unsigned>& seq) {
    std::vector<Inst<NW>> out;
    const int K = kp.K, L = (int)seq.size();
    if (L < K + 1) return out;
    for (int j = 0; j + K <= L; j++) {
        Kmer<NW> f = kzero<NW>();
        for (int t = 0; t < K; t++) f = knext(f, seq[j + t], kp);
        Kmer<NW> r = krc_n(f, K);
        unsigned pv = j > 0 ? seq[j - 1] : 4u, nx = j + K < L ? seq[j + K] : 4u;
        Inst<NW> in;
        in.j = j;
        if (kless(f, r)) { in.k = f; in.left = pv; in.right = nx; }
        else { in.k = r; in.left = nx < 4 ? (nx ^ 2u) : 4u; in.right = pv < 4 ? (pv ^ 2u) : 4u; }
        out.push_back(in);
    }
    return out;
}

struct Run { u32 b; int start, n; bool last; };
struct EmitRuns {
    std::vector<Run>* v;
    void operator()(u32 b, int s, int n, bool last) { v->push_back(Run{b, s, n, last}); }
};

template <int NW>
static int run_case(int K, int n_reads, int maxlen, u32 n_buckets, bool low_complexity) {
    KParams<NW> kp = make_kparams<NW>(K);
    SkmGeom g = make_skm_geom(K, n_buckets);
    const int W64 = (maxlen + 31) / 32;
    std::vector<u32> ring(g.w);
    typedef std::map<Kmer<NW>, Agg, KeyLess<NW>> Map;
    Map direct, half[2], merged;
    std::map<u32, u64> bucket_load;
    u64 n_inst = 0, n_runs = 0;
    int errors = 0;
    // a pool of reads drawn from a small "genome" so that k-mers repeat (and saturate) across reads
    const int G = low_complexity ? 150 : 20000;   // 150: k-mers repeat > 255 times (cov and links saturate)
    std::vector<unsigned> genome(G);
    for (auto& c : genome) c = low_complexity ? (unsigned)((rnd() % 8) < 6 ? 0 : rnd() & 3) : (unsigned)(rnd() & 3);
    for (int r = 0; r < n_reads; r++) {
        int L = (int)(rnd() % (maxlen + 1));
        if (r % 7 == 0) L = maxlen;
        if (r % 11 == 0) L = K + 1;
        if (r % 13 == 0) L = K;           // too short: no instance
        std::vector<unsigned> seq(L);
        int g0 = (int)(rnd() % G);
        bool rev = rnd() & 1;
        for (int i = 0; i < L; i++) {
            unsigned c = genome[(g0 + i) % G];
            if (rnd() % 200 == 0) c = (unsigned)(rnd() & 3);
            seq[i] = c;
        }
        if (rev) { std::vector<unsigned> t(L); for (int i = 0; i < L; i++) t[i] = seq[L - 1 - i] ^ 2u; seq = t; }
        std::vector<u64> words(W64 + 1, 0);
        for (int i = 0; i < L; i++) words[i >> 5] |= (u64)seq[i] << (2 * (i & 31));
        std::vector<Inst<NW>> ref = naive<NW>(kp, seq);
        std::vector<Run> runs;
        EmitRuns em{&runs};
        skm_scan_read(g, words.data(), L, ring.data(), 1, em);
        // tiling
        int next = 0;
        for (auto& ru : runs) {
            const u32 sp = skm_side_pack(ru.b, ru.n, ru.last);
            if (skm_side_bucket(sp) != ru.b || skm_side_n(sp) != ru.n || skm_side_last(sp) != ru.last) errors++;
            if (ru.last != (ru.start + ru.n == (int)ref.size())) { errors++; if (errors < 5) printf("bad last flag read %d\n", r); }
            if (ru.start != next || ru.n < 1 || ru.n > SKM_MAX_RUN || ru.b >= n_buckets) { errors++; if (errors < 5) printf("bad run read %d start %d n %d (expected start %d)\n", r, ru.start, ru.n, next); }
            next = ru.start + ru.n;
        }
        if (next != (int)ref.size()) { errors++; if (errors < 5) printf("runs cover %d of %zu positions (read %d, L %d)\n", next, ref.size(), r, L); }
        const u64 rank_base = (u64)r << 16;
        for (auto& in : ref) {
            Agg& a = direct[in.k];
            a.payload = payload_apply(a.payload, in.left, in.right);
            u64 rk = rank_base | (u64)in.j;
            if (rk < a.rank) a.rank = rk;
        }
        // words past the read: garbage must not matter (the record masks what it copies)
        std::vector<u64> dirty(words);
        for (int i = L; i < 32 * (int)dirty.size(); i++) dirty[i >> 5] |= (u64)(rnd() & 3) << (2 * (i & 31));
        const u64 ordinal = 1000ull + (u64)r * 3ull;
        std::vector<u32> P;   // lane packing: prefix sums of the runs' k-mer counts, padded to a power of two
        for (auto& ru : runs) {
            const SkmRec<NW> rec = skm_make_rec<NW>(K, dirty.data(), W64, ordinal, ru.start, ru.n, ru.last);
            const SkmRec<NW> rec2 = skm_make_rec<NW>(K, words.data(), W64 + 1, ordinal, ru.start, ru.n, ru.last);
            for (int x = 0; x < NW + 2; x++) if (rec.w[x] != rec2.w[x]) { errors++; if (errors < 5) printf("record depends on bytes past the read (read %d run at %d)\n", r, ru.start); break; }
            const u64 h = rec.w[0];
            if (skm_rec_n(h) != ru.n || skm_rec_start(h) != ru.start || skm_rec_last(h) != ru.last || skm_rec_ordinal(h) != ordinal || skm_rec_has_prev(h) != (ru.start > 0)) {
                errors++;
                if (errors < 5) printf("header round trip failed read %d\n", r);
            }
            u64 x[NW + 1];
            for (int i = 0; i < NW + 1; i++) x[i] = rec.w[1 + i];
            for (int t = 0; t < ru.n; t++) {
                if (ru.start + t >= (int)ref.size()) break;
                const int j = ru.start + t;
                SkmInst<NW> si = skm_instance_rec<NW>(kp, h, x, t);
                Inst<NW> a{si.canon, si.left, si.right, j};
                const Inst<NW>& b = ref[ru.start + t];
                if (!keq(a.k, b.k) || a.left != b.left || a.right != b.right || skm_rec_rank(h, t) != ((ordinal << 16) | (u64)j)) {
                    errors++;
                    if (errors < 5) printf("instance mismatch read %d pos %d: left %u/%u right %u/%u keq %d\n", r, ru.start + t, a.left, b.left, a.right, b.right, (int)keq(a.k, b.k));
                }
                u32 bk = skm_bucket_of_kmer<NW>(g, a.k), bk2 = skm_bucket_of_kmer<NW>(g, krc_n(a.k, K));
                if (bk != ru.b || bk2 != ru.b) { errors++; if (errors < 5) printf("bucket mismatch read %d pos %d: run %u kmer %u rc %u\n", r, ru.start + t, ru.b, bk, bk2); }
                Agg& ag = half[r & 1][a.k];
                ag.payload = payload_apply(ag.payload, a.left, a.right);
                u64 rk = rank_base | (u64)a.j;
                if (rk < ag.rank) ag.rank = rk;
                bucket_load[ru.b]++;
                n_inst++;
            }
            n_runs++;
        }
        if (!runs.empty() && runs.size() <= 64) {
            P.assign(65, 0);
            u32 acc = 0;
            for (size_t i = 0; i < 64; i++) { P[i] = acc; if (i < runs.size()) acc += (u32)runs[i].n; }
            P[64] = acc;
            u32 q = 0;
            for (size_t i = 0; i < runs.size(); i++)
                for (int t = 0; t < runs[i].n; t++, q++) {
                    const int rr = skm_pick_record(P.data(), 64, q);
                    if (rr != (int)i || (int)(q - P[rr]) != t) { errors++; if (errors < 5) printf("lane packing: instance %u -> record %d (expected %zu)\n", q, rr, i); }
                }
        }
    }
    for (int h = 0; h < 2; h++)
        for (auto& kv : half[h]) {
            Agg& m = merged[kv.first];
            m.payload = payload_merge(m.payload, kv.second.payload);
            if (kv.second.rank < m.rank) m.rank = kv.second.rank;
        }
    if (merged.size() != direct.size()) { errors++; printf("distinct %zu vs %zu\n", merged.size(), direct.size()); }
    u64 saturated = 0;
    for (auto& kv : direct) {
        auto it = merged.find(kv.first);
        if (it == merged.end() || it->second.payload != kv.second.payload || it->second.rank != kv.second.rank) {
            errors++;
            if (errors < 5) printf("aggregate mismatch: payload %llx vs %llx rank %llx vs %llx\n", it == merged.end() ? 0ull : it->second.payload, kv.second.payload,
                                   it == merged.end() ? 0ull : it->second.rank, kv.second.rank);
        }
        if (pl_cov(kv.second.payload) == 255) saturated++;
    }
    u64 mx = 0;
    for (auto& kv : bucket_load) if (kv.second > mx) mx = kv.second;
    printf("K=%d m=%d w=%d reads=%d instances=%llu runs=%llu distinct=%zu saturated=%llu buckets_used=%zu max_bucket=%llu errors=%d\n", K, g.m, g.w, n_reads,
           n_inst, n_runs, direct.size(), saturated, bucket_load.size(), mx, errors);
    return errors;
}

static int owner_case(u32 B, int world) {
    int errors = 0;
    u32 prev_hi = 0;
    for (int o = 0; o < world; o++) {
        const u32 lo = skm_owner_lo(B, world, o), hi = o + 1 < world ? skm_owner_lo(B, world, o + 1) : B;
        if (lo != prev_hi || hi < lo) errors++;
        prev_hi = hi;
        for (u32 b = lo; b < hi; b++) if (skm_owner_of(B, world, b) != o) { errors++; break; }
    }
    if (prev_hi != B) errors++;
    const SkmArenaGeom a = make_skm_arena_geom(world, B, 7, 1000, 4);
    if (a.bo_max < (B + world - 1) / world || a.off_recs % 256 || a.half_bytes < a.off_recs + (u64)world * 1000 * 32) errors++;
    return errors;
}

int main(int argc, char** argv) {
    if (argc > 1) rng_state ^= strtoull(argv[1], nullptr, 0) * 0x9E3779B97F4A7C15ull;   // other random reads (fuzzing); default: the fixed set
    int e = 0;
    {
        int oe = 0;
        for (u32 B : {1u, 2u, 7u, 10u, 1024u, 1000u, 262144u, 3u << 16})
            for (int w = 1; w <= 16; w++) if (B >= (u32)w) oe += owner_case(B, w);
        printf("ownership ranges: errors=%d\n", oe);
        e += oe;
    }
    e += run_case<2>(13, 3000, 60, 64, false);
    e += run_case<2>(21, 3000, 100, 256, false);
    e += run_case<2>(31, 2000, 150, 1000, false);
    e += run_case<2>(63, 2000, 150, 4096, false);
    e += run_case<2>(63, 1500, 150, 37, true);
    e += run_case<2>(63, 500, 250, 4096, false);
    e += run_case<4>(65, 1500, 150, 512, false);
    e += run_case<4>(127, 1500, 150, 4096, false);
    e += run_case<4>(127, 800, 250, 100, true);
    printf(e ? "FAILED %d\n" : "ALL OK\n", e);
    return e ? 1 : 0;
}
