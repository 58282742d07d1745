// skm.cuh -- super-k-mer partition of the reads (host + device logic of the aggregated pass 1, skm.cu).
//
// Why (profiles/r01_rmw_ubench.md): a random update of an HBM-resident slot costs one DRAM read + one write-back, 1.75e10/s on a
// B200 whatever instruction performs it, and the direct insert (k_chop_insert) pays that once per k-mer INSTANCE.  Here all
// instances of a k-mer are brought together on chip first and HBM is touched once per DISTINCT k-mer:
//   1. every k-mer is assigned to a bucket by its canonical MINIMIZER (the m-mer of smallest order value among its K-m+1 m-mers,
//      strand-symmetric), so a k-mer and its reverse complement, wherever they occur, land in the same bucket;
//   2. consecutive k-mers of a read with the same bucket form a run (a super-k-mer); a run is an 8-byte record
//      {read index, first k-mer position, count <= SKM_MAX_RUN} pointing into the resident 2-bit read store -- 0.5-0.7 B per
//      instance instead of the 32 B {key, links, rank} tuple of the exchange path;
//   3. one CTA per bucket re-chops the runs and aggregates {links, cov, first rank} per distinct k-mer in a shared-memory table,
//      then merges each distinct k-mer into the (unchanged) global table once.
// The reference has no counterpart (its P threads each scan every k-mer of every read, prlHashReads.c:79-90); what has to be
// preserved is the per-k-mer result of put_kmerset/update_kmer (newhash.c:74-140, 473-528), which is a pure function of the
// multiset of instances: saturating sums and a minimum, both associative and commutative.
#pragma once
#include "kmer.cuh"
#include "table.cuh"

namespace pgb {

constexpr int SKM_MAX_RUN = 16;      // k-mers per record = lanes of a half warp: the aggregation kernel handles one record per half-warp step
constexpr int SKM_MAX_M = 15;        // minimizer length (2m bits must fit in 32)

struct SkmGeom {
    int K = 0, m = 0, w = 0;   // w = K - m + 1 m-mers per k-mer
    u32 mmask = 0;             // low 2m bits
    u32 n_buckets = 0;
};
inline SkmGeom make_skm_geom(int K, u32 n_buckets) {
    SkmGeom g;
    g.K = K;
    g.m = K - 2 < SKM_MAX_M ? K - 2 : SKM_MAX_M;
    if (g.m < 4) g.m = K < 4 ? K : 4;
    g.w = K - g.m + 1;
    g.mmask = g.m >= 16 ? 0xFFFFFFFFu : ((1u << (2 * g.m)) - 1u);
    g.n_buckets = n_buckets;
    return g;
}

// bijective 32-bit mixer (murmur3 finaliser): equal order values <=> equal canonical m-mers, so ties cannot split a k-mer
// and its reverse complement over two buckets
PG_HD u32 skm_fmix32(u32 x) {
    x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
    return x;
}
PG_HD u32 skm_order(u32 fm, u32 rm) { return skm_fmix32(fm < rm ? fm : rm); }
// minima are concentrated near 0: hash them again before the range reduction
PG_HD u32 skm_bucket(u32 minval, u32 n_buckets) {
    u32 h = (minval ^ 0x5BD1E995u) * 0x9E3779B1u;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12;
    return (u32)(((u64)h * (u64)n_buckets) >> 32);
}

// 8-byte run record: read index in its chunk | first k-mer position | LAST flag (the run ends with the read's last k-mer) | count
PG_HD u64 skm_pack(u32 read_idx, int start, int n, bool last) {
    return ((u64)read_idx << 32) | ((u64)(unsigned)start << 8) | (last ? 0x80ull : 0ull) | (u64)(unsigned)n;
}
PG_HD u32 skm_read(u64 rec) { return (u32)(rec >> 32); }
PG_HD int skm_start(u64 rec) { return (int)((rec >> 8) & 0xFFFFFF); }
PG_HD int skm_count(u64 rec) { return (int)(rec & 0x7F); }
PG_HD bool skm_last(u64 rec) { return (rec & 0x80) != 0; }

// Split one read into runs.  scratch[slot * stride] (slot < 2 * g.w) holds, per thread, the order values of the current block of w
// m-mer positions and the suffix minima of the previous block: the minimum of a window of w positions is
// min(suffix minimum of the previous block, running minimum of the current block)  (van Herk / Gil-Werman).  All threads of a warp
// are at the same position of their reads, so the once-per-block backward pass is not divergent (a per-lane rescan whenever the
// minimum leaves the window would be: some lane rescans at almost every step).
// emit(bucket, first k-mer position, count, run ends with the last k-mer of the read).
template <class Emit>
PG_HD void skm_scan_read(const SkmGeom& g, const u64* wp, int L, u32* scratch, int stride, Emit& emit) {
    const int K = g.K, m = g.m, w = g.w;
    if (L < K + 1) return;   // reads shorter than K+1 contribute nothing (prlHashReads.c:504,642)
    u32* cur_blk = scratch;                 // raw order values of the block being filled
    u32* suf_blk = scratch + w * stride;    // suffix minima of the previous block
    u32 fm = 0, rm = 0, pref = 0xFFFFFFFFu;
    int o = 0;                              // offset of m-mer position p inside its block
    u32 cur_b = 0;
    int run_start = 0, run_len = 0;
    u64 cur = wp[0];
    for (int i = 0; i < L; i++) {
        if (i && (i & 31) == 0) cur = wp[i >> 5];
        const u32 c = (u32)((cur >> (2 * (i & 31))) & 3);
        fm = ((fm << 2) | c) & g.mmask;
        rm = (rm >> 2) | ((c ^ 2u) << (2 * (m - 1)));
        if (i < m - 1) continue;
        const u32 ov = skm_order(fm, rm);
        cur_blk[o * stride] = ov;
        pref = ov < pref ? ov : pref;
        const int j = i - K + 1;            // k-mer position; its m-mers are the positions p-w+1 .. p
        if (j >= 0) {
            u32 minval = pref;
            if (o != w - 1) {
                const u32 sv = suf_blk[(o + 1) * stride];
                minval = sv < minval ? sv : minval;
            }
            const u32 b = skm_bucket(minval, g.n_buckets);
            if (run_len == 0 || b != cur_b || run_len == SKM_MAX_RUN) {
                if (run_len) emit(cur_b, run_start, run_len, false);
                cur_b = b;
                run_start = j;
                run_len = 0;
            }
            run_len++;
        }
        if (++o == w) {                     // block complete: its suffix minima serve the next w-1 windows
            u32 run = 0xFFFFFFFFu;
            for (int q = w - 1; q >= 0; q--) {
                const u32 v = cur_blk[q * stride];
                run = v < run ? v : run;
                suf_blk[q * stride] = run;
            }
            o = 0;
            pref = 0xFFFFFFFFu;
        }
    }
    if (run_len) emit(cur_b, run_start, run_len, true);
}

// bucket of ONE k-mer given as a Kmer (used by tests: every instance of a canonical k-mer must map to the same bucket)
template <int NW>
PG_HD u32 skm_bucket_of_kmer(const SkmGeom& g, const Kmer<NW>& k) {
    u32 fm = 0, rm = 0, best = 0;
    bool have = false;
    for (int t = 0; t < g.K; t++) {
        const int bit = 2 * (g.K - 1 - t);
        const u32 c = (u32)((k.w[NW - 1 - bit / 64] >> (bit % 64)) & 3);
        fm = ((fm << 2) | c) & g.mmask;
        rm = (rm >> 2) | ((c ^ 2u) << (2 * (g.m - 1)));
        if (t < g.m - 1) continue;
        const u32 o = skm_order(fm, rm);
        if (!have || o < best) { best = o; have = true; }
    }
    return skm_bucket(best, g.n_buckets);
}

// One k-mer instance taken straight from the packed read (no rolling): the LSB-first packing of the read store IS the base-reversed
// order, so the reverse complement of the k-mer at position j is  (K bases from bit 2j) XOR 0b10...,  and the forward k-mer is the
// reverse complement of that.  `buf` holds the NW+2 consecutive words of the read starting at word (2q)>>6, q = max(j-1, 0)
// (bases j-1 .. j+K); the neighbour rules are chop_read's: left/right in the CANONICAL orientation, 4 = none.
template <int NW>
struct SkmInst {
    Kmer<NW> canon;
    unsigned left, right;
};
template <int NW>
PG_HD int skm_first_word(int j) { return (2 * (j > 0 ? j - 1 : 0)) >> 6; }
template <int NW>
PG_HD SkmInst<NW> skm_instance(const KParams<NW>& kp, const u64 (&buf)[NW + 2], int j, bool has_next) {
    const int K = kp.K;
    const int q = j > 0 ? j - 1 : 0;
    const int sh = (2 * q) & 63;
    u64 x[NW + 1];   // bases q.. from bit 0
#pragma unroll
    for (int t = 0; t < NW + 1; t++) x[t] = sh ? ((buf[t] >> sh) | (buf[t + 1] << (64 - sh))) : buf[t];
    unsigned pv = 4;
    if (j > 0) {     // drop base j-1
        pv = (unsigned)(x[0] & 3);
#pragma unroll
        for (int t = 0; t < NW + 1; t++) x[t] = (x[t] >> 2) | (t + 1 < NW + 1 ? (x[t + 1] << 62) : 0ull);
    }
    unsigned cn = 4;
    if (has_next) {
        const int bit = 2 * K, wi = bit >> 6;
        u64 v = 0;
#pragma unroll
        for (int t = 0; t < NW + 1; t++)
            if (t == wi) v = x[t];
        cn = (unsigned)((v >> (bit & 63)) & 3);
    }
    Kmer<NW> rc;
#pragma unroll
    for (int t = 0; t < NW; t++) rc.w[NW - 1 - t] = (x[t] ^ 0xAAAAAAAAAAAAAAAAull) & kp.mask.w[NW - 1 - t];
    const Kmer<NW> fwd = krc_n(rc, K);
    SkmInst<NW> r;
    const bool sm = kless(fwd, rc);          // KmerSmaller(word, bal_word); tie -> rc branch
    r.canon = sm ? fwd : rc;
    r.left = sm ? pv : (cn < 4 ? (cn ^ 2u) : 4u);
    r.right = sm ? cn : (pv < 4 ? (pv ^ 2u) : 4u);
    return r;
}

// slot index of a k-mer inside a bucket's shared-memory table: one 64-bit multiply (the 5-multiply table_hash is only needed once
// per DISTINCT k-mer, for the global table)
template <int NW>
PG_HD u32 skm_slot_hash(const Kmer<NW>& k, int log2_slots) {
    u64 x = k.w[0];
#pragma unroll
    for (int i = 1; i < NW; i++) x = ((x << 29) | (x >> 35)) ^ k.w[i];
    x ^= x >> 31;
    x *= 0x9E3779B97F4A7C15ull;
    return (u32)(x >> (64 - log2_slots));
}

// Two partial results for the same k-mer -> the result of all their instances together.  `g` may be PAYLOAD_FRESH (slot claimed, no
// instance yet); `a` holds at least one instance.  Every instance has a neighbour on at least one side (reads are >= K+1 long), so
// cov counts instances (newhash.c:74-106, 123-140) and the fields simply add with their saturation; `single` survives only when the
// total is one instance, i.e. never when both sides are non-empty.
PG_HD u64 payload_merge(u64 g, u64 a) {
    if (g == PAYLOAD_FRESH) return a;
    u64 r = 0;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        unsigned l = pl_l(g, c) + pl_l(a, c), q = pl_r(g, c) + pl_r(a, c);
        r |= (u64)(l > 63 ? 63 : l) << (6 * c);
        r |= (u64)(q > 63 ? 63 : q) << (PL_R_SHIFT + 6 * c);
    }
    unsigned cv = pl_cov(g) + pl_cov(a);
    r |= (u64)(cv > 255 ? 255 : cv) << PL_COV_SHIFT;
    return r;
}

}   // namespace pgb
