// chop.cuh -- K2: roll the canonical k-mer along one 2-bit packed read (chopKmer4read, prlHashReads.c:163-259).
// sink(canon, left, right, j) is called for every k-mer position j; left/right are the neighbour base codes in the
// CANONICAL orientation (4 = none), SURVEY.md A.2.
#pragma once
#include "kmer.cuh"

namespace pgb {

template <int NW, class Sink>
__device__ __forceinline__ void chop_read(const KParams<NW>& kp, const u64* __restrict__ wp, int L, Sink& sink) {
    const int K = kp.K;
    Kmer<NW> fwd = kzero<NW>(), rc = kzero<NW>();
    u64 cur = wp[0];
    unsigned c = (unsigned)(cur & 3);          // base i
    for (int i = 0; i < L; i++) {
        // look ahead one base (needed as the right neighbour of the k-mer ending at i)
        unsigned cn = 4;
        if (i + 1 < L) {
            if (((i + 1) & 31) == 0) cur = wp[(i + 1) >> 5];
            cn = (unsigned)((cur >> (2 * ((i + 1) & 31))) & 3);
        }
        unsigned dropped = kfirst(fwd, kp);    // base j-1 (valid when j >= 1)
        fwd = knext(fwd, c, kp);
        rc = kprev(rc, c ^ 2u, kp);
        int j = i - K + 1;
        if (j >= 0) {
            unsigned pv = j > 0 ? dropped : 4u;
            bool sm = kless(fwd, rc);          // KmerSmaller(word, bal_word); tie -> rc branch
            unsigned left = sm ? pv : (cn < 4 ? (cn ^ 2u) : 4u);
            unsigned right = sm ? cn : (pv < 4 ? (pv ^ 2u) : 4u);
            sink(sm ? fwd : rc, left, right, j);
        }
        c = cn;
    }
}


// tuple meta word: ordinal << 22 | position << 6 | left << 3 | right ; first-occurrence rank == meta >> 6
PG_HD u64 tuple_meta(u64 ord, int j, unsigned left, unsigned right) { return (ord << 22) | ((u64)j << 6) | (left << 3) | right; }

}   // namespace pgb
