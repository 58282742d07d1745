// table.cuh -- the GPU-resident open-addressing k-mer table ("ktab").
//
// Replaces the reference's KmerSet (standardPregraph/newhash.c, inc/newhash.h:77-113) on the hot path:
//   reference                                       here
//   P prime-sized sets, one thread per set          ONE power-of-two table in HBM, every thread inserts anywhere
//   2-bit flag array + 24/40 B kmer_t               32 B (K<=63) / 64 B (K<=127) sector-aligned slots, no flag array
//   put_kmerset: probe, memset+init | update        128-bit atomicCAS claim of the key, then ONE 128-bit atomicCAS on
//   (newhash.c:473-528, 74-140)                     {payload, rank}: saturating link/coverage counters + min(first-seen rank)
// The payload word has EXACTLY the reference's kmer_t bit layout (newhash.h:77-102) so that later phases and the
// parity dump can use it verbatim:
//   bits  0..23  l_links (4 x 6 bit, index = base code)   bits 24..31  cov (8 bit, saturates at 255)
//   bits 32..55  r_links (4 x 6 bit)                      bit 56 linear, 57 deleted, 58 checked, 59 single,
//   bits 60..61  twin                                     bits 62..63 inEdge
// `aux` holds, in turn: the first-occurrence rank (pass 1) -> the position in reference iteration order (after layout)
// Slot position in ktab is private to the engine; the reference's FCFS layout is rebuilt from the ranks (layout.cu).
#pragma once
#include "kmer.cuh"

namespace pgb {

struct alignas(16) U128 {
    u64 a, b;
};

constexpr u64 EMPTY64 = ~0ull;
constexpr u64 PAYLOAD_FRESH = ~0ull;          // memset(0xFF) state: "key claimed but no instance applied yet"
constexpr u64 BUSY_BIT = 1ull << 63;           // NW==4 claim protocol (w[0] uses at most 62 bits for K <= 127)

constexpr int PL_COV_SHIFT = 24;
constexpr int PL_R_SHIFT = 32;
constexpr u64 PL_LINEAR = 1ull << 56;
constexpr u64 PL_DELETED = 1ull << 57;
constexpr u64 PL_CHECKED = 1ull << 58;
constexpr u64 PL_SINGLE = 1ull << 59;
constexpr int PL_TWIN_SHIFT = 60;
constexpr int PL_INEDGE_SHIFT = 62;
constexpr u64 PL_LLINKS_MASK = 0xFFFFFFull;
constexpr u64 PL_RLINKS_MASK = 0xFFFFFFull << 32;

PG_HD unsigned pl_l(u64 p, int c) { return (unsigned)((p >> (6 * c)) & 63); }
PG_HD unsigned pl_r(u64 p, int c) { return (unsigned)((p >> (PL_R_SHIFT + 6 * c)) & 63); }
PG_HD unsigned pl_cov(u64 p) { return (unsigned)((p >> PL_COV_SHIFT) & 255); }
PG_HD int pl_nl(u64 p) { return (pl_l(p, 0) > 0) + (pl_l(p, 1) > 0) + (pl_l(p, 2) > 0) + (pl_l(p, 3) > 0); }
PG_HD int pl_nr(u64 p) { return (pl_r(p, 0) > 0) + (pl_r(p, 1) > 0) + (pl_r(p, 2) > 0) + (pl_r(p, 3) > 0); }
PG_HD int pl_first_l(u64 p) { for (int c = 0; c < 4; c++) if (pl_l(p, c)) return c; return 4; }
PG_HD int pl_first_r(u64 p) { for (int c = 0; c < 4; c++) if (pl_r(p, c)) return c; return 4; }
PG_HD u64 pl_clear_l(u64 p, int c) { return p & ~(63ull << (6 * c)); }
PG_HD u64 pl_clear_r(u64 p, int c) { return p & ~(63ull << (PL_R_SHIFT + 6 * c)); }
PG_HD unsigned pl_twin(u64 p) { return (unsigned)((p >> PL_TWIN_SHIFT) & 3); }
PG_HD unsigned pl_inedge(u64 p) { return (unsigned)((p >> PL_INEDGE_SHIFT) & 3); }

// One instance applied to a payload word: set_new_kmer (newhash.c:123-140) when fresh, else update_kmer (:74-106)
// + `single = 0` (newhash.c:509).  left/right are base codes 0..3 or 4 (= none).  Branch-free on the update path (the lanes of a
// warp hold different neighbours and different counters): one increment word, with the increments of saturated fields removed.
PG_HD u64 payload_apply(u64 p, unsigned left, unsigned right) {
    const u64 il = left < 4 ? 1ull << (6 * left) : 0ull;
    const u64 ir = right < 4 ? 1ull << (PL_R_SHIFT + 6 * right) : 0ull;
    if (p == PAYLOAD_FRESH) return (1ull << PL_COV_SHIFT) | PL_SINGLE | il | ir;
    const u64 ml = il * 63ull, mr = ir * 63ull, mc = 255ull << PL_COV_SHIFT;      // the fields the increments land in
    u64 inc = ((p & ml) == ml ? 0ull : il) | ((p & mr) == mr ? 0ull : ir);
    if ((il | ir) != 0ull && (p & mc) != mc) inc |= 1ull << PL_COV_SHIFT;
    return (p + inc) & ~PL_SINGLE;
}

// The per-entry sweeps of pass 1's end (K4): delow (thread_delow: zero every link counter <= D, deleted = 1 if nothing is left), mark
// (thread_mark: linear = 1 iff exactly one non-zero left and one non-zero right link, NO deleted check there), histogram of cov.
// Used by k_sweep (pass1.cu) and, fused, by the aggregation kernel's flush (skm.cu).  The linear flag is recomputed, not inherited.
#if defined(__CUDACC__)
__device__ __forceinline__ u64 sweep_payload(u64 p, int D, unsigned& rem, unsigned& lin, unsigned* s_hist) {
    p &= ~PL_LINEAR;
    if (D > 0) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
            unsigned l = pl_l(p, c), r = pl_r(p, c);
            if (l > 0 && l <= (unsigned)D) p = pl_clear_l(p, c);
            if (r > 0 && r <= (unsigned)D) p = pl_clear_r(p, c);
        }
        if ((p & PL_LLINKS_MASK) == 0 && (p & PL_RLINKS_MASK) == 0) { p |= PL_DELETED; rem++; }
    }
    atomicAdd(&s_hist[pl_cov(p)], 1u);
    if (pl_nl(p) == 1 && pl_nr(p) == 1) { p |= PL_LINEAR; lin++; }
    return p;
}
#endif

template <int NW>
struct Slot;
template <>
struct alignas(32) Slot<2> {
    u64 key[2];
    u64 payload;
    u64 aux;
};
template <>
struct alignas(64) Slot<4> {
    u64 key[4];
    u64 payload;
    u64 aux;
    u64 pad[2];
};

template <int NW>
struct Table {
    Slot<NW>* slots;
    u64 mask;   // capacity - 1 (capacity is a power of two)
};

#if defined(__CUDACC__)

// Table loads are STRONG (relaxed, gpu scope): a weak ld.global may legally be hoisted out of a polling loop by ptxas
// (observed: the publish-wait below was turned into a single load, producing duplicate 256-bit keys on the GPU).
PG_D U128 ldcg128(const void* p) {
    U128 r;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(r.a), "=l"(r.b) : "l"(p) : "memory");
    return r;
}
struct alignas(32) U256 {
    u64 a, b, c, d;
};
// one 32-byte sector in one request (LDG.E.256 on sm_100a): key + payload + aux of a K<=63 slot
PG_D U256 ld256(const void* p) {
    U256 r;
    asm volatile("ld.relaxed.gpu.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(r.a), "=l"(r.b), "=l"(r.c), "=l"(r.d) : "l"(p) : "memory");
    return r;
}
PG_D u64 ldcg64(const void* p) {
    u64 r;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(r) : "l"(p) : "memory");
    return r;
}
PG_D u64 ldacq64(const void* p) {
    u64 r;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(r) : "l"(p) : "memory");
    return r;
}
PG_D void stcg128(void* p, U128 v) { asm volatile("st.global.cg.v2.u64 [%0], {%1,%2};" ::"l"(p), "l"(v.a), "l"(v.b) : "memory"); }

// Find-or-claim the slot of key k.  Returns the slot index; *claimed = true if this call inserted the key.
template <int NW>
PG_D u64 table_find_or_claim(const Table<NW>& t, const Kmer<NW>& k, bool* claimed);

template <>
PG_D u64 table_find_or_claim<2>(const Table<2>& t, const Kmer<2>& k, bool* claimed) {
    u64 idx = table_hash(k) & t.mask;
    *claimed = false;
    for (;;) {
        Slot<2>* s = t.slots + idx;
        U128 cur = ldcg128(s->key);
        if (cur.a == k.w[0] && cur.b == k.w[1]) return idx;
        if (cur.a == EMPTY64 && cur.b == EMPTY64) {
            U128 want{k.w[0], k.w[1]}, empty{EMPTY64, EMPTY64};
            U128 old = atomicCAS(reinterpret_cast<U128*>(s->key), empty, want);   // ATOMG.E.CAS.128
            if (old.a == EMPTY64 && old.b == EMPTY64) { *claimed = true; return idx; }
            if (old.a == k.w[0] && old.b == k.w[1]) return idx;
        }
        idx = (idx + 1) & t.mask;
    }
}

// 256-bit keys: claim {w0|BUSY, w1} with one 128-bit CAS, publish {w2,w3}, then clear BUSY.  A reader only has to
// wait when the upper halves match (i.e. it is very likely looking at its own key being published).
template <>
PG_D u64 table_find_or_claim<4>(const Table<4>& t, const Kmer<4>& k, bool* claimed) {
    u64 idx = table_hash(k) & t.mask;
    *claimed = false;
    for (;;) {
        Slot<4>* s = t.slots + idx;
        U128 hi = ldcg128(&s->key[0]);
        if (hi.a == EMPTY64 && hi.b == EMPTY64) {
            U128 want{k.w[0] | BUSY_BIT, k.w[1]}, empty{EMPTY64, EMPTY64};
            U128 old = atomicCAS(reinterpret_cast<U128*>(&s->key[0]), empty, want);
            if (old.a == EMPTY64 && old.b == EMPTY64) {
                stcg128(&s->key[2], U128{k.w[2], k.w[3]});
                // publish with a release store (orders the lower half before it) instead of __threadfence() + exchange: the fence
                // compiles to MEMBAR.SC + CCTL.IVALL, paid by every new 256-bit key (a third of the instances at K=127)
                asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(&s->key[0]), "l"(k.w[0]) : "memory");
                *claimed = true;
                return idx;
            }
            hi = old;
        }
        if ((hi.a & ~BUSY_BIT) == k.w[0] && hi.b == k.w[1]) {
            // wait (acquire) until the lower half has been published, then compare it
            do { hi.a = ldacq64(&s->key[0]); } while (hi.a & BUSY_BIT);
            U128 lo = ldcg128(&s->key[2]);
            if (hi.a == k.w[0] && lo.a == k.w[2] && lo.b == k.w[3]) return idx;
        }
        idx = (idx + 1) & t.mask;
    }
}

// Read-only lookup (frozen table).  Returns slot index or ~0 when absent (search_kmerset, newhash.c:277-318).
template <int NW>
PG_D u64 table_find(const Table<NW>& t, const Kmer<NW>& k) {
    u64 idx = table_hash(k) & t.mask;
    for (;;) {
        const Slot<NW>* s = t.slots + idx;
        U128 hi = ldcg128(&s->key[0]);
        if (hi.a == EMPTY64 && hi.b == EMPTY64) return ~0ull;
        bool m = hi.a == k.w[0] && hi.b == k.w[1];
        if (NW == 4 && m) {
            U128 lo = ldcg128(&s->key[2]);
            m = lo.a == k.w[2] && lo.b == k.w[3];
        }
        if (m) return idx;
        idx = (idx + 1) & t.mask;
    }
}

// Apply one k-mer instance: saturating counters through a 64-bit CAS loop on `payload`, first-occurrence rank through an
// atomicMin on `aux` that is only issued when this instance is an earlier occurrence than the one on record.  EVERY writer of a
// slot's {payload, aux} words (per-instance insert, aggregated merge, spills) uses these same two 64-bit atomics: atomics of
// different widths on overlapping words are not guaranteed to be atomic with respect to each other.
// Order-independent: the final {payload, rank} is a pure function of the multiset of instances.
template <int NW>
PG_D void slot_apply(Slot<NW>* s, u64 cur_payload, u64 cur_rank, unsigned left, unsigned right, u64 rank) {
    for (;;) {
        u64 nxt = payload_apply(cur_payload, left, right);
        if (nxt == cur_payload) break;                       // saturated: read-only
        u64 old = atomicCAS(&s->payload, cur_payload, nxt);
        if (old == cur_payload) break;
        cur_payload = old;
    }
    if (rank < cur_rank) atomicMin(&s->aux, rank);
}

// Fused find-or-claim + apply for one k-mer instance.  Returns true if this call inserted the key.
template <int NW>
PG_D bool table_insert(const Table<NW>& t, const Kmer<NW>& k, unsigned left, unsigned right, u64 rank, u64* idx_out = nullptr) {
    bool claimed;
    u64 idx = table_find_or_claim(t, k, &claimed);
    if (idx_out) *idx_out = idx;
    Slot<NW>* s = t.slots + idx;
    U128 cur = claimed ? U128{PAYLOAD_FRESH, EMPTY64} : ldcg128(&s->payload);
    slot_apply(s, cur.a, cur.b, left, right, rank);
    return claimed;
}
// K <= 63: the whole 32 B slot arrives with ONE 256-bit load per probe, so a hit needs no second read before the CAS and
// a fresh claim knows the {payload, rank} it will find (the memset state).
template <>
PG_D bool table_insert<2>(const Table<2>& t, const Kmer<2>& k, unsigned left, unsigned right, u64 rank, u64* idx_out) {
    u64 idx = table_hash(k) & t.mask;
    bool claimed = false;
    U128 cur;
    Slot<2>* s;
    for (;;) {
        s = t.slots + idx;
        U256 v = ld256(s);
        if (v.a == k.w[0] && v.b == k.w[1]) { cur.a = v.c; cur.b = v.d; break; }
        if (v.a == EMPTY64 && v.b == EMPTY64) {
            U128 want{k.w[0], k.w[1]}, empty{EMPTY64, EMPTY64};
            U128 old = atomicCAS(reinterpret_cast<U128*>(s->key), empty, want);
            if (old.a == EMPTY64 && old.b == EMPTY64) { claimed = true; cur.a = PAYLOAD_FRESH; cur.b = EMPTY64; break; }
            if (old.a == k.w[0] && old.b == k.w[1]) { cur = ldcg128(&s->payload); break; }
        }
        idx = (idx + 1) & t.mask;
    }
    if (idx_out) *idx_out = idx;
    slot_apply(s, cur.a, cur.b, left, right, rank);
    return claimed;
}

template <int NW>
PG_D Kmer<NW> slot_key(const Slot<NW>* s) {
    Kmer<NW> k;
#pragma unroll
    for (int i = 0; i < NW; i++) k.w[i] = s->key[i];
    if (NW == 4) k.w[0] &= ~BUSY_BIT;
    return k;
}
template <int NW>
PG_D bool slot_occupied(const Slot<NW>* s) { return !(s->key[0] == EMPTY64 && s->key[1] == EMPTY64); }

#endif   // __CUDACC__

}   // namespace pgb
