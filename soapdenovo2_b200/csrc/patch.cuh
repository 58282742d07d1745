// patch.cuh -- the (K+1)-mer "patch" table: length-1 edges keyed by their canonical (K+1)-mer -> {edge id, twin}.
// Reference: KmerSetsPatch, node2edge.c:371-376, 481-541 (insert), prlRead2path.c:558-596 (search1kmerPlus).
// The reference only ever SEARCHES these sets (never iterates them), so their layout is free: one power-of-two table.
// A (K+1)-mer can use every bit of the key words (K = 63 -> 128 bits), so emptiness lives in `val`, not in the key.
#pragma once
#include "kmer.cuh"

namespace pgb {

template <int NW>
struct PatchSlot {
    u64 key[NW];
    u64 val;   // 0 = empty; else PATCH_VALID | twin << 32 | edge id
};
constexpr u64 PATCH_BUSY = 1ull << 62;
constexpr u64 PATCH_VALID = 1ull << 63;

#if defined(__CUDACC__)
template <int NW>
__device__ void patch_insert(PatchSlot<NW>* pt, u64 mask, const Kmer<NW>& key, u64 val) {
    u64 idx = table_hash(key) & mask;
    for (;;) {
        if (atomicCAS(&pt[idx].val, 0ull, PATCH_BUSY) == 0ull) {
            for (int i = 0; i < NW; i++) pt[idx].key[i] = key.w[i];
            __threadfence();
            atomicExch(&pt[idx].val, val);
            return;
        }
        idx = (idx + 1) & mask;
    }
}
// frozen-table lookup; returns val or 0 when absent
template <int NW>
__device__ u64 patch_find(const PatchSlot<NW>* pt, u64 mask, const Kmer<NW>& key) {
    u64 idx = table_hash(key) & mask;
    for (;;) {
        u64 v = pt[idx].val;
        if (v == 0) return 0;
        bool m = true;
        for (int i = 0; i < NW; i++) m = m && (pt[idx].key[i] == key.w[i]);
        if (m) return v;
        idx = (idx + 1) & mask;
    }
}
#endif
}   // namespace pgb
