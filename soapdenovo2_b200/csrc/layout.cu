// layout.cu -- K3b: rebuild the REFERENCE's table layout, i.e. the iteration order "set 0 slot 0..size-1, set 1 ..."
// that defines vertex order, edge ids, tip-clipping order and therefore every output file except .kmerFreq
// (SURVEY.md facts 1-2, A.4-A.5).
//
// Reference behaviour being reproduced (standardPregraph/):
//   * set of a k-mer       = crc(kmer) % P                               prlHashReads.c:83, hashFunction.c:155-158
//   * home slot in the set = kmer % prime size (or modular())            newhash.c:490-492 / 36-47
//   * slot                 = first free slot from home at insertion time, insertion order = first occurrence in the read stream
//                            (FCFS linear probing)                        newhash.c:473-528
// FCFS linear probing == priority linear probing with priority = first-occurrence rank, and that is order-independent:
// every distinct k-mer walks from its home slot doing atomicMin(slot, rank); whoever holds the larger rank moves on
// (carrying the displaced rank if it won).  The fixed point is exactly the sequential layout.  Ranks are unique
// (read ordinal << 16 | position), so the 64-bit slot word needs no payload; a second pass lets every k-mer find its
// own rank again, a third turns the rank table into "position -> ktab slot", and one scan compacts it into order[].
#include "engine_impl.cuh"
#include "scan.cuh"

namespace pgb {

struct RefGeom {
    const u64* set_size;   // [P]
    const u64* set_base;   // [P]
    int P;
    bool flavour127;
};

template <int NW>
__global__ void __launch_bounds__(256) k_count_sets(Table<NW> tab, int P, u64* set_count) {
    __shared__ unsigned s_cnt[256];
    s_cnt[threadIdx.x] = 0;
    __syncthreads();
    u64 n = tab.mask + 1;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const Slot<NW>* s = tab.slots + i;
        if (!slot_occupied(s)) continue;
        atomicAdd(&s_cnt[crc_hash(slot_key(s)) % (u64)P], 1u);
    }
    __syncthreads();
    if ((int)threadIdx.x < P && s_cnt[threadIdx.x]) atomicAdd(&set_count[threadIdx.x], (u64)s_cnt[threadIdx.x]);
}

template <int NW>
__global__ void __launch_bounds__(256) k_layout_place(Table<NW> tab, RefGeom g, u64* R) {
    u64 n = tab.mask + 1;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const Slot<NW>* s = tab.slots + i;
        if (!slot_occupied(s)) continue;
        Kmer<NW> k = slot_key(s);
        int set = (int)(crc_hash(k) % (u64)g.P);
        u64 size = g.set_size[set];
        u64* base = R + g.set_base[set];
        u64 pos = ref_home(k, size, g.flavour127);
        u64 r = s->aux;   // first-occurrence rank
        for (;;) {
            u64 old = atomicMin(&base[pos], r);
            if (old == EMPTY64) break;       // took a free slot
            if (old > r) r = old;            // displaced a later arrival: carry it on
            if (++pos == size) pos = 0;
        }
    }
}

template <int NW>
__global__ void __launch_bounds__(256) k_layout_resolve(Table<NW> tab, RefGeom g, const u64* R, u64* gpos_out) {
    u64 n = tab.mask + 1;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const Slot<NW>* s = tab.slots + i;
        if (!slot_occupied(s)) { continue; }
        Kmer<NW> k = slot_key(s);
        int set = (int)(crc_hash(k) % (u64)g.P);
        u64 size = g.set_size[set];
        const u64* base = R + g.set_base[set];
        u64 pos = ref_home(k, size, g.flavour127);
        u64 r = s->aux;
        while (base[pos] != r) { if (++pos == size) pos = 0; }
        gpos_out[i] = g.set_base[set] + pos;
    }
}

template <int NW>
__global__ void __launch_bounds__(256) k_layout_fill(Table<NW> tab, const u64* gpos, u64* R) {
    u64 n = tab.mask + 1;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        if (!slot_occupied(tab.slots + i)) continue;
        R[gpos[i]] = i;   // every resolve has finished (kernel boundary): the rank is no longer needed
    }
}

struct OccIn {
    const u64* R;
    __device__ u64 operator()(u64 i) const { return R[i] != EMPTY64; }
};
template <int NW>
struct OrderOut {
    const u64* R;
    u64* order;
    Slot<NW>* slots;
    __device__ void operator()(u64 i, u64 prefix, u64 v) const {
        if (!v) return;
        u64 slot = R[i];
        order[prefix] = slot;
        slots[slot].aux = prefix;   // aux now = index in reference iteration order
    }
};

// growth sequence of a dynamic reference set that ends up holding `count` keys (newhash.c:200-233, 340-381)
static u64 ref_dynamic_final_size(u64 count) {
    u64 size = ref_next_prime(1024);
    float lf = 0.77f;
    u64 mx = (u64)(size * lf);
    // put #c (1-based) triggers growth when (c-1)+1 > max
    while (count > mx) {
        u64 n = size;
        u64 at = mx + 1;   // the put that triggered sees count == mx
        do {
            if (n < 0xFFFFFFFULL) n <<= 1; else n += 0xFFFFFFULL;
            n = ref_next_prime(n);
        } while (n * lf < at);
        size = n;
        mx = (u64)(size * lf);
    }
    return size;
}

template <int NW>
void EngineT<NW>::build_layout() {
    const int P = prm_.P;
    if (P < 1 || P > 255) throw std::runtime_error("pgb200: -p must be in 1..255 (reference thread ids are unsigned char)");
    read_counters();
    n_nodes_ = h_cnt_[C_DISTINCT];
    std::vector<u64> set_size(P), set_base(P);
    DevBuf d_geom;
    d_geom.alloc(3 * P * sizeof(u64));
    u64* d_size = d_geom.template as<u64>();
    u64* d_base = d_size + P;
    u64* d_count = d_base + P;
    PG_CUDA(cudaMemsetAsync(d_count, 0, P * sizeof(u64), st_));
    k_count_sets<NW><<<148 * 8, 256, 0, st_>>>(tab_, P, d_count);
    PG_CUDA(cudaGetLastError());
    std::vector<u64> cnt(P);
    PG_CUDA(cudaMemcpyAsync(cnt.data(), d_count, P * sizeof(u64), cudaMemcpyDeviceToHost, st_));
    sync();
    if (prm_.initG) {
        u64 sz = ref_static_set_size(prm_.initG, P, prm_.flavour127 != 0);
        for (int i = 0; i < P; i++) {
            set_size[i] = sz;
            if (cnt[i] >= sz) throw std::runtime_error("pgb200: -a too small: a reference set would overflow (the reference spins forever here)");
        }
        layout_exact_ = true;
    } else {
        // Dynamic growth: the reference's final layout depends on the whole growth history (in-place rehash with
        // displacement chains, newhash.c:403-452).  Sizes are reproduced; the slot order is the FCFS order in the final
        // size, which equals the reference's only for sets that never grew.  Bit-exact runs use -a.
        layout_exact_ = true;
        for (int i = 0; i < P; i++) {
            set_size[i] = ref_dynamic_final_size(cnt[i]);
            if (set_size[i] != ref_next_prime(1024)) layout_exact_ = false;
        }
        if (!layout_exact_ && prm_.verbose >= 0)
            fprintf(stderr, "[pgb200] note: no -a given and the reference's sets would have grown: iteration order is FCFS in the final "
                            "set size, not the reference's growth-history order (outputs are valid but not byte-identical)\n");
    }
    u64 total = 0;
    for (int i = 0; i < P; i++) { set_base[i] = total; total += set_size[i]; }
    PG_CUDA(cudaMemcpyAsync(d_size, set_size.data(), P * sizeof(u64), cudaMemcpyHostToDevice, st_));
    PG_CUDA(cudaMemcpyAsync(d_base, set_base.data(), P * sizeof(u64), cudaMemcpyHostToDevice, st_));
    set_size_ = set_size[0];

    DevBuf Rb, gposb, scratch;
    Rb.alloc(total * sizeof(u64));
    gposb.alloc(cap_ * sizeof(u64));
    PG_CUDA(cudaMemsetAsync(Rb.p, 0xFF, total * sizeof(u64), st_));
    u64* R = Rb.template as<u64>();
    RefGeom g{d_size, d_base, P, prm_.flavour127 != 0};
    k_layout_place<NW><<<148 * 8, 256, 0, st_>>>(tab_, g, R);
    k_layout_resolve<NW><<<148 * 8, 256, 0, st_>>>(tab_, g, R, gposb.template as<u64>());
    k_layout_fill<NW><<<148 * 8, 256, 0, st_>>>(tab_, gposb.template as<u64>(), R);
    PG_CUDA(cudaGetLastError());
    order_buf_.alloc((n_nodes_ + 1) * sizeof(u64));
    scratch.alloc(scan_scratch_elems(total) * sizeof(u64));
    device_scan(OccIn{R}, OrderOut<NW>{R, order_buf_.template as<u64>(), tab_.slots}, total, scratch.template as<u64>(), d_cnt_ + C_MISC0, st_);
    read_counters();
    if (h_cnt_[C_MISC0] != n_nodes_) throw std::runtime_error("pgb200: internal error: layout lost k-mers");
}

// ---------------------------------------------------------------- parity dump (record format documented in include/pregraph_b200.h)
template <int NW>
__global__ void k_dump_nodes(Table<NW> tab, const u64* order, u64 n, int out_words, unsigned char* out) {
    const int rec = out_words * 8 + 10;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const Slot<NW>* s = tab.slots + order[i];
        Kmer<NW> k = slot_key(s);
        u64 p = s->payload;
        unsigned char* o = out + i * rec;
        for (int w = 0; w < out_words; w++) {
            int src = w - (out_words - NW);
            u64 v = src >= 0 ? k.w[src] : 0ull;
            for (int b = 0; b < 8; b++) o[w * 8 + b] = (unsigned char)(v >> (8 * b));
        }
        o += out_words * 8;
        for (int c = 0; c < 4; c++) { o[c] = (unsigned char)pl_l(p, c); o[4 + c] = (unsigned char)pl_r(p, c); }
        o[8] = (unsigned char)pl_cov(p);
        o[9] = (unsigned char)(((p & PL_SINGLE) ? 1 : 0) | ((p & PL_LINEAR) ? 2 : 0) | ((p & PL_DELETED) ? 4 : 0));
    }
}

template <int NW>
void EngineT<NW>::dump_nodes(void* host_out) {
    if (!order_buf_.p) throw std::runtime_error("pgb200: dump_nodes before build_layout");
    const int out_words = prm_.flavour127 ? 4 : 2;
    const size_t rec = out_words * 8 + 10;
    DevBuf d;
    d.alloc(n_nodes_ * rec);
    k_dump_nodes<NW><<<148 * 4, 256, 0, st_>>>(tab_, order_buf_.template as<u64>(), n_nodes_, out_words, d.template as<unsigned char>());
    PG_CUDA(cudaGetLastError());
    PG_CUDA(cudaMemcpyAsync(host_out, d.p, n_nodes_ * rec, cudaMemcpyDeviceToHost, st_));
    sync();
}

template void EngineT<2>::build_layout();
template void EngineT<4>::build_layout();
template void EngineT<2>::dump_nodes(void*);
template void EngineT<4>::dump_nodes(void*);

}   // namespace pgb
