// bucket.cu -- locality + ownership for the insert: instead of letting 3e5 threads hit random HBM sectors all over a
// 17 GB table, the (k-mer, links, rank) tuples of a chunk are first grouped by (owner GPU, table region) and then
// applied region by region, so that a region's slots (8-16 MB) stay L2-resident while its tuples stream through.
// The same grouping IS the multi-GPU exchange format: owner-major ranges of the tuple buffer are what the bucketed
// NCCL all-to-all ships (the reference's analogue is "every thread scans the batch and keeps
// hash % thrd_num == id", prlHashReads.c:79-90).
//
// Exact counting sort, no global atomics: K_count writes per-(tile, bucket) counts, one scan in bucket-major order turns
// them into offsets, K_scatter re-chops the tile and places each tuple with a shared-memory cursor.  A tile is the 256
// reads of one block.  DRAM traffic per k-mer instance: 8*(NW+1) B written + read (streaming) instead of a 128 B random
// line fetch + 32 B write-back.
#include "engine_impl.cuh"
#include "scan.cuh"
#include "chop.cuh"
#include <cstring>

namespace pgb {

constexpr int BK_THREADS = 256;
// tuple = key words + meta, padded to whole 32 B sectors so that a tuple is written with full-sector stores (no RMW fills)
template <int NW> struct TupleW { static constexpr int value = NW == 2 ? 4 : 8; };

template <int NW>
__device__ __forceinline__ unsigned bucket_of(const Kmer<NW>& k, u64 mask, int region_shift, int region_bits, int world) {
    u64 h = table_hash(k);
    // region_shift >= 0: real table regions (locality); region_shift < 0: hash-derived pseudo regions that only spread the shared-
    // memory cursors -- used for the multi-GPU exchange, where per-chunk region order would make all instances of a k-mer arrive
    // at its owner back to back and fight over one slot (measured: 1.16e10 vs 1.9e10 tuples/s)
    unsigned region = !region_bits ? 0u : region_shift >= 0 ? (unsigned)((h & mask) >> region_shift) : (unsigned)((h >> 20) & ((1u << region_bits) - 1));
    unsigned owner = world > 1 ? (unsigned)((h >> 40) % (u64)world) : 0u;
    return (owner << region_bits) | region;
}

template <int NW>
struct CountSink {
    unsigned* hist;
    u64 mask;
    int region_shift, region_bits, world;
    __device__ __forceinline__ void operator()(const Kmer<NW>& canon, unsigned, unsigned, int) {
        unsigned b = bucket_of(canon, mask, region_shift, region_bits, world);
        if (region_bits == 0) {   // few buckets (owners only): one shared-memory atomic per warp and owner instead of one per lane
            unsigned m = __match_any_sync(__activemask(), b);
            if ((int)(threadIdx.x & 31) == __ffs(m) - 1) atomicAdd(&hist[b], (unsigned)__popc(m));
        } else atomicAdd(&hist[b], 1u);
    }
};

template <int NW>
__global__ void __launch_bounds__(BK_THREADS) k_bucket_count(KParams<NW> kp, const u64* __restrict__ words, const u32* __restrict__ lens, u64 n_rec,
                                                              int W64, u64 mask, int region_shift, int region_bits, int world, int NB, u32* tilecnt,
                                                              int rpt, u64 tile0) {
    extern __shared__ unsigned s_hist[];
    for (int b = threadIdx.x; b < NB; b += BK_THREADS) s_hist[b] = 0;
    __syncthreads();
    for (int q = 0; q < rpt; q++) {   // a tile = rpt * BK_THREADS consecutive reads of the chunk
        u64 r = ((u64)blockIdx.x * rpt + q) * BK_THREADS + threadIdx.x;
        if (r < n_rec) {
            int L = (int)lens[r];
            if (L >= kp.K + 1) {
                CountSink<NW> sink{s_hist, mask, region_shift, region_bits, world};
                chop_read(kp, words + r * (u64)W64, L, sink);
            }
        }
    }
    __syncthreads();
    u32* row = tilecnt + (tile0 + blockIdx.x) * NB;
    for (int b = threadIdx.x; b < NB; b += BK_THREADS) row[b] = s_hist[b];
}

// bucket-major exclusive scan over the (tile x bucket) count matrix
struct TileCntIn {
    const u32* cnt;
    u64 n_tiles;
    int NB;
    __device__ u64 operator()(u64 i) const { u64 b = i / n_tiles, t = i - b * n_tiles; return cnt[t * NB + b]; }
};
struct TileOffOut {
    u32* off;
    u64 n_tiles;
    int NB;
    __device__ void operator()(u64 i, u64 prefix, u64) const { u64 b = i / n_tiles, t = i - b * n_tiles; off[t * NB + b] = (u32)prefix; }
};

// Fused exchange: owner o's tuples are stored STRAIGHT into rank o's receive buffer (a peer-mapped pointer, NVLink stores)
// instead of into a local send buffer that NCCL would copy later: dst[o].ptr already points at this rank's slice.
struct PeerDst {
    u64* ptr;     // where this rank's first tuple for owner o goes (peer or local memory)
    u64 start;    // local tuple index of owner o's range (offsets from the counting sort are relative to it)
};

template <int NW>
struct ScatterSink {
    unsigned* cursor;
    u64* tuples;
    u64 ord;
    u64 mask;
    int region_shift, region_bits, world;
    const PeerDst* peers;   // nullptr: plain local buffer
    __device__ __forceinline__ void operator()(const Kmer<NW>& canon, unsigned left, unsigned right, int j) {
        unsigned b = bucket_of(canon, mask, region_shift, region_bits, world);
        u64 p;
        if (region_bits == 0) {
            // warp-aggregated cursor: the lanes that go to the same owner take CONSECUTIVE tuple slots, so their 32-byte stores
            // coalesce into a few long transactions -- what makes direct peer stores over NVLink efficient
            unsigned m = __match_any_sync(__activemask(), b);
            int leader = __ffs(m) - 1, lane = threadIdx.x & 31;
            unsigned base = 0;
            if (lane == leader) base = atomicAdd(&cursor[b], (unsigned)__popc(m));
            base = __shfl_sync(m, base, leader);
            p = base + __popc(m & ((1u << lane) - 1));
        } else p = atomicAdd(&cursor[b], 1u);
        u64* t;
        if (peers) { PeerDst d = peers[b >> region_bits]; t = d.ptr + (p - d.start) * TupleW<NW>::value; }
        else t = tuples + p * TupleW<NW>::value;
        u64 m = tuple_meta(ord, j, left, right);
        if (NW == 2) {
            asm volatile("st.global.cs.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(t), "l"(canon.w[0]), "l"(canon.w[1]), "l"(m), "l"(0ull) : "memory");
        } else {
            asm volatile("st.global.cs.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(t), "l"(canon.w[0]), "l"(canon.w[1]), "l"(canon.w[2]), "l"(canon.w[NW - 1]) : "memory");
            asm volatile("st.global.cs.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(t + 4), "l"(m), "l"(0ull), "l"(0ull), "l"(0ull) : "memory");
        }
    }
};

template <int NW>
__global__ void __launch_bounds__(BK_THREADS) k_bucket_scatter(KParams<NW> kp, const u64* __restrict__ words, const u32* __restrict__ lens, u64 n_rec,
                                                                int W64, u64 ord_base, u64 ord_stride, u64 mask, int region_shift, int region_bits,
                                                                int world, int NB, const u32* tileoff, u64* tuples, int rpt, u64 tile0,
                                                                const PeerDst* peers) {
    extern __shared__ unsigned s_cur[];
    const u32* row = tileoff + (tile0 + blockIdx.x) * NB;
    for (int b = threadIdx.x; b < NB; b += BK_THREADS) s_cur[b] = row[b];
    __syncthreads();
    for (int q = 0; q < rpt; q++) {
        u64 r = ((u64)blockIdx.x * rpt + q) * BK_THREADS + threadIdx.x;
        if (r < n_rec) {
            int L = (int)lens[r];
            if (L >= kp.K + 1) {
                ScatterSink<NW> sink{s_cur, tuples, ord_base + r * ord_stride, mask, region_shift, region_bits, world, peers};
                chop_read(kp, words + r * (u64)W64, L, sink);
            }
        }
    }
}

// streaming read of the tuple (evict-first: it is touched exactly once and must not push table lines out of L2)
__device__ __forceinline__ u64 ld_stream(const u64* p) {
    u64 v;
    asm volatile("ld.global.cs.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

template <int NW>
__global__ void __launch_bounds__(BK_THREADS) k_apply_tuples(Table<NW> tab, const u64* __restrict__ tuples, u64 n, u64* counters) {
    __shared__ unsigned s_new;
    if (threadIdx.x == 0) s_new = 0;
    __syncthreads();
    unsigned my_new = 0;
    // NOT a grid-stride loop: one tuple per thread and as many blocks as needed.  The hardware hands out blocks in index
    // order, so the blocks resident at any moment cover one contiguous window of the region-sorted tuples (a persistent
    // grid drifts apart and the window -- hence the table working set -- grows without bound; measured: 180 B of DRAM reads
    // per tuple with the persistent variant).
    // (a persistent grid was also measured for the exchange path, where order does not matter: 139 ms vs 78 ms per 8.8e8 tuples)
    for (u64 i = (u64)blockIdx.x * BK_THREADS + threadIdx.x; i < n; i += (u64)gridDim.x * BK_THREADS) {
        const u64* t = tuples + i * TupleW<NW>::value;
        U256 a;
        asm volatile("ld.global.cs.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a.a), "=l"(a.b), "=l"(a.c), "=l"(a.d) : "l"(t));
        Kmer<NW> k;
        u64 m;
        if (NW == 2) { k.w[0] = a.a; k.w[NW - 1] = a.b; m = a.c; }
        else { k.w[0] = a.a; k.w[1] = a.b; k.w[NW == 4 ? 2 : 0] = a.c; k.w[NW - 1] = a.d; m = ld_stream(t + 4); }
        my_new += table_insert(tab, k, (unsigned)((m >> 3) & 7), (unsigned)(m & 7), m >> 6);
    }
    if (my_new) atomicAdd(&s_new, my_new);
    __syncthreads();
    if (threadIdx.x == 0 && s_new) atomicAdd(&counters[C_DISTINCT], (u64)s_new);
}

template <int NW>
void EngineT<NW>::bucket_chunk(const ReadChunk& ch) { bucket_chunks(&ch, 1, prm_.world > 1 ? (xchg_fused_ ? 1 : 0) : (8ull << 20), 1, !xchg_fused_); }

// Counting sort of the tuples of `n` chunks by (owner, table region of ~region_bytes).  rpt = reads per thread (tile size).
template <int NW>
void EngineT<NW>::bucket_chunks(const ReadChunk* chs, size_t n, u64 region_bytes, int rpt, bool scatter_now) {
    const int world = prm_.world > 1 ? prm_.world : 1;
    int log2cap = 0;
    while ((1ull << log2cap) < cap_) log2cap++;
    u64 table_bytes = cap_ * sizeof(Slot<NW>);
    int rb = 0;
    int region_shift;
    if (region_bytes == 0) { rb = 6; region_shift = -1; }   // pseudo regions (see bucket_of)
    else if (region_bytes == 1) { rb = 0; region_shift = 0; }   // owners only (fused peer-store exchange: warp-aggregated cursors)
    else {
        while ((table_bytes >> rb) > region_bytes && rb < 12 && world * (2 << rb) <= 8192) rb++;
        if (rb > log2cap) rb = log2cap;
        region_shift = log2cap - rb;
    }
    region_bits_ = rb;
    const int NB = world << rb;
    n_buckets_ = NB;
    const u64 tile_reads = (u64)BK_THREADS * rpt;
    std::vector<u64> tile0(n + 1, 0);
    for (size_t c = 0; c < n; c++) tile0[c + 1] = tile0[c] + (chs[c].n_rec + tile_reads - 1) / tile_reads;
    const u64 n_tiles = tile0[n];
    const u64 cells = n_tiles * (u64)NB;
    tilecnt_buf_.ensure(cells * sizeof(u32));
    tileoff_buf_.ensure(cells * sizeof(u32));
    scan_buf_.ensure(scan_scratch_elems(cells) * sizeof(u64));
    size_t smem = (size_t)NB * sizeof(unsigned);
    static bool attr_set = false;
    if (!attr_set && smem > 48 * 1024) {
        cudaFuncSetAttribute(k_bucket_count<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        cudaFuncSetAttribute(k_bucket_scatter<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        attr_set = true;
    }
    for (size_t c = 0; c < n; c++) {
        if (!chs[c].n_rec) continue;
        k_bucket_count<NW><<<(unsigned)(tile0[c + 1] - tile0[c]), BK_THREADS, smem, st_>>>(kp_, chs[c].words, chs[c].len, chs[c].n_rec, W64_, tab_.mask, region_shift,
                                                                                          rb, world, NB, tilecnt_buf_.template as<u32>(), rpt, tile0[c]);
    }
    PG_CUDA(cudaGetLastError());
    device_scan(TileCntIn{tilecnt_buf_.template as<u32>(), n_tiles, NB}, TileOffOut{tileoff_buf_.template as<u32>(), n_tiles, NB}, cells,
                scan_buf_.template as<u64>(), d_cnt_ + C_MISC1, st_);
    // total + owner range starts (tile 0's offset of an owner's first bucket == start of that owner's range)
    std::vector<u32> first_row(NB);
    PG_CUDA(cudaMemcpyAsync(first_row.data(), tileoff_buf_.p, NB * sizeof(u32), cudaMemcpyDeviceToHost, st_));
    read_counters();
    n_tuples_ = h_cnt_[C_MISC1];
    if (n_tuples_ >= 0xFFFFFFFFull) throw std::runtime_error("pgb200: more than 2^32 k-mer instances in one batch; use a smaller batch");
    owner_start_.assign(world + 1, n_tuples_);
    for (int o = 0; o < world; o++) owner_start_[o] = first_row[(size_t)o << rb];
    bk_ = BucketGeom{region_shift, rb, NB, rpt, tile0};
    if (!scatter_now) return;
    tuple_flip_ ^= 1;
    tuple_buf().ensure((n_tuples_ + 1) * TupleW<NW>::value * sizeof(u64));
    bucket_scatter(chs, n, tuple_buf().template as<u64>(), nullptr);
}

template <int NW>
void EngineT<NW>::bucket_scatter(const ReadChunk* chs, size_t n, u64* tuples, const void* peers) {
    const int world = prm_.world > 1 ? prm_.world : 1;
    size_t smem = (size_t)bk_.NB * sizeof(unsigned);
    for (size_t c = 0; c < n; c++) {
        if (!chs[c].n_rec) continue;
        k_bucket_scatter<NW><<<(unsigned)(bk_.tile0[c + 1] - bk_.tile0[c]), BK_THREADS, smem, st_>>>(
            kp_, chs[c].words, chs[c].len, chs[c].n_rec, W64_, chs[c].ord_base, chs[c].ord_stride, tab_.mask, bk_.region_shift, bk_.rb, world, bk_.NB,
            tileoff_buf_.template as<u32>(), tuples, bk_.rpt, bk_.tile0[c], reinterpret_cast<const PeerDst*>(peers));
    }
    PG_CUDA(cudaGetLastError());
    p1_.launches += 3 + 2 * n;
}

// ---------------------------------------------------------------- fused exchange (peer stores over NVLink)
template <int NW>
void EngineT<NW>::xchg_setup(uint64_t cap_tuples) {
    xchg_cap_ = cap_tuples;
    for (int b = 0; b < 2; b++) xchg_recv_[b].alloc((cap_tuples + 1) * TupleW<NW>::value * sizeof(u64));
    xchg_peer_[0].assign(prm_.world, nullptr);
    xchg_peer_[1].assign(prm_.world, nullptr);
    for (int b = 0; b < 2; b++) xchg_peer_[b][prm_.rank] = xchg_recv_[b].p;
    xchg_dst_.alloc(prm_.world * sizeof(PeerDst));
    xchg_fused_ = true;
}
template <int NW>
void EngineT<NW>::xchg_export(int buf, void* handle64) {
    cudaIpcMemHandle_t h;
    PG_CUDA(cudaIpcGetMemHandle(&h, xchg_recv_[buf].p));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle64, &h, 64);
}
template <int NW>
void EngineT<NW>::xchg_import(int peer, int buf, const void* handle64) {
    if (peer == prm_.rank) return;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    PG_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    xchg_peer_[buf][peer] = p;
}
template <int NW>
void EngineT<NW>::xchg_counts(uint64_t* counts) {
    const int world = prm_.world;
    for (int o = 0; o < world; o++) counts[o] = owner_start_.size() == (size_t)world + 1 ? owner_start_[o + 1] - owner_start_[o] : 0;
}
// base[o] = tuple index inside rank o's receive buffer where this rank's slice starts
template <int NW>
void EngineT<NW>::xchg_scatter(int buf, const uint64_t* base) {
    const int world = prm_.world;
    if (owner_start_.size() != (size_t)world + 1 || !n_tuples_) { sync(); return; }
    std::vector<PeerDst> d(world);
    for (int o = 0; o < world; o++) {
        u64 cnt = owner_start_[o + 1] - owner_start_[o];
        if (base[o] + cnt > xchg_cap_) throw std::runtime_error("pgb200: fused exchange receive buffer too small");
        if (!xchg_peer_[buf][o]) throw std::runtime_error("pgb200: fused exchange peer buffer not imported");
        d[o].ptr = reinterpret_cast<u64*>(xchg_peer_[buf][o]) + base[o] * TupleW<NW>::value;
        d[o].start = owner_start_[o];
    }
    PG_CUDA(cudaMemcpyAsync(xchg_dst_.p, d.data(), world * sizeof(PeerDst), cudaMemcpyHostToDevice, st_));
    bucket_scatter(&chunks_.back(), 1, nullptr, xchg_dst_.p);
    sync();   // all peer stores of this rank are performed; the caller's barrier makes every rank's visible
}
// Apply what the peers stored into receive buffer `buf`, on a SECOND stream and without waiting for it: the random-access apply of
// round i overlaps the decode / count / peer-store scatter of round i+1 (different bottlenecks: HBM row accesses vs ALU + NVLink).
template <int NW>
void EngineT<NW>::xchg_apply(int buf, uint64_t n) {
    sync_apply();                                   // round i-1's apply (normally long finished); settles its timing
    if (!n) return;
    create_table_if_needed();
    PG_CUDA(cudaMemcpyAsync(h_cnt_, d_cnt_, C_COUNT * sizeof(u64), cudaMemcpyDeviceToHost, st_apply_));
    PG_CUDA(cudaStreamSynchronize(st_apply_));
    ensure_table_bound(h_cnt_[C_DISTINCT], n);      // growth (rare) happens here, with nothing else touching the table
    PG_CUDA(cudaEventRecord(ev_apply_[0], st_apply_));
    unsigned blocks = (unsigned)((n + BK_THREADS - 1) / BK_THREADS);
    k_apply_tuples<NW><<<blocks, BK_THREADS, 0, st_apply_>>>(tab_, reinterpret_cast<const u64*>(xchg_recv_[buf].p), n, d_cnt_);
    PG_CUDA(cudaGetLastError());
    PG_CUDA(cudaEventRecord(ev_apply_[1], st_apply_));
    apply_inflight_ = true;
    p1_.launches += 1;
}
template <int NW>
void EngineT<NW>::sync_apply() {
    if (!apply_inflight_) return;
    PG_CUDA(cudaEventSynchronize(ev_apply_[1]));
    float ms;
    PG_CUDA(cudaEventElapsedTime(&ms, ev_apply_[0], ev_apply_[1]));
    p1_.ms_insert += ms;
    apply_inflight_ = false;
}

// Batch mode (single GPU): every chunk fed since the last flush is bucketed by 32 MB table region in ONE counting sort and applied
// region by region.  The whole table streams through L2 once per batch instead of being hit at random once per instance; the
// price is 32 B written + read per instance for the tuple buffer.
template <int NW>
void EngineT<NW>::flush_batch() {
    if (!skm_pending_.empty()) skm_flush();
    if (batch_gb_ <= 0 || prm_.world > 1 || pending_first_ >= chunks_.size()) return;   // only the single-GPU batch mode defers inserts
    settle_timing();
    read_counters();
    ensure_table_bound(h_cnt_[C_DISTINCT], pending_bound_);
    PG_CUDA(cudaEventRecord(ev_[0], st_));
    PG_CUDA(cudaEventRecord(ev_[1], st_));
    PG_CUDA(cudaEventRecord(ev_[2], st_));
    bucket_chunks(chunks_.data() + pending_first_, chunks_.size() - pending_first_, 32ull << 20, 4, true);
    apply_tuples(tuple_buf().template as<u64>(), n_tuples_);
    PG_CUDA(cudaEventRecord(ev_[3], st_));
    timing_pending_ = true;
    pending_first_ = chunks_.size();
    pending_bound_ = 0;
}

template <int NW>
void EngineT<NW>::apply_tuples(const u64* tuples, u64 n) {
    if (!n) return;
    if (!tab_.slots) ensure_table(n);
    unsigned blocks = (unsigned)((n + BK_THREADS - 1) / BK_THREADS);
    k_apply_tuples<NW><<<blocks, BK_THREADS, 0, st_>>>(tab_, tuples, n, d_cnt_);
    PG_CUDA(cudaGetLastError());
    p1_.launches += 1;
}

template void EngineT<2>::bucket_chunk(const ReadChunk&);
template void EngineT<2>::bucket_chunks(const ReadChunk*, size_t, u64, int, bool);
template void EngineT<4>::bucket_chunks(const ReadChunk*, size_t, u64, int, bool);
template void EngineT<2>::bucket_scatter(const ReadChunk*, size_t, u64*, const void*);
template void EngineT<4>::bucket_scatter(const ReadChunk*, size_t, u64*, const void*);
template void EngineT<2>::xchg_setup(uint64_t); template void EngineT<4>::xchg_setup(uint64_t);
template void EngineT<2>::xchg_export(int, void*); template void EngineT<4>::xchg_export(int, void*);
template void EngineT<2>::xchg_import(int, int, const void*); template void EngineT<4>::xchg_import(int, int, const void*);
template void EngineT<2>::xchg_counts(uint64_t*); template void EngineT<4>::xchg_counts(uint64_t*);
template void EngineT<2>::xchg_scatter(int, const uint64_t*); template void EngineT<4>::xchg_scatter(int, const uint64_t*);
template void EngineT<2>::xchg_apply(int, uint64_t); template void EngineT<4>::xchg_apply(int, uint64_t);
template void EngineT<2>::sync_apply(); template void EngineT<4>::sync_apply();
template void EngineT<2>::flush_batch();
template void EngineT<4>::flush_batch();
template void EngineT<4>::bucket_chunk(const ReadChunk&);
template void EngineT<2>::apply_tuples(const u64*, u64);
template void EngineT<4>::apply_tuples(const u64*, u64);

}   // namespace pgb
