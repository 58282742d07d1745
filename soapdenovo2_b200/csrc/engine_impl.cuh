// engine_impl.cuh -- EngineT<NW>: device state + phase methods; the methods are defined across decode.cu, pass1.cu,
// layout.cu, tips.cu, edges.cu, pass2.cu and explicitly instantiated for NW = 2 (K <= 63) and NW = 4 (K <= 127).
#pragma once
#include "engine.h"
#include "kmer.cuh"
#include "table.cuh"
#include "skm.cuh"
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

namespace pgb {

#define PG_CUDA(call)                                                                                         \
    do {                                                                                                      \
        cudaError_t e__ = (call);                                                                             \
        if (e__ != cudaSuccess) {                                                                             \
            char b__[512];                                                                                    \
            snprintf(b__, sizeof b__, "CUDA error %s at %s:%d: %s", cudaGetErrorName(e__), __FILE__, __LINE__, \
                     cudaGetErrorString(e__));                                                                \
            throw std::runtime_error(b__);                                                                    \
        }                                                                                                     \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    DevBuf() {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void alloc(size_t n) {
        release();
        if (n == 0) n = 16;
        PG_CUDA(cudaMalloc(&p, n));
        bytes = n;
    }
    void ensure(size_t n) { if (n > bytes) alloc(n + n / 4); }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// One fed text chunk, decoded: 2-bit packed reads (LSB-first, W64 words per read) + lengths.  Stays resident in HBM so
// that pass 2 re-scans the reads without touching the text again (the reference re-parses every file, prlRead2path.c:786).
struct ReadChunk {
    u64* words = nullptr;
    u32* len = nullptr;
    u64 n_rec = 0;
    u64 ord_base = 0, ord_stride = 1;
};

enum Counter { C_DISTINCT = 0, C_INSTANCES, C_KEPT, C_LINEAR, C_REMOVED, C_MISC0, C_MISC1, C_MISC2, C_COUNT = 16 };

template <int NW>
class EngineT : public IEngine {
public:
    explicit EngineT(const PgParams& p);
    ~EngineT() override;

    void feed_text(const char* text, size_t nbytes, bool on_device, int fastq, uint64_t ord_base, uint64_t ord_stride,
                   int reverse_seq, int maxlen) override;
    uint64_t last_chunk_records() const override { return last_records_; }
    void finish_pass1(Pass1Stats* st) override;
    void reset_pass1() override;
    void sweeps(SweepStats* st) override;
    void build_layout() override;
    uint64_t node_count() const override { return n_nodes_; }
    void dump_nodes(void* host_out) override;
    void remove_tips(TipStats* st) override;
    void build_edges(EdgeStats* st, std::string* edge_text) override;
    void pass2(Pass2Stats* st, std::string* prearc_text, std::string* path_bin, std::string* mark_text) override;
    void vertices(std::string* vertex_text, uint64_t* n_vertex) override;

    // ---- state
    PgParams prm_;
    KParams<NW> kp_;
    cudaStream_t st_ = nullptr;
    cudaEvent_t ev_[4] = {nullptr, nullptr, nullptr, nullptr};
    int W64_ = 0;   // packed words per read
    int insert_sm_ = 0;     // PGB200_INSERT_SM=1: per-lane state-machine variant of the insert (K <= 63); measured slower, kept for the record
    int dbg_insert_ = 0;    // PGB200_DBG_INSERT: cost-decomposition variants of k_chop_insert (profiling only)
    int l2gran_mode_ = 0;   // PGB200_L2GRAN: 0 default, 1 = 32 B globally, 2 = 32 B only around k_chop_insert

    std::vector<ReadChunk> chunks_;
    uint64_t last_records_ = 0, total_records_ = 0;
    Pass1Stats p1_;

    Table<NW> tab_{nullptr, 0};
    u64 cap_ = 0;
    DevBuf tab_buf_;
    u64* d_cnt_ = nullptr;      // Counter[]
    u64* h_cnt_ = nullptr;      // pinned mirror

    // decode scratch
    DevBuf text_bufs_[2], line_buf_, scan_buf_;
    int text_flip_ = 0;
    cudaStream_t st_copy_ = nullptr;   // H2D of chunk i+1 overlaps the insert of chunk i
    cudaEvent_t ev_copy_ = nullptr;
    bool timing_pending_ = false;
    void settle_timing();

    // layout: reference geometry
    u64 set_size_ = 0;          // prime size of every reference set (static -a mode)
    u64 n_nodes_ = 0;           // distinct k-mers == entries in iteration order
    DevBuf order_buf_;          // u64 order[n_nodes_] : iteration index -> ktab slot
    bool layout_exact_ = false;

    // edges / pass 2 state
    u64 num_ed_ = 0;
    DevBuf patch_buf_;          // (K+1)-mer patch table
    u64 patch_mask_ = 0;

    // bucketed insert path (bucket.cu): tuples {key words, meta} grouped by (owner GPU, table region)
    int bucket_mode_ = 0;       // PGB200_BUCKET=1: single-GPU inserts also go through the bucketed (region-sorted) path
    DevBuf tuple_bufs_[2], tilecnt_buf_, tileoff_buf_;   // two tuple buffers: chunk i+1 is bucketed while chunk i is on the wire
    int tuple_flip_ = 0;
    DevBuf& tuple_buf() { return tuple_bufs_[tuple_flip_]; }
    u64 n_tuples_ = 0;          // tuples currently in tuple_buf_
    int n_buckets_ = 0, region_bits_ = 0;
    std::vector<u64> owner_start_;   // [world + 1] tuple offsets of each owner's range in tuple_buf_
    void bucket_chunk(const ReadChunk& ch);
    void bucket_chunks(const ReadChunk* chs, size_t n, u64 region_bytes, int rpt, bool scatter_now);
    void bucket_scatter(const ReadChunk* chs, size_t n, u64* tuples, const void* peers);
    struct BucketGeom { int region_shift = 0, rb = 0, NB = 0, rpt = 1; std::vector<u64> tile0; } bk_;
    // fused exchange: receive buffers other ranks store into directly (CUDA IPC peer mappings over NVLink)
    bool xchg_fused_ = false;
    u64 xchg_cap_ = 0;
    DevBuf xchg_recv_[2], xchg_dst_;
    std::vector<void*> xchg_peer_[2];
    cudaStream_t st_apply_ = nullptr;
    cudaEvent_t ev_apply_[2] = {nullptr, nullptr};
    bool apply_inflight_ = false;
    void sync_apply();
    // batch mode (PGB200_BATCH_GB > 0): inserts are deferred and done region-sorted over many chunks at once
    double batch_gb_ = 0;
    size_t pending_first_ = 0;   // chunks_[pending_first_..) are decoded but not inserted yet
    u64 pending_bound_ = 0;      // upper bound of their k-mer instances
    void flush_batch();
    void apply_tuples(const u64* tuples, u64 n);
    // aggregated pass 1 (skm.cu, single GPU): super-k-mer partition per chunk, shared-memory aggregation per bucket,
    // one global update per DISTINCT k-mer
    int skm_mode_ = -1;   // -1: aggregated for device-resident text, per-chunk inserts for host text; 0 / 1: forced (PGB200_SKM)
    SkmGeom skm_geom_;
    DevBuf skm_inst_, skm_cursor_, skm_desc_, skm_side_;   // k-mers per bucket (pending chunks), scatter cursors, chunk descriptors + bucket counter
    struct SkmPending { size_t chunk = 0; u32* segoff = nullptr; u64* recs = nullptr; u64 n_recs = 0; };
    std::vector<SkmPending> skm_pending_;
    bool skm_unscattered_ = false;               // the last pending chunk is counted but its records are not written yet
    u64 skm_pending_recs_ = 0, skm_prev_total_ = 0;
    std::vector<std::pair<void*, size_t>> skm_blocks_;   // bump-allocated scratch (segment offsets, records), reused after every flush
    size_t skm_blk_ = 0, skm_used_ = 0;
    cudaEvent_t ev_skm_[2] = {nullptr, nullptr};
    int skm_part_threads_ = 128;
    void* skm_alloc(size_t bytes);
    void skm_init();
    void skm_count_chunk(size_t ci);
    void skm_scatter_last(u64 total);
    void skm_flush();
    void skm_reset();
    void skm_release();
public:
    // multi-GPU exchange surface (C-ABI: pgb200_exchange_*)
    static constexpr int tuple_words() { return NW == 2 ? 4 : 8; }
    void xchg_setup(uint64_t cap_tuples) override;
    void xchg_export(int buf, void* handle64) override;
    void xchg_import(int peer, int buf, const void* handle64) override;
    void xchg_counts(uint64_t* counts) override;
    void xchg_scatter(int buf, const uint64_t* base) override;
    void xchg_apply(int buf, uint64_t n) override;
    const void* exchange_buffer(uint64_t* ranges, int* tuple_bytes) override {
        const int world = prm_.world > 1 ? prm_.world : 1;
        for (int o = 0; o <= world; o++) ranges[o] = owner_start_.size() == (size_t)world + 1 ? owner_start_[o] : 0;
        *tuple_bytes = tuple_words() * 8;
        sync();
        return tuple_buf().p;
    }
    void exchange_clear() override { owner_start_.clear(); n_tuples_ = 0; }
    void apply_received(const void* tuples, uint64_t n) override {
        if (!n) return;
        ensure_table(0);
        ensure_table(n);
        PG_CUDA(cudaEventRecord(ev_[2], st_));
        apply_tuples(reinterpret_cast<const u64*>(tuples), n);
        PG_CUDA(cudaEventRecord(ev_[3], st_));
        sync();
        float ms;
        PG_CUDA(cudaEventElapsedTime(&ms, ev_[2], ev_[3]));
        p1_.ms_insert += ms;
    }

    // helpers
    void ensure_table(u64 need_free);
    void ensure_table_bound(u64 have, u64 incoming);
    void create_table_if_needed();
    std::vector<std::pair<void*, size_t>> arena_;   // read-store blocks
    size_t arena_used_ = 0;
    void* arena_alloc(size_t bytes);
    void grow_table(u64 new_cap);
    void alloc_table(u64 cap);
    void sync() { PG_CUDA(cudaStreamSynchronize(st_)); }
    void read_counters() {
        PG_CUDA(cudaMemcpyAsync(h_cnt_, d_cnt_, C_COUNT * sizeof(u64), cudaMemcpyDeviceToHost, st_));
        sync();
    }
};

// reference table geometry helpers (newhash.c:142-185, 200-233; prlHashReads.c:369-390)
inline bool ref_is_prime(u64 n) {
    if (n < 4) return true;
    if (n % 2 == 0) return false;
    u64 mx = (u64)__builtin_sqrtf((float)n);   // float sqrt, strict '<': squares of primes count as prime
    for (u64 i = 3; i < mx; i += 2)
        if (n % i == 0) return false;
    return true;
}
inline u64 ref_next_prime(u64 n) {
    if (n % 2 == 0) n++;
    while (!ref_is_prime(n)) n += 2;
    return n;
}
inline u64 ref_static_set_size(int initG, int P, bool flavour127) {
    u64 want = (u64)((double)initG * 1024.0f * 1024.0f * 1024.0f / (double)P / (flavour127 ? 40 : 24)), k = 0;
    do ++k; while (k * 0xFFFFFFULL < want);
    u64 init = k * 0xFFFFFFULL;
    return init < 3 ? 3 : ref_next_prime(init);
}

}   // namespace pgb
