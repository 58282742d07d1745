// skm.cu -- aggregated pass 1 (default for device-resident text; PGB200_SKM=0/1 forces): super-k-mer partition per chunk, one shared-memory aggregation per bucket, ONE global
// table update per distinct k-mer.  Replaces, for the same result, the per-instance k_chop_insert (pass1.cu), whose speed is pinned
// to the DRAM random-access rate (profiles/r01_rmw_ubench.md).  Logic shared with the host tests lives in skm.cuh.
//
//   feed_text(chunk c):   k_skm_part<false>  count runs per bucket           -> segoff_c[B+1] (k_skm_offsets), bucket_inst[B] += k-mers
//   feed_text(chunk c+1): k_skm_part<true>   write the 8-byte run records of chunk c into its per-bucket segments (the record count
//                                            of chunk c is read with chunk c+1's one host sync: no extra synchronisation)
//   flush (finish_pass1 / 64 chunks pending): k_skm_apply over bucket ranges sized to the free room of the global table
#include "engine_impl.cuh"
#include "skm.cuh"

namespace pgb {

constexpr int SKM_PART_THREADS = 128;
constexpr int SKM_APPLY_THREADS = 256;
constexpr int SKM_LOG2_SLOTS = 11;
constexpr int SKM_SLOTS = 1 << SKM_LOG2_SLOTS;                    // shared-memory table slots per CTA (32 B/slot at K<=63: 64 KB, 3 CTAs per SM)
constexpr int SKM_SOFT_LIMIT = SKM_SLOTS - SKM_APPLY_THREADS - 64;   // claims stop here: the table can never fill up completely

struct SkmChunkDev {
    const u64* words;
    const u32* lens;
    const u64* recs;
    const u32* segoff;
    u64 ord_base, ord_stride;
};

// ------------------------------------------------------------------------------------------------ partition
// The counting pass also leaves the runs of every read in a side buffer (SKM_SIDE_RUNS x 4 B per read: bucket | count << 20 |
// last << 25; start positions are the running sum of the counts), so that the scatter pass does not have to find the minimizers
// again; reads with more runs than that are re-scanned there.
constexpr int SKM_SIDE_RUNS = 16;
struct CountEmit {
    u32* cnt;
    u64* inst;
    u32* side;
    int nrun;
    __device__ __forceinline__ void operator()(u32 b, int, int n, bool last) {
        atomicAdd(&cnt[b], 1u);
        atomicAdd(&inst[b], (u64)n);
        if (nrun < SKM_SIDE_RUNS) side[nrun] = b | ((u32)n << 20) | (last ? 1u << 25 : 0u);
        nrun++;
    }
};
struct ScatterEmit {
    u32* cursor;
    const u32* segoff;
    u64* recs;
    u32 read_idx;
    __device__ __forceinline__ void operator()(u32 b, int s, int n, bool last) {
        u32 i = atomicAdd(&cursor[b], 1u);
        recs[(u64)segoff[b] + i] = skm_pack(read_idx, s, n, last);
    }
};

template <bool SCATTER>
__global__ void __launch_bounds__(SKM_PART_THREADS) k_skm_part(SkmGeom g, const u64* __restrict__ words, const u32* __restrict__ lens, u64 n_rec, int W64,
                                                              u32* cnt_or_cursor, u64* inst, const u32* __restrict__ segoff, u64* recs, u32* side, u8* nruns) {
    extern __shared__ u32 s_ring[];   // [2 * g.w][blockDim.x]: one column per thread, bank = thread -> conflict-free
    for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < n_rec; r += (u64)gridDim.x * blockDim.x) {
        const int L = (int)lens[r];
        const u64* wp = words + r * (u64)W64;
        if (SCATTER) {
            // only the reads whose runs did not fit the side buffer (k_skm_scatter_side handles the rest)
            if (nruns[r] != 255) continue;
            ScatterEmit e{cnt_or_cursor, segoff, recs, (u32)r};
            skm_scan_read(g, wp, L, s_ring + threadIdx.x, (int)blockDim.x, e);
        } else {
            CountEmit e{cnt_or_cursor, inst, side + r * SKM_SIDE_RUNS, 0};
            skm_scan_read(g, wp, L, s_ring + threadIdx.x, (int)blockDim.x, e);
            nruns[r] = e.nrun <= SKM_SIDE_RUNS ? (u8)e.nrun : (u8)255;
        }
    }
}

// scatter pass for the reads whose runs are in the side buffer: no minimizer work, just cursors and 8-byte stores
__global__ void __launch_bounds__(256) k_skm_scatter_side(const u32* __restrict__ side, const u8* __restrict__ nruns, u64 n_rec, u32* cursor,
                                                          const u32* __restrict__ segoff, u64* recs, u32* n_overflow) {
    unsigned ovf = 0;
    for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < n_rec; r += (u64)gridDim.x * blockDim.x) {
        const int nr = nruns[r];
        if (nr == 255) { ovf++; continue; }
        const uint4* row = reinterpret_cast<const uint4*>(side + r * SKM_SIDE_RUNS);
        int start = 0;
        for (int q = 0; q < nr; q += 4) {
            const uint4 v = __ldg(row + (q >> 2));
            const u32 e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int x = 0; x < 4; x++) {
                if (q + x >= nr) break;
                const u32 b = e[x] & 0xFFFFFu;
                const int n = (int)((e[x] >> 20) & 31);
                const u32 i = atomicAdd(&cursor[b], 1u);
                recs[(u64)segoff[b] + i] = skm_pack((u32)r, start, n, (e[x] >> 25) & 1);
                start += n;
            }
        }
    }
    if (ovf) atomicAdd(n_overflow, ovf);
}

// in-place exclusive scan of cnt[0..n) by ONE block; cnt[n] = total (also written to *total_out)
__global__ void __launch_bounds__(1024) k_skm_offsets(u32* cnt, u32 n, u64* total_out) {
    __shared__ u32 s_warp[32];
    const u32 per = (n + blockDim.x - 1) / blockDim.x;
    const u32 lo = threadIdx.x * per, hi = lo + per < n ? lo + per : n;
    u32 sum = 0;
    for (u32 i = lo; i < hi; i++) sum += cnt[i];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    u32 inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 v = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += v;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        u32 ws = s_warp[lane], wi = ws;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u32 v = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += v;
        }
        s_warp[lane] = wi - ws;
        if (lane == 31) { cnt[n] = wi; *total_out = wi; }
    }
    __syncthreads();
    u32 run = s_warp[wid] + inc - sum;
    for (u32 i = lo; i < hi; i++) {
        u32 c = cnt[i];
        cnt[i] = run;
        run += c;
    }
}

// ------------------------------------------------------------------------------------------------ aggregation
// Shared-memory table of one bucket, structure of arrays: key words [NW][S], payload [S], rank [S], claim list [S] (u16).
// Claim protocol: CAS(key0, EMPTY -> w0|BUSY), write the other key words + the first instance's payload and rank, fence, store w0.
// A thread that meets a BUSY key0 whose other bits match waits for the publication (independent thread scheduling: the claimer makes
// progress even inside the same warp).  The claimer also appends the slot to the claim list (the flush walks the list, not the
// table) and prefetches the k-mer's home slot in the GLOBAL table into L2, so that the flush finds it there.
template <int NW>
struct SmemTable {
    u64* key;   // [NW * S]
    u64* pay;
    u64* rnk;
    unsigned short* list;
    u32* count;
    // Two phases with a warp barrier between them (the caller's __syncwarp over the lanes that have an instance): first every lane
    // finds or claims its slot (lanes differ only in the number of probes), then the lanes that found an existing key apply their
    // instance TOGETHER.  Without the barrier the lanes that match on their first probe leave the loop and run the (long) update
    // on their own while the others keep probing (ncu: the update code ran twice per step with 9 active lanes); a formulation of
    // the loop alone does not help, ptxas produces the same code for all of them.
    // find(): 1 = key present at idx, 2 = claimed by this lane (first instance already recorded), 3 = no room (caller spills)
    __device__ __forceinline__ int find(const Table<NW>& tab, const Kmer<NW>& k, unsigned left, unsigned right, u64 rank, u32& idx) const {
        idx = skm_slot_hash(k, SKM_LOG2_SLOTS);
        volatile u64* vkey = key;
        volatile u64* vpay = pay;
        volatile u64* vrnk = rnk;
        for (int probe = 0; probe < SKM_SLOTS; probe++) {
            u64 k0 = vkey[idx];
            if (k0 == EMPTY64) {
                if (*(volatile u32*)count >= (u32)SKM_SOFT_LIMIT) return 3;
                u64 old = atomicCAS(&key[idx], EMPTY64, k.w[0] | BUSY_BIT);
                if (old == EMPTY64) {
#pragma unroll
                    for (int w = 1; w < NW; w++) vkey[w * SKM_SLOTS + idx] = k.w[w];
                    vpay[idx] = payload_apply(PAYLOAD_FRESH, left, right);
                    vrnk[idx] = rank;
                    __threadfence_block();
                    vkey[idx] = k.w[0];
                    const u32 n = atomicAdd(count, 1u);
                    list[n] = (unsigned short)idx;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(tab.slots + (table_hash(k) & tab.mask)));
                    return 2;
                }
                k0 = old;
            }
            if ((k0 & ~BUSY_BIT) == k.w[0]) {
                while (k0 & BUSY_BIT) k0 = vkey[idx];
                bool same = true;
#pragma unroll
                for (int w = 1; w < NW; w++) same = same && vkey[w * SKM_SLOTS + idx] == k.w[w];
                if (same) return 1;
            }
            idx = (idx + 1) & (SKM_SLOTS - 1);
        }
        return 3;
    }
    __device__ __forceinline__ void apply(u32 idx, unsigned left, unsigned right, u64 rank) const {
        volatile u64* vpay = pay;
        volatile u64* vrnk = rnk;
        u64 cur = vpay[idx];
        for (;;) {
            u64 nxt = payload_apply(cur, left, right);
            if (nxt == cur) break;
            u64 old = atomicCAS(&pay[idx], cur, nxt);
            if (old == cur) break;
            cur = old;
        }
        if (rank < vrnk[idx]) atomicMin(&rnk[idx], rank);
    }
};

// one aggregated k-mer -> the global table (find or claim, merge the saturating counters, keep the smallest rank)
template <int NW>
__device__ __forceinline__ bool table_merge(const Table<NW>& t, const Kmer<NW>& k, u64 agg, u64 rank) {
    bool claimed;
    u64 idx = table_find_or_claim(t, k, &claimed);
    Slot<NW>* s = t.slots + idx;
    u64 cur = claimed ? PAYLOAD_FRESH : ldcg64(&s->payload);
    for (;;) {
        u64 nxt = payload_merge(cur, agg);
        if (nxt == cur) break;
        u64 old = atomicCAS(&s->payload, cur, nxt);
        if (old == cur) break;
        cur = old;
    }
    atomicMin(&s->aux, rank);
    return claimed;
}

constexpr int SKM_MAX_CHUNKS = 64;
constexpr int SKM_HALF_WARPS = SKM_APPLY_THREADS / 16;

// words j-1 .. j+K of one read for the lane's k-mer (skm_instance): NW+2 consecutive words, zero past the read
template <int NW>
__device__ __forceinline__ void skm_load_words(const u64* __restrict__ wp, int W64, int j, u64 (&buf)[NW + 2]) {
    const int w0 = skm_first_word<NW>(j);
#pragma unroll
    for (int x = 0; x < NW + 2; x++) buf[x] = w0 + x < W64 ? __ldg(wp + w0 + x) : 0ull;
}

// One CTA per bucket (buckets handed out dynamically).  A half warp handles one run record per step, lane t the record's k-mer
// t, taken directly from the packed read (skm_instance: no rolling, so the 16 lanes are independent).  Record s+2 and the read
// words of record s+1 are loaded while record s is processed.
template <int NW>
__global__ void __launch_bounds__(SKM_APPLY_THREADS) k_skm_apply(Table<NW> tab, KParams<NW> kp, const SkmChunkDev* __restrict__ chunks, int n_chunks, int W64,
                                                                u32 b0, u32 b1, u32* next_bucket, u64* counters, int dbg) {
    extern __shared__ __align__(16) u64 s_tab[];   // key[NW][S], pay[S], rnk[S], list[S] (u16)
    __shared__ SkmChunkDev s_chunk[SKM_MAX_CHUNKS];
    __shared__ u32 s_cum[SKM_MAX_CHUNKS + 1], s_off[SKM_MAX_CHUNKS];
    __shared__ u32 s_count, s_bucket;
    __shared__ unsigned s_new, s_spill;
    SmemTable<NW> st{s_tab, s_tab + NW * SKM_SLOTS, s_tab + (NW + 1) * SKM_SLOTS, reinterpret_cast<unsigned short*>(s_tab + (NW + 2) * SKM_SLOTS), &s_count};
    if (threadIdx.x < n_chunks) s_chunk[threadIdx.x] = chunks[threadIdx.x];
    if (threadIdx.x == 0) { s_new = 0; s_spill = 0; }
    for (int i = threadIdx.x; i < SKM_SLOTS; i += SKM_APPLY_THREADS) st.key[i] = EMPTY64;   // the flush re-empties what it merges
    unsigned my_new = 0, my_spill = 0;
    const int hw = threadIdx.x >> 4, t = threadIdx.x & 15;
    for (;;) {
        __syncthreads();   // previous bucket fully flushed (and s_chunk / the empty table visible on the first trip)
        if (threadIdx.x == 0) {
            s_bucket = b0 + atomicAdd(next_bucket, 1u);
            s_count = 0;
        }
        __syncthreads();
        const u32 b = s_bucket;
        if (b >= b1) break;
        if (threadIdx.x < n_chunks) {
            const u32* so = s_chunk[threadIdx.x].segoff;
            u32 lo = so[b], hi = so[b + 1];
            s_off[threadIdx.x] = lo;
            s_cum[threadIdx.x + 1] = hi - lo;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            u32 acc = 0;
            s_cum[0] = 0;
            for (int c = 0; c < n_chunks; c++) { acc += s_cum[c + 1]; s_cum[c + 1] = acc; }
        }
        __syncthreads();
        const u32 total = s_cum[n_chunks];
        // record gq of the bucket -> (chunk, record); gq only grows, so the chunk cursor only moves forward
        int cl = 0;
        auto next_rec = [&](u32 gq, int& c) -> u64 {
            while (gq >= s_cum[cl + 1]) cl++;
            c = cl;
            return __ldg(s_chunk[cl].recs + (u64)s_off[cl] + (gq - s_cum[cl]));
        };
        u32 g = (u32)hw;
        int c_cur = 0, c_nxt = 0;
        u64 rec = 0, rec_nxt = 0;
        u64 buf[NW + 2], nbuf[NW + 2];
#pragma unroll
        for (int x = 0; x < NW + 2; x++) { buf[x] = 0; nbuf[x] = 0; }
        if (g < total) {
            rec = next_rec(g, c_cur);
            if (t < skm_count(rec)) skm_load_words<NW>(s_chunk[c_cur].words + (u64)skm_read(rec) * W64, W64, skm_start(rec) + t, buf);
        }
        if (g + SKM_HALF_WARPS < total) rec_nxt = next_rec(g + SKM_HALF_WARPS, c_nxt);
        // The two half warps run the loop in lockstep (same trip count, __syncwarp at the end): without it the lanes that leave the
        // hash insert at different times stay split into groups of ~8 for the rest of the bucket (measured: 8.2 active threads per
        // instruction, 4x the instructions issued).
        for (;;) {
            const bool act = g < total;
            if (!__any_sync(0xffffffffu, act)) break;
            // loads for the following steps first
            u64 rec_nn = 0;
            int c_nn = 0;
            if (g + SKM_HALF_WARPS < total && t < skm_count(rec_nxt))
                skm_load_words<NW>(s_chunk[c_nxt].words + (u64)skm_read(rec_nxt) * W64, W64, skm_start(rec_nxt) + t, nbuf);
            if (g + 2 * SKM_HALF_WARPS < total) rec_nn = next_rec(g + 2 * SKM_HALF_WARPS, c_nn);
            const int n = skm_count(rec);
            const bool has = act && t < n;
            const unsigned has_mask = __ballot_sync(0xffffffffu, has);   // all 32 lanes are together here
            if (has) {
                const int j = skm_start(rec) + t;
                const SkmInst<NW> in = skm_instance<NW>(kp, buf, j, !(skm_last(rec) && t == n - 1));
                const SkmChunkDev& ch = s_chunk[c_cur];
                const u64 rank = ((ch.ord_base + (u64)skm_read(rec) * ch.ord_stride) << 16) | (u64)j;
                if (dbg >= 2) { my_spill += skm_slot_hash(in.canon, SKM_LOG2_SLOTS) + in.left + in.right; }   // PGB200_SKM_DBG=2: no table work at all (timing only)
                else {
                    u32 slot;
                    const int state = st.find(tab, in.canon, in.left, in.right, rank, slot);
                    __syncwarp(has_mask);   // the lanes re-join before the counter update (see SmemTable)
                    if (state == 1) st.apply(slot, in.left, in.right, rank);
                    else if (state == 3) {
                        // bucket holds more distinct k-mers than the shared-memory table: this instance goes to HBM directly (same result)
                        my_new += table_insert(tab, in.canon, in.left, in.right, rank);
                        my_spill++;
                    }
                }
            }
            rec = rec_nxt; c_cur = c_nxt;
            rec_nxt = rec_nn; c_nxt = c_nn;
#pragma unroll
            for (int x = 0; x < NW + 2; x++) buf[x] = nbuf[x];
            g += SKM_HALF_WARPS;
            __syncwarp();
        }
        __syncthreads();
        // flush: one global update per distinct k-mer of the bucket, walking the claim list (every thread busy)
        const u32 n_claimed = s_count;
        for (u32 i = threadIdx.x; i < n_claimed; i += SKM_APPLY_THREADS) {
            const u32 idx = st.list[i];
            Kmer<NW> k;
#pragma unroll
            for (int w = 0; w < NW; w++) k.w[w] = st.key[w * SKM_SLOTS + idx];
            if (dbg == 0) my_new += table_merge(tab, k, st.pay[idx], st.rnk[idx]);   // PGB200_SKM_DBG=1: no global merge (timing only)
            st.key[idx] = EMPTY64;
        }
    }
    if (my_new) atomicAdd(&s_new, my_new);
    if (my_spill) atomicAdd(&s_spill, my_spill);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_new) atomicAdd(&counters[C_DISTINCT], (u64)s_new);
        if (s_spill) atomicAdd(&counters[C_MISC2], (u64)s_spill);
    }
}

// ------------------------------------------------------------------------------------------------ host side
template <int NW>
static constexpr size_t skm_apply_smem() { return (size_t)(NW + 2) * SKM_SLOTS * sizeof(u64) + (size_t)SKM_SLOTS * sizeof(unsigned short); }
static u64 next_pow2_u64(u64 x) { u64 p = 1; while (p < x) p <<= 1; return p; }

template <int NW>
void* EngineT<NW>::skm_alloc(size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    while (skm_blk_ < skm_blocks_.size() && skm_used_ + bytes > skm_blocks_[skm_blk_].second) { skm_blk_++; skm_used_ = 0; }
    if (skm_blk_ >= skm_blocks_.size()) {
        size_t blk = std::max<size_t>(bytes, (size_t)512 << 20);
        void* p = nullptr;
        PG_CUDA(cudaMalloc(&p, blk));
        skm_blocks_.push_back({p, blk});
        skm_blk_ = skm_blocks_.size() - 1;
        skm_used_ = 0;
    }
    void* r = static_cast<char*>(skm_blocks_[skm_blk_].first) + skm_used_;
    skm_used_ += bytes;
    return r;
}

template <int NW>
void EngineT<NW>::skm_init() {
    if (skm_geom_.n_buckets) return;
    u64 est = 0;   // expected number of distinct k-mers
    if (prm_.table_slots) est = prm_.table_slots / 2;
    else if (prm_.initG) est = (u64)((double)prm_.P * (double)ref_static_set_size(prm_.initG, prm_.P, prm_.flavour127 != 0) * 0.77);
    u64 B = est ? next_pow2_u64((est + SKM_SLOTS / 2 - 1) / (SKM_SLOTS / 2)) : (1ull << 16);
    if (B < 1024) B = 1024;
    if (B > (1ull << 20)) B = 1ull << 20;
    if (const char* e = getenv("PGB200_SKM_BUCKETS")) B = strtoull(e, nullptr, 0);
    if (B < 1) B = 1;
    if (B > (1ull << 20)) B = 1ull << 20;   // the side buffer packs the bucket in 20 bits
    skm_geom_ = make_skm_geom(prm_.K, (u32)B);
    skm_inst_.alloc(B * sizeof(u64));
    skm_cursor_.alloc((B + 1) * sizeof(u32));
    skm_desc_.alloc(SKM_MAX_CHUNKS * sizeof(SkmChunkDev) + 256);
    PG_CUDA(cudaMemsetAsync(skm_inst_.p, 0, B * sizeof(u64), st_));
    for (auto& e : ev_skm_) PG_CUDA(cudaEventCreate(&e));
    skm_part_threads_ = SKM_PART_THREADS;
    while ((size_t)skm_part_threads_ * 2 * skm_geom_.w * sizeof(u32) > 160 * 1024 && skm_part_threads_ > 32) skm_part_threads_ /= 2;
    size_t ring = (size_t)skm_part_threads_ * 2 * skm_geom_.w * sizeof(u32);
    if (ring > 48 * 1024) {
        PG_CUDA(cudaFuncSetAttribute(k_skm_part<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring));
        PG_CUDA(cudaFuncSetAttribute(k_skm_part<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring));
    }
    PG_CUDA(cudaFuncSetAttribute(k_skm_apply<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)skm_apply_smem<NW>()));
    if (prm_.verbose) fprintf(stderr, "[pgb200] aggregated pass 1: %u buckets, minimizer length %d, window %d\n", skm_geom_.n_buckets, skm_geom_.m, skm_geom_.w);
}

// count the runs of the chunk just decoded (inside the caller's "insert" event bracket)
template <int NW>
void EngineT<NW>::skm_count_chunk(size_t ci) {
    skm_init();
    const ReadChunk& ch = chunks_[ci];
    const u32 B = skm_geom_.n_buckets;
    SkmPending pd;
    pd.chunk = ci;
    pd.segoff = reinterpret_cast<u32*>(skm_alloc((B + 1) * sizeof(u32)));
    PG_CUDA(cudaMemsetAsync(pd.segoff, 0, (B + 1) * sizeof(u32), st_));
    const size_t ring = (size_t)skm_part_threads_ * 2 * skm_geom_.w * sizeof(u32);
    const unsigned blocks = (unsigned)std::min<u64>((ch.n_rec + skm_part_threads_ - 1) / skm_part_threads_, 148ull * 16);
    skm_side_.ensure(ch.n_rec * (SKM_SIDE_RUNS * sizeof(u32) + 1) + 256);
    u32* side = skm_side_.template as<u32>();
    u8* nruns = reinterpret_cast<u8*>(side + ch.n_rec * SKM_SIDE_RUNS);
    k_skm_part<false><<<blocks, skm_part_threads_, ring, st_>>>(skm_geom_, ch.words, ch.len, ch.n_rec, W64_, pd.segoff, skm_inst_.template as<u64>(), nullptr, nullptr, side, nruns);
    PG_CUDA(cudaGetLastError());
    k_skm_offsets<<<1, 1024, 0, st_>>>(pd.segoff, B, d_cnt_ + C_MISC1);
    PG_CUDA(cudaGetLastError());
    skm_pending_.push_back(pd);
    skm_unscattered_ = true;
    p1_.launches += 2;
}

// write the records of the last counted chunk; `total` = its record count (read by the caller's host sync)
template <int NW>
void EngineT<NW>::skm_scatter_last(u64 total) {
    if (!skm_unscattered_) return;
    SkmPending& pd = skm_pending_.back();
    const ReadChunk& ch = chunks_[pd.chunk];
    const u32 B = skm_geom_.n_buckets;
    pd.n_recs = total;
    pd.recs = reinterpret_cast<u64*>(skm_alloc((total ? total : 1) * sizeof(u64)));
    PG_CUDA(cudaMemsetAsync(skm_cursor_.p, 0, (B + 1) * sizeof(u32), st_));
    const size_t ring = (size_t)skm_part_threads_ * 2 * skm_geom_.w * sizeof(u32);
    const unsigned blocks = (unsigned)std::min<u64>((ch.n_rec + skm_part_threads_ - 1) / skm_part_threads_, 148ull * 16);
    u32* side = skm_side_.template as<u32>();
    u8* nruns = reinterpret_cast<u8*>(side + ch.n_rec * SKM_SIDE_RUNS);
    u32* cursor = skm_cursor_.template as<u32>();
    k_skm_scatter_side<<<(unsigned)std::min<u64>((ch.n_rec + 255) / 256, 148ull * 16), 256, 0, st_>>>(side, nruns, ch.n_rec, cursor, pd.segoff, pd.recs, cursor + B);
    PG_CUDA(cudaGetLastError());
    // reads with more than SKM_SIDE_RUNS runs (rare): full re-scan, restricted to those reads
    k_skm_part<true><<<blocks, skm_part_threads_, ring, st_>>>(skm_geom_, ch.words, ch.len, ch.n_rec, W64_, cursor, nullptr, pd.segoff, pd.recs, side, nruns);
    PG_CUDA(cudaGetLastError());
    p1_.launches += 1;
    skm_unscattered_ = false;
    skm_pending_recs_ += total;
    p1_.launches += 1;
}

// aggregate every pending chunk into the global table
template <int NW>
void EngineT<NW>::skm_flush() {
    if (skm_pending_.empty()) return;
    settle_timing();
    if (skm_unscattered_) {
        read_counters();
        PG_CUDA(cudaEventRecord(ev_skm_[0], st_));
        skm_scatter_last(h_cnt_[C_MISC1]);
    } else {
        PG_CUDA(cudaEventRecord(ev_skm_[0], st_));
    }
    const u32 B = skm_geom_.n_buckets;
    std::vector<u64> inst(B);
    PG_CUDA(cudaMemcpyAsync(inst.data(), skm_inst_.p, B * sizeof(u64), cudaMemcpyDeviceToHost, st_));
    std::vector<SkmChunkDev> desc;
    for (auto& pd : skm_pending_) {
        const ReadChunk& ch = chunks_[pd.chunk];
        desc.push_back(SkmChunkDev{ch.words, ch.len, pd.recs, pd.segoff, ch.ord_base, ch.ord_stride});
    }
    if (desc.size() > (size_t)SKM_MAX_CHUNKS) throw std::runtime_error("pgb200: internal: too many pending chunks in skm_flush");
    PG_CUDA(cudaMemcpyAsync(skm_desc_.p, desc.data(), desc.size() * sizeof(SkmChunkDev), cudaMemcpyHostToDevice, st_));
    u32* d_next = reinterpret_cast<u32*>(static_cast<char*>(skm_desc_.p) + SKM_MAX_CHUNKS * sizeof(SkmChunkDev));
    create_table_if_needed();
    PG_CUDA(cudaMemsetAsync(d_cnt_ + C_MISC2, 0, sizeof(u64), st_));   // spilled instances (bucket larger than the shared-memory table)
    read_counters();   // also completes the two copies above
    u32 b0 = 0;
    const size_t smem = skm_apply_smem<NW>();
    int ranges = 0;
    while (b0 < B) {
        // as many buckets as the table has guaranteed room for (every instance could be a new key)
        const u64 have = h_cnt_[C_DISTINCT];
        double room = 0.80 * (double)cap_ - (double)have;
        if (room < 0.25 * (double)cap_) room = 0.25 * (double)cap_;
        u64 sum = 0;
        u32 b1 = b0;
        while (b1 < B && (b1 == b0 || (double)(sum + inst[b1]) <= room)) sum += inst[b1++];
        ensure_table_bound(have, sum);
        if (getenv("PGB200_SKM_STATS")) fprintf(stderr, "[pgb200]   range %d: buckets %u..%u, %llu instance(s), %llu distinct before\n", ranges, b0, b1, (unsigned long long)sum, (unsigned long long)have);
        PG_CUDA(cudaMemsetAsync(d_next, 0, sizeof(u32), st_));
        const unsigned blocks = (unsigned)std::min<u64>((u64)(b1 - b0), 148ull * (NW == 2 ? 3 : 2));
        k_skm_apply<NW><<<blocks, SKM_APPLY_THREADS, smem, st_>>>(tab_, kp_, reinterpret_cast<const SkmChunkDev*>(skm_desc_.p), (int)desc.size(), W64_, b0, b1, d_next, d_cnt_, getenv("PGB200_SKM_DBG") ? atoi(getenv("PGB200_SKM_DBG")) : 0);
        PG_CUDA(cudaGetLastError());
        p1_.launches += 1;
        ranges++;
        b0 = b1;
        if (b0 < B) read_counters();
    }
    PG_CUDA(cudaEventRecord(ev_skm_[1], st_));
    PG_CUDA(cudaMemsetAsync(skm_inst_.p, 0, B * sizeof(u64), st_));
    PG_CUDA(cudaEventSynchronize(ev_skm_[1]));
    float ms;
    PG_CUDA(cudaEventElapsedTime(&ms, ev_skm_[0], ev_skm_[1]));
    p1_.ms_insert += ms;
    if (prm_.verbose >= 2 || getenv("PGB200_SKM_STATS")) {
        read_counters();
        fprintf(stderr, "[pgb200] aggregated %zu chunk(s), %llu records, %u buckets in %d range(s): %.2f ms; %llu instance(s) spilled past the shared-memory tables; %llu distinct, table %llu slots\n",
                skm_pending_.size(), (unsigned long long)skm_pending_recs_, B, ranges, ms, (unsigned long long)h_cnt_[C_MISC2], (unsigned long long)h_cnt_[C_DISTINCT], (unsigned long long)cap_);
    }
    skm_pending_.clear();
    skm_pending_recs_ = 0;
    skm_blk_ = 0;
    skm_used_ = 0;
}

template <int NW>
void EngineT<NW>::skm_reset() {
    skm_pending_.clear();
    skm_pending_recs_ = 0;
    skm_unscattered_ = false;
    skm_blk_ = 0;
    skm_used_ = 0;
    if (skm_geom_.n_buckets) PG_CUDA(cudaMemsetAsync(skm_inst_.p, 0, (size_t)skm_geom_.n_buckets * sizeof(u64), st_));
}

template <int NW>
void EngineT<NW>::skm_release() {
    for (auto& b : skm_blocks_) cudaFree(b.first);
    skm_blocks_.clear();
    for (auto& e : ev_skm_) if (e) cudaEventDestroy(e);
}


template void* EngineT<2>::skm_alloc(size_t); template void* EngineT<4>::skm_alloc(size_t);
template void EngineT<2>::skm_init(); template void EngineT<4>::skm_init();
template void EngineT<2>::skm_count_chunk(size_t); template void EngineT<4>::skm_count_chunk(size_t);
template void EngineT<2>::skm_scatter_last(u64); template void EngineT<4>::skm_scatter_last(u64);
template void EngineT<2>::skm_flush(); template void EngineT<4>::skm_flush();
template void EngineT<2>::skm_reset(); template void EngineT<4>::skm_reset();
template void EngineT<2>::skm_release(); template void EngineT<4>::skm_release();

}   // namespace pgb
