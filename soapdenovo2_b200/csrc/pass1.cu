// pass1.cu -- pass 1 of pregraph on the GPU: text -> 2-bit reads -> canonical k-mers -> table insert/count,
// then the per-entry sweeps (delow, mark-linear, kmerFreq histogram).
//
// Replaces (reference file:line, standardPregraph/):
//   K1  readseqInBuf / readseqfq (readseq1by1.c:138-209, 279-360), reverse2k (:788-802)    -> k_decode_pack
//   K2  chopKmer4read (prlHashReads.c:163-259)                                              -> k_chop_insert (rolling part)
//   K3  threadRoutine sig 1 + put_kmerset (prlHashReads.c:79-90, newhash.c:473-528)         -> k_chop_insert (insert part)
//   K4  thread_delow, thread_mark, freqStat (prlHashReads.c:953-996, 1020-1077, 1104-1132)  -> k_sweep
// Design differences that matter: no owner filter (the reference makes every thread scan the whole batch and keep
// hash % P == id); the CRC set hash is not computed per instance at all -- it only defines the reference's iteration
// order and is evaluated once per DISTINCT k-mer in layout.cu.
#include "engine_impl.cuh"
#include "scan.cuh"
#include "chop.cuh"
#include <ctime>

namespace pgb {

static double host_now() { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e3 + t.tv_nsec * 1e-6; }

// ------------------------------------------------------------------------------------------------ K1: line index
// element = one 16-byte group of the text; value = number of '\n' in it
struct NlIn {
    const uint4* text;
    u64 nbytes;
    __device__ __forceinline__ unsigned mask16(u64 i) const {
        uint4 v = __ldg(text + i);
        unsigned m = 0;
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            unsigned r = __vcmpeq4(w[k], 0x0A0A0A0Au) & 0x01010101u;   // exact per-byte compare
            m |= ((r | (r >> 7) | (r >> 14) | (r >> 21)) & 0xFu) << (4 * k);
        }
        u64 rem = nbytes - i * 16;
        if (rem < 16) m &= (1u << rem) - 1;
        return m;
    }
    __device__ u64 operator()(u64 i) const { return __popc(mask16(i)); }
};
// newline number g (0-based) at byte `pos`: line g+1 starts at pos+1.  Sequence line of record r is line lpr*r+1.
struct NlOut {
    NlIn in;
    u64* seq_start;
    u64* seq_end;
    u64 n_rec;
    int lpr;
    __device__ void operator()(u64 i, u64 prefix, u64 v) const {
        if (!v) return;
        unsigned m = in.mask16(i);
        u64 g = prefix;
        while (m) {
            int b = __ffs(m) - 1;
            m &= m - 1;
            u64 pos = i * 16 + b;
            // newline g ends line g and line g+1 starts at pos+1; the sequence line of record r is line lpr*r + 1
            u64 r0 = g / lpr;
            if (g - r0 * lpr == 1 && r0 < n_rec) seq_end[r0] = pos;
            u64 r1 = (g + 1) / lpr;
            if ((g + 1) - r1 * lpr == 1 && r1 < n_rec) seq_start[r1] = pos + 1;
            g++;
        }
    }
};
struct NlCountOut {
    __device__ void operator()(u64, u64, u64) const {}
};

// ------------------------------------------------------------------------------------------------ K1: decode + pack
// One warp per record.  Base code = (ch & 6) >> 1 for letters (A0 C1 T2 G3, N->3), '.' -> 0, every other byte is dropped;
// only the first min(linelen, maxlen) characters of the sequence line are considered (readseq1by1.c:177-200).
// reverse_seq: whole-read reverse complement (reverse2k).  Output: LSB-first 2-bit packing, W64 words per read.
__device__ __forceinline__ bool is_base_char(unsigned c) { return ((c | 0x20u) - 'a') < 26u || c == '.'; }
__device__ __forceinline__ unsigned base_code(unsigned c) { return c == '.' ? 0u : ((c & 6u) >> 1); }

__global__ void __launch_bounds__(256) k_decode_pack(const unsigned char* __restrict__ text, const u64* __restrict__ seq_start,
                                                     const u64* __restrict__ seq_end, u64 n_rec, int maxlen, int reverse, int K,
                                                     int W64, u64* __restrict__ words, u32* __restrict__ lens, u64* counters) {
    const int lane = threadIdx.x & 31;
    const u64 warp0 = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
    u64 inst = 0, kept = 0;
    for (u64 r = warp0; r < n_rec; r += nwarps) {
        const u64 s = seq_start[r];
        u64 e = seq_end[r];
        int raw = e > s ? (int)min((u64)(e - s), (u64)0x7fffffff) : 0;
        int use = raw < maxlen ? raw : maxlen;
        // count valid characters
        int n = 0;
        bool clean = true;
        for (int b = 0; b < use; b += 32) {
            int i = b + lane;
            bool v = i < use && is_base_char(text[s + i]);
            unsigned bal = __ballot_sync(0xffffffffu, v);
            unsigned want = (use - b) >= 32 ? 0xffffffffu : ((1u << (use - b)) - 1);
            clean = clean && (bal == want);
            n += __popc(bal);
        }
        u64* out = words + r * (u64)W64;
        if (clean) {
            for (int w = 0; w < W64; w++) {
                int oi = w * 32 + lane;   // output base index handled by this lane
                unsigned lo = 0, hi = 0;
                if (oi < n) {
                    int ii = reverse ? n - 1 - oi : oi;
                    unsigned c = base_code(text[s + ii]) ^ (reverse ? 2u : 0u);
                    if (lane < 16) lo = c << (2 * lane); else hi = c << (2 * (lane - 16));
                }
                lo = __reduce_or_sync(0xffffffffu, lo);
                hi = __reduce_or_sync(0xffffffffu, hi);
                if (lane == 0) out[w] = (u64)lo | ((u64)hi << 32);
            }
        } else {
            // rare: a byte inside the line is not a letter ('\r', digits, ...): compact with ballots + global atomicOr
            for (int w = lane; w < W64; w += 32) out[w] = 0;
            __syncwarp();
            int pos0 = 0;
            for (int b = 0; b < use; b += 32) {
                int i = b + lane;
                unsigned ch = i < use ? text[s + i] : 0;
                bool v = i < use && is_base_char(ch);
                unsigned bal = __ballot_sync(0xffffffffu, v);
                if (v) {
                    int p = pos0 + __popc(bal & ((1u << lane) - 1));
                    int oi = reverse ? n - 1 - p : p;
                    u64 c = base_code(ch) ^ (reverse ? 2u : 0u);
                    atomicOr(&out[oi >> 5], c << (2 * (oi & 31)));
                }
                pos0 += __popc(bal);
            }
        }
        if (lane == 0) {
            lens[r] = (u32)n;
            if (n >= K + 1) { inst += (u64)(n - K + 1); kept++; }   // reads shorter than K+1 are skipped (prlHashReads.c:504,642)
        }
    }
    // block-level aggregation of the counters
    __shared__ u64 s_inst, s_kept;
    if (threadIdx.x == 0) { s_inst = 0; s_kept = 0; }
    __syncthreads();
    if (lane == 0 && (inst | kept)) { atomicAdd(&s_inst, inst); atomicAdd(&s_kept, kept); }
    __syncthreads();
    if (threadIdx.x == 0 && (s_inst | s_kept)) { atomicAdd(&counters[C_INSTANCES], s_inst); atomicAdd(&counters[C_KEPT], s_kept); }
}

// ------------------------------------------------------------------------------------------------ K2+K3: chop + insert
// One thread per read: roll the forward k-mer (nextKmer) and its reverse complement (prevKmer on the complement strand)
// one base at a time, pick the canonical one, derive the left/right neighbour codes in the canonical orientation
//  and apply the instance to the table.  rank = (read ordinal << 16) | position.
// (A software-prefetch variant -- prefetch.global.L2 of the home slot 1..8 positions ahead, instances parked in a shared
//  memory ring -- was measured and dropped: 38.6 ms vs 28.0 ms per 5.3e8 instances at every distance; the kernel is bound
//  by the random-sector rate of L2/HBM, not by exposed latency.  profiles/r01_insert_ncu.md.)
constexpr int INS_THREADS = 256;
#ifndef INS_MIN_BLOCKS
#define INS_MIN_BLOCKS 5
#endif

template <int NW>
struct InsertSink {
    const Table<NW>& tab;
    u64 rank_base;
    unsigned& my_new;
    int dbg;   // 0 = the real thing; 1..4 = cost-decomposition variants for profiling (PGB200_DBG_INSERT, results are garbage)
    __device__ __forceinline__ void operator()(const Kmer<NW>& canon, unsigned left, unsigned right, int j) {
        if (dbg == 0) { my_new += table_insert(tab, canon, left, right, rank_base | (u64)j); return; }
        u64 idx = table_hash(canon) & tab.mask;
        if (dbg == 1) { my_new += (unsigned)(idx & 1) + left + right; return; }                               // ALU only
        Slot<NW>* s = tab.slots + idx;
        if (dbg == 2) { U128 v = ldcg128(s->key); my_new += (v.a == canon.w[0]); return; }                      // + one probe load
        if (dbg == 3) { atomicAdd(&s->payload, 0ull); return; }                                                // blind RED, no load
        if (dbg == 4) { U128 v = ldcg128(s->key); u64 o = atomicAdd(&s->payload, (u64)(v.a & 0)); my_new += (unsigned)(o & 1); return; }   // load -> dependent returning atomic
    }
};

template <int NW>
__global__ void __launch_bounds__(INS_THREADS, INS_MIN_BLOCKS) k_chop_insert(Table<NW> tab, KParams<NW> kp, const u64* __restrict__ words,
                                                             const u32* __restrict__ lens, u64 n_rec, int W64, u64 ord_base, u64 ord_stride,
                                                             u64* counters, int dbg, int use_tma) {
    extern __shared__ __align__(128) u64 s_words[];   // [INS_THREADS][W64] when use_tma
    __shared__ __align__(8) u64 s_bar;
    __shared__ unsigned s_new;
    if (threadIdx.x == 0) { s_new = 0; if (use_tma) mbar_init(&s_bar, 1); }
    __syncthreads();
    unsigned my_new = 0;
    const u64 n_tiles = (n_rec + INS_THREADS - 1) / INS_THREADS;
    unsigned parity = 0;
    for (u64 tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const u64 r0 = tile * INS_THREADS;
        const u64 r = r0 + threadIdx.x;
        const u64* wp = words + r * (u64)W64;
        if (use_tma) {
            u64 cnt = n_rec - r0 < (u64)INS_THREADS ? n_rec - r0 : (u64)INS_THREADS;
            unsigned bytes = (unsigned)((cnt * (u64)W64 * 8 + 15) & ~15ull);   // the arena pads every allocation to 256 B
            if (threadIdx.x == 0) {
                mbar_expect_tx(&s_bar, bytes);
                tma_bulk_g2s(s_words, words + r0 * (u64)W64, bytes, &s_bar);
            }
            mbar_wait(&s_bar, parity);
            parity ^= 1;
            wp = s_words + (u64)threadIdx.x * W64;
        }
        if (r < n_rec) {
            const int L = (int)lens[r];
            if (L >= kp.K + 1) {
                InsertSink<NW> sink{tab, (ord_base + r * ord_stride) << 16, my_new, dbg};
                chop_read(kp, wp, L, sink);
            }
        }
        if (use_tma) __syncthreads();   // the tile buffer is reused by the next bulk copy
    }
    if (my_new) atomicAdd(&s_new, my_new);
    __syncthreads();
    if (threadIdx.x == 0 && s_new) atomicAdd(&counters[C_DISTINCT], (u64)s_new);
}

// ------------------------------------------------------------------------------------------------ K2+K3, state-machine form
// ncu on k_chop_insert: ~85 % of stall samples wait on the slot load / CAS, and a quarter of them sit on instructions that only
// 1-2 lanes execute (second probes, claims, CAS retries): with one read per lane, every k-mer position costs the WARP
// max-over-lanes(probes) + claim + max-over-lanes(CAS tries) ~ 6 dependent round trips although a lane needs ~2.2 on average.
// Here every lane runs its own state machine (ADVANCE -> PROBE -> [CLAIM] -> APPLY -> ADVANCE ...) and each trip round the loop
// issues exactly ONE memory operation per lane, whatever its state; lanes drift apart by a few positions instead of
// waiting for each other, so a round trip is spent on ~32 useful operations instead of 1-2.
enum { ST_ADV = 0, ST_PROBE = 1, ST_CLAIM = 2, ST_APPLY = 3 };

__global__ void __launch_bounds__(INS_THREADS) k_chop_insert_sm2(Table<2> tab, KParams<2> kp, const u64* __restrict__ words,
                                                                 const u32* __restrict__ lens, u64 n_rec, int W64, u64 ord_base, u64 ord_stride,
                                                                 u64* counters) {
    __shared__ unsigned s_new;
    if (threadIdx.x == 0) s_new = 0;
    __syncthreads();
    const int K = kp.K;
    unsigned my_new = 0;
    for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < n_rec; r += (u64)gridDim.x * blockDim.x) {
        const int L = (int)lens[r];
        if (L < K + 1) continue;
        const u64* wp = words + r * (u64)W64;
        const u64 rank_base = (ord_base + r * ord_stride) << 16;
        Kmer<2> fwd = kzero<2>(), rc = kzero<2>();
        u64 curw = wp[0];
        int i = 0;
        // pre-roll the first K-1 bases (no k-mer yet)
        for (; i < K - 1; i++) {
            if (i && (i & 31) == 0) curw = wp[i >> 5];
            unsigned c = (unsigned)((curw >> (2 * (i & 31))) & 3);
            fwd = knext(fwd, c, kp);
            rc = kprev(rc, c ^ 2u, kp);
        }
        int state = ST_ADV;
        Kmer<2> canon = kzero<2>();
        u64 idx = 0, cur_p = 0, cur_r = 0, nxt_p = 0, rank = 0;
        unsigned left = 4, right = 4;
        for (;;) {
            if (state == ST_ADV) {
                if (i >= L) break;
                if ((i & 31) == 0) curw = wp[i >> 5];
                unsigned c = (unsigned)((curw >> (2 * (i & 31))) & 3);
                unsigned cn = 4;
                if (i + 1 < L) {
                    u64 w2 = ((i + 1) & 31) == 0 ? wp[(i + 1) >> 5] : curw;
                    cn = (unsigned)((w2 >> (2 * ((i + 1) & 31))) & 3);
                }
                unsigned dropped = kfirst(fwd, kp);
                fwd = knext(fwd, c, kp);
                rc = kprev(rc, c ^ 2u, kp);
                int j = i - K + 1;
                unsigned pv = j > 0 ? dropped : 4u;
                bool sm = kless(fwd, rc);
                canon = sm ? fwd : rc;
                left = sm ? pv : (cn < 4 ? (cn ^ 2u) : 4u);
                right = sm ? cn : (pv < 4 ? (pv ^ 2u) : 4u);
                rank = rank_base | (u64)j;
                idx = table_hash(canon) & tab.mask;
                state = ST_PROBE;
                i++;
            }
            // ---- issue: one memory operation per lane
            Slot<2>* s = tab.slots + idx;
            U256 v;
            U128 o128;
            u64 o64 = 0;
            v.a = v.b = v.c = v.d = 0;
            o128.a = o128.b = 0;
            ld256_if(state == ST_PROBE, s, v);
            cas128_if(state == ST_CLAIM, s->key, U128{EMPTY64, EMPTY64}, U128{canon.w[0], canon.w[1]}, o128);
            cas64_if(state == ST_APPLY, &s->payload, cur_p, nxt_p, o64);
            // ---- consume
            bool have_cur = false;
            if (state == ST_PROBE) {
                if (v.a == canon.w[0] && v.b == canon.w[1]) { cur_p = v.c; cur_r = v.d; have_cur = true; }
                else if (v.a == EMPTY64 && v.b == EMPTY64) state = ST_CLAIM;
                else idx = (idx + 1) & tab.mask;
            } else if (state == ST_CLAIM) {
                if (o128.a == EMPTY64 && o128.b == EMPTY64) { my_new++; cur_p = PAYLOAD_FRESH; cur_r = EMPTY64; have_cur = true; }
                else if (o128.a == canon.w[0] && o128.b == canon.w[1]) state = ST_PROBE;      // somebody else just claimed it for this key
                else { idx = (idx + 1) & tab.mask; state = ST_PROBE; }
            } else {   // ST_APPLY
                if (o64 == cur_p) { if (rank < cur_r) atomicMin(&s->aux, rank); state = ST_ADV; }
                else { cur_p = o64; have_cur = true; }
            }
            if (have_cur) {
                nxt_p = payload_apply(cur_p, left, right);
                if (nxt_p == cur_p) { if (rank < cur_r) atomicMin(&s->aux, rank); state = ST_ADV; }   // saturated: read-only
                else state = ST_APPLY;
            }
        }
    }
    if (my_new) atomicAdd(&s_new, my_new);
    __syncthreads();
    if (threadIdx.x == 0 && s_new) atomicAdd(&counters[C_DISTINCT], (u64)s_new);
}

// ------------------------------------------------------------------------------------------------ table management
template <int NW>
__global__ void k_rehash(Table<NW> oldt, Table<NW> newt) {
    u64 n = oldt.mask + 1;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const Slot<NW>* s = oldt.slots + i;
        if (!slot_occupied(s)) continue;
        Kmer<NW> k = slot_key(s);
        bool claimed;
        u64 idx = table_find_or_claim(newt, k, &claimed);
        newt.slots[idx].payload = s->payload;
        newt.slots[idx].aux = s->aux;
    }
}

static u64 next_pow2(u64 x) { u64 p = 1; while (p < x) p <<= 1; return p; }

template <int NW>
void EngineT<NW>::alloc_table(u64 cap) {
    tab_buf_.alloc(cap * sizeof(Slot<NW>));
    PG_CUDA(cudaMemsetAsync(tab_buf_.p, 0xFF, cap * sizeof(Slot<NW>), st_));
    tab_.slots = tab_buf_.template as<Slot<NW>>();
    tab_.mask = cap - 1;
    cap_ = cap;
}

template <int NW>
void EngineT<NW>::grow_table(u64 new_cap) {
    if (prm_.verbose) fprintf(stderr, "[pgb200] growing k-mer table %llu -> %llu slots\n", cap_, new_cap);
    DevBuf nb;
    nb.alloc(new_cap * sizeof(Slot<NW>));
    PG_CUDA(cudaMemsetAsync(nb.p, 0xFF, new_cap * sizeof(Slot<NW>), st_));
    Table<NW> nt{nb.template as<Slot<NW>>(), new_cap - 1};
    k_rehash<NW><<<148 * 8, 256, 0, st_>>>(tab_, nt);
    PG_CUDA(cudaGetLastError());
    sync();
    std::swap(tab_buf_.p, nb.p);
    std::swap(tab_buf_.bytes, nb.bytes);
    tab_ = nt;
    cap_ = new_cap;
}

template <int NW>
void EngineT<NW>::create_table_if_needed() {
    if (tab_.slots) return;
    u64 want = prm_.table_slots;
    if (!want) {
        if (prm_.initG) {
            // the reference's own budget: P sets of the static prime size (prlHashReads.c:369-390)
            want = (u64)prm_.P * ref_static_set_size(prm_.initG, prm_.P, prm_.flavour127 != 0);
            want = want + want / 4;
        } else {
            want = 1ull << 24;
        }
    }
    alloc_table(next_pow2(want < 1024 ? 1024 : want));
}

// have = distinct keys already in the table (exact as of the last sync), incoming = upper bound of new keys about to arrive
template <int NW>
void EngineT<NW>::ensure_table_bound(u64 have, u64 incoming) {
    create_table_if_needed();
    u64 need = have + incoming;   // every incoming instance could be a new key
    if ((double)need <= 0.80 * (double)cap_) return;
    u64 cap = cap_;
    while ((double)need > 0.80 * (double)cap) cap <<= 1;
    size_t free_b = 0, total_b = 0;
    PG_CUDA(cudaMemGetInfo(&free_b, &total_b));
    while (cap > cap_ && cap * sizeof(Slot<NW>) + (1ull << 30) > free_b) cap >>= 1;
    if (cap > cap_) { grow_table(cap); return; }
    if ((double)need > 0.97 * (double)cap_)
        throw std::runtime_error("pgb200: k-mer table cannot grow further (out of HBM); use more GPUs or a smaller batch");
}

template <int NW>
void EngineT<NW>::ensure_table(u64 incoming) {
    create_table_if_needed();
    read_counters();
    ensure_table_bound(h_cnt_[C_DISTINCT], incoming);
}

// read-store arena: chunks are carved out of large blocks (no cudaMalloc / cudaFree per chunk)
template <int NW>
void* EngineT<NW>::arena_alloc(size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (arena_.empty() || arena_used_ + bytes > arena_.back().second) {
        size_t blk = std::max<size_t>(bytes, (size_t)1 << 30);
        void* p = nullptr;
        PG_CUDA(cudaMalloc(&p, blk));
        arena_.push_back({p, blk});
        arena_used_ = 0;
    }
    void* r = static_cast<char*>(arena_.back().first) + arena_used_;
    arena_used_ += bytes;
    return r;
}

// ------------------------------------------------------------------------------------------------ feed_text

template <int NW>
void EngineT<NW>::feed_text(const char* text, size_t nbytes, bool on_device, int fastq, uint64_t ord_base, uint64_t ord_stride,
                            int reverse_seq, int maxlen) {
    double t_a = host_now(), t_b = 0, t_c = 0, t_d = 0, t_e = 0;
    last_records_ = 0;
    if (nbytes == 0) return;
    PG_CUDA(cudaSetDevice(prm_.device));
    const unsigned char* d_text;
    bool host_src = !on_device;
    if (host_src) {
        // H2D on its own stream into the buffer the previous chunk is NOT using: the copy overlaps the previous chunk's insert
        DevBuf& tb = text_bufs_[text_flip_];
        text_flip_ ^= 1;
        tb.ensure(nbytes + 16);
        PG_CUDA(cudaMemcpyAsync(tb.p, text, nbytes, cudaMemcpyHostToDevice, st_copy_));
        PG_CUDA(cudaEventRecord(ev_copy_, st_copy_));
        PG_CUDA(cudaStreamWaitEvent(st_, ev_copy_, 0));
        d_text = tb.template as<unsigned char>();
    }
    settle_timing();   // previous chunk's events (waits for its insert; the copy above is already in flight)
    PG_CUDA(cudaEventRecord(ev_[0], st_));
    if (on_device) {
        d_text = reinterpret_cast<const unsigned char*>(text);
        if ((uintptr_t)text & 15) {   // the line index reads 16-byte groups: realign with one device-to-device copy
            DevBuf& tb = text_bufs_[text_flip_];
            text_flip_ ^= 1;
            tb.ensure(nbytes + 16);
            PG_CUDA(cudaMemcpyAsync(tb.p, text, nbytes, cudaMemcpyDeviceToDevice, st_));
            d_text = tb.template as<unsigned char>();
        }
    }
    if (maxlen > prm_.max_rd_len) maxlen = prm_.max_rd_len;
    const int lpr = fastq ? 4 : 2;
    const u64 groups = (nbytes + 15) / 16;
    scan_buf_.ensure(scan_scratch_elems(groups) * sizeof(u64));
    NlIn in{reinterpret_cast<const uint4*>(d_text), (u64)nbytes};
    // pass A: count lines
    device_scan_total(in, groups, scan_buf_.template as<u64>(), d_cnt_ + C_MISC0, st_);   // tile bases stay in scan_buf_
    // ONE host sync per chunk: line count, last byte, and the counters as of the previous chunk's insert
    unsigned char* h_last = reinterpret_cast<unsigned char*>(h_cnt_ + C_COUNT);
    PG_CUDA(cudaMemcpyAsync(h_last, d_text + nbytes - 1, 1, cudaMemcpyDeviceToHost, st_));
    read_counters();
    u64 n_lines = h_cnt_[C_MISC0];
    const u64 have_distinct = h_cnt_[C_DISTINCT];
    // aggregated pass 1 (skm.cu) for text that is already resident in HBM; text arriving over PCIe is inserted chunk by chunk
    // (k_chop_insert) because those inserts hide completely under the next chunk's copy, while the aggregation would only start
    // after the last one.  PGB200_SKM=0/1 forces one path.
    const bool use_skm = prm_.world <= 1 && (skm_mode_ > 0 || (skm_mode_ < 0 && on_device));
    if (skm_unscattered_) skm_prev_total_ = h_cnt_[C_MISC1];   // record count of the previous aggregated chunk (k_skm_offsets)
    // a final line without '\n' still counts (the reference's FASTQ path tolerates it; its FASTA path does not)
    unsigned char lastc = *h_last;
    bool open_tail = lastc != '\n';
    u64 n_rec = (n_lines + (open_tail ? 1 : 0)) / lpr;
    if ((n_lines + (open_tail ? 1 : 0)) % lpr != 0)
        throw std::runtime_error("pgb200: text chunk does not hold whole FASTA/FASTQ records (line count not a multiple of 2/4)");
    if (n_rec == 0) return;
    t_b = host_now();
    line_buf_.ensure(2 * n_rec * sizeof(u64));
    u64* seq_start = line_buf_.template as<u64>();
    u64* seq_end = seq_start + n_rec;
    if (open_tail) {
        // only FASTQ can end without newline inside the quality line; the sequence line end is always a real '\n'
        PG_CUDA(cudaMemsetAsync(seq_end, 0, n_rec * sizeof(u64), st_));
    }
    device_scan_finish(in, NlOut{in, seq_start, seq_end, n_rec, lpr}, groups, scan_buf_.template as<u64>(), st_);
    if (open_tail && !fastq) {
        u64 e = nbytes;
        PG_CUDA(cudaMemcpyAsync(seq_end + n_rec - 1, &e, sizeof e, cudaMemcpyHostToDevice, st_));
        sync();
    }

    ReadChunk ch;
    ch.n_rec = n_rec;
    ch.ord_base = ord_base;
    ch.ord_stride = ord_stride;
    ch.words = reinterpret_cast<u64*>(arena_alloc(n_rec * (u64)W64_ * sizeof(u64)));
    ch.len = reinterpret_cast<u32*>(arena_alloc(n_rec * sizeof(u32)));
    chunks_.push_back(ch);
    t_c = host_now();
    {
        u64 warps = n_rec;
        unsigned blocks = (unsigned)std::min<u64>((warps + 7) / 8, 148ull * 64);
        k_decode_pack<<<blocks, 256, 0, st_>>>(d_text, seq_start, seq_end, n_rec, maxlen, reverse_seq, prm_.K, W64_, ch.words, ch.len,
                                               d_cnt_);
        PG_CUDA(cudaGetLastError());
    }
    PG_CUDA(cudaEventRecord(ev_[1], st_));
    // table capacity for the worst case of this chunk (host-side bound: no sync; growth itself syncs when it happens)
    if (xchg_fused_) create_table_if_needed();   // an apply may be in flight on the other stream: growth is decided in xchg_apply
    else if (use_skm) {
        create_table_if_needed();                 // growth is decided per bucket range in skm_flush
        if (skm_pending_.size() >= 64) { skm_scatter_last(skm_prev_total_); skm_flush(); }
    } else {
        int per_read = maxlen - prm_.K + 1;
        ensure_table_bound(have_distinct, per_read > 0 ? n_rec * (u64)per_read : 0);
    }
    t_d = host_now();
    if (l2gran_mode_ == 2) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32);
    PG_CUDA(cudaEventRecord(ev_[2], st_));
    if (prm_.world > 1) {
        bucket_chunk(ch);                       // tuples stay in the exchange buffer: caller runs the all-to-all
    } else if (use_skm) {
        skm_scatter_last(skm_prev_total_);      // records of the previous chunk (its count arrived with this chunk's host sync)
        skm_count_chunk(chunks_.size() - 1);
    } else if (batch_gb_ > 0) {
        int per_read = maxlen - prm_.K + 1;
        pending_bound_ += per_read > 0 ? n_rec * (u64)per_read : 0;
        if ((double)pending_bound_ * 32.0 >= batch_gb_ * 1e9 || pending_bound_ >= 0xF0000000ull) flush_batch();
    } else if (bucket_mode_) {
        bucket_chunk(ch);
        apply_tuples(tuple_buf().template as<u64>(), n_tuples_);
    } else {
        unsigned blocks = (unsigned)std::min<u64>((n_rec + INS_THREADS - 1) / INS_THREADS, 148ull * 64);
        if (NW == 2 && insert_sm_)
            k_chop_insert_sm2<<<blocks, INS_THREADS, 0, st_>>>(*reinterpret_cast<Table<2>*>(&tab_), *reinterpret_cast<KParams<2>*>(&kp_), ch.words, ch.len,
                                                               n_rec, W64_, ord_base, ord_stride, d_cnt_);
        else
        {
            size_t smem = (size_t)INS_THREADS * W64_ * sizeof(u64);
            int use_tma = smem <= 96 * 1024 && !getenv("PGB200_NO_TMA");
            if (use_tma && smem > 48 * 1024) cudaFuncSetAttribute(k_chop_insert<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            k_chop_insert<NW><<<blocks, INS_THREADS, use_tma ? smem : 0, st_>>>(tab_, kp_, ch.words, ch.len, n_rec, W64_, ord_base, ord_stride, d_cnt_,
                                                                              dbg_insert_, use_tma);
        }
        PG_CUDA(cudaGetLastError());
    }
    PG_CUDA(cudaEventRecord(ev_[3], st_));
    timing_pending_ = true;
    if (host_src) PG_CUDA(cudaEventSynchronize(ev_copy_));   // the caller may reuse its host buffer; the insert keeps running
    else if (prm_.world > 1) sync();                          // exchange buffer is read by the caller next
    if (l2gran_mode_ == 2) { sync(); cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 128); }
    p1_.launches += 5;   // tile sums, small scan, index apply, decode, insert
    last_records_ = n_rec;
    total_records_ += n_rec;
    t_e = host_now();
    if (prm_.verbose >= 2)
        fprintf(stderr, "[pgb200] chunk %zu: %llu rec, host ms: count %.2f alloc %.2f decode+table %.2f launch %.2f (gpu so far: decode %.2f insert %.2f)\n", chunks_.size(),
                (unsigned long long)n_rec, t_b - t_a, t_c - t_b, t_d - t_c, t_e - t_d, p1_.ms_decode, p1_.ms_insert);
}

template <int NW>
void EngineT<NW>::settle_timing() {
    if (!timing_pending_) return;
    PG_CUDA(cudaEventSynchronize(ev_[3]));
    float ms;
    PG_CUDA(cudaEventElapsedTime(&ms, ev_[0], ev_[1])); p1_.ms_decode += ms;
    PG_CUDA(cudaEventElapsedTime(&ms, ev_[2], ev_[3])); p1_.ms_insert += ms;
    timing_pending_ = false;
}

template <int NW>
void EngineT<NW>::finish_pass1(Pass1Stats* st) {
    sync_apply();
    flush_batch();
    settle_timing();
    read_counters();
    p1_.records = total_records_;
    p1_.reads_kept = h_cnt_[C_KEPT];
    p1_.instances = h_cnt_[C_INSTANCES];
    p1_.distinct = h_cnt_[C_DISTINCT];
    p1_.table_slots = cap_;
    n_nodes_ = p1_.distinct;
    if (st) *st = p1_;
}

template <int NW>
void EngineT<NW>::reset_pass1() {
    double t0 = host_now();
    sync_apply();
    settle_timing();
    sync();
    chunks_.clear();
    skm_reset();
    pending_first_ = 0;
    pending_bound_ = 0;
    // keep the first arena block for the next pass, release the rest
    while (arena_.size() > 1) { cudaFree(arena_.back().first); arena_.pop_back(); }
    arena_used_ = 0;
    total_records_ = 0;
    p1_ = Pass1Stats();
    PG_CUDA(cudaMemsetAsync(d_cnt_, 0, C_COUNT * sizeof(u64), st_));
    if (tab_.slots) PG_CUDA(cudaMemsetAsync(tab_buf_.p, 0xFF, cap_ * sizeof(Slot<NW>), st_));
    order_buf_.release();
    n_nodes_ = 0;
    sync();
    h_cnt_[C_DISTINCT] = h_cnt_[C_INSTANCES] = 0;
    if (prm_.verbose >= 2) fprintf(stderr, "[pgb200] reset_pass1: %.2f ms host\n", host_now() - t0);
}

// ------------------------------------------------------------------------------------------------ K4: sweeps
// delow (thread_delow): zero every link counter <= D, deleted=1 if nothing is left.  mark (thread_mark): linear=1 iff exactly
// one non-zero left and one non-zero right link (NO deleted check there); histogram of cov.  All per-entry => one pass.
template <int NW>
__global__ void __launch_bounds__(256) k_sweep(Table<NW> tab, int D, u64* hist, u64* counters) {
    __shared__ unsigned s_hist[256];
    __shared__ unsigned s_lin, s_rem;
    s_hist[threadIdx.x] = 0;
    if (threadIdx.x == 0) { s_lin = 0; s_rem = 0; }
    __syncthreads();
    u64 n = tab.mask + 1;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        Slot<NW>* s = tab.slots + i;
        if (!slot_occupied(s)) continue;
        u64 p = s->payload;
        if (D > 0) {
#pragma unroll
            for (int c = 0; c < 4; c++) {
                unsigned l = pl_l(p, c), r = pl_r(p, c);
                if (l > 0 && l <= (unsigned)D) p = pl_clear_l(p, c);
                if (r > 0 && r <= (unsigned)D) p = pl_clear_r(p, c);
            }
            if ((p & PL_LLINKS_MASK) == 0 && (p & PL_RLINKS_MASK) == 0) { p |= PL_DELETED; atomicAdd(&s_rem, 1u); }
        }
        atomicAdd(&s_hist[pl_cov(p)], 1u);
        if (pl_nl(p) == 1 && pl_nr(p) == 1) { p |= PL_LINEAR; atomicAdd(&s_lin, 1u); }
        s->payload = p;
    }
    __syncthreads();
    if (s_hist[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (u64)s_hist[threadIdx.x]);
    if (threadIdx.x == 0) {
        if (s_lin) atomicAdd(&counters[C_LINEAR], (u64)s_lin);
        if (s_rem) atomicAdd(&counters[C_REMOVED], (u64)s_rem);
    }
}

template <int NW>
void EngineT<NW>::sweeps(SweepStats* st) {
    double t0 = host_now();
    sync_apply();
    flush_batch();
    DevBuf hist;
    hist.alloc(256 * sizeof(u64));
    PG_CUDA(cudaMemsetAsync(hist.p, 0, 256 * sizeof(u64), st_));
    PG_CUDA(cudaMemsetAsync(d_cnt_ + C_LINEAR, 0, 2 * sizeof(u64), st_));
    int D = (int)(signed char)prm_.D;   // deLowKmer is a `char` (inc/global.h:67)
    k_sweep<NW><<<148 * 8, 256, 0, st_>>>(tab_, D, hist.template as<u64>(), d_cnt_);
    PG_CUDA(cudaGetLastError());
    u64 h[256];
    PG_CUDA(cudaMemcpyAsync(h, hist.p, sizeof h, cudaMemcpyDeviceToHost, st_));
    read_counters();
    for (int i = 0; i < 256; i++) st->hist[i] = (long long)h[i];
    st->linear = h_cnt_[C_LINEAR];
    st->removed = h_cnt_[C_REMOVED];
    if (prm_.verbose >= 2) fprintf(stderr, "[pgb200] sweeps: %.2f ms host\n", host_now() - t0);
}

// ------------------------------------------------------------------------------------------------ ctor / dtor
template <int NW>
EngineT<NW>::EngineT(const PgParams& p) : prm_(p) {
    PG_CUDA(cudaSetDevice(p.device));
    // Random 32 B slot accesses: do not let L2 promote a sector miss to a 64/128 B DRAM fetch (measured with ncu: 259 B of
    // DRAM reads per k-mer instance with the default granularity, profiles/r01_insert_ncu.md)
    if (const char* g = getenv("PGB200_L2GRAN")) l2gran_mode_ = atoi(g);
    if (const char* g = getenv("PGB200_BUCKET")) bucket_mode_ = atoi(g);
    if (const char* g = getenv("PGB200_INSERT_SM")) insert_sm_ = atoi(g);
    if (const char* g = getenv("PGB200_DBG_INSERT")) dbg_insert_ = atoi(g);
    if (const char* g = getenv("PGB200_BATCH_GB")) batch_gb_ = atof(g);
    if (const char* g = getenv("PGB200_SKM")) skm_mode_ = atoi(g) ? 1 : 0;
    if (l2gran_mode_ == 1) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32);
    kp_ = make_kparams<NW>(p.K);
    PG_CUDA(cudaStreamCreateWithFlags(&st_, cudaStreamNonBlocking));
    PG_CUDA(cudaStreamCreateWithFlags(&st_copy_, cudaStreamNonBlocking));
    PG_CUDA(cudaStreamCreateWithFlags(&st_apply_, cudaStreamNonBlocking));
    for (auto& e : ev_apply_) PG_CUDA(cudaEventCreate(&e));
    PG_CUDA(cudaEventCreateWithFlags(&ev_copy_, cudaEventDisableTiming));
    for (auto& e : ev_) PG_CUDA(cudaEventCreate(&e));
    PG_CUDA(cudaMalloc(&d_cnt_, C_COUNT * sizeof(u64)));
    PG_CUDA(cudaMemsetAsync(d_cnt_, 0, C_COUNT * sizeof(u64), st_));
    PG_CUDA(cudaHostAlloc(&h_cnt_, (C_COUNT + 2) * sizeof(u64), cudaHostAllocDefault));
    for (int i = 0; i < C_COUNT; i++) h_cnt_[i] = 0;
    W64_ = (p.max_rd_len + 31) / 32;
    if (W64_ < 1) W64_ = 1;
    sync();
}

template <int NW>
EngineT<NW>::~EngineT() {
    cudaStreamSynchronize(st_);
    for (auto& a : arena_) cudaFree(a.first);
    skm_release();
    if (d_cnt_) cudaFree(d_cnt_);
    if (h_cnt_) cudaFreeHost(h_cnt_);
    for (auto& e : ev_) if (e) cudaEventDestroy(e);
    if (st_apply_) { cudaStreamSynchronize(st_apply_); cudaStreamDestroy(st_apply_); }
    for (auto& e : ev_apply_) if (e) cudaEventDestroy(e);
    if (ev_copy_) cudaEventDestroy(ev_copy_);
    if (st_copy_) cudaStreamDestroy(st_copy_);
    if (st_) cudaStreamDestroy(st_);
}

template class EngineT<2>;
template class EngineT<4>;

IEngine* make_engine(const PgParams& p) {
    if (p.K <= 63) return new EngineT<2>(p);
    return new EngineT<4>(p);
}

}   // namespace pgb
