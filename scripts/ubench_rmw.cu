// Micro-benchmark behind profiles/r01_rmw_ubench.md: what bounds a random read-modify-write of one 32-byte table slot on this part?
// Build:  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/ubench_rmw scripts/ubench_rmw.cu
// Every thread derives its slot index from a hash of its global op number (no index traffic).  `window` < table restricts the ops in
// flight to a sliding window of the table (op i hits  base(i) + rand % window, base advancing linearly over the table): that is the
// access pattern of region-bucketed tuples applied in region order.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ u64 mix(u64 x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}
struct P { u64* tab; u64 slots; u64 win_slots; u64 n; };

__device__ __forceinline__ u64 slot_of(const P& p, u64 i) {
    u64 r = mix(i * 0x9E3779B97F4A7C15ull + 12345);
    if (p.win_slots >= p.slots) return r & (p.slots - 1);
    // sliding window: base moves over (slots - win) as i goes 0..n
    u64 base = (u64)((double)i / (double)p.n * (double)(p.slots - p.win_slots));
    return base + (r & (p.win_slots - 1));
}

template <int MODE>
__global__ void __launch_bounds__(256) k_op(P p, u64* sink) {
    u64 acc = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (u64)gridDim.x * blockDim.x) {
        u64* s = p.tab + slot_of(p, i) * 4;
        if (MODE == 0) {          // hash only
            acc += (u64)s;
        } else if (MODE == 1) {   // 32 B load
            u64 a, b, c, d;
            asm volatile("ld.relaxed.gpu.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(s));
            acc += a + b + c + d;
        } else if (MODE == 2) {   // blind RED.ADD.64
            atomicAdd(s + 2, 1ull);
        } else if (MODE == 3) {   // 32 B load -> dependent CAS.64 (the insert kernel's hit path)
            u64 a, b, c, d;
            asm volatile("ld.relaxed.gpu.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(s));
            u64 cur = c;
            for (;;) { u64 old = atomicCAS(s + 2, cur, cur + 1); if (old == cur) break; cur = old; }
            acc += a + b + d;
        } else if (MODE == 4) {   // 32 B load -> plain 8 B store (exclusive owner, no atomics)
            u64 a, b, c, d;
            asm volatile("ld.relaxed.gpu.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(s));
            asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(s + 2), "l"(c + 1) : "memory");
            acc += a + b + d;
        } else if (MODE == 5) {   // 32 B load -> 32 B store (whole slot rewritten: no partial-sector write)
            u64 a, b, c, d;
            asm volatile("ld.relaxed.gpu.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(s));
            asm volatile("st.relaxed.gpu.global.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(s), "l"(a), "l"(b), "l"(c + 1), "l"(d) : "memory");
        } else if (MODE == 6) {   // blind 32 B store
            asm volatile("st.relaxed.gpu.global.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(s), "l"(i), "l"(i), "l"(i), "l"(i) : "memory");
        }
    }
    if (acc == 0x1234567) *sink = acc;
}

// shared-memory owned region: the CTA loads `region_slots` contiguous slots, applies `per_region` random RMWs with shared-memory
// atomics, writes the region back (the round-2 candidate: exclusive ownership of table regions).
__global__ void __launch_bounds__(512) k_region(P p, u64 region_slots, u64 per_region, u64* sink) {
    extern __shared__ u64 sm[];
    const u64 n_regions = p.slots / region_slots;
    for (u64 r = blockIdx.x; r < n_regions; r += gridDim.x) {
        const uint4* src = reinterpret_cast<const uint4*>(p.tab + r * region_slots * 4);
        uint4* dst = reinterpret_cast<uint4*>(sm);
        for (u64 j = threadIdx.x; j < region_slots * 2; j += blockDim.x) dst[j] = src[j];
        __syncthreads();
        for (u64 j = threadIdx.x; j < per_region; j += blockDim.x) {
            u64 h = mix((r * per_region + j) * 0x9E3779B97F4A7C15ull + 777);
            u64* s = sm + (h & (region_slots - 1)) * 4;
            if (s[0] == h) (*sink)++;   // key compare stand-in
            atomicAdd(s + 2, 1ull);
        }
        __syncthreads();
        uint4* out = reinterpret_cast<uint4*>(p.tab + r * region_slots * 4);
        for (u64 j = threadIdx.x; j < region_slots * 2; j += blockDim.x) out[j] = dst[j];
        __syncthreads();
    }
}

template <int MODE>
static float run(P p, u64* sink, int blocks) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_op<MODE><<<blocks, 256>>>(p, sink);   // warm-up
    CK(cudaEventRecord(e0));
    k_op<MODE><<<blocks, 256>>>(p, sink);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
    return ms;
}

int main(int argc, char** argv) {
    const u64 n = argc > 1 ? strtoull(argv[1], 0, 0) : (1ull << 28);
    const u64 max_bytes = argc > 2 ? strtoull(argv[2], 0, 0) << 30 : (32ull << 30);
    u64* tab; u64* sink;
    CK(cudaMalloc(&tab, max_bytes)); CK(cudaMalloc(&sink, 8));
    CK(cudaMemset(tab, 0, max_bytes));
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int blocks = sms * 8;
    const char* names[7] = {"hash_only", "load32", "red_add64", "load32+cas64", "load32+st8", "load32+st32", "st32"};
    printf("ops per launch %llu, %d SMs\n", n, sms);
    printf("%-10s %-10s", "table", "window");
    for (int m = 0; m < 7; m++) printf(" %14s", names[m]);
    printf("   (1e9 ops/s)\n");
    struct Cfg { u64 bytes, win; };
    const u64 MB = 1ull << 20, GB = 1ull << 30;
    Cfg cfgs[] = {{16 * MB, 0}, {64 * MB, 0}, {256 * MB, 0}, {1 * GB, 0}, {8 * GB, 0}, {32 * GB, 0},
                  {32 * GB, 8 * MB}, {32 * GB, 32 * MB}, {32 * GB, 64 * MB}, {32 * GB, 256 * MB}};
    for (auto c : cfgs) {
        if (c.bytes > max_bytes) continue;
        P p{tab, c.bytes / 32, c.win ? c.win / 32 : c.bytes / 32, n};
        float ms[7];
        ms[0] = run<0>(p, sink, blocks); ms[1] = run<1>(p, sink, blocks); ms[2] = run<2>(p, sink, blocks); ms[3] = run<3>(p, sink, blocks);
        ms[4] = run<4>(p, sink, blocks); ms[5] = run<5>(p, sink, blocks); ms[6] = run<6>(p, sink, blocks);
        printf("%-10llu %-10llu", c.bytes / MB, c.win / MB);
        for (int m = 0; m < 7; m++) printf(" %14.2f", (double)n / ms[m] * 1e-6);
        printf("\n");
        fflush(stdout);
    }
    // owned regions in shared memory: table 32 GB (or max), ops = n spread evenly
    CK(cudaFuncSetAttribute(k_region, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (u64 tb : {2 * GB, 8 * GB, 32 * GB}) {
        if (tb > max_bytes) continue;
        for (u64 region_kb : {64ull, 128ull}) {
            P p{tab, tb / 32, tb / 32, n};
            u64 region_slots = region_kb * 1024 / 32;
            u64 n_regions = p.slots / region_slots;
            u64 per_region = n / n_regions;
            if (!per_region) per_region = 1;
            cudaEvent_t e0, e1;
            CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            int grid = region_kb == 64 ? sms * 3 : sms;
            k_region<<<grid, 512, region_kb * 1024>>>(p, region_slots, per_region, sink);
            CK(cudaEventRecord(e0));
            k_region<<<grid, 512, region_kb * 1024>>>(p, region_slots, per_region, sink);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("owned regions: table %llu MB, region %llu KB, %llu ops/region: %.2f ms = %.2f e9 ops/s, table stream %.0f GB/s\n",
                   tb / MB, region_kb, per_region, ms, (double)(per_region * n_regions) / ms * 1e-6, 2.0 * tb / ms * 1e-6);
            fflush(stdout);
        }
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
