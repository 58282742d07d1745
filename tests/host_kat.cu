// Host-side known-answer harness for the engine's k-mer algebra / CRC / table-geometry helpers (compiled with nvcc, runs on the CPU).
// Prints "name value" lines that tests/test_kats.py compares with tests/golden/kats.json.
#include "../soapdenovo2_b200/csrc/engine_impl.cuh"
#include <cstdio>
#include <cstring>
using namespace pgb;
template <int NW>
static Kmer<NW> from_str(const char* s, KParams<NW>& kp) {
    Kmer<NW> k = kzero<NW>();
    for (const char* p = s; *p; p++) k = knext(k, (unsigned)((*p & 6) >> 1), kp);
    return k;
}
int main() {
    const char* s = "ACGTTGCATGCAAGCTTAGCTAGGATCCATCGATCGGGCTATATCGCGATTAGCCATGCAGGT";
    KParams<2> kp = make_kparams<2>(63);
    Kmer<2> f = from_str<2>(s, kp), r = krc_n(f, 63);
    printf("hash_zero_63 0x%llx\n", crc_hash(kzero<2>()));
    printf("kmer63_fwd 0x%llx 0x%llx\n", f.w[0], f.w[1]);
    printf("kmer63_rc 0x%llx 0x%llx\n", r.w[0], r.w[1]);
    printf("kmer63_smaller_fwd_rc %d\n", kless(f, r) ? 1 : 0);
    printf("kmer63_hash_fwd 0x%llx\n", crc_hash(f) & 0xffffffffull);
    printf("kmer63_hash_rc 0x%llx\n", crc_hash(r) & 0xffffffffull);
    // the same value in the 127-mer build sits in the low words: leading zero bytes do not change an init-0 CRC
    Kmer<4> f4 = kzero<4>(); f4.w[2] = f.w[0]; f4.w[3] = f.w[1];
    printf("kmer63_hash_fwd_mer127 0x%llx\n", crc_hash(f4) & 0xffffffffull);
    // rolling update == rebuilding: prevKmer on the complement strand
    Kmer<2> rr = kzero<2>();
    for (const char* p = s; *p; p++) rr = kprev(rr, (unsigned)(((*p & 6) >> 1) ^ 2), kp);
    printf("rolling_rc_matches %d\n", keq(rr, r) ? 1 : 0);
    // table geometry: init_kmerset(1024, .77f) and init_kmerset(3*0xFFFFFF, .77f)
    unsigned long long req[2] = {1024ull, 3ull * 0xFFFFFFull};
    for (int i = 0; i < 2; i++) {
        unsigned long long sz = ref_next_prime(req[i]);
        printf("init_kmerset_%d %llu %llu\n", i, sz, (unsigned long long)(sz * 0.77f));
    }
    printf("static_set_size_a1_p3_63 %llu\n", ref_static_set_size(1, 3, false));
    printf("static_set_size_a3_p1_127 %llu\n", ref_static_set_size(3, 1, true));
    printf("home_63 %llu\n", ref_home(f, 1031, false));
    printf("home_127flavour %llu\n", ref_home(f, 1031, true));
    printf("sizeof_slot %zu %zu\n", sizeof(Slot<2>), sizeof(Slot<4>));
    // K=127 (K+1)-mer reverse complement quirk: only the lowest word is touched
    Kmer<4> q; q.w[0] = 0x1111111111111111ull; q.w[1] = 0x2222222222222222ull; q.w[2] = 0x3333333333333333ull; q.w[3] = 0x0123456789abcdefull;
    Kmer<4> qq = krc_ref(q, 128, true);
    printf("rc128_quirk %d 0x%llx\n", (qq.w[0] == q.w[0] && qq.w[1] == q.w[1] && qq.w[2] == q.w[2]) ? 1 : 0, qq.w[3]);
    return 0;
}
