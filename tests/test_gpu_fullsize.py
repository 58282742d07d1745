"""Full-size configs[1] (100 Mbp genome, 30x, 150 bp PE FASTQ, K=63: 1.76e9 k-mer instances) checked through size-independent
properties -- the oracle cannot finish this size in seconds, the properties can be checked exactly:
  * instances == sum over reads of (len - K + 1)
  * sum(hist) == distinct;  sum_c c * hist[c] == instances while nothing saturates at 255 (checksum of all coverage counters)
  * the result does not depend on how the text is cut into chunks or on the order the chunks are fed (first-occurrence ranks
    are carried by stream ordinals, not by feeding order)
  * strand symmetry: reverse-complementing every read leaves every canonical k-mer's entry unchanged (10 Mbp case, full dump)
"""
import os
import sys

import pytest

from soapdenovo2_b200 import api

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _feed_all(eng, torch, t1, t2, bench, chunk_reads, order="forward"):
    work = []
    for mate, t in enumerate((t1, t2)):
        total, off = t.numel(), 0
        while off < total:
            n = min(chunk_reads * bench.REC_BYTES, total - off)
            work.append((t.data_ptr() + off, n, (off // bench.REC_BYTES) * 2 + mate))
            off += n
    if order == "reverse":
        work.reverse()
    for ptr, n, ob in work:
        eng.feed_text(ptr, n, on_device=True, fastq=True, ord_base=ob, ord_stride=2)
    st = eng.finish_pass1()
    hist, lin, rem = eng.sweeps()
    return st, hist, lin


def test_configs1_full_size_properties():
    import torch
    import bench
    G, K = 100_000_000, 63
    n_pairs = int(G * 30 / 300)
    t1, t2 = bench.gen_pe_fastq_gpu(torch, torch.device("cuda", 0), G, n_pairs, seed=42)
    torch.cuda.synchronize()   # the engine's streams are non-blocking: they do not wait for torch's stream by themselves
    eng = api.PregraphEngine(K=K, P=8, max_rd_len=150, table_slots=1 << 29)
    st, hist, lin = _feed_all(eng, torch, t1, t2, bench, 1_000_000)
    assert st.records == 2 * n_pairs and st.reads_kept == 2 * n_pairs
    assert st.instances == 2 * n_pairs * (150 - K + 1)
    assert sum(hist) == st.distinct and hist[0] == 0
    assert hist[255] == 0 and sum(c * h for c, h in enumerate(hist)) == st.instances
    assert 1.9e8 < st.distinct < 2.3e8          # ~1e8 genomic + ~63 error k-mers per substitution
    # different chunking, reversed feeding order: identical histogram and linear-node count
    eng.reset_pass1()
    st2, hist2, lin2 = _feed_all(eng, torch, t1, t2, bench, 3_333_333, order="reverse")
    assert (st2.distinct, st2.instances, hist2, lin2) == (st.distinct, st.instances, hist, lin)
    eng.close()


def test_strand_symmetry_10mbp():
    import torch
    import bench
    G, K = 10_000_000, 63
    n_pairs = int(G * 30 / 300)
    dev = torch.device("cuda", 0)
    t1, t2 = bench.gen_pe_fastq_gpu(torch, dev, G, n_pairs, seed=7)
    comp = torch.arange(256, dtype=torch.uint8, device=dev)
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b

    def revcomp_text(t):
        rec = t.view(-1, bench.REC_BYTES).clone()
        o = 1 + bench.NAME_W + 1
        rec[:, o:o + bench.RD_LEN] = comp[rec[:, o:o + bench.RD_LEN].long()].flip(1)
        return rec.reshape(-1)

    dumps = []
    for a, b in ((t1, t2), (revcomp_text(t1), revcomp_text(t2))):
        torch.cuda.synchronize()
        eng = api.PregraphEngine(K=K, P=4, initG=4, max_rd_len=150)
        st, hist, lin = _feed_all(eng, torch, a, b, bench, 1_000_000)
        eng.build_layout()
        d = eng.dump_nodes()
        dumps.append((st.distinct, hist, sorted(d[i:i + 26] for i in range(0, len(d), 26))))
        eng.close()
    assert dumps[0][0] == dumps[1][0] and dumps[0][1] == dumps[1][1]
    assert dumps[0][2] == dumps[1][2]           # same (k-mer, 8 link counters, coverage, flags) multiset; only the ORDER may differ
