#!/usr/bin/env python
"""bench.py -- distinct k-mers hashed / second at K=63 (pregraph pass 1) on N B200s, next to the reference's CPU path.

One "step" = one complete pass 1 (prlRead2HashTable equivalent: FASTQ text -> 2-bit reads -> canonical k-mers ->
table insert/count -> delow/mark-linear/kmerFreq sweeps) over the synthetic read set of BASELINE.json configs[1]
(100 Mbp genome, 30x, 150 bp PE FASTQ, K=63), including clearing the table from the previous step.

  value  : whole-job distinct k-mers / s with the FASTQ text already resident in HBM (device pointers through the C-ABI)
  e2e    : the same metric through the C-ABI with HOST (pinned) text buffers: H2D copies inside the timed region, plus a
           D2H read of the coverage histogram / statistics every step
  roofline: the insert, HBM bound: algorithmic bytes = 64 B per k-mer instance (one 32 B slot sector read + written back),
           SURVEY.md 8(d); time = CUDA events recorded by the engine on its own stream around every launch of the insert kernels.
           With the text resident in HBM the insert is the aggregated one (k_skm_part + k_skm_apply, soapdenovo2_b200/csrc/skm.cu:
           super-k-mer buckets, one table update per DISTINCT k-mer); host text (e2e) and N>1 use the per-instance kernels
  cpu_baseline: the UNMODIFIED reference binary (oracle/_ref/SOAPdenovo-63mer pregraph, built from /root/reference by
           oracle/Makefile) timed on this box's host cores up to its "done hashing nodes" line, on a bounded sample
N>1 (torchrun): the k-mer space is sharded by an owner hash; see DESIGN.md (e).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K = 63
RD_LEN = 150
INSERT = 300
ERR = 0.001
NAME_W = 10   # "@" + 10 chars
REC_BYTES = 1 + NAME_W + 1 + RD_LEN + 1 + 2 + RD_LEN + 1


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--genome", type=int, default=int(os.environ.get("PGB200_BENCH_GENOME", 100_000_000)))
    ap.add_argument("--coverage", type=float, default=30.0)
    ap.add_argument("--chunk-reads", type=int, default=1_000_000)
    ap.add_argument("--sample-genome", type=int, default=2_500_000, help="cpu_baseline sample: sub-genome size at the same coverage")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--K", type=int, default=63, help="63 (headline, 63-mer flavour) or e.g. 127 (127-mer flavour, 256-bit keys; configs[3] shape)")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------- synthetic reads (torch = plumbing)
def gen_pe_fastq_gpu(torch, dev, genome_len, n_pairs, seed):
    """Two uint8 device tensors holding FASTQ text (fixed 316-byte records) for mates 1 and 2."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    comp = torch.zeros(256, dtype=torch.uint8, device=dev)
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    genome = acgt[torch.randint(0, 4, (genome_len,), device=dev, generator=g)]
    out = []
    ar = torch.arange(RD_LEN, device=dev)
    BATCH = 2_000_000
    bufs = [torch.empty((n_pairs, REC_BYTES), dtype=torch.uint8, device=dev) for _ in range(2)]
    for b0 in range(0, n_pairs, BATCH):
        nb = min(BATCH, n_pairs - b0)
        starts = torch.randint(0, genome_len - INSERT + 1, (nb,), device=dev, generator=g)
        r1 = genome[starts[:, None] + ar[None, :]]
        r2 = comp[genome[(starts + INSERT - RD_LEN)[:, None] + ar[None, :]].long()].flip(1)
        flip = torch.rand(nb, device=dev, generator=g) < 0.5
        m1 = torch.where(flip[:, None], r2, r1)
        m2 = torch.where(flip[:, None], r1, r2)
        for mate, reads in enumerate((m1, m2)):
            errm = torch.rand(reads.shape, device=dev, generator=g) < ERR
            code = ((reads >> 1) & 3).long()               # A(0x41)->0 C(0x43)->1 T(0x54)->2 G(0x47)->3
            lut = torch.tensor([0, 1, 3, 2], device=dev)    # -> index in "ACGT"
            sub = acgt[(lut[code] + torch.randint(1, 4, reads.shape, device=dev, generator=g)) % 4]   # always a different letter
            reads = torch.where(errm, sub, reads)
            rec = bufs[mate][b0:b0 + nb]
            rec[:, 0] = ord("@")
            ids = torch.arange(b0, b0 + nb, device=dev)
            rec[:, 1] = ord("r")
            for d in range(NAME_W - 1):
                rec[:, 1 + NAME_W - 1 - d] = ((ids // (10 ** d)) % 10 + 48).to(torch.uint8)
            o = 1 + NAME_W
            rec[:, o] = 10
            rec[:, o + 1:o + 1 + RD_LEN] = reads
            rec[:, o + 1 + RD_LEN] = 10
            rec[:, o + 2 + RD_LEN] = ord("+")
            rec[:, o + 3 + RD_LEN] = 10
            rec[:, o + 4 + RD_LEN:o + 4 + 2 * RD_LEN] = ord("I")
            rec[:, o + 4 + 2 * RD_LEN] = 10
    del genome
    return bufs[0].reshape(-1), bufs[1].reshape(-1)


# ----------------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.stop, self.index = [], False, index
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop:
            try:
                o = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(self.rows)}


# ----------------------------------------------------------------------------------------------- reference CPU arm
def write_sample_files(torch, workdir, genome_len, coverage, seed):
    n_pairs = int(genome_len * coverage / (2 * RD_LEN))
    if n_pairs % 8192 == 0:
        n_pairs -= 1   # file size must not be a multiple of 32768 B (reference AIO reader quirk)
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    t1, t2 = gen_pe_fastq_gpu(torch, dev, genome_len, n_pairs, seed)
    p1, p2 = os.path.join(workdir, "s_1.fq"), os.path.join(workdir, "s_2.fq")
    t1.cpu().numpy().tofile(p1)
    t2.cpu().numpy().tofile(p2)
    cfg = os.path.join(workdir, "s.cfg")
    with open(cfg, "w") as f:
        f.write(f"max_rd_len={RD_LEN}\n[LIB]\navg_ins={INSERT}\nreverse_seq=0\nasm_flags=3\nrank=1\nq1={p1}\nq2={p2}\n")
    return cfg, n_pairs


def time_reference_pass1(cfg, workdir, threads, tag):
    """Run the unmodified reference pregraph and time it from launch to its 'done hashing nodes' stderr line (= pass 1)."""
    ref = os.path.join(ROOT, "oracle", "_ref", "SOAPdenovo-63mer")
    kind = "reference"
    if not os.path.exists(ref):
        ref, kind = os.path.join(ROOT, "oracle", "pregraph_model_63"), "port"
        if not os.path.exists(ref):
            subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "model"], check=True)
    if kind == "reference":
        cmd = [ref, "pregraph", "-s", cfg, "-K", str(K), "-p", str(threads), "-a", "2", "-o", os.path.join(workdir, tag)]
    else:
        cmd, threads = [ref, "-1", "-s", cfg, "-K", str(K), "-p", "8", "-a", "2", "-o", os.path.join(workdir, tag)], 1
    t0 = time.perf_counter()
    p = subprocess.Popen(cmd, stderr=subprocess.PIPE, stdout=subprocess.DEVNULL, text=True)
    distinct, t_done = None, None
    for line in p.stderr:
        if "node(s) allocated" in line:
            distinct = int(line.split()[0])
            t_done = time.perf_counter()   # printed right after the last batch was inserted (prlHashReads.c:717)
            if kind == "reference":
                p.kill()   # the exact child we started; later phases are not part of the metric
                break
    p.wait()
    if distinct is None:
        raise RuntimeError("reference run produced no 'node(s) allocated' line")
    return distinct, t_done - t0, kind, threads


def main():
    args = parse_args()
    global K
    K = args.K
    import torch
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    workdir = os.environ.get("PGB200_BENCH_DIR", "/tmp/pgb200_bench")
    os.makedirs(workdir, exist_ok=True)
    workload = f"synthetic {args.genome/1e6:g} Mbp genome, {args.coverage:g}x {RD_LEN} bp PE FASTQ (insert {INSERT}, {ERR*100:g}% subst.), K={K}"

    if args.impl == "reference":
        if rank != 0:
            return
        cores = os.cpu_count() or 1
        thr = min(cores, 8) if cores < 16 else min(cores // 2, 64)
        cfg, n_pairs = write_sample_files(torch, workdir, args.sample_genome, args.coverage, seed=4242)
        vals = []
        for i in range(args.warmup + args.steps):
            d, secs, kind, used = time_reference_pass1(cfg, workdir, thr, "refarm")
            if i >= args.warmup:
                vals.append((d, secs))
        d = vals[0][0]
        secs = sum(s for _, s in vals) / len(vals)
        v = d / secs
        sample = f"{args.sample_genome/1e6:g} Mbp sub-genome at {args.coverage:g}x ({2*n_pairs} reads, {2*n_pairs*(RD_LEN-K+1)} k-mer instances, {d} distinct), pass 1 only"
        print(json.dumps({"impl": "reference", "metric": "distinct k-mers hashed/sec at K=63", "value": v, "unit": "distinct k-mers/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs * 1e3, "higher_is_better": True, "scaling": "strong",
                          "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": {"workload": workload, "sample": sample},
                          "cpu_baseline": {"value": v, "unit": "distinct k-mers/s", "cores": used, "kind": kind, "sample": sample},
                          "e2e": {"value": v, "unit": "distinct k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    from soapdenovo2_b200 import api
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (the product has no CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- workload: every rank generates the SAME read set (same seed); the k-mer space is sharded by owner hash
    n_pairs = int(args.genome * args.coverage / (2 * RD_LEN))
    t1, t2 = gen_pe_fastq_gpu(torch, dev, args.genome, n_pairs, seed=42)
    torch.cuda.synchronize()
    text_bytes = t1.numel() + t2.numel()
    est_distinct = int(args.genome * (2.3 if K <= 63 else 3.6)) + 1_000_000   # ~K error k-mers per substitution
    slots = 1 << max(20, (int(est_distinct / world * 2.2 * float(os.environ.get('PGB200_BENCH_SLOTS_MULT', '1')))).bit_length())
    eng = api.PregraphEngine(K=K, P=8, initG=0, flavour127=int(K > 63), max_rd_len=RD_LEN, device=local_rank, table_slots=slots, world=world, rank=rank,
                             verbose=int(os.environ.get("PGB200_VERBOSE", "0")))
    chunk = args.chunk_reads * REC_BYTES

    from soapdenovo2_b200 import dist as pdist

    def one_step(bufs, on_device):
        """bufs: per mate either a device tensor or (host pointer, nbytes).  N>1: chunk i is fed by rank i % N, then one
        bucketed all-to-all moves every (k-mer, links, rank) tuple to its owner rank (total work fixed: strong scaling)."""
        eng.reset_pass1()
        work = []
        for mate, t in enumerate(bufs):
            total = t.numel() if on_device else t[1]
            base = t.data_ptr() if on_device else t[0]
            off = 0
            while off < total:
                n = min(chunk, total - off)
                work.append((base + off, n, (off // REC_BYTES) * 2 + mate))
                off += n
        pipe = xchg if world > 1 else None
        for r0 in range(0, len(work), world):
            i = r0 + rank
            if i < len(work):
                ptr, n, ob = work[i]
                eng.feed_text(ptr, n, on_device=on_device, fastq=True, ord_base=ob, ord_stride=2)
            if pipe:
                pipe.round()
        if pipe:
            pipe.finish()
        st = eng.finish_pass1()
        hist, lin, rem = eng.sweeps()   # D2H of the histogram + counters: the step's result
        return st, hist

    xchg = None
    if world > 1:
        mode = os.environ.get("PGB200_XCHG", "nccl")   # "fused": peer stores over NVLink from the bucketing kernel (dist.FusedExchange)
        xchg = (pdist.FusedExchange(eng, torch, dist, dev, cap_tuples=int(2.2 * args.chunk_reads * (RD_LEN - K + 1)))
                if mode == "fused" else pdist.PipelinedExchange(eng, torch, dist, dev))

    def timed(bufs, on_device, steps, warmup):
        for _ in range(warmup):
            st, hist = one_step(bufs, on_device)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ins_ms, launches = 0.0, 0
        e0.record()
        t0 = time.perf_counter()
        for _ in range(steps):
            st, hist = one_step(bufs, on_device)
            ins_ms += st.ms_insert
            launches += st.launches + 3
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        dev_ms = e0.elapsed_time(e1)
        ms = max(dev_ms, 0.0) if dev_ms > 0 else wall * 1e3
        if world > 1:
            tt = torch.tensor([ms, float(st.distinct), float(st.instances), ins_ms], device=dev, dtype=torch.float64)
            mx = tt.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = tt.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            ms, distinct, instances, ins_ms = mx[0].item(), sm[1].item(), sm[2].item(), mx[3].item()
            hh = torch.tensor(hist, device=dev, dtype=torch.int64); dist.all_reduce(hh); hist = hh.tolist()
        else:
            distinct, instances = st.distinct, st.instances
        return ms / steps, distinct, instances, ins_ms / steps, launches // steps, hist, st

    with ClockSampler(local_rank) as cs:
        ms_step, distinct, instances, ins_ms, launches, hist, st = timed((t1, t2), True, args.steps, args.warmup)
    clocks = cs.summary()
    value = distinct / (ms_step / 1e3)

    e2e = None
    if not args.no_e2e:
        # the same step from HOST pinned buffers through the C-ABI (H2D inside), result read back every step
        lib = api.load()
        hb = []
        for t in (t1, t2):
            p = lib.pgb200_host_alloc(t.numel())
            import ctypes
            arr = (ctypes.c_ubyte * t.numel()).from_address(p)
            torch.frombuffer(arr, dtype=torch.uint8).copy_(t.cpu())
            hb.append((p, t.numel()))
        ms_e, d_e, i_e, _, _, _, _ = timed(hb, False, max(1, args.steps), 1)
        e2e = {"value": d_e / (ms_e / 1e3), "unit": "distinct k-mers/s", "ms_per_step": ms_e, "h2d_bytes_per_step": text_bytes,
               "d2h_bytes_per_step": 256 * 8 + 16 * 8 * 25}
        for p, _ in hb:
            lib.pgb200_host_free(p)

    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    inst_per_rank = instances / world
    slot_bytes = 64 if K <= 63 else 128   # one slot sector read + written back
    achieved = inst_per_rank * slot_bytes / (ins_ms / 1e3) / 1e9 if ins_ms > 0 else None
    traffic = None
    aggregated = world == 1 and os.environ.get("PGB200_SKM", "auto") != "0"
    try:
        tf = "r01_skm_apply_traffic.json" if aggregated else "r01_insert_traffic.json"
        traffic = json.load(open(os.path.join(ROOT, "profiles", tf))).get("dram_bytes_per_launch") if K == 63 and world == 1 else None
    except Exception:
        pass
    kname = "k_skm_apply<%d> (+ k_skm_part: aggregated insert)" if aggregated else ("k_chop_insert<%d>" if world == 1 else "k_apply_tuples<%d>")
    roof = {"kernel": kname % (2 if K <= 63 else 4), "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if achieved else None,
            "traffic": traffic, "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
            "algorithmic_bytes_per_instance": slot_bytes, "instances_per_step_per_gpu": inst_per_rank, "insert_kernel_ms_per_step": ins_ms}

    cpu_b = None
    if not args.no_cpu_baseline:
        try:
            cores = os.cpu_count() or 1
            thr = min(cores, 8) if cores < 16 else min(cores // 2, 64)
            cfg, sp = write_sample_files(torch, workdir, args.sample_genome, args.coverage, seed=4242)
            d, secs, kind, used = time_reference_pass1(cfg, workdir, thr, "cpub")
            cpu_b = {"value": d / secs, "unit": "distinct k-mers/s", "cores": used, "kind": kind, "host_cores_available": cores,
                     "sample": f"{args.sample_genome/1e6:g} Mbp sub-genome at {args.coverage:g}x ({2*sp} reads, {2*sp*(RD_LEN-K+1)} instances, {d} distinct), pass 1 only, {secs:.2f} s"}
        except Exception as ex:   # keep the GPU line even if the CPU leg fails
            cpu_b = {"value": None, "error": str(ex)[:200]}

    print(json.dumps({
        "metric": f"distinct k-mers hashed/sec at K={K}", "value": value, "unit": "distinct k-mers/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic",
        "config": {"workload": workload, "reads": 2 * n_pairs, "kmer_instances": int(instances), "distinct_kmers": int(distinct),
                   "instances_per_s": instances / (ms_step / 1e3), "table_slots_per_gpu": int(st.table_slots), "parallelism": f"k-mer space sharded over {world} GPU(s) by owner hash" + (", tuples stored straight into the owner GPU over NVLink by the bucketing kernel (PGB200_XCHG=nccl: NCCL all-to-all)" if world > 1 else ""),
                   "insert_mode": ("value: aggregated (super-k-mer buckets, one table update per distinct k-mer); e2e: per-instance inserts overlapped with the H2D copies" if world == 1 and os.environ.get("PGB200_SKM", "auto") == "auto" else ("PGB200_SKM=" + os.environ.get("PGB200_SKM", "") if world == 1 else "per-instance tuples exchanged between owners")),
                   "l2_policy": "inputs (6.3 GB text, 17 GB table) are far larger than the 126 MB L2; the table is cleared every step"},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu_b}))


if __name__ == "__main__":
    main()
