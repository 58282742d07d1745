import os, sys, torch
sys.path.insert(0, '/root/repo')
import torch.distributed as dist, torch.multiprocessing as mp
from soapdenovo2_b200 import api, synth, dist as pdist

def worker(rank, world, d):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT="29555", RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    eng = api.PregraphEngine(K=63, P=8, initG=1, max_rd_len=150, device=rank, world=world, rank=rank)
    work = []
    for fn, mate in ((d + "/pe_1.fq", 0), (d + "/pe_2.fq", 1)):
        data = open(fn, "rb").read(); lines = data.split(b"\n")[:-1]
        recs = [b"\n".join(lines[i:i+4]) + b"\n" for i in range(0, len(lines), 4)]
        per = (len(recs) + 2) // 3
        for c in range(3): work.append((b"".join(recs[c*per:(c+1)*per]), (c*per)*2 + mate))
    tot_sent = tot_recv = 0
    for r0 in range(0, len(work), world):
        i = r0 + rank
        if i < len(work): eng.feed_text(work[i][0], fastq=True, ord_base=work[i][1], ord_stride=2)
        ptr, ranges, tb = eng.exchange_buffer()
        n = pdist.exchange_round(eng, torch, dist, dev)
        tot_sent += ranges[-1]; tot_recv += n
        print(f"rank {rank} round {r0}: ranges {ranges} recv {n}", flush=True)
    st = eng.finish_pass1()
    print(f"rank {rank}: sent {tot_sent} recv {tot_recv} distinct {st.distinct} instances {st.instances}", flush=True)
    dist.destroy_process_group()

if __name__ == "__main__":
    d = "/tmp/dbgm"; os.makedirs(d, exist_ok=True)
    synth.scenario_pe_fastq(d)
    mp.spawn(worker, args=(2, d), nprocs=2, join=True)
