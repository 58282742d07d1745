"""2-GPU parity (run with gpurun --gpus 2): pass 1 with the k-mer space sharded by owner hash and the bucketed NCCL
all-to-all must give the same coverage histogram, distinct count and per-k-mer entries as the single-GPU oracle."""
import os

import pytest
import torch

from soapdenovo2_b200 import api, synth
from tests import util

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, d, out, mode="nccl"):
    import torch.distributed as dist
    from soapdenovo2_b200 import dist as pdist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    eng = api.PregraphEngine(K=63, P=8, initG=1, max_rd_len=150, device=rank, world=world, rank=rank)
    files = [(os.path.join(d, "pe_1.fq"), 0), (os.path.join(d, "pe_2.fq"), 1)]
    # split every file into 3 record-aligned chunks; chunk i is fed by rank i % world
    work = []
    for fn, mate in files:
        data = open(fn, "rb").read()
        lines = data.split(b"\n")[:-1]
        recs = [b"\n".join(lines[i:i + 4]) + b"\n" for i in range(0, len(lines), 4)]
        per = (len(recs) + 2) // 3
        for c in range(3):
            part = recs[c * per:(c + 1) * per]
            work.append((b"".join(part), (c * per) * 2 + mate))
    fused = pdist.FusedExchange(eng, torch, dist, dev, cap_tuples=1 << 20) if mode == "fused" else None
    for r0 in range(0, len(work), world):
        i = r0 + rank
        if i < len(work):
            eng.feed_text(work[i][0], fastq=True, ord_base=work[i][1], ord_stride=2)
        if fused:
            fused.round()
        else:
            pdist.exchange_round(eng, torch, dist, dev)
    st = eng.finish_pass1()
    hist, lin, rem = eng.sweeps()
    h = torch.tensor(hist, device=dev, dtype=torch.int64)
    dist.all_reduce(h)
    cnt = torch.tensor([st.distinct, st.instances], device=dev, dtype=torch.int64)
    dist.all_reduce(cnt)
    eng.build_layout()      # per-shard layout: records are only compared as a SET below
    out[rank] = (h.tolist(), cnt.tolist(), eng.dump_nodes())
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("mode", ["nccl", "fused"])
def test_two_gpu_sharded_pass1(tmp_path, mode):
    import torch.multiprocessing as mp
    util.build_oracle()
    d = str(tmp_path)
    cfg = synth.scenario_pe_fastq(d)
    mod, dump = os.path.join(d, "mod"), os.path.join(d, "mod.table")
    util.run_model(util.MODEL63, cfg, mod, 63, 8, ("-1", "-T", dump, "-a", "1"))
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(2, 29541 if mode == "nccl" else 29543, d, out, mode), nprocs=2, join=True)
    hist, cnt, _ = out[0]
    assert api.kmerfreq_text(hist) == open(mod + ".kmerFreq", "rb").read()
    want = open(dump, "rb").read()
    assert cnt[0] * 26 == len(want) and cnt[1] == 12000 * 88
    recs = lambda b: {b[i:i + 26] for i in range(0, len(b), 26)}
    # every k-mer lives on exactly one rank with exactly the oracle's counters (linear/deleted flags included)
    assert recs(out[0][2]) | recs(out[1][2]) == recs(want)
    assert not (recs(out[0][2]) & recs(out[1][2]))
