"""CPU (gloo, world_size 2): the host-side plumbing of the multi-GPU pass 1 -- the handle exchange and the per-epoch
fence / barrier / flush protocol of soapdenovo2_b200.dist.RecordExchange, against a mock engine that records the call order
(the data path itself is peer stores inside the CUDA library and is covered by tests/test_gpu_multi.py), and the work split."""
import os

import torch.distributed as dist
import torch.multiprocessing as mp

from soapdenovo2_b200 import dist as pdist


class MockEngine:
    def __init__(self, rank):
        self.rank, self.calls, self.imported = rank, [], {}

    def xchg_setup(self, cap):
        self.calls.append(("setup", cap))

    def xchg_export(self):
        return bytes([self.rank]) * 64

    def xchg_import(self, peer, handle):
        self.imported[peer] = handle

    def xchg_fence(self):
        self.calls.append("fence")

    def flush(self):
        self.calls.append("flush")


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = MockEngine(rank)
    x = pdist.RecordExchange(eng, dist, cap_records=1234)
    x.end_epoch()
    x.end_epoch()
    ok = eng.calls == [("setup", 1234), "fence", "flush", "fence", "flush"]
    ok = ok and eng.imported == {p: bytes([p]) * 64 for p in range(world) if p != rank}
    out[rank] = int(ok)
    dist.destroy_process_group()


def test_record_exchange_protocol_gloo_world2():
    world = 2
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(world, 29533, out), nprocs=world, join=True)
    assert dict(out) == {0: 1, 1: 1}


def test_deal_covers_every_chunk_once():
    for n in (0, 1, 7, 20, 24):
        for w in (1, 2, 3, 8):
            got = sorted(i for r in range(w) for i in pdist.deal(n, w, r))
            assert got == list(range(n))
