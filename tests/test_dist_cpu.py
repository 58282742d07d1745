"""CPU (gloo, world_size 2): the variable-size all-to-all used for the owner exchange, with host tensors standing in for the
device tuple buffers.  Checks that every byte lands on the rank that owns it, in source-rank order."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from soapdenovo2_b200 import dist as pdist


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tb = 32
    # rank r sends (r + 1) * (o + 2) tuples to owner o; payload byte = 16 * r + o
    ranges = [0]
    parts = []
    for o in range(world):
        n = (rank + 1) * (o + 2)
        ranges.append(ranges[-1] + n)
        parts.append(torch.full((n * tb,), 16 * rank + o, dtype=torch.uint8))
    send = torch.cat(parts)
    recv, recv_bytes = pdist.all_to_all_bytes(torch, dist, send, pdist.split_sizes(ranges, tb), "cpu")
    ok = recv_bytes == [(s + 1) * (rank + 2) * tb for s in range(world)]
    pos = 0
    for s in range(world):
        seg = recv[pos:pos + recv_bytes[s]]
        ok = ok and bool((seg == 16 * s + rank).all())
        pos += recv_bytes[s]
    out[rank] = int(ok)
    dist.destroy_process_group()


def test_owner_exchange_gloo_world2():
    world = 2
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(world, 29533, out), nprocs=world, join=True)
    assert dict(out) == {0: 1, 1: 1}
