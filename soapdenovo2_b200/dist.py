"""Multi-process plumbing for the multi-GPU pass 1 (one process per GPU, as torchrun launches them).

The data path is inside libpregraph_b200.so: the partition kernel stores every super-k-mer record straight into the arena of the
GPU that owns its bucket (CUDA IPC peer mappings, NVLink stores; include/pregraph_b200.h, pgb200_xchg_*).  torch.distributed only
carries the 64-byte IPC handles once, and one barrier per epoch between "my records have been delivered" (pgb200_xchg_fence) and
"aggregate what I received" (pgb200_flush).  No library collective moves k-mer data.  The reference's analogue of the exchange is
the per-thread owner filter `hash % thrd_num == id` over a shared batch (prlHashReads.c:79-90).
"""
from __future__ import annotations


def exchange_handles(dist, my_handle: bytes, world: int):
    """all-gather of the 64-byte arena handles as python objects (works with the nccl and the gloo backend alike)."""
    out = [None] * world
    dist.all_gather_object(out, bytes(my_handle))
    return out


class RecordExchange:
    """Arena setup + the per-epoch fence / barrier / flush protocol for one engine of a `world`-GPU job."""

    def __init__(self, eng, dist, cap_records: int):
        self.eng, self.dist = eng, dist
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        eng.xchg_setup(cap_records)
        handles = exchange_handles(dist, eng.xchg_export(), self.world)
        for p in range(self.world):
            if p != self.rank:
                eng.xchg_import(p, handles[p])
        dist.barrier()   # nobody stores into an arena that is not mapped everywhere yet

    def end_epoch(self):
        """Every rank calls this the same number of times: deliver, wait for the others, aggregate the owned buckets."""
        self.eng.xchg_fence()
        self.dist.barrier()
        self.eng.flush()


def deal(n_items: int, world: int, rank: int):
    """Round-robin split of work items: the indices rank `rank` processes (chunk i goes to rank i % world)."""
    return list(range(rank, n_items, world))
