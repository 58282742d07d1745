"""Multi-GPU plumbing for pass 1: the bucketed all-to-all of (k-mer, links, rank) tuples between owner ranks.

torch.distributed (NCCL over NVLink on the GPUs, gloo in the CPU tests) is only the transport; the tuples are produced by
k_bucket_scatter and consumed by k_apply_tuples inside libpregraph_b200.so (include/pregraph_b200.h, pgb200_exchange_buffer /
pgb200_apply_tuples).  One round = every rank has fed (at most) one chunk; rank r then receives every tuple whose owner
hash maps to r.  This is the exchange step SURVEY.md 8(e) describes; the reference's analogue is the per-thread owner filter
`hash % thrd_num == id` over a shared batch (prlHashReads.c:79-90).
"""
from __future__ import annotations


class DeviceMemory:
    """Expose a raw device pointer to torch (zero-copy) through __cuda_array_interface__."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def split_sizes(ranges, tuple_bytes):
    """[world+1] tuple range starts -> per-owner byte counts."""
    return [(ranges[o + 1] - ranges[o]) * tuple_bytes for o in range(len(ranges) - 1)]


def all_to_all_bytes(torch, dist, send, send_bytes, device):
    """Variable-size all-to-all of a flat uint8 tensor. Returns (recv tensor, per-source byte counts)."""
    world = dist.get_world_size()
    sc = torch.tensor(send_bytes, dtype=torch.int64, device=device)
    rc = torch.empty_like(sc)
    dist.all_to_all_single(rc, sc)
    recv_bytes = rc.tolist()
    recv = torch.empty(sum(recv_bytes), dtype=torch.uint8, device=device)
    dist.all_to_all_single(recv, send, output_split_sizes=recv_bytes, input_split_sizes=list(send_bytes))
    assert len(recv_bytes) == world
    return recv, recv_bytes


def exchange_round(eng, torch, dist, device):
    """Ship the owner-grouped tuples of the chunk this rank just fed (or nothing) and insert what this rank owns."""
    ptr, ranges, tb = eng.exchange_buffer()
    send_bytes = split_sizes(ranges, tb)
    total = ranges[-1] * tb
    send = torch.as_tensor(DeviceMemory(ptr, total), device=device) if total else torch.empty(0, dtype=torch.uint8, device=device)
    recv, recv_bytes = all_to_all_bytes(torch, dist, send, send_bytes, device)
    if recv.is_cuda:
        # the collective runs on torch's NCCL stream, the engine on its own stream: the received tuples must have landed (and
        # the send buffer must be free for the next chunk) before the engine touches either
        torch.cuda.current_stream(recv.device).synchronize()
    n = sum(recv_bytes) // tb
    if n:
        eng.apply_tuples(recv.data_ptr(), n)
    eng.exchange_clear()
    return n


class PipelinedExchange:
    """Same exchange, software-pipelined: the all-to-all of round i is in flight (NCCL stream, async) while the engine applies the
    tuples received in round i-1 and buckets the chunk of round i+1 (the engine alternates between two tuple buffers)."""

    def __init__(self, eng, torch, dist, device):
        self.eng, self.torch, self.dist, self.device = eng, torch, dist, device
        self.pending = None

    def _drain(self):
        if self.pending is None:
            return 0
        recv, n, work, send = self.pending
        work.wait()
        if recv.is_cuda:
            self.torch.cuda.current_stream(recv.device).synchronize()
        if n:
            self.eng.apply_tuples(recv.data_ptr(), n)
        self.pending = None
        return n

    def round(self):
        torch, dist = self.torch, self.dist
        ptr, ranges, tb = self.eng.exchange_buffer()
        send_bytes = split_sizes(ranges, tb)
        total = ranges[-1] * tb
        send = torch.as_tensor(DeviceMemory(ptr, total), device=self.device) if total else torch.empty(0, dtype=torch.uint8, device=self.device)
        sc = torch.tensor(send_bytes, dtype=torch.int64, device=self.device)
        rc = torch.empty_like(sc)
        dist.all_to_all_single(rc, sc)
        recv_bytes = rc.tolist()
        recv = torch.empty(sum(recv_bytes), dtype=torch.uint8, device=self.device)
        work = dist.all_to_all_single(recv, send, output_split_sizes=recv_bytes, input_split_sizes=list(send_bytes), async_op=True)
        self.eng.exchange_clear()
        done = self._drain()                       # apply the previous round while this round's bytes move
        self.pending = (recv, sum(recv_bytes) // tb, work, send)
        return done

    def finish(self):
        return self._drain()


class FusedExchange:
    """The owner exchange WITHOUT a library collective on the data path: k_bucket_scatter stores every tuple straight into its
    owner's receive buffer (CUDA IPC peer mapping, NVLink stores) while it chops the next reads.  torch.distributed only carries
    the 64-byte IPC handles once, an 8-byte-per-pair count matrix per round, and the barrier between "all stores issued" and
    "apply what arrived".  Two receive buffers alternate, so round i+1 may be scattered while a slow peer still applies round i."""

    def __init__(self, eng, torch, dist, device, cap_tuples):
        self.eng, self.torch, self.dist, self.device = eng, torch, dist, device
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        eng.xchg_setup(cap_tuples)
        for buf in (0, 1):
            mine = torch.tensor(list(eng.xchg_export(buf)), dtype=torch.uint8, device=device)
            allh = [torch.empty(64, dtype=torch.uint8, device=device) for _ in range(self.world)]
            dist.all_gather(allh, mine)
            for p in range(self.world):
                if p != self.rank:
                    eng.xchg_import(p, buf, bytes(allh[p].cpu().tolist()))
        self.buf = 0

    def round(self):
        torch, dist = self.torch, self.dist
        mine = torch.tensor(self.eng.xchg_counts(), dtype=torch.int64, device=self.device)
        rows = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(rows, mine)
        m = torch.stack(rows).cpu()                       # m[s][o] = tuples rank s holds for owner o
        base = m[: self.rank].sum(dim=0).tolist() if self.rank else [0] * self.world
        n_recv = int(m[:, self.rank].sum())
        self.eng.xchg_scatter(self.buf, [int(b) for b in base])   # returns when this rank's peer stores are performed
        dist.barrier()                                    # ... and now everybody's are
        self.eng.xchg_apply(self.buf, n_recv)
        self.eng.exchange_clear()
        self.buf ^= 1
        return n_recv

    def finish(self):
        self.dist.barrier()
        return 0
