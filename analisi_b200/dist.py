"""One-process-per-GPU plumbing for ``bench.py`` and the multi-rank tests (torchrun launch).

torch.distributed is used for the rendezvous only: broadcasting the 128-byte NCCL unique id that
``agofrt_comm_join`` needs, the barrier around the timed region and the max-over-ranks of the device
timings.  The histograms themselves never pass through torch: they are all-reduced inside
``agofrt_block`` by the library's own NCCL communicator.

The same code runs on the ``gloo`` backend without GPUs (tests/test_multirank_gloo.py, world size 2).
"""
import os

import numpy as np


class Ranks:
    """rank / world / local_rank from the torchrun environment; world == 1 needs no torch at all."""

    def __init__(self, backend=None):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.backend = None
        self._dist = None
        self._torch = None
        if self.world > 1:
            import torch
            import torch.distributed as dist
            self._torch, self._dist = torch, dist
            self.backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
            if self.backend == "nccl":
                torch.cuda.set_device(self.local_rank)
                dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            else:
                dist.init_process_group(self.backend)

    @property
    def device(self):
        return "cuda" if self.backend == "nccl" else "cpu"

    def barrier(self):
        if self._dist is None:
            return
        if self.backend == "nccl":
            self._torch.cuda.synchronize()
        self._dist.barrier()

    def broadcast_bytes(self, payload, nbytes, src=0):
        """``payload`` (bytes of length nbytes) on rank ``src`` -> the same bytes on every rank."""
        if self._dist is None:
            return bytes(payload)
        torch = self._torch
        t = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
        if self.rank == src:
            assert len(payload) == nbytes
            t.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
        self._dist.broadcast(t, src)
        return bytes(t.cpu().numpy().tobytes())

    def max_over_ranks(self, x):
        if self._dist is None:
            return float(x)
        t = self._torch.tensor([float(x)], dtype=self._torch.float64, device=self.device)
        self._dist.all_reduce(t, op=self._dist.ReduceOp.MAX)
        return float(t.item())

    def sum_counts(self, counts):
        """Integer sum of per-rank partial histograms (tests of the shard geometry on gloo; the product
        path sums on the GPUs inside agofrt_block)."""
        if self._dist is None:
            return counts
        t = self._torch.from_numpy(np.ascontiguousarray(counts).astype(np.int64)).to(self.device)
        self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM)
        return t.cpu().numpy().astype(np.uint64)

    def close(self):
        if self._dist is not None:
            self._dist.destroy_process_group()
            self._dist = None


def join_communicator(ranks, ctx):
    """Rank 0 makes the NCCL unique id, everybody joins: afterwards ``Plan.block`` shards its work
    units over all ranks and all-reduces the counts (agofrt_comm_unique_id / agofrt_comm_join)."""
    from . import cabi
    if ranks.world <= 1:
        return
    uid = cabi.Context.unique_id() if ranks.rank == 0 else b""
    uid = ranks.broadcast_bytes(uid, cabi.COMM_ID_BYTES, 0)
    ctx.join(uid, ranks.rank, ranks.world)
