"""torch.distributed plumbing for the one-process-per-GPU launch (torchrun): rank identity, the few host-side
collectives the driver needs, and shipping the 128-byte communicator id from rank 0 (what MPI_Init did for
the reference, main.cpp:78-98).  Works with the `nccl` backend on GPUs and with `gloo` on CPU (tests)."""
import os

import numpy as np


class Ranks:
    def __init__(self, backend=None):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.backend = backend
        self._dist = None
        if self.world > 1:
            import torch
            import torch.distributed as dist
            self._torch = torch
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            self.backend = backend
            if not dist.is_initialized():
                kw = {}
                if backend == "nccl":
                    kw["device_id"] = torch.device("cuda", self.local_rank)
                dist.init_process_group(backend, **kw)
            self._dist = dist

    def _dev(self):
        return "cuda" if self.backend == "nccl" else "cpu"

    def barrier(self):
        if self._dist:
            self._dist.barrier()

    def allreduce(self, x, op="sum"):
        if not self._dist:
            return x
        t = self._torch.tensor([x], dtype=self._torch.float64, device=self._dev())
        self._dist.all_reduce(t, op=self._dist.ReduceOp.MAX if op == "max" else self._dist.ReduceOp.SUM)
        return float(t.item())

    def broadcast_bytes(self, payload, nbytes, src=0):
        """rank `src` passes `payload` (bytes of length nbytes); every rank gets it back."""
        if not self._dist:
            return payload
        t = self._torch.zeros(nbytes, dtype=self._torch.uint8, device=self._dev())
        if self.rank == src:
            t.copy_(self._torch.frombuffer(bytearray(payload), dtype=self._torch.uint8))
        self._dist.broadcast(t, src)
        return bytes(t.cpu().numpy().tobytes())

    def gather_arrays(self, arr):
        """all-gather of variable-length 1-D numpy arrays (tests / small host data only)."""
        if not self._dist:
            return [arr]
        out = [None] * self.world
        self._dist.all_gather_object(out, arr)
        return out

    def shutdown(self):
        if self._dist and self._dist.is_initialized():
            self._dist.destroy_process_group()


def strip_parts(nv_total, world):
    """Vertex ranges of the reference's 1-D distribution (graph.hpp:112-113)."""
    return np.array([(nv_total * r) // world for r in range(world + 1)], dtype=np.int64)


def my_strip(nv_total, world, rank, **kw):
    """The shard reference rank `rank` of `world` would build for `miniVite -n nv_total` (only this strip is built)."""
    from . import hostgraph as hg
    return hg.generate_rgg(nv_total, world, rank, rank + 1, **kw)
