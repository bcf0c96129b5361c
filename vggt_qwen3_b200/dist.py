"""Data-parallel plumbing for stage-1 training of the projector: one process per GPU, the visual-encode path is
replicated (scenes are independent) and the ONLY exchange is the Perceiver gradient -- the reference's
equivalent is DeepSpeed ZeRO-3 / DDP (src/train/train_sft.py:119-133,165-170,217; configs/deepspeed_zero3.json:7-14).

The exchange is ZeRO-1 shaped and lives in `optim.ShardedAdamW`: reduce-scatter of the bf16 gradient -> AdamW on the
shard this rank owns -> all-gather of the updated bf16 weights.  This module provides the pieces under it:

  init_from_env / shard_batch      torchrun environment, scene sharding
  SymmetricMemoryTransport         one NVLink-symmetric allocation per rank (torch.distributed._symmetric_memory: cuMem
                                   memory mapped into every peer); peers' shards are PULLED with cudaMemcpyAsync, i.e. on
                                   the copy engines: no SM is taken from the persistent GEMM / attention grids that run
                                   concurrently (an NCCL kernel holding even a few SMs pushes part of every persistent
                                   grid into a second wave; measured in round 1: N = 2, 43.3 ms/step with NCCL against
                                   38.3 ms without any collective and 41.3 ms with copy-engine transfers).
"""
from __future__ import annotations

import datetime
import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment; initialises the process group if needed."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local)
        # a bounded timeout turns a mismatched collective into an error instead of a silent hang
        kw = {"device_id": torch.device("cuda", local)} if backend == "nccl" else {}
        dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                timeout=datetime.timedelta(seconds=int(os.environ.get("VQ_DIST_TIMEOUT_S", "180"))), **kw)
    return rank, world, local


def shard_batch(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """[start, end) of the scenes owned by `rank` when `global_batch` scenes are split as evenly as possible."""
    base, extra = divmod(global_batch, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


class LocalTransport:
    """world_size 1: a plain device buffer, no peers, barriers are no-ops."""

    def __init__(self, nbytes: int, device: torch.device) -> None:
        self.world, self.rank = 1, 0
        self.buf = torch.zeros(nbytes, dtype=torch.uint8, device=device)

    def local(self) -> torch.Tensor:
        return self.buf

    def peer(self, rank: int) -> torch.Tensor:
        assert rank == 0
        return self.buf

    def barrier(self) -> None:
        pass


class SymmetricMemoryTransport:
    """`nbytes` of NVLink-symmetric memory on every rank of `group`.  `peer(r)` is rank r's buffer mapped into this
    process (a device tensor: reading it is a load over NVLink / NVSwitch); `barrier()` is a device-side barrier on
    the CURRENT stream among all ranks (symmetric-memory signal pads, no host synchronisation)."""

    def __init__(self, nbytes: int, device: torch.device, group: Optional[dist.ProcessGroup] = None) -> None:
        import torch.distributed._symmetric_memory as symm_mem

        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.nbytes = nbytes
        self.buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        self._peers = {}
        self._channel = 0

    def local(self) -> torch.Tensor:
        return self.buf

    def peer(self, rank: int) -> torch.Tensor:
        if rank == self.rank:
            return self.buf
        if rank not in self._peers:
            self._peers[rank] = self.hdl.get_buffer(rank, (self.nbytes,), torch.uint8, 0)
        return self._peers[rank]

    def barrier(self) -> None:
        self.hdl.barrier(channel=self._channel)
        self._channel = (self._channel + 1) % 4
