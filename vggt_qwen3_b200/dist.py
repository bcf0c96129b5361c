"""Data-parallel plumbing for stage-1 training of the projector: one process per GPU, the visual-encode path is
replicated (scenes are independent) and the ONLY collective is the all-reduce of the Perceiver
gradient — the reference's equivalent is DeepSpeed ZeRO-3 / DDP (src/train/train_sft.py:119-133,165-170,217).

`BucketedGradAllReduce` overlaps that all-reduce with the backward: `vq_perceiver_backward` reports each bucket
(out_proj, layer L-1 .. 0, in_proj+latents) as soon as the kernels producing it are enqueued; the bucket is a
contiguous slice of the flat f32 gradient buffer and is reduced with NCCL on a side stream while the remaining
layers' backward GEMMs run.  Works with any torch.distributed backend (gloo on CPU in the tests).
"""
from __future__ import annotations

import datetime
import os
from typing import Callable, List, Optional, Sequence, Tuple

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


class BucketedGradAllReduce:
    """All-reduce (mean) of contiguous slices of a flat gradient buffer, bucket by bucket, on a side stream."""

    def __init__(self, group: Optional[dist.ProcessGroup] = None, grad_dtype: Optional[torch.dtype] = None) -> None:
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.grad_dtype = grad_dtype  # e.g. torch.bfloat16 to halve the bytes on the wire
        self._flat: Optional[torch.Tensor] = None
        self._ranges: List[Tuple[int, int]] = []
        self._pending: List[Tuple[object, torch.Tensor, Optional[torch.Tensor]]] = []
        self._comm_stream: Optional[torch.cuda.Stream] = None
        self.bytes_reduced = 0

    def bind(self, flat: torch.Tensor, ranges: Sequence[Tuple[int, int]]) -> None:
        """`ranges[i]` = [start, end) element range of bucket i inside `flat`."""
        self._flat = flat
        self._ranges = list(ranges)
        if flat.is_cuda and self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=flat.device)

    def bucket_ready(self, index: int) -> None:
        """Called right after the work producing bucket `index` was enqueued on the current stream."""
        if self.world == 1 or self._flat is None:
            return
        start, end = self._ranges[index]
        chunk = self._flat[start:end]
        if chunk.is_cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            with torch.cuda.stream(self._comm_stream):
                self._comm_stream.wait_event(ev)
                wire = chunk if self.grad_dtype in (None, chunk.dtype) else chunk.to(self.grad_dtype)
                work = dist.all_reduce(wire, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        else:
            wire = chunk if self.grad_dtype in (None, chunk.dtype) else chunk.to(self.grad_dtype)
            work = dist.all_reduce(wire, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self.bytes_reduced += wire.numel() * wire.element_size()
        self._pending.append((work, chunk, wire if wire is not chunk else None))

    def finish(self) -> None:
        """Wait for every bucket, scale by 1/world (and cast back if a wire dtype was used)."""
        if not self._pending:
            return
        cuda = self._flat is not None and self._flat.is_cuda
        ctx = torch.cuda.stream(self._comm_stream) if cuda else _NullCtx()
        with ctx:
            for work, chunk, wire in self._pending:
                work.wait()
                if wire is not None:
                    chunk.copy_(wire)
                chunk.mul_(1.0 / self.world)
        if cuda:
            torch.cuda.current_stream().wait_stream(self._comm_stream)
        self._pending.clear()


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def attach_projector_allreduce(projector, reducer: BucketedGradAllReduce) -> Callable[[], None]:
    """Wire `reducer` to a PerceiverProjector: buckets follow `projector.grad_buckets()` (completion order).
    Returns the `finish` callable to invoke after `backward()`."""

    def hook(layer_index: int) -> None:
        flat = projector._grad_flat
        if flat is None:
            return
        params = projector._ordered_params()
        offs, off = {}, 0
        for p in params:
            offs[id(p)] = off
            off += (p.numel() + 63) // 64 * 64
        ranges = []
        for bucket in projector.grad_buckets():
            starts = [offs[id(p)] for p in bucket]
            ends = [offs[id(p)] + p.numel() for p in bucket]
            ranges.append((min(starts), max(ends)))
        reducer.bind(flat, ranges)
        n_layers = len(projector.layers)
        index = 0 if layer_index == n_layers else (n_layers - layer_index if layer_index >= 0 else n_layers + 1)
        reducer.bucket_ready(index)

    projector.grad_ready_hook = hook
    return reducer.finish
