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

    def __init__(self, group: Optional[dist.ProcessGroup] = None, grad_dtype: Optional[torch.dtype] = None,
                 keep_wire: bool = False) -> None:
        self.group = group
        # keep_wire: leave the reduced SUM in the wire-dtype chunks (`wire_chunks`) instead of copying it back into the
        # f32 buffer and scaling it: the consumer (optim.FusedAdamW / vq_adamw_step) reads bf16 gradients and folds
        # 1/world into `grad_scale`, which saves three full passes over the 1.23 B-element gradient per step
        self.keep_wire = keep_wire
        self.wire_chunks: List[torch.Tensor] = []
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
            self.wire_chunks = []
            for work, chunk, wire in self._pending:
                work.wait()
                if self.keep_wire:
                    self.wire_chunks.append(wire if wire is not None else chunk)
                    continue
                if wire is not None:
                    chunk.copy_(wire)
                chunk.mul_(1.0 / self.world)
        if cuda:
            torch.cuda.current_stream().wait_stream(self._comm_stream)
        self._pending.clear()


class PeerCopyGradAllReduce:
    """All-reduce (sum) of the flat Perceiver gradient over NVLink peer memory with the COPY ENGINES, no SM use.

    Why: the aggregator's GEMMs are persistent kernels (one CTA pair per SM pair, static tile schedule); an NCCL
    kernel that holds even a few SMs while they run pushes part of every GEMM grid into a second wave, and the
    all-reduce "hidden" behind the next step's frozen forward then costs as much as it would exposed (measured at
    N = 2: 47 ms/step against 40 ms without the collective, for any NCCL_MAX_CTAS).  DMA copies do not take SMs.

    Layout: one symmetric bf16 buffer per rank (torch.distributed._symmetric_memory: cuMem allocations mapped into
    every peer), split in `world` shards.  finish() =
        barrier | rank r pulls shard r from every peer (cudaMemcpyAsync over NVLink) and adds it to its own shard
        | barrier | rank r pulls the reduced shard p from every peer p | barrier
    i.e. reduce-scatter + all-gather, (N-1)/N of the buffer over the wire in each phase like a ring, but on the DMA
    engines.  The result (SUM over ranks, bf16) stays in `self.buf[:numel]`; consumers fold 1/world into their
    gradient scale (vq_adamw_step's grad_scale)."""

    def __init__(self, numel: int, device: torch.device, group: Optional[dist.ProcessGroup] = None) -> None:
        import torch.distributed._symmetric_memory as symm_mem

        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.numel = numel
        self.shard = ((numel + self.world - 1) // self.world + 127) // 128 * 128
        self.buf = symm_mem.empty(self.shard * self.world, dtype=torch.bfloat16, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        self.stage = torch.empty((max(self.world - 1, 1), self.shard), dtype=torch.bfloat16, device=device)
        self.peers = [(self.rank + i) % self.world for i in range(1, self.world)]   # staggered: no hot spot
        self.stream = torch.cuda.Stream(device=device)
        self.bytes_reduced = 0
        self._armed = False

    def load(self, flat_f32: torch.Tensor, offsets: Sequence[Tuple[int, int, int]]) -> None:
        """Cast the f32 gradient into the wire buffer on the side stream, after the work enqueued so far on the current
        stream.  offsets = (src_start, dst_start, length) triples (the flat f32 buffer pads each tensor to 64)."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ev)
            for src, dst, n in offsets:
                self.buf[dst:dst + n].copy_(flat_f32[src:src + n])
        flat_f32.record_stream(self.stream)
        self._armed = True

    def finish(self) -> None:
        if not self._armed:
            return
        r, sh = self.rank, self.shard
        with torch.cuda.stream(self.stream):
            self.hdl.barrier(channel=0)
            for i, p in enumerate(self.peers):
                self.stage[i].copy_(self.hdl.get_buffer(p, (sh,), torch.bfloat16, r * sh))
            mine = self.buf[r * sh:(r + 1) * sh]
            for i in range(len(self.peers)):
                mine.add_(self.stage[i])
            self.hdl.barrier(channel=1)
            for p in self.peers:
                self.buf[p * sh:(p + 1) * sh].copy_(self.hdl.get_buffer(p, (sh,), torch.bfloat16, p * sh))
            self.hdl.barrier(channel=2)
        torch.cuda.current_stream().wait_stream(self.stream)
        self.bytes_reduced += 2 * (self.world - 1) * sh * 2
        self._armed = False


def attach_projector_peer_allreduce(projector, reducer: PeerCopyGradAllReduce) -> Callable[[], None]:
    """Wire a PeerCopyGradAllReduce to a PerceiverProjector: the whole flat gradient is handed over when the last
    bucket (in_proj + latents) of the backward has been enqueued.  Returns `finish`."""

    def hook(layer_index: int) -> None:
        if layer_index != -1 or projector._grad_flat is None:
            return
        offs, src, dst = [], 0, 0
        for p in projector._ordered_params():
            offs.append((src, dst, p.numel()))
            src += (p.numel() + 63) // 64 * 64
            dst += p.numel()
        reducer.load(projector._grad_flat, offs)

    projector.grad_ready_hook = hook
    return reducer.finish


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
