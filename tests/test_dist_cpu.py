"""CPU tests of the N>1 plumbing (one process per rank, gloo backend, world_size 2):
scene sharding and the bucketed Perceiver-gradient all-reduce that bench.py / stage-1 training use over NCCL."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vggt_qwen3_b200.dist import BucketedGradAllReduce, shard_batch


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_batch_covers_every_scene_once():
    for global_batch in (1, 7, 8, 64, 65):
        for world in (1, 2, 4, 8):
            spans = [shard_batch(global_batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == global_batch
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a1 >= a0 and b1 >= b0
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert [shard_batch(64, r, 8) for r in range(8)] == [(8 * r, 8 * r + 8) for r in range(8)]


def _worker(rank: int, world: int, port: int, wire_bf16: bool, out, keep_wire: bool = False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)
        sizes = [640, 4096, 1024, 64]                     # out_proj, two layers, in_proj+latents (completion order)
        flat = torch.randn(sum(sizes))
        mine = flat.clone()
        ranges, off = [], 0
        for n in sizes:
            ranges.append((off, off + n))
            off += n
        red = BucketedGradAllReduce(grad_dtype=torch.bfloat16 if wire_bf16 else None, keep_wire=keep_wire)
        red.bind(flat, ranges)
        for i in range(len(sizes)):                       # buckets become ready one by one during backward
            red.bucket_ready(i)
        red.finish()
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        expect = torch.stack(gathered).mean(0)
        tol = 2e-2 if wire_bf16 else 1e-6
        if keep_wire:
            # the SUM stays in the wire-dtype chunks (the consumer folds 1/world into its gradient scale); the f32
            # buffer is left untouched
            summed = torch.cat([c.float() for c in red.wire_chunks])
            ok = torch.allclose(summed / world, expect, atol=tol, rtol=tol) and torch.equal(flat, mine)
        else:
            ok = torch.allclose(flat, expect, atol=tol, rtol=tol)
        nbytes = red.bytes_reduced
        out.put((rank, bool(ok), int(nbytes)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("wire_bf16", [False, True])
def test_bucketed_allreduce_world2_gloo(wire_bf16):
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, wire_bf16, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = [out.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    total = 640 + 4096 + 1024 + 64
    for rank, ok, nbytes in results:
        assert ok, f"rank {rank}: averaged gradient mismatch"
        assert nbytes == total * (2 if wire_bf16 else 4)


def test_bucketed_allreduce_keep_wire_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, True, out, True)) for r in range(world)]
    for p in procs:
        p.start()
    results = [out.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in results)


def test_single_process_is_a_no_op():
    red = BucketedGradAllReduce()
    flat = torch.arange(10.0)
    red.bind(flat, [(0, 4), (4, 10)])
    red.bucket_ready(0)
    red.bucket_ready(1)
    red.finish()
    assert torch.equal(flat, torch.arange(10.0)) and red.bytes_reduced == 0
