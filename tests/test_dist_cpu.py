"""World-size-2 `gloo` tests (CPU) of the data-parallel host logic: scene sharding and the ZeRO-1 shaped exchange of
optim.ShardedAdamW (reduce-scatter of the wire gradient -> AdamW on the owned slice -> all-gather of the weights),
including gradient accumulation and sharded state save / load.  The CUDA pieces (NVLink-symmetric memory, the fused
reduce + AdamW kernel) are replaced by test doubles here: a transport that snapshots every rank's buffer at each
barrier (what a peer would read after it) and a torch restatement of the kernel; the GPU tests run the real ones."""
import math
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vggt_qwen3_b200.dist import shard_batch


def test_shard_batch_covers_everything_once():
    for total in (64, 7, 8, 1):
        for world in (1, 2, 4, 8):
            spans = [shard_batch(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


class SnapshotTransport:
    """Test double of SymmetricMemoryTransport over gloo: `peer(r)` is rank r's buffer as of the last barrier."""

    def __init__(self, nbytes, device, group=None):
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.buf = torch.zeros(nbytes, dtype=torch.uint8)
        self.snap = [torch.zeros(nbytes, dtype=torch.uint8) for _ in range(self.world)]

    def local(self):
        return self.buf

    def peer(self, rank):
        return self.buf if rank == self.rank else self.snap[rank]

    def barrier(self):
        dist.all_gather(self.snap, self.buf)


class TorchKernels:
    """torch restatement of vq_adamw_reduce_step / vq_cast_f32_bf16 (torch.optim.AdamW semantics, fp32 reduce)."""

    def adamw_reduce_step(self, p, m, v, srcs, w, step, *, lr, betas, eps, weight_decay, grad_scale):
        g = sum(s.float() for s in srcs) * grad_scale
        m.mul_(betas[0]).add_(g, alpha=1 - betas[0])
        v.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
        bc1, bc2 = 1 - betas[0] ** step, 1 - betas[1] ** step
        p.mul_(1 - lr * weight_decay)
        p.addcdiv_(m, v.sqrt() / math.sqrt(bc2) + eps, value=-lr / bc1)
        if w is not None:
            w.copy_(p)

    def cast_f32_bf16(self, src, dst):
        dst.copy_(src)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, accumulate, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from vggt_qwen3_b200.optim import ShardedAdamW
        from vggt_qwen3_b200.projector_perceiver import PerceiverConfig, PerceiverProjector

        torch.manual_seed(0)
        proj = PerceiverProjector(PerceiverConfig(64, 64, 2, 2, 128, 0.0), in_dim=32, out_dim=48)
        extra = torch.nn.Linear(5, 7)
        names = [n for n, _ in proj.named_parameters()]
        ref_params = [p.detach().clone().requires_grad_(True) for p in list(proj.parameters()) + list(extra.parameters())]
        ref_opt = torch.optim.AdamW(ref_params, lr=1e-2, weight_decay=0.05)
        opt = ShardedAdamW(proj, lr=1e-2, weight_decay=0.05, accumulate_steps=accumulate, extra_params=extra.parameters(),
                           transport=SnapshotTransport, kernels=TorchKernels())
        assert opt.world == world and opt.rank == rank
        lay = proj.flat_layout()
        ordered = proj._ordered_params()
        order_of = {id(p): i for i, p in enumerate(ordered)}
        for step in range(3):
            total = [torch.zeros_like(p) for p in ref_params]
            for micro in range(accumulate):
                per_rank = []
                for r in range(world):      # every rank can compute every rank's gradient: same seeds
                    g = torch.Generator().manual_seed(1000 * step + 10 * micro + r)
                    per_rank.append([torch.randn(p.shape, generator=g) for p in ref_params])
                mine = per_rank[rank]
                # "backward": the projector's kernels would write these (matrices bf16 on the wire when accumulate == 1)
                if accumulate == 1:
                    views = proj._flat_grads[2]
                    for p, gi in zip(proj.parameters(), mine):
                        views[order_of[id(p)]].copy_(gi)
                else:
                    views = proj._grad_views(torch.device("cpu"))
                    for p, gi in zip(proj.parameters(), mine):
                        v = views[order_of[id(p)]]
                        v.copy_(gi) if micro == 0 else v.add_(gi)
                for p, gi in zip(extra.parameters(), mine[len(names):]):
                    p.grad = gi.clone() if micro == 0 else p.grad + gi
                for r in range(world):
                    for i, (t, gi) in enumerate(zip(total, per_rank[r])):
                        is_mat = i < len(names) and lay["kinds"][order_of[id(list(proj.parameters())[i])]] == "mat"
                        if accumulate == 1:
                            t.add_(gi.bfloat16().float() if is_mat else gi)   # wire rounding per rank
                        else:
                            t.add_(gi)
            opt.step()
            # reference: mean over ranks and micro-steps; with accumulation the wire rounding applies to each
            # rank's accumulated matrix gradient
            if accumulate > 1:
                total = [torch.zeros_like(p) for p in ref_params]
                for r in range(world):
                    acc = [torch.zeros_like(p) for p in ref_params]
                    for micro in range(accumulate):
                        g = torch.Generator().manual_seed(1000 * step + 10 * micro + r)
                        for a, p in zip(acc, ref_params):
                            a.add_(torch.randn(p.shape, generator=g))
                    for i, (t, a) in enumerate(zip(total, acc)):
                        is_mat = i < len(names) and lay["kinds"][order_of[id(list(proj.parameters())[i])]] == "mat"
                        t.add_(a.bfloat16().float() if is_mat else a)
            for p, t in zip(ref_params, total):
                p.grad = t / (world * accumulate)
            ref_opt.step()
            opt.zero_grad()
        # compute copies (what the forward kernels read) and the synced f32 parameters
        opt.sync_parameters()
        worst = 0.0
        for p, r in zip(list(proj.parameters()) + list(extra.parameters()), ref_params):
            worst = max(worst, float((p.detach() - r.detach()).abs().max() / r.detach().abs().max().clamp_min(1e-12)))
        views_w = proj._views_into(opt.W["mat"], opt.W["vec"])
        for p in proj.parameters():
            w = views_w[order_of[id(p)]]
            tol = 1e-2 if w.dtype == torch.bfloat16 else 1e-5
            assert torch.allclose(w.float(), p.detach(), rtol=tol, atol=tol * float(p.detach().abs().max())), "compute copy"
        # sharded state round trip
        st = opt.state_dict()
        assert st["step"] == 3 and st["master"]["mat"].numel() == opt.slice["mat"]
        before = {k: v.clone() for k, v in opt.master.items()}
        opt.master["mat"].zero_()
        opt.load_state_dict(st)
        assert all(torch.equal(opt.master[k], before[k]) for k in before)
        q.put((rank, worst, opt.bytes_pulled))
    except Exception as exc:  # pragma: no cover
        import traceback

        q.put((rank, f"{exc!r}\n{traceback.format_exc()}", 0))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


@pytest.mark.parametrize("accumulate", [1, 2])
def test_sharded_adamw_two_ranks_matches_torch_adamw(accumulate):
    """Two ranks with different gradients step exactly like ONE torch.optim.AdamW fed the mean gradient (matrices
    rounded to bf16 per rank: the wire format); both ranks end with the same weights."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, accumulate, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, worst, pulled in results:
        assert not isinstance(worst, str), f"rank {rank}: {worst}"
        assert worst < 2e-5, f"rank {rank}: parameters deviate from torch.optim.AdamW by {worst:.3e}"
        assert pulled > 0
