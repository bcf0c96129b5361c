"""Multi-GPU test of the data-parallel step (needs >= 2 visible GPUs; skipped otherwise): optim.ShardedAdamW over
NVLink-symmetric memory with the real kernels.  Every rank trains the same projector on different tokens; after each
step all ranks must hold bit-identical compute copies, and the parameters must follow ONE torch.optim.AdamW fed the
fp32 mean over ranks of the fp32 oracle's gradients (run on the GPU in true fp32).  Also the raw reduce: random bf16
slices on every rank against an fp32 sum."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    try:
        import torch.distributed as dist

        from oracle import perceiver_ref as P
        from tests.parity import fp32_oracle_on_cuda
        from vggt_qwen3_b200.optim import ShardedAdamW
        from vggt_qwen3_b200.projector_perceiver import PerceiverConfig, PerceiverProjector

        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                          LOCAL_RANK=str(rank))
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        shape = P.PerceiverShape(512, 128, 4, 2, 1024, 0.0, 256, 320)
        state = P.init_perceiver_state(shape, seed=0, nonzero_bias=True)
        state["latents"] = state["latents"] * 25.0
        m = PerceiverProjector(PerceiverConfig(512, 128, 4, 2, 1024, 0.0), in_dim=256, out_dim=320)
        m.load_state_dict(state, strict=True)
        m = m.cuda().eval()
        opt = ShardedAdamW(m, lr=1e-3, weight_decay=0.01)
        ref_state = {k: v.clone().cuda().requires_grad_(True) for k, v in state.items()}
        ref_opt = torch.optim.AdamW(list(ref_state.values()), lr=1e-3, weight_decay=0.01)
        for step in range(3):
            ref_opt.zero_grad()
            for r in range(world):       # the reference sees every rank's batch
                g = torch.Generator().manual_seed(100 * step + r)
                tokens = torch.randn(2, 128, shape.in_dim, generator=g)
                grad_out = torch.randn(2, shape.num_latents, shape.out_dim, generator=g)
                with fp32_oracle_on_cuda():
                    P.perceiver_forward(ref_state, tokens.cuda(), shape.num_heads).backward(grad_out.cuda() / world)
                if r == rank:
                    mine = (tokens.cuda(), grad_out.cuda())
            ref_opt.step()
            m(mine[0]).backward(mine[1])
            opt.step()
        opt.wait()
        torch.cuda.synchronize()
        # bit-identical compute copies on every rank
        digest = torch.stack([opt.W["mat"].float().sum(), opt.W["vec"].sum(), opt.W["mat"].float().abs().max()])
        gathered = [torch.empty_like(digest) for _ in range(world)]
        dist.all_gather(gathered, digest)
        same = all(torch.equal(gathered[0], t) for t in gathered)
        full = [torch.empty_like(opt.W["mat"]) for _ in range(world)]
        dist.all_gather(full, opt.W["mat"])
        same = same and all(torch.equal(full[0], t) for t in full)
        opt.sync_parameters()
        worst = 0.0
        for name, p in m.named_parameters():
            r, p0, got = ref_state[name].detach(), state[name].cuda(), p.detach()
            if name.endswith("in_proj_bias"):   # key-bias gradient is identically zero: Adam steps on rounding noise
                E = r.numel() // 3
                keep = torch.cat([torch.arange(0, E), torch.arange(2 * E, 3 * E)]).cuda()
                r, p0, got = r[keep], p0[keep], got[keep]
            worst = max(worst, float((got - r).norm() / (r - p0).norm().clamp_min(1e-12)))
        # raw reduce numerics: bf16 slices summed in fp32
        n = opt.slice["mat"]
        gen = torch.Generator(device="cuda").manual_seed(7 + rank)
        opt.G["mat"].copy_((torch.randn(opt.region["mat"], device="cuda", generator=gen) * 3).bfloat16())
        opt.G["vec"].zero_()
        mine_all = [torch.empty_like(opt.G["mat"]) for _ in range(world)]
        dist.all_gather(mine_all, opt.G["mat"])
        expect_sum = torch.stack([t[rank * n:(rank + 1) * n].double() for t in mine_all]).sum(0)
        before = opt.master["mat"].clone()
        opt.exp_avg["mat"].zero_(); opt.exp_avg_sq["mat"].zero_()
        opt.param_groups[0].update(lr=1.0, weight_decay=0.0, betas=(0.0, 0.0), eps=1e-30)
        opt.step_count = 0
        opt.step()
        opt.wait()
        torch.cuda.synchronize()
        # betas = 0, eps ~ 0: the update is -lr * sign(g): a slice element moves by exactly -sign(sum) (0 where sum == 0)
        moved = opt.master["mat"] - before
        sign_ok = bool(torch.equal(torch.sign(-moved).double(), torch.sign(expect_sum)))
        q.put((rank, same, worst, sign_ok, opt.bytes_pulled))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as exc:  # pragma: no cover
        import traceback

        q.put((rank, f"{exc!r}\n{traceback.format_exc()}", 0, False, 0))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_adamw_over_symmetric_memory():
    import torch.multiprocessing as mp

    world = min(torch.cuda.device_count(), 4)
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, same, worst, sign_ok, pulled in results:
        assert not isinstance(same, str), f"rank {rank}: {same}"
        assert same, f"rank {rank}: compute copies differ between ranks"
        assert worst <= 0.12, f"rank {rank}: update error {worst:.3e} of the update size"
        assert sign_ok, f"rank {rank}: fp32 reduce of the bf16 slices is wrong"
        assert pulled > 0
