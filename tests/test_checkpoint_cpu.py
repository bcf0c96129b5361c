"""CPU tests of the checkpoint IO: the files written for the trainable parameters are exactly what
the reference's loader (src/inference/qa_inference.py:50-105) and HF `load_sharded_checkpoint` read, with the
reference's key names, and survive a round trip bit for bit."""
import json

import numpy as np
import torch

from vggt_qwen3_b200 import checkpoint_io as ck
from vggt_qwen3_b200.projector_perceiver import PerceiverConfig, PerceiverProjector


class _Wrapper(torch.nn.Module):
    """Stands for VGGTQwen3VLM: `projector.*`, `geom_head.*` and a frozen part that must not be written."""

    def __init__(self, seed: int):
        super().__init__()
        torch.manual_seed(seed)
        self.projector = PerceiverProjector(PerceiverConfig(64, 16, 2, 2, 128, 0.1), in_dim=32, out_dim=48)
        self.geom_head = torch.nn.Sequential(torch.nn.Linear(37, 48), torch.nn.SiLU(), torch.nn.Linear(48, 48))
        self.vision_model = torch.nn.Linear(8, 8)


def test_projector_keys_are_the_reference_keys():
    """`projector.*` names = the reference PerceiverProjector's own state-dict keys (golden file made by importing it)."""
    gold = np.load("tests/golden/perceiver_ref_small.npz")
    ref_keys = sorted(k[len("param/"):] for k in gold.files if k.startswith("param/"))
    cfg = json.loads(str(gold["cfg"])) if gold["cfg"].dtype.kind in "US" else None
    m = _Wrapper(0)
    ours = sorted(k for k in m.projector.state_dict().keys())
    layers_ref = {k for k in ref_keys if not k.startswith("layers.") or k.startswith("layers.0.") or k.startswith("layers.1.")}
    assert set(ours) == layers_ref or set(ours) == set(ref_keys), (cfg, sorted(set(ours) ^ set(ref_keys))[:6])


def test_sharded_round_trip_and_hf_loader(tmp_path):
    try:   # the reference imports it from modeling_utils (transformers 4.x); 5.x moved it to trainer_utils
        from transformers.modeling_utils import load_sharded_checkpoint
    except ImportError:
        from transformers.trainer_utils import load_sharded_checkpoint

    src, dst, hf = _Wrapper(1), _Wrapper(2), _Wrapper(3)
    root = ck.save_trainable_checkpoint(src, tmp_path, max_shard_bytes=60_000)   # forces several shards
    index = json.loads((root / ck.INDEX_NAME).read_text())
    files = sorted(set(index["weight_map"].values()))
    assert len(files) > 1 and all((root / f).exists() for f in files)
    assert all(k.startswith(("projector.", "geom_head.")) for k in index["weight_map"])
    assert set(index["weight_map"]) == {k for k in src.state_dict() if k.startswith(("projector.", "geom_head."))}
    assert index["metadata"]["total_size"] == sum(v.numel() * 4 for k, v in src.state_dict().items()
                                                  if k.startswith(("projector.", "geom_head.")))
    missing, unexpected = ck.load_checkpoint_dir(dst, tmp_path)
    assert unexpected == [] and all(k.startswith("vision_model.") for k in missing)
    for k, v in src.state_dict().items():
        same = torch.equal(v, dst.state_dict()[k])
        assert same == k.startswith(("projector.", "geom_head.")), k
    # the reference's own code path for this layout (qa_inference.py:76): HF load_sharded_checkpoint, strict=False
    load_sharded_checkpoint(hf, str(root), strict=False)
    for k, v in src.state_dict().items():
        if k.startswith(("projector.", "geom_head.")):
            assert torch.equal(v, hf.state_dict()[k]), k


def test_flat_and_legacy_layouts(tmp_path):
    src = _Wrapper(4)
    flat = tmp_path / "flat"
    flat.mkdir()
    torch.save(ck.trainable_state(src), flat / "pytorch_model.bin")
    a = _Wrapper(5)
    missing, unexpected = ck.load_checkpoint_dir(a, flat)
    assert unexpected == [] and torch.equal(a.projector.latents, src.projector.latents)
    legacy = tmp_path / "legacy"
    root = ck.save_trainable_checkpoint(src, legacy)
    root.rename(legacy / "pytorch_model_fp32.bin")
    b = _Wrapper(6)
    ck.load_checkpoint_dir(b, legacy)
    assert torch.equal(b.geom_head[2].weight, src.geom_head[2].weight)
    assert ck.load_checkpoint_dir(_Wrapper(7), tmp_path / "does_not_exist") is None
    empty = tmp_path / "empty"
    empty.mkdir()
    assert ck.load_checkpoint_dir(_Wrapper(7), empty) is None


def test_training_state_directory_round_trip(tmp_path):
    """`step_N/` directories (src/train/train_sft.py:177-190,250-257): merged f32 weights in the layout the reference's
    loader reads + per-rank optimizer shard + scheduler.bin + random_states_0.pkl; resuming restores all of them."""
    import math
    import random

    from tests.test_dist_cpu import TorchKernels
    from vggt_qwen3_b200.checkpoint_io import load_training_state, save_step
    from vggt_qwen3_b200.optim import ShardedAdamW
    from vggt_qwen3_b200.projector_perceiver import PerceiverConfig, PerceiverProjector

    class Wrap(torch.nn.Module):
        def __init__(self):
            super().__init__()
            torch.manual_seed(0)
            self.projector = PerceiverProjector(PerceiverConfig(64, 64, 2, 1, 128, 0.0), in_dim=32, out_dim=48)
            self.geom_head = torch.nn.Sequential(torch.nn.Linear(37, 48), torch.nn.SiLU(), torch.nn.Linear(48, 48))

    def train_two_steps(m, opt, sched):
        for step in range(2):
            g = torch.Generator().manual_seed(step)
            for v in m.projector._flat_grads[2]:
                v.copy_(torch.randn(v.shape, generator=g))
            for p in m.geom_head.parameters():
                p.grad = torch.randn(p.shape, generator=g)
            opt.step()
            sched.step()
            opt.zero_grad()

    def build():
        m = Wrap()
        opt = ShardedAdamW(m.projector, lr=1e-2, weight_decay=0.01, extra_params=m.geom_head.parameters(),
                           kernels=TorchKernels())
        sched = torch.optim.lr_scheduler.LambdaLR(_SchedShim(opt), lambda s: 1.0 / (1 + s))
        return m, opt, sched

    class _SchedShim(torch.optim.Optimizer):     # LR schedulers want a torch Optimizer; they only touch param_groups
        def __init__(self, opt):
            self.__dict__["_opt"] = opt
            super().__init__(opt.param_groups[0]["params"], {"lr": opt.param_groups[0]["lr"]})
            self.param_groups = opt.param_groups

        def step(self, closure=None):
            pass

    m, opt, sched = build()
    train_two_steps(m, opt, sched)
    random.seed(7)
    torch.manual_seed(11)
    root = save_step(m, tmp_path, 1500, optimizer=opt, scheduler=sched)
    assert root.name == "step_1500"
    for name in ("pytorch_model_fp32/pytorch_model.bin.index.json", "scheduler.bin", "random_states_0.pkl",
                 "optimizer_rank0.bin", "latest"):
        assert (root / name).exists(), name
    expect_rand, expect_torch = random.random(), torch.rand(3)
    want = {k: v.clone() for k, v in m.state_dict().items()}
    want_master = {k: v.clone() for k, v in opt.master.items()}
    lr_after = opt.param_groups[0]["lr"]

    m2, opt2, sched2 = build()        # fresh run, resumes
    found = load_training_state(m2, root, optimizer=opt2, scheduler=sched2)
    assert found == {"weights": True, "optimizer": True, "scheduler": True, "rng": True, "step": 1500}
    for k, v in m2.state_dict().items():
        assert torch.equal(v, want[k]), k
    assert all(torch.equal(opt2.master[k], want_master[k]) for k in want_master) and opt2.step_count == 2
    assert math.isclose(opt2.param_groups[0]["lr"], lr_after)
    assert random.random() == expect_rand and torch.equal(torch.rand(3), expect_torch)
    # and training continues identically
    train_two_steps(m, opt, sched)
    train_two_steps(m2, opt2, sched2)
    opt.sync_parameters(); opt2.sync_parameters()
    for (k, a), b in zip(m.state_dict().items(), m2.state_dict().values()):
        assert torch.equal(a, b), k
