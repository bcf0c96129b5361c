"""Checkpoint IO for the trainable side of the path, in the reference's on-disk formats.

What the reference reads (src/inference/qa_inference.py:50-105, `load_checkpoint_if_available`):
  <ckpt_dir>/pytorch_model_fp32/pytorch_model.bin.index.json + shard files   (HF sharded layout, the output of
      DeepSpeed's zero_to_fp32 merge of an Accelerate `save_state` directory -- src/train/train_sft.py:177-190)
  <ckpt_dir>/pytorch_model_fp32.bin/...                                      (legacy name of the same)
  <ckpt_dir>/*.bin | *.safetensors                                           (flat state dict)
all loaded with `strict=False`, keys as dumped from `VGGTQwen3VLM` (`projector.*`, `geom_head.*`, `text_model.*`, ...).

`save_trainable_checkpoint` writes that layout for the parameters this path trains (projector + geom_head, f32);
`load_checkpoint_dir` mirrors the reference's resolution order.  Loading goes through `load_state_dict`, i.e. in-place
`copy_` on the f32 masters, which bumps their version counters: the projector's packed bf16 compute copies are
regenerated lazily on the next forward (projector_perceiver._refresh_compute_copies).
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Dict, Iterable, List, Optional, Tuple

import torch

INDEX_NAME = "pytorch_model.bin.index.json"
TRAINABLE_PREFIXES = ("projector.", "geom_head.")


def trainable_state(model: torch.nn.Module, prefixes: Iterable[str] = TRAINABLE_PREFIXES) -> Dict[str, torch.Tensor]:
    """f32 CPU copies of the state-dict entries under `prefixes`, in state-dict order."""
    prefixes = tuple(prefixes)
    return {k: v.detach().to("cpu", torch.float32).contiguous() for k, v in model.state_dict().items()
            if k.startswith(prefixes)}


def _shard(state: Dict[str, torch.Tensor], max_shard_bytes: int) -> List[Dict[str, torch.Tensor]]:
    shards, cur, size = [], {}, 0
    for k, v in state.items():
        n = v.numel() * v.element_size()
        if cur and size + n > max_shard_bytes:
            shards.append(cur)
            cur, size = {}, 0
        cur[k] = v
        size += n
    if cur:
        shards.append(cur)
    return shards


def save_trainable_checkpoint(model: torch.nn.Module, ckpt_dir, max_shard_bytes: int = 2 << 30,
                              prefixes: Iterable[str] = TRAINABLE_PREFIXES) -> Path:
    """Write <ckpt_dir>/pytorch_model_fp32/{pytorch_model-0000i-of-0000N.bin, pytorch_model.bin.index.json}."""
    root = Path(ckpt_dir) / "pytorch_model_fp32"
    root.mkdir(parents=True, exist_ok=True)
    state = trainable_state(model, prefixes)
    if not state:
        raise RuntimeError(f"no parameters under {tuple(prefixes)} to save")
    shards = _shard(state, max_shard_bytes)
    weight_map, total = {}, 0
    for i, shard in enumerate(shards, 1):
        name = f"pytorch_model-{i:05d}-of-{len(shards):05d}.bin"
        torch.save(shard, root / name)
        for k, v in shard.items():
            weight_map[k] = name
            total += v.numel() * v.element_size()
    with (root / INDEX_NAME).open("w", encoding="utf-8") as f:
        json.dump({"metadata": {"total_size": total}, "weight_map": weight_map}, f, indent=2, sort_keys=True)
    return root


def _load_file(path: Path) -> Dict[str, torch.Tensor]:
    if path.suffix == ".safetensors":
        from safetensors.torch import load_file

        return load_file(str(path))
    return torch.load(path, map_location="cpu", weights_only=True)


def load_checkpoint_dir(model: torch.nn.Module, ckpt_dir) -> Optional[Tuple[List[str], List[str]]]:
    """Mirror of the reference's `load_checkpoint_if_available`: sharded `pytorch_model_fp32[.bin]/` with an index
    first, then flat `*.bin` / `*.safetensors`; always `strict=False`.  Returns (missing, unexpected) keys, or None
    when the directory holds no weights (the reference then keeps the base weights)."""
    if not ckpt_dir:
        return None
    path = Path(ckpt_dir)
    if not path.exists():
        return None
    for cand in (path / "pytorch_model_fp32", path / "pytorch_model_fp32.bin"):
        index = cand / INDEX_NAME
        if cand.is_dir() and index.exists():
            with index.open("r", encoding="utf-8") as f:
                files = sorted({v for v in json.load(f).get("weight_map", {}).values()})
            missing, unexpected = set(model.state_dict().keys()), []
            for name in files:
                res = model.load_state_dict(_load_file(cand / name), strict=False)
                missing &= set(res.missing_keys)
                unexpected += res.unexpected_keys
            return sorted(missing), unexpected
    merged, legacy = path / "pytorch_model_fp32", path / "pytorch_model_fp32.bin"
    if merged.exists():
        files = [merged] if merged.is_file() else sorted(merged.glob("*.bin"))
    elif legacy.is_dir():
        files = sorted(legacy.glob("*.bin"))
    else:
        files = sorted(path.glob("*.bin")) + sorted(path.glob("*.safetensors"))
    if not files:
        return None
    res = model.load_state_dict(_load_file(files[0]), strict=False)
    return list(res.missing_keys), list(res.unexpected_keys)


# ------------------------------------------------------------------------------------------------
# Training-state directories (resume): what the reference gets from `accelerator.save_state(dir)` every
# `save_every_steps` steps and at the end (src/train/train_sft.py:177-190,250-257) -- `step_N/` directories holding
# the model, the (per-rank, sharded) optimizer state, `scheduler.bin` and `random_states_{rank}.pkl`.  Same file names
# for the pieces Accelerate names; the model goes straight to the merged `pytorch_model_fp32/` layout the reference's
# inference loader reads (no `zero_to_fp32.py` pass needed: the masters are gathered at save time).
# ------------------------------------------------------------------------------------------------
SCHEDULER_NAME = "scheduler.bin"
RNG_NAME = "random_states_{rank}.pkl"
OPTIMIZER_NAME = "optimizer_rank{rank}.bin"


def _rng_states() -> dict:
    import random

    import numpy as np

    states = {"random_state": random.getstate(), "numpy_random_seed": np.random.get_state(),
              "torch_manual_seed": torch.get_rng_state()}           # Accelerate's keys (accelerate/checkpointing.py)
    if torch.cuda.is_available():
        states["torch_cuda_manual_seed"] = torch.cuda.get_rng_state_all()
    return states


def _restore_rng(states: dict) -> None:
    import random

    import numpy as np

    random.setstate(states["random_state"])
    np.random.set_state(states["numpy_random_seed"])
    torch.set_rng_state(states["torch_manual_seed"])
    if torch.cuda.is_available() and "torch_cuda_manual_seed" in states:
        torch.cuda.set_rng_state_all(states["torch_cuda_manual_seed"])


def save_training_state(model: torch.nn.Module, ckpt_dir, optimizer=None, scheduler=None, step: Optional[int] = None,
                        rank: int = 0, world: int = 1) -> Path:
    """EVERY rank calls this (like the reference's `save_checkpoint`, train_sft.py:177-190): rank 0 writes the merged
    f32 weights (`pytorch_model_fp32/`), `scheduler.bin` and `latest`; every rank writes its optimizer shard and RNG
    state.  With `optim.ShardedAdamW` the f32 masters are gathered into the module first (`sync_parameters`, a
    collective: hence every rank)."""
    root = Path(ckpt_dir)
    if optimizer is not None and hasattr(optimizer, "sync_parameters"):
        optimizer.sync_parameters()
    if rank == 0:
        root.mkdir(parents=True, exist_ok=True)
        save_trainable_checkpoint(model, root)
        if scheduler is not None:
            torch.save(scheduler.state_dict(), root / SCHEDULER_NAME)
        if step is not None:
            (root / "latest").write_text(f"step_{step}\n")
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
    root.mkdir(parents=True, exist_ok=True)
    if optimizer is not None:
        torch.save(optimizer.state_dict(), root / OPTIMIZER_NAME.format(rank=rank))
    torch.save(_rng_states(), root / RNG_NAME.format(rank=rank))
    return root


def load_training_state(model: torch.nn.Module, ckpt_dir, optimizer=None, scheduler=None, rank: int = 0) -> dict:
    """Resume from a directory written by `save_training_state`: weights (through the reference's resolution order),
    this rank's optimizer shard, the scheduler and the RNG streams (so that the dropout masks of the resumed run are
    the ones the uninterrupted run would have drawn).  Returns what was found."""
    root = Path(ckpt_dir)
    found = {"weights": load_checkpoint_dir(model, root) is not None, "optimizer": False, "scheduler": False,
             "rng": False, "step": None}
    opt_file = root / OPTIMIZER_NAME.format(rank=rank)
    if optimizer is not None and opt_file.exists():
        if hasattr(optimizer, "reload_masters"):
            optimizer.reload_masters()       # the weights just loaded
        optimizer.load_state_dict(torch.load(opt_file, map_location="cpu", weights_only=False))
        found["optimizer"] = True
    if scheduler is not None and (root / SCHEDULER_NAME).exists():
        scheduler.load_state_dict(torch.load(root / SCHEDULER_NAME, map_location="cpu", weights_only=False))
        found["scheduler"] = True
    rng_file = root / RNG_NAME.format(rank=rank)
    if rng_file.exists():
        _restore_rng(torch.load(rng_file, map_location="cpu", weights_only=False))
        found["rng"] = True
    if (root / "latest").exists():
        tag = (root / "latest").read_text().strip()
        found["step"] = int(tag.split("_")[-1]) if tag.split("_")[-1].isdigit() else None
    return found


def save_step(model: torch.nn.Module, output_dir, step: int, **kw) -> Path:
    """`<output_dir>/step_<N>/` like the reference's periodic save (train_sft.py:250-254)."""
    return save_training_state(model, Path(output_dir) / f"step_{step}", step=step, **kw)
