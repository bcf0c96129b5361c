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
