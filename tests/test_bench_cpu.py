"""CPU checks of the bench contract: the reference arm runs here (it is the CPU oracle port) and prints the JSON line
the driver parses; the B200 arm refuses to run without a GPU (no CPU fallback); the committed round-1 line carries
every key of the contract."""
import json
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"}


def test_reference_arm_prints_the_contract_line():
    import os

    # every reference step is one full-depth scene (tens of seconds of CPU work here): one step, bounded budget
    env = dict(os.environ, VQ_REF_BUDGET_S="30")
    p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "0"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["unit"] == "frames/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["config"]["workload"].startswith("c2")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert "full depth" in cb["sample"] and cb["steps_executed"] >= 1
    # ms_per_step is the time of an EXECUTED step (one scene), consistent with the run's wall clock
    assert d["ms_per_step"] * cb["steps_executed"] / 1e3 <= d["wall_seconds"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    """Under torchrun (N > 1) rank 0 alone runs the CPU arm; the other ranks print nothing and exit 0."""
    import os

    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box WITHOUT a GPU")
def test_b200_arm_has_no_cpu_fallback():
    p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1"], capture_output=True, text=True,
                       timeout=300, cwd=ROOT)
    assert p.returncode != 0 and "no CPU fallback" in (p.stderr + p.stdout)


def test_committed_round_line_carries_every_contract_key():
    d = json.loads((ROOT / "profiles" / "r01_bench_c2_final.json").read_text())
    assert BASE_KEYS | {"clocks", "gpu_launches", "roofline"} <= set(d)
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
    assert d["e2e"]["h2d_bytes_per_step"] == 8 * 3 * 518 * 518 * 4 and d["e2e"]["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "tensor"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3 and r["traffic"] > 0
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert d["gpu_launches"] == d["gpu_launches_per_step"] * d["steps"] > 0
    assert d["warmup"] >= 3 and d["dtype"] == "bf16" and "model" not in d["config"]
    c2 = json.loads((ROOT / "profiles" / "r01_bench_c2.json").read_text())
    assert {"value", "unit", "cores", "kind", "sample"} <= set(c2["cpu_baseline"])


def test_round2_line_carries_the_extended_contract():
    """The committed round-2 default line: training step with optimizer and dropout in the timed region, uint8 e2e input,
    the full result read back, c3 / c4 in the same line, the GPU-eager bar and the CPU baseline."""
    d = json.loads([l for l in (ROOT / "profiles" / "r02_bench_c2.json").read_text().splitlines() if l.startswith("{")][-1])
    assert BASE_KEYS | {"clocks", "gpu_launches", "roofline", "workloads", "gpu_eager_bf16", "cpu_baseline"} <= set(d)
    assert d["config"]["workload"].startswith("c2") and "AdamW" in d["config"]["optimizer"]
    assert "train mode" in d["config"]["dropout"] and "model" not in d["config"]
    assert d["e2e"]["h2d_bytes_per_step"] == 8 * 518 * 518 * 3            # uint8 views
    assert d["e2e"]["d2h_bytes_per_step"] == 8 * 128 * 2560 * 4            # the projected tokens, not a checksum
    assert d["e2e"]["value"] != d["value"]
    r = d["roofline"]
    assert r["bound"] == "tensor" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3 and r["traffic"] > 0
    assert set(d["workloads"]) == {"c3", "c4"}
    for w in d["workloads"].values():
        assert w["value"] > 0 and w["roofline"]["frac"] > 0 and w["e2e"]["h2d_bytes_per_step"] > 0
    assert d["gpu_eager_bf16"]["c2"]["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert "full depth" in d["cpu_baseline"]["sample"]
    assert d["warmup"] >= 3 and d["gpu_launches"] == d["gpu_launches_per_step"] * d["steps"] > 0
    ref = json.loads([l for l in (ROOT / "profiles" / "r02_bench_reference_arm.json").read_text().splitlines()
                      if l.startswith("{")][-1])
    assert ref["impl"] == "reference" and ref["config"]["workload"] == d["config"]["workload"]
    assert ref["ms_per_step"] * ref["cpu_baseline"]["steps_executed"] / 1e3 <= ref["wall_seconds"]
