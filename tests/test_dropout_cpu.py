"""CPU tests of the train-mode dropout plumbing: the Philox mask function (oracle restatement vs the published
known-answer vectors and vs the C++ header the kernels include) and the oracle's train-mode Perceiver against the
golden produced by the reference's own class under .train() (tests/golden/make_golden.py: perceiver_train_golden)."""
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import perceiver_ref as P

ROOT = Path(__file__).resolve().parent.parent
GOLD = Path(__file__).resolve().parent / "golden"


def test_philox_known_answers():
    # Random123 kat_vectors: philox4x32-10
    assert [int(x) for x in P._philox4x32_10(0, 0, 0, 0, 0, 0)] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    ones = 0xffffffff
    assert [int(x) for x in P._philox4x32_10(ones, ones, ones, ones, ones, ones)] == [
        0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert [int(x) for x in P._philox4x32_10(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0)] == [
        0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_cxx_header_matches_oracle_mask(tmp_path):
    """csrc/philox.cuh (host-compiled) and oracle.dropout_keep_mask produce the same keep bits."""
    src = tmp_path / "mask.cpp"
    src.write_text(r'''
#include <cstdio>
#include <cstdlib>
#include "philox.cuh"
int main(int argc, char** argv) {
  const long rows = atol(argv[1]); const int cols = atoi(argv[2]), ld = atoi(argv[3]);
  const float p = atof(argv[4]); const unsigned long long seed = strtoull(argv[5], 0, 10), off = strtoull(argv[6], 0, 10);
  const unsigned site = atoi(argv[7]);
  const vq::DropoutParams d = vq::make_dropout_params(p, seed, off, site);
  for (long r = 0; r < rows; ++r) {
    for (int c = 0; c < cols; ++c) {
      const uint64_t group = (uint64_t)r * (ld / 8) + c / 8;
      putchar('0' + ((vq::dropout_keep8(d, group) >> (c & 7)) & 1));
    }
    putchar('\n');
  }
}''')
    exe = tmp_path / "mask"
    subprocess.run(["g++", "-O1", "-I", str(ROOT / "vggt_qwen3_b200" / "csrc"), str(src), "-o", str(exe)], check=True)
    for rows, cols, ld, p, seed, off, site in [(37, 72, 128, 0.1, 20240611, 12, 0), (5, 256, 256, 0.5, 2**63 + 5, 2**33 + 4, 7),
                                               (64, 40, 40, 0.25, 0, 0, 23)]:
        out = subprocess.run([str(exe), *map(str, (rows, cols, ld, p, seed, off, site))], check=True,
                             capture_output=True, text=True).stdout.split()
        got = torch.tensor([[ch == "1" for ch in line] for line in out])
        assert torch.equal(got, P.dropout_keep_mask(rows, cols, ld, p, seed, off, site))


def test_mask_statistics_and_streams():
    m = P.dropout_keep_mask(512, 1024, 1024, 0.1, 7, 0, 1)
    assert abs(m.float().mean().item() - 0.9) < 2e-3
    assert abs(m.float().mean(0).min().item() - 0.9) < 0.08 and abs(m.float().mean(1).min().item() - 0.9) < 0.06
    assert not torch.equal(m, P.dropout_keep_mask(512, 1024, 1024, 0.1, 7, 0, 2))      # another site
    assert not torch.equal(m, P.dropout_keep_mask(512, 1024, 1024, 0.1, 7, 4, 1))      # another forward call
    assert not torch.equal(m, P.dropout_keep_mask(512, 1024, 1024, 0.1, 8, 0, 1))      # another seed
    assert torch.equal(m, P.dropout_keep_mask(512, 1024, 1024, 0.1, 7, 0, 1))
    assert P.dropout_keep_mask(4, 16, 16, 0.0, 1, 2, 3).all()
    assert P.dropout_threshold(0.1) == 6554


def _close(a, b, rtol, what):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    assert a.shape == b.shape, what
    assert (a - b).abs().max().item() <= rtol * (b.abs().max().item() + 1e-12), what


def test_oracle_train_mode_matches_reference_class_golden():
    """The reference's own PerceiverProjector under .train() (its dropout calls replaced by the deterministic masks in
    call order) against the oracle's train-mode restatement: pins the four dropout sites and their numbering."""
    z = np.load(GOLD / "perceiver_ref_train.npz")
    state = {k[len("param/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param/")}
    grads = {k[len("grad/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("grad/")}
    p, seed, offset = float(z["dropout"][0]), int(z["dropout"][1]), int(z["dropout"][2])
    heads = int(z["cfg"][2])
    out, g, gt = P.perceiver_forward_backward(state, torch.from_numpy(z["tokens"]), torch.from_numpy(z["grad_out"]),
                                              heads, dtype=torch.float64, dropout=(p, seed, offset))
    _close(out, z["out"], 1e-5, "train-mode forward")
    _close(gt, z["grad_tokens"], 2e-4, "grad tokens")
    for k in grads:
        _close(g[k], grads[k], 2e-4, f"grad {k}")
    # and it is NOT the eval-mode result
    ev = P.perceiver_forward(state, torch.from_numpy(z["tokens"]), heads)
    assert (ev - torch.from_numpy(z["out"])).abs().max() > 1e-2
