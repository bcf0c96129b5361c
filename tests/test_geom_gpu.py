"""GPU parity of the geometry-token head and the fused training splice through the C ABI:
forward against the oracle restatement, backward against fp32 autograd through the oracle."""
import pytest
import torch

from oracle import perceiver_ref as P

pytestmark = pytest.mark.gpu
IMAGE_ID = 7


def _geom_inputs(B, V, g, drop=()):
    return {k: torch.randn(B, V, n, generator=g) for k, n in (("R", 9), ("t", 3), ("K", 9), ("depth_hist", 16))
            if k not in drop}


@pytest.mark.parametrize("B,V,H", [(2, 3, 192), (8, 8, 2560), (11, 1, 512)])
def test_geom_head_forward_backward(B, V, H):
    from vggt_qwen3_b200.vggt_qwen3_vlm import _GeomHeadFunction

    g = torch.Generator().manual_seed(B + H)
    head = torch.nn.Sequential(torch.nn.Linear(37, H), torch.nn.SiLU(), torch.nn.Linear(H, H))
    geom = _geom_inputs(B, V, g)
    state = {"geom_head." + k: v.detach().clone().requires_grad_(True) for k, v in head.state_dict().items()}
    ref = P.encode_geom(state, geom, 1)[:, 0]
    dy = torch.randn(B, H, generator=g)
    ref.backward(dy)
    cuda_head = head.cuda()
    parts = [geom[k].cuda() for k in ("R", "t", "K", "depth_hist")]
    y = _GeomHeadFunction.apply(parts, [9, 3, 9, 16], B, V, cuda_head[0].weight, cuda_head[0].bias, cuda_head[2].weight,
                                cuda_head[2].bias)
    assert y.shape == (B, H) and y.dtype == torch.float32
    assert torch.allclose(y.detach().cpu(), ref.detach(), atol=2e-5, rtol=1e-5)
    y.backward(dy.cuda())
    for name, p in (("0.weight", cuda_head[0].weight), ("0.bias", cuda_head[0].bias), ("2.weight", cuda_head[2].weight),
                    ("2.bias", cuda_head[2].bias)):
        r = state["geom_head." + name].grad
        assert torch.allclose(p.grad.cpu(), r, atol=1e-4 * r.abs().max().item() + 1e-6, rtol=1e-4), name


def test_geom_missing_keys_are_zeros():
    """`geom_token.get(key, zeros)` of the reference (src/models/vggt_qwen3_vlm.py:168-173): a missing K contributes
    zeros of R's shape."""
    from vggt_qwen3_b200.vggt_qwen3_vlm import _GeomHeadFunction

    g = torch.Generator().manual_seed(0)
    B, V, H = 3, 2, 128
    head = torch.nn.Sequential(torch.nn.Linear(37, H), torch.nn.SiLU(), torch.nn.Linear(H, H)).cuda()
    geom = _geom_inputs(B, V, g, drop=("K",))
    full = dict(geom, K=torch.zeros(B, V, 9))
    state = {"geom_head." + k: v.detach().cpu() for k, v in head.state_dict().items()}
    ref = P.encode_geom(state, full, 1)[:, 0]
    parts = [geom["R"].cuda(), geom["t"].cuda(), None, geom["depth_hist"].cuda()]
    y = _GeomHeadFunction.apply(parts, [9, 3, 9, 16], B, V, head[0].weight, head[0].bias, head[2].weight, head[2].bias)
    assert torch.allclose(y.detach().cpu(), ref, atol=2e-5, rtol=1e-5)


def _splice_oracle(emb, ids, geom, n_geom, vis):
    feats = vis if geom is None else torch.cat([geom.unsqueeze(1).expand(-1, n_geom, -1).to(vis.dtype), vis], dim=1)
    return P.splice_overwrite(emb, ids, feats, IMAGE_ID)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("n_geom", [0, 8])
def test_splice_features_forward_backward(dtype, n_geom):
    from vggt_qwen3_b200.vggt_qwen3_vlm import _SpliceFunction

    g = torch.Generator().manual_seed(5)
    B, L, H, n_vis = 3, 200, 256, 24
    ids = torch.randint(8, 64, (B, L), generator=g)
    ids[0, 3] = IMAGE_ID
    ids[1, 150] = IMAGE_ID
    ids[2, 10] = IMAGE_ID
    ids[2, 20] = IMAGE_ID          # overlapping spans: the later position overwrites, its head rows win
    ids[2, 100] = IMAGE_ID
    emb = torch.randn(B, L, H, generator=g).to(dtype)
    vis = torch.randn(B, n_vis, H, generator=g).to(dtype)
    geom = torch.randn(B, H, generator=g) if n_geom else None
    # oracle forward + autograd backward in fp32/fp64 on the same (dtype-rounded) values
    e = emb.double().requires_grad_(True)
    v = vis.double().requires_grad_(True)
    gm = geom.to(dtype).double().requires_grad_(True) if n_geom else None
    want = _splice_oracle(e, ids, gm, n_geom, v)
    w = torch.randn(B, L, H, generator=g).to(dtype)
    want.backward(w.double())
    ec, vc = emb.cuda().requires_grad_(True), vis.cuda().requires_grad_(True)
    gc = geom.cuda().requires_grad_(True) if n_geom else None
    got = _SpliceFunction.apply(ec, ids.cuda(), gc, vc, IMAGE_ID, n_geom)
    assert torch.equal(got.detach().cpu().double(), want.detach().to(dtype).double())
    got.backward(w.cuda())
    assert torch.equal(ec.grad.cpu().double(), e.grad.to(dtype).double())
    tol = 1e-6 if dtype == torch.float32 else 1e-2   # f32 atomics: up to three addends per feature row, any order
    assert (vc.grad.cpu().double() - v.grad).abs().max() <= tol * v.grad.abs().max()
    if n_geom:
        assert (gc.grad.cpu().double() - gm.grad).abs().max() <= max(tol, 1e-6) * gm.grad.abs().max()
