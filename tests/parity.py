"""Parity metrics of the north star: bf16 build vs fp32 oracle on identical inputs/weights.

  per-token cosine >= 0.999
  per-token max relative error <= 2e-2, defined as max|a-b| / max|b| over the token's features
  token order and counts identical (checked exactly on shapes; see also the index-tagged tests)
"""
import torch

COS_MIN = 0.999
REL_MAX = 2e-2


def token_metrics(test: torch.Tensor, ref: torch.Tensor):
    assert test.shape == ref.shape, f"token count/order mismatch: {tuple(test.shape)} vs {tuple(ref.shape)}"
    a = test.detach().double().cpu().reshape(-1, test.shape[-1])
    b = ref.detach().double().cpu().reshape(-1, ref.shape[-1])
    cos = torch.nn.functional.cosine_similarity(a, b, dim=-1)
    rel = (a - b).abs().amax(dim=-1) / b.abs().amax(dim=-1).clamp_min(1e-30)
    return cos.min().item(), rel.max().item()


def assert_token_parity(test: torch.Tensor, ref: torch.Tensor, what: str, cos_min: float = COS_MIN, rel_max: float = REL_MAX):
    assert torch.isfinite(test.float()).all(), f"{what}: non-finite values"
    cos, rel = token_metrics(test, ref)
    assert cos >= cos_min, f"{what}: per-token cosine {cos:.6f} < {cos_min}"
    assert rel <= rel_max, f"{what}: per-token max relative error {rel:.4e} > {rel_max}"
    return cos, rel


class fp32_oracle_on_cuda:
    """Context for running the (device-agnostic, plain PyTorch) oracle restatements on `cuda` in TRUE fp32 for the
    BASELINE-size cases the CPU would take minutes for: TF32 off for matmul and cuDNN, SDPA forced to the math
    backend (explicit softmax(QK^T)V in fp32).  The oracle stays the checker; nothing of the product runs here."""

    def __enter__(self):
        from torch.nn.attention import SDPBackend, sdpa_kernel

        self._mm, self._cudnn = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        self._prec = torch.get_float32_matmul_precision()
        torch.set_float32_matmul_precision("highest")
        self._ctx = sdpa_kernel(SDPBackend.MATH)
        self._ctx.__enter__()
        return self

    def __exit__(self, *exc):
        self._ctx.__exit__(*exc)
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = self._mm, self._cudnn
        torch.set_float32_matmul_precision(self._prec)
        return False
