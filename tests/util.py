"""Shared helpers for the parity tests (test plumbing only)."""
import torch


def to_internal(x, cp=None, dtype=torch.bfloat16):
    """NCDHW fp32 -> (N,D,H,W,Cp) with zero channel padding (torch as test plumbing): bf16 for activations and
    gradients, ``dtype=torch.float16`` for the raw conv output y of a conv -> norm block."""
    n, c, d, h, w = x.shape
    cp = cp or (c + 31) // 32 * 32
    out = torch.zeros((n, d, h, w, cp), dtype=dtype, device=x.device)
    out[..., :c] = x.permute(0, 2, 3, 4, 1).to(dtype)
    return out.contiguous()


def from_internal(x, c):
    """(N,D,H,W,Cp) bf16 -> NCDHW fp32."""
    return x[..., :c].permute(0, 4, 1, 2, 3).float().contiguous()


def bf16_round(x):
    return x.to(torch.bfloat16).float()


def rel_to_max(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def rel_l2(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def strict_fp32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def no_dropout(*modules):
    """Set p = 0 on every nn.Dropout inside the modules (oracle and drop-in alike: both read the layer's own p)."""
    for mod in modules:
        for m in mod.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
