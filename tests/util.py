"""Shared helpers for the parity tests (test plumbing only)."""
import torch


def to_internal(x, cp=None):
    """NCDHW fp32 -> (N,D,H,W,Cp) bf16 with zero channel padding (torch as test plumbing)."""
    n, c, d, h, w = x.shape
    cp = cp or (c + 31) // 32 * 32
    out = torch.zeros((n, d, h, w, cp), dtype=torch.bfloat16, device=x.device)
    out[..., :c] = x.permute(0, 2, 3, 4, 1).to(torch.bfloat16)
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
