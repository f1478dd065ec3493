"""Shared helpers for the parity tests (error metrics, small synthetic configs)."""
import math

import torch


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def max_abs(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.detach().double().cpu() - b.detach().double().cpu()).abs().max())


def mse(a: torch.Tensor, b: torch.Tensor) -> float:
    return float(((a.detach().double().cpu() - b.detach().double().cpu()) ** 2).mean())


def psnr_255(a: torch.Tensor, b: torch.Tensor) -> float:
    """PSNR on the 0..255 scale after the reference's postprocess (t2v_pipeline.rs:147-155)."""
    pa = (a.detach().double().cpu() * 0.5 + 0.5).clamp(0, 1) * 255
    pb = (b.detach().double().cpu() * 0.5 + 0.5).clamp(0, 1) * 255
    m = float(((pa - pb) ** 2).mean())
    return 99.0 if m == 0 else 10.0 * math.log10(255.0 ** 2 / m)
