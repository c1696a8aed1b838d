"""Image-agreement metrics used by the parity criterion (oracle; test infrastructure only).

Restates ``masked_mse`` / ``masked_mae`` / ``psnr`` / ``ssim_simple`` -- Evaluation/DDIM_Multi-step.py:72-101
(the reference's SSIM is global, not windowed), ``sam`` / ``ergas`` --
Evaluation_Updated/Evaluation_Pure_Generation.py:229-254.
"""
import math

import torch


def _w(pred, mask):
    if mask is None:
        return torch.ones_like(pred[:, :1])
    m = mask.unsqueeze(1) if mask.ndim == 3 else mask
    return (m.float() > 0).float()


def masked_mae(pred, tgt, mask=None) -> float:
    w = _w(pred, mask)
    return ((w * (pred - tgt).abs()).sum() / (w.sum() * pred.size(1) + 1e-8)).item()


def masked_mse(pred, tgt, mask=None) -> float:
    w = _w(pred, mask)
    return ((w * (pred - tgt) ** 2).sum() / (w.sum() * pred.size(1) + 1e-8)).item()


def psnr(pred, tgt, mask=None) -> float:
    m = masked_mse(pred, tgt, mask)
    return 99.0 if m <= 1e-12 else 10.0 * math.log10(1.0 / m)


def ssim_simple(pred, tgt, C1=0.01 ** 2, C2=0.03 ** 2) -> float:
    mx, my = pred.mean().item(), tgt.mean().item()
    vx, vy = pred.var().item(), tgt.var().item()
    cxy = ((pred - pred.mean()) * (tgt - tgt.mean())).mean().item()
    return ((2 * mx * my + C1) * (2 * cxy + C2)) / ((mx ** 2 + my ** 2 + C1) * (vx + vy + C2) + 1e-8)


def sam(pred, tgt, mask=None) -> float:
    p, g = pred.squeeze(0), tgt.squeeze(0)
    m = (mask.squeeze(0) > 0) if mask is not None else torch.ones_like(p[0], dtype=torch.bool)
    p, g = p[:, m], g[:, m]
    cosv = (p * g).sum(0) / (p.norm(dim=0).clamp(min=1e-8) * g.norm(dim=0).clamp(min=1e-8))
    return torch.arccos(cosv.clamp(-1.0, 1.0)).mean().item()


def ergas(pred, tgt, mask=None, scale_ratio: float = 4.0) -> float:
    C = pred.size(1)
    acc = 0.0
    for c in range(C):
        rmse = math.sqrt(max(masked_mse(pred[:, c:c + 1], tgt[:, c:c + 1], mask), 0.0))
        acc += (rmse / (tgt[:, c:c + 1].mean().item() + 1e-8)) ** 2
    return 100.0 * (acc / C) ** 0.5 * scale_ratio
