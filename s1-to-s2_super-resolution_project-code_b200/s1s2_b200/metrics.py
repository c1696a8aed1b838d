"""Patch-level agreement metrics with the reference's definitions (product-side copy used by the sampler
wrappers' return values and the drivers): masked_mae / masked_mse / psnr / ssim_simple
(Evaluation/DDIM_Multi-step.py:72-101), sam / ergas (Evaluation_Updated/Evaluation_Pure_Generation.py:229-254)."""
import math

import torch


def _weights(pred, mask):
    if mask is None:
        return torch.ones_like(pred[:, :1])
    m = mask if mask.ndim == 4 else mask.unsqueeze(1)
    return (m.to(pred.device).float() > 0).float()


def masked_mae(pred, tgt, mask=None) -> float:
    w = _weights(pred, mask)
    return ((w * (pred - tgt).abs()).sum() / (w.sum() * pred.size(1) + 1e-8)).item()


def masked_mse(pred, tgt, mask=None) -> float:
    w = _weights(pred, mask)
    return ((w * (pred - tgt) ** 2).sum() / (w.sum() * pred.size(1) + 1e-8)).item()


def psnr(pred, tgt, mask=None) -> float:
    mse = masked_mse(pred, tgt, mask)
    return 99.0 if mse <= 1e-12 else 10.0 * math.log10(1.0 / mse)


def ssim_simple(pred, tgt, C1=0.01 ** 2, C2=0.03 ** 2) -> float:
    mu_x, mu_y = pred.mean(), tgt.mean()
    var_x, var_y = pred.var(), tgt.var()
    cov = ((pred - mu_x) * (tgt - mu_y)).mean()
    val = ((2 * mu_x * mu_y + C1) * (2 * cov + C2)) / ((mu_x ** 2 + mu_y ** 2 + C1) * (var_x + var_y + C2) + 1e-8)
    return float(val.item())


def sam(pred, tgt, mask=None) -> float:
    p, g = pred.squeeze(0), tgt.squeeze(0)
    sel = (mask.squeeze(0) > 0) if mask is not None else torch.ones_like(p[0], dtype=torch.bool)
    p, g = p[:, sel], g[:, sel]
    cos = (p * g).sum(0) / (p.norm(dim=0).clamp(min=1e-8) * g.norm(dim=0).clamp(min=1e-8))
    return torch.arccos(cos.clamp(-1.0, 1.0)).mean().item()


def ergas(pred, tgt, mask=None, scale_ratio: float = 4.0) -> float:
    n = pred.size(1)
    tot = 0.0
    for c in range(n):
        rmse = math.sqrt(max(masked_mse(pred[:, c:c + 1], tgt[:, c:c + 1], mask), 0.0))
        tot += (rmse / (tgt[:, c:c + 1].mean().item() + 1e-8)) ** 2
    return 100.0 * (tot / n) ** 0.5 * scale_ratio
