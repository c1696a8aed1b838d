"""Image-agreement metrics used by the parity criterion (oracle; test infrastructure only).

Restates ``masked_mse`` / ``masked_mae`` / ``psnr`` / ``ssim_simple`` -- Evaluation/DDIM_Multi-step.py:72-101
(the reference's SSIM is global, not windowed), ``sam`` / ``ergas`` --
Evaluation_Updated/Evaluation_Pure_Generation.py:229-254, and the dataset-level pixel-weighted aggregation
``channelwise_error_sums`` / ``aggregate_final`` -- Evaluation/Limitation_Test.py:118-159.
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


def channelwise_error_sums(pred, tgt, mask=None):
    """Per-channel sums of |pred - tgt| and (pred - tgt)^2 over the valid pixels of a batch, and the number of valid
    pixels (Limitation_Test.py:118-133).  Returns (abs_sum[C], sq_sum[C], valid_pixels) as fp32 tensors."""
    w = _w(pred, mask)
    d = pred - tgt
    return (w * d.abs()).sum(dim=(0, 2, 3)), (w * d ** 2).sum(dim=(0, 2, 3)), w.sum()


def aggregate_final(abs_sum_c, sq_sum_c, w_pix_sum, band_weights=None):
    """Dataset-level MAE / MSE / PSNR from sums accumulated over all batches (Limitation_Test.py:135-159): per-channel
    means over the valid pixels, combined with equal or normalised band weights; PSNR = 99 when MSE <= 1e-12.
    Returns (mae, mse, psnr, mae_c, mse_c, psnr_c)."""
    n = w_pix_sum.clamp_min(1e-8)
    mae_c, mse_c = abs_sum_c / n, sq_sum_c / n
    if band_weights is None:
        mae, mse = mae_c.mean().item(), mse_c.mean().item()
    else:
        bw = torch.tensor(band_weights, dtype=mae_c.dtype)
        bw = bw / bw.sum().clamp_min(1e-8)
        mae, mse = (mae_c * bw).sum().item(), (mse_c * bw).sum().item()
    ps = 99.0 if mse <= 1e-12 else 10.0 * math.log10(1.0 / mse)
    ps_c = torch.where(mse_c <= 1e-12, torch.full_like(mse_c, 99.0), 10.0 * torch.log10(1.0 / mse_c))
    return mae, mse, ps, mae_c.numpy(), mse_c.numpy(), ps_c.numpy()
