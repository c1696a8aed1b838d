"""Patch-level agreement metrics with the reference's definitions (product-side copy used by the sampler
wrappers' return values and the drivers): masked_mae / masked_mse / psnr / ssim_simple
(Evaluation/DDIM_Multi-step.py:72-101), sam / ergas (Evaluation_Updated/Evaluation_Pure_Generation.py:229-254)."""
import ctypes as C
import math

import torch

METRIC_NAMES = ("mae", "mse", "psnr", "ssim_simple", "sam", "ergas", "valid")


METRIC_ROW = 24        # doubles per patch written by s1s2_patch_metrics (include/s1s2_b200.h)


def _patch_metrics_raw(pred, tgt, mask=None):
    """f64[N,24] rows of s1s2_patch_metrics: one fused CUDA pass per patch over (pred, tgt, mask)."""
    from . import _lib
    if pred.device.type != "cuda":
        raise _lib.S1S2Error("patch_metrics runs on a CUDA device only (no CPU fallback)")
    N, Cn, H, W = pred.shape
    pred = pred.to(torch.float32).contiguous()
    tgt = tgt.to(device=pred.device, dtype=torch.float32).contiguous()
    m = None
    if mask is not None:
        m = mask.to(pred.device).reshape(N, H * W)
        if m.dtype != torch.uint8:                 # tile_extract's u8 masks pass straight through (kernel tests != 0)
            m = (m > 0).to(torch.uint8)
        m = m.contiguous()
    out = torch.empty((N, METRIC_ROW), device=pred.device, dtype=torch.float64)
    idx = pred.device.index if pred.device.index is not None else torch.cuda.current_device()
    stream = torch.cuda.current_stream(pred.device).cuda_stream
    _lib.check(_lib.lib().s1s2_patch_metrics(idx, pred.data_ptr(), tgt.data_ptr(), m.data_ptr() if m is not None else None,
                                             N, Cn, H * W, out.data_ptr(), C.c_void_p(stream)))
    return out


def patch_metrics(pred, tgt, mask=None):
    """All six metrics of N patches in one fused CUDA pass each (s1s2_patch_metrics): f64[N,7] in METRIC_NAMES order.
    pred, tgt f32[N,C,H,W] on the same CUDA device; mask [N,H,W] / [N,1,H,W] (non-zero = valid) or None.
    One device->host copy of the result replaces ~10 `.item()` synchronisations per patch of the reference drivers."""
    return _patch_metrics_raw(pred, tgt, mask)[:, :7]


def channelwise_error_sums(pred, tgt, mask=None):
    """Limitation_Test.py:118-133 on the device: (abs_sum[C], sq_sum[C], valid_pixels) of a batch as float64 CUDA
    tensors -- add them up over the batches of a dataset and hand the totals to ``aggregate_final``."""
    raw = _patch_metrics_raw(pred, tgt, mask)
    Cn = pred.shape[1]
    return raw[:, 8:8 + Cn].sum(0), raw[:, 16:16 + Cn].sum(0), raw[:, 6].sum()


def aggregate_final(abs_sum_c, sq_sum_c, w_pix_sum, band_weights=None):
    """Limitation_Test.py:135-159: dataset-level MAE / MSE / PSNR from the accumulated sums: per-channel means over the
    valid pixels, combined with equal or normalised band weights.  Returns (mae, mse, psnr, mae_c, mse_c, psnr_c)."""
    a, q = abs_sum_c.detach().double().cpu(), sq_sum_c.detach().double().cpu()
    n = torch.as_tensor(w_pix_sum).detach().double().cpu().clamp_min(1e-8)
    mae_c, mse_c = a / n, q / n
    if band_weights is None:
        mae, mse = float(mae_c.mean()), float(mse_c.mean())
    else:
        bw = torch.tensor(band_weights, dtype=torch.float64)
        bw = bw / bw.sum().clamp_min(1e-8)
        mae, mse = float((mae_c * bw).sum()), float((mse_c * bw).sum())
    ps = 99.0 if mse <= 1e-12 else 10.0 * math.log10(1.0 / mse)
    ps_c = torch.where(mse_c <= 1e-12, torch.full_like(mse_c, 99.0), 10.0 * torch.log10(1.0 / mse_c))
    return mae, mse, ps, mae_c.numpy(), mse_c.numpy(), ps_c.numpy()


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
