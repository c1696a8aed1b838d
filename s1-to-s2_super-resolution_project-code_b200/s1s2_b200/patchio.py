"""Patch.py's on-disk products from in-memory rasters: `patch_XXXXXX.npz` files and `manifest.json`.

Reference: Patch.py:195-255 (window loop, filters, normalisation, np.savez_compressed with the keys inputs / target /
mask + folder,row,col,transform,crs,patch_size,stride,valid_ratio) and :289-305 (manifest).  The window filters and the
tile extraction run on the device (s1s2_tile_filter, s1s2_tile_extract); raster reading (rasterio) and the PNG previews
are outside the hot path and not reproduced.  The evaluation drivers read these files back (drivers.load_npz_as_tensors).
"""
import json
import os

import numpy as np
import torch

from . import patch

DEFAULTS = dict(valid_ratio_threshold=0.80, variance_threshold=1e-4, dark_thr=0.10, dark_max_ratio=0.60, texture_thr=5e-5)


def write_patches(inputs: torch.Tensor, target: torch.Tensor, out_dir: str, patch_size=256, stride=32, colloc=None,
                  folder="scene", transform=(), crs="", max_patches=0, start_count=0, **thresholds):
    """inputs f32[4,H,W] (HH dB, HV dB, incidence deg, elevation m), target f32[4,H,W] (B2,B3,B4,B8 in [0,1]) on a CUDA
    device; optional colloc u8[H,W].  Writes the kept windows in Patch.py's order and returns
    (manifest entries, counters dict with the four skip counts)."""
    th = dict(DEFAULTS, **thresholds)
    os.makedirs(out_dir, exist_ok=True)
    dev = inputs.device
    H, W = int(inputs.shape[1]), int(inputs.shape[2])
    origins = patch.tile_origins(H, W, patch_size, stride)
    stats = patch.tile_filter(inputs, target, origins, patch_size, colloc=colloc, **th).cpu().numpy()
    codes = stats[:, 7].astype(int)
    counters = dict(validratio_skipped=int((codes == 1).sum()), var_skipped=int((codes == 2).sum()),
                    dark_skipped=int((codes == 3).sum()), texture_skipped=int((codes == 4).sum()))
    keep = np.nonzero(codes == 0)[0]
    if max_patches:
        keep = keep[:max(0, max_patches - start_count)]
    vmask = torch.isfinite(target).all(0)                      # build_mask's target / collocation terms (Patch.py:41-49)
    if colloc is not None:
        vmask &= colloc.to(dev) > 0
    cond, mask, ratio = patch.tile_extract(inputs, origins[keep], patch_size, vmask=vmask.to(torch.uint8))
    cond, mask, ratio = cond.cpu().numpy(), mask.cpu().numpy(), ratio.cpu().numpy()
    tgt = target.cpu().numpy()
    entries = []
    for k, w in enumerate(keep):
        r, c = int(origins[w, 0]), int(origins[w, 1])
        M = mask[k].astype(bool)
        Y = tgt[:, r:r + patch_size, c:c + patch_size].copy()
        Y[:, ~M] = 0.0                                          # Patch.py:241-244
        Y = np.nan_to_num(Y, nan=0.0, posinf=0.0, neginf=0.0).astype(np.float32)
        pid = f"{start_count + k:06d}"
        path = os.path.join(out_dir, f"patch_{pid}.npz")
        np.savez_compressed(path, inputs=cond[k], target=Y, mask=mask[k].astype("uint8"), folder=folder, row=r, col=c,
                            transform=list(transform), crs=str(crs), patch_size=patch_size, stride=stride,
                            valid_ratio=float(ratio[k]))
        entries.append({"patch_id": pid, "folder": folder, "npz": os.path.relpath(path, out_dir),
                        "preview_dir": os.path.join("preview_patches", f"patch_{pid}"), "row": r, "col": c,
                        "valid_ratio": float(ratio[k])})
    return entries, counters


def write_manifest(out_dir, entries, counters, base_dir="", patch_size=256, stride=32, **thresholds):
    """manifest.json with Patch.py's keys (Patch.py:289-305)."""
    th = dict(DEFAULTS, **thresholds)
    doc = {"total_patches": len(entries), "dark_skipped": counters.get("dark_skipped", 0),
           "texture_skipped": counters.get("texture_skipped", 0), "validratio_skipped": counters.get("validratio_skipped", 0),
           "var_skipped": counters.get("var_skipped", 0), "base_dir": base_dir, "patch_size": patch_size, "stride": stride,
           "valid_ratio_threshold": th["valid_ratio_threshold"], "variance_threshold": th["variance_threshold"],
           "dark_thr": th["dark_thr"], "dark_max_ratio": th["dark_max_ratio"], "texture_thr": th["texture_thr"],
           "patches": entries[:2000]}
    with open(os.path.join(out_dir, "manifest.json"), "w") as f:
        json.dump(doc, f, indent=2, ensure_ascii=False)
    return doc
