"""Patch.py's tiling on the device: window enumeration (host, integer-exact), tile extraction with per-patch
normalisation and the overlap-blend stitch (CUDA gather kernels behind s1s2_tile_extract / s1s2_stitch).

Reference: patch_iter Patch.py:80-84; slicing :201-203; build_mask :41-49; zscore_inplace :51-62 applied at
:228-229; incidence / elevation scaling :231-232; invalid / non-finite -> 0 :236-239.  The stitch does not exist
in the reference: uniform-weight average of all patches covering a pixel, patches
visited in ascending index, uncovered pixels 0 with cover mask 0.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


def patch_iter(H: int, W: int, ps: int, stride: int):
    """Window origins (row, col) in the reference's order: rows outer, columns inner; remainder not covered."""
    for r in range(0, H - ps + 1, stride):
        for c in range(0, W - ps + 1, stride):
            yield r, c


def tile_origins(H: int, W: int, ps: int, stride: int) -> np.ndarray:
    return np.array(list(patch_iter(H, W, ps, stride)), dtype=np.int32).reshape(-1, 2)


def shard_range(n: int, rank: int, world: int):
    """Contiguous near-equal split of n units: rank r owns [lo, hi)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _dev_index(t: torch.Tensor) -> int:
    if t.device.type != "cuda":
        raise _lib.S1S2Error("patch kernels run on a CUDA device only (no CPU fallback)")
    return t.device.index if t.device.index is not None else torch.cuda.current_device()


def tile_extract(scene: torch.Tensor, origins, ps: int, vmask: torch.Tensor = None):
    """scene f32[4,SH,SW] (cuda) -> (cond f32[N,4,ps,ps], mask u8[N,ps,ps], valid_ratio f32[N])."""
    dev = scene.device
    idx = _dev_index(scene)
    scene = scene.to(torch.float32).contiguous()
    if scene.ndim != 3 or scene.shape[0] != 4:
        raise ValueError(f"scene must be f32[4,H,W], got {tuple(scene.shape)}")
    org = torch.as_tensor(np.ascontiguousarray(origins, dtype=np.int32)).reshape(-1, 2)
    N = org.shape[0]
    SH, SW = scene.shape[1:]
    if N and (int(org.min()) < 0 or int(org[:, 0].max()) + ps > SH or int(org[:, 1].max()) + ps > SW):
        raise ValueError("window outside the scene")
    org_d = org.to(dev)
    cond = torch.empty((N, 4, ps, ps), device=dev, dtype=torch.float32)
    mask = torch.empty((N, ps, ps), device=dev, dtype=torch.uint8)
    ratio = torch.empty((N,), device=dev, dtype=torch.float32)
    vm = None
    if vmask is not None:            # any non-zero value counts as valid (a bare uint8 cast would wrap 256 -> 0, -1 -> 255)
        vm = (vmask.to(dev) != 0).to(torch.uint8).contiguous()
    stream = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(_lib.lib().s1s2_tile_extract(idx, scene.data_ptr(), vm.data_ptr() if vm is not None else None, SH, SW,
                                            org_d.data_ptr(), N, ps, cond.data_ptr(), mask.data_ptr(), ratio.data_ptr(),
                                            C.c_void_p(stream)))
    return cond, mask, ratio


FILTER_CODES = ("keep", "valid_ratio", "variance", "dark", "texture")


def tile_filter(scene: torch.Tensor, target: torch.Tensor, origins, ps: int, colloc: torch.Tensor = None,
                valid_ratio_threshold=0.80, variance_threshold=1e-4, dark_thr=0.10, dark_max_ratio=0.60, texture_thr=5e-5):
    """Patch.py's window filters (Patch.py:205-224, defaults :327-337) on the device: stats f32[N,8] =
    valid_ratio, var[0..3], dark_fraction, laplacian_var, code (index into FILTER_CODES; 0 = the window is kept)."""
    dev = scene.device
    idx = _dev_index(scene)
    scene = scene.to(torch.float32).contiguous()
    target = target.to(device=dev, dtype=torch.float32).contiguous()
    if target.ndim != 3 or target.shape[0] != 4 or target.shape[1:] != scene.shape[1:]:
        raise ValueError(f"target must be f32[4,H,W] on the scene's grid, got {tuple(target.shape)}")
    org = torch.as_tensor(np.ascontiguousarray(origins, dtype=np.int32)).reshape(-1, 2)
    N = org.shape[0]
    SH, SW = scene.shape[1:]
    if N and (int(org.min()) < 0 or int(org[:, 0].max()) + ps > SH or int(org[:, 1].max()) + ps > SW):
        raise ValueError("window outside the scene")
    org_d = org.to(dev)
    # Patch.py:41-49 tests `colloc > 0` on the float raster: 0.5 is valid, -1 is not, 256 does not wrap
    cl = (colloc.to(dev) > 0).to(torch.uint8).contiguous() if colloc is not None else None
    stats = torch.empty((N, 8), device=dev, dtype=torch.float32)
    th = (C.c_float * 5)(valid_ratio_threshold, variance_threshold, dark_thr, dark_max_ratio, texture_thr)
    stream = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(_lib.lib().s1s2_tile_filter(idx, scene.data_ptr(), scene.shape[0], target.data_ptr(),
                                           cl.data_ptr() if cl is not None else None, SH, SW, org_d.data_ptr(), N, ps, th,
                                           stats.data_ptr(), C.c_void_p(stream)))
    return stats


def hann_window(ps: int, device=None) -> torch.Tensor:
    """Separable blend window f32[ps]: w[i] = 0.5 - 0.5 cos(2 pi (i + 0.5) / ps) -- the Hann window sampled at pixel
    centres, strictly positive, so every covered pixel keeps a non-zero weight sum."""
    i = torch.arange(ps, dtype=torch.float64)
    w = (0.5 - 0.5 * torch.cos(2.0 * torch.pi * (i + 0.5) / ps)).to(torch.float32)
    return w.to(device) if device is not None else w


def stitch(preds: torch.Tensor, origins, ps: int, stride: int, SH: int, SW: int, window=None):
    """preds f32[N,C,ps,ps] (cuda) + origins on the stride grid -> (canvas f32[C,SH,SW], cover u8[SH,SW]).

    window: None = uniform weights (the primary, bit-reproducible definition), "hann" or a f32[ps] tensor = separable
    per-pixel weights w[ly] * w[lx] (SURVEY.md section 8 a9's optional variant)."""
    dev = preds.device
    idx = _dev_index(preds)
    preds = preds.to(torch.float32).contiguous()
    org = np.ascontiguousarray(origins, dtype=np.int32).reshape(-1, 2)
    N, Cn = preds.shape[0], preds.shape[1]
    if org.shape[0] != N:
        raise ValueError("one origin per patch")
    if N:
        if (org % stride).any() or org.min() < 0 or org[:, 0].max() + ps > SH or org[:, 1].max() + ps > SW:
            raise ValueError("origins must lie on the stride grid inside the scene")
        key = org[:, 0].astype(np.int64) * (SW + 1) + org[:, 1]
        if (np.diff(key) <= 0).any():
            raise ValueError("origins must be strictly ascending in (row, col) order (Patch.py iteration order)")
    win = None
    if window is not None:
        win = hann_window(ps, dev) if isinstance(window, str) and window == "hann" else window
        if not isinstance(win, torch.Tensor) or win.numel() != ps:
            raise ValueError("window must be None, 'hann' or a tensor of ps weights")
        win = win.to(device=dev, dtype=torch.float32).contiguous()
    org_d = torch.as_tensor(org).to(dev)
    canvas = torch.empty((Cn, SH, SW), device=dev, dtype=torch.float32)
    cover = torch.empty((SH, SW), device=dev, dtype=torch.uint8)
    stream = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(_lib.lib().s1s2_stitch_weighted(idx, preds.data_ptr(), org_d.data_ptr(), N, Cn, ps, stride, SH, SW,
                                               win.data_ptr() if win is not None else None, canvas.data_ptr(),
                                               cover.data_ptr(), C.c_void_p(stream)))
    return canvas, cover
