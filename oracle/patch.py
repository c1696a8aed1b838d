"""Tiling, per-patch normalisation and overlap-blend stitch (oracle; test infrastructure only).

Restates (numpy, like the reference):
  * ``patch_iter``      -- Patch.py:80-84   (row-major sliding window; right/bottom remainder not covered)
  * ``build_mask``      -- Patch.py:41-49   (all inputs finite [and target finite, collocation > 0])
  * window slicing      -- Patch.py:201-203
  * valid-ratio filter  -- Patch.py:205-209
  * ``zscore_inplace``  -- Patch.py:51-62 applied to HH, HV over valid pixels (Patch.py:228-229)
  * incidence/90, elevation/1000 with nan->0 -- Patch.py:231-232
  * invalid -> 0 and nan/inf -> 0 -- Patch.py:236-239

``stitch`` is NOT in the reference: parity unpinned.  Definition adopted:
canvas[c,y,x] = sum_p w * pred_p[c, y-r_p, x-c_p] / sum_p w with uniform w = 1, patches accumulated in
index order; pixels covered by no patch are 0 and flagged 0 in the coverage mask.
"""
import numpy as np


def tile_origins(H: int, W: int, ps: int, stride: int) -> np.ndarray:
    """int32[N,2] (row, col) in the reference's iteration order."""
    rows = range(0, H - ps + 1, stride)
    cols = range(0, W - ps + 1, stride)
    return np.array([(r, c) for r in rows for c in cols], dtype=np.int32).reshape(-1, 2)


def valid_mask(inputs: np.ndarray, target: np.ndarray = None, colloc: np.ndarray = None) -> np.ndarray:
    m = np.isfinite(inputs).all(axis=0)
    if target is not None:
        m &= np.isfinite(target).all(axis=0)
    if colloc is not None:
        m &= colloc > 0
    return m


def _zscore(x: np.ndarray, m: np.ndarray) -> None:
    if m is None or not m.any():
        mu, sd = np.nanmean(x), np.nanstd(x)
    else:
        mu, sd = float(np.nanmean(x[m])), float(np.nanstd(x[m]))
    if not np.isfinite(mu):
        mu = 0.0
    if not np.isfinite(sd) or sd < 1e-6:
        sd = 1.0
    x -= mu
    x /= sd


def extract_patch(inputs: np.ndarray, vmask: np.ndarray, row: int, col: int, ps: int):
    """One window -> (cond f32[4,ps,ps], mask u8[ps,ps], valid_ratio).  inputs = (HH dB, HV dB, incidence deg,
    elevation m) as float32 [4,H,W]."""
    X = inputs[:, row:row + ps, col:col + ps].astype(np.float32).copy()
    M = vmask[row:row + ps, col:col + ps].copy()
    vr = float(M.mean()) if M.size else 0.0
    with np.errstate(all="ignore"):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            _zscore(X[0], M)
            _zscore(X[1], M)
    X[2] = np.nan_to_num(X[2], nan=0.0) / 90.0
    X[3] = np.nan_to_num(X[3], nan=0.0) / 1000.0
    for ch in range(X.shape[0]):
        a = X[ch]
        a[~M] = 0.0
        X[ch] = np.nan_to_num(a, nan=0.0, posinf=0.0, neginf=0.0).astype(np.float32)
    return X, M.astype(np.uint8), vr


def extract_all(inputs: np.ndarray, vmask: np.ndarray, ps: int, stride: int, valid_ratio_threshold: float = 0.0):
    """All windows in order; windows whose valid ratio is below the threshold are skipped (Patch.py:205-209)."""
    H, W = inputs.shape[1:]
    idx, conds, masks = [], [], []
    for r, c in tile_origins(H, W, ps, stride):
        X, M, vr = extract_patch(inputs, vmask, int(r), int(c), ps)
        if vr < valid_ratio_threshold:
            continue
        idx.append((r, c)); conds.append(X); masks.append(M)
    n = len(idx)
    return (np.array(idx, dtype=np.int32).reshape(n, 2),
            np.stack(conds) if n else np.zeros((0, inputs.shape[0], ps, ps), np.float32),
            np.stack(masks) if n else np.zeros((0, ps, ps), np.uint8))


# ---------------------------------------------------------------------------------------------- quality filters
def band_variances(Y: np.ndarray, M: np.ndarray) -> np.ndarray:
    """Per-band variance of the target over valid pixels (Patch.py:211-214 uses np.nanvar(Y[ch][M]))."""
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return np.array([np.nanvar(Y[ch][M]) for ch in range(Y.shape[0])], dtype=np.float64)


def dark_fraction(Y: np.ndarray, M: np.ndarray, thr: float = 0.10) -> float:
    """Patch.py:88-98: share of valid pixels with mean(B2,B3,B4) < thr and B8 < thr; 1.0 for an empty mask."""
    if not M.any():
        return 1.0
    vis = (Y[0] + Y[1] + Y[2]) / 3.0
    return float(((vis < thr) & (Y[3] < thr) & M).sum()) / float(M.sum())


def laplacian_var(img: np.ndarray, M: np.ndarray) -> float:
    """Patch.py:100-114: variance over valid pixels of the 5-point Laplacian with a symmetric patch boundary
    (scipy.signal.convolve2d(mode="same", boundary="symm") restated with np.pad); NaN stencils are ignored (nanvar)."""
    import warnings
    a = img.astype(np.float32).copy()
    bad = ~np.isfinite(a)
    if (bad & M).any():
        a[bad] = np.nanmean(a[M])
    q = np.pad(a, 1, mode="symmetric")
    with np.errstate(invalid="ignore"):
        L = q[:-2, 1:-1] + q[2:, 1:-1] + q[1:-1, :-2] + q[1:-1, 2:] - 4.0 * q[1:-1, 1:-1]
        # convolve2d multiplies the four corner samples by the kernel's zeros: a non-finite corner still poisons L
        L = L + 0.0 * (q[:-2, :-2] + q[:-2, 2:] + q[2:, :-2] + q[2:, 2:])
    if not M.any():
        return 0.0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return float(np.nanvar(L[M]))


def filter_decision(Y, M, valid_ratio_threshold=0.80, variance_threshold=1e-4, dark_thr=0.10, dark_max_ratio=0.60,
                    texture_thr=5e-5):
    """The four tests of Patch.py:205-224 in order; returns (code, stats): code 0 = keep, 1 = valid ratio, 2 = flat
    target, 3 = too dark, 4 = no texture; stats = (valid_ratio, var[0..C-1], dark_fraction, laplacian_var)."""
    vr = float(M.mean()) if M.size else 0.0
    var = band_variances(Y, M)
    dk = dark_fraction(Y, M, dark_thr)
    lv = laplacian_var(Y[3], M)
    code = 0
    if vr < valid_ratio_threshold:
        code = 1
    elif all(v < variance_threshold for v in var):
        code = 2
    elif dk > dark_max_ratio:
        code = 3
    elif lv < texture_thr:
        code = 4
    return code, (vr, var, dk, lv)


def hann_window(ps: int) -> np.ndarray:
    """w[i] = 0.5 - 0.5 cos(2 pi (i + 0.5) / ps), computed in float64 and rounded to float32 (strictly positive)."""
    i = np.arange(ps, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * (i + 0.5) / ps)).astype(np.float32)


def stitch(preds: np.ndarray, origins: np.ndarray, H: int, W: int, window: np.ndarray = None):
    """Overlap blend.  preds f32[N,C,ps,ps], origins i32[N,2] -> (canvas f32[C,H,W], cover u8[H,W]).

    window None: uniform weights.  window f32[ps]: separable weights w[ly] * w[lx].  Accumulates in fp32 in patch-index
    order and divides once, which is the exact arithmetic the CUDA gather kernel performs per output pixel.
    PARITY UNPINNED: the reference has no stitch; this states the definition of DESIGN.md."""
    N, C, ps, _ = preds.shape
    acc = np.zeros((C, H, W), np.float32)
    cnt = np.zeros((H, W), np.float32)
    w2 = None if window is None else (window.astype(np.float32)[:, None] * window.astype(np.float32)[None, :]).astype(np.float32)
    for p in range(N):
        r, c = int(origins[p, 0]), int(origins[p, 1])
        if w2 is None:
            acc[:, r:r + ps, c:c + ps] += preds[p]
            cnt[r:r + ps, c:c + ps] += 1.0
        else:
            acc[:, r:r + ps, c:c + ps] += (w2[None] * preds[p]).astype(np.float32)
            cnt[r:r + ps, c:c + ps] += w2
    out = np.where(cnt > 0, acc / np.where(cnt > 0, cnt, 1.0).astype(np.float32), 0.0).astype(np.float32)
    return out, (cnt > 0).astype(np.uint8)
