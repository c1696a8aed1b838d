"""Functional fp32 restatement of the reference denoiser (oracle; test infrastructure only).

Restates ``UNetSmall.__init__/forward`` -- Evaluation/DDIM_Multi-step.py:19-53 (the same class is
re-declared in every reference script).  The network is a plain 3-level UNet: 3x3 conv + bias + ReLU
pairs, 2x2 max-pool, 2x2/stride-2 transposed conv, channel-concat skips and a 1x1 head.  The integer
timestep enters as one extra constant input plane.

Written against a flat ``state_dict`` (the 34 tensors the reference checkpoint holds) instead of an
``nn.Module`` so that the same weights can drive the oracle and the CUDA path.
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F

# (state_dict prefix, kind, Cin multiplier, Cout multiplier) in execution order; multipliers are in units of
# base_ch except where noted.  kind: c3 = 3x3 conv pad 1, up = ConvTranspose2d k2 s2, c1 = 1x1 conv.
_ENC = ("down1", "down2", "down3")
_DEC = (("up3", "conv3"), ("up2", "conv2"), ("up1", "conv1"))


def param_shapes(in_ch: int, out_ch: int, base_ch: int = 96) -> "OrderedDict[str, tuple]":
    """Names and shapes of the reference checkpoint, in ``state_dict()`` order."""
    b = base_ch
    s = OrderedDict()

    def conv(name, co, ci, k):
        s[name + ".weight"] = (co, ci, k, k)
        s[name + ".bias"] = (co,)

    conv("inc.0", b, in_ch + 1, 3)
    c = b
    for lvl in _ENC:
        conv(f"{lvl}.0.0", 2 * c, c, 3)
        conv(f"{lvl}.0.2", 2 * c, 2 * c, 3)
        c *= 2
    for up, blk in _DEC:
        s[up + ".weight"] = (c, c // 2, 2, 2)       # ConvTranspose2d layout: [Cin, Cout, kh, kw]
        s[up + ".bias"] = (c // 2,)
        conv(f"{blk}.0", c // 2, c, 3)
        conv(f"{blk}.2", c // 2, c // 2, 3)
        c //= 2
    conv("outc", out_ch, b, 1)
    return s


def init_state_dict(in_ch: int = 8, out_ch: int = 4, base_ch: int = 96, seed: int = 1234,
                    generator: torch.Generator = None) -> "OrderedDict[str, torch.Tensor]":
    """Random weights with PyTorch's default Conv init statistics (uniform +-1/sqrt(fan_in)).

    This is NOT bit-identical to ``torch.manual_seed(seed); UNetSmall(...)`` (a synthetic stand-in for the
    missing ``Models/*.pth`` blobs); it only has to be a plausible checkpoint that both sides load.
    """
    g = generator or torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    shapes = param_shapes(in_ch, out_ch, base_ch)
    for name, shp in shapes.items():
        if name.endswith(".weight"):
            fan_in = shp[1] * shp[2] * shp[3]           # torch takes fan_in from dim 1 (also for ConvTranspose)
            bound = 1.0 / fan_in ** 0.5
            sd[name] = (torch.rand(shp, generator=g) * 2 - 1) * bound
        else:
            sd[name] = (torch.rand(shp, generator=g) * 2 - 1) * bound
    return sd


def _pair(sd, prefix, x):
    x = F.relu(F.conv2d(x, sd[f"{prefix}.0.weight"], sd[f"{prefix}.0.bias"], padding=1))
    return F.relu(F.conv2d(x, sd[f"{prefix}.2.weight"], sd[f"{prefix}.2.bias"], padding=1))


def forward(sd, xt_and_cond: torch.Tensor, t_idx: torch.Tensor, taps: dict = None) -> torch.Tensor:
    """eps/v prediction f32[B,out_ch,H,W] for f32[B,in_ch,H,W] and i64[B] (fp32 throughout).

    ``taps`` (optional dict) receives every intermediate activation, keyed by the layer that produced it,
    for per-layer parity checks of the CUDA kernels.
    """
    B, _, H, W = xt_and_cond.shape
    tplane = t_idx.reshape(B, 1, 1, 1).to(torch.float32).expand(B, 1, H, W)
    x = torch.cat([xt_and_cond, tplane], dim=1)
    rec = (lambda k, v: taps.__setitem__(k, v)) if taps is not None else (lambda k, v: None)

    e = [F.relu(F.conv2d(x, sd["inc.0.weight"], sd["inc.0.bias"], padding=1))]
    rec("inc", e[0])
    for lvl in _ENC:
        h = _pair(sd, f"{lvl}.0", e[-1])
        rec(lvl + ".pre_pool", h)
        e.append(F.max_pool2d(h, 2))
        rec(lvl, e[-1])
    d = e[3]
    for i, (up, blk) in enumerate(_DEC):
        u = F.conv_transpose2d(d, sd[up + ".weight"], sd[up + ".bias"], stride=2)
        rec(up, u)
        d = _pair(sd, blk, torch.cat([u, e[2 - i]], dim=1))
        rec(blk, d)
    out = F.conv2d(d, sd["outc.weight"], sd["outc.bias"])
    rec("outc", out)
    return out


class OracleModel:
    """Callable with the reference's model surface: ``m(xt_and_cond, t_idx)`` and ``m.outc.out_channels``."""

    class _Head:
        def __init__(self, n):
            self.out_channels = n

    def __init__(self, sd):
        self.sd = sd
        self.outc = OracleModel._Head(sd["outc.weight"].shape[0])
        self.calls = []          # [(x_in, t_idx, out)] when recording
        self.record = False

    def __call__(self, xt_and_cond, t_idx):
        with torch.no_grad():
            out = forward(self.sd, xt_and_cond, t_idx)
        if self.record:
            self.calls.append((xt_and_cond.clone(), t_idx.clone(), out.clone()))
        return out
