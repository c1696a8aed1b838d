"""Noise schedule and timestep grids (oracle; test infrastructure only).

Restates:
  * ``cosine_beta_schedule``  -- Evaluation/DDIM_Multi-step.py:9-16
  * ``linear_beta_schedule``  -- Evaluation/Limitation_Test.py:22-23
  * alphas / alpha_bar derivation -- Evaluation/DDIM_Multi-step.py:210-212
  * grid convention A (eps scripts)  -- Evaluation/DDIM_Multi-step.py:124,
    Evaluation_Updated/Evaluation_Pure_Generation.py:282
  * grid convention B (v scripts / Limitation_Test*) --
    Evaluation/DDIM_Multi-step_v_Prediction.py:147-151,
    Evaluation/Limitation_Test_v_Prediction.py:233-239,
    Evaluation/Limitation_Test.py:234-236
"""
import math

import torch


def cosine_betas(T: int, s: float = 0.008) -> torch.Tensor:
    """Nichol-Dhariwal cosine betas: fp64 internally, clipped to [1e-5, 0.999], returned fp32."""
    u = torch.linspace(0, T, T + 1, dtype=torch.float64)
    g = torch.cos(((u / T + s) / (1 + s)) * math.pi / 2) ** 2
    abar64 = g / g[0]
    b = 1 - abar64[1:] / abar64[:-1]
    return torch.clip(b, 1e-5, 0.999).float()


def linear_betas(T: int, beta_start: float = 1e-4, beta_end: float = 2e-2) -> torch.Tensor:
    return torch.linspace(beta_start, beta_end, T, dtype=torch.float32)


def make_schedule(T: int = 1000, kind: str = "cosine"):
    """Returns (betas, alphas, alpha_bar), all fp32; the cumprod runs in fp32 like the reference."""
    betas = cosine_betas(T) if kind == "cosine" else linear_betas(T)
    alphas = 1.0 - betas
    return betas, alphas, torch.cumprod(alphas, dim=0)


def grid_a(t_start: int, steps: int) -> torch.Tensor:
    """Convention A: integer-dtype linspace t_start -> 0 with steps+1 entries (truncation toward zero).

    The model is evaluated at entries [0..steps-1]; entry [steps] (=0) only supplies alpha_bar_next.
    """
    return torch.linspace(t_start, 0, steps + 1, dtype=torch.long)


def grid_b(K: int, steps: int, force_append: bool = True) -> torch.Tensor:
    """Convention B: ascending unique(round(linspace(0, K, steps))), with K appended if missing."""
    g = torch.unique(torch.round(torch.linspace(0, K, steps)).to(torch.long), sorted=True)
    if force_append and int(g[-1]) != K:
        g = torch.unique(torch.cat([g, torch.tensor([K], dtype=torch.long)]), sorted=True)
    return g
