"""Noise schedules and timestep grids, computed on the host exactly as the reference scripts do, plus the
translation of a sampler configuration into the per-step coefficient records the fused epilogue consumes.

Reference: cosine_beta_schedule (Evaluation/DDIM_Multi-step.py:9-16), linear_beta_schedule / make_schedule
(Evaluation/Limitation_Test.py:22-31), alpha_bar derivation (Evaluation/DDIM_Multi-step.py:210-212), grid
convention A (Evaluation/DDIM_Multi-step.py:124) and B (Evaluation/DDIM_Multi-step_v_Prediction.py:147-151).
All coefficient arithmetic runs in float32 torch on the CPU so that it rounds like the reference's own
elementwise code (sqrt of float32 values in float32).
"""
import math

import torch

from . import _lib


def cosine_beta_schedule(T: int, s: float = 0.008) -> torch.Tensor:
    steps = T + 1
    x = torch.linspace(0, T, steps, dtype=torch.float64)
    f = torch.cos(((x / T) + s) / (1 + s) * math.pi * 0.5) ** 2
    f = f / f[0]
    betas = 1 - (f[1:] / f[:-1])
    return torch.clip(betas, 1e-5, 0.999).float()


def linear_beta_schedule(T: int, beta_start: float = 1e-4, beta_end: float = 0.02) -> torch.Tensor:
    return torch.linspace(beta_start, beta_end, T, dtype=torch.float32)


def make_schedule(T: int, kind: str = "cosine") -> torch.Tensor:
    """betas, like the reference's make_schedule (Limitation_Test.py:25-31)."""
    if kind == "cosine":
        return cosine_beta_schedule(T)
    if kind == "linear":
        return linear_beta_schedule(T)
    raise ValueError(f"unknown schedule '{kind}'")


def derive(betas: torch.Tensor):
    """(betas, alphas, alpha_bar) in float32; cumprod in float32 like the reference."""
    alphas = 1.0 - betas
    return betas, alphas, torch.cumprod(alphas, dim=0)


def grid_a(t_start: int, steps: int) -> torch.Tensor:
    """Convention A: integer linspace t_start..0 with steps+1 entries (CPU torch, the reference's own call)."""
    return torch.linspace(int(t_start), 0, int(steps) + 1, dtype=torch.long)


def grid_b(K: int, steps: int, force_append: bool = True) -> torch.Tensor:
    """Convention B: ascending unique(round(linspace(0, K, steps))) [+ K]."""
    idxs = torch.unique(torch.round(torch.linspace(0, int(K), int(steps))).to(torch.long), sorted=True)
    if force_append and idxs[-1].item() != K:
        idxs = torch.unique(torch.cat([idxs, torch.tensor([int(K)], dtype=torch.long)]), sorted=True)
    return idxs


def _f(x) -> float:
    return float(x)


def _abar_cpu(alpha_bar: torch.Tensor) -> torch.Tensor:
    return alpha_bar.detach().to("cpu", torch.float32)


def steps_eps_grid_a(alpha_bar, t_start: int, steps: int):
    """ddpm_ddim_generate / ddim_multistep_eval: `steps` calls at ts[0..steps-1]; result = clamp(last x0_hat)."""
    ab = _abar_cpu(alpha_bar)
    ts = grid_a(t_start, steps)
    out = []
    for i in range(steps):
        a_cur, a_next = ab[ts[i]], ab[ts[i + 1]]
        out.append(_lib.Step(int(ts[i]), _lib.STEP_EPS_DDIM, _lib.STEP_FINAL if i == steps - 1 else 0, -1,
                             _f(torch.sqrt(1 - a_cur)), _f(torch.sqrt(a_cur + 1e-8)),
                             _f(torch.sqrt(a_next)), _f(torch.sqrt(1 - a_next)), 0.0))
    return out


def steps_grid_b(alpha_bar, idxs, param: str, eta: float = 0.0, stochastic_form: bool = False, philox: bool = False):
    """ddim_sample (eps) / ddim_multistep_eval_v / sample_ddim_v (v): descending walk over idxs, last call at
    idxs[0] returns clamp(x0).  With eta > 0 (or stochastic_form) every non-final step adds sigma * z where
    z = step_noise[k], k counting the non-final steps in execution order (philox=True: z is generated in the kernel,
    k is its Philox stream index)."""
    ab = _abar_cpu(alpha_bar)
    kind = _lib.STEP_EPS_DDIM if param == "eps" else _lib.STEP_V_DDIM
    out, k = [], 0
    for i in reversed(range(len(idxs))):
        t = int(idxs[i])
        a_t = ab[t]
        if param == "eps":
            c0, c1 = _f(torch.sqrt(1 - a_t)), _f(torch.sqrt(a_t + 1e-8))
        else:
            c0, c1 = _f(torch.sqrt(a_t)), _f(torch.sqrt(1.0 - a_t))
        if i == 0:
            out.append(_lib.Step(t, kind, _lib.STEP_FINAL, -1, c0, c1, 0.0, 0.0, 0.0))
            continue
        a_prev = ab[int(idxs[i - 1])]
        if eta == 0.0 and not stochastic_form:
            out.append(_lib.Step(t, kind, 0, -1, c0, c1, _f(torch.sqrt(a_prev)), _f(torch.sqrt(1 - a_prev)), 0.0))
        else:
            sigma = eta * torch.sqrt((1 - a_prev) / (1 - a_t + 1e-8) * (1 - a_t / a_prev).clamp_min(0))
            dirc = torch.sqrt((1 - a_prev) - sigma ** 2).clamp_min(0)
            out.append(_lib.Step(t, kind, _lib.STEP_PHILOX if philox else _lib.STEP_NOISE, k, c0, c1,
                                 _f(torch.sqrt(a_prev)), _f(dirc), _f(sigma)))
            k += 1
    return out


def steps_ddpm(betas, alphas, alpha_bar, param: str, t_list=None, philox: bool = False):
    """ddpm_sample / sample_ddpm_v: ancestral chain over t_list (default T-1..0); z for step t>0 is
    step_noise[k], k counting in execution order; result = clamp(x after the last step)."""
    b, a, ab = (x.detach().to("cpu", torch.float32) for x in (betas, alphas, alpha_bar))
    ts = list(reversed(range(len(b)))) if t_list is None else [int(t) for t in t_list]
    kind = _lib.STEP_EPS_DDPM if param == "eps" else _lib.STEP_V_DDPM
    out, k = [], 0
    for n, t in enumerate(ts):
        flags = (_lib.STEP_FINAL if n == len(ts) - 1 else 0) | ((_lib.STEP_PHILOX if philox else _lib.STEP_NOISE) if t > 0 else 0)
        c0, c1 = (_f(torch.sqrt(ab[t])), _f(torch.sqrt(1.0 - ab[t]))) if param == "v" else (0.0, 0.0)
        out.append(_lib.Step(t, kind, flags, k if t > 0 else -1, c0, c1, _f(1 / torch.sqrt(a[t])),
                             _f(b[t] / torch.sqrt(1 - ab[t] + 1e-8)), _f(torch.sqrt(b[t]))))
        if t > 0:
            k += 1
    return out
