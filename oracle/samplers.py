"""Sampler loops (oracle; test infrastructure only).

Every function takes ``model`` = any callable ``model(xt_and_cond f32[B,8,H,W], t_idx i64[B]) -> f32[B,4,H,W]``
and the *initial noise as an explicit tensor* (the reference draws it from torch's global generator; the
north-star contract is "same supplied initial-noise tensor").  Stochastic variants take the per-step noise
as a callable ``step_noise(i) -> tensor`` for the same reason.

Restates:
  * eps DDIM, grid A, result = clamp(last x0_hat)      -- Evaluation/DDIM_Multi-step.py:116-137 (from noised GT),
                                                          Evaluation_Updated/Evaluation_Pure_Generation.py:277-292 (from noise)
  * eps DDIM, grid B (K=T-1), result = clamp(x0 @ t=0)  -- Evaluation/Limitation_Test.py:227-249
  * v   DDIM, grid B, eta >= 0                          -- Evaluation/DDIM_Multi-step_v_Prediction.py:137-178,
                                                          Evaluation/Limitation_Test_v_Prediction.py:229-254
  * v -> (x0, eps)                                      -- Evaluation/DDIM_Multi-step_v_Prediction.py:59-65
  * DDPM ancestral (eps and v)                          -- Evaluation/Limitation_Test.py:209-224,
                                                          Evaluation/Limitation_Test_v_Prediction.py:210-226
  * partial DDIM from GT                                -- Evaluation/Limitation_Test.py:252-270
  * one-step reconstruction (eps / v)                   -- Evaluation/Onestep.py:149-160,
                                                          Evaluation/Onestep_v_Prediction.py:58-71
All arithmetic is fp32 in torch with the reference's operation order (no fused multiply-add).
"""
import torch

from .schedule import grid_a, grid_b


def _bc(a):
    return a.reshape(-1, 1, 1, 1)


def v_to_x0_eps(x_t, v, abar_t):
    sa, sb = _bc(torch.sqrt(abar_t)), _bc(torch.sqrt(1.0 - abar_t))
    return sa * x_t - sb * v, sb * x_t + sa * v


def eps_to_x0(x_t, eps, abar_t):
    return (x_t - torch.sqrt(1 - abar_t) * eps) / torch.sqrt(abar_t + 1e-8)


def _tvec(t, B):
    return torch.full((B,), int(t), dtype=torch.long)


@torch.no_grad()
def ddim_eps_grid_a(model, cond, alpha_bar, x_init, t_start, steps, trace=None):
    """x_init is x_{t_start} (unit normal for pure generation; the noised GT for the recon evaluators)."""
    ts = grid_a(t_start, steps)
    x_t, B = x_init.clone(), cond.shape[0]
    x0 = None
    for i in range(steps):
        a_cur, a_nxt = alpha_bar[ts[i]], alpha_bar[ts[i + 1]]
        eps = model(torch.cat([x_t, cond], 1), _tvec(ts[i], B))
        x0 = eps_to_x0(x_t, eps, a_cur)
        x_new = torch.sqrt(a_nxt) * x0 + torch.sqrt(1 - a_nxt) * eps
        if trace is not None:
            trace.append(dict(t=int(ts[i]), x_in=x_t, pred=eps, x0=x0, x_out=x_new))
        x_t = x_new
    return torch.clamp(x0, 0.0, 1.0)


def noise_gt(x_gt, alpha_bar, t, noise):
    """x_t = sqrt(abar_t) x_gt + sqrt(1-abar_t) noise  (Evaluation/DDIM_Multi-step.py:120-123)."""
    a = _bc(alpha_bar[torch.as_tensor([int(t)])])
    return torch.sqrt(a) * x_gt + torch.sqrt(1 - a) * noise


@torch.no_grad()
def ddim_eps_grid_b(model, cond, alpha_bar, x_init, steps, trace=None):
    T, B = len(alpha_bar), cond.shape[0]
    idx = grid_b(T - 1, steps, force_append=False)
    x_t = x_init.clone()
    for i in reversed(range(len(idx))):
        t = int(idx[i])
        eps = model(torch.cat([x_t, cond], 1), _tvec(t, B))
        x0 = eps_to_x0(x_t, eps, alpha_bar[t])
        if i == 0:
            x_new = x0
        else:
            a_prev = alpha_bar[int(idx[i - 1])]
            x_new = torch.sqrt(a_prev) * x0 + torch.sqrt(1 - a_prev) * eps
        if trace is not None:
            trace.append(dict(t=t, x_in=x_t, pred=eps, x0=x0, x_out=x_new))
        x_t = x_new
    return torch.clamp(x_t, 0.0, 1.0)


@torch.no_grad()
def ddim_v_grid_b(model, cond, alpha_bar, noise, steps, t_start=None, eta=0.0, step_noise=None, trace=None):
    """``noise`` is the unit-normal draw; it is scaled by sqrt(1-abar_K) here like the reference does."""
    T, B = len(alpha_bar), cond.shape[0]
    K = T - 1 if t_start is None else int(max(1, min(int(t_start), T - 1)))
    idx = grid_b(K, steps)
    x_t = noise * torch.sqrt(1 - alpha_bar[K])
    for i in reversed(range(len(idx))):
        t = int(idx[i])
        a_t = alpha_bar[t]
        v = model(torch.cat([x_t, cond], 1), _tvec(t, B))
        x0, eps = v_to_x0_eps(x_t, v, a_t)
        if i == 0:
            x_new = x0
        else:
            a_prev = alpha_bar[int(idx[i - 1])]
            if eta == 0.0 and step_noise is None:
                x_new = torch.sqrt(a_prev) * x0 + torch.sqrt(1 - a_prev) * eps
            else:
                sigma = eta * torch.sqrt((1 - a_prev) / (1 - a_t + 1e-8) * (1 - a_t / a_prev).clamp_min(0))
                dirc = torch.sqrt((1 - a_prev) - sigma ** 2).clamp_min(0)
                x_new = torch.sqrt(a_prev) * x0 + dirc * eps + sigma * step_noise(i)
        if trace is not None:
            trace.append(dict(t=t, x_in=x_t, pred=v, x0=x0, x_out=x_new))
        x_t = x_new
    return torch.clamp(x_t, 0.0, 1.0)


@torch.no_grad()
def ddpm_ancestral(model, cond, betas, alphas, alpha_bar, x_init, step_noise, param="eps", t_list=None, trace=None):
    """Full-chain ancestral sampling; ``step_noise(t)`` supplies z for every t > 0.  ``t_list`` (descending)
    restricts the chain for tests (the reference always runs T-1 .. 0)."""
    B = cond.shape[0]
    x_t = x_init.clone()
    ts = list(reversed(range(len(betas)))) if t_list is None else list(t_list)
    for t in ts:
        out = model(torch.cat([x_t, cond], 1), _tvec(t, B))
        eps = out if param == "eps" else v_to_x0_eps(x_t, out, alpha_bar[t])[1]
        mean = (1 / torch.sqrt(alphas[t])) * (x_t - (betas[t] / torch.sqrt(1 - alpha_bar[t] + 1e-8)) * eps)
        x_new = mean + torch.sqrt(betas[t]) * step_noise(t) if t > 0 else mean
        if trace is not None:
            trace.append(dict(t=t, x_in=x_t, pred=out, x_out=x_new))
        x_t = x_new
    return torch.clamp(x_t, 0.0, 1.0)


@torch.no_grad()
def partial_ddim_from_gt(model, x_gt, cond, alpha_bar, k, noise):
    k = int(max(0, min(k, len(alpha_bar) - 1)))
    a = alpha_bar[k].reshape(1, 1, 1, 1)
    x_t = torch.sqrt(a) * x_gt + torch.sqrt(1 - a) * noise
    B = cond.shape[0]
    for cur in range(k, 0, -1):
        eps = model(torch.cat([x_t, cond], 1), _tvec(cur, B))
        x0 = eps_to_x0(x_t, eps, alpha_bar[cur])
        a_prev = alpha_bar[cur - 1]
        x_t = torch.sqrt(a_prev) * x0 + torch.sqrt(1 - a_prev) * eps
    return torch.clamp(x_t, 0.0, 1.0)


@torch.no_grad()
def one_step_eps(model, x_gt, cond, alpha_bar, t_small, noise):
    T = len(alpha_bar)
    t = max(1, min(int(t_small), T - 1))
    x_t = noise_gt(x_gt, alpha_bar, t, noise)
    eps = model(torch.cat([x_t, cond], 1), _tvec(t, cond.shape[0]))
    a = _bc(alpha_bar[torch.as_tensor([t])])
    return torch.clamp(eps_to_x0(x_t, eps, a), 0.0, 1.0), eps, x_t


@torch.no_grad()
def one_step_v(model, x_gt, cond, alpha_bar, t_small, noise):
    T = len(alpha_bar)
    t = max(0, min(int(t_small), T - 1))
    x_t = noise_gt(x_gt, alpha_bar, t, noise)
    v = model(torch.cat([x_t, cond], 1), _tvec(t, cond.shape[0]))
    x0, _ = v_to_x0_eps(x_t, v, alpha_bar[torch.as_tensor([t])])
    return torch.clamp(x0, 0.0, 1.0), v, x_t
