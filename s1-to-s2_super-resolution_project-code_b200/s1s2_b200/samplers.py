"""The reference's sampler entry points, routed to the fused CUDA sampling loop (s1s2_sample).

Same names, argument order and return values as the reference functions; every function additionally accepts
``noise=`` (the initial unit-normal draw; drawn with ``torch.randn`` on the input's device when omitted, like the
reference) and, for the stochastic samplers, ``step_noise=`` (f32[n,B,C,H,W], one slice per noisy step in
execution order).  Without ``step_noise`` the per-step noise is generated inside the last kernel of each model call
(Philox4x32-10, ``seed=``; drawn from torch's generator when omitted): a DDPM-1000 chain at batch 64 would otherwise
need 64 GB of pre-drawn z.  One library call enqueues the whole loop: no per-step host synchronisation, no torch ops.

  ddpm_ddim_generate     Evaluation_Updated/Evaluation_Pure_Generation.py:277-292
  ddim_multistep_eval    Evaluation/DDIM_Multi-step.py:116-137
  ddim_multistep_eval_v  Evaluation/DDIM_Multi-step_v_Prediction.py:137-178
  one_step_recon[_v]     Evaluation/DDIM_Multi-step.py:155-170, DDIM_Multi-step_v_Prediction.py:211-227
  ddim_sample / ddpm_sample / partial_ddim_from_gt   Evaluation/Limitation_Test.py:209-270
  sample_ddim_v / sample_ddpm_v                      Evaluation/Limitation_Test_v_Prediction.py:210-254
"""
import ctypes as C

import torch

from . import _lib, schedule
from .metrics import masked_mae, masked_mse
from .model import UNetSmallB200


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _f32c(t, dev):
    return t.to(device=dev, dtype=torch.float32).contiguous()


def run_steps(model: UNetSmallB200, steps, cond, x_init, init_scale=1.0, step_noise=None, tap_pred=False, tap_x=False,
              seed=None, patch_base=0):
    """Enqueue one sampling loop (list of _lib.Step) on the current stream.

    Returns the result tensor f32[B,C,H,W] (valid when the stream drains), or (result, taps) when taps are asked
    for: taps["pred"] f32[n,B,C,H,W] = network output of every call, taps["x"] = state after every update."""
    if not isinstance(model, UNetSmallB200):
        raise TypeError("the fused samplers need a UNetSmallB200 model (drop-in for the reference's UNetSmall)")
    dev = cond.device
    B, _, H, W = cond.shape
    eng = model.engine(dev, H, W, B)
    cond = _f32c(cond, dev)
    x_init = _f32c(x_init, dev)
    if tuple(x_init.shape) != (B, model.out_ch, H, W):
        raise ValueError(f"initial state must be f32[{B},{model.out_ch},{H},{W}], got {tuple(x_init.shape)}")
    n = len(steps)
    arr = (_lib.Step * n)(*steps)
    n_noise = sum(1 for s in steps if s.flags & _lib.STEP_NOISE)
    if n_noise:
        if step_noise is None:
            step_noise = torch.randn((n_noise, B, model.out_ch, H, W), device=dev)
        step_noise = _f32c(step_noise, dev)
        if step_noise.shape[0] < n_noise:
            raise ValueError(f"step_noise holds {step_noise.shape[0]} slices, the chain needs {n_noise}")
    out = torch.empty_like(x_init)
    taps = {}
    if tap_pred:
        taps["pred"] = torch.empty((n,) + tuple(x_init.shape), device=dev, dtype=torch.float32)
    if tap_x:
        taps["x"] = torch.empty((n,) + tuple(x_init.shape), device=dev, dtype=torch.float32)
    if any(s.flags & _lib.STEP_PHILOX for s in steps):      # in-kernel noise: key it (drawn from torch's generator if not given)
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        _lib.check(_lib.lib().s1s2_set_noise_seed(eng.h, C.c_uint64(int(seed)), C.c_uint32(int(patch_base))), eng.h)
    stream = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(_lib.lib().s1s2_sample(eng.h, arr, n, _ptr(cond), _ptr(x_init), float(init_scale),
                                      _ptr(step_noise) if n_noise else None, _ptr(out), _ptr(taps.get("pred")),
                                      _ptr(taps.get("x")), B, C.c_void_p(stream)), eng.h)
    return (out, taps) if taps else out


def host_result_buffer(model: UNetSmallB200, shape):
    """The pinned host tensor run_steps_host returns results of this shape in (allocated on first use, then reused)."""
    key = ("host_out", tuple(shape))
    cache = model.__dict__.setdefault("_host_buffers", {})
    out = cache.get(key)
    if out is None:
        out = cache[key] = torch.empty(tuple(shape), dtype=torch.float32, pin_memory=True)
    return out


def run_steps_host(model: UNetSmallB200, steps, cond_host, x_init_host, init_scale=1.0, device=None, batch=None):
    """End-to-end entry with HOST tensors (pinned recommended): N patches in batches of `batch` (default: all at once, at
    most the model's max_batch) through s1s2_sample_host_stream -- upload of batch i+1 and download of batch i-1 overlap the
    model calls of batch i; returns a pinned host tensor (reused across calls) once everything has drained."""
    dev = torch.device(device if device is not None else "cuda")
    N, _, H, W = cond_host.shape
    if batch is None:
        batch = min(N, max(model.max_batch, 1))
    eng = model.engine(dev, H, W, batch)
    assert cond_host.device.type == "cpu" and x_init_host.device.type == "cpu"
    cond_host = cond_host.to(torch.float32).contiguous()
    x_init_host = x_init_host.to(torch.float32).contiguous()
    out = host_result_buffer(model, x_init_host.shape)     # reused across calls (the caller copies out of it if it keeps it)
    n = len(steps)
    arr = (_lib.Step * n)(*steps)
    stream = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(_lib.lib().s1s2_sample_host_stream(eng.h, arr, n, _ptr(cond_host), _ptr(x_init_host), float(init_scale),
                                                  _ptr(out), N, int(batch), C.c_void_p(stream)), eng.h)
    return out


def _randn(shape, dev, noise):
    return noise if noise is not None else torch.randn(shape, device=dev)


# ---------------------------------------------------------------------------------------------- eps, grid A
@torch.no_grad()
def ddpm_ddim_generate(model, x_cond, alpha_bar, t_start=200, steps=20, noise=None):
    Ct = model.outc.out_channels
    x_T = _randn((x_cond.size(0), Ct, x_cond.size(2), x_cond.size(3)), x_cond.device, noise)
    return run_steps(model, schedule.steps_eps_grid_a(alpha_bar, t_start, steps), x_cond, x_T)


@torch.no_grad()
def ddim_multistep_eval(model, x_gt, x_cond, alpha_bar, mask, t_start=200, steps=20, noise=None):
    t_start = max(1, min(int(t_start), len(alpha_bar) - 1))
    a_t = alpha_bar[t_start].to(x_gt.device).view(-1, 1, 1, 1)
    x_t = torch.sqrt(a_t) * x_gt + torch.sqrt(1 - a_t) * _randn(x_gt.shape, x_gt.device, noise)
    x0 = run_steps(model, schedule.steps_eps_grid_a(alpha_bar, t_start, steps), x_cond, x_t)
    return masked_mae(x0, x_gt, mask), masked_mse(x0, x_gt, mask), x0


def _one_step_noise(x_gt, rng_seed, noise):
    """The reference's draw: ``if rng_seed is not None: torch.manual_seed(rng_seed)`` then ``torch.randn_like(x_gt)``
    (DDIM_Multi-step.py:157,162) -- the global generator is re-seeded, exactly as the scripts do."""
    if noise is not None:
        return noise
    if rng_seed is not None:
        torch.manual_seed(int(rng_seed))
    return torch.randn_like(x_gt)


@torch.no_grad()
def one_step_recon(model, x_gt, x_cond, alpha_bar, mask, t_small, rng_seed=None, noise=None):
    T = len(alpha_bar)
    t = max(1, min(int(t_small), T - 1))
    noise = _one_step_noise(x_gt, rng_seed, noise)
    a = alpha_bar[t].to(x_gt.device).view(-1, 1, 1, 1)
    x_t = torch.sqrt(a) * x_gt + torch.sqrt(1 - a) * noise
    ab = schedule._abar_cpu(alpha_bar)
    st = [_lib.Step(t, _lib.STEP_EPS_DDIM, _lib.STEP_FINAL, -1, float(torch.sqrt(1 - ab[t])),
                    float(torch.sqrt(ab[t] + 1e-8)), 0.0, 0.0, 0.0)]
    x0 = run_steps(model, st, x_cond, x_t)
    return masked_mae(x0, x_gt, mask), masked_mse(x0, x_gt, mask), x0


# ---------------------------------------------------------------------------------------------- v, grid B
@torch.no_grad()
def ddim_multistep_eval_v(model, x_gt, x_cond, alpha_bar, mask, t_start=200, steps=20, eta: float = 0.0, noise=None,
                          step_noise=None, seed=None):
    T = len(alpha_bar)
    t_start = max(1, min(int(t_start), T - 1))
    idxs = schedule.grid_b(t_start, steps)
    ab = schedule._abar_cpu(alpha_bar)
    x0 = run_steps(model, schedule.steps_grid_b(alpha_bar, idxs, "v", eta=float(eta), philox=step_noise is None), x_cond,
                   _randn(x_gt.shape, x_gt.device, noise), init_scale=float(torch.sqrt(1 - ab[t_start])),
                   step_noise=step_noise, seed=seed)
    return masked_mae(x0, x_gt, mask), masked_mse(x0, x_gt, mask), x0


@torch.no_grad()
def one_step_recon_v(model, x_gt, x_cond, alpha_bar, mask, t_small, rng_seed=None, noise=None, allow_t0=False):
    """DDIM_Multi-step_v_Prediction.py:211-227: t_small is clamped to [1, T-1] like the reference.  ``allow_t0=True`` (not
    in the reference's signature) lets the t = 0 identity check of Onestep_v_Prediction.py:184-197 run through here."""
    T = len(alpha_bar)
    t = max(0 if allow_t0 else 1, min(int(t_small), T - 1))
    noise = _one_step_noise(x_gt, rng_seed, noise)
    a = alpha_bar[t].to(x_gt.device).view(-1, 1, 1, 1)
    x_t = torch.sqrt(a) * x_gt + torch.sqrt(1 - a) * noise
    ab = schedule._abar_cpu(alpha_bar)
    st = [_lib.Step(t, _lib.STEP_V_DDIM, _lib.STEP_FINAL, -1, float(torch.sqrt(ab[t])), float(torch.sqrt(1.0 - ab[t])),
                    0.0, 0.0, 0.0)]
    x0 = run_steps(model, st, x_cond, x_t)
    return masked_mae(x0, x_gt, mask), masked_mse(x0, x_gt, mask), x0


@torch.no_grad()
def sample_ddim_v(model, cond, alpha_bar, C_tgt, steps=250, eta=0.05, t_start=None, noise=None, step_noise=None, seed=None):
    T = len(alpha_bar)
    B, _, H, W = cond.shape
    K = T - 1 if t_start is None else int(max(1, min(int(t_start), T - 1)))
    ab = schedule._abar_cpu(alpha_bar)
    idxs = schedule.grid_b(K, steps)
    return run_steps(model, schedule.steps_grid_b(alpha_bar, idxs, "v", eta=float(eta), philox=step_noise is None), cond,
                     _randn((B, C_tgt, H, W), cond.device, noise), init_scale=float(torch.sqrt(1 - ab[K])),
                     step_noise=step_noise, seed=seed)


# ---------------------------------------------------------------------------------------------- eps, grid B / DDPM
@torch.no_grad()
def ddim_sample(model, cond, alphas, alpha_bar, C_tgt, steps=50, noise=None):
    T = len(alphas)
    B, _, H, W = cond.shape
    idxs = schedule.grid_b(T - 1, steps, force_append=False)
    return run_steps(model, schedule.steps_grid_b(alpha_bar, idxs, "eps"), cond, _randn((B, C_tgt, H, W), cond.device, noise))


@torch.no_grad()
def ddpm_sample(model, cond, betas, alphas, alpha_bar, C_tgt, noise=None, step_noise=None, t_list=None, seed=None):
    B, _, H, W = cond.shape
    return run_steps(model, schedule.steps_ddpm(betas, alphas, alpha_bar, "eps", t_list, philox=step_noise is None), cond,
                     _randn((B, C_tgt, H, W), cond.device, noise), step_noise=step_noise, seed=seed)


@torch.no_grad()
def sample_ddpm_v(model, cond, betas, alphas, alpha_bar, C_tgt, noise=None, step_noise=None, t_list=None, seed=None):
    B, _, H, W = cond.shape
    return run_steps(model, schedule.steps_ddpm(betas, alphas, alpha_bar, "v", t_list, philox=step_noise is None), cond,
                     _randn((B, C_tgt, H, W), cond.device, noise), step_noise=step_noise, seed=seed)


@torch.no_grad()
def partial_ddim_from_gt(model, x_gt, cond, alpha_bar, k: int, noise=None):
    k = int(max(0, min(k, len(alpha_bar) - 1)))
    a_t = alpha_bar[k].to(x_gt.device).view(1, 1, 1, 1)
    x_t = torch.sqrt(a_t) * x_gt + torch.sqrt(1 - a_t) * _randn(x_gt.shape, x_gt.device, noise)
    if k == 0:
        return torch.clamp(x_t, 0.0, 1.0)
    ab = schedule._abar_cpu(alpha_bar)
    st = []
    for cur in range(k, 0, -1):
        a_cur, a_prev = ab[cur], ab[cur - 1]
        # every step keeps x' = sqrt(a_prev) x0 + sqrt(1-a_prev) eps; the result is clamp(x') of the last one
        st.append(_lib.Step(cur, _lib.STEP_EPS_DDIM, 0, -1, float(torch.sqrt(1 - a_cur)), float(torch.sqrt(a_cur + 1e-8)),
                            float(torch.sqrt(a_prev)), float(torch.sqrt(1 - a_prev)), 0.0))
    x = run_steps(model, st, cond, x_t)
    return torch.clamp(x, 0.0, 1.0)
