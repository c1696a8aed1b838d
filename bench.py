#!/usr/bin/env python
"""DDIM-50 patches/sec of the S1->S2 sampling path on N B200s (one process per GPU), next to the reference's CPU path and
the library path (PyTorch eager + cuDNN) it would take on the same GPU.

A "step" is one complete DDIM-50 sampling (50 fused model calls + scheduler updates) of one batch of synthetic 256x256
patches per GPU.  Workloads (BASELINE.json configs):
  v64     v-prediction UNet, grid B (0..999, 50 entries), eta=0, batch 64 per GPU     [default; north-star target]
  eps16   eps-prediction UNet, grid A (999 -> 0, 50 calls), batch 16 per GPU
  sweep   v-prediction UNet, grid B with 10 / 25 / 50 / 100 / 250 steps (config 4, DDIM_Sweep): one JSON line with a `sweep`
          table of ms per model call and patches/s per step count (`value` = the 50-step row)
  scene   one 4x2048x2048 scene tiled by Patch.py's rule, patch-sharded over the ranks, gathered and stitched (config 5)
  latency the reference scripts' own operating point: batch 1 (and 2 / 4 / 8 / 16), DDIM-50, through (i) the fused
          s1s2_sample loop and (ii) the 2-line drop-in of INTEGRATION.md (UNetSmallB200.forward + the script's own
          torch scheduler), next to PyTorch eager + cuDNN at batch 1
Patches are independent units: N GPUs = N x batch patches per step, no data-path collective ("weak" scaling).

The default line carries, besides the contract keys:
  e2e       the same metric through the host-buffer entry s1s2_sample_host_stream (pinned host cond + noise in, image out)
  roofline  tensor-pipe roofline of the conv kernel family (every launch in the timed region is one instantiation of it):
            algorithmic FLOPs (301 851 475 968 per patch per model call) / device time, against
            MEASURED_PEAKS.json's sustained bf16 figure; `layers` = every launch of one model call timed with CUDA event pairs
  cpu_baseline          oracle/ (a port of the reference's PyTorch sampler) on this box's host cores, one whole chain
  gpu_library_baseline  the same oracle module on `cuda` = PyTorch eager + cuDNN, the path the unmodified reference takes on
                        this GPU: TF32 batch 1 (the scripts as written) and fp16 autocast + channels_last batch 64 (best case)
  latency   the batch 1..16 table of the `latency` workload (rank 0, N = 1)
  other_configs  BASELINE configs 1, 2 and 4 in short form (N = 1): one denoiser call at batch 1, eps16, the step-count sweep
  scene     config 5 at this N: one 2048^2 scene at stride 64 and at Patch.py's default stride 32, with per-phase times
            and the sha256 of the stitched canvas (must not depend on N)

`--impl reference` times only the CPU oracle port (the reference is Python/PyTorch and does not travel to the GPU box;
oracle/ is its restatement, pinned against vectors the real reference produced: tests/golden/).
"""
import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "s1-to-s2_super-resolution_project-code_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

FLOP_PER_CALL = 301_851_475_968          # per 256x256 patch per model call (measured on the reference)
H = W = 256
N_CALLS = 50
SEED_V, SEED_EPS = 1235, 1234            # synthetic stand-ins for Models/ddpm_s1_to_s2_upgraded_v.pth / _v3.pth


def layer_flops():
    """2*M*N*K of every launch of one model call (execution order), per patch; sums to FLOP_PER_CALL."""
    b = 96
    rows = [("inc.0", 256, b, 9 * 9)]
    c, s = b, 256
    for lvl in ("down1", "down2", "down3"):
        rows += [(f"{lvl}.0.0", s, 2 * c, 9 * c), (f"{lvl}.0.2", s, 2 * c, 9 * 2 * c)]
        c, s = 2 * c, s // 2
    for up, blk in (("up3", "conv3"), ("up2", "conv2"), ("up1", "conv1")):
        rows += [(up, s, 4 * (c // 2), c)]
        s *= 2
        rows += [(f"{blk}.0", s, c // 2, 9 * c), (f"{blk}.2", s, c // 2, 9 * (c // 2))]
        c //= 2
    out = [(n, 2 * side * side * N * K) for n, side, N, K in rows]
    out[-1] = (out[-1][0], out[-1][1] + 2 * 256 * 256 * 4 * 96)      # outc rides in conv1.2's epilogue
    assert sum(f for _, f in out) == FLOP_PER_CALL
    return out


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["bf16_tflops_sustained"]), float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


def csrc_fingerprint():
    """sha256 (16 hex) over the kernel sources with comments and whitespace removed: an ncu capture is only quoted while
    the code it profiled is unchanged (editing a comment does not invalidate it; editing a kernel does)."""
    import re
    h = hashlib.sha256()
    d = os.path.join(PKG, "csrc")
    for name in sorted(os.listdir(d)):
        if name.endswith((".cu", ".cuh")):
            text = open(os.path.join(d, name), encoding="utf-8").read()
            text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)          # block comments
            text = re.sub(r"//[^\n]*", "", text)                       # line comments (no string literal here holds "//")
            h.update(name.encode())
            h.update(re.sub(r"\s+", "", text).encode())
    return h.hexdigest()[:16]


def synthetic_batch(B, seed):
    import torch
    g = torch.Generator().manual_seed(seed)
    cond = torch.randn((B, 4, H, W), generator=g)
    cond[:, 2] = torch.rand((B, H, W), generator=g) * 0.4 + 0.2
    cond[:, 3] = (torch.randn((B, H, W), generator=g) * 0.3 + 0.3).abs()
    noise = torch.randn((B, 4, H, W), generator=g)
    return cond, noise


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                      stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
            out, _ = self.p.communicate()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [c for c, w in zip(sm, power) if w >= 0.5 * max(power)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "power_w_max": max(power), "samples": len(sm)}


def make_steps(workload, abar):
    from s1s2_b200 import schedule
    import torch
    if workload == "eps16":
        return schedule.steps_eps_grid_a(abar, 999, 50), 1.0
    steps = schedule.steps_grid_b(abar, schedule.grid_b(999, 50), "v")
    return steps, float(torch.sqrt(1 - abar[999]))


def build_model(seed, max_batch, dev):
    """UNetSmallB200 with the synthetic checkpoint of `seed` loaded the way the reference scripts load theirs."""
    import s1s2_b200
    sd = s1s2_b200.synthetic_checkpoint(seed)
    model = s1s2_b200.UNetSmallB200(8, 4, 96, max_batch=max_batch).to(dev)
    model.load_state_dict({k: v.to(dev) for k, v in sd.items()}, strict=True)
    return model.eval()


# ====================================================================================== baselines (measurement only)
def cpu_oracle_sample(workload, n_calls, threads):
    """Times `n_calls` model calls + scheduler updates of the DDIM-50 chain of ONE patch in the CPU oracle; returns
    seconds.  (bench.py's cpu_baseline / reference-arm leg: the one place outside tests/ that executes oracle/.)"""
    import torch
    from oracle import samplers as osamplers, schedule as osched, unet as ounet
    torch.set_num_threads(threads)
    sd = ounet.init_state_dict(8, 4, 96, seed=SEED_EPS if workload == "eps16" else SEED_V)
    model = ounet.OracleModel(sd)
    _, _, abar = osched.make_schedule(1000)
    cond, noise = synthetic_batch(1, 2024)
    calls = {"n": 0}

    class Stop(Exception):
        pass

    def counted(x, t):
        if calls["n"] >= n_calls:
            raise Stop()
        calls["n"] += 1
        return model(x, t)
    counted.outc = model.outc
    t0 = time.perf_counter()
    try:
        if workload == "eps16":
            osamplers.ddim_eps_grid_a(counted, cond, abar, noise, 999, 50)
        else:
            osamplers.ddim_v_grid_b(counted, cond, abar, noise, 50)
    except Stop:
        pass
    return time.perf_counter() - t0


def gpu_library_baseline(dev, budget_s=12.0):
    """The path the UNMODIFIED reference takes on this GPU: eager PyTorch dispatching to cuDNN.
    The reference tree does not travel to the GPU box, so this runs oracle/'s restatement of the same module and v-DDIM loop
    (pinned against the reference's outputs) on `cuda`.  Measurement only -- nothing here is on the product path.
      tf32_b1      fp32 tensors, cudnn.allow_tf32 = True (PyTorch default), batch 1: the batch-1 scripts as written
      fp16_cl_b64  torch.autocast(float16) (Limitation_Test_v_Prediction.py:205-207) + channels_last, batch 64: the best
                   cuDNN configuration found (tools/cudnn_baseline.py tries the others)"""
    import torch
    from oracle import samplers as osamplers, schedule as osched, unet as ounet
    sd = ounet.init_state_dict(8, 4, 96, seed=SEED_V)
    _, _, abar = osched.make_schedule(1000)
    abar_d = abar.to(dev)
    rows = []
    t_begin = time.perf_counter()
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    try:
        for mode, B, n_calls in (("tf32", 1, 50), ("fp16_cl", 64, 50)):
            if time.perf_counter() - t_begin > budget_s:
                rows.append({"mode": mode, "batch": B, "skipped": "time budget"})
                continue
            torch.backends.cudnn.allow_tf32 = True
            torch.backends.cuda.matmul.allow_tf32 = True
            torch.backends.cudnn.benchmark = True
            cond, noise = synthetic_batch(B, 2024)
            cond, noise = cond.to(dev), noise.to(dev)
            sdd = {k: v.to(dev) for k, v in sd.items()}
            if mode == "fp16_cl":
                sdd = {k: (v.contiguous(memory_format=torch.channels_last) if v.ndim == 4 else v) for k, v in sdd.items()}
                cond = cond.contiguous(memory_format=torch.channels_last)
                noise = noise.contiguous(memory_format=torch.channels_last)
            base = ounet.OracleModel(sdd)
            calls = {"n": 0, "limit": 2}

            class Stop(Exception):
                pass

            def model(x, t, base=base, calls=calls, mode=mode):
                if calls["n"] >= calls["limit"]:
                    raise Stop()
                calls["n"] += 1
                if mode == "fp16_cl":
                    with torch.autocast("cuda", dtype=torch.float16):
                        return base(x, t.to(x.device)).float()
                return base(x, t.to(x.device))
            model.outc = base.outc

            def chain(limit):
                calls["n"], calls["limit"] = 0, limit
                try:
                    osamplers.ddim_v_grid_b(model, cond, abar_d, noise, 50)
                except Stop:
                    pass
            try:
                chain(3)                                       # cuDNN autotune + warm-up
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                chain(n_calls)
                e1.record()
                torch.cuda.synchronize()
                ms_call = e0.elapsed_time(e1) / calls["n"]
                rows.append({"mode": mode, "batch": B, "model_calls_timed": calls["n"], "ms_per_model_call": round(ms_call, 3),
                             "patches_per_s": round(B / (ms_call * N_CALLS / 1e3), 3),
                             "tflops": round(FLOP_PER_CALL * B / (ms_call / 1e3) / 1e12, 1)})
            except RuntimeError as e:                          # e.g. out of memory next to the product arena
                rows.append({"mode": mode, "batch": B, "error": str(e)[:160]})
            del base, sdd, cond, noise
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = saved
    return {"what": f"oracle/ restatement of the reference's UNetSmall + v-DDIM loop on cuda: PyTorch {torch.__version__} "
                    f"eager + cuDNN {torch.backends.cudnn.version()} (the library path of the unmodified reference); DDIM-50 patches/s",
            "rows": rows, "wall_s": round(time.perf_counter() - t_begin, 1)}


def run_reference(args, rank, world):
    """Reference arm: the reference's own CPU implementation of the path (oracle/ port), all host threads.  Each step is a
    bounded sample -- the first `n_calls` of the 50 model calls (+ scheduler updates) of one patch's chain; `ms_per_step` is
    what was really timed per step and `value` is the DDIM-50 patches/s that per-call time implies."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_calls = 10
    cpu_oracle_sample(args.workload, 1, threads)                       # page in torch / oneDNN
    for _ in range(args.warmup):
        cpu_oracle_sample(args.workload, 1, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_oracle_sample(args.workload, n_calls, threads)
    dt = time.perf_counter() - t0
    s_per_call = dt / (args.steps * n_calls)
    pps = 1.0 / (s_per_call * N_CALLS)
    sample = (f"each step = the first {n_calls} of the 50 model calls (+ scheduler updates) of one 256x256 patch's chain, "
              f"oracle/ fp32 PyTorch CPU, {threads} threads; ms_per_step is the time of that sample; value = 1 / (50 x the "
              f"measured time per model call) -- a whole chain measured the same way agrees within 4 % (cpu_baseline of the main arm)")
    cfg = config_block(args, world)
    cfg["arithmetic"] = "f32 (PyTorch CPU: oneDNN convolutions, ATen elementwise)"
    line = {"impl": "reference", "metric": "DDIM-50 patches/sec", "value": pps, "unit": "patches/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg, "sample_model_calls_per_step": n_calls, "model_calls_per_patch": N_CALLS,
            "cpu_baseline": {"value": pps, "unit": "patches/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": pps, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ====================================================================================== blocks of the default line
def script_style_v_ddim(model, cond, abar_d, noise, idxs):
    """The reference script's own sampling loop with only the model swapped ("2-line change"): what
    DDIM_Multi-step_v_Prediction.py:153-175 executes per step -- torch.cat, one model call, the v -> (x0, eps) conversion and
    the DDIM update as eager torch elementwise ops -- here around UNetSmallB200.forward (s1s2_forward)."""
    import torch
    B = cond.shape[0]
    K = int(idxs[-1])
    x = noise * torch.sqrt(1 - abar_d[K])
    for i in reversed(range(len(idxs))):
        t = int(idxs[i])
        t_idx = torch.full((B,), t, dtype=torch.long, device=cond.device)
        a_t = abar_d[t]
        v = model(torch.cat([x, cond], dim=1), t_idx)
        sa, sb = torch.sqrt(a_t), torch.sqrt(1 - a_t)
        x0 = sa * x - sb * v
        eps = sb * x + sa * v
        if i == 0:
            x = x0
            break
        a_prev = abar_d[int(idxs[i - 1])]
        x = torch.sqrt(a_prev) * x0 + torch.sqrt(1 - a_prev) * eps
    return torch.clamp(x, 0.0, 1.0)


def latency_table(model, abar, dev, batches=(1, 2, 4, 8, 16), target_s=0.8):
    """DDIM-50 at small batches (the reference scripts run batch 1): per batch the fused loop (s1s2_sample: one library call
    enqueues 50 model calls) and the script-style drop-in loop; device time by CUDA events over whole chains, plus the host
    time spent enqueueing (when it exceeds the device time the path is launch-bound)."""
    import torch
    from s1s2_b200 import samplers, schedule
    peak_tf, _, _ = peaks()
    idxs = schedule.grid_b(999, 50)
    steps = schedule.steps_grid_b(abar, idxs, "v")
    init_scale = float(torch.sqrt(1 - abar[999]))
    abar_d = abar.to(dev)
    rows = []
    for B in batches:
        cond_h, noise_h = synthetic_batch(B, 4000 + B)
        cond, noise = cond_h.to(dev), noise_h.to(dev)
        row = {"batch": B}
        for name, fn in (("fused", lambda: samplers.run_steps(model, steps, cond, noise, init_scale=init_scale)),
                         ("dropin", lambda: script_style_v_ddim(model, cond, abar_d, noise, idxs))):
            for _ in range(2):
                out = fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            out = fn()
            e1.record()
            host_one = time.perf_counter() - t0
            torch.cuda.synchronize()
            one = e0.elapsed_time(e1) / 1e3
            reps = max(2, min(80, int(target_s / max(one, 1e-4))))
            clocks = ClockSampler(dev.index or 0)
            t0 = time.perf_counter()
            e0.record()
            for _ in range(reps):
                out = fn()
            e1.record()
            host = time.perf_counter() - t0
            torch.cuda.synchronize()
            clk = clocks.stop()
            ms_chain = e0.elapsed_time(e1) / reps
            assert bool(torch.isfinite(out).all())
            tf = FLOP_PER_CALL * N_CALLS * B / (ms_chain / 1e3) / 1e12
            row[name] = {"chains_timed": reps, "ms_per_chain": round(ms_chain, 3), "ms_per_model_call": round(ms_chain / N_CALLS, 4),
                         "patches_per_s": round(B / (ms_chain / 1e3), 2), "tflops": round(tf, 1), "frac_of_peak": round(tf / peak_tf, 4),
                         "host_enqueue_ms_per_chain": round(host / reps * 1e3, 3), "first_chain_host_ms": round(host_one * 1e3, 3),
                         "sm_mhz": clk.get("sm_mhz"), "power_w_max": clk.get("power_w_max"), "clock_reasons": clk.get("reasons")}
        if B == 1:
            res_f = samplers.run_steps(model, steps, cond, noise, init_scale=init_scale)
            res_d = script_style_v_ddim(model, cond, abar_d, noise, idxs)
            row["fused_vs_dropin_max_abs_diff"] = float((res_f - res_d).abs().max())
        rows.append(row)
    return rows


def other_configs_block(model_v, abar, dev):
    """The remaining BASELINE configs in the driver-visible line (N = 1): config 1 as the host-visible latency of one
    denoiser call at batch 1 (Onestep.py:149-164 runs exactly one), config 2 (eps model, grid A 999 -> 0, batch 16) and
    config 4 (v model, 10 / 25 / 50 / 100 / 250 steps at batch 64); one timed chain each, device time by CUDA events."""
    import torch
    from s1s2_b200 import samplers, schedule
    out = {}
    # config 1
    cond, noise = synthetic_batch(1, 5001)
    x = torch.cat([noise, cond], 1).to(dev)
    t = torch.full((1,), 20, dtype=torch.long, device=dev)
    for _ in range(5):
        model_v(x, t)
    torch.cuda.synchronize()
    lat = []
    for _ in range(30):
        t0 = time.perf_counter()
        y = model_v(x, t)
        torch.cuda.synchronize()
        lat.append((time.perf_counter() - t0) * 1e3)
    out["onestep_b1"] = {"what": "one model(torch.cat([x_t, x_cond], 1), t_idx) call at batch 1, host-visible (call + synchronize)",
                         "ms_median": round(statistics.median(lat), 4), "ms_min": round(min(lat), 4), "finite": bool(torch.isfinite(y).all())}
    # config 4
    B = model_v.max_batch
    cond_h, noise_h = synthetic_batch(B, 2024)
    cond_d, noise_d = cond_h.to(dev), noise_h.to(dev)
    init_scale = float(torch.sqrt(1 - abar[999]))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rows = []
    samplers.run_steps(model_v, schedule.steps_grid_b(abar, schedule.grid_b(999, 50), "v"), cond_d, noise_d, init_scale=init_scale)
    torch.cuda.synchronize()                       # (back at the sustained clock after the batch-1 calls above)
    for n_steps in (10, 25, 50, 100, 250):
        steps = schedule.steps_grid_b(abar, schedule.grid_b(999, n_steps), "v")
        e0.record()
        res = samplers.run_steps(model_v, steps, cond_d, noise_d, init_scale=init_scale)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        rows.append({"ddim_steps": n_steps, "model_calls": len(steps), "ms_per_chain": round(ms, 2),
                     "ms_per_model_call": round(ms / len(steps), 4), "patches_per_s": round(B / (ms / 1e3), 2)})
        assert bool(torch.isfinite(res).all())
    out["sweep_v64"] = {"what": "DDIM_Sweep (config 4): v model, grid B from 999, batch %d, one chain per step count" % B, "rows": rows}
    del cond_d, noise_d
    # config 2
    model_e = build_model(SEED_EPS, 16, dev)
    steps = schedule.steps_eps_grid_a(abar, 999, 50)
    cond_h, noise_h = synthetic_batch(16, 2025)
    cond_d, noise_d = cond_h.to(dev), noise_h.to(dev)
    samplers.run_steps(model_e, steps, cond_d, noise_d)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        res = samplers.run_steps(model_e, steps, cond_d, noise_d)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    out["eps16"] = {"what": "DDIM_Multi-step / Evaluation_Pure_Generation (config 2): eps model, grid A 999 -> 0 (50 calls), batch 16",
                    "ms_per_chain": round(ms, 2), "patches_per_s": round(16 / (ms / 1e3), 2),
                    "frac_of_peak": round(FLOP_PER_CALL * N_CALLS * 16 / (ms / 1e3) / 1e12 / peaks()[0], 4),
                    "finite": bool(torch.isfinite(res).all())}
    del model_e
    torch.cuda.empty_cache()
    return out


def scene_block(model, abar, dev, rank, world, batch):
    """BASELINE config 5 at this world size: one synthetic 4 x 2048 x 2048 Sentinel-1 scene held in pinned HOST memory, uploaded,
    tiled by Patch.py's rule (256 / stride 64 = 841 windows, and Patch.py's default stride 32 = 3249), patch-sharded over the
    ranks, v-DDIM-50, NCCL gather to rank 0, overlap-blend stitch, canvas downloaded.  Strong scaling: the whole scene is the
    unit.  canvas_sha256 must be the same for every world size (pure sharding, noise keyed by global patch index)."""
    import torch
    import torch.distributed as dist
    from s1s2_b200 import scene as sc
    scn = sc.synthetic_scene(2048, 2048, seed=0).pin_memory()
    if world > 1:                                  # communicator set-up (first gather) stays outside the timed scenes
        sc.gather_to_rank0(torch.zeros((1, 4, 8, 8), device=dev), [1] * world, rank, world)
        torch.cuda.synchronize()
    out = []
    for stride in (64, 32):
        tm = {}
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res, up, down = sc.generate_scene_host(model, scn, abar, dev, rank=rank, world=world, ps=256, stride=stride, param="v",
                                               steps=50, t_start=999, batch=batch, timings=tm)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        keys = ("extract_ms", "noise_ms", "sample_ms", "sync_ms", "gather_ms", "stitch_ms")
        vec = torch.tensor([ms] + [tm.get(k, 0.0) for k in keys], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(vec, op=dist.ReduceOp.MAX)
        vec = vec.tolist()
        if rank == 0:
            n = int(res["kept"].sum())
            sha = hashlib.sha256(res["canvas"].numpy().tobytes()).hexdigest()
            row = {"stride": stride, "patches": n, "ms_total": round(vec[0], 2), "patches_per_s": round(n / (vec[0] / 1e3), 2)}
            row.update({k: round(v, 3) for k, v in zip(keys, vec[1:])})
            row.update({"h2d_bytes": up, "d2h_bytes": down, "canvas_sha256": sha,
                        "covered_fraction": round(float(res["cover"].float().mean()), 6)})
            out.append(row)
        del res
        torch.cuda.empty_cache()
    return {"what": "Evaluation_Pure_Generation over a whole scene (config 5): 4x2048x2048 synthetic Sentinel-1 scene in pinned host "
                    "memory -> upload -> Patch.py tiling -> patch-sharded v-DDIM-50 -> NCCL gather -> overlap-blend stitch -> download; "
                    "phase times = max over ranks (CUDA events)", "n_gpus": world, "batch_per_gpu": batch, "rows": out} if rank == 0 else None


# ====================================================================================== stand-alone workloads
def run_scene(args, rank, world, dev):
    import torch
    import torch.distributed as dist
    from s1s2_b200 import schedule
    model = build_model(SEED_V, args.batch, dev)
    _, _, abar = schedule.derive(schedule.cosine_beta_schedule(1000))
    model.engine(dev, 256, 256, args.batch)
    steps, init_scale = make_steps("v64", abar)
    from s1s2_b200 import samplers
    cond_h, noise_h = synthetic_batch(args.batch, 2024 + rank)
    for _ in range(max(1, args.warmup // 3)):                         # clocks and caches warm before the scene is timed
        samplers.run_steps(model, steps, cond_h.to(dev), noise_h.to(dev), init_scale=init_scale)
    torch.cuda.synchronize()
    clocks = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    l0 = model.launch_count()
    blk = scene_block(model, abar, dev, rank, world, args.batch)
    clk = clocks.stop()
    if rank == 0:
        r64 = blk["rows"][0]
        peak_tf, _, peak_src = peaks()
        achieved = r64["patches"] / (r64["sample_ms"] / 1e3) * N_CALLS * FLOP_PER_CALL / 1e12 / world
        line = {"metric": "DDIM-50 patches/sec", "value": r64["patches_per_s"], "unit": "patches/s", "n_gpus": world, "steps": 1,
                "warmup": args.warmup, "ms_per_step": r64["ms_total"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f16", "data": "synthetic",
                "config": {"workload": "Evaluation_Pure_Generation over a whole scene: 4x2048x2048 synthetic Sentinel-1 scene, Patch.py "
                           "tiling 256/stride 64 (841 patches), v-DDIM-50, patch-sharded, NCCL gather to rank 0, overlap-blend stitch",
                           "batch_per_gpu": args.batch, "patches": r64["patches"], "parallelism": f"patch-sharded x{world}, one gather",
                           "l2": "working set exceeds L2"},
                "e2e": {"value": r64["patches_per_s"], "unit": "patches/s", "h2d_bytes_per_step": r64["h2d_bytes"],
                        "d2h_bytes_per_step": r64["d2h_bytes"],
                        "entry": "s1s2_b200.scene.generate_scene_host (pinned host scene in, host canvas out)"},
                "gpu_launches": int(model.launch_count() - l0), "clocks": clk, "scene": blk,
                "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                             "peak_source": peak_src, "traffic": None,
                             "kernel": "conv kernel family, per GPU, over the sampling phase of the stride-64 scene (slowest rank)"},
                "cpu_baseline": None}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_sweep(args, rank, world, dev):
    """BASELINE config 4: v-prediction model, grid B from K = 999 with 10 / 25 / 50 / 100 / 250 steps (len(idxs) model calls
    each, t = 0 included), batch 64 per GPU; per-call latency and patches/s per step count."""
    import torch
    import torch.distributed as dist
    from s1s2_b200 import samplers, schedule
    B = args.batch
    model = build_model(SEED_V, B, dev)
    _, _, abar = schedule.derive(schedule.cosine_beta_schedule(1000))
    cond_h, noise_h = synthetic_batch(B, 2024 + rank)
    cond_d, noise_d = cond_h.to(dev), noise_h.to(dev)
    init_scale = float(torch.sqrt(1 - abar[999]))

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
    table, clocks_all, launches = [], [], 0
    for n_steps in (10, 25, 50, 100, 250):
        steps = schedule.steps_grid_b(abar, schedule.grid_b(999, n_steps), "v")
        reps = max(1, min(args.steps, 500 // len(steps)))
        for _ in range(max(1, 150 // len(steps))):
            out = samplers.run_steps(model, steps, cond_d, noise_d, init_scale=init_scale)
        sync()
        clocks = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
        l0 = model.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = samplers.run_steps(model, steps, cond_d, noise_d, init_scale=init_scale)
        e1.record()
        sync()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        launches += int(model.launch_count() - l0)
        clocks_all.append(clocks.stop())
        assert bool(torch.isfinite(out).all())
        table.append({"ddim_steps": n_steps, "model_calls": len(steps), "chains_timed": reps, "ms_per_chain": ms / reps,
                      "ms_per_model_call": ms / reps / len(steps), "patches_per_s": B * world * reps / (ms / 1e3),
                      "tflops_per_gpu": FLOP_PER_CALL * len(steps) * B * reps / (ms / 1e3) / 1e12})
    if rank == 0:
        peak_tf, _, peak_src = peaks()
        r50 = next(r for r in table if r["ddim_steps"] == 50)
        line = {"metric": "DDIM-50 patches/sec", "value": r50["patches_per_s"], "unit": "patches/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": r50["ms_per_chain"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f16",
                "data": "synthetic", "config": dict(config_block(args, world), workload="DDIM_Sweep (BASELINE config 4): "
                "v-prediction UNetSmall(8,4,96), grid B from 999 with 10/25/50/100/250 steps, eta=0"),
                "sweep": table, "gpu_launches": launches, "clocks": clocks_all[2],
                "roofline": {"bound": "tensor", "achieved": r50["tflops_per_gpu"], "peak": peak_tf, "unit": "TFLOP/s",
                             "frac": r50["tflops_per_gpu"] / peak_tf, "peak_source": peak_src, "traffic": None},
                "e2e": None, "cpu_baseline": None}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_latency(args, rank, world, dev):
    """The reference scripts' operating point (batch 1) as its own line: `value` = fused DDIM-50 patches/s at batch 1."""
    import torch
    from s1s2_b200 import schedule
    if rank != 0:
        return
    model = build_model(SEED_V, 16, dev)
    _, _, abar = schedule.derive(schedule.cosine_beta_schedule(1000))
    clocks = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    l0 = model.launch_count()
    rows = latency_table(model, abar, dev)
    clk = clocks.stop()
    lib = None if args.no_library_baseline else gpu_library_baseline(dev)
    peak_tf, _, peak_src = peaks()
    b1 = rows[0]["fused"]
    layers = None
    if not args.no_layers:
        fl = dict(layer_flops())
        layers = {}
        for B in (1, 4):
            lt = model.profile_layers(dev, H, W, B, reps=20)
            layers[f"b{B}"] = [{"layer": n, "ms": round(t, 4), "tflops": round(fl[n] * B / (t / 1e3) / 1e12, 1)} for n, t in lt]
    line = {"metric": "DDIM-50 patches/sec", "value": b1["patches_per_s"], "unit": "patches/s", "n_gpus": 1, "steps": b1["chains_timed"],
            "warmup": 2, "ms_per_step": b1["ms_per_chain"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16", "data": "synthetic",
            "config": {"workload": "latency: DDIM_Multi-step_v_Prediction at the scripts' own batch 1 (and 2/4/8/16), v-DDIM-50, one "
                                   "256x256 patch per chain; fused s1s2_sample loop and the INTEGRATION.md 2-line drop-in loop",
                       "batch_per_gpu": 1, "ddim_steps": 50, "l2": "at batch 1 the 123 MB activation arena is L2-sized: warm-cache "
                       "numbers, as in the reference's own back-to-back steps"},
            "latency": rows, "gpu_library_baseline": lib, "layers": layers, "gpu_launches": int(model.launch_count() - l0), "clocks": clk,
            "roofline": {"bound": "tensor", "achieved": b1["tflops"], "peak": peak_tf, "unit": "TFLOP/s", "frac": b1["tflops"] / peak_tf,
                         "peak_source": peak_src, "traffic": None}, "e2e": None, "cpu_baseline": None}
    print(json.dumps(line), flush=True)


def config_block(args, world):
    B = args.batch
    return {"workload": ("DDIM_Multi-step / Evaluation_Pure_Generation: eps-prediction UNetSmall(8,4,96), grid A 999->0 (50 calls)"
                         if args.workload == "eps16" else
                         "DDIM_Multi-step_v_Prediction: v-prediction UNetSmall(8,4,96), grid B 0..999 (50 calls), eta=0"),
            "patch": "4x256x256 cond + 4x256x256 noise (Patch.py shape)", "batch_per_gpu": B, "global_batch": B * world,
            "ddim_steps": 50, "weights": "random init (Models/*.pth absent from the reference tree)",
            "arithmetic": "f16 operands, f32 accumulate (TMEM), f32 bias / epilogue / scheduler state (x_t carried as an f16 hi/lo pair)",
            "parallelism": f"patch-sharded x{world}, no data-path collective",
            "l2": "working set per step (activation arena ~123 MB/patch) exceeds the 126 MB L2; no flush needed"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="v64", choices=["v64", "eps16", "scene", "sweep", "latency"])
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-layers", action="store_true")
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--no-library-baseline", action="store_true")
    ap.add_argument("--no-scene", action="store_true")
    ap.add_argument("--no-configs", action="store_true")
    ap.add_argument("--quick", action="store_true", help="main arm only (no cpu / library baselines, latency table, scene block)")
    args = ap.parse_args()
    if args.quick:
        args.no_cpu_baseline = args.no_latency = args.no_library_baseline = args.no_scene = args.no_configs = True
    if args.batch is None:
        args.batch = 16 if args.workload == "eps16" else 64
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torch.distributed.run, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 2000)] + sys.argv
        sys.exit(subprocess.call(cmd))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from s1s2_b200 import _build
    if rank == 0:
        _build.build()                              # no-op when the in-tree .so is current (stdout stays one JSON line)
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: the product path has no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()

    from s1s2_b200 import samplers, schedule

    if args.workload == "scene":
        run_scene(args, rank, world, dev)
        return
    if args.workload == "sweep":
        run_sweep(args, rank, world, dev)
        return
    if args.workload == "latency":
        run_latency(args, rank, world, dev)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    B = args.batch
    model = build_model(SEED_EPS if args.workload == "eps16" else SEED_V, B, dev)
    _, _, abar = schedule.derive(schedule.cosine_beta_schedule(1000))
    steps, init_scale = make_steps(args.workload, abar)
    assert len(steps) == N_CALLS
    cond_h, noise_h = synthetic_batch(B, 2024 + rank)
    cond_h, noise_h = cond_h.pin_memory(), noise_h.pin_memory()
    cond_d, noise_d = cond_h.to(dev), noise_h.to(dev)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------------------------------------------------------- device-resident arm
    for _ in range(args.warmup):
        out = samplers.run_steps(model, steps, cond_d, noise_d, init_scale=init_scale)
    sync()
    clocks = ClockSampler(local)
    l0 = model.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = samplers.run_steps(model, steps, cond_d, noise_d, init_scale=init_scale)
    e1.record()
    sync()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = model.launch_count() - l0
    clk = clocks.stop()
    assert bool(torch.isfinite(out).all())
    patches = B * world * args.steps
    value = patches / (ms / 1e3)

    # ---------------------------------------------------------------- end-to-end arm (host buffers, pipelined entry)
    # one library call per timed region: all `steps` batches go through s1s2_sample_host_stream, which uploads batch i+1 and
    # downloads batch i-1 under the model calls of batch i; every byte in and out is a real pinned-host <-> device copy.
    n_e2e = B * args.steps
    cond_all = cond_h.repeat(args.steps, 1, 1, 1).pin_memory()
    noise_all = noise_h.repeat(args.steps, 1, 1, 1).pin_memory()
    samplers.run_steps_host(model, steps, cond_h, noise_h, init_scale=init_scale, device=dev, batch=B)
    samplers.host_result_buffer(model, noise_all.shape)          # pinned result buffer exists before the clock starts
    sync()
    e0.record()
    res = samplers.run_steps_host(model, steps, cond_all, noise_all, init_scale=init_scale, device=dev, batch=B)
    e1.record()
    sync()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    img_bytes = B * 4 * H * W * 4
    e2e = {"value": patches / (ms_e2e / 1e3), "unit": "patches/s", "h2d_bytes_per_step": 2 * img_bytes * world,
           "d2h_bytes_per_step": img_bytes * world,
           "entry": "s1s2_sample_host_stream (pinned host cond + noise -> host image; copies of neighbouring batches overlap the model calls)"}
    assert res.shape[0] == n_e2e
    assert bool(torch.equal(res[-B:], out.cpu())), "host-buffer entry disagrees with the device-resident entry"
    del cond_all, noise_all

    # ---------------------------------------------------------------- roofline of the conv kernel family
    peak_tf, _, peak_src = peaks()
    flop_step = FLOP_PER_CALL * N_CALLS * B                      # per GPU per step
    achieved = flop_step * args.steps / (ms / 1e3) / 1e12       # TFLOP/s per GPU over the timed region
    roof = {"bound": "tensor", "kernel": "tcgen05 implicit-GEMM conv family: conv_umma_kernel<...> (14 launches per model call) + "
            "conv_px_kernel<...> (conv1.0, conv1.2+outc+scheduler); timed region = 50 model calls x steps, nothing else launches",
            "achieved": achieved, "peak": peak_tf,
            "unit": "TFLOP/s", "frac": achieved / peak_tf, "peak_source": peak_src, "traffic": None,
            "flop_per_launch_avg": FLOP_PER_CALL * B / 16}
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp) and B == 64:           # ncu --set full capture of one model call at batch 64 (profiles/)
        tj = json.load(open(tp))
        fp = csrc_fingerprint()
        if tj.get("csrc_fingerprint") == fp:      # quoted only while the profiled kernels are the ones being timed
            roof["traffic"] = tj["dram_total_GB"] * 1e9 / 16                     # DRAM bytes per launch (mean of the 16)
            roof["traffic_detail"] = {k: tj.get(k) for k in ("source", "per", "dram_read_GB", "dram_write_GB",
                                                             "algorithmic_activation_GB", "git_head", "csrc_fingerprint")}
        else:
            roof["traffic_stale"] = (f"profiles/traffic.json was captured for kernel sources {tj.get('csrc_fingerprint')}, "
                                     f"the tree is at {fp}: re-run tools/gpu_round.sh")
    if not args.no_layers and rank == 0:
        lt = model.profile_layers(dev, H, W, B, reps=3)
        fl = dict(layer_flops())
        roof["layers"] = [{"layer": n, "ms": round(t, 4), "tflops": round(fl[n] * B / (t / 1e3) / 1e12, 1)} for n, t in lt]
        top = max(roof["layers"], key=lambda r: r["ms"])
        roof["dominant_launch"] = top

    lat = lib = other = None
    if rank == 0 and world == 1:
        if not args.no_configs and args.workload == "v64":
            other = other_configs_block(model, abar, dev)
        if not args.no_latency:
            lat = latency_table(model, abar, dev)
        if not args.no_library_baseline:
            lib = gpu_library_baseline(dev)
    scene = None
    if not args.no_scene and args.workload == "v64":
        scene = scene_block(model, abar, dev, rank, world, B)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cpu_oracle_sample(args.workload, 1, threads)
        dt = cpu_oracle_sample(args.workload, N_CALLS, threads)
        cpu = {"value": 1.0 / dt, "unit": "patches/s", "cores": threads, "kind": "port",
               "sample": f"one complete DDIM-50 chain (all {N_CALLS} model calls + scheduler updates) of one 256x256 patch "
                         f"({dt:.1f} s), oracle/ fp32 PyTorch CPU, {threads} threads"}

    if rank == 0:
        line = {"metric": "DDIM-50 patches/sec", "value": value, "unit": "patches/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f16",
                "data": "synthetic", "config": config_block(args, world), "e2e": e2e, "gpu_launches": int(launches),
                "clocks": clk, "roofline": roof, "cpu_baseline": cpu, "gpu_library_baseline": lib, "latency": lat, "scene": scene,
                "other_configs": other}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
