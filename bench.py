#!/usr/bin/env python
"""DDIM-50 patches/sec of the S1->S2 sampling path on N B200s (one process per GPU), next to the CPU oracle.

A "step" is one complete DDIM-50 sampling (50 fused model calls + scheduler updates) of one batch of synthetic
256x256 patches per GPU.  Workloads (BASELINE.json configs):
  v64   v-prediction UNet, grid B (0..999, 50 entries), eta=0, batch 64 per GPU     [default; north-star target]
  eps16 eps-prediction UNet, grid A (999 -> 0, 50 calls), batch 16 per GPU
  sweep v-prediction UNet, grid B with 10 / 25 / 50 / 100 / 250 steps (BASELINE config 4, DDIM_Sweep): one JSON line with a
        `sweep` table of ms per model call and patches/s per step count (`value` = the 50-step row)
  scene one 4x2048x2048 scene tiled by Patch.py's rule, patch-sharded over the ranks, gathered and stitched (config 5)
Patches are independent units: N GPUs = N x batch patches per step, no data-path collective ("weak" scaling).

  value     patches/s with inputs resident in HBM (CUDA events, barrier + synchronize on both sides, max over ranks)
  e2e       the same through the host-buffer entry s1s2_sample_host (pinned host cond + noise in, image out)
  roofline  tensor-pipe roofline of the conv kernel family (every launch in the timed region is one instantiation of
            conv kernel family): algorithmic FLOPs (301 851 475 968 per patch per model call)
            / device time, against MEASURED_PEAKS.json's sustained bf16 figure; `layers` lists every launch of one
            model call timed with a CUDA event pair on the launching stream.
  cpu_baseline  oracle/ (a port of the reference's PyTorch sampler) timed on this box's host cores, bounded sample.

`--impl reference` times only the CPU oracle port (the reference is Python/PyTorch and does not travel to the GPU
box; oracle/ is its restatement, pinned against vectors the real reference produced: tests/golden/).
"""
import argparse
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
    if workload == "v64":
        steps = schedule.steps_grid_b(abar, schedule.grid_b(999, 50), "v")
        return steps, float(torch.sqrt(1 - abar[999]))
    steps = schedule.steps_eps_grid_a(abar, 999, 50)
    return steps, 1.0


def cpu_oracle_sample(workload, n_calls, threads):
    """Times `n_calls` model calls + scheduler updates of the DDIM-50 chain of ONE patch in the CPU oracle; returns
    (seconds, patches/s extrapolated to the 50-call chain)."""
    import torch
    from oracle import samplers as osamplers, schedule as osched, unet as ounet
    torch.set_num_threads(threads)
    sd = ounet.init_state_dict(8, 4, 96, seed=1234 if workload == "eps16" else 1235)
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
        if workload == "v64":
            osamplers.ddim_v_grid_b(counted, cond, abar, noise, 50)
        else:
            osamplers.ddim_eps_grid_a(counted, cond, abar, noise, 999, 50)
    except Stop:
        pass
    dt = time.perf_counter() - t0
    return dt, 1.0 / (dt * N_CALLS / calls["n"])


def run_reference(args, rank, world):
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
    pps = args.steps / (dt * N_CALLS / n_calls)
    sample = (f"each step = {n_calls} of the 50 model calls (+ scheduler updates) of one 256x256 patch, oracle/ fp32 "
              f"PyTorch CPU, {threads} threads; patches/s extrapolated x{N_CALLS // n_calls}")
    line = {"impl": "reference", "metric": "DDIM-50 patches/sec", "value": pps, "unit": "patches/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3 * N_CALLS / n_calls,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_block(args, world),
            "cpu_baseline": {"value": pps, "unit": "patches/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": pps, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_scene(args, rank, world, dev):
    """BASELINE config 5: one synthetic 4 x 2048 x 2048 Sentinel-1 scene, Patch.py tiling (256 / stride 64 = 841 patches)
    sharded patch-wise over the ranks, v-DDIM-50, NCCL gather, overlap-blend stitch on rank 0.  Strong scaling: a step
    is the whole scene; value = patches / max-over-ranks device time."""
    import torch
    import torch.distributed as dist
    import s1s2_b200
    from s1s2_b200 import scene as sc, schedule
    from oracle import unet as ounet
    sd = ounet.init_state_dict(8, 4, 96, seed=1235)
    model = s1s2_b200.UNetSmallB200(8, 4, 96, max_batch=args.batch).to(dev)
    model.load_state_dict(sd, strict=True)
    model.eval()
    _, _, abar = schedule.derive(schedule.cosine_beta_schedule(1000))
    scn = sc.synthetic_scene(2048, 2048, seed=0).to(dev)
    model.engine(dev, 256, 256, args.batch)

    def once():
        return sc.generate_scene(model, scn, abar, ps=256, stride=64, param="v", steps=50, t_start=999, batch=args.batch,
                                 rank=rank, world=world)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
    for _ in range(max(1, args.warmup // 3)):
        res = once()
    sync()
    clocks = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    l0 = model.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        res = once()
    e1.record()
    sync()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clk = clocks.stop()
    if rank == 0:
        n = int(res["kept"].sum())
        value = n * args.steps / (ms / 1e3)
        peak_tf, _, peak_src = peaks()
        achieved = value * N_CALLS * FLOP_PER_CALL / 1e12 / world
        line = {"metric": "DDIM-50 patches/sec", "value": value, "unit": "patches/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f16", "data": "synthetic",
                "config": {"workload": "Evaluation_Pure_Generation over a whole scene: 4x2048x2048 synthetic Sentinel-1 scene, Patch.py "
                           "tiling 256/stride 64 (841 patches), v-DDIM-50, patch-sharded, NCCL gather to rank 0, overlap-blend stitch",
                           "batch_per_gpu": args.batch, "patches": n, "parallelism": f"patch-sharded x{world}, one gather",
                           "l2": "working set exceeds L2"},
                "e2e": {"value": value, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                        "entry": "s1s2_b200.scene.generate_scene (scene resident on the device; tile extract -> sample -> gather -> stitch)"},
                "gpu_launches": int(model.launch_count() - l0), "clocks": clk,
                "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                             "peak_source": peak_src, "traffic": None,
                             "kernel": "conv kernel family, per GPU, over the whole scene step (extract/stitch/gather included in time)"},
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
    import s1s2_b200
    from s1s2_b200 import samplers, schedule
    from oracle import unet as ounet
    B = args.batch
    model = s1s2_b200.UNetSmallB200(8, 4, 96, max_batch=B).to(dev)
    model.load_state_dict(ounet.init_state_dict(8, 4, 96, seed=1235), strict=True)
    model.eval()
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


def config_block(args, world):
    B = args.batch
    return {"workload": ("DDIM_Multi-step_v_Prediction: v-prediction UNetSmall(8,4,96), grid B 0..999 (50 calls), eta=0"
                         if args.workload == "v64" else
                         "DDIM_Multi-step / Evaluation_Pure_Generation: eps-prediction UNetSmall(8,4,96), grid A 999->0 (50 calls)"),
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
    ap.add_argument("--workload", default="v64", choices=["v64", "eps16", "scene", "sweep"])
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-layers", action="store_true")
    args = ap.parse_args()
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

    import s1s2_b200
    from s1s2_b200 import samplers, schedule
    from oracle import unet as ounet               # weights only: the synthetic checkpoint both sides load

    if args.workload == "scene":
        run_scene(args, rank, world, dev)
        return
    if args.workload == "sweep":
        run_sweep(args, rank, world, dev)
        return

    B = args.batch
    sd = ounet.init_state_dict(8, 4, 96, seed=1234 if args.workload == "eps16" else 1235)
    model = s1s2_b200.UNetSmallB200(8, 4, 96, max_batch=B).to(dev)
    model.load_state_dict(sd, strict=True)
    model.eval()
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

    # ---------------------------------------------------------------- end-to-end arm (host buffers)
    samplers.run_steps_host(model, steps, cond_h, noise_h, init_scale=init_scale, device=dev)
    sync()
    e0.record()
    for _ in range(args.steps):
        res = samplers.run_steps_host(model, steps, cond_h, noise_h, init_scale=init_scale, device=dev)
    e1.record()
    sync()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    img_bytes = B * 4 * H * W * 4
    e2e = {"value": patches / (ms_e2e / 1e3), "unit": "patches/s", "h2d_bytes_per_step": 2 * img_bytes * world,
           "d2h_bytes_per_step": img_bytes * world, "entry": "s1s2_sample_host (pinned host cond + noise -> host image)"}
    assert bool(torch.equal(res, out.cpu())), "host-buffer entry disagrees with the device-resident entry"

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
        roof["traffic"] = tj["dram_total_GB"] * 1e9 / 16                     # DRAM bytes per launch (mean of the 16)
        roof["traffic_detail"] = {k: tj[k] for k in ("source", "per", "dram_read_GB", "dram_write_GB", "algorithmic_activation_GB")}
    if not args.no_layers and rank == 0:
        lt = model.profile_layers(dev, H, W, B, reps=3)
        fl = dict(layer_flops())
        roof["layers"] = [{"layer": n, "ms": round(t, 4), "tflops": round(fl[n] * B / (t / 1e3) / 1e12, 1)} for n, t in lt]
        top = max(roof["layers"], key=lambda r: r["ms"])
        roof["dominant_launch"] = top

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cpu_oracle_sample(args.workload, 1, threads)
        n_calls = N_CALLS
        dt, pps = cpu_oracle_sample(args.workload, n_calls, threads)
        cpu = {"value": pps, "unit": "patches/s", "cores": threads, "kind": "port",
               "sample": f"one complete DDIM-50 chain (all {n_calls} model calls + scheduler updates) of one 256x256 patch "
                         f"({dt:.1f} s), oracle/ fp32 PyTorch CPU, {threads} threads"}

    if rank == 0:
        line = {"metric": "DDIM-50 patches/sec", "value": value, "unit": "patches/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f16",
                "data": "synthetic", "config": config_block(args, world), "e2e": e2e, "gpu_launches": int(launches),
                "clocks": clk, "roofline": roof, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
