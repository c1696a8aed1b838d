"""GPU bring-up diagnostic (run under gpurun): per-layer isolated errors, then a quick timing.
usage: python tools/diag.py layers B H W | time B steps"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "s1-to-s2_super-resolution_project-code_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from oracle import schedule as osched, unet as ounet  # noqa: E402
import s1s2_b200  # noqa: E402
from s1s2_b200 import samplers, schedule  # noqa: E402


BASE_CH = int(os.environ.get("S1S2_BASE_CH", "96"))      # 64: the reference's class default


def make(max_batch):
    sd = ounet.init_state_dict(8, 4, BASE_CH, seed=1234)
    m = s1s2_b200.UNetSmallB200(8, 4, BASE_CH, max_batch=max_batch).to("cuda")
    m.load_state_dict(sd)
    return sd, m.eval()


def main():
    mode = sys.argv[1]
    if mode == "layers":
        from layer_ref import check_layers
        B, H, W = (int(v) for v in sys.argv[2:5])
        sd, m = make(B)
        g = torch.Generator().manual_seed(1)
        x = torch.randn((B, 8, H, W), generator=g)
        t = torch.tensor([999, 20, 501, 0][:B], dtype=torch.long)
        y = m(x.cuda(), t.cuda())
        torch.cuda.synchronize()
        print(f"forward ok B={B} {H}x{W}; isolated per-layer errors:", flush=True)
        check_layers(m, sd, y, B, verbose=True)
        ref = ounet.OracleModel(sd)(x, t)
        print("end-to-end rel-L2 vs fp32 oracle:", float((y.cpu() - ref).norm() / ref.norm()), flush=True)
    elif mode == "time":
        B, nsteps = int(sys.argv[2]), int(sys.argv[3])
        sd, m = make(B)
        _, _, ab = osched.make_schedule(1000)
        steps = schedule.steps_grid_b(ab, schedule.grid_b(999, nsteps), "v")
        cond = torch.randn((B, 4, 256, 256), device="cuda")
        x = torch.randn((B, 4, 256, 256), device="cuda")
        samplers.run_steps(m, steps[:2], cond, x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        samplers.run_steps(m, steps, cond, x)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        # 2*MAC per patch per call: 301.851 GFLOP at base_ch 96; the 3x3 / transposed convs scale with base_ch^2
        fl = B * len(steps) * (301.851e9 if BASE_CH == 96 else 134.394e9)
        print(f"base_ch={BASE_CH} B={B} steps={len(steps)}: {ms:.1f} ms, {ms/len(steps):.2f} ms/step, {fl/ms/1e9:.1f} TFLOP/s, "
              f"{B/(ms/1e3)*len(steps)/50:.2f} DDIM-50-equivalent patches/s", flush=True)


def loop_mode():
    """python tools/diag.py loop B layer_idx[,layer_idx..] reps : each layer alone, perf modes 0..3, with clocks."""
    sys.path.insert(0, ROOT)
    import bench
    B, reps = int(sys.argv[2]), int(sys.argv[4])
    layers = [int(v) for v in sys.argv[3].split(",")]
    modes = [int(v) for v in sys.argv[5].split(",")] if len(sys.argv) > 5 else [0, 1, 2, 3]
    sd, m = make(B)
    fl = bench.layer_flops()
    m(torch.randn((B, 8, 256, 256), device="cuda"), torch.zeros(B, dtype=torch.long, device="cuda"))
    torch.cuda.synchronize()
    for li in layers:
        for mode in modes:
            cs = bench.ClockSampler(0)
            ms = m.loop_layer("cuda", 256, 256, B, li, reps, mode)
            clk = cs.stop()
            tf = fl[li][1] * B / (ms / 1e3) / 1e12
            mhz = clk.get("sm_mhz") or 0
            util = tf * 1e12 / (148 * 8192 * mhz * 1e6) if mhz else 0
            print(f"layer {li:2d} {fl[li][0]:10s} mode {mode}: {ms:7.3f} ms  {tf:7.1f} TFLOP/s  sm {mhz:6.0f} MHz  "
                  f"util@clk {util:5.3f}  power {clk.get('power_w_max')} reasons {clk.get('reasons')}", flush=True)


def overhead_mode():
    """python tools/diag.py overhead : per-model-call time of the fused chain at batch 1 for shrinking patch sizes -- at 16 x 16
    the arithmetic is negligible, so the time per call / 16 is the fixed cost of one launch in the chain (prologue, first
    loads, tail, hand-over to the next kernel)."""
    sd, m = make(1)
    _, _, ab = osched.make_schedule(1000)
    steps = schedule.steps_grid_b(ab, schedule.grid_b(999, 50), "v")
    for hw in (16, 32, 64, 128, 256):
        cond = torch.randn((1, 4, hw, hw), device="cuda")
        x = torch.randn((1, 4, hw, hw), device="cuda")
        for _ in range(3):
            samplers.run_steps(m, steps, cond, x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            samplers.run_steps(m, steps, cond, x)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 10 / len(steps) * 1e3
        print(f"{hw:3d} x {hw:3d}, batch 1: {us:7.1f} us per model call = {us / 16:5.2f} us per launch", flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "overhead":
        overhead_mode()
        sys.exit(0)
    if sys.argv[1] == "loop":
        loop_mode()
        sys.exit(0)
    main()
