"""The library path the reference itself would take on this GPU: PyTorch eager + cuDNN (run under gpurun).

SURVEY.md section 2a / 8d: the reference ships no kernels; on CUDA its sampler is eager PyTorch dispatching to cuDNN
(conv / conv-transpose) and ATen elementwise kernels -- fp32 with TF32 allowed (cuDNN default) in the batch-1 scripts,
fp16 autocast in Limitation_Test*.py.  The reference tree does not travel to the GPU box, so this times oracle/'s
restatement of the same module and v-DDIM loop (pinned against the reference's outputs, tests/golden/) on `cuda`
through exactly those library calls.  MEASUREMENT ONLY: nothing here is on the product path.

Reports DDIM-50 patches/s (v sampler, grid B 0..999, eta=0) for
  tf32_b1     fp32 tensors, torch.backends.cudnn.allow_tf32 = True (PyTorch default), batch 1   (the reference scripts as written)
  fp32_b1     the same with TF32 off (strict fp32)
  fp16_b16/64 torch.autocast(float16), batch 16 / 64                                  (Limitation_Test_v_Prediction.py:205-207)
  fp16_cl_b64 the same with channels_last tensors (best case for cuDNN's tensor-core engines)

usage: python tools/cudnn_baseline.py [out.json]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import samplers as osamplers, schedule as osched, unet as ounet  # noqa: E402

FLOP_PER_CALL = 301_851_475_968


def run(mode, B, n_calls, sd, abar):
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(2024)
    cond = torch.randn((B, 4, 256, 256), generator=g).to(dev)
    noise = torch.randn((B, 4, 256, 256), generator=g).to(dev)
    torch.backends.cudnn.allow_tf32 = mode != "fp32"
    torch.backends.cuda.matmul.allow_tf32 = mode != "fp32"
    torch.backends.cudnn.benchmark = True
    sdd = {k: v.to(dev) for k, v in sd.items()}
    if mode == "fp16_cl":
        sdd = {k: (v.contiguous(memory_format=torch.channels_last) if v.ndim == 4 else v) for k, v in sdd.items()}
        cond = cond.contiguous(memory_format=torch.channels_last)
        noise = noise.contiguous(memory_format=torch.channels_last)
    base = ounet.OracleModel(sdd)
    calls = {"n": 0}

    class Stop(Exception):
        pass

    def model(x, t):
        if calls["n"] >= n_calls:
            raise Stop()
        calls["n"] += 1
        t = t.to(x.device)
        if mode.startswith("fp16"):
            with torch.autocast("cuda", dtype=torch.float16):
                return base(x, t).float()
        return base(x, t)
    model.outc = base.outc

    def chain():
        calls["n"] = 0
        try:
            osamplers.ddim_v_grid_b(model, cond, abar.to(dev), noise, 50)
        except Stop:
            pass
    chain()                                   # cuDNN autotune + warm-up
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    chain()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ms_call = ms / n_calls
    return {"mode": mode, "batch": B, "model_calls_timed": n_calls, "ms_per_model_call": round(ms_call, 3),
            "patches_per_s_ddim50": round(B / (ms_call * 50 / 1e3), 3),
            "tflops": round(FLOP_PER_CALL * B / (ms_call / 1e3) / 1e12, 1)}


def main():
    sd = ounet.init_state_dict(8, 4, 96, seed=1235)
    _, _, abar = osched.make_schedule(1000)
    rows = []
    for mode, B, n in (("tf32", 1, 50), ("fp32", 1, 20), ("fp16", 16, 50), ("fp16", 64, 50), ("fp16_cl", 64, 50)):
        t0 = time.time()
        try:
            r = run(mode, B, n, sd, abar)
        except RuntimeError as e:             # e.g. out of memory at batch 64 in fp32
            r = {"mode": mode, "batch": B, "error": str(e)[:200]}
        r["wall_s"] = round(time.time() - t0, 1)
        rows.append(r)
        print(r, flush=True)
        torch.cuda.empty_cache()
    out = {"what": "oracle/ restatement of the reference's v-DDIM sampler run on cuda through PyTorch eager + cuDNN "
                   f"(torch {torch.__version__}, cuDNN {torch.backends.cudnn.version()}); DDIM-50-equivalent patches/s",
           "gpu": torch.cuda.get_device_name(0), "rows": rows}
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
