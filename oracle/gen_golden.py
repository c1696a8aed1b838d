#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE's own code (imported by path from /root/reference).

Run in the build container only (the GPU box has no /root/reference):

    python oracle/gen_golden.py [--ref /root/reference] [--out tests/golden]

The reference has no tests or fixtures for this path, so these vectors -- outputs of the unmodified reference
functions on seeded inputs -- are what pins the oracle (tests/test_oracle_golden.py) and, through it, the CUDA
path.  Everything is kept tiny (base_ch=4, 16x16 / 32x32 patches, T<=1000) so the fixtures stay small.

RNG handling: the reference draws its noise from torch's global CPU generator.  Before each reference call we
``torch.manual_seed(s)``; afterwards we replay the same seed and the same sequence of ``torch.randn`` shapes to
recover the tensors the reference consumed, and store them next to the outputs.
"""
import argparse
import hashlib
import importlib.util
import os
import sys
import types

import numpy as np
import torch


def load_ref(ref_root, rel, name):
    path = os.path.join(ref_root, rel)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def sd_to_np(sd, prefix="sd/"):
    return {prefix + k: v.detach().numpy() for k, v in sd.items()}


def gen_metrics_agg(lt, out_dir):
    """Dataset-level, pixel-weighted aggregation of Limitation_Test.py:118-159: channelwise_error_sums over three
    batches (masked, unmasked, one empty mask row) accumulated the way run_eval does (:330-334), then aggregate_final
    with equal and with explicit band weights."""
    g = torch.Generator().manual_seed(31)
    out = {}
    abs_tot, sq_tot, w_tot = torch.zeros(4), torch.zeros(4), torch.zeros(())
    for b, (B, masked) in enumerate(((3, True), (2, False), (2, True))):
        pred, tgt = torch.rand(B, 4, 24, 20, generator=g), torch.rand(B, 4, 24, 20, generator=g)
        mask = (torch.rand(B, 24, 20, generator=g) > 0.25).float() if masked else None
        if b == 2:
            mask[1] = 0.0                                  # a patch without a single valid pixel
        a, q, w = lt.channelwise_error_sums(pred, tgt, mask)
        out[f"agg/pred{b}"], out[f"agg/tgt{b}"] = pred.numpy(), tgt.numpy()
        if mask is not None:
            out[f"agg/mask{b}"] = mask.numpy()
        out[f"agg/abs{b}"], out[f"agg/sq{b}"], out[f"agg/w{b}"] = a.numpy(), q.numpy(), w.numpy()
        abs_tot += a; sq_tot += q; w_tot += w
    for tag, bw in (("eq", None), ("bw", [1.0, 1.0, 1.0, 2.0])):
        mae, mse, ps, mae_c, mse_c, ps_c = lt.aggregate_final(abs_tot, sq_tot, w_tot, bw)
        out[f"agg/final_{tag}"] = np.array([mae, mse, ps], np.float64)
        out[f"agg/final_{tag}_c"] = np.stack([mae_c, mse_c, ps_c]).astype(np.float64)
    np.savez_compressed(os.path.join(out_dir, "metrics_agg.npz"), **out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "..", "tests", "golden"))
    ap.add_argument("--only", default=None, choices=[None, "metrics_agg"], help="write just this fixture file")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    torch.set_num_threads(1)           # one summation order, reproducible fixtures
    torch.use_deterministic_algorithms(True)
    if args.only == "metrics_agg":
        gen_metrics_agg(load_ref(args.ref, "Evaluation/Limitation_Test.py", "ref_lt"), args.out)
        print("metrics_agg.npz", os.path.getsize(os.path.join(args.out, "metrics_agg.npz")))
        return

    if "rasterio" not in sys.modules:  # Patch.py imports rasterio at module scope; only raster I/O uses it
        sys.modules["rasterio"] = types.ModuleType("rasterio")

    ms = load_ref(args.ref, "Evaluation/DDIM_Multi-step.py", "ref_ms")
    msv = load_ref(args.ref, "Evaluation/DDIM_Multi-step_v_Prediction.py", "ref_msv")
    lt = load_ref(args.ref, "Evaluation/Limitation_Test.py", "ref_lt")
    ltv = load_ref(args.ref, "Evaluation/Limitation_Test_v_Prediction.py", "ref_ltv")
    pg = load_ref(args.ref, "Evaluation_Updated/Evaluation_Pure_Generation.py", "ref_pg")
    pt = load_ref(args.ref, "Patch.py", "ref_patch")

    # ---------------------------------------------------------------- schedule + grids
    out = {}
    betas = ms.cosine_beta_schedule(1000)
    alphas = 1.0 - betas
    abar = torch.cumprod(alphas, dim=0)
    out["cosine_betas"] = betas.numpy()
    out["cosine_alpha_bar"] = abar.numpy()
    out["cosine_alpha_bar_sha256"] = np.frombuffer(
        hashlib.sha256(abar.numpy().astype("<f4").tobytes()).hexdigest().encode(), dtype=np.uint8)
    lbetas = lt.make_schedule(1000, "linear")
    out["linear_betas"] = lbetas.numpy()
    out["linear_alpha_bar"] = torch.cumprod(1.0 - lbetas, dim=0).numpy()
    b16 = ms.cosine_beta_schedule(16)
    out["cosine16_betas"] = b16.numpy()
    for (ts, st) in [(999, 50), (200, 20), (999, 10), (999, 25), (999, 100), (999, 250), (300, 100), (150, 10)]:
        out[f"gridA_{ts}_{st}"] = torch.linspace(ts, 0, st + 1, dtype=torch.long).numpy()
    for (K, st) in [(999, 50), (200, 20), (999, 10), (999, 25), (999, 100), (999, 250), (30, 100), (999, 2)]:
        # the v scripts' construction (DDIM_Multi-step_v_Prediction.py:147-151), verbatim call sequence
        g = torch.linspace(0, K, st)
        idxs = torch.unique(torch.round(g).to(torch.long), sorted=True)
        if idxs[-1].item() != K:
            idxs = torch.unique(torch.cat([idxs, torch.tensor([K])]), sorted=True)
        out[f"gridB_{K}_{st}"] = idxs.numpy()
    np.savez_compressed(os.path.join(args.out, "schedule.npz"), **out)

    # ---------------------------------------------------------------- UNet forward
    out = {}
    for tag, (bc, B, H, seed) in {"bc4": (4, 2, 32, 11), "bc8": (8, 1, 16, 12)}.items():
        torch.manual_seed(seed)
        net = ms.UNetSmall(in_ch=8, out_ch=4, base_ch=bc).eval()
        x = torch.randn(B, 8, H, H)
        t = torch.tensor([999, 20][:B], dtype=torch.long)
        acts = {}
        hooks = [m.register_forward_hook(lambda mod, i, o, n=n: acts.__setitem__(n, o.detach()))
                 for n, m in net.named_children()]
        with torch.no_grad():
            y = net(x, t)
        for h in hooks:
            h.remove()
        out.update(sd_to_np(net.state_dict(), f"{tag}/sd/"))
        out[f"{tag}/x"] = x.numpy(); out[f"{tag}/t"] = t.numpy(); out[f"{tag}/y"] = y.numpy()
        for n, a in acts.items():
            out[f"{tag}/act/{n}"] = a.numpy()
    np.savez_compressed(os.path.join(args.out, "unet.npz"), **out)

    # ---------------------------------------------------------------- samplers
    out = {}
    torch.manual_seed(21)
    net = ms.UNetSmall(in_ch=8, out_ch=4, base_ch=4).eval()
    # scale the head down so trajectories stay O(1) with random weights (still the reference's code path)
    out.update(sd_to_np(net.state_dict(), "sd/"))
    H = 16
    g = torch.Generator().manual_seed(5)
    cond1 = torch.randn(1, 4, H, H, generator=g)
    cond2 = torch.randn(2, 4, H, H, generator=g)
    x_gt = torch.rand(1, 4, H, H, generator=g)
    out["cond1"], out["cond2"], out["x_gt"] = cond1.numpy(), cond2.numpy(), x_gt.numpy()
    betas = ms.cosine_beta_schedule(1000); alphas = 1 - betas; abar = torch.cumprod(alphas, 0)

    def replay(seed, shapes):
        torch.manual_seed(seed)
        return [torch.randn(*s) for s in shapes]

    # (1) eps DDIM from noise, grid A  -- Evaluation_Pure_Generation.ddpm_ddim_generate
    for (ts, st, seed) in [(999, 10, 101), (200, 20, 102)]:
        torch.manual_seed(seed)
        y = pg.ddpm_ddim_generate(net, cond1, abar, t_start=ts, steps=st)
        (z,) = replay(seed, [(1, 4, H, H)])
        out[f"pg_generate_{ts}_{st}/noise"] = z.numpy(); out[f"pg_generate_{ts}_{st}/y"] = y.numpy()
    # (2) eps DDIM from noised GT, grid A -- DDIM_Multi-step.ddim_multistep_eval
    torch.manual_seed(103)
    mae, mse, y = ms.ddim_multistep_eval(net, x_gt, cond1, abar, None, t_start=200, steps=20)
    (z,) = replay(103, [(1, 4, H, H)])
    out["ms_eval_200_20/noise"] = z.numpy(); out["ms_eval_200_20/y"] = y.numpy()
    out["ms_eval_200_20/mae_mse"] = np.array([mae, mse], np.float64)
    # (3) v DDIM grid B, eta = 0 -- DDIM_Multi-step_v_Prediction.ddim_multistep_eval_v
    for (ts, st, seed) in [(999, 10, 104), (200, 20, 105)]:
        torch.manual_seed(seed)
        mae, mse, y = msv.ddim_multistep_eval_v(net, x_gt, cond1, abar, None, t_start=ts, steps=st, eta=0.0)
        (z,) = replay(seed, [(1, 4, H, H)])
        out[f"msv_eval_{ts}_{st}/noise"] = z.numpy(); out[f"msv_eval_{ts}_{st}/y"] = y.numpy()
    # (4) batched eps DDIM grid B -- Limitation_Test.ddim_sample
    torch.manual_seed(106)
    y = lt.ddim_sample(net, cond2, alphas, abar, 4, steps=12)
    (z,) = replay(106, [(2, 4, H, H)])
    out["lt_ddim_12/noise"] = z.numpy(); out["lt_ddim_12/y"] = y.numpy()
    # (5) batched v DDIM grid B, eta = 0 and eta = 0.05 -- Limitation_Test_v_Prediction.sample_ddim_v
    for eta, seed in [(0.0, 107), (0.05, 108)]:
        torch.manual_seed(seed)
        y = ltv.sample_ddim_v(net, cond2, abar, 4, steps=12, eta=eta, t_start=None)
        zs = replay(seed, [(2, 4, H, H)] * 12)     # init, then one draw per loop iteration with i > 0
        tag = f"ltv_ddim_12_eta{eta}"
        out[f"{tag}/noise"] = zs[0].numpy(); out[f"{tag}/y"] = y.numpy()
        out[f"{tag}/step_noise"] = torch.stack(zs[1:]).numpy()   # in loop order (i = n-1 .. 1)
    # (6) DDPM ancestral on a 16-step schedule -- Limitation_Test.ddpm_sample / sample_ddpm_v
    b16 = ms.cosine_beta_schedule(16); a16 = 1 - b16; ab16 = torch.cumprod(a16, 0)
    for tag, fn, seed in [("lt_ddpm16", lt.ddpm_sample, 109), ("ltv_ddpm16", ltv.sample_ddpm_v, 110)]:
        torch.manual_seed(seed)
        y = fn(net, cond2, b16, a16, ab16, 4)
        zs = replay(seed, [(2, 4, H, H)] * 16)     # init, then z for t = 15 .. 1
        out[f"{tag}/noise"] = zs[0].numpy(); out[f"{tag}/y"] = y.numpy()
        out[f"{tag}/step_noise"] = torch.stack(zs[1:]).numpy()
    # (7) partial DDIM from GT -- Limitation_Test.partial_ddim_from_gt
    torch.manual_seed(111)
    y = lt.partial_ddim_from_gt(net, x_gt, cond1, abar, 6)
    (z,) = replay(111, [(1, 4, H, H)])
    out["lt_partial_6/noise"] = z.numpy(); out["lt_partial_6/y"] = y.numpy()
    # (8) one-step eps recon -- DDIM_Multi-step.one_step_recon (same arithmetic as Onestep.py:149-160)
    mae, mse, y = ms.one_step_recon(net, x_gt, cond1, abar, None, 20, rng_seed=112)
    (z,) = replay(112, [(1, 4, H, H)])
    out["ms_onestep_20/noise"] = z.numpy(); out["ms_onestep_20/y"] = y.numpy()
    # (9) v <-> x0, eps conversion -- DDIM_Multi-step_v_Prediction.v_to_x0_eps
    xv, vv = torch.randn(2, 4, H, H, generator=g), torch.randn(2, 4, H, H, generator=g)
    x0c, epc = msv.v_to_x0_eps(xv, vv, abar[torch.tensor([500, 20])])
    out["v2x0/x"], out["v2x0/v"], out["v2x0/x0"], out["v2x0/eps"] = xv.numpy(), vv.numpy(), x0c.numpy(), epc.numpy()
    # (10) metrics -- DDIM_Multi-step / Evaluation_Pure_Generation
    p = torch.rand(1, 4, H, H, generator=g); mk = (torch.rand(1, H, H, generator=g) > 0.2).float()
    out["metrics/pred"], out["metrics/mask"] = p.numpy(), mk.numpy()
    out["metrics/vals"] = np.array([pg.masked_mae(p, x_gt, mk), pg.masked_mse(p, x_gt, mk), pg.psnr(p, x_gt, mk),
                                    pg.ssim_simple(p, x_gt), pg.sam(p, x_gt, mk), pg.ergas(p, x_gt, mk)], np.float64)
    np.savez_compressed(os.path.join(args.out, "samplers.npz"), **out)

    # ---------------------------------------------------------------- Patch.py tiling + normalisation
    out = {}
    for (Hh, Ww, ps, st) in [(2048, 2048, 256, 64), (2048, 2048, 256, 32), (2048, 2048, 256, 128),
                             (300, 420, 64, 32), (256, 256, 256, 32), (255, 400, 256, 32), (700, 513, 256, 100)]:
        lst = np.array(list(pt.patch_iter(Hh, Ww, ps, st)), dtype=np.int32).reshape(-1, 2)
        key = f"iter_{Hh}_{Ww}_{ps}_{st}"
        out[key + "/n"] = np.array([len(lst)], np.int64)
        out[key + "/sha256"] = np.frombuffer(hashlib.sha256(lst.astype("<i4").tobytes()).hexdigest().encode(), np.uint8)
        if len(lst) <= 200:
            out[key + "/list"] = lst
    rng = np.random.default_rng(7)
    Hh, Ww, ps, st = 96, 128, 32, 16
    scene = np.stack([rng.normal(-12, 4, (Hh, Ww)), rng.normal(-19, 4, (Hh, Ww)),
                      rng.uniform(20, 45, (Hh, Ww)), np.abs(rng.normal(300, 300, (Hh, Ww)))]).astype(np.float32)
    holes = rng.random((4, Hh, Ww)) < 0.02
    scene[holes] = np.nan
    scene[0, 40:44, 60:70] = np.inf
    scene[:, 0:32, 0:32] = np.nan                     # one fully invalid window
    scene[0, 64:96, 96:128] = -7.25                   # constant HH over a window -> sigma < 1e-6 branch
    target = np.zeros_like(scene)                     # finite everywhere: mask is decided by the inputs alone
    vmask = pt.build_mask(scene, target, None)
    conds, masks, idx = [], [], []
    import warnings
    for (r, c) in pt.patch_iter(Hh, Ww, ps, st):
        X = scene[:, r:r + ps, c:c + ps].copy(); M = vmask[r:r + ps, c:c + ps].copy()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            pt.zscore_inplace(X[0], M); pt.zscore_inplace(X[1], M)          # Patch.py:228-229
        X[2] = np.nan_to_num(X[2], nan=0.0) / 90.0                           # Patch.py:231
        X[3] = np.nan_to_num(X[3], nan=0.0) / 1000.0                         # Patch.py:232
        for ch in range(4):                                                  # Patch.py:236-239
            a = X[ch]; a[~M] = 0.0
            X[ch] = np.nan_to_num(a, nan=0.0, posinf=0.0, neginf=0.0).astype(np.float32)
        conds.append(X); masks.append(M.astype(np.uint8)); idx.append((r, c))
    out["norm/scene"] = scene; out["norm/vmask"] = vmask.astype(np.uint8)
    out["norm/cond"] = np.stack(conds); out["norm/mask"] = np.stack(masks)
    out["norm/origins"] = np.array(idx, np.int32); out["norm/ps_stride"] = np.array([ps, st], np.int32)
    np.savez_compressed(os.path.join(args.out, "patch.npz"), **out)

    # ---------------------------------------------------------------- Patch.py quality filters (Patch.py:88-114,205-224)
    out = {}
    rng = np.random.default_rng(11)
    Hh, Ww, ps, st = 96, 160, 32, 16
    yy, xx = np.mgrid[0:Hh, 0:Ww].astype(np.float32)
    base = 0.35 + 0.25 * np.sin(yy / 9.0) * np.cos(xx / 13.0)
    target = np.stack([np.clip(base * s_ + rng.normal(0, 0.03, (Hh, Ww)), 0, 1) for s_ in (0.8, 0.9, 1.0, 1.3)]).astype(np.float32)
    target[:, 0:48, 0:48] = 0.01 + rng.random((4, 48, 48)).astype(np.float32) * 0.08      # dark block (var > 1e-4)
    target[:, 48:96, 96:160] = 0.4                                                          # flat block (var < 1e-4, no texture)
    target[:, 0:48, 96:160] = (0.5 + 0.10 * np.sin(yy[0:48, 96:160] / 15.0)).astype(np.float32)     # smooth: varied but textureless
    target[rng.random((4, Hh, Ww)) < 0.01] = np.nan
    inputs = rng.normal(-12, 4, (4, Hh, Ww)).astype(np.float32)
    inputs[:, 60:96, 0:30] = np.nan                                                        # low valid ratio windows
    colloc = (rng.random((Hh, Ww)) > 0.03).astype(np.uint8)
    M_all = pt.build_mask(inputs, target, colloc)
    rows = []
    import warnings
    for (r, c) in pt.patch_iter(Hh, Ww, ps, st):
        Y = target[:, r:r + ps, c:c + ps].copy(); M = M_all[r:r + ps, c:c + ps].copy()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            vr = float(M.mean())
            var = [float(np.nanvar(Y[ch][M])) for ch in range(4)]
            dk = float(pt.dark_fraction(Y, M, thr=0.10))
            lv = float(pt.laplacian_var(Y[3], M))
        code = 0                                                                           # Patch.py:205-224, default thresholds
        if vr < 0.80: code = 1
        elif all(v < 1e-4 for v in var): code = 2
        elif dk > 0.60: code = 3
        elif lv < 5e-5: code = 4
        rows.append([r, c, code, vr] + var + [dk, lv])
    out["filt/inputs"] = inputs; out["filt/target"] = target; out["filt/colloc"] = colloc
    out["filt/mask"] = M_all.astype(np.uint8); out["filt/ps_stride"] = np.array([ps, st], np.int32)
    out["filt/rows"] = np.array(rows, np.float64)      # row, col, code, valid_ratio, var0..3, dark_fraction, laplacian_var
    np.savez_compressed(os.path.join(args.out, "filters.npz"), **out)
    gen_metrics_agg(lt, args.out)

    for f in sorted(os.listdir(args.out)):
        print(f, os.path.getsize(os.path.join(args.out, f)))


if __name__ == "__main__":
    main()
