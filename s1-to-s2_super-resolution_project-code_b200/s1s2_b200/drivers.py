"""Drop-in command-line drivers with the reference scripts' flags and output files, running on the fused CUDA path.

  python -m s1s2_b200.drivers onestep      ...   Evaluation/Onestep.py:93-175 (and Onestep_v_Prediction.py with --param v)
  python -m s1s2_b200.drivers ddim         ...   Evaluation/DDIM_Multi-step.py --mode ddim (:173-232)
  python -m s1s2_b200.drivers ddim_v       ...   Evaluation/DDIM_Multi-step_v_Prediction.py --mode ddim
  python -m s1s2_b200.drivers ddim_sweep   ...   Evaluation/DDIM_Sweep.py --mode ddim_sweep (:387-416); --param v sweeps the
                                                 step counts of BASELINE.json's config 4 on the v model
  python -m s1s2_b200.drivers true_infer   ...   Evaluation_Updated/Evaluation_Pure_Generation.py --mode ddim --true_infer (:539-574)
  python -m s1s2_b200.drivers night_demo   ...   the same script's --mode night_demo (:720-727): generation without a target;
                                                 writes viz/NNN_night_pred.npy instead of the PNG panel
  python -m s1s2_b200.drivers limitation   ...   Evaluation/Limitation_Test.py run_eval (:273-400) and, with --param v,
                                                 Limitation_Test_v_Prediction.py run_eval (:258-372): batched DDPM / DDIM
                                                 sampling, dataset-level pixel-weighted MAE / MSE / PSNR, *_pred.npy / *_gt.npy
  python -m s1s2_b200.drivers scene        ...   whole-scene generation (Patch.py tiling + stitch, s1s2_b200.scene), under torchrun

Same flag names (--patch_dir --ckpt --out_dir --T --base_ch --max_files --t_start --ddim_steps --ddim_eta --t_small
--n_seeds --seed_base), same CSV / summary column sets.  PNG previews are not produced (visualisation is outside the
hot path).  Patches are read from Patch.py's `patch_*.npz` files (keys inputs / target / mask, Patch.py:249-255) and
processed `--batch` at a time; per-patch results do not depend on the batch size.
"""
import argparse
import csv
import os
import time

import numpy as np
import torch

from . import metrics, samplers, schedule
from .model import UNetSmallB200


def load_npz_as_tensors(path, device):
    """Evaluation/DDIM_Multi-step.py:104-111."""
    d = np.load(path)
    x_cond = torch.from_numpy(np.nan_to_num(d["inputs"].astype(np.float32))).unsqueeze(0).to(device)
    x_gt = torch.from_numpy(np.nan_to_num(d["target"].astype(np.float32))).unsqueeze(0).to(device)
    mask = torch.from_numpy(np.nan_to_num(d["mask"].astype(np.float32))).unsqueeze(0).to(device) if "mask" in d else None
    return x_cond, x_gt, mask, x_cond.size(1), x_gt.size(1)


def load_model(ckpt, in_ch, out_ch, base_ch, device, max_batch):
    """Evaluation/DDIM_Multi-step_v_Prediction.py:263-271 (optional {"model"| "state_dict"} wrapper, strict load)."""
    model = UNetSmallB200(in_ch=in_ch, out_ch=out_ch, base_ch=base_ch, max_batch=max_batch).to(device)
    state = torch.load(ckpt, map_location=device)
    if isinstance(state, dict) and "model" in state and isinstance(state["model"], dict):
        state = state["model"]
    elif isinstance(state, dict) and "state_dict" in state and isinstance(state["state_dict"], dict):
        state = state["state_dict"]
    model.load_state_dict(state, strict=True)
    return model.eval()


def _files(args):
    files = sorted(f for f in os.listdir(args.patch_dir) if f.endswith(".npz"))
    assert files, "No .npz found"
    return files[:args.max_files] if args.max_files > 0 else files


def _setup(args):
    os.makedirs(args.out_dir, exist_ok=True)
    device = torch.device("cuda")
    files = _files(args)
    print(f"[INFO] Evaluating {len(files)} files")
    x_cond0, x_gt0, _, Cc, Ct = load_npz_as_tensors(os.path.join(args.patch_dir, files[0]), device)
    model = load_model(args.ckpt, Cc + Ct, Ct, args.base_ch, device, args.batch)
    _, _, alpha_bar = schedule.derive(schedule.cosine_beta_schedule(args.T))
    return device, files, model, alpha_bar.to(device)


def _load_cpu(path):
    """The arrays of one patch file as CPU tensors (same conversions as load_npz_as_tensors)."""
    d = np.load(path)
    cond = torch.from_numpy(np.nan_to_num(d["inputs"].astype(np.float32))).unsqueeze(0)
    gt = torch.from_numpy(np.nan_to_num(d["target"].astype(np.float32))).unsqueeze(0)
    mask = torch.from_numpy(np.nan_to_num(d["mask"].astype(np.float32))).unsqueeze(0) if "mask" in d else None
    return cond, gt, mask


def _batches(args, files, device, workers=4):
    """Batches of `args.batch` patch files as device tensors.  The files of the NEXT batch are read and decompressed by a
    small thread pool (and staged in pinned memory) while the caller samples the current one, so file I/O -- ~5 ms per
    compressed 256x256 patch, a third of a DDIM-50 batch if done serially -- stays off the GPU's critical path."""
    from concurrent.futures import ThreadPoolExecutor
    pin = device.type == "cuda"

    def load(lo):
        names = files[lo:lo + args.batch]
        items = [_load_cpu(os.path.join(args.patch_dir, f)) for f in names]
        cond, gt = torch.cat([it[0] for it in items], 0), torch.cat([it[1] for it in items], 0)
        if pin:
            cond, gt = cond.pin_memory(), gt.pin_memory()
        return lo, names, cond, gt, [it[2] for it in items]

    starts = list(range(0, len(files), args.batch))
    with ThreadPoolExecutor(max_workers=1) as pool:          # one batch ahead; np.load / zlib release the GIL
        nxt = pool.submit(load, starts[0]) if starts else None
        for k in range(len(starts)):
            lo, names, cond, gt, mask = nxt.result()
            nxt = pool.submit(load, starts[k + 1]) if k + 1 < len(starts) else None
            yield (lo, names, cond.to(device, non_blocking=True), gt.to(device, non_blocking=True),
                   [m.to(device) if m is not None else None for m in mask])


def _batch_mask(mask, gt):
    """u8/f32[B,H,W] mask of a batch of files; a file without a 'mask' key counts as all-valid (the reference handles
    masks per file: masked_mae(..., mask=None) weighs every pixel)."""
    if all(m is None for m in mask):
        return None
    ones = torch.ones((1,) + tuple(gt.shape[2:]), device=gt.device, dtype=torch.float32)
    return torch.cat([m if m is not None else ones for m in mask], 0)


def _mstd(a):
    t = torch.tensor(a)
    return t.mean().item(), t.std(unbiased=False).item()


def _recon_batch(args, model, alpha_bar, cond, gt, noise, param, t_start, steps):
    """x0 of the from-noised-GT (eps) / from-noise (v) multistep evaluators for a batch, per-patch noise supplied."""
    if param == "v":
        ab = schedule._abar_cpu(alpha_bar)
        T = len(ab)
        K = max(1, min(int(t_start), T - 1))
        st = schedule.steps_grid_b(alpha_bar, schedule.grid_b(K, steps), "v", eta=float(args.ddim_eta))
        return samplers.run_steps(model, st, cond, noise, init_scale=float(torch.sqrt(1 - ab[K])))
    t_start = max(1, min(int(t_start), len(alpha_bar) - 1))
    a_t = alpha_bar[t_start].view(-1, 1, 1, 1)
    x_t = torch.sqrt(a_t) * gt + torch.sqrt(1 - a_t) * noise
    return samplers.run_steps(model, schedule.steps_eps_grid_a(alpha_bar, t_start, steps), cond, x_t)


def cmd_ddim(args, param):
    device, files, model, alpha_bar = _setup(args)
    maes, mses = [], []
    t0 = time.perf_counter()
    with open(os.path.join(args.out_dir, "ddim_metrics.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["file", "t_start", "ddim_steps", "MAE", "MSE"] if param == "eps" else
                   ["file", "t_start", "ddim_steps", "eta", "MAE", "MSE"])
        for lo, names, cond, gt, mask in _batches(args, files, device):
            noise = torch.cat([torch.randn_like(gt[i:i + 1]) for i in range(len(names))], 0)   # one draw per file
            x0 = _recon_batch(args, model, alpha_bar, cond, gt, noise, param, args.t_start, args.ddim_steps)
            vals = metrics.patch_metrics(x0, gt, _batch_mask(mask, gt)).cpu()     # one fused pass + one copy per batch
            for i, fname in enumerate(names):
                mae, mse = float(vals[i, 0]), float(vals[i, 1])
                maes.append(mae); mses.append(mse)
                row = [fname, args.t_start, args.ddim_steps] + ([args.ddim_eta] if param == "v" else [])
                w.writerow(row + [f"{mae:.6f}", f"{mse:.6f}"])
    with open(os.path.join(args.out_dir, "ddim_summary.txt"), "w") as f:
        f.write(f"files: {len(files)}  t_start: {args.t_start}  steps: {args.ddim_steps}\n")
        f.write(f"MAE mean/std: {_mstd(maes)[0]:.6f} / {_mstd(maes)[1]:.6f}\n")
        f.write(f"MSE mean/std: {_mstd(mses)[0]:.6f} / {_mstd(mses)[1]:.6f}\n")
    print(f"[DONE] DDIM ({len(files) / (time.perf_counter() - t0):.2f} patches/s incl. file I/O)")


def cmd_sweep(args):
    device, files, model, alpha_bar = _setup(args)
    t_list = [int(x) for x in args.t_start_grid.split(",")]
    k_list = [int(x) for x in args.ddim_steps_grid.split(",")]
    with open(os.path.join(args.out_dir, "ddim_sweep_summary.csv"), "w", newline="") as fsum:
        wsum = csv.writer(fsum)
        wsum.writerow(["t_start", "steps", "files", "MAE_mean", "MAE_std", "MSE_mean", "MSE_std", "ms_per_step", "patches_per_s"])
        for t_start in t_list:
            for steps in k_list:
                maes, mses, n_calls = [], [], 0
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for lo, names, cond, gt, mask in _batches(args, files, device):
                    noise = []
                    for i in range(len(names)):       # same file -> same starting noise across all configs (:404)
                        torch.manual_seed(args.seed_base + lo + i)
                        noise.append(torch.randn_like(gt[i:i + 1]))
                    x0 = _recon_batch(args, model, alpha_bar, cond, gt, torch.cat(noise, 0), args.param, t_start, steps)
                    n_calls = steps if args.param == "eps" else len(schedule.grid_b(max(1, min(t_start, args.T - 1)), steps))
                    vals = metrics.patch_metrics(x0, gt, _batch_mask(mask, gt)).cpu()
                    maes += [float(v) for v in vals[:, 0]]
                    mses += [float(v) for v in vals[:, 1]]
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                nb = (len(files) + args.batch - 1) // args.batch
                wsum.writerow([t_start, steps, len(files), f"{_mstd(maes)[0]:.6f}", f"{_mstd(maes)[1]:.6f}",
                               f"{_mstd(mses)[0]:.6f}", f"{_mstd(mses)[1]:.6f}", f"{dt / (nb * max(n_calls, 1)) * 1e3:.3f}",
                               f"{len(files) / dt:.3f}"])
    print("[DONE] DDIM sweep")


def cmd_true_infer(args):
    device, files, model, alpha_bar = _setup(args)
    agg = {k: [] for k in ("mae", "mse", "psnr", "sam", "ergas")}
    with open(os.path.join(args.out_dir, "ddim_true_infer_metrics.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["file", "t_start", "ddim_steps", "seeds", "MAE_mean", "MAE_std", "MSE_mean", "MSE_std", "PSNR_mean",
                    "SAM_mean", "ERGAS_mean"])
        for lo, names, cond, gt, mask in _batches(args, files, device):
            per = [{k: [] for k in agg} for _ in names]
            for s in range(args.n_seeds):
                noise = []
                for i in range(len(names)):           # the reference re-seeds before every generate call (:553)
                    torch.manual_seed(args.seed_base + s)
                    noise.append(torch.randn((1, gt.size(1), gt.size(2), gt.size(3)), device=device))
                x0 = samplers.ddpm_ddim_generate(model, cond, alpha_bar, t_start=args.t_start, steps=args.ddim_steps,
                                                 noise=torch.cat(noise, 0))
                vals = metrics.patch_metrics(x0, gt, _batch_mask(mask, gt)).cpu()   # one fused pass + one copy per batch
                for i in range(len(names)):
                    for k, col in (("mae", 0), ("mse", 1), ("psnr", 2), ("sam", 4), ("ergas", 5)):
                        per[i][k].append(float(vals[i, col]))
            for i, fname in enumerate(names):
                mu = {k: float(np.mean(v)) for k, v in per[i].items()}
                sd = {k: float(np.std(v, ddof=0)) for k, v in per[i].items()}
                w.writerow([fname, args.t_start, args.ddim_steps, args.n_seeds, f"{mu['mae']:.6f}", f"{sd['mae']:.6f}",
                            f"{mu['mse']:.6f}", f"{sd['mse']:.6f}", f"{mu['psnr']:.3f}", f"{mu['sam']:.4f}", f"{mu['ergas']:.2f}"])
                for k in agg:
                    agg[k].append(mu[k])
    with open(os.path.join(args.out_dir, "ddim_true_infer_summary.txt"), "w") as f:
        f.write(f"files: {len(files)}  t_start: {args.t_start}  steps: {args.ddim_steps}  seeds: {args.n_seeds}\n")
        f.write(f"MAE mean/std:  {_mstd(agg['mae'])[0]:.6f} / {_mstd(agg['mae'])[1]:.6f}\n")
        f.write(f"MSE mean/std:  {_mstd(agg['mse'])[0]:.6f} / {_mstd(agg['mse'])[1]:.6f}\n")
        f.write(f"PSNR mean/std: {_mstd(agg['psnr'])[0]:.3f} / {_mstd(agg['psnr'])[1]:.3f}\n")
        f.write(f"SAM  mean/std: {_mstd(agg['sam'])[0]:.4f} / {_mstd(agg['sam'])[1]:.4f}\n")
        f.write(f"ERGAS mean/std:{_mstd(agg['ergas'])[0]:.2f} / {_mstd(agg['ergas'])[1]:.2f}\n")
    print("[DONE] DDIM (TRUE-INFER)")


def cmd_night_demo(args):
    """Evaluation_Pure_Generation.py --mode night_demo (:720-727): pure generation from the conditioning alone (no target
    is read) for the first max(1, --save_viz_n) files.  The reference renders a PNG panel per file; rendering is outside
    the hot path, so the generated patch is saved as `viz/NNN_night_pred.npy` (f32[4,H,W]) next to the conditioning
    `viz/NNN_night_cond.npy` -- what save_panel would have drawn."""
    device, files, model, alpha_bar = _setup(args)
    viz_dir = os.path.join(args.out_dir, "viz")
    os.makedirs(viz_dir, exist_ok=True)
    files = files[:max(1, args.save_viz_n)]
    for lo, names, cond, gt, mask in _batches(args, files, device):
        noise = torch.cat([torch.randn_like(gt[i:i + 1]) for i in range(len(names))], 0)      # one draw per file (:281)
        x0 = samplers.ddpm_ddim_generate(model, cond, alpha_bar, t_start=args.t_start, steps=args.ddim_steps, noise=noise)
        for i in range(len(names)):
            np.save(os.path.join(viz_dir, f"{lo + i:03d}_night_pred.npy"), x0[i].cpu().numpy())
            np.save(os.path.join(viz_dir, f"{lo + i:03d}_night_cond.npy"), cond[i].cpu().numpy())
    print("[DONE] NIGHT_DEMO")


def cmd_limitation(args):
    """run_eval of Limitation_Test.py (:273-400, eps) / Limitation_Test_v_Prediction.py (:258-372, v): --mode ddpm|ddim,
    --time_schedule cosine|linear, --batch_size, --band_weights, --partial_reverse_k (eps), --save_n; prints the reference's
    report (equal-channel, optional band-weighted, per-channel, all pixel-weighted over the whole dataset) and also writes it
    to limitation_summary.txt.  PNG previews are not produced."""
    os.makedirs(args.out_dir, exist_ok=True)
    device = torch.device("cuda")
    torch.manual_seed(args.seed)
    files = _files(args)
    _, x_gt0, _, Cc, Ct = load_npz_as_tensors(os.path.join(args.patch_dir, files[0]), device)
    print(f"[INFO] inputs={Cc}, target={Ct}")
    args.batch = args.batch_size
    model = load_model(args.ckpt, Cc + Ct, Ct, args.base_ch, device, args.batch_size)
    betas = schedule.make_schedule(args.T, args.time_schedule)
    betas, alphas, alpha_bar = schedule.derive(betas)
    betas, alphas, alpha_bar = betas.to(device), alphas.to(device), alpha_bar.to(device)
    tot, saved, lines = None, 0, []
    for bi, (lo, names, cond, gt, mask) in enumerate(_batches(args, files, device)):
        if args.param == "v":
            if args.mode == "ddpm":
                x_pred = samplers.sample_ddpm_v(model, cond, betas, alphas, alpha_bar, Ct, seed=args.seed + bi)
            else:
                x_pred = samplers.sample_ddim_v(model, cond, alpha_bar, Ct, steps=args.ddim_steps, eta=args.ddim_eta,
                                                t_start=args.lim_t_start, seed=args.seed + bi)
        elif args.mode == "ddpm":
            x_pred = samplers.ddpm_sample(model, cond, betas, alphas, alpha_bar, Ct, seed=args.seed + bi)
        else:
            x_pred = samplers.ddim_sample(model, cond, alphas, alpha_bar, Ct, steps=args.ddim_steps)
        mk = _batch_mask(mask, gt)
        sums = metrics.channelwise_error_sums(x_pred, gt, mk)
        tot = sums if tot is None else tuple(a + b for a, b in zip(tot, sums))
        for b in range(len(names)):
            if saved >= args.save_n:
                break
            stem = f"{args.mode}_{bi:04d}_{b:02d}"
            np.save(os.path.join(args.out_dir, f"{stem}_pred.npy"), x_pred[b].cpu().numpy())
            np.save(os.path.join(args.out_dir, f"{stem}_gt.npy"), gt[b].cpu().numpy())
            saved += 1
        if args.partial_reverse_k and bi == 0 and args.param == "eps":
            for k in args.partial_reverse_k:
                xr = samplers.partial_ddim_from_gt(model, gt, cond, alpha_bar, int(k))
                mae_k, mse_k, ps_k, _, _, _ = metrics.aggregate_final(*metrics.channelwise_error_sums(xr, gt, mk))
                lines.append(f"[partial-reverse k={k}] MAE={mae_k:.6f}  MSE={mse_k:.6f}  PSNR={ps_k:.3f} dB")
    mae, mse, ps, mae_c, mse_c, ps_c = metrics.aggregate_final(*tot)
    lines += ["", "==== Unweighted (equal-channel) ====", f"MAE:  {mae:.6f}", f"MSE:  {mse:.6f}", f"PSNR: {ps:.3f} dB"]
    if args.band_weights:
        mae_w, mse_w, ps_w, _, _, _ = metrics.aggregate_final(*tot, band_weights=args.band_weights)
        lines += ["", "==== Weighted (band_weights) ====", f"band_weights = {args.band_weights}", f"MAE_w:  {mae_w:.6f}",
                  f"MSE_w:  {mse_w:.6f}", f"PSNR_w: {ps_w:.3f} dB"]
    bands = ["B2", "B3", "B4", "B8"] if len(mae_c) == 4 else [f"Band{i}" for i in range(len(mae_c))]
    lines += ["", "-- Per-channel metrics (pixel-weighted) --"]
    lines += [f"{nm:>3s}:  MAE={mae_c[i]:.6f}  MSE={mse_c[i]:.6f}  PSNR={ps_c[i]:.3f} dB" for i, nm in enumerate(bands)]
    print("\n".join(lines))
    with open(os.path.join(args.out_dir, "limitation_summary.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    print(f"\n[INFO] Results saved to: {args.out_dir}")


def cmd_onestep(args):
    os.makedirs(args.out_dir, exist_ok=True)
    device = torch.device("cuda")
    files = _files(args)
    x_cond, x_gt, mask, Cc, Ct = load_npz_as_tensors(os.path.join(args.patch_dir, files[0]), device)
    model = load_model(args.ckpt, Cc + Ct, Ct, args.base_ch, device, 1)
    _, _, alpha_bar = schedule.derive(schedule.cosine_beta_schedule(args.T))
    alpha_bar = alpha_bar.to(device)
    fn = samplers.one_step_recon_v if args.param == "v" else samplers.one_step_recon
    if args.param == "v":       # the v script's t=0 identity check is real (Onestep_v_Prediction.py:184-197)
        mae0, mse0, _ = fn(model, x_gt, x_cond, alpha_bar, mask, 0, noise=torch.zeros_like(x_gt), allow_t0=True)
    else:                       # the eps script's is vacuous: x0_hat_t0 = x_t0 (Onestep.py:139)
        mae0, mse0 = metrics.masked_mae(x_gt, x_gt, mask), metrics.masked_mse(x_gt, x_gt, mask)
    print(f"[t=0 identity] MAE={mae0:.6f}  MSE={mse0:.6f}  (should be ~0.0)")
    mae, mse, _ = fn(model, x_gt, x_cond, alpha_bar, mask, args.t_small)
    print(f"[one-step@t={max(1, min(args.t_small, args.T - 1))}] MAE={mae:.6f}  MSE={mse:.6f}")


def cmd_scene(args):
    import torch.distributed as dist
    from . import scene as sc
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    if args.scene_npy:
        scene = torch.from_numpy(np.load(args.scene_npy).astype(np.float32))
    else:
        scene = sc.synthetic_scene(args.scene_size, args.scene_size, seed=0)
    model = load_model(args.ckpt, 8, 4, args.base_ch, device, args.batch) if args.ckpt else None
    if model is None:
        torch.manual_seed(1235)
        model = UNetSmallB200(8, 4, args.base_ch, max_batch=args.batch).to(device).eval()
    _, _, alpha_bar = schedule.derive(schedule.cosine_beta_schedule(args.T))
    scene = scene.to(device)
    # one-off costs outside the timed region: activation arena + weight repack, NCCL communicator
    model.engine(device, args.patch_size, args.patch_size, args.batch)
    if world > 1:
        dist.all_reduce(torch.zeros(1, device=device))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = sc.generate_scene(model, scene, alpha_bar, ps=args.patch_size, stride=args.stride, param=args.param,
                            steps=args.ddim_steps, t_start=args.t_start, batch=args.batch, seed_base=args.seed_base,
                            valid_ratio_threshold=args.valid_ratio_threshold, rank=rank, world=world,
                            window=None if args.blend == "uniform" else args.blend)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    if rank == 0:
        os.makedirs(args.out_dir, exist_ok=True)
        np.save(os.path.join(args.out_dir, "scene_pred.npy"), res["canvas"].cpu().numpy())
        np.save(os.path.join(args.out_dir, "scene_cover.npy"), res["cover"].cpu().numpy())
        n = int(res["kept"].sum())
        print(f"[DONE] scene {tuple(scene.shape)}: {n} patches on {world} GPU(s) in {dt:.2f} s = {n / dt:.2f} patches/s")
    if world > 1:
        dist.destroy_process_group()


def main(argv=None):
    ap = argparse.ArgumentParser("s1s2_b200 drivers")
    ap.add_argument("cmd", choices=["onestep", "ddim", "ddim_v", "ddim_sweep", "true_infer", "night_demo", "limitation", "scene"])
    ap.add_argument("--patch_dir")
    ap.add_argument("--ckpt")
    ap.add_argument("--out_dir", required=True)
    ap.add_argument("--T", type=int, default=1000)
    ap.add_argument("--base_ch", type=int, default=96)
    ap.add_argument("--max_files", type=int, default=0)
    ap.add_argument("--t_start", type=int, default=200)
    ap.add_argument("--ddim_steps", type=int, default=20)
    ap.add_argument("--ddim_eta", type=float, default=0.0)
    ap.add_argument("--t_small", type=int, default=20)
    ap.add_argument("--n_seeds", type=int, default=8)
    ap.add_argument("--seed_base", type=int, default=1234)
    ap.add_argument("--t_start_grid", default="300,200,150,100")
    ap.add_argument("--ddim_steps_grid", default="10,20,50,100")
    ap.add_argument("--true_infer", action="store_true")
    ap.add_argument("--save_viz_n", type=int, default=6)
    # Limitation_Test*.py flags (the v script's --t_start is --lim_t_start here: default None = start from T-1)
    ap.add_argument("--mode", default="ddim", choices=["ddpm", "ddim"])
    ap.add_argument("--time_schedule", default="cosine", choices=["cosine", "linear"])
    ap.add_argument("--batch_size", type=int, default=2)
    ap.add_argument("--save_n", type=int, default=16)
    ap.add_argument("--band_weights", nargs="*", type=float, default=None)
    ap.add_argument("--partial_reverse_k", nargs="*", type=int, default=None)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--lim_t_start", type=int, default=None)
    # additions
    ap.add_argument("--param", default="eps", choices=["eps", "v"])
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--scene_npy", default=None, help="f32[4,H,W] scene (HH dB, HV dB, incidence deg, elevation m)")
    ap.add_argument("--scene_size", type=int, default=2048)
    ap.add_argument("--patch_size", type=int, default=256)
    ap.add_argument("--stride", type=int, default=64)
    ap.add_argument("--valid_ratio_threshold", type=float, default=0.0)
    ap.add_argument("--blend", default="uniform", choices=["uniform", "hann"], help="overlap blend of the scene stitch")
    args = ap.parse_args(argv)
    if args.cmd == "onestep":
        cmd_onestep(args)
    elif args.cmd == "ddim":
        cmd_ddim(args, "eps")
    elif args.cmd == "ddim_v":
        cmd_ddim(args, "v")
    elif args.cmd == "ddim_sweep":
        cmd_sweep(args)
    elif args.cmd == "true_infer":
        cmd_true_infer(args)
    elif args.cmd == "night_demo":
        cmd_night_demo(args)
    elif args.cmd == "limitation":
        cmd_limitation(args)
    else:
        cmd_scene(args)


if __name__ == "__main__":
    main()
