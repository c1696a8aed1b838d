"""GPU parity, second tier: every sampler entry the header names against the CPU oracle at the BASELINE sizes, the
small-batch tilings, the trained-scale (fp16 range) stress with the saturation counter, the pipelined host entry and the
patch-side additions (keyed noise, Hann-weighted stitch, out-of-scene windows).

Tolerances as in test_gpu_parity.py: model call vs fp32 oracle rel-L2 <= 5e-3 and max|err| <= 2e-2 max|ref|; scheduler
updates bit-exact given the network output; one-step / few-step images within 2e-3 absolute of the oracle's.
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import metrics as ometrics
from oracle import patch as opatch
from oracle import samplers as osamplers
from oracle import schedule as osched
from oracle import unet as ounet
from test_gpu_parity import _inputs, _ref_update

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import s1s2_b200
    assert torch.cuda.is_available()
    dev = torch.device("cuda:0")
    sd = ounet.init_state_dict(8, 4, 96, seed=1234)
    model = s1s2_b200.UNetSmallB200(8, 4, 96, max_batch=4).to(dev)
    model.load_state_dict(sd, strict=True)
    model.eval()
    _, alphas, abar = osched.make_schedule(1000)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    return dict(dev=dev, sd=sd, model=model, oracle=ounet.OracleModel(sd), abar=abar, alphas=alphas)


def _gt(B, H, W, seed):
    return torch.rand((B, 4, H, W), generator=torch.Generator().manual_seed(seed))


# ------------------------------------------------------------------------------------------------ one-step (config 1)
def test_one_step_recon_eps_matches_oracle(env):
    """BASELINE config 1 (Evaluation/Onestep.py:149-164, DDIM_Multi-step.py:155-170): one denoise at t_small = 20, batch 1,
    256 x 256, same noise; also the reference's clamp of t_small to [1, T-1]."""
    from s1s2_b200 import samplers
    dev, ab = env["dev"], env["abar"]
    z, cond = _inputs(1, 256, 256, seed=601)
    x_gt = _gt(1, 256, 256, 61)
    mask = (torch.rand((1, 256, 256), generator=torch.Generator().manual_seed(62)) > 0.1).float()
    for t_small in (20, 0):                       # 0 is clamped to 1 like the reference
        ref_x0, ref_eps, _ = osamplers.one_step_eps(env["oracle"], x_gt, cond, ab, t_small, z)
        mae, mse, x0 = samplers.one_step_recon(env["model"], x_gt.to(dev), cond.to(dev), ab.to(dev), mask.to(dev), t_small,
                                               noise=z.to(dev))
        x0 = x0.cpu()
        assert float((x0 - ref_x0).abs().max()) <= 2e-3
        assert abs(mae - ometrics.masked_mae(ref_x0, x_gt, mask)) <= 1e-5
        assert abs(mse - ometrics.masked_mse(ref_x0, x_gt, mask)) <= 1e-5


def test_one_step_recon_v_matches_oracle_and_t0_identity(env):
    """Onestep_v_Prediction.py:184-227: the t = 0 identity check (x0 = sqrt(abar_0) x_gt - sqrt(1-abar_0) v, so the error is
    0.0064 |v|) and the one-step reconstruction at t = 20; the public function clamps t_small to [1, T-1] like the
    reference, the identity check goes through allow_t0."""
    from s1s2_b200 import samplers
    dev, ab = env["dev"], env["abar"]
    z, cond = _inputs(1, 256, 256, seed=602)
    x_gt = _gt(1, 256, 256, 63)
    mask = torch.ones((1, 256, 256))
    ref_x0, _, _ = osamplers.one_step_v(env["oracle"], x_gt, cond, ab, 20, z)
    mae, mse, x0 = samplers.one_step_recon_v(env["model"], x_gt.to(dev), cond.to(dev), ab.to(dev), mask.to(dev), 20, noise=z.to(dev))
    assert float((x0.cpu() - ref_x0).abs().max()) <= 2e-3
    assert abs(mae - ometrics.masked_mae(ref_x0, x_gt, mask)) <= 1e-5
    # t = 0: identity up to sqrt(1 - abar_0) * v
    ref0, v0, _ = osamplers.one_step_v(env["oracle"], x_gt, cond, ab, 0, torch.zeros_like(x_gt))
    mae0, mse0, x00 = samplers.one_step_recon_v(env["model"], x_gt.to(dev), cond.to(dev), ab.to(dev), mask.to(dev), 0,
                                                noise=torch.zeros_like(x_gt).to(dev), allow_t0=True)
    assert float((x00.cpu() - ref0).abs().max()) <= 1e-4
    assert mae0 <= 0.0065 * float(v0.abs().mean()) + 1e-4
    # without allow_t0 the call runs at t = 1, as DDIM_Multi-step_v_Prediction.py:213 does
    ref1, _, _ = osamplers.one_step_v(env["oracle"], x_gt, cond, ab, 1, z)
    _, _, x01 = samplers.one_step_recon_v(env["model"], x_gt.to(dev), cond.to(dev), ab.to(dev), mask.to(dev), 0, noise=z.to(dev))
    assert float((x01.cpu() - ref1).abs().max()) <= 2e-3


def test_one_step_rng_seed_reseeds_the_global_generator_like_the_reference(env):
    from s1s2_b200 import samplers
    dev, ab = env["dev"], env["abar"]
    _, cond = _inputs(1, 32, 32, seed=603)
    x_gt = _gt(1, 32, 32, 64).to(dev)
    mask = torch.ones((1, 32, 32), device=dev)
    _, _, a = samplers.one_step_recon(env["model"], x_gt, cond.to(dev), ab.to(dev), mask, 20, rng_seed=77)
    after = torch.randn(3, device=dev)
    torch.manual_seed(77)
    z = torch.randn_like(x_gt)
    want_after = torch.randn(3, device=dev)
    _, _, b = samplers.one_step_recon(env["model"], x_gt, cond.to(dev), ab.to(dev), mask, 20, noise=z)
    assert torch.equal(a, b) and torch.equal(after, want_after)


def test_partial_ddim_from_gt_k5(env):
    """Limitation_Test.py:252-270 with k = 5 (five consecutive timesteps 5..1), free-running against the oracle."""
    from s1s2_b200 import samplers
    dev, ab = env["dev"], env["abar"]
    z, cond = _inputs(2, 64, 64, seed=604)
    x_gt = _gt(2, 64, 64, 65)
    ref = osamplers.partial_ddim_from_gt(env["oracle"], x_gt, cond, ab, 5, z)
    got = samplers.partial_ddim_from_gt(env["model"], x_gt.to(dev), cond.to(dev), ab.to(dev), 5, noise=z.to(dev)).cpu()
    assert float((got - ref).abs().max()) <= 2e-3
    assert ometrics.psnr(got, ref) >= 50.0
    ref0 = osamplers.partial_ddim_from_gt(env["oracle"], x_gt, cond, ab, 0, z)
    got0 = samplers.partial_ddim_from_gt(env["model"], x_gt.to(dev), cond.to(dev), ab.to(dev), 0, noise=z.to(dev)).cpu()
    assert torch.equal(got0, ref0)                 # k = 0: no model call, clamp of the noised ground truth


# ------------------------------------------------------------------------------------------------ full-size chains
def _teacher_forced_sparse(env, steps, cond, x_init, init_scale, every):
    """All scheduler updates bit-exact given our network output; the network output itself compared with the fp32 oracle
    on OUR state at every `every`-th call (and the first and last)."""
    from s1s2_b200 import samplers
    dev = env["dev"]
    out, taps = samplers.run_steps(env["model"], steps, cond.to(dev), x_init.to(dev), init_scale=init_scale,
                                   tap_pred=True, tap_x=True)
    torch.cuda.synchronize()
    B = cond.shape[0]
    x_in = x_init * torch.tensor(init_scale, dtype=torch.float32)
    worst, checked = 0.0, 0
    preds, xs = taps["pred"].cpu(), taps["x"].cpu()
    for i, st in enumerate(steps):
        got = preds[i]
        if i % every == 0 or i == len(steps) - 1:
            ref = env["oracle"](torch.cat([x_in, cond], 1), torch.full((B,), st.t, dtype=torch.long))
            rel = float((got - ref).norm() / ref.norm())
            assert rel <= 5e-3, (i, st.t, rel)
            assert float((got - ref).abs().max()) <= 2e-2 * float(ref.abs().max()), (i, st.t)
            worst, checked = max(worst, rel), checked + 1
        want = _ref_update(st, x_in, got, None)
        assert torch.equal(xs[i], want), (i, st.t, float((xs[i] - want).abs().max()))
        x_in = xs[i]
    assert torch.equal(out.cpu(), x_in)
    return worst, checked


def test_ddim_sample_eps_grid_b_full_size_teacher_forced(env):
    """Limitation_Test.py:227-249 (`ddim_sample`: eps model on grid B from T-1 = 999, 50 steps, t = 0 included) at 256 x 256."""
    from s1s2_b200 import schedule
    x, cond = _inputs(1, 256, 256, seed=605)
    steps = schedule.steps_grid_b(env["abar"], schedule.grid_b(999, 50, force_append=False), "eps")
    worst, n = _teacher_forced_sparse(env, steps, cond, x, 1.0, every=1)
    print(f"[eps grid B parity] {n} teacher-forced model calls at 256x256: worst per-step eps rel-L2 {worst:.2e}")


@pytest.mark.parametrize("n_steps,every", [(10, 1), (25, 1), (100, 4), (250, 10)])
def test_sweep_grids_teacher_forced_full_size(env, n_steps, every):
    """BASELINE config 4's other step counts (v model, grid B from 999) at 256 x 256: every scheduler update bit-exact, the
    predicted v checked against the oracle on every `every`-th call."""
    from s1s2_b200 import schedule
    ab = env["abar"]
    x, cond = _inputs(1, 256, 256, seed=610 + n_steps)
    steps = schedule.steps_grid_b(ab, schedule.grid_b(999, n_steps), "v")
    worst, n = _teacher_forced_sparse(env, steps, cond, x, float(torch.sqrt(1 - ab[999])), every)
    print(f"[sweep parity] {n_steps} steps ({len(steps)} calls, {n} checked): worst per-step v rel-L2 {worst:.2e}")


# ------------------------------------------------------------------------------------------------ tilings / range
def test_small_batch_tilings_do_not_change_bits(env):
    """The narrow tilings picked at small batches (pick_variant: N = 96 / 192 column tiles when the wide ones cannot fill
    the 74 CTA pairs) accumulate every output element in the same order as the default ones: a model with the alternates
    disabled produces bit-identical activations and outputs, at batch 1 and 3, 256 x 256."""
    import s1s2_b200
    dev = env["dev"]
    os.environ["S1S2_NO_ALTS"] = "1"
    try:
        plain = s1s2_b200.UNetSmallB200(8, 4, 96, max_batch=3).to(dev)
        plain.load_state_dict(env["sd"], strict=True)
        plain.eval()
        plain.engine(dev, 256, 256, 3)            # the handle reads the switch at creation
    finally:
        del os.environ["S1S2_NO_ALTS"]
    for B in (1, 3):
        x, cond = _inputs(B, 256, 256, seed=620 + B)
        t = torch.tensor([999, 501, 20][:B], dtype=torch.long)
        xin = torch.cat([x, cond], 1).to(dev)
        a = env["model"](xin, t.to(dev))
        b = plain(xin, t.to(dev))
        assert torch.equal(a, b), B
        for name in ("down3", "conv3", "conv2", "down1"):
            assert torch.equal(env["model"].activation(name, B), plain.activation(name, B)), (B, name)
    del plain


def test_tile_width_choice_by_batch(env):
    """pick_variant: the widest tile at large batches (least weight traffic per flop), narrower ones where the wide tiles
    would leave most of the 74 CTA pairs idle (the 64 x 64 layers of a single 256 x 256 patch have 16 pairs of M tiles)."""
    import s1s2_b200
    dev = env["dev"]
    m = s1s2_b200.UNetSmallB200(8, 4, 96, max_batch=64).to(dev)
    m.load_state_dict(env["sd"], strict=True)
    w64 = dict(m.tile_widths(dev, 256, 256, 64))
    w1 = dict(m.tile_widths(dev, 256, 256, 1))
    assert w64["down3.0.2"] == 256 and w64["conv3.0"] == 192 and w64["down2.0.2"] == 192 and w64["conv1.0"] == 96
    assert w64["up3"] == 256 and w64["up1"] == 192 and w64["inc.0"] == 96 and w64["conv1.2"] == 96
    assert w1["down3.0.2"] == 192 and w1["down3.0.0"] == 192 and w1["conv3.0"] == 96 and w1["conv3.2"] == 96
    assert all(w1[k] <= w64[k] for k in w64)
    assert len(w64) == 16
    del m


def _scaled_sd(sd, g):
    return {k: (v * g if k.endswith(".weight") and not k.startswith("outc") else v.clone()) for k, v in sd.items()}


def test_trained_scale_stress_and_saturation_counter(env):
    """fp16 range of the activation arena ("fp16 overflow").  Random-init weights shrink activations
    layer by layer; trained checkpoints do not.  Every conv / transposed-conv weight x3 makes the activations GROW through
    the network (up to ~1e3-1e4 from the t = 999 plane): still inside fp16, parity with the fp32 oracle holds and the
    saturation counter reads zero everywhere.  x8 pushes them past 65504: the epilogues' cvt.satfinite clamps, and the
    debug counter reports which layers did."""
    import s1s2_b200
    dev = env["dev"]
    x, cond = _inputs(2, 64, 64, seed=630)
    t = torch.tensor([999, 20], dtype=torch.long)
    xin = torch.cat([x, cond], 1)
    m = s1s2_b200.UNetSmallB200(8, 4, 96, max_batch=2).to(dev)
    peak = {}
    for g in (3.0, 8.0):
        sd = _scaled_sd(env["sd"], g)
        m.load_state_dict(sd, strict=True)
        taps = {}
        ref = ounet.forward(sd, xin, t, taps=taps)
        got = m(xin.to(dev), t.to(dev)).cpu()
        counts = m.saturation_counts(2)
        peak[g] = max(float(v.abs().max()) for k, v in taps.items() if k != "outc")
        if g == 3.0:
            assert 1e3 <= peak[g] <= 6e4, peak[g]
            assert sum(counts.values()) == 0, counts
            rel = float((got - ref).norm() / ref.norm())
            assert rel <= 5e-3, rel
        else:
            assert peak[g] > 65504.0, peak[g]
            hit = {k: v for k, v in counts.items() if v}
            assert hit, counts
            assert counts["xin16"] == 0 and counts["inc"] == 0       # the input record and the first layer stay in range
    print(f"[fp16 range] peak |activation| of the oracle: x3 weights {peak[3.0]:.3g}, x8 weights {peak[8.0]:.3g}")
    del m


def test_forward_poisons_out_of_range_timesteps(env):
    """t outside [0, 2048] cannot be represented exactly in the fp16 time planes: the affected patch returns NaN instead
    of a silently rounded timestep; the other patches of the batch are untouched."""
    dev = env["dev"]
    x, cond = _inputs(2, 32, 32, seed=640)
    xin = torch.cat([x, cond], 1).to(dev)
    good = env["model"](xin, torch.tensor([999, 20], device=dev))
    bad = env["model"](xin, torch.tensor([999, 5000], device=dev))
    assert torch.equal(bad[0], good[0])
    assert bool(torch.isnan(bad[1]).all())


# ------------------------------------------------------------------------------------------------ host pipeline
def test_sample_host_stream_pipeline_matches_device_entry(env):
    """s1s2_sample_host_stream: 7 host patches in batches of 3 (two full batches + a ragged one; three staging rotations)
    equal the device-resident entry patch by patch; called twice to exercise buffer reuse."""
    from s1s2_b200 import samplers, schedule
    dev, ab = env["dev"], env["abar"]
    x, cond = _inputs(7, 32, 32, seed=650)
    steps = schedule.steps_grid_b(ab, schedule.grid_b(999, 4), "v")
    s = float(torch.sqrt(1 - ab[999]))
    want = torch.cat([samplers.run_steps(env["model"], steps, cond[i:i + 1].to(dev), x[i:i + 1].to(dev), init_scale=s).cpu()
                      for i in range(7)], 0)
    ch, xh = cond.pin_memory(), x.pin_memory()
    for _ in range(2):
        got = samplers.run_steps_host(env["model"], steps, ch, xh, init_scale=s, device=dev, batch=3)
        assert torch.equal(got, want)
    got1 = samplers.run_steps_host(env["model"], steps, cond, x, init_scale=s, device=dev, batch=4)      # pageable memory
    assert torch.equal(got1, want)


# ------------------------------------------------------------------------------------------------ patch side
@pytest.mark.parametrize("H,W,ps,st", [(40, 56, 16, 8), (70, 93, 32, 19), (96, 96, 32, 16)])
def test_stitch_hann_bitexact_vs_oracle(env, H, W, ps, st):
    """SURVEY.md section 8 a9's optional Hann-weighted blend (parity unpinned: no reference stitch): separable window,
    sum(w * pred) / sum(w) in ascending patch order -- bit-exact against the numpy statement, vector and scalar paths."""
    from s1s2_b200 import patch
    org = patch.tile_origins(H, W, ps, st)
    g = np.random.default_rng(H * 7 + W)
    preds = g.random((len(org), 4, ps, ps), dtype=np.float32)
    keep = g.random(len(org)) > 0.2
    win = opatch.hann_window(ps)
    assert np.array_equal(patch.hann_window(ps).numpy(), win) and float(win.min()) > 0.0
    ref, cov = opatch.stitch(preds[keep], org[keep], H, W, window=win)
    canvas, cover = patch.stitch(torch.from_numpy(preds[keep]).to(env["dev"]), org[keep], ps, st, H, W, window="hann")
    assert np.array_equal(cover.cpu().numpy(), cov)
    assert np.array_equal(canvas.cpu().numpy(), ref)
    one = np.ones(ps, np.float32)                 # an all-ones window is the uniform blend
    c1, _ = patch.stitch(torch.from_numpy(preds[keep]).to(env["dev"]), org[keep], ps, st, H, W, window=torch.from_numpy(one))
    c0, _ = patch.stitch(torch.from_numpy(preds[keep]).to(env["dev"]), org[keep], ps, st, H, W)
    assert torch.equal(c0, c1)


def test_patch_noise_is_keyed_by_global_index(env):
    from s1s2_b200 import scene as sc
    dev = env["dev"]
    a = sc.patch_noise([3, 7, 11], (4, 64, 64), 1234, dev)
    b = sc.patch_noise([11], (4, 64, 64), 1234, dev)
    c = sc.patch_noise([7, 3], (4, 64, 64), 1234, dev)
    assert torch.equal(a[2], b[0]) and torch.equal(a[1], c[0]) and torch.equal(a[0], c[1])
    assert not torch.equal(a[0], a[1])
    assert not torch.equal(sc.patch_noise([3], (4, 64, 64), 1235, dev)[0], a[0])
    z = sc.patch_noise(list(range(16)), (4, 128, 128), 99, dev).double().flatten()
    n = z.numel()
    assert abs(float(z.mean())) < 5.0 / n ** 0.5 and abs(float(z.var()) - 1.0) < 5.0 * (2.0 / n) ** 0.5
    assert abs(float((z ** 4).mean()) - 3.0) < 0.05 and abs(float((z ** 3).mean())) < 0.03
    # the per-file torch seeding of DDIM_Sweep.py:193,404 stays available
    t = sc.patch_noise([5], (4, 32, 32), 1234, dev, method="torch")
    g = torch.Generator(device=dev).manual_seed(1239)
    assert torch.equal(t[0], torch.randn((4, 32, 32), generator=g, device=dev))


def test_out_of_scene_windows_yield_invalid_patches(env):
    """C ABI robustness: origins are device data the host entry cannot range-check; a window outside the scene produces an
    all-invalid patch (cond 0, mask 0, ratio 0) / a failed valid-ratio test, never an out-of-bounds read."""
    from s1s2_b200 import _lib
    dev = env["dev"]
    scn = torch.randn((4, 64, 80), device=dev)
    tgt = torch.rand((4, 64, 80), device=dev)
    org = torch.tensor([[0, 0], [48, 64], [49, 64], [-1, 0], [0, 65], [1 << 20, 0]], dtype=torch.int32, device=dev)
    N, ps = org.shape[0], 16
    cond = torch.full((N, 4, ps, ps), 7.0, device=dev)
    mask = torch.full((N, ps, ps), 9, dtype=torch.uint8, device=dev)
    ratio = torch.full((N,), 5.0, device=dev)
    L = _lib.lib()
    _lib.check(L.s1s2_tile_extract(0, scn.data_ptr(), None, 64, 80, org.data_ptr(), N, ps, cond.data_ptr(), mask.data_ptr(),
                                   ratio.data_ptr(), None))
    stats = torch.full((N, 8), -1.0, device=dev)
    th = (C.c_float * 5)(0.8, 1e-4, 0.1, 0.6, 5e-5)
    _lib.check(L.s1s2_tile_filter(0, scn.data_ptr(), 4, tgt.data_ptr(), None, 64, 80, org.data_ptr(), N, ps, th,
                                  stats.data_ptr(), None))
    torch.cuda.synchronize()
    assert ratio[:2].tolist() == [1.0, 1.0] and bool((mask[:2] == 1).all())
    assert ratio[2:].tolist() == [0.0] * 4 and bool((mask[2:] == 0).all()) and bool((cond[2:] == 0).all())
    assert stats[2:, 7].tolist() == [1.0] * 4 and bool((stats[2:, :7] == 0).all())
    assert bool((stats[:2, 0] == 1.0).all())


def test_filter_colloc_uses_greater_than_zero_like_build_mask(env):
    """Patch.py:41-49 tests `colloc > 0` on the float raster: 0.5 is valid, -1 and 0 are not, 256 does not wrap to 0."""
    from s1s2_b200 import patch
    dev = env["dev"]
    scn = torch.randn((4, 32, 32), device=dev)
    tgt = torch.rand((4, 32, 32), device=dev)
    colloc = torch.ones((32, 32), device=dev)
    colloc[:8] = 0.5
    colloc[8:16] = 256.0
    colloc[16:24] = -1.0
    colloc[24:] = 0.0
    st = patch.tile_filter(scn, tgt, np.array([[0, 0]], np.int32), 32, colloc=colloc)
    assert abs(float(st[0, 0]) - 0.5) < 1e-6
    _, mask, ratio = patch.tile_extract(scn, np.array([[0, 0]], np.int32), 32, vmask=colloc)   # any non-zero value is valid
    assert abs(float(ratio[0]) - 0.75) < 1e-6 and int(mask.sum()) == 24 * 32


# ------------------------------------------------------------------------------------------------ base_ch = 64
@pytest.fixture(scope="module")
def env64():
    """UNetSmall(8, 4, 64): the class default of the reference (Train_Orignal.py:99); the scripts pass --base_ch 96."""
    import s1s2_b200
    dev = torch.device("cuda:0")
    sd = ounet.init_state_dict(8, 4, 64, seed=4321)
    model = s1s2_b200.UNetSmallB200(8, 4, 64, max_batch=4).to(dev)
    model.load_state_dict(sd, strict=True)
    model.eval()
    _, alphas, abar = osched.make_schedule(1000)
    return dict(dev=dev, sd=sd, model=model, oracle=ounet.OracleModel(sd), abar=abar, alphas=alphas)


@pytest.mark.parametrize("B,H,W", [(2, 32, 32), (1, 48, 80), (3, 64, 64), (1, 256, 256)])
def test_base64_layers_isolated(env64, B, H, W):
    from layer_ref import check_layers
    x, cond = _inputs(B, H, W, seed=700 + B * 10 + H)
    t = torch.tensor([999, 20, 501][:B], dtype=torch.long)
    y = env64["model"](torch.cat([x, cond], 1).to(env64["dev"]), t.to(env64["dev"]))
    torch.cuda.synchronize()
    rows = check_layers(env64["model"], env64["sd"], y, B)
    bad = [(n, e, m) for n, e, m, _ in rows if not e <= 2.0 ** -9 * m + 1e-6]
    assert not bad, bad


def test_base64_model_call_and_chain_vs_oracle(env64):
    from s1s2_b200 import schedule
    dev, ab = env64["dev"], env64["abar"]
    x, cond = _inputs(1, 256, 256, seed=711)
    xin = torch.cat([x, cond], 1)
    t = torch.tensor([999], dtype=torch.long)
    ref = env64["oracle"](xin, t)
    got = env64["model"](xin.to(dev), t.to(dev)).cpu()
    rel = float((got - ref).norm() / ref.norm())
    assert rel <= 5e-3, rel
    assert float((got - ref).abs().max()) <= 2e-2 * float(ref.abs().max())
    # teacher-forced chains: v on grid B (12 calls) and eps on grid A from 999 (8 calls), every update bit-exact
    x, cond = _inputs(2, 64, 64, seed=712)
    worst, _ = _teacher_forced_sparse(env64, schedule.steps_grid_b(ab, schedule.grid_b(999, 12), "v"), cond, x,
                                      float(torch.sqrt(1 - ab[999])), every=1)
    worst2, _ = _teacher_forced_sparse(env64, schedule.steps_eps_grid_a(ab, 999, 8), cond, x, 1.0, every=1)
    print(f"[base_ch 64] model call rel-L2 {rel:.2e}; chains: worst per-step rel-L2 {worst:.2e} (v), {worst2:.2e} (eps)")


def test_base64_batch_independent(env64):
    dev = env64["dev"]
    x, cond = _inputs(4, 64, 64, seed=713)
    t = torch.tensor([999, 979, 20, 0], dtype=torch.long)
    xin = torch.cat([x, cond], 1).to(dev)
    y = env64["model"](xin, t.to(dev)).clone()
    for i in range(4):
        assert torch.equal(env64["model"](xin[i:i + 1], t[i:i + 1].to(dev))[0], y[i]), i
    x, cond = _inputs(3, 256, 256, seed=714)            # small-batch tilings at full size
    xin = torch.cat([x, cond], 1).to(dev)
    t = torch.tensor([999, 501, 20], dtype=torch.long).to(dev)
    y = env64["model"](xin, t).clone()
    assert torch.equal(env64["model"](xin[1:2], t[1:2])[0], y[1])


# ------------------------------------------------------------------------------------------------ DDPM, full length
def test_ddpm_1000_full_chain(env):
    """Limitation_Test.py:209-224 at its real length: the ancestral chain over all T = 1000 timesteps (999 noisy steps), one
    32 x 32 patch.  (a) supplied z: every one of the 1000 scheduler updates bit-exact, the predicted eps checked against the
    fp32 oracle on our state at every 20th call; (b) in-kernel Philox noise: finite, clamped, reproducible for a seed."""
    from s1s2_b200 import samplers, schedule
    dev = env["dev"]
    betas = osched.cosine_betas(1000)
    alphas = 1 - betas
    ab = torch.cumprod(alphas, 0)
    assert torch.equal(ab, env["abar"])
    x, cond = _inputs(1, 32, 32, seed=800)
    zs = torch.randn((999, 1, 4, 32, 32), generator=torch.Generator().manual_seed(801))
    steps = schedule.steps_ddpm(betas, alphas, ab, "eps")
    assert len(steps) == 1000 and steps[0].t == 999 and steps[-1].t == 0
    out, taps = samplers.run_steps(env["model"], steps, cond.to(dev), x.to(dev), step_noise=zs.to(dev), tap_pred=True, tap_x=True)
    torch.cuda.synchronize()
    preds, xs = taps["pred"].cpu(), taps["x"].cpu()
    x_in, worst = x.clone(), 0.0
    for i, st in enumerate(steps):
        if i % 20 == 0 or i == 999:
            ref = env["oracle"](torch.cat([x_in, cond], 1), torch.full((1,), st.t, dtype=torch.long))
            rel = float((preds[i] - ref).norm() / ref.norm())
            assert rel <= 5e-3, (i, st.t, rel)
            worst = max(worst, rel)
        z = zs[st.noise_index] if st.noise_index >= 0 else None
        want = _ref_update(st, x_in, preds[i], z)
        assert torch.equal(xs[i], want), (i, st.t)
        x_in = xs[i]
    assert torch.equal(out.cpu(), x_in) and float(out.min()) >= 0.0 and float(out.max()) <= 1.0
    a = samplers.ddpm_sample(env["model"], cond.to(dev), betas, alphas, ab, 4, noise=x.to(dev), seed=5).clone()
    b = samplers.ddpm_sample(env["model"], cond.to(dev), betas, alphas, ab, 4, noise=x.to(dev), seed=5)
    c = samplers.ddpm_sample(env["model"], cond.to(dev), betas, alphas, ab, 4, noise=x.to(dev), seed=6)
    assert torch.equal(a, b) and not torch.equal(a, c) and bool(torch.isfinite(a).all())
    print(f"[DDPM-1000] 1000 updates bit-exact, eps checked on 51 calls: worst rel-L2 {worst:.2e}")
