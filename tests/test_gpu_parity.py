"""GPU parity tests: the CUDA path (through the C ABI, via the ctypes mirror) against the CPU oracle.

Tolerances:
  * per-layer, isolated, fp16-rounded weights and inputs: max|err| <= 2^-9 * max|ref| (fp16 store rounding + order)
  * model call vs fp32 oracle: rel-L2 <= 5e-3 and max|err| <= 2e-2 * max|ref| (fp16 operands, fp32 accumulate)
  * scheduler update given the network output: bit-exact (fp32, same operation order, no FMA contraction)
  * free-running v-DDIM image: PSNR(ours, oracle) >= 40 dB, |PSNR/SSIM vs target| within 0.1 dB / 0.002
  * tiling indices, masks, stitch: bit-exact; z-score values <= 2e-6 abs
"""
import os

import numpy as np
import pytest
import torch

from oracle import metrics as ometrics
from oracle import patch as opatch
from oracle import samplers as osamplers
from oracle import schedule as osched
from oracle import unet as ounet

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(__file__), "golden")


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
    return dict(dev=dev, sd=sd, model=model, oracle=ounet.OracleModel(sd), abar=abar, alphas=alphas)


def _inputs(B, H, W, seed=0):
    g = torch.Generator().manual_seed(seed)
    cond = torch.randn((B, 4, H, W), generator=g)
    cond[:, 2] = torch.rand((B, H, W), generator=g) * 0.4 + 0.2
    cond[:, 3] = (torch.randn((B, H, W), generator=g) * 0.3 + 0.3).abs()
    x = torch.randn((B, 4, H, W), generator=g)
    return x, cond


# ------------------------------------------------------------------------------------------------ denoiser
@pytest.mark.parametrize("B,H,W", [(2, 32, 32), (3, 64, 64), (1, 48, 80), (1, 256, 256), (3, 16, 16), (2, 128, 384)])
def test_layers_isolated(env, B, H, W):
    from layer_ref import check_layers
    x, cond = _inputs(B, H, W, seed=B * 1000 + H)
    t = torch.tensor([999, 20, 501][:B], dtype=torch.long)
    y = env["model"](torch.cat([x, cond], 1).to(env["dev"]), t.to(env["dev"]))
    torch.cuda.synchronize()
    rows = check_layers(env["model"], env["sd"], y, B)
    bad = [(n, e, m) for n, e, m, _ in rows if not e <= 2.0 ** -9 * m + 1e-6]
    assert not bad, bad


@pytest.mark.parametrize("B,H,W", [(2, 64, 64), (1, 256, 256)])
def test_model_call_vs_oracle(env, B, H, W):
    x, cond = _inputs(B, H, W, seed=11)
    t = torch.tensor([999, 20][:B], dtype=torch.long)
    xin = torch.cat([x, cond], 1)
    ref = env["oracle"](xin, t)
    got = env["model"](xin.to(env["dev"]), t.to(env["dev"])).cpu()
    rel = float((got - ref).norm() / ref.norm())
    assert rel <= 5e-3, rel
    assert float((got - ref).abs().max()) <= 2e-2 * float(ref.abs().max())


def test_model_call_large_dynamic_range(env):
    """States far outside fp16 range (the eps sampler from t=999 reaches |x_t| ~ 1e5 with random weights, SURVEY.md
    M8): the per-patch power-of-two range scale keeps fp16 activations in range; patches in one batch are scaled
    independently."""
    B, H, W = 3, 32, 32
    x, cond = _inputs(B, H, W, seed=12)
    x[0] *= 1.0e5
    x[1] *= 40.0
    t = torch.tensor([499, 832, 20], dtype=torch.long)
    xin = torch.cat([x, cond], 1)
    ref = env["oracle"](xin, t)
    got = env["model"](xin.to(env["dev"]), t.to(env["dev"])).cpu()
    for i in range(B):
        rel = float((got[i] - ref[i]).norm() / ref[i].norm())
        assert rel <= 5e-3, (i, rel)
        assert float((got[i] - ref[i]).abs().max()) <= 2e-2 * float(ref[i].abs().max()), i


def test_batch_independent_and_deterministic(env):
    B, H, W = 4, 64, 64
    x, cond = _inputs(B, H, W, seed=5)
    t = torch.tensor([999, 979, 20, 0], dtype=torch.long)
    xin = torch.cat([x, cond], 1).to(env["dev"])
    y1 = env["model"](xin, t.to(env["dev"])).clone()
    y2 = env["model"](xin, t.to(env["dev"])).clone()
    assert torch.equal(y1, y2)
    for i in range(B):
        yi = env["model"](xin[i:i + 1], t[i:i + 1].to(env["dev"]))
        assert torch.equal(yi[0], y1[i]), i


def test_drop_in_surface(env):
    import s1s2_b200
    m = env["model"]
    assert m.outc.out_channels == 4
    assert list(m.state_dict().keys()) == list(ounet.param_shapes(8, 4, 96).keys())
    with pytest.raises(s1s2_b200.S1S2Error):
        m(torch.zeros(1, 8, 32, 32, device="cuda")[:, :, :, :].cpu(), torch.zeros(1, dtype=torch.long))   # CPU tensor
    bad = dict(env["sd"])
    bad.pop("outc.bias")
    with pytest.raises(RuntimeError):
        s1s2_b200.UNetSmallB200(8, 4, 96).load_state_dict(bad, strict=True)
    with pytest.raises(s1s2_b200.S1S2Error):
        s1s2_b200.UNetSmallB200(8, 4, 48).to(env["dev"])(torch.zeros(1, 8, 32, 32, device="cuda"),
                                                      torch.zeros(1, dtype=torch.long, device="cuda"))
    with pytest.raises(s1s2_b200.S1S2Error):
        m(torch.zeros(1, 8, 30, 32, device="cuda"), torch.zeros(1, dtype=torch.long, device="cuda"))     # H % 16


# ------------------------------------------------------------------------------------------------ samplers
def _ref_update(st, x_in, pred, z):
    """The reference's elementwise scheduler code for one step record (fp32 torch, reference operation order)."""
    from s1s2_b200 import _lib
    c0, c1, c2, c3, c4 = (torch.tensor(v, dtype=torch.float32) for v in (st.c0, st.c1, st.c2, st.c3, st.c4))
    if st.kind == _lib.STEP_EPS_DDIM:
        x0, e = (x_in - c0 * pred) / c1, pred
    elif st.kind in (_lib.STEP_V_DDIM, _lib.STEP_V_DDPM):
        x0, e = c0 * x_in - c1 * pred, c1 * x_in + c0 * pred
    else:
        x0, e = None, pred
    ddpm = st.kind in (_lib.STEP_EPS_DDPM, _lib.STEP_V_DDPM)
    xn = c2 * (x_in - c3 * e) if ddpm else c2 * x0 + c3 * e
    if st.flags & _lib.STEP_NOISE:
        xn = xn + c4 * z
    if st.flags & _lib.STEP_FINAL:
        xn = torch.clamp(xn if ddpm else x0, 0.0, 1.0)
    return xn


def _teacher_forced(env, steps, cond, x_init, init_scale=1.0, step_noise=None):
    from s1s2_b200 import _lib, samplers
    dev = env["dev"]
    out, taps = samplers.run_steps(env["model"], steps, cond.to(dev), x_init.to(dev), init_scale=init_scale,
                                   step_noise=None if step_noise is None else step_noise.to(dev),
                                   tap_pred=True, tap_x=True)
    torch.cuda.synchronize()
    B = cond.shape[0]
    x_in = x_init * torch.tensor(init_scale, dtype=torch.float32)
    worst = 0.0
    for i, st in enumerate(steps):
        ref = env["oracle"](torch.cat([x_in, cond], 1), torch.full((B,), st.t, dtype=torch.long))
        got = taps["pred"][i].cpu()
        rel = float((got - ref).norm() / ref.norm())
        worst = max(worst, rel)
        assert rel <= 5e-3, (i, st.t, rel)
        assert float((got - ref).abs().max()) <= 2e-2 * float(ref.abs().max()), (i, st.t)
        z = step_noise[st.noise_index] if st.flags & _lib.STEP_NOISE else None
        want = _ref_update(st, x_in, got, z)
        x_out = taps["x"][i].cpu()
        assert torch.equal(x_out, want), (i, st.t, float((x_out - want).abs().max()))
        x_in = x_out
    assert torch.equal(out.cpu(), x_in)
    return worst


def test_eps_ddim_grid_a_teacher_forced(env):
    from s1s2_b200 import schedule
    x, cond = _inputs(2, 32, 32, seed=21)
    _teacher_forced(env, schedule.steps_eps_grid_a(env["abar"], 999, 6), cond, x)
    _teacher_forced(env, schedule.steps_eps_grid_a(env["abar"], 200, 5), cond, x)


def test_eps_ddim50_config2_teacher_forced_full_size(env):
    """BASELINE config 2 at full size: eps-prediction DDIM-50 on grid A from t = 999 (Evaluation_Pure_Generation.py:277-292),
    one 256x256 patch.  The chain is chaotic in fp32 itself (1/sqrt(abar_999) = 8970 on the first step), so
    parity is teacher-forced: every one of the 50 model calls is compared with the fp32 oracle on OUR state (rel-L2 <=
    5e-3, max-abs <= 2e-2 max|ref|, with |x_t| reaching ~1e5 early on) and every scheduler update must be bit-exact."""
    from s1s2_b200 import schedule
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    x, cond = _inputs(1, 256, 256, seed=502)
    worst = _teacher_forced(env, schedule.steps_eps_grid_a(env["abar"], 999, 50), cond, x)
    print(f"[config 2 parity] 50 teacher-forced model calls at 256x256: worst per-step eps rel-L2 {worst:.2e}")


def test_v_ddim_grid_b_teacher_forced(env):
    from s1s2_b200 import schedule
    x, cond = _inputs(2, 32, 32, seed=22)
    ab = env["abar"]
    _teacher_forced(env, schedule.steps_grid_b(ab, schedule.grid_b(999, 6), "v"), cond, x,
                    init_scale=float(torch.sqrt(1 - ab[999])))


def test_stochastic_chains_teacher_forced(env):
    from s1s2_b200 import schedule
    x, cond = _inputs(2, 32, 32, seed=23)
    ab = env["abar"]
    g = torch.Generator().manual_seed(99)
    zs = torch.randn((8, 2, 4, 32, 32), generator=g)
    _teacher_forced(env, schedule.steps_grid_b(ab, schedule.grid_b(999, 5), "v", eta=0.05), cond, x,
                    init_scale=float(torch.sqrt(1 - ab[999])), step_noise=zs)
    b16 = osched.cosine_betas(16)
    a16 = 1 - b16
    ab16 = torch.cumprod(a16, 0)
    for param in ("eps", "v"):
        _teacher_forced(env, schedule.steps_ddpm(b16, a16, ab16, param, t_list=[15, 9, 3, 1, 0]), cond, x, step_noise=zs)


def test_v_ddim_free_running_image_agreement(env):
    from s1s2_b200 import samplers
    B, H, W = 2, 32, 32
    x, cond = _inputs(B, H, W, seed=31)
    ab = env["abar"]
    ref = osamplers.ddim_v_grid_b(env["oracle"], cond, ab, x, 10)
    got = samplers.sample_ddim_v(env["model"], cond.to(env["dev"]), ab, 4, steps=10, eta=0.0, noise=x.to(env["dev"])).cpu()
    assert ometrics.psnr(got, ref) >= 40.0
    tgt = torch.rand((B, 4, H, W), generator=torch.Generator().manual_seed(3))
    assert abs(ometrics.psnr(got, tgt) - ometrics.psnr(ref, tgt)) <= 0.1
    assert abs(ometrics.ssim_simple(got, tgt) - ometrics.ssim_simple(ref, tgt)) <= 0.002


@pytest.mark.parametrize("seed", [501, 502, 503])
def test_v_ddim50_headline_config_free_running(env, seed):
    """The headline configuration itself (BASELINE config 3 / north-star target): DDIM-50 of the v-prediction model from
    t = 999 on a 256x256 patch, eta = 0, same weights / conditioning / supplied initial noise -- free-running against the
    fp32 CPU oracle, three seeds.  Stated tolerance: per-step predicted-v rel-L2 <= 5e-3 and max-abs <= 2e-2 max|ref|
    (teacher-forced on the oracle's own trajectory, EVERY one of the 50 steps), final image PSNR(ours, oracle) >= 40 dB,
    PSNR / SSIM against a target within 0.1 dB / 0.002."""
    from s1s2_b200 import samplers
    dev = env["dev"]
    x, cond = _inputs(1, 256, 256, seed=seed)
    ab = env["abar"]
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    trace = []
    ref = osamplers.ddim_v_grid_b(env["oracle"], cond, ab, x, 50, trace=trace)
    assert len(trace) == 50
    got = samplers.sample_ddim_v(env["model"], cond.to(dev), ab, 4, steps=50, eta=0.0, noise=x.to(dev)).cpu()
    assert ometrics.psnr(got, ref) >= 40.0, ometrics.psnr(got, ref)
    tgt = torch.rand((1, 4, 256, 256), generator=torch.Generator().manual_seed(9))
    d_psnr = abs(ometrics.psnr(got, tgt) - ometrics.psnr(ref, tgt))
    d_ssim = abs(ometrics.ssim_simple(got, tgt) - ometrics.ssim_simple(ref, tgt))
    assert d_psnr <= 0.1 and d_ssim <= 0.002
    worst = (0.0, 0.0)
    for rec in trace:                                        # per-step predicted v, teacher-forced on the oracle's states
        t = torch.full((1,), rec["t"], dtype=torch.long)
        v = env["model"](torch.cat([rec["x_in"], cond], 1).to(dev), t.to(dev)).cpu()
        rel = float((v - rec["pred"]).norm() / rec["pred"].norm())
        mabs = float((v - rec["pred"]).abs().max()) / float(rec["pred"].abs().max())
        assert rel <= 5e-3 and mabs <= 2e-2, (rec["t"], rel, mabs)
        worst = (max(worst[0], rel), max(worst[1], mabs))
    print(f"[headline parity seed {seed}] PSNR(ours, oracle) {ometrics.psnr(got, ref):.2f} dB, max|ours - oracle| "
          f"{float((got - ref).abs().max()):.2e}, dPSNR vs target {d_psnr:.2e} dB, dSSIM {d_ssim:.2e}, per-step v over all 50 "
          f"steps: worst rel-L2 {worst[0]:.2e}, worst max-abs/max|ref| {worst[1]:.2e}")


def test_eps_recon_free_running_image_agreement(env):
    from s1s2_b200 import samplers
    x, cond = _inputs(1, 32, 32, seed=32)
    ab = env["abar"]
    x_gt = torch.rand((1, 4, 32, 32), generator=torch.Generator().manual_seed(4))
    x_init = osamplers.noise_gt(x_gt, ab, 200, x)
    ref = osamplers.ddim_eps_grid_a(env["oracle"], cond, ab, x_init, 200, 20)
    mask = torch.ones((1, 32, 32))
    mae, mse, got = samplers.ddim_multistep_eval(env["model"], x_gt.to(env["dev"]), cond.to(env["dev"]), ab,
                                                 mask.to(env["dev"]), t_start=200, steps=20, noise=x.to(env["dev"]))
    got = got.cpu()
    assert ometrics.psnr(got, ref) >= 40.0
    assert abs(ometrics.psnr(got, x_gt) - ometrics.psnr(ref, x_gt)) <= 0.1
    assert abs(mae - ometrics.masked_mae(ref, x_gt)) <= 1e-3


def test_sample_host_matches_device_entry(env):
    from s1s2_b200 import samplers, schedule
    x, cond = _inputs(2, 32, 32, seed=41)
    steps = schedule.steps_grid_b(env["abar"], schedule.grid_b(999, 4), "v")
    s = float(torch.sqrt(1 - env["abar"][999]))
    a = samplers.run_steps(env["model"], steps, cond.to(env["dev"]), x.to(env["dev"]), init_scale=s).cpu()
    b = samplers.run_steps_host(env["model"], steps, cond.pin_memory(), x.pin_memory(), init_scale=s, device=env["dev"])
    assert torch.equal(a, b)


def test_full_size_batch_properties(env):
    """BASELINE-size property checks (256x256): clamp range, determinism, independence of patches in a batch."""
    from s1s2_b200 import samplers, schedule
    B = 4
    x, cond = _inputs(B, 256, 256, seed=51)
    ab = env["abar"]
    steps = schedule.steps_grid_b(ab, schedule.grid_b(999, 3), "v")
    s = float(torch.sqrt(1 - ab[999]))
    y = samplers.run_steps(env["model"], steps, cond.to(env["dev"]), x.to(env["dev"]), init_scale=s).clone()
    y2 = samplers.run_steps(env["model"], steps, cond.to(env["dev"]), x.to(env["dev"]), init_scale=s).clone()
    assert torch.equal(y, y2)
    assert float(y.min()) >= 0.0 and float(y.max()) <= 1.0 and torch.isfinite(y).all()
    y1 = samplers.run_steps(env["model"], steps, cond[2:3].to(env["dev"]), x[2:3].to(env["dev"]), init_scale=s)
    assert torch.equal(y1[0], y[2])


def test_max_batch_64_matches_batch_1(env):
    """BASELINE config 3's batch (64 patches of 256x256, the largest single-GPU configuration): every patch of the batch
    equals the same patch sampled alone (bit-exact), phantom tiles and the last partial cluster wave included."""
    import s1s2_b200
    from s1s2_b200 import samplers, schedule
    dev = env["dev"]
    big = s1s2_b200.UNetSmallB200(8, 4, 96, max_batch=64).to(dev)
    big.load_state_dict(env["sd"], strict=True)
    big.eval()
    x, cond = _inputs(64, 256, 256, seed=77)
    ab = env["abar"]
    steps = schedule.steps_grid_b(ab, schedule.grid_b(999, 2), "v")
    s = float(torch.sqrt(1 - ab[999]))
    y = samplers.run_steps(big, steps, cond.to(dev), x.to(dev), init_scale=s).clone()
    assert torch.isfinite(y).all() and float(y.min()) >= 0.0 and float(y.max()) <= 1.0
    for k in (0, 31, 63):
        y1 = samplers.run_steps(big, steps, cond[k:k + 1].to(dev), x[k:k + 1].to(dev), init_scale=s)
        assert torch.equal(y1[0], y[k]), k
    y61 = samplers.run_steps(big, steps, cond[:61].to(dev), x[:61].to(dev), init_scale=s)      # odd batch: phantom tiles
    assert torch.equal(y61, y[:61])
    del big


# ------------------------------------------------------------------------------------------------ patch I/O
def test_tile_extract_matches_golden(env):
    from s1s2_b200 import patch
    z = np.load(os.path.join(G, "patch.npz"))
    ps, st = (int(v) for v in z["norm/ps_stride"])
    scene = z["norm/scene"]
    org = patch.tile_origins(scene.shape[1], scene.shape[2], ps, st)
    assert np.array_equal(org, z["norm/origins"])
    cond, mask, ratio = patch.tile_extract(torch.from_numpy(scene).to(env["dev"]), org, ps)
    assert np.array_equal(mask.cpu().numpy(), z["norm/mask"])
    assert np.abs(cond.cpu().numpy() - z["norm/cond"]).max() <= 2e-6
    assert np.allclose(ratio.cpu().numpy(), z["norm/mask"].reshape(len(org), -1).mean(1), atol=1e-7)


def test_tile_extract_edge_cases(env):
    from s1s2_b200 import patch
    rng = np.random.default_rng(3)
    scene = rng.normal(-12, 4, (4, 70, 90)).astype(np.float32)
    scene[:, :40, :40] = np.nan                     # one window fully invalid
    scene[0, 38:, 57:] = 5.0                        # window (38,57): constant channel -> sigma < 1e-6 -> 1
    scene[1, 45, 70] = np.inf
    vm = opatch.valid_mask(scene)
    org = opatch.tile_origins(70, 90, 32, 19)
    cond, mask, _ = patch.tile_extract(torch.from_numpy(scene).to(env["dev"]), org, 32)
    for i, (r, c) in enumerate(org):
        X, M, _ = opatch.extract_patch(scene, vm, int(r), int(c), 32)
        assert np.array_equal(mask[i].cpu().numpy(), M), i
        assert np.abs(cond[i].cpu().numpy() - X).max() <= 2e-6, i
    c0, m0, _ = patch.tile_extract(torch.from_numpy(scene).to(env["dev"]), np.zeros((0, 2), np.int32), 32)
    assert c0.shape == (0, 4, 32, 32) and m0.shape == (0, 32, 32)


@pytest.mark.parametrize("H,W,ps,st", [(40, 56, 16, 8), (70, 93, 32, 19), (64, 64, 64, 64)])
def test_stitch_bitexact_vs_oracle(env, H, W, ps, st):
    from s1s2_b200 import patch
    rng = np.random.default_rng(H)
    org = opatch.tile_origins(H, W, ps, st)
    preds = rng.random((len(org), 4, ps, ps)).astype(np.float32)
    keep = np.ones(len(org), bool)
    if len(org) > 3:
        keep[[1, len(org) // 2]] = False            # skipped patches are simply absent
    ref, cov = opatch.stitch(preds[keep], org[keep], H, W)
    canvas, cover = patch.stitch(torch.from_numpy(preds[keep]).to(env["dev"]), org[keep], ps, st, H, W)
    assert np.array_equal(cover.cpu().numpy(), cov)
    assert np.array_equal(canvas.cpu().numpy(), ref)


def test_extract_sample_stitch_roundtrip(env):
    """Size-independent property: identical overlaps blend to themselves (stitch o extract == identity on covered
    pixels for the un-normalised channels)."""
    from s1s2_b200 import patch
    H, W, ps, st = 512, 768, 256, 64
    g = torch.Generator().manual_seed(8)
    truth = torch.rand((4, H, W), generator=g).to(env["dev"])
    org = patch.tile_origins(H, W, ps, st)
    tiles = torch.stack([truth[:, r:r + ps, c:c + ps] for r, c in org])
    canvas, cover = patch.stitch(tiles, org, ps, st, H, W)
    assert bool(cover.all())
    assert float((canvas - truth).abs().max()) <= 1e-6


def test_patch_kernels_vector_and_scalar_paths_agree(env):
    """The patch I/O kernels move float4 / uchar4 when the geometry is 4-element aligned and scalars otherwise.  The same
    windows embedded one column to the right (unaligned origins and row pitch -> scalar path) must give the same masks,
    decisions and (to summation-order rounding) the same values; metrics with an odd pixel count and C = 5 take the
    scalar 8-channel path and are checked against the oracle's definitions."""
    from s1s2_b200 import metrics, patch
    dev = env["dev"]
    rng = np.random.default_rng(11)
    H, W, ps, st = 96, 128, 32, 16
    scene = rng.normal(-12, 4, (4, H, W)).astype(np.float32)
    scene[rng.random((4, H, W)) < 0.02] = np.nan
    target = rng.random((4, H, W)).astype(np.float32)
    target[:, :40, :40] *= 0.05                                            # a dark, flat corner: exercises codes 2 / 3
    org = patch.tile_origins(H, W, ps, st)
    sc2 = np.full((4, H, W + 3), np.nan, np.float32); sc2[:, :, 1:W + 1] = scene
    tg2 = np.zeros((4, H, W + 3), np.float32); tg2[:, :, 1:W + 1] = target
    org2 = org + np.array([0, 1], np.int32)
    c_a, m_a, r_a = patch.tile_extract(torch.from_numpy(scene).to(dev), org, ps)
    c_b, m_b, r_b = patch.tile_extract(torch.from_numpy(sc2).to(dev), org2, ps)
    assert torch.equal(m_a, m_b) and torch.equal(r_a, r_b)
    assert float((c_a - c_b).abs().max()) <= 1e-6
    f_a = patch.tile_filter(torch.from_numpy(scene).to(dev), torch.from_numpy(target).to(dev), org, ps).cpu().numpy()
    f_b = patch.tile_filter(torch.from_numpy(sc2).to(dev), torch.from_numpy(tg2).to(dev), org2, ps).cpu().numpy()
    assert np.array_equal(f_a[:, 7], f_b[:, 7]) and len(set(f_a[:, 7])) > 1
    assert np.allclose(f_a[:, :7], f_b[:, :7], rtol=1e-5, atol=1e-9, equal_nan=True)
    gen = torch.Generator().manual_seed(23)
    pred, tgt = torch.rand((3, 5, 7, 9), generator=gen), torch.rand((3, 5, 7, 9), generator=gen)
    mask = (torch.rand((3, 7, 9), generator=gen) > 0.3).float()
    got = metrics.patch_metrics(pred.to(dev), tgt.to(dev), mask.to(dev)).cpu().numpy()
    for i in range(3):
        a, b, mk = pred[i:i + 1], tgt[i:i + 1], mask[i:i + 1]
        want = [ometrics.masked_mae(a, b, mk), ometrics.masked_mse(a, b, mk), ometrics.psnr(a, b, mk),
                ometrics.ssim_simple(a, b), ometrics.sam(a, b, mk), ometrics.ergas(a, b, mk)]
        assert np.allclose(got[i, :6], want, rtol=5e-6, atol=1e-7), (i, got[i], want)


# ------------------------------------------------------------------------------------------------ scene + drivers
def test_scene_pipeline_single_rank(env):
    """extract -> sample -> stitch composed by s1s2_b200.scene equals the pieces run by hand; stitch of the returned
    patches is bit-exact against the oracle's numpy stitch."""
    from s1s2_b200 import patch, scene as sc
    scn = sc.synthetic_scene(96, 160, seed=5, nan_fraction=0.03)
    scn[:, :40, :48] = float("nan")
    ab = env["abar"]
    res = sc.generate_scene(env["model"], scn.to(env["dev"]), ab, ps=32, stride=16, param="v", steps=3, batch=4,
                            valid_ratio_threshold=0.5)
    org = patch.tile_origins(96, 160, 32, 16)
    assert np.array_equal(res["origins"], org)
    vm = opatch.valid_mask(scn.numpy())
    want_keep = np.array([vm[r:r + 32, c:c + 32].mean() >= 0.5 for r, c in org])
    assert np.array_equal(res["kept"], want_keep) and 0 < want_keep.sum() < len(org)
    ref_canvas, ref_cover = opatch.stitch(res["preds"].cpu().numpy(), org[want_keep], 96, 160)
    assert np.array_equal(res["canvas"].cpu().numpy(), ref_canvas)
    assert np.array_equal(res["cover"].cpu().numpy(), ref_cover)
    # one patch by hand: same conditioning, same keyed noise, same sampler
    k = int(np.nonzero(want_keep)[0][3])
    cond, _, _ = patch.tile_extract(scn.to(env["dev"]), org[k:k + 1], 32)
    z = sc.patch_noise([k], (4, 32, 32), 1234, env["dev"])
    y = sc.sample_patches(env["model"], cond, ab, z, param="v", steps=3, batch=1)
    assert torch.equal(y[0], res["preds"][3])


def test_drivers_write_reference_outputs(env, tmp_path):
    from s1s2_b200 import drivers
    pdir, odir = tmp_path / "patches", tmp_path / "out"
    pdir.mkdir()
    rng = np.random.default_rng(0)
    for i in range(5):
        np.savez_compressed(pdir / f"patch_{i:06d}.npz", inputs=rng.normal(size=(4, 32, 32)).astype(np.float32),
                            target=rng.random((4, 32, 32)).astype(np.float32), mask=(rng.random((32, 32)) > 0.1).astype(np.uint8))
    ckpt = tmp_path / "ddpm_s1_to_s2_v3.pth"
    torch.save({"model": env["sd"]}, ckpt)                    # wrapped form (DDIM_Multi-step_v_Prediction.py:264-270)
    common = ["--patch_dir", str(pdir), "--ckpt", str(ckpt), "--batch", "2"]
    drivers.main(["ddim", "--out_dir", str(odir / "a"), "--t_start", "200", "--ddim_steps", "3"] + common)
    rows = list(__import__("csv").reader(open(odir / "a" / "ddim_metrics.csv")))
    assert rows[0] == ["file", "t_start", "ddim_steps", "MAE", "MSE"] and len(rows) == 6
    assert open(odir / "a" / "ddim_summary.txt").read().startswith("files: 5  t_start: 200  steps: 3\n")
    drivers.main(["ddim_v", "--out_dir", str(odir / "b"), "--t_start", "999", "--ddim_steps", "3"] + common)
    assert list(__import__("csv").reader(open(odir / "b" / "ddim_metrics.csv")))[0][3] == "eta"
    drivers.main(["ddim_sweep", "--out_dir", str(odir / "c"), "--t_start_grid", "200,100", "--ddim_steps_grid", "2,3"] + common)
    rows = list(__import__("csv").reader(open(odir / "c" / "ddim_sweep_summary.csv")))
    assert rows[0][:7] == ["t_start", "steps", "files", "MAE_mean", "MAE_std", "MSE_mean", "MSE_std"] and len(rows) == 5
    drivers.main(["true_infer", "--out_dir", str(odir / "d"), "--t_start", "999", "--ddim_steps", "3", "--n_seeds", "2"] + common)
    rows = list(__import__("csv").reader(open(odir / "d" / "ddim_true_infer_metrics.csv")))
    assert rows[0][-3:] == ["PSNR_mean", "SAM_mean", "ERGAS_mean"] and len(rows) == 6
    drivers.main(["onestep", "--out_dir", str(odir / "e"), "--t_small", "20"] + common)
    # Evaluation_Pure_Generation.py --mode night_demo: generation without a target, first --save_viz_n files; .npy instead of PNG
    drivers.main(["night_demo", "--out_dir", str(odir / "n"), "--t_start", "999", "--ddim_steps", "3", "--save_viz_n", "3"] + common)
    night = sorted(os.listdir(odir / "n" / "viz"))
    assert night == [f"{i:03d}_night_{k}.npy" for i in range(3) for k in ("cond", "pred")]
    pred = np.load(odir / "n" / "viz" / "001_night_pred.npy")
    assert pred.shape == (4, 32, 32) and pred.min() >= 0.0 and pred.max() <= 1.0
    assert np.array_equal(np.load(odir / "n" / "viz" / "001_night_cond.npy"), np.load(pdir / "patch_000001.npz")["inputs"])
    # a file without a 'mask' key counts as all-valid; the masks of the other files of its batch stay in force
    mdir = tmp_path / "mixed"
    mdir.mkdir()
    for i in range(2):
        d = dict(np.load(pdir / f"patch_{i:06d}.npz"))
        if i == 1:
            d.pop("mask")
        np.savez_compressed(mdir / f"patch_{i:06d}.npz", **d)
    mixed = ["--patch_dir", str(mdir), "--ckpt", str(ckpt), "--batch", "2", "--t_start_grid", "200", "--ddim_steps_grid", "3"]
    drivers.main(["ddim_sweep", "--out_dir", str(odir / "m2")] + mixed)
    for i in range(2):                                      # the same two files one at a time: same per-file seeds and noise
        one = tmp_path / f"one{i}"
        one.mkdir()
        os.link(mdir / f"patch_{i:06d}.npz", one / f"patch_{i:06d}.npz")
    rows2 = list(__import__("csv").reader(open(odir / "m2" / "ddim_sweep_summary.csv")))
    drivers.main(["ddim_sweep", "--out_dir", str(odir / "m0"), "--patch_dir", str(tmp_path / "one0"), "--ckpt", str(ckpt), "--batch", "1",
                  "--t_start_grid", "200", "--ddim_steps_grid", "3"])
    drivers.main(["ddim_sweep", "--out_dir", str(odir / "m1"), "--patch_dir", str(tmp_path / "one1"), "--ckpt", str(ckpt), "--batch", "1",
                  "--t_start_grid", "200", "--ddim_steps_grid", "3", "--seed_base", "1235"])
    r0 = list(__import__("csv").reader(open(odir / "m0" / "ddim_sweep_summary.csv")))[1]
    r1 = list(__import__("csv").reader(open(odir / "m1" / "ddim_sweep_summary.csv")))[1]
    assert abs(float(rows2[1][3]) - 0.5 * (float(r0[3]) + float(r1[3]))) <= 2e-6       # MAE_mean of the mixed batch
    # per-patch results do not depend on the batch size
    drivers.main(["ddim_sweep", "--out_dir", str(odir / "f"), "--t_start_grid", "200", "--ddim_steps_grid", "3",
                  "--patch_dir", str(pdir), "--ckpt", str(ckpt), "--batch", "1"])
    drivers.main(["ddim_sweep", "--out_dir", str(odir / "g"), "--t_start_grid", "200", "--ddim_steps_grid", "3",
                  "--patch_dir", str(pdir), "--ckpt", str(ckpt), "--batch", "4"])
    a = list(__import__("csv").reader(open(odir / "f" / "ddim_sweep_summary.csv")))[1][:7]
    b = list(__import__("csv").reader(open(odir / "g" / "ddim_sweep_summary.csv")))[1][:7]
    assert a == b
    # Limitation_Test.py run_eval: batched DDIM, dataset-level pixel-weighted report; the report must equal the oracle's
    # aggregation (Limitation_Test.py:118-159) of the predictions the driver saved
    lim = ["--patch_dir", str(pdir), "--ckpt", str(ckpt), "--batch_size", "2", "--save_n", "16"]
    drivers.main(["limitation", "--out_dir", str(odir / "h"), "--mode", "ddim", "--ddim_steps", "3", "--band_weights", "1", "1", "1",
                  "2", "--partial_reverse_k", "2"] + lim)
    txt = open(odir / "h" / "limitation_summary.txt").read()
    assert "==== Unweighted (equal-channel) ====" in txt and "==== Weighted (band_weights) ====" in txt and "[partial-reverse k=2]" in txt
    tot = None
    for i in range(5):
        d = np.load(pdir / f"patch_{i:06d}.npz")
        pred = torch.from_numpy(np.load(odir / "h" / f"ddim_{i // 2:04d}_{i % 2:02d}_pred.npy"))[None]
        gt = torch.from_numpy(np.load(odir / "h" / f"ddim_{i // 2:04d}_{i % 2:02d}_gt.npy"))[None]
        assert np.array_equal(gt[0].numpy(), d["target"])
        sums = ometrics.channelwise_error_sums(pred, gt, torch.from_numpy(d["mask"].astype(np.float32))[None])
        tot = sums if tot is None else tuple(a_ + b_ for a_, b_ in zip(tot, sums))
    mae, mse, ps, mae_c, _, _ = ometrics.aggregate_final(*tot)
    assert f"MAE:  {mae:.6f}" in txt and f"MSE:  {mse:.6f}" in txt and f"PSNR: {ps:.3f} dB" in txt
    assert f" B8:  MAE={mae_c[3]:.6f}" in txt
    # the v script's run_eval in DDPM mode (in-kernel noise), short schedule
    drivers.main(["limitation", "--out_dir", str(odir / "i"), "--mode", "ddpm", "--param", "v", "--T", "16"] + lim)
    assert (odir / "i" / "ddpm_0002_00_pred.npy").exists() and "-- Per-channel metrics (pixel-weighted) --" in open(odir / "i" / "limitation_summary.txt").read()


def test_fused_patch_metrics_match_oracle(env):
    """s1s2_patch_metrics vs the oracle's (reference-pinned) metric definitions, incl. the golden vector the reference
    itself produced; tolerance 2e-6 relative (different summation order, fp64 vs torch fp32 sums)."""
    from s1s2_b200 import metrics
    z = np.load(os.path.join(G, "samplers.npz"))
    p, g, m = (torch.from_numpy(z[k]) for k in ("metrics/pred", "x_gt", "metrics/mask"))
    got = metrics.patch_metrics(p.to(env["dev"]), g.to(env["dev"]), m.to(env["dev"])).cpu().numpy()[0]
    want = z["metrics/vals"]                      # mae, mse, psnr, ssim_simple, sam, ergas from the reference's code
    assert np.allclose(got[:6], want, rtol=5e-6, atol=1e-7), (got[:6], want)
    gen = torch.Generator().manual_seed(17)
    N, C, H, W = 5, 4, 64, 48
    pred = torch.rand((N, C, H, W), generator=gen)
    tgt = torch.rand((N, C, H, W), generator=gen)
    mask = (torch.rand((N, H, W), generator=gen) > 0.2).float()
    mask[3] = 1.0
    got = metrics.patch_metrics(pred.to(env["dev"]), tgt.to(env["dev"]), mask.to(env["dev"])).cpu().numpy()
    for i in range(N):
        a, b, mk = pred[i:i + 1], tgt[i:i + 1], mask[i:i + 1]
        want = [ometrics.masked_mae(a, b, mk), ometrics.masked_mse(a, b, mk), ometrics.psnr(a, b, mk),
                ometrics.ssim_simple(a, b), ometrics.sam(a, b, mk), ometrics.ergas(a, b, mk)]
        assert np.allclose(got[i, :6], want, rtol=5e-6, atol=1e-7), (i, got[i], want)
        assert got[i, 6] == float(mk.sum())
    nomask = metrics.patch_metrics(pred.to(env["dev"]), tgt.to(env["dev"])).cpu().numpy()
    assert np.allclose(nomask[0, :2], [ometrics.masked_mae(pred[:1], tgt[:1]), ometrics.masked_mse(pred[:1], tgt[:1])], rtol=5e-6)
    identical = metrics.patch_metrics(pred.to(env["dev"]), pred.to(env["dev"])).cpu().numpy()
    assert identical[0, 2] == 99.0 and identical[0, 0] == 0.0          # psnr's mse <= 1e-12 branch


def test_pixel_weighted_aggregation_matches_golden(env):
    """metrics.channelwise_error_sums / aggregate_final (per-channel sums from the fused metrics kernel) vs the values
    Limitation_Test.py:118-159 produced (tests/golden/metrics_agg.npz): three batches accumulated like run_eval does.
    fp64 sums vs torch fp32 sums: 2e-6 relative."""
    from s1s2_b200 import metrics
    z = np.load(os.path.join(G, "metrics_agg.npz"))
    dev = env["dev"]
    tot = None
    for b in range(3):
        mask = torch.from_numpy(z[f"agg/mask{b}"]).to(dev) if f"agg/mask{b}" in z.files else None
        a, q, w = metrics.channelwise_error_sums(torch.from_numpy(z[f"agg/pred{b}"]).to(dev), torch.from_numpy(z[f"agg/tgt{b}"]).to(dev), mask)
        assert np.allclose(a.cpu().numpy(), z[f"agg/abs{b}"], rtol=2e-6) and np.allclose(q.cpu().numpy(), z[f"agg/sq{b}"], rtol=2e-6)
        assert float(w) == float(z[f"agg/w{b}"])
        tot = (a, q, w) if tot is None else (tot[0] + a, tot[1] + q, tot[2] + w)
    for tag, bw in (("eq", None), ("bw", [1.0, 1.0, 1.0, 2.0])):
        mae, mse, ps, mae_c, mse_c, ps_c = metrics.aggregate_final(*tot, bw)
        assert np.allclose([mae, mse, ps], z[f"agg/final_{tag}"], rtol=2e-6)
        assert np.allclose(np.stack([mae_c, mse_c, ps_c]), z[f"agg/final_{tag}_c"], rtol=2e-6)


def test_quality_filters_match_golden(env):
    """s1s2_tile_filter vs values the reference's own Patch.py functions produced (tests/golden/filters.npz): decision
    codes exact, statistics within 1e-4 relative (fp64 one-pass sums vs numpy float32 pairwise sums)."""
    from s1s2_b200 import patch
    z = np.load(os.path.join(G, "filters.npz"))
    ps, st = (int(v) for v in z["filt/ps_stride"])
    rows = z["filt/rows"]
    org = rows[:, :2].astype(np.int32)
    stats = patch.tile_filter(torch.from_numpy(z["filt/inputs"]).to(env["dev"]), torch.from_numpy(z["filt/target"]).to(env["dev"]),
                              org, ps, colloc=torch.from_numpy(z["filt/colloc"]).to(env["dev"])).cpu().numpy()
    assert np.array_equal(stats[:, 7].astype(int), rows[:, 2].astype(int))
    assert set(stats[:, 7].astype(int)) == {0, 1, 2, 3, 4}
    assert np.allclose(stats[:, 0], rows[:, 3], rtol=0, atol=1e-7)                      # valid ratio
    assert np.allclose(stats[:, 1:5], rows[:, 4:8], rtol=1e-4, atol=1e-9)               # band variances
    assert np.allclose(stats[:, 5], rows[:, 8], rtol=0, atol=1e-7)                      # dark fraction
    assert np.allclose(stats[:, 6], rows[:, 9], rtol=1e-4, atol=1e-10)                  # Laplacian variance
    # a window with no valid pixel: ratio 0, dark fraction 1, texture 0 (Patch.py:93-94,114), code 1
    scene = torch.full((4, 40, 40), float("nan"))
    s0 = patch.tile_filter(scene.to(env["dev"]), torch.rand(4, 40, 40).to(env["dev"]), np.array([[4, 4]], np.int32), 32).cpu().numpy()[0]
    assert s0[0] == 0.0 and s0[5] == 1.0 and s0[6] == 0.0 and int(s0[7]) == 1


def test_in_kernel_philox_noise(env):
    """S1S2_STEP_PHILOX: the z of a stochastic step generated inside the head kernel.  A DDPM step with c2 = 0, c4 = 1
    returns z itself: unit-normal statistics, determined by (seed, patch id, step index, pixel) only - independent of the
    batch a patch sits in - and different for different seeds / steps.  Parity with the reference's torch generator is
    statistical by construction."""
    from s1s2_b200 import _lib, samplers, schedule
    B, H, W = 2, 64, 64
    x, cond = _inputs(B, H, W, seed=61)
    st = [_lib.Step(500, _lib.STEP_EPS_DDPM, _lib.STEP_PHILOX, 0, 0.0, 0.0, 0.0, 0.0, 1.0),
          _lib.Step(400, _lib.STEP_EPS_DDPM, _lib.STEP_PHILOX, 1, 0.0, 0.0, 0.0, 0.0, 1.0)]
    dev = env["dev"]
    _, taps = samplers.run_steps(env["model"], st, cond.to(dev), x.to(dev), tap_x=True, seed=1234)
    z = taps["x"].cpu()                                     # [2 steps, B, 4, H, W]
    n = z[0].numel()
    assert abs(float(z.mean())) < 4.0 / (2 * n) ** 0.5 and abs(float(z.var()) - 1.0) < 0.03
    assert abs(float((z ** 4).mean()) - 3.0) < 0.25         # kurtosis of a normal
    assert float(z.abs().max()) < 6.5 and torch.isfinite(z).all()
    assert abs(float((z[0] * z[1]).mean())) < 0.02          # steps are independent streams
    assert abs(float((z[0, 0] * z[0, 1]).mean())) < 0.03    # patches are independent streams
    for c in range(3):
        assert abs(float((z[0, :, c] * z[0, :, c + 1]).mean())) < 0.03      # channels (Box-Muller pairs) uncorrelated
    _, taps2 = samplers.run_steps(env["model"], st, cond.to(dev), x.to(dev), tap_x=True, seed=1234)
    assert torch.equal(taps2["x"].cpu(), z)                 # same seed, same noise
    _, taps3 = samplers.run_steps(env["model"], st, cond.to(dev), x.to(dev), tap_x=True, seed=99)
    assert not torch.equal(taps3["x"].cpu(), z)
    # patch 1 alone with patch_base = 1 draws what it drew as slot 1 of the batch
    _, taps4 = samplers.run_steps(env["model"], st, cond[1:2].to(dev), x[1:2].to(dev), tap_x=True, seed=1234, patch_base=1)
    assert torch.equal(taps4["x"].cpu()[:, 0], z[:, 1])
    # a whole ancestral chain without supplied noise runs on it and stays in range
    b16 = osched.cosine_betas(16)
    y = samplers.ddpm_sample(env["model"], cond.to(dev), b16, 1 - b16, torch.cumprod(1 - b16, 0), 4, noise=x.to(dev), seed=5)
    assert torch.isfinite(y).all() and float(y.min()) >= 0.0 and float(y.max()) <= 1.0
    y2 = samplers.ddpm_sample(env["model"], cond.to(dev), b16, 1 - b16, torch.cumprod(1 - b16, 0), 4, noise=x.to(dev), seed=5)
    assert torch.equal(y, y2)


def test_patch_files_match_patch_py_format(env, tmp_path):
    """patchio.write_patches: the .npz files Patch.py would write (keys, dtypes, order, filter decisions, contents) from
    in-memory rasters, checked against the oracle's restatement window by window; the drivers read them back."""
    from s1s2_b200 import drivers, patchio
    z = np.load(os.path.join(G, "filters.npz"))
    ps, st = (int(v) for v in z["filt/ps_stride"])
    inputs, target, colloc = z["filt/inputs"], z["filt/target"], z["filt/colloc"]
    dev = env["dev"]
    entries, counters = patchio.write_patches(torch.from_numpy(inputs).to(dev), torch.from_numpy(target).to(dev), str(tmp_path),
                                              patch_size=ps, stride=st, colloc=torch.from_numpy(colloc).to(dev), folder="S1S2_x")
    rows = z["filt/rows"]
    codes = rows[:, 2].astype(int)
    assert counters == dict(validratio_skipped=int((codes == 1).sum()), var_skipped=int((codes == 2).sum()),
                            dark_skipped=int((codes == 3).sum()), texture_skipped=int((codes == 4).sum()))
    kept = rows[codes == 0]
    assert len(entries) == len(kept)
    M_all = opatch.valid_mask(inputs, target, colloc)
    for k, e in enumerate(entries):
        d = np.load(os.path.join(tmp_path, e["npz"]))
        assert set(d.files) == {"inputs", "target", "mask", "folder", "row", "col", "transform", "crs", "patch_size", "stride",
                                "valid_ratio"}
        r, c = int(kept[k, 0]), int(kept[k, 1])
        assert (int(d["row"]), int(d["col"]), int(d["patch_size"]), int(d["stride"])) == (r, c, ps, st)
        X, M, vr = opatch.extract_patch(inputs, M_all, r, c, ps)
        assert d["mask"].dtype == np.uint8 and np.array_equal(d["mask"], M)
        assert d["inputs"].dtype == np.float32 and np.abs(d["inputs"] - X).max() <= 2e-6
        Y = target[:, r:r + ps, c:c + ps].copy()
        Y[:, ~M.astype(bool)] = 0.0
        assert np.array_equal(d["target"], np.nan_to_num(Y).astype(np.float32))
        assert abs(float(d["valid_ratio"]) - vr) <= 1e-7
    doc = patchio.write_manifest(str(tmp_path), entries, counters, patch_size=ps, stride=st)
    assert doc["total_patches"] == len(kept) and doc["patches"][0]["patch_id"] == "000000"
    x_cond, x_gt, mask, Cc, Ct = drivers.load_npz_as_tensors(os.path.join(tmp_path, entries[0]["npz"]), dev)
    assert (Cc, Ct) == (4, 4) and tuple(x_cond.shape) == (1, 4, ps, ps) and mask is not None
