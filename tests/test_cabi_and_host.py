"""CPU-side checks: the C-ABI library builds, loads and exports every symbol include/s1s2_b200.h declares (no
compute without a GPU), it fails loudly without a device, and the host logic (schedules, grids, step records,
tiling, sharding) agrees with the oracle / golden vectors."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from oracle import patch as opatch
from oracle import schedule as osched

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from s1s2_b200 import _lib
    return _lib


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "s1s2_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(s1s2_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(lib.SIGNATURES), declared ^ set(lib.SIGNATURES)
    L = lib.lib()
    for name in declared:
        assert getattr(L, name) is not None
    assert L.s1s2_abi_version() == 3


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib):
    import s1s2_b200
    h = C.c_void_p()
    rc = lib.lib().s1s2_create(C.byref(h), 0, 8, 4, 96, 32, 32, 1)
    assert rc != 0 and not h.value
    assert b"no CUDA device" in lib.lib().s1s2_global_error() or b"CPU fallback" in lib.lib().s1s2_global_error()
    m = s1s2_b200.UNetSmallB200(8, 4, 96)
    with pytest.raises(s1s2_b200.S1S2Error):
        m(torch.zeros(1, 8, 32, 32), torch.zeros(1, dtype=torch.long))
    assert lib.lib().s1s2_create(C.byref(h), 0, 8, 4, 48, 32, 32, 1) == lib.ERR_INVALID     # architecture check first


def test_step_struct_layout(lib):
    assert C.sizeof(lib.Step) == 36
    assert [f[0] for f in lib.Step._fields_] == ["t", "kind", "flags", "noise_index", "c0", "c1", "c2", "c3", "c4"]


def test_module_matches_reference_checkpoint_layout():
    import s1s2_b200
    from oracle import unet as ounet
    m = s1s2_b200.UNetSmallB200(8, 4, 96)
    shapes = ounet.param_shapes(8, 4, 96)
    sd = m.state_dict()
    assert list(sd.keys()) == list(shapes.keys())
    assert all(tuple(sd[k].shape) == shapes[k] for k in shapes)
    assert m.outc.out_channels == 4
    assert sum(v.numel() for v in sd.values()) == 17_237_668


def test_schedule_and_grids_match_oracle_and_golden():
    from s1s2_b200 import schedule
    z = np.load(os.path.join(G, "schedule.npz"))
    b, a, ab = schedule.derive(schedule.cosine_beta_schedule(1000))
    assert np.array_equal(b.numpy(), z["cosine_betas"]) and np.array_equal(ab.numpy(), z["cosine_alpha_bar"])
    assert np.array_equal(schedule.make_schedule(1000, "linear").numpy(), z["linear_betas"])
    for k in z.files:
        if k.startswith("gridA_"):
            _, ts, st = k.split("_")
            assert np.array_equal(schedule.grid_a(int(ts), int(st)).numpy(), z[k]), k
        if k.startswith("gridB_"):
            _, K, st = k.split("_")
            assert np.array_equal(schedule.grid_b(int(K), int(st)).numpy(), z[k]), k


def test_step_records():
    from s1s2_b200 import _lib, schedule
    _, alphas, ab = osched.make_schedule(1000)
    st = schedule.steps_eps_grid_a(ab, 999, 50)
    assert [s.t for s in st] == osched.grid_a(999, 50).tolist()[:-1] and len(st) == 50
    assert all(s.kind == _lib.STEP_EPS_DDIM for s in st)
    assert [s.flags for s in st] == [0] * 49 + [_lib.STEP_FINAL]
    assert st[0].c1 == float(torch.sqrt(ab[999] + 1e-8)) and st[0].c2 == float(torch.sqrt(ab[979]))
    assert st[-1].c2 == float(torch.sqrt(ab[0]))
    sv = schedule.steps_grid_b(ab, schedule.grid_b(999, 50), "v")
    assert [s.t for s in sv] == osched.grid_b(999, 50).tolist()[::-1] and sv[-1].t == 0
    assert sv[-1].flags == _lib.STEP_FINAL and all(s.flags == 0 for s in sv[:-1])
    se = schedule.steps_grid_b(ab, schedule.grid_b(999, 12), "v", eta=0.05)
    assert [s.noise_index for s in se] == list(range(11)) + [-1]
    sd = schedule.steps_ddpm(1 - alphas, alphas, ab, "eps")
    assert len(sd) == 1000 and sd[0].t == 999 and sd[-1].t == 0 and sd[-1].flags == _lib.STEP_FINAL
    assert sd[0].flags == _lib.STEP_NOISE and sd[-2].noise_index == 998


def test_tiling_and_sharding_host_logic():
    from s1s2_b200 import patch
    for H, W, ps, st in [(2048, 2048, 256, 64), (300, 420, 64, 32), (255, 400, 256, 32), (700, 513, 256, 100)]:
        assert np.array_equal(patch.tile_origins(H, W, ps, st), opatch.tile_origins(H, W, ps, st))
    for n in (0, 1, 7, 841, 3249):
        for world in (1, 2, 3, 8):
            r = [patch.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n and all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1


def test_driver_batches_prefetch_keeps_order_and_contents(tmp_path):
    """drivers._batches reads the next batch in a background thread; order, grouping and contents must be those of a
    plain serial read (Evaluation/DDIM_Multi-step.py:104-111 conversions)."""
    import types
    import numpy as np
    import torch
    from s1s2_b200 import drivers
    rng = np.random.default_rng(5)
    for i in range(7):
        x = rng.normal(size=(4, 8, 8)).astype(np.float32)
        x[0, 0, 0] = np.nan
        kw = dict(inputs=x, target=rng.random((4, 8, 8)).astype(np.float32))
        if i != 3:
            kw["mask"] = (rng.random((8, 8)) > 0.2).astype(np.uint8)
        np.savez_compressed(tmp_path / f"patch_{i:06d}.npz", **kw)
    files = sorted(f for f in os.listdir(tmp_path) if f.endswith(".npz"))
    args = types.SimpleNamespace(patch_dir=str(tmp_path), batch=3)
    seen = []
    for lo, names, cond, gt, mask in drivers._batches(args, files, torch.device("cpu")):
        assert names == files[lo:lo + 3] and cond.shape[0] == len(names) == len(mask)
        for j, f in enumerate(names):
            c, g, m, _, _ = drivers.load_npz_as_tensors(os.path.join(tmp_path, f), torch.device("cpu"))
            assert torch.equal(cond[j:j + 1], c) and torch.equal(gt[j:j + 1], g) and not torch.isnan(cond).any()
            assert (m is None and mask[j] is None) or torch.equal(mask[j], m)
        seen += names
    assert seen == files
