"""The N>1 path on CPU: world_size-2 gloo processes run s1s2_b200.scene.generate_scene with the CUDA stages replaced
by the oracle's numpy restatements (tile extraction, stitch) and a deterministic stand-in sampler, and must reproduce
the single-process result bit for bit (patch-wise sharding, per-patch noise keyed by global index, one gather)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "s1-to-s2_super-resolution_project-code_b200")


class _Head:
    out_channels = 4


class _FakeModel:
    outc = _Head()


def _extract(scene, origins, ps, vmask=None):
    from oracle import patch as opatch
    sc = scene.numpy()
    vm = opatch.valid_mask(sc)
    conds, masks, ratios = [], [], []
    for r, c in origins:
        X, M, vr = opatch.extract_patch(sc, vm, int(r), int(c), ps)
        conds.append(X); masks.append(M); ratios.append(vr)
    n = len(origins)
    cond = torch.from_numpy(np.stack(conds)) if n else torch.zeros((0, 4, ps, ps))
    mask = torch.from_numpy(np.stack(masks)) if n else torch.zeros((0, ps, ps), dtype=torch.uint8)
    return cond, mask, torch.tensor(ratios, dtype=torch.float32)


def _sample(model, cond, alpha_bar, noise, param="v", steps=50, t_start=999, batch=64):
    # any deterministic per-patch map built from exactly-rounded elementwise ops (independent of vector width)
    return torch.clamp(cond * 0.5 + noise * 0.25 + float(alpha_bar[3]), 0.0, 1.0)


def _noise(indices, shape, seed_base, device):
    # per-patch torch generator keyed by the global index (the product default is the library's Philox kernel: CUDA only)
    from s1s2_b200 import scene as sc
    return sc.patch_noise(indices, shape, seed_base, device, method="torch")


def _stitch(preds, origins, ps, stride, SH, SW):
    from oracle import patch as opatch
    c, m = opatch.stitch(preds.numpy(), np.asarray(origins), SH, SW)
    return torch.from_numpy(c), torch.from_numpy(m)


def _run(rank, world, port, out_path, thr):
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    from s1s2_b200 import scene as sc
    from oracle import schedule as osched
    if world > 1:
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    _, _, abar = osched.make_schedule(1000)
    scene = sc.synthetic_scene(96, 160, seed=3, nan_fraction=0.05)
    scene[:, :40, :48] = float("nan")                       # a few windows fall under the valid-ratio threshold
    res = sc.generate_scene(_FakeModel(), scene, abar, ps=32, stride=16, batch=3, valid_ratio_threshold=thr, rank=rank,
                            world=world, extract_fn=_extract, sample_fn=_sample, stitch_fn=_stitch, noise_fn=_noise)
    if rank == 0:
        torch.save({k: (v if isinstance(v, torch.Tensor) else torch.as_tensor(v)) for k, v in res.items()}, out_path)
    else:
        assert res is None
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("thr", [0.0, 0.8])
def test_world2_matches_world1(tmp_path, thr):
    one, two = str(tmp_path / "w1.pt"), str(tmp_path / "w2.pt")
    _run(0, 1, 0, one, thr)
    mp.spawn(_run, args=(2, _free_port(), two, thr), nprocs=2, join=True)
    a, b = torch.load(one), torch.load(two)
    assert set(a) == set(b)
    for k in a:
        assert torch.equal(a[k], b[k]), k
    assert 0 < int(a["kept"].sum()) <= a["kept"].numel()
    if thr > 0:
        assert int(a["kept"].sum()) < a["kept"].numel()      # the filter removed windows, shards stay consistent
    assert a["preds"].shape[0] == int(a["kept"].sum())
