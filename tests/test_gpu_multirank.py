"""The N > 1 path on real GPUs: two NCCL ranks (one process per GPU) run s1s2_b200.scene.generate_scene with the CUDA
kernels -- tile extract, keyed noise, fused sampler, NCCL gather, stitch -- and must reproduce the one-rank result bit for
bit (pure patch sharding: SURVEY.md section 8e).  Skipped on a one-GPU box; the CPU twin with stand-in stages is
tests/test_scene_sharding.py."""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "s1-to-s2_super-resolution_project-code_b200")

CFG = dict(ps=64, stride=32, param="v", steps=4, t_start=999, batch=5, valid_ratio_threshold=0.5)


def _scene():
    from s1s2_b200 import scene as sc
    scn = sc.synthetic_scene(192, 320, seed=11, nan_fraction=0.03)
    scn[:, :70, :90] = float("nan")                      # some windows fall under the valid-ratio threshold
    return scn


def _run(rank, world, port, out_path):
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    import s1s2_b200
    from s1s2_b200 import scene as sc, schedule
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    if world > 1:
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=dev)
    sd = s1s2_b200.synthetic_checkpoint(1235)
    model = s1s2_b200.UNetSmallB200(8, 4, 96, max_batch=CFG["batch"]).to(dev)
    model.load_state_dict({k: v.to(dev) for k, v in sd.items()}, strict=True)
    model.eval()
    _, _, abar = schedule.derive(schedule.cosine_beta_schedule(1000))
    res = sc.generate_scene(model, _scene().to(dev), abar, rank=rank, world=world, **CFG)
    for window in ("hann",):
        res_w = sc.generate_scene(model, _scene().to(dev), abar, rank=rank, world=world, window=window, **CFG)
    torch.cuda.synchronize()
    if rank == 0:
        torch.save({"canvas": res["canvas"].cpu(), "cover": res["cover"].cpu(), "preds": res["preds"].cpu(),
                    "kept": torch.as_tensor(res["kept"]), "canvas_hann": res_w["canvas"].cpu()}, out_path)
    else:
        assert res is None and res_w is None
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_nccl_ranks_reproduce_one_rank_bit_for_bit(tmp_path):
    import torch.multiprocessing as mp
    one, two = str(tmp_path / "w1.pt"), str(tmp_path / "w2.pt")
    mp.spawn(_run, args=(1, 0, one), nprocs=1, join=True)
    mp.spawn(_run, args=(2, _free_port(), two), nprocs=2, join=True)
    a, b = torch.load(one), torch.load(two)
    n = int(a["kept"].sum())
    assert 0 < n < a["kept"].numel() and a["preds"].shape[0] == n
    for k in a:
        assert torch.equal(a[k], b[k]), k
    assert float(a["cover"].float().mean()) > 0.5 and bool(torch.isfinite(a["canvas"]).all())
    assert not torch.equal(a["canvas"], a["canvas_hann"])
