"""Whole-scene generation: Patch.py tiling -> patch-wise sharding over the GPUs of one box -> fused DDIM sampling ->
one gather to rank 0 -> overlap-blend stitch.

Reference call pattern: Evaluation_Updated/Evaluation_Pure_Generation.py:539-574 (`--mode ddim --true_infer`) runs
``ddpm_ddim_generate`` once per ``patch_*.npz`` that Patch.py:195-255 wrote; here the patches come straight from the
scene rasters on the device.  Patches are independent units: rank r owns the contiguous range
``shard_range(N, r, world)`` of the row-major ``patch_iter`` list, no collective touches the data path until the final
gather of the predicted patches (f32[n_r,4,ps,ps]) to rank 0, which stitches.  The result does not depend on the
world size: the initial noise of a patch is keyed by its global index and every kernel is batch-independent.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib, patch, samplers


def patch_noise(indices, shape, seed_base, device, method="philox"):
    """Unit-normal initial noise f32[len(indices), *shape] keyed by GLOBAL patch index, so the draw of a patch does not
    depend on the rank or batch slot it lands in.

    method "philox" (default): one library kernel for all patches (s1s2_patch_noise: Philox4x32-10 keyed by seed_base,
    counter = (element, patch index)).  method "torch": one torch generator per patch seeded seed_base + index -- the
    per-file seeding of Evaluation/DDIM_Sweep.py:193,404, bit-identical to what that script would draw on this device
    (one generator re-seed + one launch per patch: fine for a file list, slow for the 3 249 windows of a stride-32 scene)."""
    n = len(indices)
    out = torch.empty((n,) + tuple(shape), device=device, dtype=torch.float32)
    if n == 0:
        return out
    if method == "torch":
        g = torch.Generator(device=device)
        for k, idx in enumerate(indices):
            g.manual_seed(int(seed_base) + int(idx))
            out[k] = torch.randn(shape, generator=g, device=device, dtype=torch.float32)
        return out
    if method != "philox":
        raise ValueError(f"unknown noise method '{method}'")
    dev = torch.device(device)
    didx = dev.index if dev.index is not None else torch.cuda.current_device()
    ids = torch.as_tensor(np.asarray(indices, dtype=np.int64)).to(dev)
    elems = int(np.prod(shape))
    stream = torch.cuda.current_stream(dev).cuda_stream
    for lo in range(0, n, 65535):
        m = min(65535, n - lo)
        _lib.check(_lib.lib().s1s2_patch_noise(didx, C.c_uint64(int(seed_base) & (2 ** 64 - 1)), C.c_void_p(ids[lo:].data_ptr()), m,
                                               elems, C.c_void_p(out[lo:].data_ptr()), C.c_void_p(stream)))
    return out


def sample_patches(model, cond, alpha_bar, noise, param="v", steps=50, t_start=999, batch=64):
    """DDIM (eta=0) over N patches in batches of `batch`: v -> sample_ddim_v (grid B), eps -> ddpm_ddim_generate (grid A)."""
    outs = []
    for lo in range(0, cond.shape[0], batch):
        c, z = cond[lo:lo + batch], noise[lo:lo + batch]
        if param == "v":
            y = samplers.sample_ddim_v(model, c, alpha_bar, z.shape[1], steps=steps, eta=0.0, t_start=t_start, noise=z)
        elif param == "eps":
            y = samplers.ddpm_ddim_generate(model, c, alpha_bar, t_start=t_start, steps=steps, noise=z)
        else:
            raise ValueError(f"unknown parameterisation '{param}'")
        outs.append(y)
    if not outs:
        return torch.empty((0,) + tuple(noise.shape[1:]), device=cond.device, dtype=torch.float32)
    return torch.cat(outs, 0)


def gather_to_rank0(local: torch.Tensor, counts, rank: int, world: int, group=None):
    """Concatenate per-rank shards (counts[r] rows each, rank order) on rank 0; other ranks get None.  Shards are
    padded to the largest count so that one fixed-size collective moves everything (NCCL gather over NVLink)."""
    if world == 1:
        return local
    import torch.distributed as dist
    cmax = max(counts)
    pad = torch.zeros((cmax,) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, bufs, dst=0, group=group)
    if rank != 0:
        return None
    return torch.cat([bufs[r][:counts[r]] for r in range(world)], 0)


class _PhaseClock:
    """CUDA-event stopwatch over the phases of one generate_scene call (no host synchronisation until read)."""

    def __init__(self, enabled):
        self.enabled = enabled
        self.marks = []

    def mark(self, name):
        if self.enabled:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.marks.append((name, e))

    def read(self, into):
        if not self.enabled or not self.marks:
            return
        torch.cuda.synchronize()
        for (_, e0), (name, e1) in zip(self.marks[:-1], self.marks[1:]):
            into[name + "_ms"] = into.get(name + "_ms", 0.0) + e0.elapsed_time(e1)


def generate_scene(model, scene, alpha_bar, ps=256, stride=64, param="v", steps=50, t_start=999, batch=64,
                   seed_base=1234, vmask=None, valid_ratio_threshold=0.0, rank=0, world=1, group=None,
                   extract_fn=None, sample_fn=None, stitch_fn=None, noise_fn=None, window=None, timings=None):
    """scene f32[4,SH,SW] on this rank's device (HH dB, HV dB, incidence deg, elevation m; NaN = no data).

    Returns on rank 0 a dict(canvas f32[C,SH,SW], cover u8[SH,SW], origins i32[N,2], kept bool[N], preds f32[Nk,C,ps,ps]);
    None on the other ranks.  The *_fn hooks exist for the CPU tests of the sharding logic (gloo, no GPU): the
    product path uses the CUDA kernels and has no fallback.  `window`: None = uniform overlap blend, "hann" = Hann-weighted
    (patch.stitch).  `timings` (dict, CUDA only): receives extract_ms / noise_ms / sample_ms / sync_ms / gather_ms / stitch_ms of
    this rank (sync = waiting for the slowest rank at the keep-flag all-reduce; gather = the transfer itself), measured with CUDA events on the current stream."""
    extract_fn = extract_fn or patch.tile_extract
    sample_fn = sample_fn or sample_patches
    stitch_fn = stitch_fn or patch.stitch
    noise_fn = noise_fn or patch_noise
    clock = _PhaseClock(timings is not None and scene.device.type == "cuda")
    SH, SW = int(scene.shape[1]), int(scene.shape[2])
    origins = patch.tile_origins(SH, SW, ps, stride)
    N = len(origins)
    lo, hi = patch.shard_range(N, rank, world)
    clock.mark("start")
    cond, mask, ratio = extract_fn(scene, origins[lo:hi], ps, vmask)
    keep = ratio >= valid_ratio_threshold if (hi - lo) else torch.zeros((0,), dtype=torch.bool, device=scene.device)
    kept_idx = (torch.nonzero(keep).flatten().cpu().numpy() + lo).astype(np.int64)      # global patch indices
    clock.mark("extract")
    C_tgt = model.outc.out_channels
    noise = noise_fn(kept_idx, (C_tgt, ps, ps), seed_base, scene.device)
    clock.mark("noise")
    preds = sample_fn(model, cond[keep], alpha_bar, noise, param=param, steps=steps, t_start=t_start, batch=batch)
    clock.mark("sample")

    # which patches every rank kept (tiny, host-side): ranks agree on counts before the fixed-size gather
    if world > 1:
        import torch.distributed as dist
        flags = torch.zeros((N,), dtype=torch.uint8, device=scene.device)
        flags[lo:hi] = keep.to(torch.uint8)
        dist.all_reduce(flags, op=dist.ReduceOp.MAX, group=group)
        kept_all = flags.bool().cpu().numpy()
    else:
        kept_all = keep.cpu().numpy().astype(bool)
    clock.mark("sync")                  # (world > 1: the all-reduce is where a fast rank waits for the slowest one)
    counts = [int(kept_all[slice(*patch.shard_range(N, r, world))].sum()) for r in range(world)]
    allp = gather_to_rank0(preds, counts, rank, world, group)
    clock.mark("gather")
    if rank != 0:
        clock.read(timings)
        return None
    if window is None:
        canvas, cover = stitch_fn(allp, origins[kept_all], ps, stride, SH, SW)
    else:
        canvas, cover = stitch_fn(allp, origins[kept_all], ps, stride, SH, SW, window=window)
    clock.mark("stitch")
    clock.read(timings)
    return dict(canvas=canvas, cover=cover, origins=origins, kept=kept_all, preds=allp)


def generate_scene_host(model, scene_host, alpha_bar, device, rank=0, world=1, **kw):
    """generate_scene for a scene raster held in HOST memory (pinned recommended), the way a caller that has just read the
    rasters from disk holds it (Patch.py:152-187): uploads the scene to `device`, runs generate_scene there, and on rank 0
    downloads canvas and cover into pinned host tensors.  Returns (result dict with host `canvas` / `cover`, bytes up,
    bytes down); the other ranks get (None, bytes up, 0)."""
    scene_d = scene_host.to(device, non_blocking=True)
    res = generate_scene(model, scene_d, alpha_bar, rank=rank, world=world, **kw)
    up = scene_host.numel() * scene_host.element_size()
    if res is None:
        return None, up, 0
    canvas = torch.empty(res["canvas"].shape, dtype=torch.float32, pin_memory=True)
    cover = torch.empty(res["cover"].shape, dtype=torch.uint8, pin_memory=True)
    canvas.copy_(res["canvas"], non_blocking=True)
    cover.copy_(res["cover"], non_blocking=True)
    torch.cuda.current_stream(device).synchronize()
    res = dict(res, canvas=canvas, cover=cover)
    return res, up, canvas.numel() * 4 + cover.numel()


def synthetic_scene(SH=2048, SW=2048, seed=0, nan_fraction=0.02):
    """Synthetic Sentinel-1 scene of the shape Patch.py reads: HH ~ N(-12,4) dB,
    HV ~ N(-19,4) dB, incidence U(20,45) deg, smooth elevation 0..1500 m, `nan_fraction` no-data holes."""
    g = torch.Generator().manual_seed(seed)
    hh = torch.randn((SH, SW), generator=g) * 4.0 - 12.0
    hv = torch.randn((SH, SW), generator=g) * 4.0 - 19.0
    inc = torch.rand((SH, SW), generator=g) * 25.0 + 20.0
    yy = torch.linspace(0, 3.0, SH).view(-1, 1)
    xx = torch.linspace(0, 2.0, SW).view(1, -1)
    elev = 750.0 * (1.0 + torch.sin(yy) * torch.cos(xx))
    scene = torch.stack([hh, hv, inc, elev.expand(SH, SW)], 0).contiguous()
    if nan_fraction > 0:
        holes = torch.rand((SH, SW), generator=g) < nan_fraction
        scene[:, holes] = float("nan")
    return scene
