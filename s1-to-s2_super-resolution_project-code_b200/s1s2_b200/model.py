"""Drop-in for the reference's ``UNetSmall`` whose forward runs in libs1s2_b200 (sm_100a tcgen05 kernels).

Boundary being mirrored: ``UNetSmall(in_ch, out_ch, base_ch)`` -- constructor
Evaluation/DDIM_Multi-step.py:21-40, call ``model(torch.cat([x_t, x_cond], 1), t_idx)`` :42-53, checkpoint
loading ``model.load_state_dict(torch.load(ckpt), strict=True); model.eval()``
(Evaluation/DDIM_Multi-step_v_Prediction.py:263-271) and ``model.outc.out_channels``
(Evaluation_Updated/Evaluation_Pure_Generation.py:280).

The module owns the same 34 float32 parameters under the same names, created in the same order (so a fixed
``torch.manual_seed`` yields the reference's random init); they are repacked into the library's fp16 K-major
layout whenever they change.  The forward pass never touches PyTorch operators: without the CUDA library on an
sm_100 device it raises.
"""
import ctypes as C
from collections import OrderedDict

import torch
import torch.nn as nn

from . import _lib

# (attribute, kind, cin multiple, cout multiple) in the reference's construction order; multiples of base_ch.
_BLOCKS = (("down1", 1, 2), ("down2", 2, 4), ("down3", 4, 8))
_DECODER = (("up3", "conv3", 8, 4), ("up2", "conv2", 4, 2), ("up1", "conv1", 2, 1))


def _pair(cin, cout):
    return nn.Sequential(nn.Conv2d(cin, cout, 3, padding=1), nn.ReLU(), nn.Conv2d(cout, cout, 3, padding=1), nn.ReLU())


class _Engine:
    """One library handle: arena for (H, W, max_batch) on one device + the weights it was loaded with."""

    def __init__(self, device_index, H, W, max_batch, base_ch=96):
        self.h = C.c_void_p()
        self.key = (device_index, H, W)
        self.max_batch = max_batch
        self.weights_version = None
        _lib.check(_lib.lib().s1s2_create(C.byref(self.h), device_index, 8, 4, base_ch, H, W, max_batch))

    def close(self):
        if self.h:
            _lib.lib().s1s2_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def synthetic_checkpoint(seed: int, in_ch: int = 8, out_ch: int = 4, base_ch: int = 96):
    """state_dict of a freshly initialised network under ``torch.manual_seed(seed)`` -- the stand-in for the reference's
    ``Models/*.pth`` blobs, which are absent from its tree.  The module creates its parameters in the
    reference's order with PyTorch's default Conv init, i.e. what ``torch.manual_seed(seed); UNetSmall(8, 4, 96)`` yields."""
    with torch.random.fork_rng(devices=[]):
        torch.manual_seed(int(seed))
        m = UNetSmallB200(in_ch, out_ch, base_ch)
    return OrderedDict((k, v.detach().clone()) for k, v in m.state_dict().items())


class UNetSmallB200(nn.Module):
    def __init__(self, in_ch: int, out_ch: int, base_ch: int = 96, max_batch: int = 16):
        super().__init__()
        b = base_ch
        self.in_ch, self.out_ch, self.base_ch = in_ch, out_ch, base_ch
        self.inc = nn.Sequential(nn.Conv2d(in_ch + 1, b, 3, padding=1), nn.ReLU())
        for name, ci, co in _BLOCKS:
            setattr(self, name, nn.Sequential(_pair(b * ci, b * co), nn.MaxPool2d(2)))
        for up, blk, ci, co in _DECODER:
            setattr(self, up, nn.ConvTranspose2d(b * ci, b * co, 2, stride=2))
            setattr(self, blk, _pair(b * ci, b * co))
        self.outc = nn.Conv2d(b, out_ch, 1)
        self.requires_grad_(False)
        self.max_batch = int(max_batch)
        self._engines = OrderedDict()
        self._weights_epoch = 0

    # ------------------------------------------------------------------ library plumbing
    def _weights_version(self):
        return (self._weights_epoch,) + tuple((p.data_ptr(), p._version) for p in self.parameters())

    def invalidate_weights(self):
        """Force the next call to repack the parameters into the library.  Needed after writes the autograd version
        counter does not see -- ``p.data.copy_(...)``, ``p.data.mul_(...)``, EMA swaps through ``.data`` -- which would
        otherwise leave the engine sampling with the previous fp16 copy.  ``load_state_dict``, ``.to()`` / ``.cuda()`` /
        ``.float()`` and ordinary in-place parameter updates are detected automatically."""
        self._weights_epoch += 1

    def load_state_dict(self, *args, **kwargs):
        res = super().load_state_dict(*args, **kwargs)
        self.invalidate_weights()
        return res

    def _apply(self, fn, *args, **kwargs):
        res = super()._apply(fn, *args, **kwargs)
        self.invalidate_weights()
        return res

    def engine(self, device: torch.device, H: int, W: int, batch: int) -> "_Engine":
        """Handle for this geometry, (re)created when the batch outgrows it and (re)loaded when parameters change."""
        if device.type != "cuda":
            raise _lib.S1S2Error("UNetSmallB200 runs only on a CUDA (sm_100a) device; there is no CPU fallback. "
                                 "Move the model and its inputs to 'cuda'.")
        if (self.in_ch, self.out_ch) != (8, 4) or self.base_ch not in (64, 96):
            raise _lib.S1S2Error("libs1s2_b200 implements UNetSmall(in_ch=8, out_ch=4, base_ch=96 | 64) only "
                                 f"(got {self.in_ch}, {self.out_ch}, {self.base_ch})")
        idx = device.index if device.index is not None else torch.cuda.current_device()
        key = (idx, H, W)
        eng = self._engines.get(key)
        if eng is not None and eng.max_batch < batch:
            eng.close()
            eng = None
        if eng is None:
            eng = _Engine(idx, H, W, max(batch, self.max_batch), self.base_ch)
            self._engines[key] = eng
        self._engines.move_to_end(key)           # activation() / saturation_counts() read the engine used last
        ver = self._weights_version()
        if eng.weights_version != ver:
            sd = self.state_dict()
            for k, v in sd.items():
                if v.device.type != "cuda" or v.device.index != idx or v.dtype != torch.float32:
                    raise _lib.S1S2Error(f"parameter {k} lives on {v.device} ({v.dtype}); call model.to('cuda:{idx}') "
                                         "and keep float32 parameters (the library repacks them to fp16 itself)")
            keep = [v.contiguous() for v in sd.values()]
            n = len(sd)
            names = (C.c_char_p * n)(*[k.encode() for k in sd.keys()])
            ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in keep])
            numel = (C.c_int64 * n)(*[t.numel() for t in keep])
            stream = torch.cuda.current_stream(device).cuda_stream
            _lib.check(_lib.lib().s1s2_load_weights(eng.h, n, names, ptrs, numel, C.c_void_p(stream)), eng.h)
            eng.weights_version = ver
        return eng

    def launch_count(self) -> int:
        return sum(int(_lib.lib().s1s2_launch_count(e.h)) for e in self._engines.values())

    def profile_layers(self, device, H: int, W: int, batch: int, reps: int = 3):
        """[(state_dict prefix, mean ms per launch)] of one model call at this geometry (s1s2_profile_layers)."""
        eng = self.engine(torch.device(device), H, W, batch)
        L = _lib.lib()
        n = C.c_int()
        _lib.check(L.s1s2_profile_layers(eng.h, batch, reps, None, 0, C.byref(n), None), eng.h)
        ms = (C.c_float * n.value)()
        stream = torch.cuda.current_stream(torch.device(device)).cuda_stream
        _lib.check(L.s1s2_profile_layers(eng.h, batch, reps, ms, n.value, C.byref(n), C.c_void_p(stream)), eng.h)
        return [(L.s1s2_layer_name(eng.h, i).decode(), float(ms[i])) for i in range(n.value)]

    def tile_widths(self, device, H: int, W: int, batch: int):
        """[(state_dict prefix, GEMM-N tile width)] the launches of a model call use at this batch (s1s2_debug_tile_width)."""
        eng = self.engine(torch.device(device), H, W, batch)
        L = _lib.lib()
        out, i = [], 0
        while True:
            name = L.s1s2_layer_name(eng.h, i)
            if name is None:
                return out
            out.append((name.decode(), int(L.s1s2_debug_tile_width(eng.h, i, batch))))
            i += 1

    def loop_layer(self, device, H: int, W: int, batch: int, layer: int, reps: int, perf_mode: int = 0) -> float:
        """Mean ms per launch of one layer run alone `reps` times (s1s2_debug_loop_layer; measurement aid)."""
        eng = self.engine(torch.device(device), H, W, batch)
        ms = C.c_float()
        stream = torch.cuda.current_stream(torch.device(device)).cuda_stream
        _lib.check(_lib.lib().s1s2_debug_loop_layer(eng.h, batch, layer, reps, perf_mode, C.byref(ms), C.c_void_p(stream)), eng.h)
        return float(ms.value)

    def saturation_counts(self, batch: int, device=None):
        """{activation name: number of fp16 outputs of the LAST model call that sit at the epilogues' saturation value
        (|v| = 65504) or are not finite} (s1s2_debug_saturation_count; debug aid for trained-scale checkpoints)."""
        eng = next(reversed(self._engines.values()))
        L = _lib.lib()
        n = C.c_int()
        _lib.check(L.s1s2_debug_saturation_count(eng.h, batch, None, 0, C.byref(n), None), eng.h)
        counts = (C.c_uint64 * n.value)()
        stream = torch.cuda.current_stream(torch.device("cuda", eng.key[0])).cuda_stream
        _lib.check(L.s1s2_debug_saturation_count(eng.h, batch, counts, n.value, C.byref(n), C.c_void_p(stream)), eng.h)
        return {L.s1s2_view_name(eng.h, i).decode(): int(counts[i]) for i in range(n.value)}

    # ------------------------------------------------------------------ the reference's call
    @torch.no_grad()
    def forward(self, xt_and_cond: torch.Tensor, t_idx: torch.Tensor) -> torch.Tensor:
        if xt_and_cond.ndim != 4 or xt_and_cond.shape[1] != self.in_ch:
            raise ValueError(f"expected f32[B,{self.in_ch},H,W], got {tuple(xt_and_cond.shape)}")
        B, _, H, W = xt_and_cond.shape
        dev = xt_and_cond.device
        eng = self.engine(dev, H, W, B)
        x = xt_and_cond.to(torch.float32).contiguous()
        t = t_idx.to(device=dev, dtype=torch.int64).reshape(-1)
        if t.numel() == 1 and B > 1:
            t = t.expand(B)
        t = t.contiguous()
        if t.numel() != B:
            raise ValueError(f"t_idx must hold {B} timesteps, got {t.numel()}")
        out = torch.empty((B, self.out_ch, H, W), device=dev, dtype=torch.float32)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(_lib.lib().s1s2_forward(eng.h, x.data_ptr(), t.data_ptr(), out.data_ptr(), B, C.c_void_p(stream)), eng.h)
        return out

    @torch.no_grad()
    def activation(self, name: str, batch: int, device=None) -> torch.Tensor:
        """Per-layer parity tap (s1s2_debug_activation): activation `name` of the last call as f32 NCHW."""
        eng = next(reversed(self._engines.values()))
        idx = eng.key[0]
        c, hh, ww = C.c_int(), C.c_int(), C.c_int()
        L = _lib.lib()
        _lib.check(L.s1s2_debug_activation(eng.h, name.encode(), None, batch, C.byref(c), C.byref(hh), C.byref(ww), None), eng.h)
        out = torch.empty((batch, c.value, hh.value, ww.value), device=f"cuda:{idx}", dtype=torch.float32)
        stream = torch.cuda.current_stream(out.device).cuda_stream
        _lib.check(L.s1s2_debug_activation(eng.h, name.encode(), out.data_ptr(), batch, None, None, None, C.c_void_p(stream)), eng.h)
        return out
