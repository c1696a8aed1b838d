"""Per-layer isolated references for the CUDA denoiser (test infrastructure).

Each layer of libs1s2_b200 is checked on ITS OWN input as produced by the library (the fp16 activation of the
previous layer, read back exactly through s1s2_debug_activation), against the oracle's arithmetic
(oracle/unet.py: conv3x3 pad 1 + bias + ReLU, 2x2 max-pool, ConvTranspose2d k2 s2, channel cat, 1x1 head) evaluated
in float64 with the weights rounded to fp16 like the library's repack does.  What remains is accumulation order
(fp32 in TMEM) and the fp16 rounding of the stored output, so the bound is tight: a wrong descriptor, tap, swizzle
or channel offset shows up as O(1) error in exactly one layer.
"""
import torch
import torch.nn.functional as F


def h16(w):
    return w.to(torch.float16).to(torch.float64)


# (tap name of the output, reference op, input tap names, state_dict prefix)
LAYERS = [
    ("inc", "inc", ("xin16",), "inc.0"),
    ("down1.0", "c3", ("inc",), "down1.0.0"),
    ("down1", "c3p", ("down1.0",), "down1.0.2"),
    ("down2.0", "c3", ("down1",), "down2.0.0"),
    ("down2", "c3p", ("down2.0",), "down2.0.2"),
    ("down3.0", "c3", ("down2",), "down3.0.0"),
    ("down3", "c3p", ("down3.0",), "down3.0.2"),
    ("up3", "up", ("down3",), "up3"),
    ("conv3.0", "c3", ("up3", "down2"), "conv3.0"),
    ("conv3", "c3", ("conv3.0",), "conv3.2"),
    ("up2", "up", ("conv3",), "up2"),
    ("conv2.0", "c3", ("up2", "down1"), "conv2.0"),
    ("conv2", "c3", ("conv2.0",), "conv2.2"),
    ("up1", "up", ("conv2",), "up1"),
    ("conv1.0", "c3", ("up1", "inc"), "conv1.0"),
    ("out", "head", ("conv1.0",), "conv1.2"),
]


def layer_reference(sd, kind, prefix, inputs):
    """float64 reference of one layer on the given (already fp16-valued) inputs."""
    x = torch.cat([i.to(torch.float64) for i in inputs], dim=1)
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"].to(torch.float64)
    if kind == "inc":
        # pixel record: [xlo0..3 | t t 0 0 | c0..c3 | xhi0..3], x = 4096*hi + lo; the t weight is an fp16 hi+lo pair
        rec = x
        xin = torch.cat([4096.0 * rec[:, 12:16] + rec[:, 0:4], rec[:, 8:12], rec[:, 4:5]], dim=1)
        w64 = h16(w)
        hi = w[:, 8].to(torch.float16).to(torch.float32)
        lo = (w[:, 8] - hi).to(torch.float16).to(torch.float32)
        w64[:, 8] = hi.to(torch.float64) + lo.to(torch.float64)
        return F.relu(F.conv2d(xin, w64, b, padding=1))
    if kind == "c3":
        return F.relu(F.conv2d(x, h16(w), b, padding=1))
    if kind == "c3p":
        return F.max_pool2d(F.relu(F.conv2d(x, h16(w), b, padding=1)), 2)
    if kind == "up":
        return F.conv_transpose2d(x, h16(w), b, stride=2)
    if kind == "head":
        hmid = F.relu(F.conv2d(x, h16(w), b, padding=1))
        return F.conv2d(hmid, sd["outc.weight"].to(torch.float64), sd["outc.bias"].to(torch.float64))
    raise ValueError(kind)


def check_layers(model, sd, y, B, verbose=False):
    """Returns [(name, max_abs_err, max_abs_ref, rel_l2)] for every layer; `y` is the forward's output."""
    acts = {}

    def act(name):
        if name not in acts:
            acts[name] = model.activation(name, B).cpu()
        return acts[name]

    rows = []
    for name, kind, ins, prefix in LAYERS:
        ref = layer_reference(sd, kind, prefix, [act(i) for i in ins])
        got = (y.cpu() if name == "out" else act(name)).to(torch.float64)
        err = (got - ref).abs().max().item()
        mag = ref.abs().max().item()
        rel = ((got - ref).norm() / ref.norm().clamp_min(1e-30)).item()
        rows.append((name, err, mag, rel))
        if verbose:
            print(f"  {name:10s} max|err| {err:.3e}  max|ref| {mag:.3e}  rel-L2 {rel:.3e}", flush=True)
    return rows
