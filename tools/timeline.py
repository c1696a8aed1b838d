"""Where the time of a small-batch model call goes, launch by launch (run under gpurun).

Needs the timeline build of the library (never the shipped one):
    nvcc ... -DS1S2_TIMELINE -o ab_libs/lib_timeline.so s1s2_lib.cu        (tools/build_timeline.sh)
    S1S2_LIB=ab_libs/lib_timeline.so python tools/timeline.py [B] [H]

Runs a few model calls back to back exactly like the sampling loop (programmatic dependent launches) and prints, for every
launch of the last call, the %globaltimer stamps of CTA 0's roles relative to the grid's first CTA entry, plus the gap
to the previous grid: how much of a launch is launch hand-over, prologue, first round trip to memory, MMA work, epilogue
and drain.  Times in microseconds.
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "s1-to-s2_super-resolution_project-code_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import s1s2_b200  # noqa: E402
from s1s2_b200 import _lib  # noqa: E402

SLOTS = ["entry", "prologue", "dep.wait", "loads req", "act landed", "wgt landed", "last MMA", "1st acc", "stores iss", "drained", "cta0 done"]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    hw = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    assert "S1S2_LIB" in os.environ, "point S1S2_LIB at the -DS1S2_TIMELINE build"
    dev = torch.device("cuda:0")
    m = s1s2_b200.UNetSmallB200(8, 4, 96, max_batch=B).to(dev)
    m.load_state_dict({k: v.to(dev) for k, v in s1s2_b200.synthetic_checkpoint(1235).items()})
    x = torch.randn((B, 8, hw, hw), device=dev)
    m(x, torch.full((B,), 500, device=dev))
    eng = m.engine(dev, hw, hw, B)
    L = _lib.lib()
    fn = L.s1s2_debug_timeline
    fn.restype, fn.argtypes = C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.c_void_p]
    nl = 16
    buf = (C.c_uint64 * (nl * 16))()
    _lib.check(fn(eng.h, B, 6, buf, None), eng.h)
    rows = [[int(buf[l * 16 + k]) for k in range(16)] for l in range(nl)]
    print(f"# batch {B}, {hw} x {hw}: last of 6 back-to-back model calls; us relative to the grid's first CTA entry (slot 12)")
    print(f"{'layer':10s} {'gap':>6s} {'grid':>7s} | " + " ".join(f"{s:>10s}" for s in SLOTS))
    prev_end, t_first, tot_gap = None, rows[0][12], 0.0
    for l in range(nl):
        r = rows[l]
        g0, g1 = r[12], r[13]
        gap = (g0 - prev_end) / 1e3 if prev_end is not None else 0.0
        tot_gap += gap
        cells = " ".join(f"{(r[k] - g0) / 1e3:10.2f}" if r[k] else f"{'-':>10s}" for k in range(11))
        print(f"{L.s1s2_layer_name(eng.h, l).decode():10s} {gap:6.2f} {(g1 - g0) / 1e3:7.2f} | {cells}")
        prev_end = g1
    print(f"# model call: {(rows[-1][13] - t_first) / 1e3:.1f} us from the first entry to the last exit; "
          f"sum of grid-to-grid gaps {tot_gap:.1f} us (negative gap = the next grid's CTAs entered before the last CTA of this one left)")


if __name__ == "__main__":
    main()
