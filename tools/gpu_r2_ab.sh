#!/bin/bash
# same-box A/B of two builds of the library (batch 64 quick line + batch 1..16 latency): ab_libs/lib_r2n.so vs the in-tree one
set -u
O=gpurun_out
if [ "${1:-}" = "tests" ]; then timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3; fi
for V in "" "S1S2_LIB=$PWD/ab_libs/lib_r2n.so" "" "S1S2_LIB=$PWD/ab_libs/lib_r2n.so"; do
  env $V python bench.py --quick --steps 4 > $O/ab.json 2>> $O/ab.err
  env $V python bench.py --workload latency --no-library-baseline --no-layers > $O/ab_lat.json 2>> $O/ab.err
  python - "$V" <<PY
import json,sys
d=json.loads(open("$O/ab.json").read().strip().splitlines()[-1])
l=json.loads(open("$O/ab_lat.json").read().strip().splitlines()[-1])
L={r["layer"]:r["ms"] for r in d["roofline"]["layers"]}
print("%-24s v64 %.2f clk %s | d1.0.2 %.3f d2.0.2 %.3f d3.0.2 %.3f c1.0 %.3f d1.0.0 %.3f | lat " % (sys.argv[1][-20:] or "(in-tree)", d["value"], d["clocks"]["sm_mhz"], L["down1.0.2"], L["down2.0.2"], L["down3.0.2"], L["conv1.0"], L["down1.0.0"]) + " ".join("%.2f" % r["fused"]["patches_per_s"] for r in l["latency"]))
PY
done
