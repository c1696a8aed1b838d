#!/bin/bash
# Perf iteration pass: parity tests, fixed cost per launch, latency table, batch-64 quick line.   usage: tools/gpu_r2_perf.sh TAG [notests]
set -u
T=${1:-r2x}
O=gpurun_out
mkdir -p $O
if [ "${2:-}" != "notests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > $O/pytest_$T.log; echo "pytest rc=$? $(tail -1 $O/pytest_$T.log)"
fi
python tools/diag.py overhead > $O/overhead_$T.log 2>&1; echo "overhead rc=$?"; cat $O/overhead_$T.log
timeout 300 python bench.py --workload latency --no-library-baseline > $O/bench_latency_$T.json 2> $O/bench_latency_$T.err; echo "latency rc=$?"
timeout 300 python bench.py --quick --steps 5 > $O/bench_v64_quick_$T.json 2> $O/bench_v64_quick_$T.err; echo "v64 quick rc=$?"
python - <<PY
import json
d=json.loads(open("$O/bench_latency_$T.json").read().strip().splitlines()[-1])
for r in d["latency"]:
    print("B=%2d fused %.4f ms/call %7.2f p/s frac %.3f | dropin %.4f ms/call %7.2f p/s" % (r["batch"], r["fused"]["ms_per_model_call"], r["fused"]["patches_per_s"], r["fused"]["frac_of_peak"], r["dropin"]["ms_per_model_call"], r["dropin"]["patches_per_s"]))
d=json.loads(open("$O/bench_v64_quick_$T.json").read().strip().splitlines()[-1])
print("v64", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), d["clocks"]["sm_mhz"], " ".join("%s:%.3f" % (r["layer"], r["ms"]) for r in d["roofline"]["layers"]))
PY
