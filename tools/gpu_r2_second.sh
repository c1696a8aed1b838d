#!/bin/bash
# Round-2 second GPU pass: tests, 64B-swizzle experiment, fixed launch overhead, the full default bench line, ncu captures.
set -u
T=${1:-r2b}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 > $O/pytest_$T.log; echo "pytest rc=$? $(tail -1 $O/pytest_$T.log)"
python tools/diag.py overhead > $O/overhead_$T.log 2>&1; echo "overhead rc=$?"
# same layer (conv2.2: 192 -> 192 at 128 x 128), 64-channel chunks / 128B swizzle vs 32-channel chunks / 64B swizzle with three taps per stage
for F in "" "conv2.2=HC96IN" "conv2.2=HC96IN,conv2.0=HC96IN,down2.0.0=HC96IN"; do
  S1S2_FORCE="$F" timeout 300 python bench.py --quick --steps 3 > "$O/bench_force_${T}_$(echo $F | tr '=,.' '___').json" 2>> $O/bench_force_$T.err; echo "force '$F' rc=$?"
done
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_default_$T.json 2> $O/bench_default_$T.err; echo "default rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference_$T.json 2>> $O/bench_default_$T.err; echo "reference rc=$?"
du -sh $O
