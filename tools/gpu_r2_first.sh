#!/bin/bash
# Round-2 first GPU pass: parity tests, latency workload (with / without the small-batch tilings), Cout-on-N A/B for conv1.x.
set -u
T=${1:-r2a}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/smi_$T.txt
timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 > $O/pytest_$T.log; echo "pytest rc=$? $(tail -1 $O/pytest_$T.log)"
timeout 300 python bench.py --workload latency > $O/bench_latency_$T.json 2> $O/bench_latency_$T.err; echo "latency rc=$?"
S1S2_NO_ALTS=1 timeout 300 python bench.py --workload latency --no-library-baseline > $O/bench_latency_noalts_$T.json 2>> $O/bench_latency_$T.err; echo "latency noalts rc=$?"
timeout 300 python bench.py --quick --steps 5 > $O/bench_v64_quick_$T.json 2> $O/bench_v64_quick_$T.err; echo "v64 quick rc=$?"
S1S2_C10_UMMA=1 timeout 300 python bench.py --quick --steps 5 > $O/bench_v64_c10umma_$T.json 2>> $O/bench_v64_quick_$T.err; echo "c10 rc=$?"
S1S2_C12_UMMA=1 timeout 300 python bench.py --quick --steps 5 > $O/bench_v64_c12umma_$T.json 2>> $O/bench_v64_quick_$T.err; echo "c12 rc=$?"
du -sh $O
