#!/bin/bash
# Round-2 third GPU pass: parity tests, ncu launch list of the bench command, ncu --set full per-layer tables at batch 64 and 1.
set -u
T=${1:-r2c}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -60 > $O/pytest_$T.log; echo "pytest rc=$? $(tail -1 $O/pytest_$T.log)"
python bench.py --quick --steps 2 --warmup 3 > $O/bench_quick_$T.json 2> $O/bench_quick_$T.err; echo "quick rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file $O/launches_$T.csv \
    python bench.py --quick --steps 1 --warmup 3 --no-layers > $O/ncu_launches_$T.log 2>&1; echo "ncu launches rc=$?"
python tools/diag.py time 64 2 > $O/diag_time_$T.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_ -s 32 -c 16 -o $O/prof_b64_$T python tools/diag.py time 64 2 > $O/ncu_full_b64_$T.log 2>&1; echo "ncu full b64 rc=$?"
python tools/ncu_summary.py $O/prof_b64_$T.ncu-rep $O/ncu_full_per_layer_b64_$T.txt --layers --batch 64 --traffic $O/traffic_$T.json \
    --title "ncu --set full --clock-control none -k regex:conv_ -s 32 -c 16 python tools/diag.py time 64 2 (one model call, B=64, 256x256; serialised, cold)" > /dev/null
python tools/diag.py time 1 2 >> $O/diag_time_$T.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_ -s 32 -c 16 -o $O/prof_b1_$T python tools/diag.py time 1 2 > $O/ncu_full_b1_$T.log 2>&1; echo "ncu full b1 rc=$?"
python tools/ncu_summary.py $O/prof_b1_$T.ncu-rep $O/ncu_full_per_layer_b1_$T.txt --layers --batch 1 \
    --title "ncu --set full --clock-control none -k regex:conv_ -s 32 -c 16 python tools/diag.py time 1 2 (one model call, B=1, 256x256; serialised, cold)" > /dev/null
ncu -i $O/prof_b1_$T.ncu-rep --page raw --csv > $O/prof_b1_${T}_raw.csv 2>/dev/null
ncu -i $O/prof_b64_$T.ncu-rep --page raw --csv > $O/prof_b64_${T}_raw.csv 2>/dev/null
rm -f $O/prof_b64_$T.ncu-rep            # gpurun brings back at most 64 MiB: keep the batch-1 report only
python tools/hbm_kernels.py $O/hbm_kernels_$T.json > $O/hbm_kernels_$T.log 2>&1; echo "hbm rc=$?"
ls -la $O | tail -20; du -sh $O
