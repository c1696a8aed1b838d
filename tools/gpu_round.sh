#!/bin/bash
# One-GPU measurement pass of a round (run under gpurun from the repo root): parity tests, the bench lines of every
# BASELINE config, the ncu launch list and the ncu --set full captures summarised under profiles/.
# usage: bash tools/gpu_round.sh TAG        (outputs: gpurun_out/*_TAG.*)
set -u
T=${1:-rX}
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > $O/bench_v64_$T.json 2> $O/bench_v64_$T.err; echo "v64 rc=$?"
python bench.py --workload eps16 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_eps16_$T.json 2>> $O/bench_v64_$T.err; echo "eps16 rc=$?"
python bench.py --workload sweep --steps 3 --warmup 3 > $O/bench_sweep_$T.json 2>> $O/bench_v64_$T.err; echo "sweep rc=$?"
python bench.py --workload scene --steps 1 --warmup 3 > $O/bench_scene1_$T.json 2>> $O/bench_v64_$T.err; echo "scene rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_$T.json 2>> $O/bench_v64_$T.err; echo "reference rc=$?"
python tools/hbm_kernels.py $O/hbm_kernels_$T.json > $O/hbm_kernels_$T.log 2>&1; echo "hbm rc=$?"
# launch list of the default bench command (only after it exited 0 without ncu, above): the first 1000 launches
ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file $O/launches_$T.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-layers > $O/ncu_launches_$T.log 2>&1; echo "ncu launches rc=$?"
# one model call at batch 64, every launch with the full set
python tools/diag.py time 64 2 > $O/diag_time_$T.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_ -s 32 -c 16 -o $O/prof_$T python tools/diag.py time 64 2 > $O/ncu_full_$T.log 2>&1; echo "ncu full rc=$?"
python tools/ncu_summary.py $O/prof_$T.ncu-rep $O/ncu_full_per_layer_$T.txt --layers --traffic $O/traffic_$T.json \
    --title "ncu --set full --clock-control none -k regex:conv_ -s 32 -c 16 python tools/diag.py time 64 2 (one model call, B=64, 256x256; serialised, cold)" > /dev/null
ncu -i $O/prof_$T.ncu-rep --page raw --csv > $O/prof_${T}_raw.csv 2>/dev/null
HBM_K=1 HBM_REPS=1 HBM_STRIDES=64 ncu --set full --clock-control none --import-source on -k regex:"tile_|stitch_gather|patch_metrics" -c 8 \
    -o $O/prof_hbm_$T python tools/hbm_kernels.py > $O/ncu_hbm_$T.log 2>&1; echo "ncu hbm rc=$?"
python tools/ncu_summary.py $O/prof_hbm_$T.ncu-rep $O/ncu_full_hbm_kernels_$T.txt \
    --title "ncu --set full --clock-control none: patch I/O kernels, 2048^2 scene, 256/stride 64 (841 windows), one launch each x2" > /dev/null
ncu -i $O/prof_hbm_$T.ncu-rep --page raw --csv > $O/prof_hbm_${T}_raw.csv 2>/dev/null
rm -f $O/prof_hbm_$T.ncu-rep            # gpurun brings back at most 64 MiB: keep the conv report, drop this one
du -sh $O
