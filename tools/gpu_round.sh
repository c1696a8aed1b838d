#!/bin/bash
# One-GPU measurement pass of a round (run under gpurun from the repo root): parity tests, the bench lines of every BASELINE
# config, the ncu launch list and the ncu --set full captures summarised under profiles/.
# usage: bash tools/gpu_round.sh TAG        (outputs: gpurun_out/*_TAG.*)
set -u
T=${1:-rX}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 > $O/pytest_$T.log; echo "pytest rc=$? $(tail -1 $O/pytest_$T.log)"
python bench.py --steps 20 --warmup 5 > $O/bench_default_$T.json 2> $O/bench_$T.err; echo "default rc=$?"
python bench.py --workload eps16 --steps 5 --warmup 3 --quick > $O/bench_eps16_$T.json 2>> $O/bench_$T.err; echo "eps16 rc=$?"
python bench.py --workload sweep --steps 3 --warmup 3 > $O/bench_sweep_$T.json 2>> $O/bench_$T.err; echo "sweep rc=$?"
python bench.py --workload latency > $O/bench_latency_$T.json 2>> $O/bench_$T.err; echo "latency rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference_$T.json 2>> $O/bench_$T.err; echo "reference rc=$?"
python tools/cudnn_baseline.py $O/cudnn_baseline_$T.json > $O/cudnn_baseline_$T.log 2>&1; echo "cudnn rc=$?"
python tools/hbm_kernels.py $O/hbm_kernels_$T.json > $O/hbm_kernels_$T.log 2>&1; echo "hbm rc=$?"
# launch list of the bench command (only after it exited 0 without ncu, above): the first 1000 launches
ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file $O/launches_$T.csv \
    python bench.py --quick --steps 1 --warmup 3 --no-layers > $O/ncu_launches_$T.log 2>&1; echo "ncu launches rc=$?"
# one model call at batch 64 and at batch 1, every launch with the full set
for B in 64 1; do
  python tools/diag.py time $B 2 >> $O/diag_time_$T.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:conv_ -s 32 -c 16 -o $O/prof_b${B}_$T python tools/diag.py time $B 2 > $O/ncu_full_b${B}_$T.log 2>&1; echo "ncu full b$B rc=$?"
  TR=""; if [ $B = 64 ]; then TR="--traffic $O/traffic_$T.json"; fi
  python tools/ncu_summary.py $O/prof_b${B}_$T.ncu-rep $O/ncu_full_per_layer_b${B}_$T.txt --layers --batch $B $TR \
      --title "ncu --set full --clock-control none -k regex:conv_ -s 32 -c 16 python tools/diag.py time $B 2 (one model call, B=$B, 256x256; serialised, cold)" > /dev/null
  rm -f $O/prof_b${B}_$T.ncu-rep          # gpurun brings back at most 64 MiB: summaries only
done
HBM_K=1 HBM_REPS=1 HBM_STRIDES=64 ncu --set full --clock-control none -k regex:"tile_|stitch_gather|patch_metrics" -c 8 \
    -o $O/prof_hbm_$T python tools/hbm_kernels.py > $O/ncu_hbm_$T.log 2>&1; echo "ncu hbm rc=$?"
python tools/ncu_summary.py $O/prof_hbm_$T.ncu-rep $O/ncu_full_hbm_kernels_$T.txt \
    --title "ncu --set full --clock-control none: patch I/O kernels, 2048^2 scene, 256/stride 64 (841 windows), one launch each x2" > /dev/null
rm -f $O/prof_hbm_$T.ncu-rep
if [ -f ab_libs/lib_timeline.so ]; then
  for C in "1 256" "4 256" "1 16"; do
    S1S2_LIB=$PWD/ab_libs/lib_timeline.so python tools/timeline.py $C > $O/timeline_b$(echo $C | tr ' ' '_')_$T.txt 2>&1
  done; echo "timeline done"
fi
du -sh $O
