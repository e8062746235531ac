#!/bin/bash
# One gpurun call worth of ncu evidence for a round.  usage: tools/gpu_profiles.sh TAG
# Outputs land in gpurun_out/ (tools/make_profiles.py TAG copies the judged ones into profiles/).
TAG=${1:-r02}
PART=${2:-all}   # a | b | all  (gpurun merges at most 64 MiB back per call: four .ncu-rep files do not fit)
mkdir -p gpurun_out
B="--no-cpu-baseline --no-phase-profile --steps 40 --warmup 10"
if [ "$PART" != b ]; then
python bench.py $B > gpurun_out/bench_${TAG}_quick.json 2> gpurun_out/bench_err.log || { tail -20 gpurun_out/bench_err.log; exit 1; }
# launch list of the same command (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
  python bench.py $B > gpurun_out/ncu1.log 2>&1
# full captures of the dominant kernel of every config, inside the timed window
ncu --set full --clock-control none --import-source on -k regex:step_small -s 150 -c 1 -f -o gpurun_out/prof_${TAG}_step \
  python bench.py $B > gpurun_out/ncu2.log 2>&1; tail -1 gpurun_out/ncu2.log
ncu --set full --clock-control none --import-source on -k regex:step_small -s 150 -c 1 -f -o gpurun_out/prof_${TAG}_cfg3 \
  python bench.py --config cfg3 $B > gpurun_out/ncu3.log 2>&1; tail -1 gpurun_out/ncu3.log
fi
[ "$PART" = a ] && exit 0
ncu --set full --clock-control none --import-source on -k regex:step_small -s 80 -c 1 -f -o gpurun_out/prof_${TAG}_cfg4 \
  python bench.py --config cfg4 $B > gpurun_out/ncu4.log 2>&1; tail -1 gpurun_out/ncu4.log
ncu --set full --clock-control none --import-source on -k regex:step_grid -s 80 -c 1 -f -o gpurun_out/prof_${TAG}_grid \
  python bench.py --config cfg5 $B > gpurun_out/ncu5.log 2>&1; tail -1 gpurun_out/ncu5.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 100 --csv --log-file gpurun_out/launches_${TAG}_grid.csv \
  python bench.py --config cfg5 $B > gpurun_out/ncu6.log 2>&1
python tools/bench_configs.py > gpurun_out/configs_${TAG}.jsonl 2> gpurun_out/configs_err.log; cat gpurun_out/configs_${TAG}.jsonl | cut -c1-120
