#!/bin/bash
# One gpurun call worth of evidence: GPU tests, smoke, bench (both arms), per-config numbers,
# ncu launch list and one --set full capture of the step kernel.  usage: tools/gpu_round.sh TAG [quick]
# Outputs land in gpurun_out/ (tools/make_profiles.py TAG copies the judged ones into profiles/).
TAG=${1:-r01}
MODE=${2:-full}
mkdir -p gpurun_out
set -o pipefail
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 || exit 1
python -c "import __graft_entry__ as g; g.smoke()" || exit 1
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_err.log || { tail -20 gpurun_out/bench_err.log; exit 1; }
cat gpurun_out/bench_${TAG}.json
[ "$MODE" = quick ] && exit 0
python bench.py --impl reference --steps 50 --warmup 5 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/err_ref.log
python tools/bench_configs.py > gpurun_out/configs_${TAG}.jsonl 2> gpurun_out/configs_err.log; cat gpurun_out/configs_${TAG}.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 330 --csv --log-file gpurun_out/launches_${TAG}.csv \
  python bench.py --steps 150 --warmup 20 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_small -s 120 -c 2 -f -o gpurun_out/prof_${TAG}_step \
  python bench.py --steps 150 --warmup 20 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
python tools/bench_policy.py > gpurun_out/policy_${TAG}.jsonl 2> gpurun_out/policy_err.log; cat gpurun_out/policy_${TAG}.jsonl
ncu --set full --clock-control none --import-source on -k regex:policy_mlp_tc -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_policy_tc \
  python tools/bench_policy.py > gpurun_out/ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:observe -s 100 -c 1 -f -o gpurun_out/prof_${TAG}_obs \
  python tools/bench_configs.py env > gpurun_out/ncu4.log 2>&1
BC_STEPS=60 BC_WARMUP=20 ncu --set full --clock-control none --import-source on -k regex:step_grid -s 70 -c 1 -f -o gpurun_out/prof_${TAG}_grid \
  python tools/bench_configs.py cfg5 > gpurun_out/ncu5.log 2>&1
python tools/obs_time.py 2>/dev/null | tail -1
