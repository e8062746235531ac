ncu --set full --clock-control none --import-source on -k regex:step_small -s 120 -c 1 -f -o gpurun_out/prof_cfg2p python bench.py --steps 150 --warmup 20 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
tail -n 1 gpurun_out/ncu2.log
