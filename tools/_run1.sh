ncu --set full --clock-control none --import-source on -k regex:step_small -s 120 -c 1 -f -o gpurun_out/prof_cfg4g python tools/bench_configs.py cfg4 > gpurun_out/ncu4.log 2>&1
tail -n 1 gpurun_out/ncu4.log
