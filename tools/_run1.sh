python tools/bench_configs.py cfg2 cfg4 2>&1 | cut -c1-110
for f in _alt/liblp3_*.so; do echo $f; ORCA_B200_LIB=$PWD/$f python tools/bench_configs.py cfg2 cfg4 2>&1 | cut -c1-110; done
