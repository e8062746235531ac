python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/pcie_probe.py
python tools/bench_configs.py cfg2 env 2>&1 | tail -3
for c in 4 8 16; do ORCA_B200_HOST_CHUNKS=$c python bench.py --steps 100 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunks $c', d['ms_per_step'], d['e2e']['value'])"; done
