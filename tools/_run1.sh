python -m pytest tests/test_gpu_edge_cases.py -m gpu -x -q -k host 2>&1 | tail -3
for i in 1 2 3; do python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['ms_per_step'], d['e2e']['value'])"; done
ORCA_B200_HOST_NO_AUTOTUNE=1 python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('direct only', d['ms_per_step'], d['e2e']['value'])"
ORCA_B200_HOST_NO_MAPPED=1 python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('staged only', d['ms_per_step'], d['e2e']['value'])"
