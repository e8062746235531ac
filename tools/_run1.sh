timeout 300 python -m pytest tests/test_gpu_policy.py -m gpu -x -q 2>&1 | tail -3
timeout 120 python tools/bench_policy.py
