python -m pytest tests/test_gpu_policy.py -m gpu -x -q 2>&1 | tail -5
python tools/bench_policy.py
