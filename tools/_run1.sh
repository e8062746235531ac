python tools/pcie_probe.py
nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current --format=csv
for i in 1 2; do
python bench.py --steps 100 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('mapped', d['ms_per_step'], d['e2e']['value'])"
ORCA_B200_HOST_NO_MAPPED=1 python bench.py --steps 100 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('staged', d['ms_per_step'], d['e2e']['value'])"
done
