timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -1
for i in 1 2; do timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'])"; done
timeout 600 python bench.py --cfg 3 --docs 10000000 --vocab 500000 --queries 2000 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('C3', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])"
