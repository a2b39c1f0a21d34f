timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -1
export MICRO='[["w1",2400],["w1 w300",2400],["w3 w40 w500",2400],["w300 w301 w302 w303",8000],["text:w300 AND text:w301",10000]]'
timeout 120 python tools/micro.py 2>&1 | cut -c1-90
for i in 1 2 3; do timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'])"; done
