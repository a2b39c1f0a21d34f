python -m pytest tests -m gpu -x -q 2>&1 | tail -1
for nc in 1 2 3 4 6; do echo -n "chunks=$nc "; FG_PIPELINE_CHUNKS=$nc python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['value'])"; done
