for i in 1 2 3; do for lib in libfugu_base.so libfugu_pref.so; do
  export FG_LIB=$PWD/fugu_b200/$lib
  echo -n "$lib "; python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['index']['work_items'])"
done; done
