for lib in libfugu_old.so libfugu_gpu.so; do
  export FG_LIB=$PWD/fugu_b200/$lib
  echo "== $lib"
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-160
  FG_BENCH_COUNTS=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-160
done
