export MICRO='[["text:w1",600],["w1",600],["w1 w2",400],["w1 w300",600],["w3 w40 w500",600]]'
for lib in libfugu_gpu.so libfugu_cw6144.so libfugu_cw8192.so; do
  export FG_LIB=$PWD/fugu_b200/$lib
  echo "== $lib"
  python tools/micro.py 2>&1 | cut -c1-100
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-160
done
