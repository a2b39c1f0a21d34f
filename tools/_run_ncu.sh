python -m pytest tests -m gpu -x -q 2>&1 | tail -1
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain_c2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain_c2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"search_kernel|colscan|merge" -s 12 -c 6 -o gpurun_out/prof_r1b -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_f.log 2>&1
tail -1 gpurun_out/ncu_f.log | cut -c1-100
