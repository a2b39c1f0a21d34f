export MICRO='[["text:w1",600]]'
python tools/micro.py > gpurun_out/plain_w1c.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:colscan -s 2 -c 1 -o gpurun_out/prof_w1f -f python tools/micro.py > gpurun_out/ncu_w1f.log 2>&1
tail -2 gpurun_out/ncu_w1f.log
