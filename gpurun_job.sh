CMD="python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
$CMD > gpurun_out/plain_r1k.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1k.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_probe2 -s 3 -c 1 -o gpurun_out/prof_probe_r1k $CMD > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
