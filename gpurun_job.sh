CMD="python bench.py --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_probe2 -s 3 -c 1 -o gpurun_out/prof_probe_r1h $CMD > gpurun_out/ncu1.log 2>&1
tail -2 gpurun_out/ncu1.log
