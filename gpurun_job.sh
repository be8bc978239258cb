CMD="python bench.py --coverage 6 --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_validate -s 3 -c 3 -o gpurun_out/prof_valwarp $CMD > gpurun_out/ncu1.log 2>&1
tail -2 gpurun_out/ncu1.log
