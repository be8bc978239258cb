python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --e2e-steps 3 --no-cpu-baseline > gpurun_out/bench_h3100_p.json 2> gpurun_out/bench_h3100_p.err; echo rc=$?; tail -3 gpurun_out/bench_h3100_p.err
python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['config']['stage_ms'], d['roofline']['frac'], d['e2e'])" gpurun_out/bench_h3100_p.json
CMD="python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_r1p.csv $CMD > gpurun_out/ncu_launches_p.log 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:k_probe2 -s 3 -c 1 -o gpurun_out/prof_probe_h3100_r1p -f $CMD > gpurun_out/ncu_full_p.log 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:k_validate -s 3 -c 1 -o gpurun_out/prof_validate_h3100_r1p -f $CMD > gpurun_out/ncu_full_v.log 2>&1; echo rc=$?
