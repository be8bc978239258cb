python bench.py --steps 5 --warmup 3 --e2e-steps 3 --no-cpu-baseline > gpurun_out/bench_h3100_ag.json 2> gpurun_out/bench_h3100_ag.err; echo rc=$?; tail -3 gpurun_out/bench_h3100_ag.err
python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e'])" gpurun_out/bench_h3100_ag.json
