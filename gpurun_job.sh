(time python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r1ak.json 2> gpurun_out/bench_ref_r1ak.err); echo rc=$?; tail -3 gpurun_out/bench_ref_r1ak.err
python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(d['value'], d['steps'], d['ms_per_step'], d['cpu_baseline']['cores'], d['cpu_baseline']['sample'][:200])" gpurun_out/bench_ref_r1ak.json
(time python bench.py --steps 5 --warmup 3 > gpurun_out/bench_default_r1ak.json 2> gpurun_out/bench_default_r1ak.err); echo rc=$?; tail -3 gpurun_out/bench_default_r1ak.err
python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['cpu_baseline']['value'], d['e2e']['value'])" gpurun_out/bench_default_r1ak.json
