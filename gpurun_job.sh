python -m pytest tests/test_gpu_rule_shims.py -m gpu -x -q 2>&1 | tail -30
(time python bench.py > gpurun_out/bench_default_r1s.json 2> gpurun_out/bench_default_r1s.err); echo rc=$?; tail -3 gpurun_out/bench_default_r1s.err
python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['config']['stage_ms'], d['roofline']['frac'], d['e2e'], d['cpu_baseline']['value'], d['clocks'])" gpurun_out/bench_default_r1s.json
