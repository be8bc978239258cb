python -m pytest tests/test_gpu_match.py -m gpu -x -q 2>&1 | tail -3
for wl in h3100 s150; do
python bench.py --workload $wl --steps 5 --warmup 3 --e2e-steps 3 --no-cpu-baseline > gpurun_out/bench_${wl}_ah.json 2> gpurun_out/bench_${wl}_ah.err; echo rc=$?; tail -3 gpurun_out/bench_${wl}_ah.err
python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['config']['stage_ms'], d['e2e'])" gpurun_out/bench_${wl}_ah.json
done
