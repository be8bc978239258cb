python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for wl in h3100 s150; do
python bench.py --workload $wl --steps 5 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/bench_${wl}_o.json 2> gpurun_out/bench_${wl}_o.err; echo rc=$?
python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['config']['stage_ms'], d['roofline']['frac'])" gpurun_out/bench_${wl}_o.json
done
