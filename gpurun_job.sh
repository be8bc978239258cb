python -m pytest tests/test_gpu_db.py -m gpu -x -q 2>&1 | tail -3
for wl in h3100 s150; do
python bench.py --workload $wl --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/bench_${wl}_ai.json 2> gpurun_out/bench_${wl}_ai.err; echo rc=$?; tail -3 gpurun_out/bench_${wl}_ai.err
python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], 'db_build_ms', d['config']['db_build_ms'], d['config']['n_sunks'], d['config']['results_per_step']['rows'])" gpurun_out/bench_${wl}_ai.json
done
