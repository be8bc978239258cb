for k in 16 24 31; do
python bench.py --workload s150 --k $k --steps 5 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/bench_s150_k${k}_r1al.json 2> gpurun_out/bench_s150_k${k}_r1al.err; echo rc=$?
python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(d['config']['k'], d['value'], d['ms_per_step'], d['config']['stage_ms'], d['config']['n_sunks'], d['config']['results_per_step']['rows'])" gpurun_out/bench_s150_k${k}_r1al.json
done
