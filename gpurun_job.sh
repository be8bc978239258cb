for n in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_h3100_n${n}_r1aj.json 2> gpurun_out/bench_h3100_n${n}_r1aj.err; echo rc=$?; tail -3 gpurun_out/bench_h3100_n${n}_r1aj.err | cut -c1-300
python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['config']['stage_ms'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['packed_host_input'].get('value'), d['config']['results_per_step'])" gpurun_out/bench_h3100_n${n}_r1aj.json
done
