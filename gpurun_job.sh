nvidia-smi topo -m 2>&1 | head -16
lscpu | grep -i -E "numa|socket|^CPU\(s\)" | head
for n in 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_h3100_n${n}_numa.json 2> gpurun_out/bench_h3100_n${n}_numa.err; echo rc=$?; tail -5 gpurun_out/bench_h3100_n${n}_numa.err | cut -c1-300
python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e'])" gpurun_out/bench_h3100_n${n}_numa.json
done
