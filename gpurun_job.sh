python - <<PY
import torch
p=torch.cuda.get_device_properties(0)
print("L2", p.L2_cache_size)
import ctypes
rt=ctypes.CDLL("libcudart.so.12")
for name,aid in (("MaxPersistingL2CacheSize",108),("MaxAccessPolicyWindowSize",109)):
    v=ctypes.c_int(); rt.cudaDeviceGetAttribute(ctypes.byref(v), aid, 0); print(name, v.value)
PY
python -m pytest tests/test_gpu_match.py -m gpu -x -q 2>&1 | tail -2
for wl in h3100 s150; do
CMD="python bench.py --workload $wl --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
$CMD > gpurun_out/plain_$wl.log 2>&1; python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['config']['stage_ms']['probe'], d['roofline']['frac'])" gpurun_out/plain_$wl.log
ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:k_probe2 -s 3 -c 1 --csv --log-file gpurun_out/probe_l2_$wl.csv $CMD > gpurun_out/ncu_$wl.log 2>&1
done
