for wl in h3100 s150; do
CMD="python bench.py --workload $wl --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:k_probe2 -s 3 -c 1 -o gpurun_out/prof_probe_${wl}_r1ah -f $CMD > gpurun_out/ncu_full_${wl}_ah.log 2>&1; echo rc=$?
done
CMD="python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r1ah.csv $CMD > gpurun_out/ncu_launches_ah.log 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:k_validate -s 9 -c 3 -o gpurun_out/prof_validate_h3100_r1ah -f $CMD > gpurun_out/ncu_full_v_ah.log 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:k_cand_vote_warp -s 3 -c 1 -o gpurun_out/prof_vote_h3100_r1ah -f $CMD > gpurun_out/ncu_full_d_ah.log 2>&1; echo rc=$?
(time python bench.py > gpurun_out/bench_default_r1ah.json 2> gpurun_out/bench_default_r1ah.err); echo rc=$?
(time python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_r1ah.json 2> gpurun_out/bench_ref_r1ah.err); echo rc=$?
