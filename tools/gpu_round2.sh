mkdir -p gpurun_out
Q="--no-extras --no-cpu --no-check --steps 10 --warmup 3"
for cap in 128 256; do MCG_FUSED_CAP=$cap python bench.py $Q > gpurun_out/b_cap$cap.json 2>gpurun_out/b_cap$cap.err; python -c "
import json,sys
d=json.load(open('gpurun_out/b_cap$cap.json'))
print('cap $cap', 'dev ms', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'launches/step', d['gpu_launches']/d['steps'], d['phases_ms_per_step_rank0'])
"; done
MCG_TRACE=1 python bench.py --no-extras --no-cpu --no-check --steps 2 --warmup 1 2>&1 >/dev/null | grep -E "mcg finalize|mcg query" | tail -40 > gpurun_out/trace.log; cat gpurun_out/trace.log | tail -30
CMD="python bench.py --steps 2 --warmup 1 --no-extras --no-cpu --no-check"
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(q_|rs_|runs_|table_|tile_|sketch_|cta_seq|window_count|count_sizes|set_|exp_|imp_|ovp_|remove_keys)' -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv $CMD > gpurun_out/ncu1.log 2>&1; echo ncu1=$?
