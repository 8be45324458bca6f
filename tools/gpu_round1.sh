mkdir -p gpurun_out; python -m pytest tests -m gpu -x -q 2>&1 | tail -4; Q="--no-extras --no-cpu --no-check --steps 10 --warmup 3"; for sub in 262144 131072 65536; do MCG_QUERY_SUB=$sub python bench.py $Q > gpurun_out/b_sub$sub.json 2>gpurun_out/b_sub$sub.err; python -c "
import json,sys
d=json.load(open('gpurun_out/b_sub$sub.json'))
print($sub, 'dev ms', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'launches/step', d['gpu_launches']/d['steps'], d['phases_ms_per_step_rank0'])
"; done; CMD="python bench.py --steps 2 --warmup 1 --no-extras --no-cpu --no-check"; ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(q_|rs_|runs_|table_|tile_|sketch_|cta_seq|window_count|count_sizes|set_|exp_|imp_|ovp_|remove_keys)' -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv $CMD > gpurun_out/ncu1.log 2>&1; echo ncu1=$?
