mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r02_gputests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "ref rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench_n1.json'))
print('N=1 dev ms', round(d['ms_per_step'],3), 'value M', round(d['value']/1e6,1), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'e2e M', round(d['e2e']['value']/1e6,1), d['phases_ms_per_step_rank0'], d['parity']['verdict'], 'frac', round(d['roofline']['frac'],3), 'launches/step', d['gpu_launches']/d['steps'])
print('build', d['build']['ms'], 'sketch', d['sketch']['value']/1e9, 'long', d['query_long'].get('ms_per_step'), 'cpu', d['cpu_baseline']['value'])
r=json.load(open('gpurun_out/r02_bench_reference.json'))
print('reference arm', r['value'], r['ms_per_step'])
PY
CMD="python bench.py --steps 2 --warmup 1 --no-extras --no-cpu --no-check"
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(q_|rs_|runs_|table_|tile_|sketch_|cta_seq|window_count|count_sizes|set_|exp_|imp_|ovp_|remove_keys|tgt_)' -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv $CMD > gpurun_out/ncu1.log 2>&1; echo ncu1=$?
