# quick check of a kernel change: query parity tests + a short bench (no extras, no CPU legs; parity sample on)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_query.py tests/test_gpu_cli_dropin.py -q -x 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-extras --no-cpu > gpurun_out/quick.json 2> gpurun_out/quick.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/quick.json'))
print('dev ms', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), d['phases_ms_per_step_rank0'], d['parity']['verdict'], 'frac', round(d['roofline']['frac'],3))
PY
