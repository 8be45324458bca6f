mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --no-extras --no-cpu --steps 10 --warmup 3 > gpurun_out/b_inl.json 2>gpurun_out/b_inl.err; python -c "
import json
d=json.load(open('gpurun_out/b_inl.json'))
print('dev ms', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'launches/step', d['gpu_launches']/d['steps'], d['phases_ms_per_step_rank0'], d['parity']['verdict'], 'build', d['build']['ms'], d['build']['ms_finalize'])
"
