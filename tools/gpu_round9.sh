mkdir -p gpurun_out
python -m pytest tests/test_gpu_query.py tests/test_gpu_build.py -m gpu -x -q 2>&1 | tail -3
for load in 0.45 0.6 0.75; do MCG_TABLE_LOAD=$load python bench.py --no-extras --no-cpu --no-check --steps 10 --warmup 3 > gpurun_out/b_load$load.json 2>/dev/null; python -c "
import json
d=json.load(open('gpurun_out/b_load$load.json'))
print('load $load dev ms', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'fused', round(d['phases_ms_per_step_rank0']['fused_kernel'],3), 'build ms', round(d['build']['ms'],2), round(d['build']['ms_finalize'],2))
"; done
