mkdir -p gpurun_out
python -m pytest tests/test_gpu_query.py -m gpu -x -q 2>&1 | tail -3
for carve in 36 44; do MCG_FUSED_CARVE=$carve python bench.py --no-extras --no-cpu --steps 10 --warmup 3 > gpurun_out/b_carve$carve.json 2>/dev/null; python -c "
import json
d=json.load(open('gpurun_out/b_carve$carve.json'))
print('carve $carve dev ms', round(d['ms_per_step'],3), 'fused', round(d['phases_ms_per_step_rank0']['fused_kernel'],3), d['parity']['verdict'])
"; done
