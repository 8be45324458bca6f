mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 10 --warmup 3 --no-extras --no-cpu > gpurun_out/bench_n2c.json 2>/dev/null; python -c "
import json
d=json.load(open('gpurun_out/bench_n2c.json'))
print('N=2 dev ms', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), d['phases_ms_per_step_rank0'], d['parity']['verdict'])
"
CUDA_VISIBLE_DEVICES=0 python bench.py --no-extras --no-cpu --no-check --steps 10 --warmup 3 > gpurun_out/b_gal.json 2>/dev/null; python -c "
import json
d=json.load(open('gpurun_out/b_gal.json'))
print('N=1 dev ms', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), d['phases_ms_per_step_rank0'])
"
