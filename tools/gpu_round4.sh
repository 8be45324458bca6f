mkdir -p gpurun_out
python -m pytest tests/test_gpu_nccl.py -m gpu -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 10 --warmup 3 --no-extras --no-cpu > gpurun_out/bench_n2b.json 2>/dev/null; python -c "
import json
d=json.load(open('gpurun_out/bench_n2b.json'))
print('N=2 dev ms', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), d['phases_ms_per_step_rank0'], d['parity']['verdict'])
"
export CUDA_VISIBLE_DEVICES=0
CMD="python bench.py --steps 2 --warmup 1 --no-extras --no-cpu --no-check"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(q_|rs_|runs_|table_|tile_|sketch_|cta_seq|window_count|count_sizes|set_|exp_|imp_|ovp_|remove_keys)' -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv $CMD > gpurun_out/ncu1.log 2>&1; echo ncu1=$?
MCG_TRACE=1 python bench.py --no-extras --no-cpu --no-check --steps 2 --warmup 1 2>&1 >/dev/null | grep -E "mcg finalize" | tail -12
